"""bench.py contract (CPU part): the reference arm — the reference's own compiled units (oracle/_ref) stepping a bounded
sample on the host cores, or the oracle port of the same loops where they are not available — prints ONE JSON line
with the keys the driver reads, for N = 1 and, under torchrun with two ranks, from rank 0 only."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]


def _check(line, n_gpus):
    d = json.loads(line)
    for k in KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "atom-timesteps/s" and d["n_gpus"] == n_gpus
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "buck/coul/long" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    # the reference's own compiled units where oracle/_ref exists (this container, and prebuilt on the GPU box)
    assert cb["kind"] == ("reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref.so")) else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if cb["kind"] == "reference":
        ph = cb["phase_s"]
        assert ph["ref_pair"] > 0 and ph["ref_kspace"] > 0 and ph["ref_nve"] >= 0
        # the reported time is the reference's members + the upstream glue; marshalling is left out and stated
        timed = sum(v for k, v in ph.items() if k != "harness")
        assert abs(d["ms_per_step"] * d["steps"] / 1e3 - timed) < 5e-3 * max(d["steps"], 1) + 1e-2 * timed
        assert d["wall_seconds_with_marshalling"] >= timed
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def _json_lines(out):
    return [l for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_prints_once():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1, lines          # rank 0 alone runs and prints; the other rank exits 0 without work
    _check(lines[0], 2)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_own_arm_json_line_on_a_small_workload():
    """bench.py on one GPU at a reduced size (--rep 4: 76 800 atoms): one JSON line carrying every key of the contract,
    with the roofline, cpu_baseline, e2e and clocks objects filled in"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--rep", "4", "--steps", "6", "--warmup", "3",
                        "--cpu-rep", "2", "--cpu-steps", "2"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in [k for k in KEYS if k != "impl"] + ["clocks", "gpu_launches", "roofline", "roofline_kernels",
                                                 "step_roofline_frac", "parity"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    rf = d["roofline"]
    assert rf["achieved"] > 0 and rf["peak"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 76800 * 24 and e["d2h_bytes_per_step"] == 76800 * 48
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["cores"] >= 1
    if cb["kind"] == "reference":
        # the reference's own compiled units (oracle/_ref travels to the GPU box prebuilt) run the step; the oracle port
        # of the same loops is timed beside them
        assert cb["phase_s"]["ref_pair"] > 0 and cb["phase_s"]["ref_kspace"] > 0 and cb["port"]["value"] > 0
    assert d["clocks"]["sm_max_mhz"] > 0
    # one roofline entry per kernel with >= 1 % of the step, measured in this run
    names = [r["kernel"] for r in d["roofline_kernels"]]
    assert "k_pair" in names and any(n.startswith("k_fft") for n in names)
    for r in d["roofline_kernels"]:
        assert r["avg_launch_ms"] > 0 and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3
    assert 0.0 < d["step_roofline_frac"] <= 1.0
    # parity of the same styles on the bounded sample (here data.aC x 2^3), GPU vs oracle, inside the bench line
    p = d["parity"]
    assert p["ok"] is True and p["pair_set_equal"] is True
    assert p["max_rel_force_err"] <= 1e-9 and p["epair_rel"] <= 1e-10 and p["ekspace_rel"] <= 1e-9


def test_kernel_roofline_models_of_the_half_spectrum_passes():
    """bench.kernel_rooflines is pure arithmetic: with the half-spectrum transforms on, the spectral passes are charged
    Fs = F (nx/2 + 1) / nx points and three fields come back; the ideal step time is the sum of the rows"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N, entries, grid = 4050000, 2087561650, (250, 250, 270)
    F = grid[0] * grid[1] * grid[2]
    timers = {"pair": (200.0, 20), "k_fft_x_fwd": (2.0, 20), "k_fft_y_fwd": (2.4, 20), "k_fft_z_poisson": (9.2, 20),
              "k_fft_y_inv": (6.4, 20), "k_fft_x_inv": (6.0, 60), "fieldforce": (18.0, 20)}
    cfg = {"flops_key": "buck_coul_long"}
    rows, ideal, nbar = bench.kernel_rooflines(cfg, timers, 20, N, N + 800000, entries, F, 0, 34.7, 6551.4, "double",
                                               fp32_peak=65.0, half_nx=grid[0])
    by = {r["kernel"]: r for r in rows}
    Fs = F * (grid[0] // 2 + 1) / grid[0]
    assert by["k_fft_x_r2c (x fwd, real in)"]["work_per_launch"] == 8.0 * F + 16.0 * Fs
    assert by["k_fft_z_poisson"]["work_per_launch"] == (16.0 + 8.0 + 48.0) * Fs
    assert by["k_fft_x_c2r (x inv, real out)"]["launches"] == 60
    assert abs(ideal - sum(r["ideal_ms_per_step"] for r in rows)) < 1e-12 and nbar == entries / N
    full, _, _ = bench.kernel_rooflines(cfg, timers, 20, N, N + 800000, entries, F, 0, 34.7, 6551.4, "double", fp32_peak=65.0)
    assert {r["kernel"] for r in full} >= {"k_fft_pass x fwd (real in)", "k_fft_pass x inv (real out)"}
    assert bench.pppm_half_spectrum(1) is (os.environ.get("B200MD_R2C", "1")[:1] != "0")
