#!/usr/bin/env python
"""bench.py — headline benchmark: atom-timesteps/s for buck/coul/long + PPPM (BASELINE.json).

A "step" is one MD timestep of the hot path: fix nve/intel initial_integrate -> Neighbor::decide (ghost refresh,
or re-bin + ghost rebuild + full neighbour-list build when an atom moved skin/2) -> PairBuckCoulLongIntel::compute
-> PPPMIntel::compute -> final_integrate, atoms resident in HBM.  Workload (N=1): data.aC replicated 15^3 =
4.05 M atoms, `buck/coul/long 12.0`, `kspace_style pppm 1e-4` order 5 (grid 250x250x270), skin 0.3, check yes,
double precision (`package intel mode double`) — SURVEY.md §8d S3 / §6.2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

`--impl reference` times the CPU restatement of the reference (oracle/, the reference itself cannot be compiled:
DESIGN.md) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "atom-timesteps/s buck/coul/long+PPPM"
UNIT = "atom-timesteps/s"
PAIR_FLOPS = 125.0          # SURVEY §8d official work per pair evaluation, buck/coul/long analytic
CUT, SKIN, ACC, ORDER = 12.0, 0.3, 1.0e-4, 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rep", type=int, default=15, help="data.aC replication per dimension per GPU (15 -> 4.05M atoms)")
    ap.add_argument("--prec", default="double", choices=["double", "mixed"])
    ap.add_argument("--table", type=int, default=0, help="1: Coulomb lookup tables (INTEL_ALLOW_TABLE path)")
    ap.add_argument("--cpu-rep", type=int, default=8, help="replication of the CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md clocks line)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu=0):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [v.strip() for v in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload(W, rep, rank=0, nranks=1):
    """data.aC x rep^3 per GPU; ranks stack along z (weak scaling: per-GPU work fixed).  Every rank generates only
    its own block (its z slab of the rep x rep x rep*nranks system) and the global box."""
    # as the shipped script: the crystal of data.aC as read, `velocity all create 300.0` (no displacement)
    s = W.aC_system((rep, rep, rep), jitter=0.0, seed=1281937 + rank)
    lz = s["boxhi"][2] - s["boxlo"][2]
    s["x"][:, 2] += rank * lz
    s["boxhi"] = s["boxhi"].copy()
    s["boxhi"][2] = s["boxlo"][2] + nranks * lz
    return s


def pair_setup_args(pkg, W, s, g_ewald, table):
    co = W.coeffs_aC(CUT, CUT)
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    ct = pkg.init_coul_tables(CUT, g_ewald, W.UNITS["metal"]["qqrd2e"]) if table else None
    return co, cf, ct


def run_b200(args):
    import torch
    import torch.distributed as dist
    pkg = graft.load_package()
    W = importlib.import_module("lammps_buck_intel_b200.workloads")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        if rank == 0:
            print("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world),
                  file=sys.stderr)
        sys.exit(2)
    u = W.UNITS["metal"]
    prec = pkg.PREC_DOUBLE if args.prec == "double" else pkg.PREC_MIXED
    # each rank builds the whole (deterministic) system; the library keeps its slab
    s = workload(W, args.rep, rank, world)
    nlocal0 = len(s["x"])
    natoms = nlocal0 * world            # identical blocks: global count and charge sums follow from one block
    prd = s["boxhi"] - s["boxlo"]
    grid, g_ewald = pkg.pppm_init(ACC, u["qqrd2e"], float(np.sum(s["q"] ** 2)) * world, natoms, CUT, prd, order=ORDER)
    co, cf, ct = pair_setup_args(pkg, W, s, g_ewald, args.table)
    ctx = pkg.Context(local_rank, prec)
    ctx.set_units(u["qqrd2e"], u["ftm2v"])
    ctx.set_box(s["boxlo"], s["boxhi"])
    if world > 1:
        ctx.comm_init_torch(dist, rank, world)
    ctx.atoms_upload(s["x"], s["type"], s["mass"], v=s["v"], q=s["q"])
    ctx.neigh_setup(SKIN, every=1, delay=0, check=1)
    ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=g_ewald, coul_tables=ct)
    ctx.pppm_setup(*grid, ORDER, g_ewald)
    ctx.nve_setup(u["dt"])
    ctx.setup_forces(0, 0)
    ctx.neigh_build()          # a second build settles every capacity-grown buffer before anything is timed
    st0 = ctx.neigh_stats()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed region: EXACTLY K steps, device-timed, max over ranks ---------------
    ctx.run(max(args.warmup, 3))
    ctx.timers_enable(True)
    ctx.timers_reset()
    l0 = ctx.launch_count()
    nb0 = ctx.neigh_stats()["nbuilds"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    ms_dev = ctx.run_timed(args.steps)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - l0
    timers = ctx.timers()
    ctx.timers_enable(False)
    st1 = ctx.neigh_stats()
    if world > 1:
        t = torch.tensor([ms_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
    ms_per_step = ms_dev / args.steps
    value = natoms * args.steps / (ms_dev * 1e-3)

    # ---- roofline of the dominant kernel (pair, FP64-bound: SURVEY §8d) ---------------------------------
    pair_ms, pair_calls = timers["pair"]
    entries = st1["total"]
    if world > 1:
        t = torch.tensor([float(entries)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        entries = int(t.item())
    fp64_peak = ctx.microbench(0) if prec == pkg.PREC_DOUBLE else ctx.microbench(1)
    hbm_peak, hbm_src = peaks()
    pair_avg_ms = pair_ms / max(pair_calls, 1)
    local_entries = st1["total"]
    local_atoms = nlocal0
    achieved_tf = PAIR_FLOPS * local_entries / (pair_avg_ms * 1e-3) / 1e12 if pair_avg_ms > 0 else 0.0
    bytes_per_atom = (4.0 * local_entries / max(local_atoms, 1) + 8 + 32 + 32) if prec == pkg.PREC_DOUBLE else \
        (4.0 * local_entries / max(local_atoms, 1) + 8 + 16 + 32)
    pair_gbs = bytes_per_atom * local_atoms / (pair_avg_ms * 1e-3) / 1e9 if pair_avg_ms > 0 else 0.0
    kname = "k_pair<buck/coul/long,%s>" % args.prec
    traffic = None   # DRAM bytes per launch of this kernel from the committed ncu capture (profiles/traffic.json)
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(kname, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {"kernel": kname, "bound": "fp64" if prec == pkg.PREC_DOUBLE else "fp32",
                "achieved": round(achieved_tf, 3), "peak": round(fp64_peak, 3), "unit": "TFLOP/s",
                "frac": round(achieved_tf / fp64_peak, 4) if fp64_peak else None, "traffic": traffic,
                "peak_source": "FMA microbenchmark in this run (b200md_microbench); MEASURED_PEAKS.json holds no FP64 figure",
                "work_model": "125 flop per neighbour-list entry (SURVEY 8d) x %d entries per launch" % local_entries,
                "avg_launch_ms": round(pair_avg_ms, 4), "share_of_step": round(pair_avg_ms / ms_per_step, 4)}
    roofline_hbm = {"kernel": roofline["kernel"], "bound": "hbm", "achieved": round(pair_gbs, 1), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(pair_gbs / hbm_peak, 4), "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json (%s)" % hbm_src,
                    "work_model": "(4*nbar + 8 + 32 + 32) B per atom-step, nbar = %.1f" % (local_entries / max(local_atoms, 1))}

    # ---- e2e: the same step through host buffers (pinned), H2D x and D2H x,f every step ------------------
    e2e = None
    if not args.no_e2e and world == 1:
        xin = torch.empty((natoms, 3), dtype=torch.float64).pin_memory()
        xout = torch.empty((natoms, 3), dtype=torch.float64).pin_memory()
        fout = torch.empty((natoms, 3), dtype=torch.float64).pin_memory()
        xin_n, xout_n, fout_n = xin.numpy(), xout.numpy(), fout.numpy()
        xin_n[:] = ctx.atoms_download(("x",))["x"]
        ne = max(3, min(args.steps, 10))
        for _ in range(2):
            ctx.step_host(xin_n, xout_n, fout_n)
            xin_n[:] = xout_n
        torch.cuda.synchronize()
        te = time.perf_counter()
        for _ in range(ne):
            ctx.step_host(xin_n, xout_n, fout_n)
            xin_n, xout_n = xout_n, xin_n      # next step uploads what was just downloaded
        torch.cuda.synchronize()
        te = time.perf_counter() - te
        e2e = {"value": natoms * ne / te, "unit": UNIT, "h2d_bytes_per_step": natoms * 24,
               "d2h_bytes_per_step": natoms * 48, "steps": ne,
               "note": "b200md_step_host: pinned host x in, x and f out every step; includes the host-side un-permute"}
    elif world > 1:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "host-buffer stepping is single-GPU only"}

    # ---- CPU baseline: the oracle's whole-step loop on the host cores, bounded sample ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args, W)

    if rank == 0:
        phase = {k: round(v[0] / args.steps, 4) for k, v in timers.items() if v[1] > 0}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if prec == pkg.PREC_DOUBLE else "f32 compute / f64 accumulate", "data": "synthetic",
            "config": {"workload": "data.aC x %dx%dx%d (%d atoms) buck/coul/long %.1f + pppm %g order %d grid %dx%dx%d "
                                   "g_ewald %.4f, skin %.1f check yes, nve dt 1 fs" %
                                   (args.rep, args.rep, args.rep * world, natoms, CUT, ACC, ORDER, *grid, g_ewald, SKIN),
                       "precision": args.prec, "coulomb": "table" if args.table else "analytic erfc",
                       "neighbor_entries": int(entries), "nghost": st1["nghost"],
                       "rebuilds_in_timed_region": int(st1["nbuilds"] - nb0),
                       "l2": "inputs larger than L2 (neighbour list %.1f GB per GPU)" % (4.0 * local_entries / 1e9),
                       "parallelism": "z-slab x%d" % world if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu, "phase_ms_per_step": phase, "wall_s_timed_region": wall,
        }
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_setup(W, rep):
    orc = graft.load_oracle()
    pkg = graft.load_package()
    u = W.UNITS["metal"]
    s = W.aC_system(rep, jitter=0.0)
    natoms = len(s["x"])
    grid, g_ewald = pkg.pppm_init(ACC, u["qqrd2e"], s["q"], natoms, CUT, s["boxhi"] - s["boxlo"], order=ORDER)
    co = W.coeffs_aC(CUT, CUT)
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=g_ewald)
    pp = orc.PPPM(*grid, ORDER, g_ewald, s["boxlo"], s["boxhi"], u["qqrd2e"])
    md = orc.MD(s, P, prec=orc.DOUBLE, skin=SKIN, every=1, delay=0, check=1, dt=u["dt"], ftm2v=u["ftm2v"], pppm=pp)
    return orc, s, md, natoms, grid


def cpu_baseline(args, W):
    """oracle/ whole-step loop (half list, newton on, thread-private force arrays and grids, OpenMP over all host
    cores) on a bounded sample: data.aC x cpu_rep^3, same styles/accuracy; the metric is intensive in N."""
    cores = os.cpu_count() or 1
    orc, s, md, natoms, grid = cpu_setup(W, args.cpu_rep)
    md.run(1, cores)                      # builds the list, first forces
    t0 = time.perf_counter()
    tm = md.run(args.cpu_steps, cores)
    dt = time.perf_counter() - t0
    return {"value": natoms * args.cpu_steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "data.aC x %d^3 = %d atoms, grid %dx%dx%d, %d steps, %.1f s; restatement of the reference loops "
                      "(g++ -O3 -march=native -fopenmp), not the ICC USER-INTEL build" % (args.cpu_rep, natoms, *grid, args.cpu_steps, dt),
            "phase_s": {k: round(v, 3) for k, v in tm.items() if k != "nbuilds"}}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference cannot be built
    here), all host threads, same metric/config, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = importlib.import_module("lammps_buck_intel_b200.workloads") if graft.load_package() else None
    cores = os.cpu_count() or 1
    orc, s, md, natoms, grid = cpu_setup(W, args.cpu_rep)
    md.run(max(args.warmup, 1), cores)
    t0 = time.perf_counter()
    md.run(args.steps, cores)
    dt = time.perf_counter() - t0
    value = natoms * args.steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": max(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "data.aC x %d^3 (%d atoms per step sample) buck/coul/long %.1f + pppm %g order %d grid "
                                  "%dx%dx%d, skin %.1f check yes, nve; bounded sample of the 4.05 M-atom workload "
                                  "(metric is intensive in N)" % (args.cpu_rep, natoms, CUT, ACC, ORDER, *grid, SKIN)},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "oracle/ restatement (half list, newton on, OpenMP %d threads); /root/reference needs "
                                      "LAMMPS core + MPI + ICC and cannot be compiled" % cores},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
