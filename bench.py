#!/usr/bin/env python
"""bench.py — headline benchmark: atom-timesteps/s for buck/coul/long + PPPM (BASELINE.json).

A "step" is one MD timestep of the hot path: fix nve/intel initial_integrate -> Neighbor::decide (ghost refresh,
or re-bin + ghost rebuild + full neighbour-list build when an atom moved skin/2) -> Pair*Intel::compute
-> PPPMIntel::compute -> final_integrate, atoms resident in HBM.  Default workload (N=1): data.aC replicated 15^3 =
4.05 M atoms, `buck/coul/long 12.0`, `kspace_style pppm 1e-4` order 5 (grid 250x250x270), skin 0.3, check yes,
double precision (`package intel mode double`) — SURVEY.md §8d S3 / §6.2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--config selects one of BASELINE.json's other configurations (parity-test cases, not the headline):
    buck_coul_long  (default)   examples/in.buck_coul_long semantics, data.aC x rep^3
    buck                        examples/in.buck: fcc LJ-unit melt, pair buck 2.5, every 20 check no
    buck_big                    examples/in.buck_big: pair buck 5.0, delay 5 every 1
    buck_coul_cut               examples/in.buck_coul_cut: data.aC x rep^3, buck/coul/cut 10.0, no kspace
    spce_pppm                   examples/in.spce electrostatics only: data.spce x rep^3, PPPMIntel::compute alone
    spce                        examples/in.spce pair + k-space path: lj/cut/coul/long 6.8 8.8 with the molecules' special
                                bonds in the device-built list + pppm, fix nve (SHAKE / bonds / NVT are off that path)
    buck_big_disp               BASELINE config 5: fcc melt, buck/long/coul/long long off 5.0 + pppm/disp (geometric)

The JSON line carries, beside the contract's keys: `roofline` (the dominant kernel), `roofline_kernels` (one entry per
kernel with >= 1 % of the step: live CUDA-event launch time, algorithmic bytes|flops of SURVEY 8d, achieved, frac),
`step_roofline_frac` (sum of the kernels' ideal times / measured step), and `parity` (the same styles on the bounded
sample data.aC x 8^3, GPU path against the CPU oracle: forces, energies, pair set; at N > 1 the sample is decomposed over
the N ranks, so the line proves the multi-GPU path too).

`--impl reference` (and the `cpu_baseline` leg of the own arm) steps a bounded sample of the same workload on the host
cores with every function the reference ships running from the reference's OWN translation units, compiled unchanged
(oracle/_ref: PairBuckCoulLongIntel::compute, PPPMIntel::compute, FixNVEIntel), and the upstream pieces around them
(neighbour list, ghosts, communication) from the oracle — tests/refmd.py; where oracle/_ref is missing, the oracle's
restatement of the same loops (bit-identical to it) is timed instead.
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

UNIT = "atom-timesteps/s"
ORDER = 5
# SURVEY §8d official work per pair evaluation (one neighbour-list entry)
PAIR_FLOPS = {"buck": 58.0, "buck_coul_cut": 69.0, "buck_coul_long": 125.0, "buck_long_coul_long": 98.0,
              # not in SURVEY 8d's list; counted the same way from pair_lj_long_coul_long_intel.cpp:540-690 (ORDER1, cut LJ):
              # the buck/coul/long count without exp(-r/rho) and its two multiplies
              "lj_long_coul_long": 103.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="buck_coul_long",
                    choices=["buck_coul_long", "buck", "buck_big", "buck_coul_cut", "spce_pppm", "spce", "buck_big_disp"])
    ap.add_argument("--rep", type=int, default=0, help="replication per dimension per GPU (0: the config's default; "
                                                       "buck_coul_long 15 -> 4.05 M atoms)")
    ap.add_argument("--rep3", type=int, nargs=3, default=None, help="explicit replication / cells per dimension (one GPU), "
                                                                     "e.g. 30 40 40 = the in.buck_big box")
    ap.add_argument("--geometry", default="cube", choices=["slab", "cube"],
                    help="N > 1: 'cube' (default) replicates the global system isotropically, round(rep * N^(1/3)) per "
                         "dimension (SURVEY S3: data.aC x 30^3 = 32.4 M atoms on 8 GPUs; 19^3 / 24^3 on 2 / 4: per-GPU "
                         "work within 2.5 %% of the single-GPU run); 'slab' stacks the per-GPU blocks along z "
                         "(rep x rep x rep*N), which keeps the halo area per GPU constant and flatters a slab decomposition")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --rep is the GLOBAL replication, split over the ranks")
    ap.add_argument("--acc", type=float, default=1.0e-4, help="kspace accuracy (the shipped in.buck_coul_long says 1e-6)")
    ap.add_argument("--prec", default="double", choices=["double", "mixed"])
    ap.add_argument("--table", type=int, default=0, help="1: Coulomb lookup tables (INTEL_ALLOW_TABLE path)")
    ap.add_argument("--cpu-rep", type=int, default=8, help="replication of the CPU-baseline / parity sample")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


# ---------------------------------------------------------------------------------------------------------------------
# configurations

def split_reps(rep, world, geometry, scaling):
    """(rx, ry, rz) of the GLOBAL system for `world` ranks"""
    if scaling == "strong":
        return rep, rep, rep
    if geometry == "cube" and world > 1:
        r = round(rep * world ** (1.0 / 3.0))
        return r, r, r
    return rep, rep, rep * world


def my_slab(s, rank, world):
    """keep this rank's z slab of the global system `s` (atoms and velocities); returns (system, first global id)"""
    if world == 1:
        return s, 0
    lo, hi = s["boxlo"][2], s["boxhi"][2]
    slab = (hi - lo) / world
    owner = np.floor((s["x"][:, 2] - lo) / slab).astype(np.int64).clip(0, world - 1)
    own = owner == rank
    first = int(np.sum(owner < rank))
    out = dict(s)
    for k in ("x", "v", "type", "q"):
        if s.get(k) is not None:
            out[k] = s[k][own]
    out["gid"] = np.nonzero(own)[0]
    return out, first


def rank_system(make, reps, rank, world):
    """this rank's part of the global system make(reps): it generates only the z blocks that overlap its slab
    [rank, rank + 1) * Lz / world and keeps the atoms inside the slab (the blocks need not align with the slabs: the
    cubic geometries split 19, 24 or 30 cells over 2, 4 or 8 ranks); the box is the global one"""
    rx, ry, rz = reps
    if world == 1:
        return make((rx, ry, rz))
    b0 = int(np.floor(rz * rank / world))
    b1 = int(np.ceil(rz * (rank + 1) / world))
    s = make((rx, ry, b1 - b0))
    lz_cell = (s["boxhi"][2] - s["boxlo"][2]) / (b1 - b0)
    s["x"][:, 2] += b0 * lz_cell
    s["boxhi"] = s["boxhi"].copy()
    s["boxhi"][2] = s["boxlo"][2] + rz * lz_cell
    slab = (s["boxhi"][2] - s["boxlo"][2]) / world
    own = np.floor((s["x"][:, 2] - s["boxlo"][2]) / slab).astype(np.int64).clip(0, world - 1) == rank
    for k in ("x", "v", "type", "q"):
        if s.get(k) is not None:
            s[k] = np.ascontiguousarray(s[k][own])
    return s


class Config:
    """one BASELINE configuration: system, styles, the work model of its pair kernel"""

    def __init__(self, args, pkg, W, world):
        self.args, self.pkg, self.W, self.world = args, pkg, W, world
        c = args.config
        self.name = c
        self.kspace = c in ("buck_coul_long", "spce_pppm", "spce", "buck_big_disp")
        self.pair = c != "spce_pppm"
        dflt = {"buck_coul_long": 15, "buck_coul_cut": 15, "buck": 100, "buck_big": 100, "spce_pppm": 4, "spce": 8,
                "buck_big_disp": 100}[c]
        self.rep = args.rep or dflt
        self.metric = {"buck_coul_long": "atom-timesteps/s buck/coul/long+PPPM", "buck": "atom-timesteps/s buck",
                       "buck_big": "atom-timesteps/s buck (in.buck_big)",
                       "buck_coul_cut": "atom-timesteps/s buck/coul/cut",
                       "spce_pppm": "atom-timesteps/s PPPM only (data.spce)",
                       "spce": "atom-timesteps/s lj/cut/coul/long+PPPM (data.spce)",
                       "buck_big_disp": "atom-timesteps/s buck/long/coul/long+pppm/disp"}[c]

    def system(self, rep3):
        W, c = self.W, self.name
        if c in ("buck_coul_long", "buck_coul_cut"):
            return W.aC_system(rep3, jitter=0.0)   # the crystal as read, `velocity all create 300.0`
        if c in ("buck", "buck_big", "buck_big_disp"):
            return W.fcc_system(*rep3, jitter=0.0)
        return W.spce_system(rep3, temperature=30.0 if c == "spce" else 300.0)   # no bonded forces: keep the molecules together

    def global_reps(self):
        if self.args.rep3 and self.world == 1:
            return tuple(self.args.rep3)
        return split_reps(self.rep, self.world, self.args.geometry, self.args.scaling)

    def setup(self, ctx, s, natoms_global, qsq_global):
        """styles on the context; returns a description dict"""
        pkg, W, a, c = self.pkg, self.W, self.args, self.name
        u = W.UNITS[s["units"]]
        prd = s["boxhi"] - s["boxlo"]
        d = {}
        if c == "buck_coul_long":
            cut, skin = 12.0, 0.3
            grid, g = pkg.pppm_init(a.acc, u["qqrd2e"], qsq_global, natoms_global, cut, prd, order=ORDER)
            co = W.coeffs_aC(cut, cut)
            cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
            ct = pkg.init_coul_tables(cut, g, u["qqrd2e"]) if a.table else None
            ctx.neigh_setup(skin, every=1, delay=0, check=1)
            ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=g, coul_tables=ct)
            ctx.pppm_setup(*grid, ORDER, g)
            d = dict(style="buck/coul/long %.1f + pppm %g order %d" % (cut, a.acc, ORDER), grid=list(grid), g_ewald=g,
                     neigh="skin 0.3 delay 0 every 1 check yes", flops_key="buck_coul_long")
        elif c == "buck_coul_cut":
            cut, skin = 10.0, 0.3
            co = W.coeffs_aC(cut, cut)
            cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_CUT, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
            ctx.neigh_setup(skin, every=1, delay=0, check=1)
            ctx.pair_setup(pkg.PAIR_BUCK_COUL_CUT, 2, cf)
            d = dict(style="buck/coul/cut %.1f" % cut, neigh="skin 0.3 delay 0 every 1 check yes",
                     flops_key="buck_coul_cut")
        elif c in ("buck", "buck_big"):
            cut = 2.5 if c == "buck" else 5.0
            co = W.coeffs_in_buck(cut)
            cf = pkg.pair_coeffs(pkg.PAIR_BUCK, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
            if c == "buck":
                ctx.neigh_setup(0.3, every=20, delay=0, check=0)
                neigh = "skin 0.3 delay 0 every 20 check no"
            else:
                ctx.neigh_setup(0.3, every=1, delay=5, check=1)
                neigh = "skin 0.3 delay 5 every 1 check yes"
            ctx.pair_setup(pkg.PAIR_BUCK, 1, cf)
            d = dict(style="buck %.1f" % cut, neigh=neigh, flops_key="buck")
        elif c == "buck_big_disp":
            # BASELINE config 5: buck/long/coul/long long off 5.0 + pppm/disp, geometric mixing (B = sqrt|C|).
            # Dispersion mesh: h = 1 / g_ewald_6-scale of the in.buck_big cell (explicit, like `kspace_modify mesh/disp`)
            cut = 5.0
            co = W.coeffs_in_buck(cut)
            cf = pkg.pair_coeffs(pkg.PAIR_BUCK_LONG_COUL_LONG, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
            g6 = 0.68
            nmesh = [self._mesh(p, 0.84) for p in prd]
            ctx.neigh_setup(0.3, every=1, delay=5, check=1)
            ctx.pair_setup(pkg.PAIR_BUCK_LONG_COUL_LONG, 1, cf, g_ewald_6=g6, ewald_order=1 << 6)
            B = np.array([0.0, np.sqrt(abs(co["C"][1, 1]))])
            ctx.pppm_setup(*nmesh, ORDER, g6, dispersion=1, B=B)
            d = dict(style="buck/long/coul/long long off %.1f + pppm/disp (geometric) order %d" % (cut, ORDER),
                     grid=nmesh, g_ewald_6=g6, neigh="skin 0.3 delay 5 every 1 check yes",
                     flops_key="buck_long_coul_long")
        elif c == "spce":
            cl, cc, skin = 6.8, 8.8, 2.0
            grid, g = pkg.pppm_init(a.acc, u["qqrd2e"], qsq_global, natoms_global, cc, prd, order=ORDER)
            co = W.coeffs_spce(cl, cc)
            cf = pkg.pair_coeffs(pkg.PAIR_LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
            ct = pkg.init_coul_tables(cc, g, u["qqrd2e"]) if a.table else None
            sp = (1, 0.0, 0.0, 0.5)
            # SHAKE and the bonded terms are off this path, so nothing holds the sites of a molecule together: the positions
            # advance with a 0.01 fs step (same work per step) and the list is rebuilt every 10 steps, the cadence in.spce
            # reaches with its 2 A skin and `delay 10`
            ctx.neigh_setup(skin, every=10, delay=0, check=0)
            self.dt_override = 0.01
            ctx.pair_setup(pkg.PAIR_LJ_LONG_COUL_LONG, 2, cf, special_lj=sp, special_coul=sp, g_ewald=g,
                           ewald_order=1 << 1, coul_tables=ct)
            n = len(s["x"])
            nsp = np.zeros((n, 3), np.int32)
            spl = np.zeros((n, 2), np.int32)
            o = np.arange(0, n, 3)            # O, H, H per molecule (examples/data.spce)
            nsp[o] = (2, 2, 2)
            spl[o, 0], spl[o, 1] = o + 1, o + 2
            for h, other in ((o + 1, o + 2), (o + 2, o + 1)):
                nsp[h] = (1, 2, 2)
                spl[h, 0], spl[h, 1] = o, other
            ctx.atoms_set_special(nsp, spl)
            ctx.pppm_setup(*grid, ORDER, g)
            d = dict(style="lj/cut/coul/long %.1f %.1f + special_bonds lj/coul 0 0 0.5 + pppm %g order %d" % (cl, cc, a.acc, ORDER),
                     grid=list(grid), g_ewald=g, neigh="skin 2.0, rebuilt every 10 steps, dt 0.01 fs (no SHAKE / bonds on the path)",
                     flops_key="lj_long_coul_long")
        elif c == "spce_pppm":
            cut = 8.8
            grid, g = pkg.pppm_init(a.acc, u["qqrd2e"], qsq_global, natoms_global, cut, prd, order=ORDER)
            ctx.neigh_setup(2.0, every=1, delay=10, check=1)
            ctx.pppm_setup(*grid, ORDER, g)
            d = dict(style="pppm %g order %d (lj/cut/coul/long 6.8 8.8 not evaluated: electrostatics only)" % (a.acc, ORDER),
                     grid=list(grid), g_ewald=g, neigh="none (k-space only)", flops_key=None)
        ctx.nve_setup(getattr(self, "dt_override", None) or u["dt"])
        return d

    @staticmethod
    def _mesh(prd, h):
        n = max(int(prd / h) + 1, 2 * ORDER)
        while True:
            m = n
            for p in (2, 3, 5):
                while m % p == 0:
                    m //= p
            if m == 1:
                return n
            n += 1


# ---------------------------------------------------------------------------------------------------------------------
# per-kernel roofline (SURVEY §8d work models; launch times from CUDA events recorded on the launching stream)

def pppm_half_spectrum(world):
    """mirrors b200md_pppm_setup: real-to-complex transforms unless B200MD_R2C=0 (orthogonal boxes)"""
    return os.environ.get("B200MD_R2C", "1")[:1] != "0"


def kernel_rooflines(cfg, timers, steps, N, nall, entries, F, G_tiles, fp_peak, hbm_peak, prec, ncomp_packs=2,
                     fp32_peak=None, half_nx=0):
    """timers: name -> (ms, calls) accumulated over `steps` timed steps on this rank.  Returns (list, ideal ms/step).
    half_nx > 0: the half-spectrum (real-to-complex) PPPM transforms are on; the spectral arrays hold
    Fs = F (nx/2 + 1) / nx points and the three gradient fields are transformed one by one."""
    flt = 8 if prec == "double" else 4
    fl = PAIR_FLOPS.get(cfg["flops_key"]) if cfg.get("flops_key") else None
    rows = []

    def add(kernel, tname, bound, work, unit_note):
        ms, calls = timers.get(tname, (0.0, 0))
        if calls == 0 or ms <= 0.0:
            return
        avg = ms / calls
        if bound == "hbm":
            ach = work / (avg * 1e-3) / 1e9
            peak, unit = hbm_peak, "GB/s"
        else:
            ach = work / (avg * 1e-3) / 1e12
            peak, unit = fp_peak, "TFLOP/s"
        rows.append({"kernel": kernel, "bound": bound, "avg_launch_ms": round(avg, 4), "launches": int(calls),
                     "ms_per_step": round(ms / steps, 4), "work_per_launch": work, "work_model": unit_note,
                     "achieved": round(ach, 2), "peak": round(peak, 2), "unit": unit, "frac": round(ach / peak, 4),
                     "ideal_ms_per_step": round(work / (peak * (1e9 if bound == "hbm" else 1e12)) * 1e3 * calls / steps, 4)})

    nbar = entries / max(N, 1)
    if fl:
        add("k_pair", "pair", "fp64" if prec == "double" else "fp32", fl * entries,
            "%g flop per list entry (SURVEY 8d) x %d entries" % (fl, entries))
    # the mask kernel is arithmetic: ~3.8 candidates per neighbour (5^3-bin stencil over the cut-off sphere), 9 flop per
    # distance test (3 sub, 3 mul/fma, compare), FP32 on bin-relative coordinates with an exact FP64 redo in the guard band
    ms_m, calls_m = timers.get("k_nb_mask", (0.0, 0))
    if calls_m and fp32_peak:
        work = 9.0 * 3.8 * entries
        avg = ms_m / calls_m
        ach = work / (avg * 1e-3) / 1e12
        rows.append({"kernel": "k_nb_mask", "bound": "fp32", "avg_launch_ms": round(avg, 4), "launches": int(calls_m),
                     "ms_per_step": round(ms_m / steps, 4), "work_per_launch": work,
                     "work_model": "9 flop x 3.8 candidates per list entry (distance tests of the 5^3-bin stencil)",
                     "achieved": round(ach, 2), "peak": round(fp32_peak, 2), "unit": "TFLOP/s",
                     "frac": round(ach / fp32_peak, 4),
                     "ideal_ms_per_step": round(work / (fp32_peak * 1e12) * 1e3 * calls_m / steps, 4)})
    add("k_nb_fill", "k_nb_fill", "hbm", 4.0 * entries + 8.0 * N + entries / 8.0 * 3.8,
        "writes 4 B x entries + 8 B x N, reads the hit masks")
    add("k_rho_tiles", "k_rho_tiles", "hbm", 40.0 * N + 8.0 * G_tiles, "40 B x N + 8 B x tile-block points")
    add("k_rho_fold", "k_rho_fold", "hbm", 8.0 * G_tiles + 8.0 * F, "reads every tile-block point once, writes 8 B x F")
    if half_nx:
        Fs = F * (half_nx // 2 + 1) / half_nx
        nf = 3 if ncomp_packs == 2 else 1          # fields transformed back: Ex, Ey, Ez (ik) or u (ad)
        add("k_fft_x_r2c (x fwd, real in)", "k_fft_x_fwd", "hbm", 8.0 * F + 16.0 * Fs, "8 B x F in, 16 B x Fs out (half spectrum)")
        add("k_fft_pass y fwd", "k_fft_y_fwd", "hbm", 32.0 * Fs, "16 B x Fs in and out")
        add("k_fft_z_poisson", "k_fft_z_poisson", "hbm", (16.0 + 8.0 + 16.0 * nf) * Fs,
            "16 B x Fs in, 8 B x Fs Green's function, 16 B x Fs out per field")
        add("k_fft_pass y inv", "k_fft_y_inv", "hbm", 32.0 * nf * Fs, "16 B x Fs in and out per field")
        add("k_fft_x_c2r (x inv, real out)", "k_fft_x_inv", "hbm", 16.0 * Fs + 8.0 * F, "16 B x Fs in, 8 B x F out per field")
    else:
        add("k_fft_pass x fwd (real in)", "k_fft_x_fwd", "hbm", 24.0 * F, "8 B x F in, 16 B x F out")
        add("k_fft_pass y fwd", "k_fft_y_fwd", "hbm", 32.0 * F, "16 B x F in and out")
        add("k_fft_z_poisson", "k_fft_z_poisson", "hbm", (16.0 + 8.0 + 16.0 * ncomp_packs) * F,
            "16 B x F in, 8 B x F Green's function, 16 B x F out per packed field")
        add("k_fft_pass y inv", "k_fft_y_inv", "hbm", 32.0 * ncomp_packs * F, "16 B x F in and out per packed field")
        xinv_ms, xinv_calls = timers.get("k_fft_x_inv", (0.0, 0))
        if xinv_calls:
            per = 56.0 * F / 2.0 if ncomp_packs == 2 else 24.0 * F
            add("k_fft_pass x inv (real out)", "k_fft_x_inv", "hbm", per,
                "16 B x F in per packed field, 8 B x F out per field component (average of the two launches)")
    add("k_fieldforce", "fieldforce", "hbm", 104.0 * N + 24.0 * F,
        "40 B x N sorted atoms + 64 B x N force read-modify-write + 24 B x F field bricks (the L1 data pipe binds, "
        "see profiles/r02_ncu_k_fieldforce.txt)")
    add("k_nve_initial", "k_nve_initial", "hbm", (120.0 + (16.0 if flt == 4 else 0.0)) * N, "reads f,v,x writes v,x")
    add("k_nve_final", "k_nve_final", "hbm", 72.0 * N, "reads f,v writes v")
    ideal = sum(r["ideal_ms_per_step"] for r in rows)
    return rows, ideal, nbar


# ---------------------------------------------------------------------------------------------------------------------
# parity on the bounded sample: GPU path (decomposed over the N ranks) against the CPU oracle

def pair_set_signature(n, i, own):
    """order-independent per-atom signature of a set of (i, owner(j)) pairs: count, sum and sum of squares of the
    partner ids (exact in float64 for ids < 2^26)"""
    cnt = np.bincount(i, minlength=n).astype(np.float64)
    s1 = np.bincount(i, weights=own.astype(np.float64), minlength=n)
    s2 = np.bincount(i, weights=(own.astype(np.float64) % 4093.0) ** 2, minlength=n)
    return cnt, s1, s2


def parity_block(args, pkg, W, dist, rank, world, local_rank, prec):
    """styles of the selected config on data.aC x cpu_rep^3 (or the fcc sample): max force error relative to the largest
    force component, pair / k-space energy errors, pair-set equality (N = 1).  The oracle runs on rank 0."""
    import torch
    cfgname = args.config
    if cfgname not in ("buck_coul_long", "buck_coul_cut"):
        return {"skipped": "parity block covers the data.aC configurations; this configuration is parity-tested in tests/"}
    orc = graft.load_oracle()
    u = W.UNITS["metal"]
    rep = args.cpu_rep
    s = W.aC_system(rep)   # seeded displacement 0.02 A: a perfect crystal has no net forces
    n = len(s["x"])
    prd = s["boxhi"] - s["boxlo"]
    long_ = cfgname == "buck_coul_long"
    cut = 12.0 if long_ else 10.0
    skin = 0.3
    co = W.coeffs_aC(cut, cut)
    style = pkg.PAIR_BUCK_COUL_LONG if long_ else pkg.PAIR_BUCK_COUL_CUT
    grid, g = (None, 0.0)
    if long_:
        grid, g = pkg.pppm_init(args.acc, u["qqrd2e"], s["q"], n, cut, prd, order=ORDER)
    cf = pkg.pair_coeffs(style, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    ctx = pkg.Context(local_rank, prec)
    ctx.set_units(u["qqrd2e"], u["ftm2v"])
    ctx.set_box(s["boxlo"], s["boxhi"])
    if world > 1:
        ctx.comm_init_torch(dist, rank, world)
    mine, first = my_slab(s, rank, world)
    ctx.atoms_upload(mine["x"], mine["type"], s["mass"], v=mine["v"], q=mine["q"])
    ctx.neigh_setup(skin)
    ctx.pair_setup(style, 2, cf, g_ewald=g)
    if long_:
        ctx.pppm_setup(*grid, ORDER, g)
    ctx.nve_setup(u["dt"])
    th = ctx.setup_forces(1, 1)
    if world > 1:
        d = ctx.atoms_download_ids(("f",))
        gid = mine["gid"][d["ids"] - first]
        parts = [None] * world if rank == 0 else None
        dist.gather_object((gid, d["f"]), parts, dst=0)
        f_gpu = None
        if rank == 0:
            f_gpu = np.zeros((n, 3))
            for gi, fi in parts:
                f_gpu[gi] = fi
    else:
        f_gpu = ctx.atoms_download(("f",))["f"]
    out = None
    if rank == 0:
        oprec = orc.DOUBLE if prec == pkg.PREC_DOUBLE else orc.MIXED
        P = orc.Params(orc.BUCK_COUL_LONG if long_ else orc.BUCK_COUL_CUT, 2, co["A"], co["rho"], co["C"], co["cut_lj"],
                       co["cut_coul"], qqrd2e=u["qqrd2e"], g_ewald=g)
        t0 = time.perf_counter()
        fo, evo, aux = orc.pair_forces_periodic(P, oprec, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], skin)
        ftot = fo[:, :3].copy()
        ek = None
        if long_:
            pp = orc.PPPM(*grid, ORDER, g, s["boxlo"], s["boxhi"], u["qqrd2e"], prec=oprec)
            fk, ek, vk = pp.compute(s["x"], s["q"])
            ftot += fk
        t_cpu = time.perf_counter() - t0
        scale = np.abs(ftot).max()
        out = {"sample": "data.aC x %d^3 = %d atoms, seeded displacement 0.02 A%s; GPU path over %d rank(s) vs oracle/ "
                         "(half list, newton on; pinned bit for bit against oracle/_ref)" %
                         (rep, n, ", grid %dx%dx%d" % tuple(grid) if long_ else "", world),
               "max_rel_force_err": float(np.abs(f_gpu - ftot).max() / scale),
               "epair_rel": float(abs((th[0] + th[1]) - (evo[0] + evo[1])) / abs(evo[0] + evo[1])),
               "virial_rel": float(np.abs(th[2:8] - evo[2:8]).max() / np.abs(evo[2:8]).max()),
               "ekspace_rel": float(abs(th[8] - ek) / abs(ek)) if long_ else None,
               # mixed mode: this arm is the reference's DEFAULT newton-on half list; the device evaluates a cross-boundary
               # pair from both sides with each side's own float image (the reference's NEWTON_PAIR = 0 arrangement), so
               # float image noise (~5e-5 here, the size of the reference's own mixed-vs-double gap) remains against this
               # arm; the 1e-5 bar is met against the NEWTON_PAIR = 0 arm (tests/test_gpu_pair.py), which needs an
               # O(N^2) list and is not run at this size
               "tolerance": {"force": 1e-9 if prec == pkg.PREC_DOUBLE else 1e-4,
                             "energy": 1e-10 if prec == pkg.PREC_DOUBLE else 1e-5},
               "oracle_seconds": round(t_cpu, 2)}
        if world == 1:
            # pair set: the oracle's binned half list (each pair once) symmetrised, against the device's full list
            nn, off, ent, gsrc, gshift = ctx.neigh_download()
            i_g = np.repeat(np.arange(n, dtype=np.int64), nn)
            j_g = ent.astype(np.int64) & 0x3FFFFFFF
            own_g = np.where(j_g >= n, gsrc[np.maximum(j_g - n, 0)], j_g) if len(gsrc) else j_g
            sig_g = pair_set_signature(n, i_g, own_g)
            hn, hoff, hent, src = aux["numneigh"], aux["offsets"], aux["entries"], aux["src"]
            i_h = np.repeat(np.arange(n, dtype=np.int64), hn)
            j_h = hent.astype(np.int64) & 0x3FFFFFFF
            own_h = j_h.copy()
            gh = own_h >= n
            for _ in range(4):   # staged ghosts: a ghost's source may itself be a ghost
                if not gh.any():
                    break
                own_h[gh] = src[own_h[gh] - n]
                gh = own_h >= n
            sig_h = pair_set_signature(n, np.concatenate([i_h, own_h]), np.concatenate([own_h, i_h]))
            out["pair_set_equal"] = bool(all(np.array_equal(a, b) for a, b in zip(sig_g, sig_h)))
            if prec != pkg.PREC_DOUBLE and not out["pair_set_equal"]:
                # float distance test: a pair exactly at the list cut-off (outside the force cut-off by the skin) can be
                # in range seen from one periodic image and out of range seen from the other
                diff = int(np.abs(sig_g[0] - sig_h[0]).sum())
                out["pair_set_rim_entries"] = diff
                out["pair_set_equal"] = None if diff <= 1e-6 * len(ent) else False
            out["pair_set_entries"] = int(len(ent))
            out["pair_set_check"] = "per-atom count, sum and sum of squares of partner ids, full list vs symmetrised half list"
        else:
            out["pair_set_equal"] = None
        tol = out["tolerance"]
        out["ok"] = bool(out["max_rel_force_err"] <= tol["force"] and out["epair_rel"] <= tol["energy"] and
                         (out["ekspace_rel"] is None or out["ekspace_rel"] <= max(tol["energy"], 1e-9)) and
                         out["pair_set_equal"] is not False)
    ctx.close()
    del torch
    return out


# ---------------------------------------------------------------------------------------------------------------------

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
    cfg = Config(args, pkg, W, world)
    prec = pkg.PREC_DOUBLE if args.prec == "double" else pkg.PREC_MIXED
    rx, ry, rz = cfg.global_reps()
    blocks_z = rz
    s = rank_system(cfg.system, (rx, ry, rz), rank, world)
    u = W.UNITS[s["units"]]
    nlocal0 = len(s["x"])
    qsq_local = float(np.sum(s["q"] ** 2)) if s.get("q") is not None else 0.0
    if world > 1:
        t = torch.tensor([float(nlocal0), qsq_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        natoms, qsq = int(round(t[0].item())), float(t[1].item())
    else:
        natoms, qsq = nlocal0, qsq_local
    ctx = pkg.Context(local_rank, prec)
    ctx.set_units(u["qqrd2e"], u["ftm2v"])
    ctx.set_box(s["boxlo"], s["boxhi"])
    if world > 1:
        ctx.comm_init_torch(dist, rank, world)
    ctx.atoms_upload(s["x"], s["type"], s["mass"], v=s["v"], q=s.get("q"))
    desc = cfg.setup(ctx, s, natoms, qsq)
    kspace_only = not cfg.pair

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if kspace_only:
        ctx.pppm_compute(0, 0)
        st1 = dict(total=0, nghost=0, nbuilds=0)
        nb0 = 0
    else:
        ctx.setup_forces(0, 0)
        ctx.neigh_build()          # a second build settles every capacity-grown buffer before anything is timed
        nb0 = None

    # ---- warm-up, then the timed region: EXACTLY K steps, device-timed, max over ranks ---------------
    warm = max(args.warmup, 3)
    if kspace_only:
        for _ in range(warm):
            ctx.pppm_compute(0, 0)
    else:
        ctx.run(warm)
    ctx.timers_enable(True)
    ctx.timers_reset()
    l0 = ctx.launch_count()
    if not kspace_only:
        nb0 = ctx.neigh_stats()["nbuilds"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    if kspace_only:
        # PPPMIntel::compute alone: the library's phase timers (CUDA events on its stream) bracket every kernel of it
        for _ in range(args.steps):
            ctx.pppm_compute(0, 0)
        torch.cuda.synchronize()
        tms = ctx.timers()
        ms_dev = sum(tms[k][0] for k in ("make_rho", "fft", "poisson", "fieldforce", "comm"))
    else:
        ms_dev = ctx.run_timed(args.steps)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - l0
    timers = ctx.timers()
    ctx.timers_enable(False)
    if not kspace_only:
        st1 = ctx.neigh_stats()
    if world > 1:
        t = torch.tensor([ms_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev = float(t.item())
    ms_per_step = ms_dev / args.steps
    value = natoms * args.steps / (ms_dev * 1e-3)

    # ---- rooflines ---------------------------------------------------------------------------------------
    entries_local = int(st1["total"])
    entries = entries_local
    if world > 1:
        t = torch.tensor([float(entries_local)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        entries = int(t.item())
    fp_peak = ctx.microbench(0) if prec == pkg.PREC_DOUBLE else ctx.microbench(1)
    hbm_peak, hbm_src = peaks()
    grid = desc.get("grid")
    F = int(np.prod(grid)) // world if grid else 0
    G_tiles = 0
    if grid:
        E = 8 + ORDER - 1
        G_tiles = int(np.prod([-(-g // 8) for g in (grid[0], grid[1], max(grid[2] // world, 8))])) * E ** 3
    fp32_peak = ctx.microbench(1)
    rows, ideal_ms, nbar = kernel_rooflines(desc, timers, args.steps, nlocal0, nlocal0 + int(st1["nghost"]), entries_local,
                                            F, G_tiles, fp_peak, hbm_peak, args.prec, fp32_peak=fp32_peak,
                                            half_nx=(grid[0] if grid and pppm_half_spectrum(world) else 0))
    for r in rows:
        r["share_of_step"] = round(r["ms_per_step"] / ms_per_step, 4)
    rows = [r for r in rows if r["share_of_step"] >= 0.01 or r["kernel"] == "k_pair"]
    dominant = max(rows, key=lambda r: r["ms_per_step"]) if rows else None
    roofline = None
    if dominant:
        roofline = {"kernel": dominant["kernel"], "bound": "hbm" if dominant["bound"] == "hbm" else "tensor",
                    "bound_detail": dominant["bound"] + (" pipe (no tensor cores on this path: 'tensor' stands for the "
                                                         "compute roof)" if dominant["bound"] != "hbm" else ""),
                    "achieved": dominant["achieved"], "peak": dominant["peak"], "unit": dominant["unit"],
                    "frac": dominant["frac"], "traffic": None,
                    "traffic_note": "dram bytes per launch come from ncu --set full captures (profiles/); none exists for "
                                    "this configuration",
                    "peak_source": ("FMA microbenchmark in this run (b200md_microbench); MEASURED_PEAKS.json holds no "
                                    "FP64/FP32 figure") if dominant["bound"] != "hbm" else hbm_src,
                    "work_model": dominant["work_model"], "avg_launch_ms": dominant["avg_launch_ms"],
                    "share_of_step": dominant["share_of_step"]}

    if roofline:
        # dram__bytes_read + dram__bytes_write of one launch from the committed ncu --set full capture - only for the
        # very configuration the capture was taken on (same kernel, atoms, precision, one GPU; the list size drifts by
        # about 1 % with the step count at which the last rebuild falls, hence the 2 % window)
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as fh:
                cap = json.load(fh).get(roofline["kernel"])
            a = cap["applies_to"] if cap else None
            if (a and a["config"] == args.config and a["n_gpus"] == world and a["precision"] == args.prec and
                    not args.table and abs(entries_local - a["neighbor_entries"]) <= 2e-2 * a["neighbor_entries"]):
                roofline["traffic"] = cap["dram_bytes_per_launch"]
                roofline["traffic_note"] = cap["source"]
        except (OSError, ValueError, KeyError):
            pass

    # ---- e2e: the same step through host buffers (pinned), H2D x and D2H x,f every step ------------------
    e2e = None
    if not args.no_e2e and not kspace_only:
        nl = nlocal0
        xin = torch.empty((nl, 3), dtype=torch.float64).pin_memory()
        xout = torch.empty((nl, 3), dtype=torch.float64).pin_memory()
        fout = torch.empty((nl, 3), dtype=torch.float64).pin_memory()
        xin_n, xout_n, fout_n = xin.numpy(), xout.numpy(), fout.numpy()
        ok = True
        try:
            if world == 1:
                xin_n[:] = ctx.atoms_download(("x",))["x"]
            ne = max(3, min(args.steps, 10))
            if world > 1:
                # several GPUs: atoms migrate between ranks, so the host owns no fixed slice; every step downloads this
                # rank's ids, x and f in device order (b200md_step_host_ids); positions stay resident (no H2D)
                cap = int(nlocal0 * 1.1) + 4096
                ids = torch.empty((cap,), dtype=torch.int32).pin_memory().numpy()
                xo = torch.empty((cap, 3), dtype=torch.float64).pin_memory().numpy()
                fo = torch.empty((cap, 3), dtype=torch.float64).pin_memory().numpy()
                for _ in range(2):
                    ctx.step_host_ids(ids, xo, fo)
                barrier()
                te = time.perf_counter()
                for _ in range(ne):
                    ctx.step_host_ids(ids, xo, fo)
                barrier()
                te = time.perf_counter() - te
                h2d = 0
            else:
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
                h2d = natoms * 24
        except Exception as ex:   # noqa: BLE001  (reported, never silently dropped)
            ok = False
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)}
        if ok:
            if world > 1:
                t = torch.tensor([te], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                te = float(t.item())
            e2e = {"value": natoms * ne / te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": natoms * 48,
                   "steps": ne,
                   "note": ("b200md_step_host: pinned host x in, x and f out every step; includes the host-side un-permute"
                            if world == 1 else
                            "b200md_step_host_ids: pinned host ids, x and f out every step in device order (atoms migrate "
                            "between the ranks, positions stay resident: no upload)")}
            if world > 1:
                e2e["d2h_bytes_per_step"] = natoms * 52
    elif kspace_only:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "k-space-only configuration: see tests/ for b200md_pppm_compute_host"}

    phase = {k: round(v[0] / args.steps, 4) for k, v in timers.items() if v[1] > 0 and not k.startswith("k_")}
    ctx.close()

    # ---- parity on the bounded sample (all ranks take part) ------------------------------------------------
    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(args, pkg, W, dist, rank, world, local_rank, prec)
        except Exception as ex:   # noqa: BLE001
            parity = {"error": str(ex)}

    # ---- CPU baseline: the oracle's whole-step loop on the host cores, bounded sample ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and cfg.name == "buck_coul_long":
        cpu = cpu_baseline(args, W)

    if rank == 0:
        rebuilds = int(st1["nbuilds"] - nb0) if not kspace_only else 0
        out = {
            "metric": cfg.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if prec == pkg.PREC_DOUBLE else "f32 compute / f64 accumulate", "data": "synthetic",
            "config": {"workload": "%s x %dx%dx%d (%d atoms) %s, %s, nve dt %g" %
                                   ("data.aC" if "aC" in str(cfg.name) or cfg.name.startswith("buck_coul") else
                                    ("data.spce" if cfg.name.startswith("spce") else "fcc rho* 0.8442"),
                                    rx, ry, blocks_z, natoms, desc["style"], desc["neigh"], u["dt"]),
                       "name": cfg.name, "precision": args.prec, "coulomb": "table" if args.table else "analytic erfc",
                       "grid": desc.get("grid"), "g_ewald": desc.get("g_ewald", desc.get("g_ewald_6")),
                       "neighbor_entries": int(entries), "nbar": round(nbar, 1), "nghost": int(st1["nghost"]),
                       "rebuilds_in_timed_region": rebuilds,
                       "pppm_transforms": (("half spectrum (real-to-complex)" if pppm_half_spectrum(world) else
                                            "complex-to-complex (B200MD_R2C=0)") if desc.get("grid") else None),
                       "l2": "inputs larger than L2 (neighbour list %.1f GB, grids %.2f GB per GPU)" %
                             (4.0 * entries_local / 1e9, 56.0 * F / 1e9),
                       "parallelism": ("z-slab x%d (%s geometry)" % (world, args.geometry)) if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "roofline_kernels": rows,
            "step_roofline_frac": round(ideal_ms / ms_per_step, 4) if ms_per_step > 0 else None,
            "step_ideal_ms": round(ideal_ms, 4),
            "parity": parity, "cpu_baseline": cpu,
            "same_config_as_cpu_arm": False if cpu else None,
            "cpu_arm_note": ("the CPU arms run data.aC x %d^3; the GPU arm runs x %d^3: the metric is intensive in N, the "
                             "ratio compares throughput, not wall time of one job" % (args.cpu_rep, cfg.rep)) if cpu else None,
            "phase_ms_per_step": phase, "wall_s_timed_region": wall,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_setup(W, rep, acc):
    orc = graft.load_oracle()
    pkg = graft.load_package()
    u = W.UNITS["metal"]
    s = W.aC_system(rep, jitter=0.0)
    natoms = len(s["x"])
    cut, skin = 12.0, 0.3
    grid, g_ewald = pkg.pppm_init(acc, u["qqrd2e"], s["q"], natoms, cut, s["boxhi"] - s["boxlo"], order=ORDER)
    co = W.coeffs_aC(cut, cut)
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=g_ewald)
    pp = orc.PPPM(*grid, ORDER, g_ewald, s["boxlo"], s["boxhi"], u["qqrd2e"])
    kw = dict(prec=orc.DOUBLE, skin=skin, every=1, delay=0, check=1, dt=u["dt"], ftm2v=u["ftm2v"], pppm=pp)
    return orc, s, P, kw, natoms, grid


def reference_step_loop(args, W, steps, warm):
    """the whole Verlet step with every function the reference ships — PairBuckCoulLongIntel::compute, PPPMIntel::compute,
    FixNVEIntel::initial_integrate / final_integrate — running from the reference's OWN translation units, compiled
    unchanged into oracle/_ref/libref.so (oracle/Makefile.ref), all host cores; only what the reference inherits from
    upstream LAMMPS (neighbour list, ghosts, forward / reverse communication, the re-neighbouring decision) comes from the
    oracle (tests/refmd.py).  Returns None where oracle/_ref is not available."""
    graft.load_oracle()
    import refc
    if not refc.available():
        return None
    import refmd
    cores = os.cpu_count() or 1
    orc, s, P, kw, natoms, grid = cpu_setup(W, args.cpu_rep, args.acc)
    rm = refmd.RefMD(s, P, nthreads=cores, **kw)
    rm.run(max(warm, 1))                   # builds the list, first forces, warm-up steps
    t0, w0, nb0 = dict(rm.t), time.perf_counter(), rm.nbuilds
    rm.run(steps)
    wall = time.perf_counter() - w0
    ph = {k: rm.t[k] - t0[k] for k in rm.t}
    secs = sum(v for k, v in ph.items() if k != "harness")
    return {"value": natoms * steps / secs, "unit": UNIT, "cores": cores, "kind": "reference", "seconds": secs,
            "steps": steps, "atoms": natoms, "grid": list(grid), "rebuilds": rm.nbuilds - nb0,
            "wall_seconds_with_marshalling": round(wall, 3),
            "phase_s": {k: round(v, 3) for k, v in ph.items()},
            "sample": "data.aC x %d^3 = %d atoms, grid %dx%dx%d, %d steps, %.1f s: PairBuckCoulLongIntel::compute, "
                      "PPPMIntel::compute and FixNVEIntel::initial/final_integrate of /root/reference compiled UNCHANGED "
                      "(oracle/_ref: g++ -O3 -fopenmp, %d threads; not the ICC build), the upstream neighbour list / ghosts / "
                      "communication around them from the oracle; timed = the reference's members + that glue, without the "
                      "marshalling between the two libraries (%.1f s of wall time in all)"
                      % (args.cpu_rep, natoms, *grid, steps, secs, cores, wall)}


def port_step_loop(args, W, steps, warm):
    """oracle/ whole-step loop (restatement of the same loops; bit-identical to oracle/_ref) — the fallback where
    oracle/_ref is not available, and a second opinion beside it"""
    cores = os.cpu_count() or 1
    orc, s, P, kw, natoms, grid = cpu_setup(W, args.cpu_rep, args.acc)
    md = orc.MD(s, P, **kw)
    md.run(max(warm, 1), cores)
    t0 = time.perf_counter()
    tm = md.run(steps, cores)
    dt = time.perf_counter() - t0
    return {"value": natoms * steps / dt, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt, "steps": steps,
            "atoms": natoms, "grid": list(grid), "rebuilds": tm["nbuilds"],
            "phase_s": {k: round(v, 3) for k, v in tm.items() if k != "nbuilds"},
            "sample": "data.aC x %d^3 = %d atoms, grid %dx%dx%d, %d steps, %.1f s; oracle/ restatement of the reference "
                      "loops (g++ -O3 -march=native -fopenmp, %d threads; bit-identical to the reference's own compiled "
                      "loops, oracle/_ref), not the ICC USER-INTEL build" % (args.cpu_rep, natoms, *grid, steps, dt, cores)}


def cpu_baseline(args, W):
    """the CPU arm on a bounded sample (data.aC x cpu_rep^3, same styles / accuracy; the metric is intensive in N): the
    reference's own compiled units where oracle/_ref exists (kind "reference"), else the oracle port (kind "port")"""
    try:
        ref = reference_step_loop(args, W, args.cpu_steps, 1)
    except Exception as ex:   # noqa: BLE001
        ref = None
        err = str(ex)
    else:
        err = None
    port = port_step_loop(args, W, args.cpu_steps, 1)
    out = dict(ref or port)
    out["port"] = {k: port[k] for k in ("value", "seconds", "phase_s", "rebuilds")} if ref else None
    if err:
        out["reference_error"] = err
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads, same metric / config, each step
    a bounded sample of the workload.  The reference's own compiled translation units (oracle/_ref) run every function
    it ships; where they are not available the oracle port of the same loops stands in (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = importlib.import_module("lammps_buck_intel_b200.workloads") if graft.load_package() else None
    warm = max(args.warmup, 1)
    try:
        r = reference_step_loop(args, W, args.steps, warm)
    except Exception as ex:   # noqa: BLE001
        print("bench.py: oracle/_ref failed (%s); timing the oracle port" % ex, file=sys.stderr)
        r = None
    if r is None:
        r = port_step_loop(args, W, args.steps, warm)
    value, natoms, grid = r["value"], r["atoms"], r["grid"]
    out = {"impl": "reference", "metric": "atom-timesteps/s buck/coul/long+PPPM", "value": value, "unit": UNIT,
           "n_gpus": args.gpus, "steps": args.steps,
           "warmup": warm, "ms_per_step": r["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "data.aC x %d^3 (%d atoms per step sample) buck/coul/long 12.0 + pppm %g order %d grid "
                                  "%dx%dx%d, skin 0.3 check yes, nve; bounded sample of the 4.05 M-atom workload "
                                  "(metric is intensive in N)" % (args.cpu_rep, natoms, args.acc, ORDER, *grid)},
           "same_config": False, "sample_atoms": natoms,
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "phase_s", "rebuilds") if k in r},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if "wall_seconds_with_marshalling" in r:
        out["wall_seconds_with_marshalling"] = r["wall_seconds_with_marshalling"]
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
