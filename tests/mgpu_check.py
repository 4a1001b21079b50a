"""2..N-rank correctness check of the slab decomposition: every rank uploads its z slab of a small charged system,
runs setup_forces + a few steps; forces / energies / positions are compared with a single-GPU run of the same system
on rank 0 (a second context).  Launch: python -m torch.distributed.run --nproc-per-node N tests/mgpu_check.py (or pytest tests/test_gpu_multi.py)"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft

pkg = graft.load_package()
W = importlib.import_module("lammps_buck_intel_b200.workloads")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
rep = (2, 2, int(os.environ.get("REPZ", "2")) * world)
nsteps = int(os.environ.get("NSTEPS", "12"))
DIFF = int(os.environ.get("DIFF", "0"))   # kspace_modify diff: 0 ik, 1 ad
DISP = int(os.environ.get("DISP", "0"))   # 1: add the geometric-mixing dispersion grid of pppm/disp (second PPPM state)
s = W.aC_system(rep, jitter=0.05)
u = W.UNITS["metal"]
n = len(s["x"])
prd = s["boxhi"] - s["boxlo"]
cut = 8.0
grid, g = pkg.pppm_init(1e-4, u["qqrd2e"], s["q"], n, cut, prd)
co = W.coeffs_aC(cut, cut)
cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
s["v"] = s["v"] * 3.0   # hot, so that atoms migrate between the slabs within a few steps


def setup(ctx, sel):
    ctx.atoms_upload(s["x"][sel], s["type"][sel], s["mass"], v=s["v"][sel], q=s["q"][sel])
    ctx.neigh_setup(0.6)
    ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=g)
    ctx.pppm_setup(*grid, 5, g, differentiation=DIFF)
    if DISP:
        ctx.pppm_setup(60, 60, 32 * rep[2], 5, 0.31, dispersion=1, B=np.array([0.0, 9.0, 13.2]))
    ctx.nve_setup(u["dt"])
    return ctx.setup_forces(1, 1)


ctx = pkg.Context(lr)
ctx.set_units(u["qqrd2e"], u["ftm2v"])
ctx.set_box(s["boxlo"], s["boxhi"])
ctx.comm_init_torch(dist, rank, world)
slab = prd[2] / world
own = np.floor((s["x"][:, 2] - s["boxlo"][2]) / slab).astype(int).clip(0, world - 1) == rank
first = int(np.sum(np.floor((s["x"][:, 2] - s["boxlo"][2]) / slab).astype(int).clip(0, world - 1) < rank))
gid = np.nonzero(own)[0]    # global id k of this rank's upload = first + k  ->  original index gid[k]
th0 = setup(ctx, own)
d0 = ctx.atoms_download_ids(("f",))
th1 = ctx.run(nsteps, thermo=True)
d1 = ctx.atoms_download_ids(("x", "f"))
st = ctx.neigh_stats()

# gather (original index, x, f) to rank 0
def gather(d, keys):
    # ids are global: rank r's uploads occupy [first_r, first_r + n_r); map back to the original atom index
    firsts = [None] * world
    dist.all_gather_object(firsts, (first, gid))
    allf = [None] * world
    dist.all_gather_object(allf, {k: d[k] for k in keys + ["ids"]})
    out = {k: np.zeros((n, 3)) for k in keys}
    seen = np.zeros(n, int)
    starts = np.array([f[0] for f in firsts])
    for blk in allf:
        r = np.searchsorted(starts, blk["ids"], side="right") - 1
        orig = np.array([firsts[rr][1][i - firsts[rr][0]] for rr, i in zip(r, blk["ids"])], int)
        seen[orig] += 1
        for k in keys:
            out[k][orig] = blk[k]
    assert (seen == 1).all(), "atoms lost or duplicated by migration: %s" % np.bincount(seen)
    return out

g0 = gather(d0, ["f"])
g1 = gather(d1, ["x", "f"])
nb = [None] * world
dist.all_gather_object(nb, (st["nbuilds"], int(len(d1["ids"]))))
if rank == 0:
    ref = pkg.Context(lr)
    ref.set_units(u["qqrd2e"], u["ftm2v"])
    ref.set_box(s["boxlo"], s["boxhi"])
    r0 = setup(ref, slice(None))
    f0 = ref.atoms_download(("f",))["f"]
    r1 = ref.run(nsteps, thermo=True)
    dr = ref.atoms_download(("x", "f"))
    fs = np.abs(f0).max()
    e = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    print("ranks %d atoms %d grid %s diff %s builds/owned per rank %s" % (world, n, grid, ("ad" if DIFF else "ik") + (" + disp grid" if DISP else ""), nb))
    print("step 0: force err %.3e  epair err %.3e  ekspace err %.3e  virial err %.3e" %
          (np.abs(g0["f"] - f0).max() / fs, abs(th0[0] + th0[1] - r0[0] - r0[1]) / abs(r0[0] + r0[1]),
           abs(th0[8] - r0[8]) / abs(r0[8]), e(th0[2:8] + th0[9:15], r0[2:8] + r0[9:15])))
    dx = g1["x"] - dr["x"]
    dx -= np.round(dx / prd) * prd
    print("step %d: x err %.3e  force err %.3e  etot err %.3e  ke err %.3e" %
          (nsteps, np.abs(dx).max(), np.abs(g1["f"] - dr["f"]).max() / fs,
           abs(th1[0] + th1[1] + th1[8] - r1[0] - r1[1] - r1[8]) / abs(r1[0] + r1[1] + r1[8]), abs(th1[15] - r1[15]) / r1[15]))
    ok = np.abs(g0["f"] - f0).max() / fs < 1e-9 and np.abs(dx).max() < 1e-8
    print("MGPU CHECK", "OK" if ok else "FAILED")
    ref.close()
ctx.close()
dist.destroy_process_group()
