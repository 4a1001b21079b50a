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
# DISP: 1 adds the geometric-mixing dispersion grid of pppm/disp (second PPPM state), 2 the arithmetic-mixing one
# (seven components), 3 the no-mixing one (eigen-components)
DISP = int(os.environ.get("DISP", "0"))
# MODE=spce: the pair + k-space path of in.spce on data.spce x (1,1,N): lj/long/coul/long cut long, special bonds on the
# lists built on the device (b200md_atoms_set_special with global partner ids)
MODE = os.environ.get("MODE", "aC")
GROUP = int(os.environ.get("GROUP", "0"))   # 1: fix nve on a sub-group with per-atom masses (b200md_nve_set_group)
if MODE == "spce":
    s = W.spce_system((1, 1, world))
    u = W.UNITS["real"]
    cut, skin = 8.8, 2.0
    co = W.coeffs_spce()
    style = pkg.PAIR_LJ_LONG_COUL_LONG
else:
    s = W.aC_system(rep, jitter=0.05)
    u = W.UNITS["metal"]
    cut, skin = 8.0, 0.6
    co = W.coeffs_aC(cut, cut)
    style = pkg.PAIR_BUCK_COUL_LONG
n = len(s["x"])
prd = s["boxhi"] - s["boxlo"]
grid, g = pkg.pppm_init(1e-4, u["qqrd2e"], s["q"], n, cut, prd)
cf = pkg.pair_coeffs(style, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
s["v"] = s["v"] * 3.0   # hot, so that atoms migrate between the slabs within a few steps
slab = prd[2] / world
owner = np.floor((s["x"][:, 2] - s["boxlo"][2]) / slab).astype(int).clip(0, world - 1)
# global id of an atom = (atoms of lower ranks) + its position in its rank's upload
orig2gid = np.zeros(n, np.int64)
for r in range(world):
    sel_r = np.nonzero(owner == r)[0]
    orig2gid[sel_r] = int((owner < r).sum()) + np.arange(len(sel_r))
ingroup_all = (np.arange(n) % 3 != 0).astype(np.int32)
rmass_all = s["mass"][s["type"]] * (1.0 + 0.1 * (np.arange(n) % 5))
if MODE == "spce" and not GROUP:
    # without the bonded terms and SHAKE of in.spce the hydrogens have no repulsive wall: they stay in place (fix nve
    # on the oxygens only, as bench.py --config spce does), the oxygens move and migrate
    GROUP = 2
    ingroup_all = (s["type"] == 1).astype(np.int32)


def water_specials(nn):
    nspecial = np.zeros((nn, 3), np.int32)
    special = np.zeros((nn, 2), np.int32)
    o = np.arange(0, nn, 3)
    nspecial[o] = (2, 2, 2)
    special[o, 0], special[o, 1] = o + 1, o + 2
    for h, other in ((o + 1, o + 2), (o + 2, o + 1)):
        nspecial[h] = (1, 2, 2)
        special[h, 0], special[h, 1] = o, other
    return nspecial, special


def setup(ctx, sel, global_ids):
    ctx.atoms_upload(s["x"][sel], s["type"][sel], s["mass"], v=s["v"][sel], q=s["q"][sel])
    ctx.neigh_setup(skin)
    if MODE == "spce":
        sl = (1, 0.0, 0.0, 0.5)
        ctx.pair_setup(style, 2, cf, special_lj=sl, special_coul=sl, g_ewald=g, ewald_order=1 << 1)
        nsp, sp = water_specials(n)
        # rows of this upload, partners as global ids (one GPU: upload indices = the original ones)
        ctx.atoms_set_special(nsp[sel], (orig2gid[sp[sel]] if global_ids else sp[sel]).astype(np.int32))
    else:
        ctx.pair_setup(style, 2, cf, g_ewald=g)
    ctx.pppm_setup(*grid, 5, g, differentiation=DIFF)
    eps = np.array([0.0, 0.8, 2.1]); sig = np.array([0.0, 2.9, 3.6])
    mesh6 = (60, 60, 32 * rep[2])
    if DISP == 1:
        ctx.pppm_setup(*mesh6, 5, 0.31, dispersion=1, B=np.array([0.0, 9.0, 13.2]), differentiation=DIFF)
    elif DISP == 2:
        c = np.sqrt([1.0, 6.0, 15.0, 20.0, 15.0, 6.0, 1.0])
        B7 = np.array([np.sqrt(eps[i]) / 4.0 * c * sig[i] ** np.arange(7) for i in range(3)])
        ctx.pppm_setup(*mesh6, 5, 0.31, dispersion=2, B=B7, differentiation=DIFF)
    elif DISP == 3:
        Cij = 4.0 * np.sqrt(np.outer(eps, eps)) * ((sig[:, None] + sig[None, :]) / 2.0) ** 6
        Cij[1, 2] = Cij[2, 1] = 0.7 * Cij[1, 2]
        ctx.pppm_setup(*mesh6, 5, 0.31, dispersion=3, B=Cij, differentiation=DIFF)
    if GROUP:
        ctx.nve_set_group(ingroup_all[sel], rmass_all[sel] if GROUP == 1 else None)
    ctx.nve_setup(u["dt"])
    return ctx.setup_forces(1, 1)


ctx = pkg.Context(lr)
ctx.set_units(u["qqrd2e"], u["ftm2v"])
ctx.set_box(s["boxlo"], s["boxhi"])
ctx.comm_init_torch(dist, rank, world)
own = owner == rank
first = int((owner < rank).sum())
gid = np.nonzero(own)[0]    # global id k of this rank's upload = first + k  ->  original index gid[k]
th0 = setup(ctx, own, True)
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
    r0 = setup(ref, slice(None), False)
    f0 = ref.atoms_download(("f",))["f"]
    r1 = ref.run(nsteps, thermo=True)
    dr = ref.atoms_download(("x", "f"))
    fs = np.abs(f0).max()
    e = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    what = ("ad" if DIFF else "ik") + ("", " + disp grid (geometric)", " + disp grid (arithmetic, 7 components)",
                                       " + disp grid (no mixing rule)")[DISP]
    what += (" + special bonds (data.spce)" if MODE == "spce" else "") + ("", " + nve group / rmass", " + nve on the oxygens")[GROUP]
    print("ranks %d atoms %d grid %s diff %s builds/owned per rank %s" % (world, n, grid, what, nb))
    print("step 0: force err %.3e  epair err %.3e  ekspace err %.3e  virial err %.3e" %
          (np.abs(g0["f"] - f0).max() / fs, abs(th0[0] + th0[1] - r0[0] - r0[1]) / abs(r0[0] + r0[1]),
           abs(th0[8] - r0[8]) / abs(r0[8]), e(th0[2:8] + th0[9:15], r0[2:8] + r0[9:15])))
    dx = g1["x"] - dr["x"]
    dx -= np.round(dx / prd) * prd
    print("step %d: x err %.3e  force err %.3e  etot err %.3e  ke err %.3e" %
          (nsteps, np.abs(dx).max(), np.abs(g1["f"] - dr["f"]).max() / fs,
           abs(th1[0] + th1[1] + th1[8] - r1[0] - r1[1] - r1[8]) / abs(r1[0] + r1[1] + r1[8]), abs(th1[15] - r1[15]) / r1[15]))
    ev_err = max(abs(th0[0] + th0[1] - r0[0] - r0[1]) / abs(r0[0] + r0[1]), abs(th0[8] - r0[8]) / abs(r0[8]),
                 e(th0[2:8] + th0[9:15], r0[2:8] + r0[9:15]))
    ok = np.abs(g0["f"] - f0).max() / fs < 1e-9 and np.abs(dx).max() < 1e-8 and ev_err < 1e-10
    print("MGPU CHECK", "OK" if ok else "FAILED")
    ref.close()
ctx.close()
dist.destroy_process_group()
