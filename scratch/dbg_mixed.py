import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package(); orc = g.load_oracle()
import importlib
W = importlib.import_module("lammps_buck_intel_b200.workloads")
u = W.UNITS["metal"]
co = W.coeffs_aC(12.0, 12.0)
ge = 0.2776
P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"], g_ewald=ge)
cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
ctx = pkg.Context(0, 1)
ctx.set_units(u["qqrd2e"], u["ftm2v"])
ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=ge)
rng = np.random.default_rng(0)
n = 20000
# n dimers: atom 2k (type 1) and 2k+1 (type 2), far apart from other dimers
base = rng.uniform(0, 25, (n, 3)) + np.arange(n)[:, None] * 100.0
r = rng.uniform(1.4, 11.9, n)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
x = np.empty((2 * n, 3)); x[0::2] = base; x[1::2] = base + r[:, None] * d
t = np.tile(np.array([1, 2], np.int32), n)
q = np.tile(np.array([2.96653, -1.483265]), n)
nn = np.ones(2 * n, np.int32); off = np.arange(2 * n + 1, dtype=np.int64)
ent = np.arange(2 * n, dtype=np.int32) ^ 1
# make all neighbours "ghost-like" for the oracle newton=0 semantics: duplicate
x2 = np.concatenate([x, x]); t2 = np.concatenate([t, t]); q2 = np.concatenate([q, q])
ent2 = ent + 2 * n
f, ev = ctx.pair_eval_host(1, 1, 2 * n, x2, t2, q2, nn, off[:-1], ent2)
fo, evo = orc.pair_eval(P, 1, 1, 1, 2 * n, x2, t2, q2, nn, off, ent2, newton=0)
fd, evd = orc.pair_eval(P, 0, 1, 1, 2 * n, x2, t2, q2, nn, off, ent2, newton=0)
fm = np.linalg.norm(fd[:2 * n, :3], axis=1)
e_go = np.linalg.norm(f[:, :3] - fo[:2 * n, :3], axis=1) / fm
e_od = np.linalg.norm(fo[:2 * n, :3] - fd[:2 * n, :3], axis=1) / fm
print("per-pair rel err gpu-vs-oracle(mixed): median %.2e  99%% %.2e max %.2e ; exact-equal fraction %.3f" % (np.median(e_go), np.percentile(e_go, 99), e_go.max(), (e_go == 0).mean()))
print("per-pair rel err oracle(mixed)-vs-double: median %.2e 99%% %.2e max %.2e" % (np.median(e_od), np.percentile(e_od, 99), e_od.max()))
w = np.argsort(e_go)[-5:]
for k in w:
    print("r=%.4f type %d  f_gpu %s f_orc %s" % (r[k // 2], t[k], f[k, :3], fo[k, :3]))
