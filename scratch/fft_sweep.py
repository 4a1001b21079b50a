"""PPPM kernel times on the bench grid (4.05 M atoms, 250x250x270) for the FFT tuning knobs given in the environment."""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
W = importlib.import_module(ge.PKG_NAME + ".workloads")
s = W.aC_system(15)
ctx = pkg.make_context(s)
ctx.neigh_setup(0.3)
ctx.pppm_setup(250, 250, 270, 5, 0.248)
ctx.timers_enable(True)
for _ in range(3):
    ctx.pppm_compute(0, 0)
ctx.timers_reset()
n = 10
for _ in range(n):
    ctx.pppm_compute(0, 0)
t = ctx.timers()
keys = [k for k in t if k.startswith("k_fft") or k in ("k_rho_tiles", "k_rho_fold", "fieldforce", "make_rho", "fft")]
print(" ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("B200MD_")) or "defaults",
      "|", "  ".join("%s %.3f" % (k.replace("k_fft_", ""), t[k][0] / n) for k in keys))
