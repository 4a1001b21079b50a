"""Freeze the CPU oracle's outputs as golden vectors (the reference ships none, the build authors its own).

    python tests/golden/make_golden.py        # writes tests/golden/*.npz

Inputs are the seeded systems of lammps-buck-intel_b200/workloads.py; outputs are what oracle/ (the restatement of
pair_buck*_intel.cpp / pppm_intel.cpp / pppm_disp_intel.cpp pinned by tests/test_oracle_kat.py) computes for them.
tests/test_golden.py checks (CPU) that the oracle still reproduces them and (GPU) that the device path matches them.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

CASES = {
    # name: (system, style, cut_lj, cut_coul, g_ewald, g_ewald_6, order1, order6, tables)
    "buck": ("fcc6", "BUCK", 2.5, None, 0.0, 0.0, 0, 0, False),
    "buck_coul_cut": ("aC1", "BUCK_COUL_CUT", 7.5, 10.0, 0.0, 0.0, 0, 0, False),
    "buck_coul_long": ("aC1", "BUCK_COUL_LONG", 12.0, 12.0, 0.2776, 0.0, 0, 0, False),
    "buck_coul_long_table": ("aC1", "BUCK_COUL_LONG", 12.0, 12.0, 0.2776, 0.0, 0, 0, True),
    "buck_long_coul_long": ("aC1", "BUCK_LONG_COUL_LONG", 10.0, 10.0, 0.29, 0.33, 1, 1, False),
    "buck_long_coul_long_table": ("aC1", "BUCK_LONG_COUL_LONG", 10.0, 10.0, 0.29, 0.33, 1, 1, True),
}
PPPM_CASES = {
    # name: (grid, order, g_ewald, diff_ad, dispersion)
    "pppm_ik5": ((24, 24, 27), 5, 0.28, 0, False),
    "pppm_ad4": ((24, 24, 27), 4, 0.28, 1, False),
    "pppm_ik7": ((30, 30, 32), 7, 0.31, 0, False),
    "pppm_disp_g5": ((30, 30, 32), 5, 0.31, 0, True),
}
B_DISP = np.array([0.0, 9.0, 13.2])
# later additions, kept apart so that the first ten files stay byte-identical:
#   name: (grid, order, g_ewald, diff_ad, dispersion, slab_volfactor, per-atom tallies)
PPPM_CASES2 = {
    "pppm_ik5_peratom": ((24, 24, 27), 5, 0.28, 0, False, 1.0, True),
    "pppm_disp_g5_ad": ((30, 30, 32), 5, 0.31, 1, True, 1.0, False),
    "pppm_slab3_ik5_peratom": ((24, 24, 80), 5, 0.28, 0, False, 3.0, True),
}


def pppm_case2(W, orc, name, prec=0):
    """-> (system, f, e, v, eatom or None, vatom or None) from the oracle"""
    grid, order, g, ad, disp, slab, peratom = PPPM_CASES2[name]
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    if disp:
        pp = orc.PPPM.dispersion(*grid, order, g, s["boxlo"], s["boxhi"], diff_ad=ad, prec=prec)
        w = B_DISP[s["type"]]
    else:
        pp = orc.PPPM(*grid, order, g, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad, slab=slab, prec=prec)
        w = s["q"]
    f, e, v = pp.compute(s["x"], w, eflag=3 if peratom else 1, vflag=5 if peratom else 1)
    ea, va = pp.peratom() if peratom else (None, None)
    return s, f, e, v, ea, va


def system(W, name):
    return W.fcc_system(6, 6, 6) if name == "fcc6" else W.aC_system(1)


def pair_case(pkg, W, orc, name, prec=0):
    sysn, style, cl, cc, ge, g6, o1, o6, tables = CASES[name]
    s = system(W, sysn)
    u = W.UNITS[s["units"]]
    co = W.coeffs_in_buck(cl) if sysn == "fcc6" else W.coeffs_aC(cl, cc)
    P = orc.Params(getattr(orc, style), s["ntypes"], co["A"], co["rho"], co["C"], co["cut_lj"], co.get("cut_coul"),
                   qqrd2e=u["qqrd2e"], g_ewald=ge, g_ewald_6=g6, order1=o1, order6=o6)
    ct = dt = None
    if tables and (style == "BUCK_COUL_LONG" or o1):
        ct = pkg.init_coul_tables(cc, ge, u["qqrd2e"])
        P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
    if tables and o6:
        dt = pkg.init_disp_tables(cl, g6)
        P.set_disp_tables(dt[0], 12, dt[1], dt[2], dt[3])
    return s, u, co, P, ct, dt


def main():
    pkg = graft.load_package()
    orc = graft.load_oracle()
    W = importlib.import_module("lammps_buck_intel_b200.workloads")
    for name in PPPM_CASES2:
        s, f, e, v, ea, va = pppm_case2(W, orc, name)
        extra = {} if ea is None else dict(eatom=ea, vatom=va)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), f=f, e=e, v=v, **extra)
        print(name, "e", e)
    if "--new-only" in sys.argv:
        return
    for name in CASES:
        s, u, co, P, ct, dt = pair_case(pkg, W, orc, name)
        f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3, eflag=3, vflag=1,
                                            eatom=1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), f=f[:, :3], eatom=f[:, 3], ev=ev)
        print(name, "ev", ev[:2])
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    for name, (grid, order, g, ad, disp) in PPPM_CASES.items():
        if disp:
            pp = orc.PPPM.dispersion(*grid, order, g, s["boxlo"], s["boxhi"])
            f, e, v = pp.compute(s["x"], B_DISP[s["type"]])
        else:
            pp = orc.PPPM(*grid, order, g, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad)
            f, e, v = pp.compute(s["x"], s["q"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), f=f, e=e, v=v)
        print(name, "e", e)


if __name__ == "__main__":
    main()
