"""Runs the drop-in translation units of lammps-buck-intel_b200/integration/ on the GPU (through tests/integration_harness.cpp:
stand-in LAMMPS objects, the classes of the reference's own headers, init_style / init / compute) and compares atom->f, the
pair tallies and the k-space energy / virial with the oracle.  Prints INTEGRATION OK and exits 0 when everything agrees.

    python tests/integration_check.py          (on a B200; oracle/_ref/libinteg.so prebuilt or /root/reference present)

Three force evaluations per case with the positions displaced in between: the first uploads the atoms and builds the
device list (neighbor->ago == 0), the others take the positions-only path of the host-stepped deployment."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import __graft_entry__ as g  # noqa: E402


def main():
    pkg = g.load_package()
    pkg.load()
    import integ
    import orc
    W = importlib.import_module("lammps_buck_intel_b200.workloads")
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    co = W.coeffs_aC(8.0, 8.0)
    n = len(s["x"])
    ge, grid, skin = 0.30, (24, 24, 27), 0.3
    rng = np.random.default_rng(5)
    dx = rng.uniform(-0.01, 0.01, (n, 3))
    nsteps = 3
    xfin = W.wrap(s["x"] + (nsteps - 1) * dx, s["boxlo"], s["boxhi"])   # the oracle bins wrapped positions; forces do not care
    worst = 0.0
    for table in (False, True):
        for use_grid in (True, False):
            P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                           g_ewald=ge)
            if table:
                ct = pkg.init_coul_tables(8.0, ge, u["qqrd2e"])
                P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
            f, ev, ek, vk = integ.buck_coul_long(P, orc.DOUBLE, s, co["A"], co["rho"], co["C"], co["cut_lj"], 8.0, skin,
                                                 grid=grid if use_grid else None, nsteps=nsteps, dx=dx)
            fo, evo, _ = orc.pair_forces_periodic(P, orc.DOUBLE, xfin, s["type"], s["q"], s["boxlo"], s["boxhi"], skin)
            fref = fo[:, :3].copy()
            eko, vko = 0.0, np.zeros(6)
            if use_grid:
                fk, eko, vko = orc.PPPM(*grid, 5, ge, s["boxlo"], s["boxhi"], u["qqrd2e"]).compute(xfin, s["q"])
                fref += fk
            ferr = np.abs(f - fref).max() / np.abs(fref).max()
            eerr = max(abs(ev[0] - evo[0]) / abs(evo[0]), abs(ev[1] - evo[1]) / abs(evo[1]))
            verr = np.abs(ev[2:] - evo[2:8]).max() / np.abs(evo[2:8]).max()
            kerr = abs(ek - eko) / abs(eko) if use_grid else 0.0
            kverr = np.abs(vk - vko).max() / np.abs(vko).max() if use_grid else 0.0
            print("table %d kspace %d: force %.2e  pair energy %.2e  pair virial %.2e  k-space energy %.2e  virial %.2e"
                  % (table, use_grid, ferr, eerr, verr, kerr, kverr), flush=True)
            assert ferr <= 1e-9 and eerr <= 1e-10 and verr <= 1e-10 and kerr <= 1e-9 and kverr <= 1e-9
            worst = max(worst, ferr)
    print("INTEGRATION OK (max relative force error %.2e)" % worst)


if __name__ == "__main__":
    main()
