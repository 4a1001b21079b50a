"""-m gpu parity at the sizes the reference's scripts really run, through the C ABI:
  in.buck 32 000 atoms, in.buck_coul_cut 76 800, in.buck_big 192 000 — forces, energies, virial, pair set;
  in.spce electrostatics on the REAL examples/data.spce (tests/golden/data_spce.npz): x 2^3 at 72^3, x 4^3 at 80^3 and 135^3;
  one neighbour list with more than 2^31 entries, checked through a size-independent property (every replica of the
  perfect data.aC crystal feels the forces of the 2^3 system the oracle computes).
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _signature(n, i, own):
    """order-independent per-atom signature of a pair set: count, sum and sum of squares of the partner ids"""
    w = own.astype(np.float64)
    return (np.bincount(i, minlength=n), np.bincount(i, weights=w, minlength=n),
            np.bincount(i, weights=(w % 4093.0) ** 2, minlength=n))


def _owners(j, n, src):
    own = j.astype(np.int64).copy()
    g = own >= n
    for _ in range(4):
        if not g.any():
            break
        own[g] = src[own[g] - n]
        g = own >= n
    return own


@pytest.mark.parametrize("script", ["in.buck", "in.buck_coul_cut", "in.buck_big"])
def test_script_size_parity(pkg, W, orc, script):
    if script == "in.buck":                 # examples/in.buck:7-30: 20^3 fcc cells, buck 2.5
        s, co, style, ostyle, nt, natoms = W.fcc_system(20, 20, 20), W.coeffs_in_buck(2.5), pkg.PAIR_BUCK, orc.BUCK, 1, 32000
    elif script == "in.buck_big":           # examples/in.buck_big:4-13: 30 x 40 x 40 cells, buck 5.0
        s, co, style, ostyle, nt, natoms = W.fcc_system(30, 40, 40), W.coeffs_in_buck(5.0), pkg.PAIR_BUCK, orc.BUCK, 1, 192000
    else:                                   # examples/in.buck_coul_cut:2-11: data.aC x 4^3, buck/coul/cut 10.0
        s, co, style, ostyle, nt, natoms = (W.aC_system(4), W.coeffs_aC(10.0, 10.0), pkg.PAIR_BUCK_COUL_CUT,
                                            orc.BUCK_COUL_CUT, 2, 76800)
    n = len(s["x"])
    assert n == natoms
    u = W.UNITS[s["units"]]
    cutc = co.get("cut_coul")
    P = orc.Params(ostyle, nt, co["A"], co["rho"], co["C"], co["cut_lj"], cutc, qqrd2e=u["qqrd2e"])
    cf = pkg.pair_coeffs(style, nt, co["A"], co["rho"], co["C"], co["cut_lj"], cutc)
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3)
    ctx.pair_setup(style, nt, cf)
    ctx.neigh_build()
    ev = ctx.pair_compute(1, 1)
    f = ctx.atoms_download(("f",))["f"]
    fo, evo, aux = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    assert util.rel_force_err(f, fo[:, :3]) <= 1e-9
    escale = max(abs(evo[0]), abs(evo[1]))
    assert abs(ev[0] - evo[0]) <= 1e-10 * escale and abs(ev[1] - evo[1]) <= 1e-10 * escale
    assert np.abs(ev[2:8] - evo[2:8]).max() <= 1e-10 * np.abs(evo[2:8]).max()
    # pair set: device full list against the symmetrised binned half list of the oracle
    nn, off, ent, gsrc, gshift = ctx.neigh_download()
    i_g = np.repeat(np.arange(n, dtype=np.int64), nn)
    own_g = _owners(ent & 0x3FFFFFFF, n, gsrc)
    i_h = np.repeat(np.arange(n, dtype=np.int64), aux["numneigh"])
    own_h = _owners(aux["entries"] & 0x3FFFFFFF, n, aux["src"])
    assert 2 * len(own_h) == len(own_g) == ctx.neigh_stats()["total"]
    sg = _signature(n, i_g, own_g)
    sh = _signature(n, np.concatenate([i_h, own_h]), np.concatenate([own_h, i_h]))
    for a, b in zip(sg, sh):
        assert np.array_equal(a, b)
    ctx.close()


@pytest.mark.parametrize("rep,grid,acc", [(2, (72, 72, 72), 1e-5), (4, (80, 80, 80), 1e-4), (4, (135, 135, 135), 1e-5)])
def test_data_spce_pppm_parity(pkg, W, orc, rep, grid, acc):
    """BASELINE config 4 (examples/in.spce:6-11): PPPM on the real data.spce, replicated, at the grids stock
    set_grid_global picks for pppm 1e-4 / 1e-5 with the script's 8.8 A Coulomb cut-off"""
    s = W.spce_system(rep)
    u = W.UNITS["real"]
    n = len(s["x"])
    assert n == 4500 * rep ** 3
    g_auto, g = pkg.pppm_init(acc, u["qqrd2e"], s["q"], n, 8.8, s["boxhi"] - s["boxlo"])
    if rep == 4:
        assert tuple(g_auto) == grid          # the sizing itself reproduces SURVEY 6.2
    ctx = pkg.make_context(s)
    ctx.neigh_setup(2.0)
    ctx.pppm_setup(*grid, 5, g)
    pp = orc.PPPM(*grid, 5, g, s["boxlo"], s["boxhi"], u["qqrd2e"])
    fo, eo, vo = pp.compute(s["x"], s["q"])
    f, e, v = ctx.pppm_compute_host(s["x"], s["q"], 1, 1)
    assert util.rel_force_err(f, fo) <= 1e-9
    assert abs(e - eo) <= 1e-10 * abs(eo)
    assert np.abs(v - vo).max() <= 1e-10 * np.abs(vo).max()
    d = ctx.pppm_download()
    rho = pp.density()
    assert np.abs(d["density"] - rho).max() <= 1e-11 * np.abs(rho).max()
    ctx.close()


def test_list_beyond_2_31_entries(pkg, W, orc):
    """data.aC x 16^3 = 4.9 M atoms: 2.55 G list entries (> INT_MAX; 10 GB).  The perfect crystal is periodic in the
    1 200-atom cell, so atom k of every replica must feel the force the oracle computes for atom k of the 2^3 system."""
    rep = 16
    co = W.coeffs_aC(12.0, 12.0)
    u = W.UNITS["metal"]
    g = 0.28
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    s = W.aC_system(rep, jitter=0.0)
    n = len(s["x"])
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3)
    ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=g)
    ctx.neigh_build()
    st = ctx.neigh_stats()
    assert st["total"] > 2 ** 31, st
    ev = ctx.pair_compute(1, 1)
    f = ctx.atoms_download(("f",))["f"]
    ctx.close()
    s2 = W.aC_system(2, jitter=0.0)
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=g)
    fo, evo, _ = orc.pair_forces_periodic(P, 0, s2["x"], s2["type"], s2["q"], s2["boxlo"], s2["boxhi"], 0.3)
    base = fo[:1200, :3]                       # replicate keeps the cell's 1 200 atoms inner-most
    scale = np.abs(base).max()
    err = np.abs(f.reshape(-1, 1200, 3) - base[None]).max() / scale
    assert err <= 1e-9, err
    per_atom = (evo[0] + evo[1]) / len(s2["x"])
    assert abs((ev[0] + ev[1]) / n - per_atom) <= 1e-10 * abs(per_atom)
