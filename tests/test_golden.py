"""Golden vectors (tests/golden/*.npz, frozen by tests/golden/make_golden.py from the KAT-pinned oracle):
CPU — the oracle still reproduces them (guards the checker itself against compiler / refactoring drift);
GPU — the device path, through the C ABI, matches them at the parity bars (1e-9 forces, 1e-10 energy/virial)."""
import importlib.util
import os

import numpy as np
import pytest

import util

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_oracle_reproduces_golden_pair(pkg, W, orc, name):
    s, u, co, P, ct, dt = mg.pair_case(pkg, W, orc, name)
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3, eflag=3, vflag=1,
                                        eatom=1)
    g = _load(name)
    assert util.rel_force_err(f[:, :3], g["f"]) <= 1e-13
    assert np.allclose(ev, g["ev"], rtol=1e-13, atol=1e-13 * np.abs(g["ev"]).max())
    assert np.allclose(f[:, 3], g["eatom"], rtol=0, atol=1e-12 * np.abs(g["eatom"]).max())


@pytest.mark.parametrize("name", sorted(mg.PPPM_CASES))
def test_oracle_reproduces_golden_pppm(W, orc, name):
    grid, order, gew, ad, disp = mg.PPPM_CASES[name]
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    if disp:
        f, e, v = orc.PPPM.dispersion(*grid, order, gew, s["boxlo"], s["boxhi"]).compute(s["x"], mg.B_DISP[s["type"]])
    else:
        f, e, v = orc.PPPM(*grid, order, gew, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad).compute(s["x"], s["q"])
    g = _load(name)
    assert util.rel_force_err(f, g["f"]) <= 1e-12
    assert e == pytest.approx(float(g["e"]), rel=1e-12)
    assert np.allclose(v, g["v"], rtol=0, atol=1e-12 * np.abs(g["v"]).max())


@pytest.mark.parametrize("name", sorted(mg.PPPM_CASES2))
def test_oracle_reproduces_golden_pppm_variants(W, orc, name):
    """per-atom tallies, the dispersion grid with ad differentiation, the slab correction"""
    s, f, e, v, ea, va = mg.pppm_case2(W, orc, name)
    g = _load(name)
    assert util.rel_force_err(f, g["f"]) <= 1e-12
    assert e == pytest.approx(float(g["e"]), rel=1e-12)
    assert np.allclose(v, g["v"], rtol=0, atol=1e-12 * np.abs(g["v"]).max())
    if ea is not None:
        assert np.allclose(ea, g["eatom"], rtol=0, atol=1e-12 * np.abs(g["eatom"]).max())
        assert np.allclose(va, g["vatom"], rtol=0, atol=1e-12 * np.abs(g["vatom"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.PPPM_CASES2))
def test_device_matches_golden_pppm_variants(pkg, W, name):
    grid, order, gew, ad, disp, slab, peratom = mg.PPPM_CASES2[name]
    s = W.aC_system(1)
    if slab > 1.0:
        s = dict(s, periodic=(1, 1, 0))
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3)
    if disp:
        ctx.pppm_setup(*grid, order, gew, dispersion=1, B=mg.B_DISP, differentiation=ad)
    else:
        ctx.pppm_setup(*grid, order, gew, differentiation=ad, slab=slab if slab > 1.0 else 0.0)
    e, v = ctx.pppm_compute(3 if peratom else 1, 5 if peratom else 1)
    f = ctx.atoms_download(("f",))["f"]
    g = _load(name)
    assert util.rel_force_err(f, g["f"]) <= 1e-9
    assert abs(e - float(g["e"])) <= 1e-10 * abs(float(g["e"]))
    assert np.abs(v - g["v"]).max() <= 1e-10 * np.abs(g["v"]).max()
    if peratom:
        ea, va = ctx.pppm_peratom()
        assert np.abs(ea - g["eatom"]).max() <= 1e-10 * np.abs(g["eatom"]).max()
        assert np.abs(va - g["vatom"]).max() <= 1e-10 * np.abs(g["vatom"]).max()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_device_matches_golden_pair(pkg, W, orc, name):
    sysn, style, cl, cc, ge, g6, o1, o6, tables = mg.CASES[name]
    s, u, co, P, ct, dt = mg.pair_case(pkg, W, orc, name)
    pstyle = getattr(pkg, "PAIR_" + style)
    cf = pkg.pair_coeffs(pstyle, s["ntypes"], co["A"], co["rho"], co["C"], co["cut_lj"], co.get("cut_coul"))
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3)
    ctx.pair_setup(pstyle, s["ntypes"], cf, g_ewald=ge, g_ewald_6=g6, ewald_order=(o1 << 1) | (o6 << 6), coul_tables=ct,
                   disp_tables=dt)
    ctx.neigh_build()
    ev = ctx.pair_compute(3, 1)
    d = ctx.atoms_download(("f", "eatom"))
    g = _load(name)
    assert util.rel_force_err(d["f"], g["f"]) <= 1e-9
    escale = np.abs(g["ev"][:2]).max()
    assert np.abs(ev[:2] - g["ev"][:2]).max() <= 1e-10 * escale
    assert np.abs(ev[2:] - g["ev"][2:]).max() <= 1e-10 * np.abs(g["ev"][2:]).max()
    assert np.abs(d["eatom"] - g["eatom"]).max() <= 1e-9 * np.abs(g["eatom"]).max()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.PPPM_CASES))
def test_device_matches_golden_pppm(pkg, W, name):
    grid, order, gew, ad, disp = mg.PPPM_CASES[name]
    s = W.aC_system(1)
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3)
    if disp:
        ctx.pppm_setup(*grid, order, gew, dispersion=1, B=mg.B_DISP)
    else:
        ctx.pppm_setup(*grid, order, gew, differentiation=ad)
    e, v = ctx.pppm_compute(1, 1)       # a fresh upload left f = 0: the accumulated force is the k-space force
    f = ctx.atoms_download(("f",))["f"]
    g = _load(name)
    assert util.rel_force_err(f, g["f"]) <= 1e-9
    assert abs(e - float(g["e"])) <= 1e-10 * abs(float(g["e"]))
    assert np.abs(v - g["v"]).max() <= 1e-10 * np.abs(g["v"]).max()
    ctx.close()
