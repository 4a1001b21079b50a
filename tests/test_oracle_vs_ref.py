"""Pins the oracle: oracle/ (the restatement) against oracle/_ref/libref.so — the reference's OWN translation units
(pair_buck*_intel.cpp, pppm_intel.cpp, fix_nve_intel.cpp) compiled unchanged from /root/reference against the stand-in
headers of oracle/ref_shim/ (recipe oracle/Makefile.ref).  Same inputs to both; single-threaded results must agree BIT
FOR BIT, threaded ones to rounding of the thread reductions.  Also: the committed golden vectors (tests/golden/*.npz)
are reproduced by the reference's code itself, not only by the restatement.

Runs wherever libref.so exists (built here from /root/reference; it travels to the GPU box prebuilt)."""
import importlib.util
import os

import numpy as np
import pytest

import refc
import util

pytestmark = pytest.mark.skipif(not refc.available(), reason="oracle/_ref not built and /root/reference absent")

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def _listed(orc, P, prec, s, skin=0.3):
    """ghosts + binned half list (newton on) of the periodic system, as the reference's caller would hand them over"""
    cutneighmax = P.cutmax() + skin
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cutneighmax)
    nn, off, ent = orc.neigh_half_bin(len(s["x"]), xa, ta, P.ntypes, P.cutneighsq(skin), s["boxlo"], s["boxhi"],
                                      cutneighmax, prec)
    return xa, ta, qa, src, nn, off, ent


@pytest.mark.parametrize("prec", [0, 1], ids=["double", "mixed"])
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_pair_eval_bitwise(pkg, W, orc, name, prec):
    """eval<EVFLAG,EFLAG,NEWTON_PAIR> of all four styles (analytic and table branches): forces, per-atom energies,
    energies and both virial forms, one thread, bit for bit"""
    s, u, co, P, ct, dt = mg.pair_case(pkg, W, orc, name, prec=prec)
    n = len(s["x"])
    xa, ta, qa, src, nn, off, ent = _listed(orc, P, prec, s)
    for eflag, vflag, eatom in ((0, 0, 0), (1, 1, 0), (1, 1, 1), (1, 2, 0), (0, 1, 0)):
        fo, evo = orc.pair_eval(P, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=1, eatom=eatom, nthreads=1)
        fr, evr = refc.pair_eval(P, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=1, eatom=eatom, nthreads=1,
                                 skin=0.3)
        assert np.array_equal(fo, fr), (name, prec, eflag, vflag, np.abs(fo - fr).max())
        assert np.array_equal(evo, evr), (name, prec, eflag, vflag, evo, evr)


@pytest.mark.parametrize("name", ["buck", "buck_coul_long", "buck_long_coul_long"])
def test_pair_eval_newton_off_bitwise(pkg, W, orc, name):
    """the NEWTON_PAIR = 0 instantiation (f[j] and tallies only for owned j): the configuration the device path mirrors"""
    s, u, co, P, ct, dt = mg.pair_case(pkg, W, orc, name)
    n = len(s["x"])
    xa, ta, qa, src, nn, off, ent = _listed(orc, P, 0, s)
    fo, evo = orc.pair_eval(P, 0, 1, 1, n, xa, ta, qa, nn, off, ent, newton=0, eatom=1, nthreads=1)
    fr, evr = refc.pair_eval(P, 0, 1, 1, n, xa, ta, qa, nn, off, ent, newton=0, eatom=1, nthreads=1, skin=0.3)
    assert np.array_equal(fo[:n], fr[:n])
    assert np.array_equal(evo, evr)


def test_pair_special_bonds_bitwise(pkg, W, orc):
    """special-bond bits in the list entries scale Buckingham by special_lj and subtract (1 - special_coul) prefactor"""
    s, u, co, P0, ct, dt = mg.pair_case(pkg, W, orc, "buck_coul_long")
    P = orc.Params(orc.BUCK_COUL_LONG, s["ntypes"], co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"],
                   qqrd2e=u["qqrd2e"], g_ewald=0.2776, special_lj=(1, 0.0, 0.5, 0.25), special_coul=(1, 0.0, 0.3, 0.8))
    n = len(s["x"])
    xa, ta, qa, src, nn, off, ent = _listed(orc, P, 0, s)
    ent = ent.copy()
    rng = np.random.default_rng(5)
    pick = rng.random(len(ent)) < 0.05
    ent[pick] |= (rng.integers(1, 4, pick.sum()).astype(np.int32) << 30).astype(np.int32)
    fo, evo = orc.pair_eval(P, 0, 1, 1, n, xa, ta, qa, nn, off, ent, newton=1, eatom=1, nthreads=1)
    fr, evr = refc.pair_eval(P, 0, 1, 1, n, xa, ta, qa, nn, off, ent, newton=1, eatom=1, nthreads=1, skin=0.3)
    assert np.array_equal(fo, fr)
    assert np.array_equal(evo, evr)


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_reference_reproduces_golden_pair(pkg, W, orc, name):
    """the golden vectors the GPU tests compare against come out of the reference's own code (4 threads here: the
    thread-private force arrays are summed in thread order, so only rounding of that reduction differs)"""
    s, u, co, P, ct, dt = mg.pair_case(pkg, W, orc, name)
    n = len(s["x"])
    xa, ta, qa, src, nn, off, ent = _listed(orc, P, 0, s)
    fr, evr = refc.pair_eval(P, 0, 1, 1, n, xa, ta, qa, nn, off, ent, newton=1, eatom=1, nthreads=4, skin=0.3)
    fr = orc.reverse_comm(n, src, fr)[:n]
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert util.rel_force_err(fr[:, :3], g["f"]) <= 1e-13
    assert np.allclose(evr, g["ev"], rtol=1e-12, atol=1e-13 * np.abs(g["ev"]).max())
    assert np.allclose(fr[:, 3], g["eatom"], rtol=0, atol=1e-12 * np.abs(g["eatom"]).max())


@pytest.mark.parametrize("prec", [0, 1], ids=["double", "mixed"])
@pytest.mark.parametrize("name", ["pppm_ik5", "pppm_ad4", "pppm_ik7"])
def test_pppm_bitwise(W, orc, name, prec):
    """PPPMIntel::compute (particle_map, make_rho, brick2fft, poisson_ik/ad, fieldforce_ik/ad, energy / virial
    post-factors) on the oracle's base-class state: density, fields, forces, energy, virial — bit for bit"""
    grid, order, gew, ad, disp = mg.PPPM_CASES[name]
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    pp = orc.PPPM(*grid, order, gew, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad, prec=prec)
    fo, eo, vo = pp.compute(s["x"], s["q"], nthreads=1)
    dens_o = pp.density()
    fields_o = [pp.field(d) for d in range(1 if ad else 3)]
    fr, er, vr, dens_r, fields_r = refc.pppm_compute(pp, s["x"], s["q"], prec=prec, nthreads=1)
    assert np.array_equal(dens_o, dens_r), np.abs(dens_o - dens_r).max()
    for d in range(len(fields_o)):
        assert np.array_equal(fields_o[d], fields_r[d]), (d, np.abs(fields_o[d] - fields_r[d]).max())
    assert np.array_equal(fo, fr), np.abs(fo - fr).max()
    assert eo == er
    assert np.array_equal(vo, vr)


def test_pppm_threaded_and_golden(W, orc):
    """4 OpenMP threads in the reference's make_rho / fieldforce (thread-private grids summed in thread order), against
    the committed golden vector"""
    grid, order, gew, ad, disp = mg.PPPM_CASES["pppm_ik5"]
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    pp = orc.PPPM(*grid, order, gew, s["boxlo"], s["boxhi"], u["qqrd2e"])
    fr, er, vr, _, _ = refc.pppm_compute(pp, s["x"], s["q"], nthreads=4, want_grids=False)
    g = np.load(os.path.join(HERE, "golden", "pppm_ik5.npz"))
    assert util.rel_force_err(fr, g["f"]) <= 1e-12
    assert er == pytest.approx(float(g["e"]), rel=1e-12)
    assert np.allclose(vr, g["v"], rtol=0, atol=1e-12 * np.abs(g["v"]).max())


def test_pppm_out_of_range_message(W, orc):
    """error->one text of particle_map (pppm_intel.cpp:385) comes out of the reference itself"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    pp = orc.PPPM(24, 24, 27, 5, 0.28, s["boxlo"], s["boxhi"], u["qqrd2e"])
    x = s["x"].copy()
    x[0, 0] += 3.0 * (s["boxhi"][0] - s["boxlo"][0])
    with pytest.raises(RuntimeError, match="Out of range atoms - cannot compute PPPM"):
        refc.pppm_compute(pp, x, s["q"], want_grids=False)


def test_nve_bitwise(W, orc):
    """FixNVEIntel::initial_integrate / final_integrate: single type, several types, rmass, sub-group"""
    rng = np.random.default_rng(11)
    n = 500
    x, v, f = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 3)) * 3
    dt, ftm2v = 0.001, 1.0 / 1.0364269e-4
    # one type: scalar dtfm branch (fix_nve_intel.cpp:68-77)
    t1, m1 = np.ones(n, np.int32), np.array([0.0, 12.011])
    d1 = orc.nve_dtfm(t1, m1, dt, ftm2v)
    for which in (0, 1):
        xr, vr = refc.nve(which, x, v, f, t1, m1, dt, ftm2v)
        if which == 0:
            xo, vo = orc.nve_initial(x, v, f, d1, dt)
        else:
            xo, vo = x, orc.nve_final(v, f, d1)
        assert np.array_equal(xo, xr) and np.array_equal(vo, vr)
    # two types: _dtfm array branch
    t2, m2 = rng.integers(1, 3, n).astype(np.int32), np.array([0.0, 12.011, 15.9994])
    d2 = orc.nve_dtfm(t2, m2, dt, ftm2v)
    xr, vr = refc.nve(0, x, v, f, t2, m2, dt, ftm2v)
    xo, vo = orc.nve_initial(x, v, f, d2, dt)
    assert np.array_equal(xo, xr) and np.array_equal(vo, vr)
    xr, vr = refc.nve(1, x, v, f, t2, m2, dt, ftm2v)
    assert np.array_equal(orc.nve_final(v, f, d2), vr)
    # per-atom masses and a sub-group (:88-97, :147-190)
    rmass = rng.uniform(1.0, 20.0, n)
    ingroup = (rng.random(n) < 0.6).astype(np.int32)
    dg = orc.nve_dtfm_group(t2, m2, dt, ftm2v, ingroup=ingroup, rmass=rmass)
    xr, vr = refc.nve(0, x, v, f, t2, m2, dt, ftm2v, rmass=rmass, ingroup=ingroup)
    xo, vo = orc.nve_initial_group(x, v, f, dg, dt)
    assert np.array_equal(xo, xr) and np.array_equal(vo, vr)
    assert np.array_equal(xr[ingroup == 0], x[ingroup == 0])


@pytest.mark.parametrize("prec", [0, 1], ids=["double", "mixed"])
@pytest.mark.parametrize("variant", ["cut_long", "cut_long_table", "long_long", "long_long_table"])
def test_lj_long_coul_long_bitwise(pkg, W, orc, variant, prec):
    """pair_lj_long_coul_long_intel.cpp on the real data.spce water box: `lj/long/coul/long cut long 6.8 8.8`
    (= the lj/cut/coul/long of examples/in.spce:7) and `long long`, analytic and tabulated, with the special-bond bits of
    the water molecules (special_bonds lj/coul 0.0 0.0 0.5) in the list — forces, per-atom energies, energies, virial"""
    s = W.spce_system(1)
    u = W.UNITS["real"]
    co = W.coeffs_spce()
    o6 = 1 if variant.startswith("long_long") else 0
    ge, g6 = 0.32, 0.30 if o6 else 0.0
    P = orc.Params(orc.LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"],
                   qqrd2e=u["qqrd2e"], g_ewald=ge, g_ewald_6=g6, order1=1, order6=o6, special_lj=(1, 0.0, 0.0, 0.5),
                   special_coul=(1, 0.0, 0.0, 0.5))
    if variant.endswith("table"):
        ct = pkg.init_coul_tables(8.8, ge, u["qqrd2e"])
        P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
        # the reference copies the dispersion tables whenever the Coulomb ones are on (:850-866)
        dt = pkg.init_disp_tables(6.8, g6 if o6 else 0.3)
        P.set_disp_tables(dt[0], 12 if o6 else 0, dt[1], dt[2], dt[3])
    n = len(s["x"])
    skin = 2.0
    cutneighmax = P.cutmax() + skin
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cutneighmax)
    nn, off, ent = orc.neigh_half_bin(n, xa, ta, 2, P.cutneighsq(skin), s["boxlo"], s["boxhi"], cutneighmax, prec)
    ent = util.water_special_bits(n, nn, ent, src, s["mol"], s["type"])
    assert ((ent.view(np.uint32) >> 30) != 0).sum() == 3 * (n // 3)      # two O-H and one H-H pair per molecule
    for eflag, vflag, eatom in ((0, 0, 0), (1, 1, 1), (1, 2, 0)):
        fo, evo = orc.pair_eval(P, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=1, eatom=eatom, nthreads=1)
        fr, evr = refc.pair_eval(P, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=1, eatom=eatom, nthreads=1,
                                 skin=skin)
        assert np.array_equal(fo, fr), (variant, prec, eflag, vflag, np.abs(fo - fr).max())
        assert np.array_equal(evo, evr), (variant, prec, evo, evr)
    assert np.abs(fo[:n, :3]).max() > 1.0


@pytest.mark.parametrize("prec", [0, 1], ids=["double", "mixed"])
@pytest.mark.parametrize("table", [0, 12], ids=["analytic", "table"])
def test_lj_long_off_on_hexane_bitwise(pkg, W, orc, table, prec):
    """the instantiation examples/in.hexane selects — `lj/long/coul/long long off 9.8`: eval<..., ORDER1 = 0, ORDER6 = 1>
    (pair_lj_long_coul_long_intel.cpp:426-747), analytic and with the dispersion table of the stock default
    `table/disp 12` — on the real equilibrated_data.hexane (no charges, no special bonds), g_ewald_6 as PPPMDisp sizes it
    for the script's `force/disp/real 0.0001`"""
    s = W.hexane_system()
    co = W.coeffs_hexane()
    g6 = 0.3044751226
    P = orc.Params(orc.LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], g_ewald_6=g6,
                   order1=0, order6=1)
    if table:
        # pack_force_const copies every table — the dispersion ones too — only `if (ncoultablebits)` (:839-860), from the
        # Coulomb arrays' length: as shipped (both defaults 12, Coulomb off, so stock init_style never builds rtable)
        # the reference reads a null rtable.  The harness hands it Coulomb tables as well; ORDER1 = 0 never reads them
        ct = pkg.init_coul_tables(9.8, 0.3, 1.0)
        P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
        dt = pkg.init_disp_tables(9.8, g6)
        P.set_disp_tables(dt[0], 12, dt[1], dt[2], dt[3])
    n = len(s["x"])
    skin = 2.0
    cutneighmax = P.cutmax() + skin
    assert abs(cutneighmax - 11.8) < 1e-12
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cutneighmax)
    nn, off, ent = orc.neigh_half_bin(n, xa, ta, 2, P.cutneighsq(skin), s["boxlo"], s["boxhi"], cutneighmax, prec)
    for eflag, vflag, eatom in ((0, 0, 0), (1, 1, 1), (1, 2, 0)):
        fo, evo = orc.pair_eval(P, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=1, eatom=eatom, nthreads=1)
        fr, evr = refc.pair_eval(P, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=1, eatom=eatom, nthreads=1,
                                 skin=skin)
        assert np.array_equal(fo, fr), (table, prec, eflag, vflag, np.abs(fo - fr).max())
        assert np.array_equal(evo, evr), (table, prec, evo, evr)
    assert np.abs(fo[:n, :3]).max() > 1.0 and evo[1] == 0.0


@pytest.mark.parametrize("prec", [0, 1], ids=["double", "mixed"])
def test_whole_step_from_the_reference_units_equals_the_oracle_loop(pkg, W, orc, prec):
    """tests/refmd.py (what `bench.py --impl reference` times): the Verlet step with PairBuckCoulLongIntel::compute,
    PPPMIntel::compute and FixNVEIntel::initial/final_integrate running from the reference's own compiled units and the
    upstream pieces from the oracle — the same trajectory, bit for bit at one thread, as oracle/md.cpp over steps that
    include list rebuilds and ghost refreshes"""
    import refmd
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    co = W.coeffs_aC(8.0, 8.0)
    g = 0.30
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=g)
    kw = dict(prec=prec, skin=0.3, every=1, delay=0, check=1, dt=0.001, ftm2v=u["ftm2v"])   # 4 builds in 20 steps
    pp = orc.PPPM(24, 24, 27, 5, g, s["boxlo"], s["boxhi"], u["qqrd2e"], prec=prec)
    md = orc.MD(s, P, pppm=pp, **kw)
    tm = md.run(20, 1)
    pp2 = orc.PPPM(24, 24, 27, 5, g, s["boxlo"], s["boxhi"], u["qqrd2e"], prec=prec)
    rm = refmd.RefMD(s, P, pppm=pp2, nthreads=1, **kw)
    tr = rm.run(20)
    xo, vo, fo = md.get()
    assert tr["nbuilds"] == tm["nbuilds"] and 3 <= tr["nbuilds"] < 10
    assert np.array_equal(rm.x, xo) and np.array_equal(rm.v, vo) and np.array_equal(rm.f, fo)
    assert tr["ref_pair"] > 0.0 and tr["ref_kspace"] > 0.0 and tr["ref_nve"] > 0.0 and rm.seconds() > 0.0
