"""-m gpu: the non-bonded + k-space force of examples/in.spce (BASELINE config 4) on the REAL data.spce:
`lj/long/coul/long cut long 6.8 8.8` (= the script's lj/cut/coul/long 6.8 8.8; pair_lj_long_coul_long_intel.cpp:426-747),
special_bonds lj/coul 0.0 0.0 0.5 carried as bits 30-31 of a list built ON THE DEVICE (b200md_atoms_set_special), pppm 1e-4 —
against the oracle (bit-identical to the reference's compiled loops, tests/test_oracle_vs_ref.py)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def water_specials(n):
    """O, H, H per molecule: O has two 1-2 partners; each H has the O as 1-2 and the other H as 1-3 partner"""
    nspecial = np.zeros((n, 3), np.int32)
    special = np.zeros((n, 2), np.int32)
    o = np.arange(0, n, 3)
    nspecial[o] = (2, 2, 2)
    special[o, 0], special[o, 1] = o + 1, o + 2
    for h, other in ((o + 1, o + 2), (o + 2, o + 1)):
        nspecial[h] = (1, 2, 2)
        special[h, 0], special[h, 1] = o, other
    return nspecial, special


@pytest.mark.parametrize("prec", [0, 1], ids=["double", "mixed"])
@pytest.mark.parametrize("table", [False, True], ids=["analytic", "table"])
def test_in_spce_nonbonded_and_kspace_forces(pkg, W, orc, table, prec):
    s = W.spce_system(1)
    n = len(s["x"])
    assert n == 4500 and np.array_equal(s["type"][:3], [1, 2, 2]) and len(set(s["mol"][:3])) == 1
    u = W.UNITS["real"]
    co = W.coeffs_spce()
    skin = 2.0
    grid, ge = pkg.pppm_init(1e-4, u["qqrd2e"], s["q"], n, 8.8, s["boxhi"] - s["boxlo"])
    sl, sc = (1, 0.0, 0.0, 0.5), (1, 0.0, 0.0, 0.5)
    P = orc.Params(orc.LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=ge, order1=1, special_lj=sl, special_coul=sc)
    cf = pkg.pair_coeffs(pkg.PAIR_LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    ct = None
    if table:
        ct = pkg.init_coul_tables(8.8, ge, u["qqrd2e"])
        P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
        dt = pkg.init_disp_tables(6.8, 0.3)       # copied by the reference whenever the Coulomb tables are on; unused
        P.set_disp_tables(dt[0], 0, dt[1], dt[2], dt[3])
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(skin, every=1, delay=10, check=1)
    ctx.pair_setup(pkg.PAIR_LJ_LONG_COUL_LONG, 2, cf, special_lj=sl, special_coul=sc, g_ewald=ge, ewald_order=1 << 1,
                   coul_tables=ct)
    ctx.atoms_set_special(*water_specials(n))
    ctx.pppm_setup(*grid, 5, ge)
    ctx.nve_setup(u["dt"])
    th = ctx.setup_forces(1, 1)
    f = ctx.atoms_download(("f",))["f"]
    # the list the device built: six flagged entries per molecule (O-H twice each way, H-H each way)
    nn, off, ent, gsrc, gshift = ctx.neigh_download()
    bits = (ent.view(np.uint32) >> 30)
    assert (bits == 1).sum() == 4 * (n // 3) and (bits == 2).sum() == 2 * (n // 3) and (bits == 3).sum() == 0
    # oracle: binned half list with the same bits, reference-shaped evaluation, reverse comm, PPPM
    cutneighmax = P.cutmax() + skin
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cutneighmax)
    hn, hoff, hent = orc.neigh_half_bin(n, xa, ta, 2, P.cutneighsq(skin), s["boxlo"], s["boxhi"], cutneighmax, prec)
    hent = util.water_special_bits(n, hn, hent, src, s["mol"], s["type"])
    fo, evo = orc.pair_eval(P, prec, 1, 1, n, xa, ta, qa, hn, hoff, hent, newton=1)
    fo = orc.reverse_comm(n, src, fo)[:n]
    pp = orc.PPPM(*grid, 5, ge, s["boxlo"], s["boxhi"], u["qqrd2e"], prec=prec)
    fk, ek, vk = pp.compute(s["x"], s["q"])
    ftot = fo[:, :3] + fk
    tol_f, tol_e = (1e-9, 1e-10) if prec == 0 else (5e-5, 1e-5)   # mixed: newton-on arm, see test_gpu_pair.py
    assert util.rel_force_err(f, ftot) <= tol_f
    escale = max(abs(evo[0]), abs(evo[1]))
    assert abs(th[0] - evo[0]) <= tol_e * escale and abs(th[1] - evo[1]) <= tol_e * escale
    assert np.abs(th[2:8] - evo[2:8]).max() <= tol_e * np.abs(evo[2:8]).max()
    assert abs(th[8] - ek) <= max(tol_e, 1e-9) * abs(ek)
    # the same pairs through the host-list entry point (type gathered, bits supplied by the caller)
    if prec == 0:
        fn_, foff_, fent_ = orc.neigh_full_brute(n, xa, ta, 2, P.cutneighsq(skin), prec)
        fent_ = util.water_special_bits(n, fn_, fent_, src, s["mol"], s["type"])
        fh, evh = ctx.pair_eval_host(1, 1, n, xa, ta, qa, fn_, foff_[:-1], fent_)
        assert util.rel_force_err(fh[:, :3] + fk, ftot) <= 1e-9
    ctx.run(3)    # a few steps with the bits re-made at rebuilds
    ctx.close()


@pytest.mark.parametrize("prec", [0, 1], ids=["double", "mixed"])
@pytest.mark.parametrize("table", [False, True], ids=["analytic", "table"])
def test_in_hexane_pair_and_dispersion_kspace_forces(pkg, W, orc, table, prec):
    """examples/in.hexane on the device: `lj/long/coul/long long off 9.8` (ORDER1 = 0, ORDER6 = 1: the instantiation the
    oracle equals the reference on bit for bit, test_lj_long_off_on_hexane_bitwise) + the geometric dispersion grid of
    `pppm/disp`, g_ewald_6 and mesh as the host class sizes them from the script's accuracies (test_host_sizing.py), on
    the real equilibrated_data.hexane — forces, energies, virial against the oracle"""
    s = W.hexane_system()
    n = len(s["x"])
    co = W.coeffs_hexane()
    skin, g6, grid6 = 2.0, 0.3044751226, (50, 24, 20)
    P = orc.Params(orc.LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], g_ewald_6=g6,
                   order1=0, order6=1)
    cf = pkg.pair_coeffs(pkg.PAIR_LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    dt = None
    if table:
        dt = pkg.init_disp_tables(9.8, g6)
        P.set_disp_tables(dt[0], 12, dt[1], dt[2], dt[3])
    B = np.sqrt(4.0 * np.array([0.0, co["A"][1, 1], co["A"][2, 2]]) * 3.97 ** 6)
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(skin, every=1, delay=10, check=1)
    ctx.pair_setup(pkg.PAIR_LJ_LONG_COUL_LONG, 2, cf, g_ewald_6=g6, ewald_order=1 << 6, disp_tables=dt)
    ctx.pppm_setup(*grid6, 5, g6, dispersion=1, B=B)
    ctx.nve_setup(1.0e-4)   # fix nve instead of the script's rigid bodies: see scripts.IN_HEXANE_NVE
    th = ctx.setup_forces(1, 1)
    f = ctx.atoms_download(("f",))["f"]
    cutneighmax = P.cutmax() + skin
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cutneighmax)
    hn, hoff, hent = orc.neigh_half_bin(n, xa, ta, 2, P.cutneighsq(skin), s["boxlo"], s["boxhi"], cutneighmax, prec)
    fo, evo = orc.pair_eval(P, prec, 1, 1, n, xa, ta, qa, hn, hoff, hent, newton=1)
    fo = orc.reverse_comm(n, src, fo)[:n]
    fk, ek, vk = orc.PPPM.dispersion(*grid6, 5, g6, s["boxlo"], s["boxhi"], prec=prec).compute(s["x"], B[s["type"]])
    ftot = fo[:, :3] + fk
    tol_f, tol_e = (1e-9, 1e-10) if prec == 0 else (5e-5, 1e-5)   # mixed: newton-on arm, see test_gpu_pair.py
    assert util.rel_force_err(f, ftot) <= tol_f
    assert abs(th[0] - evo[0]) <= tol_e * abs(evo[0]) and th[1] == 0.0 == evo[1]
    assert np.abs(th[2:8] - evo[2:8]).max() <= tol_e * np.abs(evo[2:8]).max()
    assert abs(th[8] - ek) <= max(tol_e, 1e-9) * abs(ek)
    assert np.abs(th[9:15] - vk).max() <= max(tol_e, 1e-9) * np.abs(vk).max()
    ctx.run(3)
    # the k-space forces alone (the overlapping united atoms of a molecule dominate the total: 4e5 against 0.3): a fresh
    # upload zeroes the force array, the dispersion grid accumulates onto it
    ctx.atoms_upload(s["x"], s["type"], s["mass"], v=s["v"], q=s["q"])
    e2, v2 = ctx.pppm_compute(1, 1)
    f2 = ctx.atoms_download(("f",))["f"]
    assert np.abs(f2 - fk).max() <= (1e-9 if prec == 0 else 2e-5) * np.abs(fk).max()
    assert abs(e2 - ek) <= max(tol_e, 1e-9) * abs(ek)
    ctx.close()
