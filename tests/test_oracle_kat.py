"""Known-answer tests that pin the CPU oracle (the reference ships no tests or golden vectors, SURVEY.md §4,
§8c: "parity unpinned").  Each KAT checks a restated formula against an independent closed form / library."""
import math
import os

import numpy as np
import pytest

import util


def _dimer(orc, style, r, qq=(1.3, -0.7), **kw):
    """two atoms at distance r along a skewed axis, no periodic images: returns (f on atom 0, ev)"""
    d = np.array([0.36, -0.48, 0.8])
    d /= np.linalg.norm(d)
    x = np.array([[0.1, 0.2, 0.3], [0.1, 0.2, 0.3] + r * d])
    t = np.array([1, 2], np.int32)
    q = np.array(qq)
    P = kw["P"]
    nn = np.array([1, 0], np.int32)
    off = np.array([0, 1, 1], np.int64)
    ent = np.array([1], np.int32)
    f, ev = orc.pair_eval(P, kw.get("prec", 0), 1, 1, 2, x, t, q, nn, off, ent, newton=1)
    return f, ev, d


def _params(orc, style, **kw):
    A = np.zeros((3, 3)); rho = np.ones((3, 3)); C = np.zeros((3, 3))
    A[1, 2] = A[2, 1] = 1800.0; rho[1, 2] = rho[2, 1] = 0.29; C[1, 2] = C[2, 1] = 130.0
    A[1, 1] = A[2, 2] = 10.0; C[1, 1] = C[2, 2] = 1.0
    return orc.Params(style, 2, A, rho, C, np.full((3, 3), 9.0), np.full((3, 3), 10.0), qqrd2e=14.399645, **kw)


@pytest.mark.parametrize("r", [1.1, 2.5, 6.0, 8.9])
def test_dimer_closed_form_all_styles(orc, r):
    A, rho, C, qq, k = 1800.0, 0.29, 130.0, 1.3 * -0.7, 14.399645
    e_b = A * math.exp(-r / rho) - C / r ** 6
    f_b = A / rho * math.exp(-r / rho) - 6 * C / r ** 7          # -dE/dr
    # buck
    f, ev, d = _dimer(orc, orc.BUCK, r, P=_params(orc, orc.BUCK))
    assert ev[0] == pytest.approx(e_b, rel=1e-13)
    assert np.allclose(f[0, :3], -f_b * d, rtol=1e-12) and np.allclose(f[1, :3], f_b * d, rtol=1e-12)
    # buck/coul/cut
    f, ev, d = _dimer(orc, orc.BUCK_COUL_CUT, r, P=_params(orc, orc.BUCK_COUL_CUT))
    assert ev[1] == pytest.approx(k * qq / r, rel=1e-13)
    assert np.allclose(f[0, :3], -(f_b + k * qq / r ** 2) * d, rtol=1e-12)
    # buck/coul/long analytic: erfc via the A&S 5-term polynomial (<= 1.5e-7 absolute)
    g = 0.28
    f, ev, d = _dimer(orc, orc.BUCK_COUL_LONG, r, P=_params(orc, orc.BUCK_COUL_LONG, g_ewald=g))
    assert ev[1] == pytest.approx(k * qq * math.erfc(g * r) / r, abs=abs(k * qq / r) * 2e-7)
    fc = k * qq * (math.erfc(g * r) / r ** 2 + 2 * g / math.sqrt(math.pi) * math.exp(-(g * r) ** 2) / r)
    assert np.allclose(f[0, :3], -(f_b + fc) * d, rtol=0, atol=abs(k * qq / r ** 2) * 3e-7 + 1e-12)
    # buck/long/coul/long, ORDER6 only: E = A e^{-r/rho} - C g6^6 e^{-x} (1 + x + x^2/2) / x^3, x = (g6 r)^2
    g6 = 0.31
    P = _params(orc, orc.BUCK_LONG_COUL_LONG, g_ewald_6=g6, order6=1)
    f, ev, d = _dimer(orc, orc.BUCK_LONG_COUL_LONG, r, P=P)
    xx = (g6 * r) ** 2
    e6 = A * math.exp(-r / rho) - C * g6 ** 6 * math.exp(-xx) * (1 + xx + xx * xx / 2) / xx ** 3
    assert ev[0] == pytest.approx(e6, rel=1e-12)


@pytest.mark.parametrize("style,kw", [(0, {}), (1, {}), (2, dict(g_ewald=0.28)),
                                       (3, dict(g_ewald=0.28, g_ewald_6=0.31, order1=1, order6=1)),
                                       (3, dict(g_ewald_6=0.31, order6=1))])
def test_force_is_minus_energy_gradient(orc, style, kw):
    P = _params(orc, style, **kw)
    for r in (1.3, 3.0, 7.7):
        h = 1e-5
        _, evp, _ = _dimer(orc, style, r + h, P=P)
        _, evm, _ = _dimer(orc, style, r - h, P=P)
        f, _, d = _dimer(orc, style, r, P=P)
        dEdr = ((evp[0] + evp[1]) - (evm[0] + evm[1])) / (2 * h)
        fr = float(f[1, :3] @ d)              # force on atom 1 along +d = -dE/dr
        # the polynomial erfc has a 1e-7-level derivative inconsistency by construction
        assert fr == pytest.approx(-dEdr, rel=2e-6, abs=1e-6)


def test_as_erfc_polynomial(orc):
    """pair_buck_coul_long_intel.cpp:296-307 constants reproduce erfc to 1.5e-7 absolute"""
    A = (0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429)
    p = 0.3275911
    x = np.linspace(0.0, 6.0, 2001)
    t = 1.0 / (1.0 + p * x)
    poly = t * (A[0] + t * (A[1] + t * (A[2] + t * (A[3] + t * A[4])))) * np.exp(-x * x)
    ref = np.array([math.erfc(v) for v in x])
    assert np.abs(poly - ref).max() <= 1.5e-7


def test_special_bonds_and_tables_dimer(orc, pkg):
    g = 0.28
    P = _params(orc, orc.BUCK_COUL_LONG, g_ewald=g, special_lj=(1, 0.0, 0.5, 0.25), special_coul=(1, 0.0, 0.3, 0.8))
    r = 3.3
    x = np.array([[0.0, 0, 0], [r, 0, 0]]); t = np.array([1, 2], np.int32); q = np.array([1.3, -0.7])
    k, qq = 14.399645, -0.91
    for sb in (1, 2, 3):
        ent = np.array([1 | (sb << 30)], np.int64).astype(np.uint32).view(np.int32)
        f, ev = orc.pair_eval(P, 0, 1, 0, 2, x, t, q, np.array([1, 0], np.int32), np.array([0, 1, 1]), ent)
        fl, fc = P.p.special_lj[sb], P.p.special_coul[sb]
        e_b = 1800.0 * math.exp(-r / 0.29) - 130.0 / r ** 6
        assert ev[0] == pytest.approx(fl * e_b, rel=1e-12)
        assert ev[1] == pytest.approx(k * qq / r * (math.erfc(g * r) - (1 - fc)), abs=abs(k * qq / r) * 2e-7)
    # table branch: interpolation error of the 12-bit table is ~1e-6 relative to the analytic value
    tabs = P.make_coul_tables(10.0)
    f_t, ev_t = orc.pair_eval(P, 0, 1, 0, 2, x, t, q, np.array([1, 0], np.int32), np.array([0, 1, 1]), np.array([1], np.int32))
    assert ev_t[1] == pytest.approx(k * qq * math.erfc(g * r) / r, rel=5e-6)
    # the product's host-side table builder (Pair::init_tables in Python) matches the oracle's (to libm rounding)
    t2, mask, shift, inner = pkg.init_coul_tables(10.0, g, 14.399645)
    assert (mask, shift) == (P.p.ncoulmask, P.p.ncoulshiftbits) and inner == P.p.tabinnersq
    for kk in tabs:
        assert np.allclose(tabs[kk], t2[kk], rtol=1e-9, atol=1e-13 * np.abs(tabs[kk]).max()), kk


def test_half_list_equals_brute_force_pair_set(orc, W):
    for s, co, st, nt in ((W.fcc_system(6, 6, 6), W.coeffs_in_buck(2.5), orc.BUCK, 1),
                          (W.aC_system(1), W.coeffs_aC(10.0, 10.0), orc.BUCK_COUL_CUT, 2)):
        P = orc.Params(st, nt, co["A"], co["rho"], co["C"], co["cut_lj"], co.get("cut_coul"))
        n = len(s["x"])
        cm = P.cutmax() + 0.3
        for prec in (0, 1):
            xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cm)
            cns = P.cutneighsq(0.3)
            hn, hoff, hent = orc.neigh_half_bin(n, xa, ta, nt, cns, s["boxlo"], s["boxhi"], cm, prec)
            fn, foff, fent = orc.neigh_full_brute(n, xa, ta, nt, cns, prec)
            assert np.array_equal(util.pair_keys(n, hn, hent, src, shift, symmetrize=True),
                                  util.pair_keys(n, fn, fent, src, shift))
            # i in N(j) <=> j in N(i)
            kf = util.pair_keys(n, fn, fent, src, shift)
            assert 2 * hoff[-1] == foff[-1] == len(kf)


def test_newton_sum_and_counts(orc, W):
    s = W.fcc_system(8, 8, 8)
    co = W.coeffs_in_buck(2.5)
    P = orc.Params(orc.BUCK, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
    f, ev, aux = orc.pair_forces_periodic(P, 0, s["x"], s["type"], None, s["boxlo"], s["boxhi"], 0.3)
    assert np.abs(f[:, :3].sum(0)).max() < 1e-10
    # 38.8 half-list neighbours per atom at cut+skin for in.buck
    assert aux["offsets"][-1] / len(s["x"]) == pytest.approx(38.8, abs=0.6)
    # virial by pair tally equals f.r virial
    _, ev2, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], None, s["boxlo"], s["boxhi"], 0.3, vflag=2)
    assert np.allclose(ev[2:], ev2[2:], rtol=1e-9, atol=1e-9)


def test_fft3d_matches_numpy(orc):
    rng = np.random.default_rng(3)
    for shape in ((8, 6, 5), (12, 10, 9), (27, 25, 16), (40, 36, 36)):
        a = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        fwd = orc.fft3d(a, 1)
        # LAMMPS FFT3d: flag=+1 is exp(+ikx) = numpy's unnormalised inverse
        assert np.abs(fwd - np.fft.ifftn(a) * a.size).max() <= 1e-11 * np.abs(fwd).max()
        back = orc.fft3d(fwd, -1) / a.size
        assert np.abs(back - a).max() <= 1e-12


def test_pppm_plus_real_space_equals_ewald(orc, W):
    """what in.buck_coul_long:12 really runs is kspace_style ewald: a tight PPPM + erfc real space must give
    the same Coulomb forces/energy as a direct Ewald sum"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    g = 0.30
    fe, ee, ve = orc.ewald_recip(s["x"], s["q"], s["boxlo"], s["boxhi"], g, 14, u["qqrd2e"])
    pp = orc.PPPM(60, 60, 64, 7, g, s["boxlo"], s["boxhi"], u["qqrd2e"])
    fp, ep, vp = pp.compute(s["x"], s["q"])
    scale = np.abs(fe).max()
    assert np.abs(fp - fe).max() / scale < 2e-5
    assert ep == pytest.approx(ee, rel=1e-6)
    assert np.allclose(vp, ve, rtol=1e-4, atol=1e-4 * np.abs(ve).max())
    # ad differentiation converges to the same answer
    pa = orc.PPPM(60, 60, 64, 7, g, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=1)
    fa, ea, va = pa.compute(s["x"], s["q"])
    assert np.abs(fa - fe).max() / scale < 2e-4
    assert ea == pytest.approx(ee, rel=1e-6)


def _sheared(s, tilt):
    """the atoms of an orthogonal system carried along with a shear of its box (same lamda coordinates)"""
    xy, xz, yz = tilt
    prd = s["boxhi"] - s["boxlo"]
    lam = (s["x"] - s["boxlo"]) / prd
    x = np.empty_like(s["x"])
    x[:, 0] = s["boxlo"][0] + prd[0] * lam[:, 0] + xy * lam[:, 1] + xz * lam[:, 2]
    x[:, 1] = s["boxlo"][1] + prd[1] * lam[:, 1] + yz * lam[:, 2]
    x[:, 2] = s["boxlo"][2] + prd[2] * lam[:, 2]
    return x


def test_triclinic_pppm_known_answers(orc, W):
    """PPPMIntel::compute on a triclinic box (pppm_intel.cpp:151-156, 878-883 -> stock setup_triclinic,
    compute_gf_ik_triclinic, poisson_ik_triclinic, restated): (1) zero tilt is the orthogonal path; (2) a tight mesh
    reproduces the direct Ewald sum over the reciprocal lattice of the tilted cell; (3) tilting the box by one whole
    lattice vector (xy = xprd) describes the same crystal: same energy and forces"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    g = 0.30
    lo, hi = s["boxlo"], s["boxhi"]
    f0, e0, v0 = orc.PPPM(24, 24, 27, 5, g, lo, hi, u["qqrd2e"]).compute(s["x"], s["q"])
    ft, et, vt = orc.PPPM.triclinic(24, 24, 27, 5, g, lo, hi, (0.0, 0.0, 0.0), u["qqrd2e"]).compute(s["x"], s["q"])
    assert et == pytest.approx(e0, rel=1e-12) and np.abs(ft - f0).max() <= 1e-11 * np.abs(f0).max()
    assert np.allclose(vt, v0, rtol=0, atol=1e-11 * np.abs(v0).max())
    # (2) a genuinely tilted cell
    tilt = (3.0, -2.0, 4.0)
    x = _sheared(s, tilt)
    fe, ee, ve = orc.ewald_recip_tri(x, s["q"], lo, hi, tilt, g, 14, u["qqrd2e"])
    fp, ep, vp = orc.PPPM.triclinic(60, 60, 64, 7, g, lo, hi, tilt, u["qqrd2e"]).compute(x, s["q"])
    scale = np.abs(fe).max()
    assert np.abs(fp - fe).max() / scale < 2e-5
    assert ep == pytest.approx(ee, rel=1e-6)
    assert np.allclose(vp, ve, rtol=1e-4, atol=1e-4 * np.abs(ve).max())
    assert np.abs(fe - orc.ewald_recip(s["x"], s["q"], lo, hi, g, 14, u["qqrd2e"])[0]).max() > 0.05 * scale   # not the cube again
    # (3) xy = xprd: the same lattice; atoms wrapped into the tilted cell
    prd = hi - lo
    tilt1 = (prd[0], 0.0, 0.0)
    lam = np.empty_like(s["x"])
    d = s["x"] - lo
    lam[:, 2] = d[:, 2] / prd[2]
    lam[:, 1] = d[:, 1] / prd[1]
    lam[:, 0] = (d[:, 0] - tilt1[0] * lam[:, 1]) / prd[0]
    lam -= np.floor(lam)
    xw = np.column_stack([lo[0] + prd[0] * lam[:, 0] + tilt1[0] * lam[:, 1], lo[1] + prd[1] * lam[:, 1],
                          lo[2] + prd[2] * lam[:, 2]])
    fo, eo, _ = orc.ewald_recip(s["x"], s["q"], lo, hi, g, 14, u["qqrd2e"])
    f1, e1, _ = orc.ewald_recip_tri(xw, s["q"], lo, hi, tilt1, g, 16, u["qqrd2e"])
    assert e1 == pytest.approx(eo, rel=1e-10) and np.abs(f1 - fo).max() <= 1e-9 * np.abs(fo).max()
    # the 45-degree cell is a different (coarser) discretisation of the same crystal: a tight mesh converges to it
    fq, eq, _ = orc.PPPM.triclinic(60, 60, 64, 7, g, lo, hi, tilt1, u["qqrd2e"]).compute(xw, s["q"])
    assert eq == pytest.approx(eo, rel=1e-8) and np.abs(fq - fo).max() <= 2e-6 * np.abs(fo).max()


def test_madelung_constant_rocksalt(orc, pkg):
    """NaCl rock salt: E per ion pair = -M q^2 / r0 with M = 1.747565, from PPPM + real-space erfc"""
    nc, a = 4, 5.64
    r0 = a / 2
    idx = np.stack(np.meshgrid(*[np.arange(2 * nc)] * 3, indexing="ij"), -1).reshape(-1, 3)
    x = idx * r0 + 0.25
    q = np.where(idx.sum(1) % 2 == 0, 1.0, -1.0)
    t = np.where(q > 0, 1, 2).astype(np.int32)
    lo, hi = np.zeros(3), np.full(3, nc * a)
    k = 14.399645
    g = 0.40
    A = np.zeros((3, 3)); rho = np.ones((3, 3)); C = np.zeros((3, 3))
    P = orc.Params(orc.BUCK_COUL_LONG, 2, A, rho, C, np.full((3, 3), 9.0), np.full((3, 3), 9.0), qqrd2e=k, g_ewald=g)
    P.arr["cut_ljsq"][:] = 0.0  # Coulomb only
    f, ev, _ = orc.pair_forces_periodic(P, 0, x, t, q, lo, hi, 0.3)
    pp = orc.PPPM(48, 48, 48, 7, g, lo, hi, k)
    fk, ek, _ = pp.compute(x, q)
    npairs = len(x) / 2
    M = -(ev[1] + ek) / npairs * r0 / k
    assert M == pytest.approx(1.747565, abs=3e-5)
    assert np.abs(f[:, :3] + fk).max() < 1e-3  # perfect lattice: forces vanish


def test_pppm_sizing_heuristic(orc, W):
    """PPPM::set_grid_global restated: grids are 2,3,5-smooth and land near the SURVEY §6.2 estimates"""
    s = W.aC_system(1)
    x, t, q, lo, hi = W.replicate(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 2, 2, 2)
    qsq = float((q * q).sum())
    for acc, expect in ((1e-4, (36, 36, 40)), (1e-5, (60, 60, 64))):
        grid, g = orc.pppm_size(acc, 14.399645, qsq, len(x), 12.0, hi - lo, order=5)
        for n in grid:
            m = n
            for p in (2, 3, 5):
                while m % p == 0:
                    m //= p
            assert m == 1
        assert all(abs(a - b) <= 0.25 * b for a, b in zip(grid, expect)), (grid, expect)
        assert 0.2 < g < 0.4


def test_nve_matches_closed_form(orc):
    x = np.array([[0.0, 1.0, 2.0]]); v = np.array([[0.5, -0.25, 0.125]]); f = np.array([[2.0, 4.0, -8.0]])
    dtfm = orc.nve_dtfm(np.array([1], np.int32), np.array([0.0, 4.0]), 0.01, 2.0)
    assert np.allclose(dtfm, 0.5 * 0.01 * 2.0 / 4.0)
    x1, v1 = orc.nve_initial(x, v, f, dtfm, 0.01)
    assert np.allclose(v1, v + dtfm.reshape(1, 3) * f) and np.allclose(x1, x + 0.01 * v1)
    assert np.allclose(orc.nve_final(v1, f, dtfm), v1 + dtfm.reshape(1, 3) * f)


def _disp_system(W):
    s = W.aC_system(1)
    B = np.array([0.0, 9.0, 13.2])                 # geometric mixing: C_ij = B_i B_j
    C = np.outer(B, B)
    A = np.zeros((3, 3)); rho = np.ones((3, 3))    # pure dispersion (no repulsive wall: only -C/r^6 is checked)
    return s, B, C, A, rho


def _disp_total(orc, s, B, C, A, rho, g6, rc=11.0, grid=(54, 54, 60), order=7):
    P = orc.Params(orc.BUCK_LONG_COUL_LONG, 2, A, rho, C, np.full((3, 3), rc), np.full((3, 3), rc), qqrd2e=14.399645,
                   g_ewald_6=g6, order6=1)
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    pp = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"])
    fk, ek, vk = pp.compute(s["x"], B[s["type"]])
    return f[:, :3] + fk, ev[0] + ek, ev[2:] + vk


def test_dispersion_ewald_is_independent_of_g_ewald_6(orc, W):
    """pppm/disp 'g' path + buck/long/coul/long ORDER6 real space: the split parameter must drop out of the total
    energy, forces and virial (pins compute_gf_6, the self terms of pppm_disp_intel.cpp:486-510 and vg_6), and the
    virial trace of an r^-6 potential is 6 E."""
    s, B, C, A, rho = _disp_system(W)
    f1, e1, v1 = _disp_total(orc, s, B, C, A, rho, 0.28)
    f2, e2, v2 = _disp_total(orc, s, B, C, A, rho, 0.36)
    assert e1 == pytest.approx(e2, rel=2e-6)
    assert np.abs(f1 - f2).max() <= 2e-5 * np.abs(f1).max()
    assert np.allclose(v1, v2, rtol=0, atol=2e-5 * np.abs(v1).max())
    assert v1[0] + v1[1] + v1[2] == pytest.approx(6.0 * e1, rel=2e-5)


def test_dispersion_ewald_equals_direct_lattice_sum(orc, W):
    """... and equals the direct sum of -B_i B_j / r^6 over periodic images (R = 38 A + continuum tail)"""
    s, B, C, A, rho = _disp_system(W)
    ft, et, _ = _disp_total(orc, s, B, C, A, rho, 0.30)
    x, w = s["x"], B[s["type"]]
    prd = s["boxhi"] - s["boxlo"]
    R = 38.0
    nimg = [int(np.ceil(R / p)) for p in prd]
    shifts = np.array([[i, j, k] for i in range(-nimg[0], nimg[0] + 1) for j in range(-nimg[1], nimg[1] + 1)
                       for k in range(-nimg[2], nimg[2] + 1)], float) * prd
    e = 0.0
    f = np.zeros_like(x)
    for sh in shifts:
        d = x[:, None, :] - (x[None, :, :] + sh)          # r_i - r_j'
        r2 = (d * d).sum(-1)
        m = (r2 < R * R) & (r2 > 1e-12)
        ww = w[:, None] * w[None, :]
        r2m = np.where(m, r2, 1.0)
        inv6 = np.where(m, 1.0 / r2m ** 3, 0.0)
        e += -0.5 * (ww * inv6).sum()
        # F_i = -dU/dr_i, U = -ww/r^6  ->  F_i = -6 ww d / r^8
        f += (-6.0 * ww * inv6 / r2m)[:, :, None].__mul__(d).sum(1)
    vol = prd.prod()
    e += -0.5 * w.sum() ** 2 / vol * 4.0 * np.pi / (3.0 * R ** 3)
    assert et == pytest.approx(e, rel=2e-5)
    assert np.abs(ft - f).max() <= 1e-4 * np.abs(f).max()


def test_dispersion_grid_ad_differentiation_converges_to_ik(orc, W):
    """kspace_modify diff ad on the geometric dispersion grid (fieldforce_g_ad, compute_sf_coeff_6 [UPSTREAM]): the
    energy and virial are those of ik (same poisson sums), the forces converge to the ik forces on a fine mesh, and
    the self-force coefficients scale like the Coulomb ones (zero net force is restored to the mesh accuracy)"""
    s, B, C, A, rho = _disp_system(W)
    w = B[s["type"]]
    g6, grid = 0.30, (54, 54, 60)
    fi, ei, vi = orc.PPPM.dispersion(*grid, 7, g6, s["boxlo"], s["boxhi"]).compute(s["x"], w)
    pa = orc.PPPM.dispersion(*grid, 7, g6, s["boxlo"], s["boxhi"], diff_ad=1)
    fa, ea, va = pa.compute(s["x"], w)
    assert ea == pytest.approx(ei, rel=1e-12) and np.allclose(va, vi, rtol=1e-12, atol=0)
    scale = np.abs(fi).max()
    assert np.abs(fa - fi).max() < 2e-4 * scale
    assert np.abs(fa.sum(0)).max() < 1e-3 * scale
    assert np.abs(pa.sf_coeff()).max() > 0.0
    # a coarse mesh separates the two schemes: they are different discretisations, not the same code path
    fc_i = orc.PPPM.dispersion(24, 24, 27, 5, g6, s["boxlo"], s["boxhi"]).compute(s["x"], w)[0]
    fc_a = orc.PPPM.dispersion(24, 24, 27, 5, g6, s["boxlo"], s["boxhi"], diff_ad=1).compute(s["x"], w)[0]
    assert np.abs(fc_a - fc_i).max() > 10 * np.abs(fa - fi).max()


def _lb_matrix(eps, sig):
    """lj4_ij = 4 eps_ij sigma_ij^6 with Lorentz-Berthelot mixing (pair lj/long/coul/long, mix arithmetic)"""
    return 4.0 * np.sqrt(np.outer(eps, eps)) * ((sig[:, None] + sig[None, :]) / 2.0) ** 6


def _mixed_total(orc, s, Cij, g6, kspace, rc=11.0, grid=(54, 54, 60), order=7):
    """real space (buck/long/coul/long ORDER6 with A = 0: only -C_ij/r^6) + the k-space part given by kspace(pp)"""
    A = np.zeros((3, 3)); rho = np.ones((3, 3))
    P = orc.Params(orc.BUCK_LONG_COUL_LONG, 2, A, rho, Cij, np.full((3, 3), rc), np.full((3, 3), rc), qqrd2e=14.399645,
                   g_ewald_6=g6, order6=1)
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    pp = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"])
    fk, ek, vk = kspace(pp)
    return f[:, :3] + fk, ev[0] + ek, ev[2:] + vk


def test_arithmetic_mixing_coefficients(orc):
    """PPPMDisp::init_coeffs, function[2] restated: sum_k B_i[k] B_j[6-k] is lj4_ij of Lorentz-Berthelot mixing"""
    eps = np.array([0.0, 0.8, 2.1, 0.05]); sig = np.array([0.0, 2.9, 3.6, 1.7])
    B7 = orc.disp_B_arithmetic(eps, sig)
    C = np.einsum("ik,jk->ij", B7, B7[:, ::-1])
    assert np.allclose(C[1:, 1:], _lb_matrix(eps, sig)[1:, 1:], rtol=1e-14, atol=0)


@pytest.mark.parametrize("ad", [0, 1])
def test_arithmetic_mixing_with_equal_sigma_is_geometric_mixing(orc, W, ad):
    """function[2] (seven coupled grids, poisson_2s) against function[1] (one grid): with one sigma for every type
    C_ij = B_i B_j, B_i = 2 sqrt(eps_i) sigma^3, and both paths must give the same energy, virial and forces"""
    s = W.aC_system(1)
    eps = np.array([0.0, 0.8, 2.1]); sig = np.array([0.0, 3.1, 3.1])
    B7 = orc.disp_B_arithmetic(eps, sig)
    Bg = 2.0 * np.sqrt(eps) * sig ** 3
    grid, order, g6 = (24, 24, 27), 5, 0.30
    fa, ea, va = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"], diff_ad=ad).compute_arith(s["x"], B7[s["type"]])
    fg, eg, vg = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"], diff_ad=ad).compute(s["x"], Bg[s["type"]])
    assert ea == pytest.approx(eg, rel=1e-11)
    assert np.allclose(va, vg, rtol=0, atol=1e-11 * np.abs(vg).max())
    assert np.abs(fa - fg).max() <= 2e-9 * np.abs(fg).max()   # seven grids whose contributions largely cancel
    # without energy / virial poisson_2s packs two densities into one transform: same forces
    fa0 = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"], diff_ad=ad).compute_arith(
        s["x"], B7[s["type"]], eflag=0, vflag=0)[0]
    assert np.abs(fa0 - fa).max() <= 2e-9 * np.abs(fa).max()


@pytest.mark.parametrize("ad", [0, 1])
def test_no_mixing_rule_on_the_arithmetic_matrix_equals_arithmetic_mixing(orc, W, ad):
    """function[3] (eigen-grids of C_ij) fed with the Lorentz-Berthelot matrix against function[2] (seven grids): two
    different decompositions of the same coefficients; and a rank-one matrix reduces to geometric mixing"""
    s = W.aC_system(1)
    eps = np.array([0.0, 0.8, 2.1]); sig = np.array([0.0, 2.9, 3.6])
    B7 = orc.disp_B_arithmetic(eps, sig)
    Bn, lam = orc.disp_B_none(_lb_matrix(eps, sig))
    assert len(lam) == 2 and lam.min() < 0 < lam.max()      # indefinite: a negative eigen-grid is exercised
    grid, order, g6 = (24, 24, 27), 5, 0.30
    mk = lambda: orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"], diff_ad=ad)
    fa, ea, va = mk().compute_arith(s["x"], B7[s["type"]])
    fn, en, vn = mk().compute_none(s["x"], Bn[s["type"]], lam)
    assert en == pytest.approx(ea, rel=1e-10)
    assert np.allclose(vn, va, rtol=0, atol=1e-10 * np.abs(va).max())
    assert np.abs(fn - fa).max() <= 2e-9 * np.abs(fa).max()
    Bg = np.array([0.0, 9.0, 13.2])
    B1, lam1 = orc.disp_B_none(np.outer(Bg, Bg))
    assert len(lam1) == 1
    f1, e1, v1 = mk().compute_none(s["x"], B1[s["type"]], lam1)
    fg, eg, vg = mk().compute(s["x"], Bg[s["type"]])
    assert e1 == pytest.approx(eg, rel=1e-11) and np.abs(f1 - fg).max() <= 1e-11 * np.abs(fg).max()
    assert np.allclose(v1, vg, rtol=0, atol=1e-11 * np.abs(vg).max())


def test_arithmetic_mixing_ewald_is_independent_of_g_ewald_6(orc, W):
    """function[2] + the ORDER6 real space with lj4_ij of Lorentz-Berthelot mixing: the split parameter drops out of
    energy, forces and virial - this fixes the absolute scale of B[7 i + k] (a wrong prefactor leaves a g-dependent
    remainder) - and the virial trace of an r^-6 potential is 6 E"""
    s = W.aC_system(1)
    eps = np.array([0.0, 0.8, 2.1]); sig = np.array([0.0, 2.9, 3.6])
    Cij = _lb_matrix(eps, sig)
    w7 = orc.disp_B_arithmetic(eps, sig)[s["type"]]
    f1, e1, v1 = _mixed_total(orc, s, Cij, 0.28, lambda pp: pp.compute_arith(s["x"], w7))
    f2, e2, v2 = _mixed_total(orc, s, Cij, 0.36, lambda pp: pp.compute_arith(s["x"], w7))
    assert e1 == pytest.approx(e2, rel=2e-6)
    assert np.abs(f1 - f2).max() <= 2e-5 * np.abs(f1).max()
    assert np.allclose(v1, v2, rtol=0, atol=2e-5 * np.abs(v1).max())
    assert v1[0] + v1[1] + v1[2] == pytest.approx(6.0 * e1, rel=2e-5)


def test_nve_group_branch_freezes_atoms_outside_the_group(orc):
    """FixNVEIntel with igroup != all (fix_nve_intel.cpp:88-97, 173-190): dtfm is 0 outside the group and those
    atoms keep x and v; inside, v += dtf/m f and x += dt v with rmass overriding the per-type mass"""
    type_ = np.array([1, 2, 1, 2], dtype=np.int32)
    mass = np.array([0.0, 2.0, 4.0])
    rmass = np.array([1.0, 8.0, 16.0, 32.0])
    ingroup = np.array([1, 0, 1, 1], dtype=np.int32)
    dt, ftm2v = 0.5, 2.0            # dtf = 0.5
    d = orc.nve_dtfm_group(type_, mass, dt, ftm2v, ingroup, rmass)
    assert np.array_equal(d.reshape(4, 3)[:, 0], [0.5, 0.0, 0.5 / 16.0, 0.5 / 32.0])
    d2 = orc.nve_dtfm_group(type_, mass, dt, ftm2v, ingroup, None)
    assert np.array_equal(d2.reshape(4, 3)[:, 0], [0.25, 0.0, 0.25, 0.125])
    assert np.array_equal(orc.nve_dtfm_group(type_, mass, dt, ftm2v), orc.nve_dtfm(type_, mass, dt, ftm2v))
    x = np.arange(12.0).reshape(4, 3)
    v = np.ones((4, 3))
    f = np.full((4, 3), 4.0)
    xo, vo = orc.nve_initial_group(x, v, f, d, dt)
    assert np.array_equal(vo[1], v[1]) and np.array_equal(xo[1], x[1])          # frozen
    assert np.array_equal(vo[0], [3.0, 3.0, 3.0]) and np.array_equal(xo[0], x[0] + 1.5)
    assert np.array_equal(vo[2], 1.0 + 4.0 * 0.5 / 16.0 * np.ones(3))


def test_pppm_peratom_sums_to_the_global_tallies(orc, W):
    """stock poisson_peratom / fieldforce_peratom restated: interpolation uses the assignment function the density was
    spread with, so by Parseval sum_i eatom_i = E and sum_i vatom_i = virial to round-off, for ik and ad"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    for ad in (0, 1):
        pp = orc.PPPM(24, 24, 27, 5, 0.28, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad)
        f, e, v = pp.compute(s["x"], s["q"], eflag=3, vflag=5)
        ea, va = pp.peratom()
        assert ea.sum() == pytest.approx(e, rel=1e-11)
        assert np.allclose(va.sum(0), v, rtol=1e-10, atol=1e-10 * np.abs(v).max())
        # the flags only add the tallies: forces, energy and virial are those of a plain compute
        f0, e0, v0 = pp.compute(s["x"], s["q"])
        assert np.array_equal(f, f0) and e == e0 and np.array_equal(v, v0)
        assert not pp.peratom()[0].any()   # nothing tallied without the flags


def test_peratom_energy_of_rocksalt_is_half_the_madelung_energy_per_ion(orc):
    """every ion of NaCl carries -M k q^2 / (2 r0): pins the per-atom self-energy term and the factors 1/2 of
    PPPM::compute's eatom post-processing together with the pair style's eatom"""
    nc, a = 4, 5.64
    r0 = a / 2
    idx = np.stack(np.meshgrid(*[np.arange(2 * nc)] * 3, indexing="ij"), -1).reshape(-1, 3)
    x = idx * r0 + 0.25
    q = np.where(idx.sum(1) % 2 == 0, 1.0, -1.0)
    t = np.where(q > 0, 1, 2).astype(np.int32)
    lo, hi = np.zeros(3), np.full(3, nc * a)
    k, g = 14.399645, 0.40
    A = np.zeros((3, 3)); rho = np.ones((3, 3)); C = np.zeros((3, 3))
    P = orc.Params(orc.BUCK_COUL_LONG, 2, A, rho, C, np.full((3, 3), 9.0), np.full((3, 3), 9.0), qqrd2e=k, g_ewald=g)
    P.arr["cut_ljsq"][:] = 0.0
    f, ev, _ = orc.pair_forces_periodic(P, 0, x, t, q, lo, hi, 0.3, eatom=1)
    assert f[:, 3].sum() == pytest.approx(ev[0] + ev[1], rel=1e-12)
    pp = orc.PPPM(48, 48, 48, 7, g, lo, hi, k)
    pp.compute(x, q, eflag=3, vflag=0)
    ek, _ = pp.peratom(vatom=False)
    e_ion = f[:len(x), 3] + ek
    assert np.allclose(e_ion, -1.747565 * k / (2 * r0), rtol=0, atol=2e-4)


def _slab_system(n=160, L=12.0, seed=3):
    rng = np.random.default_rng(seed)
    x = np.column_stack([rng.uniform(0, L, n), rng.uniform(0, L, n), rng.uniform(2.5, 8.5, n)])
    q = np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
    q[x[:, 2] > 5.5] *= 1.5            # a net dipole along z ...
    q -= q.mean()                      # ... in a neutral cell
    return x, q, np.zeros(3), np.full(3, L)


def test_slab_correction_known_answers(orc):
    """kspace_modify slab (PPPM::setup with zprd_slab, PPPM::slabcorr at pppm_intel.cpp:305), four pins:
    (1) against the plain periodic solver on the same extended cell the energy differs by exactly
        qqrd2e 2 pi M_z^2 / V and the forces by -4 pi qqrd2e q M_z / V (neutral cell);
    (2) the corrected energy no longer depends on the amount of vacuum (volfactor 3 vs 5), the uncorrected one does;
    (3) it is invariant under a rigid shift along z inside the box;
    (4) f_z = -dE/dz."""
    x, q, lo, hi = _slab_system()
    k, g, L = 14.399645, 0.55, hi[0]
    f3, e3, _ = orc.PPPM(36, 36, 108, 7, g, lo, hi, k, slab=3.0).compute(x, q)
    # (1) the same mesh as a fully periodic cell of height 3 L
    hi3 = hi * np.array([1, 1, 3.0])
    fp, ep, _ = orc.PPPM(36, 36, 108, 7, g, lo, hi3, k).compute(x, q)
    M, V = (q * x[:, 2]).sum(), L * L * 3 * L
    assert abs(M) > 2.0
    assert e3 - ep == pytest.approx(k * 2 * math.pi * M * M / V, rel=1e-9)
    df = f3 - fp
    assert np.abs(df[:, :2]).max() == 0.0
    assert np.allclose(df[:, 2], -4 * math.pi * k * q * M / V, rtol=1e-9, atol=1e-12)
    # (2)
    f5, e5, _ = orc.PPPM(36, 36, 180, 7, g, lo, hi, k, slab=5.0).compute(x, q)
    hi5 = hi * np.array([1, 1, 5.0])
    _, ep5, _ = orc.PPPM(36, 36, 180, 7, g, lo, hi5, k).compute(x, q)
    assert abs(e5 - e3) < 2e-5 * abs(e3)
    assert abs(ep5 - ep) > 50 * abs(e5 - e3)
    assert np.abs(f5 - f3).max() < 1e-4 * np.abs(f3).max()
    # (3)
    xs = x + np.array([0.0, 0.0, 1.25])
    fs, es, _ = orc.PPPM(36, 36, 108, 7, g, lo, hi, k, slab=3.0).compute(xs, q)
    assert abs(es - e3) < 2e-5 * abs(e3) and np.abs(fs - f3).max() < 2e-4 * np.abs(f3).max()
    # (4)
    i, d = 7, 1e-4
    pp = orc.PPPM(36, 36, 108, 7, g, lo, hi, k, slab=3.0)
    xp, xm = x.copy(), x.copy()
    xp[i, 2] += d
    xm[i, 2] -= d
    num = -(pp.compute(xp, q)[1] - pp.compute(xm, q)[1]) / (2 * d)
    assert abs(num - f3[i, 2]) < 5e-4 * np.abs(f3).max()


def test_md_loop_stops_when_the_model_blows_up(orc, W):
    """Buckingham's -C/r^6 wins over A exp(-r/rho) at short range: a system that is hot enough collapses, atoms fly out of
    the box, and the oracle's MD loop reports it instead of binning coordinates that are out of range"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    co = W.coeffs_aC(8.0, 8.0)
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=0.3)
    s["v"] = s["v"] * 6.0
    pp = orc.PPPM(24, 24, 27, 5, 0.3, s["boxlo"], s["boxhi"], u["qqrd2e"])
    md = orc.MD(s, P, pppm=pp, skin=0.3, every=1, delay=0, check=1, dt=0.002, ftm2v=u["ftm2v"])
    with pytest.raises(RuntimeError, match="left the box"):
        md.run(40, 1)
