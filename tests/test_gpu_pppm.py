"""-m gpu parity: PPPM on the device vs the CPU oracle, same inputs, explicit (nx,ny,nz,order,g_ewald)
("PPPM parity at explicit grid ... so both sides use the same grid")."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(16, 16, 16), (8, 6, 5), (27, 25, 32), (40, 36, 36), (64, 60, 60), (108, 96, 96),
                                   (135, 50, 45), (12, 250, 270)])
def test_fft3d_matches_numpy(pkg, shape):
    rng = np.random.default_rng(11)
    a = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    ctx = pkg.Context(0, 0)
    fwd = ctx.fft3d(a, 1)      # LAMMPS flag=+1 == exp(+ikx)
    ref = np.fft.ifftn(a) * a.size
    assert np.abs(fwd - ref).max() <= 1e-12 * np.abs(ref).max()
    bwd = ctx.fft3d(a, -1)
    ref = np.fft.fftn(a)
    assert np.abs(bwd - ref).max() <= 1e-12 * np.abs(ref).max()
    back = ctx.fft3d(fwd, -1) / a.size
    assert np.abs(back - a).max() <= 1e-12
    with pytest.raises(pkg.B200MDError):
        ctx.fft3d(np.zeros((7, 8, 8), complex), 1)   # 7 is not 2^a 3^b 5^c
    ctx.close()


def _case(W, name):
    if name == "aC1":
        return W.aC_system(1), (24, 24, 27), 0.28
    if name == "aC2_1e-4":           # 9600 atoms, 36x36x40 @ 1e-4
        return W.aC_system(2), (36, 36, 40), 0.2472
    if name == "aC2_1e-5":
        return W.aC_system(2), (60, 60, 64), 0.2776
    if name == "water":              # S4 stand-in: 5184 atoms, neutral
        return W.water_like_system(12), (40, 40, 40), 0.30
    raise ValueError(name)


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("name,order,ad", [("aC1", 5, 0), ("aC1", 3, 0), ("aC1", 4, 0), ("aC1", 7, 0), ("aC1", 2, 0),
                                           ("aC2_1e-4", 5, 0), ("aC2_1e-5", 5, 0), ("water", 5, 0),
                                           ("aC1", 5, 1), ("aC1", 4, 1), ("aC2_1e-4", 5, 1), ("aC1", 7, 1)])
def test_pppm_matches_oracle(pkg, W, orc, name, order, ad, prec):
    s, grid, g = _case(W, name)
    u = W.UNITS[s["units"]]
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(2.0)
    ctx.pppm_setup(*grid, order, g, differentiation=ad)
    pp = orc.PPPM(*grid, order, g, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad, prec=prec)
    fo, eo, vo = pp.compute(s["x"], s["q"])
    # host-buffer form of PPPMIntel::compute
    f, e, v = ctx.pppm_compute_host(s["x"], s["q"], 1, 1)
    d = ctx.pppm_download()
    tg = 1e-12
    assert np.abs(d["greensfn"] - pp.greensfn()).max() <= tg * np.abs(pp.greensfn()).max()
    tol_grid = 1e-11 if prec == 0 else 2e-5
    rho = pp.density()
    if prec == 1:
        print("mixed grid: density err %.2e field err %.2e" % (np.abs(d["density"] - rho).max() / np.abs(rho).max(),
              np.abs(d["fx"] - pp.field(0)).max() / np.abs(pp.field(0)).max()))
    assert np.abs(d["density"] - rho).max() <= tol_grid * np.abs(rho).max()
    for k, c in enumerate(("fx", "fy", "fz")):
        ref = pp.field(k)
        assert np.abs(d[c] - ref).max() <= tol_grid * np.abs(ref).max(), c
        if ad:
            break
    tol_f = 1e-9 if prec == 0 else 1e-5
    tol_e = 1e-10 if prec == 0 else 1e-5
    assert util.rel_force_err(f, fo) <= tol_f
    assert abs(e - eo) <= tol_e * abs(eo)
    assert np.abs(v - vo).max() <= tol_e * np.abs(vo).max()
    if ad:
        assert np.abs(d["sf_coeff"] - pp.sf_coeff()).max() <= 1e-11 * np.abs(pp.sf_coeff()).max()
    # resident form: accumulates onto whatever the force array holds (f +=), host order preserved
    co = W.coeffs_aC(6.0, 6.0) if s["units"] == "metal" else None
    if co is not None:
        cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
        ctx.neigh_setup(0.3)
        ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=g)
        ctx.neigh_build()
        ctx.pair_compute(0, 0)
        fp = ctx.atoms_download(("f",))["f"]
        e2, v2 = ctx.pppm_compute(1, 1)
        ft = ctx.atoms_download(("f",))["f"]
        assert util.rel_force_err(ft - fp, fo) <= tol_f
        assert abs(e2 - eo) <= tol_e * abs(eo)
    ctx.close()


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("grid,order,tilt", [((24, 24, 27), 5, (3.0, -2.0, 4.0)), ((30, 32, 36), 7, (-6.0, 5.0, 1.5)),
                                             ((24, 25, 27), 4, (0.0, 0.0, 7.0)), ((24, 24, 27), 5, (0.0, 0.0, 0.0))])
def test_pppm_triclinic_matches_oracle(pkg, W, orc, grid, order, tilt, prec):
    """PPPMIntel::compute on a triclinic box (pppm_intel.cpp:151-156 x2lamda, :878-883 poisson_ik_triclinic, stock
    setup_triclinic / compute_gf_ik_triclinic): Green's function, density, field bricks, forces, energy and virial
    against the oracle (pinned by the Ewald sum over the tilted cell's reciprocal lattice); even and odd grid sizes
    exercise the Nyquist handling of the packed Ex + i Ey transform with the cross terms of the wave vector"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    g = 0.28
    lo, hi = s["boxlo"], s["boxhi"]
    prd = hi - lo
    lam = (s["x"] - lo) / prd
    x = np.column_stack([lo[0] + prd[0] * lam[:, 0] + tilt[0] * lam[:, 1] + tilt[1] * lam[:, 2],
                         lo[1] + prd[1] * lam[:, 1] + tilt[2] * lam[:, 2], lo[2] + prd[2] * lam[:, 2]])
    ctx = pkg.Context(0, prec)
    ctx.set_units(u["qqrd2e"], u["ftm2v"])
    ctx.set_box_triclinic(lo, hi, tilt)
    ctx.atoms_upload(x, s["type"], s["mass"], q=s["q"])
    ctx.neigh_setup(2.0)
    ctx.pppm_setup(*grid, order, g)
    # zero tilt through the triclinic entry point is the orthogonal path (float positions, not float lamda coordinates)
    pp = (orc.PPPM.triclinic(*grid, order, g, lo, hi, tilt, u["qqrd2e"], prec=prec) if any(tilt) else
          orc.PPPM(*grid, order, g, lo, hi, u["qqrd2e"], prec=prec))
    fo, eo, vo = pp.compute(x, s["q"])
    f, e, v = ctx.pppm_compute_host(x, s["q"], 1, 1)
    d = ctx.pppm_download()
    assert np.abs(d["greensfn"] - pp.greensfn()).max() <= 1e-12 * np.abs(pp.greensfn()).max()
    tol_grid = 1e-11 if prec == 0 else 2e-5
    assert np.abs(d["density"] - pp.density()).max() <= tol_grid * np.abs(pp.density()).max()
    for k, cc in enumerate(("fx", "fy", "fz")):
        ref = pp.field(k)
        assert np.abs(d[cc] - ref).max() <= tol_grid * np.abs(ref).max(), cc
    tol_f, tol_e = (1e-9, 1e-10) if prec == 0 else (1e-5, 1e-5)
    assert util.rel_force_err(f, fo) <= tol_f
    assert abs(e - eo) <= tol_e * abs(eo)
    assert np.abs(v - vo).max() <= tol_e * np.abs(vo).max()
    # resident form on the uploaded atoms: same numbers
    e2, v2 = ctx.pppm_compute(1, 1)
    assert util.rel_force_err(ctx.atoms_download(("f",))["f"], fo) <= tol_f and abs(e2 - eo) <= tol_e * abs(eo)
    if any(tilt):
        # what stock PPPM::init refuses on a triclinic box, and the parts of the path that need an orthogonal one
        for kw, msg in ((dict(differentiation=1), "kspace_modify diff ad"), (dict(dispersion=1, B=np.array([0.0, 1.0, 2.0])), "PPPMDisp")):
            with pytest.raises(pkg.B200MDError, match=msg):
                ctx.pppm_setup(*grid, order, g, **kw)
        co = W.coeffs_aC(6.0, 6.0)
        cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
        ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=g)
        with pytest.raises(pkg.B200MDError, match="triclinic"):
            ctx.neigh_build()
    else:
        # zero tilt through the triclinic entry point is the orthogonal path
        c2 = pkg.make_context(s, precision=prec)
        c2.neigh_setup(2.0)
        c2.pppm_setup(*grid, order, g)
        f0, e0, v0 = c2.pppm_compute_host(x, s["q"], 1, 1)
        assert util.rel_force_err(f, f0) <= (1e-13 if prec == 0 else 1e-6) and abs(e - e0) <= 1e-10 * abs(e0)
        c2.close()
    ctx.close()


@pytest.mark.parametrize("name,order,ad,prec", [("aC1", 5, 0, 0), ("aC2_1e-4", 5, 1, 0), ("water", 7, 0, 0),
                                                ("aC1", 4, 0, 1)])
def test_pppm_peratom_matches_oracle(pkg, W, orc, name, order, ad, prec):
    """eflag & 2 / vflag & 4: per-atom k-space energy and virial (stock poisson_peratom / fieldforce_peratom and the
    eatom / vatom post-factors of PPPM::compute) against the oracle; they sum to the global tallies"""
    s, grid, g = _case(W, name)
    u = W.UNITS[s["units"]]
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(2.0)
    ctx.pppm_setup(*grid, order, g, differentiation=ad)
    pp = orc.PPPM(*grid, order, g, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad, prec=prec)
    fo, eo, vo = pp.compute(s["x"], s["q"], eflag=3, vflag=5)
    eao, vao = pp.peratom()
    f0 = ctx.atoms_download(("f",))["f"]
    e, v = ctx.pppm_compute(3, 5)
    ea, va = ctx.pppm_peratom()
    tol = 1e-10 if prec == 0 else 2e-5
    assert np.abs(ea - eao).max() <= tol * np.abs(eao).max()
    assert np.abs(va - vao).max() <= tol * np.abs(vao).max()
    assert ea.sum() == pytest.approx(e, rel=1e-9 if prec == 0 else 1e-4)
    assert np.allclose(va.sum(0), v, rtol=1e-8 if prec == 0 else 1e-4, atol=1e-8 * np.abs(v).max())
    # the forces of the same call are the usual ones
    f1 = ctx.atoms_download(("f",))["f"]
    assert util.rel_force_err(f1 - f0, fo) <= (1e-9 if prec == 0 else 1e-5)
    # energy only: the virial bricks are skipped and asking for them is an error; a plain compute clears the tallies
    ctx.pppm_compute(3, 1)
    ea2, _ = ctx.pppm_peratom(vatom=False)
    assert np.array_equal(ea2, ea)
    with pytest.raises(pkg.B200MDError):
        ctx.pppm_peratom()
    ctx.pppm_compute(1, 1)
    with pytest.raises(pkg.B200MDError):
        ctx.pppm_peratom(vatom=False)
    ctx.close()


@pytest.mark.parametrize("ad,prec", [(0, 0), (1, 0), (0, 1)])
def test_slab_pppm_matches_oracle(pkg, W, orc, ad, prec):
    """`boundary p p f` + `kspace_modify slab 3.0`: mesh over zprd * 3 and PPPM::slabcorr (pppm_intel.cpp:305) —
    energy, virial, forces and per-atom energies against the oracle (itself pinned by four known answers)"""
    rng = np.random.default_rng(3)
    n, L = 400, 14.0
    x = np.column_stack([rng.uniform(0, L, n), rng.uniform(0, L, n), rng.uniform(2.5, 10.5, n)])
    q = np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
    q[x[:, 2] > 6.5] *= 1.5
    q -= q.mean() - 0.01          # slightly non-neutral: the qsum terms of slabcorr are exercised too
    s = dict(x=x, q=q, type=np.ones(n, np.int32), mass=np.array([0.0, 12.0]), boxlo=np.zeros(3), boxhi=np.full(3, L),
             units="metal", periodic=(1, 1, 0))
    u = W.UNITS["metal"]
    grid, g = (30, 30, 90), 0.35
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(1.0)
    ctx.pppm_setup(*grid, 5, g, differentiation=ad, slab=3.0)
    pp = orc.PPPM(*grid, 5, g, s["boxlo"], s["boxhi"], u["qqrd2e"], diff_ad=ad, prec=prec, slab=3.0)
    fo, eo, vo = pp.compute(x, q, eflag=3, vflag=1)
    eao, _ = pp.peratom(vatom=False)
    e, v = ctx.pppm_compute(3, 1)
    f = ctx.atoms_download(("f",))["f"]
    ea, _ = ctx.pppm_peratom(vatom=False)
    tf, te = (1e-9, 1e-10) if prec == 0 else (1e-5, 1e-5)
    assert util.rel_force_err(f, fo) <= tf
    assert abs(e - eo) <= te * abs(eo) and np.abs(v - vo).max() <= te * np.abs(vo).max()
    assert np.abs(ea - eao).max() <= (1e-10 if prec == 0 else 2e-5) * np.abs(eao).max()
    ctx.close()
    # a fully periodic box refuses the slab option, a p p f box refuses plain PPPM
    ctx = pkg.make_context(dict(s, periodic=(1, 1, 1)), precision=prec)
    with pytest.raises(pkg.B200MDError) as ei:
        ctx.pppm_setup(*grid, 5, g, slab=3.0)
    assert "Incorrect boundaries with slab PPPM" in str(ei.value)
    ctx.close()
    ctx = pkg.make_context(s, precision=prec)
    with pytest.raises(pkg.B200MDError) as ei:
        ctx.pppm_setup(*grid, 5, g)
    assert "Cannot use nonperiodic boundaries with PPPM" in str(ei.value)
    ctx.close()


def test_pppm_deterministic_and_flags(pkg, W):
    s, grid, g = _case(W, "aC1")
    outs = []
    for _ in range(2):
        ctx = pkg.make_context(s)
        ctx.pppm_setup(*grid, 5, g)
        f, e, v = ctx.pppm_compute_host(s["x"], s["q"], 1, 1)
        outs.append((f, e, v, ctx.pppm_download()["density"]))
        f0, e0, v0 = ctx.pppm_compute_host(s["x"], s["q"], 0, 0)
        assert e0 == 0.0 and not v0.any()
        assert np.array_equal(f0, f)    # same kernels for forces with or without tallies
        ctx.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])


def test_pppm_errors(pkg, W):
    s, grid, g = _case(W, "aC1")
    ctx = pkg.make_context(s)
    with pytest.raises(pkg.B200MDError) as ei:
        ctx.pppm_setup(*grid, 8, g)
    assert "PPPM order greater than supported" in str(ei.value)
    ctx.neigh_setup(0.3)
    ctx.pppm_setup(*grid, 5, g)
    x = s["x"].copy()
    x[17, 0] += 9.0 * (s["boxhi"][0] - s["boxlo"][0])
    with pytest.raises(pkg.B200MDError) as ei:
        ctx.pppm_compute_host(x, s["q"], 0, 0)
    assert "Out of range atoms - cannot compute PPPM" in str(ei.value)
    # uncharged system: returns without touching forces ("return if there are no charges", pppm_intel.cpp:149)
    f, e, v = ctx.pppm_compute_host(s["x"], np.zeros(len(x)), 1, 1)
    assert not f.any() and e == 0.0
    ctx.close()


def test_full_step_buck_coul_long_pppm(pkg, W, orc):
    """the north-star step: buck/coul/long + PPPM forces and energies vs the oracle, then a short NVE run
    conserves energy"""
    s = W.aC_system(2)
    u = W.UNITS["metal"]
    g, grid = 0.2776, (60, 60, 64)
    co = W.coeffs_aC(12.0, 12.0)
    ctx = pkg.make_context(s)
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    ctx.neigh_setup(0.3)
    ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=g)
    ctx.pppm_setup(*grid, 5, g)
    ctx.nve_setup(u["dt"])
    th = ctx.setup_forces(1, 1)
    f = ctx.atoms_download(("f",))["f"]
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=g)
    fo, evo, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    pp = orc.PPPM(*grid, 5, g, s["boxlo"], s["boxhi"], u["qqrd2e"])
    fk, ek, vk = pp.compute(s["x"], s["q"])
    assert util.rel_force_err(f, fo[:, :3] + fk) <= 1e-9
    assert abs(th[0] - evo[0]) <= 1e-10 * abs(evo[0]) and abs(th[1] - evo[1]) <= 1e-10 * abs(evo[1])
    assert abs(th[8] - ek) <= 1e-10 * abs(ek)
    assert np.abs(th[9:15] - vk).max() <= 1e-10 * np.abs(vk).max()
    e0 = th[0] + th[1] + th[8] + u["mvv2e"] * th[15]
    th1 = ctx.run(40, thermo=True)
    e1 = th1[0] + th1[1] + th1[8] + u["mvv2e"] * th1[15]
    # velocity Verlet at dt = 1 fs: the total energy wobbles by O((w dt)^2) of the kinetic energy, no drift
    assert abs(e1 - e0) <= 0.03 * u["mvv2e"] * th1[15], (e0, e1, u["mvv2e"] * th1[15])
    ctx.close()


def _resident(ctx, s):
    """PPPM alone on device-resident atoms: a fresh upload zeroes the force array, compute accumulates onto it"""
    ctx.atoms_upload(s["x"], s["type"], s["mass"], v=s.get("v"), q=s.get("q"))
    e, v = ctx.pppm_compute(1, 1)
    return ctx.atoms_download(("f",))["f"], e, v


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("ad", [0, 1])
@pytest.mark.parametrize("rule", ["arithmetic", "none"])
def test_pppm_disp_arithmetic_and_no_mixing_match_oracle(pkg, W, orc, rule, ad, prec):
    """PPPMDispIntel::compute function[2] (arithmetic mixing, seven coupled grids, pppm_disp_intel.cpp:315-407) and
    function[3] (no mixing rule, eigen-grids, :409-467): the device runs them as signed self-coupled components of the
    one-density kernels; forces, energy and virial against the oracle's restatement of the coupled-grid algorithm
    (make_rho_a / poisson_2s_ik|ad / fieldforce_a_ik|ad and the _none members), ik and ad, with and without the
    energy / virial tallies"""
    s = W.aC_system(1)
    eps = np.array([0.0, 0.8, 2.1]); sig = np.array([0.0, 2.9, 3.6])
    Cij = 4.0 * np.sqrt(np.outer(eps, eps)) * ((sig[:, None] + sig[None, :]) / 2.0) ** 6
    g6, grid, order = 0.31, (30, 30, 32), 5
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(0.3)
    pp = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"], prec=prec, diff_ad=ad)
    if rule == "arithmetic":
        B7 = orc.disp_B_arithmetic(eps, sig)
        ctx.pppm_setup(*grid, order, g6, dispersion=2, B=B7, differentiation=ad)
        fo, eo, vo = pp.compute_arith(s["x"], B7[s["type"]])
    else:
        Cn = Cij.copy()
        Cn[1, 2] = Cn[2, 1] = 0.7 * Cij[1, 2]          # not expressible by any mixing rule
        Bn, lam = orc.disp_B_none(Cn)
        ctx.pppm_setup(*grid, order, g6, dispersion=3, B=Cn, differentiation=ad)
        fo, eo, vo = pp.compute_none(s["x"], Bn[s["type"]], lam)
    f, e, v = _resident(ctx, s)
    # seven grids whose contributions largely cancel: 1e-8 of the largest force in double mode
    tol_f, tol_e = (1e-8, 1e-10) if prec == 0 else (2e-5, 1e-5)
    assert np.abs(f - fo).max() <= tol_f * np.abs(fo).max()
    assert abs(e - eo) <= tol_e * abs(eo)
    assert np.abs(v - vo).max() <= tol_e * np.abs(vo).max()
    # without the tallies the forces are the same bits
    ctx.atoms_upload(s["x"], s["type"], s["mass"], v=s.get("v"), q=s.get("q"))
    ctx.pppm_compute(0, 0)
    assert np.array_equal(ctx.atoms_download(("f",))["f"], f)
    # Coulomb grid + mixed dispersion grid in one compute (function[0] + function[2|3]): the sums
    u = W.UNITS["metal"]
    ctx.pppm_setup(24, 24, 27, 5, 0.28, differentiation=ad)
    f2, e2, v2 = _resident(ctx, s)
    fc, ec, vc = orc.PPPM(24, 24, 27, 5, 0.28, s["boxlo"], s["boxhi"], u["qqrd2e"], prec=prec, diff_ad=ad).compute(
        s["x"], s["q"])
    assert np.abs(f2 - (fo + fc)).max() <= 2 * tol_f * np.abs(fo + fc).max()
    assert abs(e2 - (eo + ec)) <= 2 * tol_e * abs(eo + ec)
    ctx.close()


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("order", [5, 3, 7])
def test_pppm_disp_geometric_matches_oracle(pkg, W, orc, order, prec):
    """PPPMDispIntel 'g' grid (pppm_disp_intel.cpp:245-313 with the per-atom weight B[type]):
    forces, energy (incl. the self terms :486-492) and virial (:498-510) against the oracle"""
    s = W.aC_system(1)
    B = np.array([0.0, 9.0, 13.2])
    g6 = 0.31
    grid = (30, 30, 32)
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(0.3)
    ctx.pppm_setup(*grid, order, g6, dispersion=1, B=B)
    f, e, v = _resident(ctx, s)
    pp = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"], prec=prec)
    fo, eo, vo = pp.compute(s["x"], B[s["type"]])
    tol_f, tol_e = (1e-9, 1e-10) if prec == 0 else (1e-5, 1e-5)
    assert np.abs(f - fo).max() <= tol_f * np.abs(fo).max()
    assert abs(e - eo) <= tol_e * abs(eo)
    assert np.abs(v - vo).max() <= tol_e * np.abs(vo).max()
    # kspace_modify diff ad on the dispersion grid (fieldforce_g_ad, compute_sf_coeff_6)
    ctx.pppm_setup(*grid, order, g6, dispersion=1, B=B, differentiation=1)
    fa, ea, va = _resident(ctx, s)
    pa = orc.PPPM.dispersion(*grid, order, g6, s["boxlo"], s["boxhi"], prec=prec, diff_ad=1)
    fao, eao, vao = pa.compute(s["x"], B[s["type"]])
    assert np.abs(fa - fao).max() <= tol_f * np.abs(fao).max()
    assert abs(ea - eao) <= tol_e * abs(eao) and np.abs(va - vao).max() <= tol_e * np.abs(vao).max()
    assert np.abs(fa - f).max() > 1e-6 * np.abs(f).max()      # not the ik path again
    ctx.pppm_setup(*grid, order, g6, dispersion=1, B=B)
    # Coulomb and dispersion grids together (pppm/disp with function[0] and function[1]): sums of the two
    u = W.UNITS["metal"]
    ctx.pppm_setup(24, 24, 27, 5, 0.28)
    f2, e2, v2 = _resident(ctx, s)
    pc = orc.PPPM(24, 24, 27, 5, 0.28, s["boxlo"], s["boxhi"], u["qqrd2e"], prec=prec)
    fc, ec, vc = pc.compute(s["x"], s["q"])
    assert np.abs(f2 - (fo + fc)).max() <= tol_f * np.abs(fo + fc).max() * 2
    assert abs(e2 - (eo + ec)) <= tol_e * abs(eo + ec) * 2
    assert np.abs(v2 - (vo + vc)).max() <= tol_e * np.abs(vo + vc).max() * 2
    ctx.close()
