"""-m gpu parity tests proper: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): neighbour pair sets bit-exact; double mode per-atom forces <= 1e-9
relative (to the largest force component of the system), energy/virial <= 1e-10 relative; mixed <= 1e-5.
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

TOL_F = {0: 1e-9, 1: 1e-5}
TOL_E = {0: 1e-10, 1: 1e-5}


def _setup(pkg, W, orc, case, prec, tables=False):
    u = None
    if case == "buck":
        s = W.fcc_system(8, 8, 8)
        co = W.coeffs_in_buck(2.5)
        style, ostyle, nt, ge = pkg.PAIR_BUCK, orc.BUCK, 1, 0.0
    elif case == "buck_big":
        s = W.fcc_system(9, 10, 10)
        co = W.coeffs_in_buck(5.0)
        style, ostyle, nt, ge = pkg.PAIR_BUCK, orc.BUCK, 1, 0.0
    elif case == "coul_cut":
        s = W.aC_system(1)
        co = W.coeffs_aC(10.0, 10.0)
        style, ostyle, nt, ge = pkg.PAIR_BUCK_COUL_CUT, orc.BUCK_COUL_CUT, 2, 0.0
    elif case == "coul_cut_split":   # different lj / coul cutoffs (three cutoffs, pair_buck_coul_cut_intel.cpp:275,295,319)
        s = W.aC_system(1)
        co = W.coeffs_aC(7.5, 10.0)
        style, ostyle, nt, ge = pkg.PAIR_BUCK_COUL_CUT, orc.BUCK_COUL_CUT, 2, 0.0
    elif case == "coul_long":
        s = W.aC_system(1)
        co = W.coeffs_aC(12.0, 12.0)
        style, ostyle, nt, ge = pkg.PAIR_BUCK_COUL_LONG, orc.BUCK_COUL_LONG, 2, 0.2776
    elif case == "coul_long_r2":
        s = W.aC_system(2)
        co = W.coeffs_aC(12.0, 12.0)
        style, ostyle, nt, ge = pkg.PAIR_BUCK_COUL_LONG, orc.BUCK_COUL_LONG, 2, 0.3051
    elif case == "lcl_disp":        # BASELINE config 5 variant: in.buck_big with `buck/long/coul/long long off 5.0`
        s = W.fcc_system(7, 7, 8)
        co = W.coeffs_in_buck(5.0)
        co["C"] = -co["C"]          # attractive dispersion, as a long-range r^-6 term needs
        style, ostyle, nt, ge, g6, o1, o6 = pkg.PAIR_BUCK_LONG_COUL_LONG, orc.BUCK_LONG_COUL_LONG, 1, 0.0, 0.85, 0, 1
    elif case == "lcl_both":        # long Coulomb + long dispersion on the charged aC cell
        s = W.aC_system(1)
        co = W.coeffs_aC(10.0, 10.0)
        style, ostyle, nt, ge, g6, o1, o6 = pkg.PAIR_BUCK_LONG_COUL_LONG, orc.BUCK_LONG_COUL_LONG, 2, 0.29, 0.33, 1, 1
    elif case == "lcl_coul":        # long Coulomb, cut dispersion (ORDER1 only)
        s = W.aC_system(1)
        co = W.coeffs_aC(10.0, 10.0)
        style, ostyle, nt, ge, g6, o1, o6 = pkg.PAIR_BUCK_LONG_COUL_LONG, orc.BUCK_LONG_COUL_LONG, 2, 0.29, 0.0, 1, 0
    else:
        raise ValueError(case)
    u = W.UNITS[s["units"]]
    lcl = style == pkg.PAIR_BUCK_LONG_COUL_LONG
    if not lcl:
        g6, o1, o6 = 0.0, 0, 0
    cutc = co.get("cut_coul", co["cut_lj"] if lcl else None)
    P = orc.Params(ostyle, nt, co["A"], co["rho"], co["C"], co["cut_lj"], cutc, qqrd2e=u["qqrd2e"],
                   g_ewald=ge, g_ewald_6=g6, order1=o1, order6=o6)
    cf = pkg.pair_coeffs(style, nt, co["A"], co["rho"], co["C"], co["cut_lj"], cutc)
    ct = dt = None
    if tables and (not lcl or o1):
        cc = float(cutc[1, 1])
        ct = pkg.init_coul_tables(cc, ge, u["qqrd2e"])
        P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
    if tables and lcl and o6:
        dt = pkg.init_disp_tables(float(co["cut_lj"][1, 1]), g6)
        P.set_disp_tables(dt[0], 12, dt[1], dt[2], dt[3])
    ctx = pkg.make_context(s, precision=prec)
    ctx.neigh_setup(0.3)
    ctx.pair_setup(style, nt, cf, g_ewald=ge, g_ewald_6=g6, ewald_order=(o1 << 1) | (o6 << 6), coul_tables=ct,
                   disp_tables=dt)
    return s, P, ctx


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("case", ["buck", "coul_cut", "coul_long"])
def test_pair_set_bit_exact(pkg, W, orc, case, prec):
    s, P, ctx = _setup(pkg, W, orc, case, prec)
    n = len(s["x"])
    ctx.neigh_build()
    nn, off, ent, gsrc, gshift = ctx.neigh_download()
    st = ctx.neigh_stats()
    assert off[-1] == st["total"] == nn.sum()
    assert nn.max() == st["max_numneigh"]
    gkeys = util.pair_keys(n, nn, ent, gsrc, gshift)
    assert len(np.unique(gkeys)) == len(gkeys), "duplicate entries in a row"
    # oracle: same ghosts rule, brute-force full list and LAMMPS-style binned half list
    cutneighmax = P.cutmax() + 0.3
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cutneighmax)
    assert len(src) == st["nghost"], "ghost count differs from Comm::borders"
    cns = P.cutneighsq(0.3)
    fn, foff, fent = orc.neigh_full_brute(n, xa, ta, P.ntypes, cns, prec)
    okeys = util.pair_keys(n, fn, fent, src, shift)
    assert np.array_equal(gkeys, okeys), "full-list pair set differs from the O(N^2) oracle"
    hn, hoff, hent = orc.neigh_half_bin(n, xa, ta, P.ntypes, cns, s["boxlo"], s["boxhi"], cutneighmax, prec)
    hkeys = util.pair_keys(n, hn, hent, src, shift, symmetrize=True)
    assert np.array_equal(gkeys, hkeys), "full list != symmetrised reference half list"
    assert 2 * hoff[-1] == st["total"]
    ctx.close()


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("case,tables", [("buck", False), ("buck_big", False), ("coul_cut", False),
                                         ("coul_cut_split", False), ("coul_long", False), ("coul_long", True),
                                         ("coul_long_r2", False), ("lcl_disp", False), ("lcl_disp", True),
                                         ("lcl_both", False), ("lcl_both", True), ("lcl_coul", False)])
def test_pair_forces_energy_virial(pkg, W, orc, case, tables, prec):
    s, P, ctx = _setup(pkg, W, orc, case, prec, tables)
    n = len(s["x"])
    ctx.neigh_build()
    ev = ctx.pair_compute(3, 1)
    d = ctx.atoms_download(("f", "eatom"))
    fo, evo, aux = orc.pair_forces_periodic(P, prec, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3,
                                            eflag=3, vflag=1, eatom=1)
    err = util.rel_force_err(d["f"], fo[:, :3])
    if prec == 0:
        assert err <= TOL_F[0], "force error %g" % err
    else:
        # Mixed mode stores positions in float (thr_pack), so a pair that straddles the periodic boundary has a
        # slightly different float distance seen from either side (float(x_k + L) - float(x_i) vs
        # float(x_k) - float(x_i - L)); exp(-r/rho) amplifies that ulp ~50x.  The newton-on half list evaluates such
        # a pair once, the device (newton off) from both sides — exactly like the reference's own NEWTON_PAIR=0
        # path, which is therefore the 1e-5 comparison; against the newton-on arm the image noise remains.
        fn_, evn = orc.pair_forces_periodic_newtoff(P, prec, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3,
                                                    eflag=3, vflag=1, eatom=1)
        errn = util.rel_force_err(d["f"], fn_[:, :3])
        fd, _, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3, 0, 0)
        print("mixed: gpu-vs-oracle(mixed,newton off) %.2e  gpu-vs-oracle(mixed,newton on) %.2e  "
              "oracle(mixed)-vs-oracle(double) %.2e" % (errn, err, util.rel_force_err(fo[:, :3], fd[:, :3])))
        assert errn <= TOL_F[1], "force error %g" % errn
        assert err <= 5e-5
        assert np.abs(evn[:2] - evo[:2]).max() <= 1e-6 * np.abs(evo[:2]).max()
    escale = max(abs(evo[0]), abs(evo[1]))
    assert abs(ev[0] - evo[0]) <= TOL_E[prec] * escale
    assert abs(ev[1] - evo[1]) <= TOL_E[prec] * escale
    vscale = np.abs(evo[2:]).max()
    assert np.abs(ev[2:] - evo[2:]).max() <= TOL_E[prec] * vscale
    # per-atom energy (f[].w)
    assert np.abs(d["eatom"] - fo[:, 3]).max() <= TOL_F[prec] * np.abs(fo[:, 3]).max()
    assert abs(d["eatom"].sum() - (evo[0] + evo[1])) <= 10 * TOL_E[prec] * escale
    # f.r virial of the reference (vflag 2, newton on) is the same tensor to rounding
    if prec == 0:
        _, evo2, _ = orc.pair_forces_periodic(P, prec, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3,
                                              eflag=1, vflag=2)
        assert np.abs(ev[2:] - evo2[2:]).max() <= 1e-8 * vscale
    # the EVFLAG=0 instantiation gives the same forces (different FMA contraction allowed: 1e-13)
    ctx.pair_compute(0, 0)
    f0 = ctx.atoms_download(("f",))["f"]
    assert util.rel_force_err(f0, d["f"]) <= (1e-13 if prec == 0 else 1e-6)
    # and is itself bitwise reproducible
    ctx.pair_compute(0, 0)
    assert np.array_equal(f0, ctx.atoms_download(("f",))["f"])
    # Newton's third law on a full list: net force vanishes
    assert np.abs(d["f"].sum(0)).max() <= 1e-6 * np.abs(d["f"]).max() * (1 if prec == 0 else 1e3)
    ctx.close()


def test_deterministic_rerun(pkg, W, orc):
    """no FP atomics anywhere: two independent contexts give bit-identical forces and tallies"""
    out = []
    for _ in range(2):
        s, P, ctx = _setup(pkg, W, orc, "coul_long", 0)
        ctx.neigh_build()
        ev = ctx.pair_compute(1, 1)
        out.append((ctx.atoms_download(("f",))["f"], ev))
        ctx.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("style", ["coul_long", "coul_cut", "buck"])
def test_eval_host_with_special_bonds(pkg, W, orc, style, prec):
    """the reference eval<> signature with host buffers, including special-bond bits in the list
    (jlist[jj] >> SBBITS & 3, pair_buck_intel.cpp:246-247) — oracle list + random special flags"""
    s, P, ctx = _setup(pkg, W, orc, style, prec, tables=False)
    n = len(s["x"])
    cutneighmax = P.cutmax() + 0.3
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cutneighmax)
    fn, foff, fent = orc.neigh_full_brute(n, xa, ta, P.ntypes, P.cutneighsq(0.3), prec)
    rng = np.random.default_rng(5)
    # symmetric special flags keyed on the unordered owner pair so both directions agree
    own, code, _ = util.owner_and_shift(fent, n, src, shift)
    i = np.repeat(np.arange(n), fn)
    lo, hi = np.minimum(i, own), np.maximum(i, own)
    sb = ((lo * 7919 + hi * 104729) % 11)
    sb = np.where(sb < 3, sb + 1, 0).astype(np.int64)
    ent_sb = (fent.astype(np.int64) | (sb << 30)).astype(np.uint32).view(np.int32)
    sl, sc = (1.0, 0.0, 0.5, 0.25), (1.0, 0.0, 0.3, 0.8)
    for k in range(4):
        P.p.special_lj[k] = sl[k]
        P.p.special_coul[k] = sc[k]
    # re-run setup with the special factors
    co = dict(buck=W.coeffs_in_buck(2.5), coul_cut=W.coeffs_aC(10.0, 10.0), coul_long=W.coeffs_aC(12.0, 12.0))[style]
    st = dict(buck=pkg.PAIR_BUCK, coul_cut=pkg.PAIR_BUCK_COUL_CUT, coul_long=pkg.PAIR_BUCK_COUL_LONG)[style]
    cf = pkg.pair_coeffs(st, P.ntypes, co["A"], co["rho"], co["C"], co["cut_lj"], co.get("cut_coul"))
    ctx.pair_setup(st, P.ntypes, cf, special_lj=sl, special_coul=sc, g_ewald=P.p.g_ewald)
    f, ev = ctx.pair_eval_host(1, 1, n, xa, ta, qa, fn, foff[:-1], ent_sb)
    # oracle: newton-off eval on the same full list double counts f[j]; use its half: evaluate with
    # newton=0 on a list restricted to j>i-equivalent entries is awkward, so compare per-atom instead:
    # reference NEWTON_PAIR=0 semantics on a FULL list = only f[i] is meaningful if f[j] updates are dropped.
    # The oracle accumulates f[j] too; subtract by evaluating the transpose: total = 2 * f_i, energies = 2x.
    fo, evo = orc.pair_eval(P, prec, 1, 1, n, xa, ta, qa, fn, foff, ent_sb, newton=0)
    # owned-owned pairs are double counted, owned-ghost are not: rebuild the oracle answer from a list
    # where every entry is made to look like a ghost interaction (j >= nlocal): shift indices
    xa2 = np.concatenate([xa[:n], xa]); ta2 = np.concatenate([ta[:n], ta]); qa2 = np.concatenate([qa[:n], qa])
    ent2 = ((fent.astype(np.int64) + n) | (sb << 30)).astype(np.uint32).view(np.int32)
    fo, evo = orc.pair_eval(P, prec, 1, 1, n, xa2, ta2, qa2, fn, foff, ent2, newton=0)
    err = util.rel_force_err(f[:, :3], fo[:n, :3])
    assert err <= TOL_F[prec], "force error %g" % err
    escale = max(abs(evo[0]), abs(evo[1]), 1e-300)
    assert abs(ev[0] - evo[0]) <= TOL_E[prec] * escale and abs(ev[1] - evo[1]) <= TOL_E[prec] * escale
    assert np.abs(ev[2:] - evo[2:]).max() <= TOL_E[prec] * np.abs(evo[2:]).max()
    ctx.close()


def test_list_row_layouts_agree(pkg, W, orc, monkeypatch):
    """the device list uses fixed-pitch, 128 B-aligned rows, or packed CSR rows when the padding would cost more than
    half the list (forced here with B200MD_LIST_CSR=1): same rows, bit-identical forces and tallies"""
    out = []
    for csr in ("0", "1"):
        monkeypatch.setenv("B200MD_LIST_CSR", csr)
        s, P, ctx = _setup(pkg, W, orc, "coul_long", 0)
        ctx.neigh_build()
        ev = ctx.pair_compute(1, 1)
        out.append((ctx.neigh_download(), ctx.atoms_download(("f",))["f"], ev))
        ctx.close()
    (l0, f0, e0), (l1, f1, e1) = out
    for a, b in zip(l0, l1):     # numneigh, CSR offsets, entries, ghost sources and shifts of the exported list
        assert np.array_equal(a, b)
    assert l0[0].sum() > 0 and np.array_equal(f0, f1) and np.array_equal(np.asarray(e0), np.asarray(e1))


def test_nve_bit_exact(pkg, W, orc):
    """fix nve/intel: x and v after initial/final integrate are bit-identical to the oracle given the same
    forces (un-fused mul+add, fix_nve_intel.cpp:74-77,116-117)"""
    s, P, ctx = _setup(pkg, W, orc, "coul_cut", 0)
    u = W.UNITS["metal"]
    ctx.nve_setup(u["dt"])
    ctx.neigh_build()
    ctx.pair_compute(0, 0)
    d0 = ctx.atoms_download(("x", "v", "f"))
    dtfm = orc.nve_dtfm(s["type"], s["mass"], u["dt"], u["ftm2v"])
    ctx.nve_initial_integrate()
    d1 = ctx.atoms_download(("x", "v"))
    xo, vo = orc.nve_initial(d0["x"], d0["v"], d0["f"], dtfm, u["dt"])
    assert np.array_equal(d1["x"], xo) and np.array_equal(d1["v"], vo)
    ctx.nve_final_integrate()
    d2 = ctx.atoms_download(("v",))
    vo2 = orc.nve_final(vo, d0["f"], dtfm)
    assert np.array_equal(d2["v"], vo2)
    ctx.close()


def test_nve_group_and_rmass_bit_exact(pkg, W, orc):
    """fix nve/intel on a sub-group with per-atom masses (fix_nve_intel.cpp:88-97, 147-190): atoms outside the
    group keep x and v bit for bit, the others match the oracle bit for bit; the kinetic energy uses rmass"""
    s, P, ctx = _setup(pkg, W, orc, "coul_cut", 0)
    u = W.UNITS["metal"]
    n = len(s["type"])
    rng = np.random.default_rng(5)
    ingroup = (rng.random(n) < 0.6).astype(np.int32)
    rmass = rng.uniform(5.0, 30.0, n)
    ctx.nve_set_group(ingroup, rmass)
    ctx.nve_setup(u["dt"])
    ctx.neigh_build()            # the sort must carry the per-atom dtfm along
    ctx.pair_compute(0, 0)
    d0 = ctx.atoms_download(("x", "v", "f"))
    dtfm = orc.nve_dtfm_group(s["type"], s["mass"], u["dt"], u["ftm2v"], ingroup, rmass)
    ctx.nve_initial_integrate()
    d1 = ctx.atoms_download(("x", "v"))
    xo, vo = orc.nve_initial_group(d0["x"], d0["v"], d0["f"], dtfm, u["dt"])
    assert np.array_equal(d1["x"], xo) and np.array_equal(d1["v"], vo)
    frozen = ingroup == 0
    assert frozen.any() and np.array_equal(d1["x"][frozen], d0["x"][frozen])
    assert not np.array_equal(d1["x"][~frozen], d0["x"][~frozen])
    ctx.nve_final_integrate()
    d2 = ctx.atoms_download(("v",))
    assert np.array_equal(d2["v"], orc.nve_final(vo, d0["f"], dtfm))
    # group all again, per-type masses: the plain path is back
    ctx.nve_set_group(None, None)
    ctx.nve_setup(u["dt"])
    ctx.nve_initial_integrate()
    d3 = ctx.atoms_download(("x", "v"))
    dt0 = orc.nve_dtfm(s["type"], s["mass"], u["dt"], u["ftm2v"])
    xo3, vo3 = orc.nve_initial(d1["x"], d2["v"], d0["f"], dt0, u["dt"])
    assert np.array_equal(d3["x"], xo3) and np.array_equal(d3["v"], vo3)
    ctx.close()


def test_run_energy_conservation_and_rebuilds(pkg, W, orc):
    """in.buck semantics: 100 NVE steps, rebuild every 20 steps without check (neigh_modify delay 0 every 20
    check no): total energy drift stays small and exactly 5 rebuilds happen"""
    s = W.fcc_system(8, 8, 8, jitter=0.0)
    co = W.coeffs_in_buck(2.5)
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3, every=20, delay=0, check=0)
    ctx.pair_setup(pkg.PAIR_BUCK, 1, cf)
    ctx.nve_setup(0.005)
    th0 = ctx.setup_forces(1, 1)
    e0 = th0[0] + th0[15]
    th1 = ctx.run(100, thermo=True)
    e1 = th1[0] + th1[15]
    st = ctx.neigh_stats()
    assert st["nbuilds"] == 1 + 5
    n = len(s["x"])
    assert abs(e1 - e0) / n < 2e-3, (e0, e1)
    assert th1[15] > 0
    # positions are host-ordered and inside the periodic images
    x = ctx.atoms_download(("x",))["x"]
    assert np.isfinite(x).all()
    ctx.close()


def test_trajectory_matches_oracle_steps(pkg, W, orc):
    """5 velocity-Verlet steps with check-yes reneighbouring: positions track the CPU oracle to 1e-11 rel"""
    s, P, ctx = _setup(pkg, W, orc, "coul_cut", 0)
    u = W.UNITS["metal"]
    dt = u["dt"]
    ctx.nve_setup(dt)
    ctx.setup_forces(0, 0)
    ctx.run(5)
    d = ctx.atoms_download(("x", "v"))
    # oracle loop
    x, v = s["x"].copy(), s["v"].copy()
    dtfm = orc.nve_dtfm(s["type"], s["mass"], dt, u["ftm2v"])
    f, _, _ = orc.pair_forces_periodic(P, 0, x, s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3, 0, 0)
    for _ in range(5):
        x, v = orc.nve_initial(x, v, f[:, :3].copy(), dtfm, dt)
        f, _, _ = orc.pair_forces_periodic(P, 0, W.wrap(x, s["boxlo"], s["boxhi"]), s["type"], s["q"], s["boxlo"],
                                           s["boxhi"], 0.3, 0, 0)
        v = orc.nve_final(v, f[:, :3].copy(), dtfm)
    prd = s["boxhi"] - s["boxlo"]
    dx = d["x"] - x
    dx -= np.round(dx / prd) * prd
    assert np.abs(dx).max() <= 1e-11 * prd.max()
    assert np.abs(d["v"] - v).max() <= 1e-9 * np.abs(v).max()
    ctx.close()


def test_error_paths(pkg, W):
    s = W.fcc_system(4, 4, 4)
    ctx = pkg.make_context(s)
    with pytest.raises(pkg.B200MDError):
        ctx.pair_compute(0, 0)           # before pair_setup
    co = W.coeffs_in_buck(2.5)
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK, 1, co["A"], co["rho"], co["C"], np.full((2, 2), 7.0))
    ctx.neigh_setup(0.3)
    ctx.pair_setup(pkg.PAIR_BUCK, 1, cf)
    with pytest.raises(pkg.B200MDError) as ei:   # box (6.7) shorter than the ghost cutoff (7.3)
        ctx.neigh_build()
    assert "ghost cutoff" in str(ei.value)
    ctx.close()
    with pytest.raises(pkg.B200MDError):
        pkg.Context(0, 7)               # PREC_MODE_SINGLE etc. are not provided


def test_step_host_matches_resident_run(pkg, W, orc):
    """b200md_step_host (host buffers in and out every step, the plug-in deployment / bench.py's e2e leg) walks the
    same trajectory, bit for bit, as the resident b200md_run"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    co = W.coeffs_aC(8.0, 8.0)
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])

    def make():
        c = pkg.make_context(s)
        c.neigh_setup(0.3)
        c.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=0.3)
        c.pppm_setup(24, 24, 27, 5, 0.3)
        c.nve_setup(u["dt"])
        c.setup_forces(0, 0)
        return c

    a, b = make(), make()
    a.run(4)
    da = a.atoms_download(("x", "f"))
    n = len(s["x"])
    xin = np.ascontiguousarray(b.atoms_download(("x",))["x"])
    xout, fout = np.zeros((n, 3)), np.zeros((n, 3))
    for _ in range(4):
        b.step_host(xin, xout, fout)
        xin, xout = xout, xin
    assert np.array_equal(xin, da["x"]) and np.array_equal(fout, da["f"])
    # x_in = NULL keeps the device positions; outputs may be NULL
    b.step_host(None, None, fout)
    a.run(1)
    assert np.array_equal(fout, a.atoms_download(("f",))["f"])
    a.close(); b.close()


def test_edge_cases_empty_ragged_nonperiodic(pkg, W, orc):
    """empty system; atoms far outside the box (Domain::pbc wraps them); a very inhomogeneous system (empty bins, one
    crowded bin, an isolated atom with no neighbours); a non-periodic box (no ghosts, FP64 test for every candidate)"""
    co = W.coeffs_in_buck(2.5)
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
    P = orc.Params(orc.BUCK, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
    lo, hi = np.zeros(3), np.array([14.0, 15.0, 16.0])
    mass = np.array([0.0, 1.0])
    # --- empty
    ctx = pkg.Context()
    ctx.set_units(1.0, 1.0); ctx.set_box(lo, hi)
    ctx.atoms_upload(np.zeros((0, 3)), np.zeros(0, np.int32), mass)
    ctx.neigh_setup(0.3); ctx.pair_setup(pkg.PAIR_BUCK, 1, cf); ctx.nve_setup(0.005)
    th = ctx.setup_forces(1, 1)
    assert not th.any() and ctx.neigh_stats()["total"] == 0
    ctx.run(3)
    ctx.close()
    # --- ragged: a dense cluster, a sparse gas, one isolated atom; some coordinates several box lengths away
    rng = np.random.default_rng(5)
    cluster = np.array([7.0, 7.5, 8.0]) + rng.uniform(-1.4, 1.4, (300, 3))
    gas = rng.uniform(0, 1, (60, 3)) * (hi - lo) * np.array([1.0, 1.0, 0.3])
    x = np.concatenate([cluster, gas, [[0.7, 14.2, 15.1]]])
    # keep every pair farther apart than 0.75 (the r^-6 term of in.buck's C = -0.8 is repulsive: no overflow, but keep
    # forces in a sane range for the relative comparison)
    keep = np.ones(len(x), bool)
    for i in range(len(x)):
        if keep[i]:
            d = x[i + 1:] - x[i]
            d -= np.round(d / (hi - lo)) * (hi - lo)
            keep[i + 1:] &= (d * d).sum(1) > 0.75 ** 2
    x = x[keep]
    t = np.ones(len(x), np.int32)
    xs = x.copy()
    xs[::7] += np.array([3, -2, 1]) * (hi - lo)       # far outside: remapped by the first build only within one image,
    xs[::7] -= np.array([2, -3, 1]) * (hi - lo)       # so bring them back to +-1 box length like a real trajectory
    ctx = pkg.Context()
    ctx.set_units(1.0, 1.0); ctx.set_box(lo, hi)
    ctx.atoms_upload(xs, t, mass)
    ctx.neigh_setup(0.3); ctx.pair_setup(pkg.PAIR_BUCK, 1, cf)
    ctx.neigh_build()
    ev = ctx.pair_compute(1, 1)
    f = ctx.atoms_download(("f",))["f"]
    xw = W.wrap(W.wrap(xs, lo, hi), lo, hi)
    fo, evo, _ = orc.pair_forces_periodic(P, 0, xw, t, None, lo, hi, 0.3)
    assert util.rel_force_err(f, fo[:, :3]) <= 1e-9
    assert abs(ev[0] - evo[0]) <= 1e-10 * abs(evo[0])
    nn = ctx.neigh_download()[0]
    assert nn.min() == 0 and nn.max() > 20, "expected an isolated atom and a crowded bin"
    ctx.close()
    # --- non-periodic box: no ghosts at all, pair set = brute force over owned atoms
    ctx = pkg.Context()
    ctx.set_units(1.0, 1.0); ctx.set_box(lo, hi, periodic=(0, 0, 0))
    ctx.atoms_upload(x, t, mass)
    ctx.neigh_setup(0.3); ctx.pair_setup(pkg.PAIR_BUCK, 1, cf)
    ctx.neigh_build()
    st = ctx.neigh_stats()
    assert st["nghost"] == 0
    nn, off, ent, gsrc, gshift = ctx.neigh_download()
    fn, foff, fent = orc.neigh_full_brute(len(x), x, t, 1, P.cutneighsq(0.3), 0)
    assert np.array_equal(nn, fn)
    for i in (0, len(x) // 2, len(x) - 1):
        assert np.array_equal(np.sort(ent[off[i]:off[i + 1]] & pkg.NEIGHMASK), np.sort(fent[foff[i]:foff[i + 1]] & pkg.NEIGHMASK))
    ctx.close()


def test_table_inner_cutoff_and_list_flavours_agree(pkg, W, orc):
    """Coulomb tables on a device-built list with a pair inside the tables' inner cut-off (r < sqrt 2: the reference takes
    the analytic branch there, pair_buck_coul_long_intel.cpp:294) against the oracle, and the packed-type flavour of the
    kernel against the type-gathering flavour used for host-supplied lists (same pairs)"""
    s = W.aC_system(1)
    x = s["x"].copy()
    # bring one oxygen to 1.25 A of another atom (inside tabinner = sqrt 2)
    d = x[1] - x[0]
    d -= np.round(d / (s["boxhi"] - s["boxlo"])) * (s["boxhi"] - s["boxlo"])
    x[1] = x[0] + d / np.linalg.norm(d) * 1.25
    s = dict(s, x=W.wrap(x, s["boxlo"], s["boxhi"]))
    u = W.UNITS["metal"]
    co = W.coeffs_aC(12.0, 12.0)
    ge = 0.2776
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=ge)
    ct = pkg.init_coul_tables(12.0, ge, u["qqrd2e"])
    P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3)
    ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=ge, coul_tables=ct)
    ctx.neigh_build()
    ev = ctx.pair_compute(1, 1)
    f = ctx.atoms_download(("f",))["f"]
    fo, evo, aux = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    assert util.rel_force_err(f, fo[:, :3]) <= 1e-9
    assert abs(ev[1] - evo[1]) <= 1e-10 * abs(evo[1]) and abs(ev[0] - evo[0]) <= 1e-10 * max(abs(evo[0]), abs(evo[1]))
    # the generic flavour (host-supplied list of the same pairs, type gathered, branches as the reference has them)
    nn, off, ent, gsrc, gshift = ctx.neigh_download()
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], P.cutmax() + 0.3)
    hn, hoff, hent = orc.neigh_full_brute(len(x), xa, ta, 2, P.cutneighsq(0.3), 0)
    fg, evg = ctx.pair_eval_host(1, 1, len(x), xa, ta, qa, hn, hoff[:-1], hent)
    assert util.rel_force_err(f, fg[:, :3]) <= 1e-12
    assert abs(ev[1] - evg[1]) <= 1e-12 * abs(evg[1])
    ctx.close()


def test_setup_guards_and_host_list_validation(pkg, W, orc):
    """errors instead of silent garbage (round-1 advisor findings): exponent arguments outside fast_exp's range are
    refused at setup; b200md_pair_eval_host checks the caller's types and list entries; a k-space state is dropped by a
    later b200md_neigh_setup with a larger skin (its brick halo is sized from skin/2)"""
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    co = W.coeffs_aC(12.0, 12.0)
    cf = pkg.pair_coeffs(pkg.PAIR_BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    ctx = pkg.make_context(s)
    ctx.neigh_setup(0.3)
    with pytest.raises(pkg.B200MDError, match="g_ewald"):
        ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=3.0)          # (3 * 12)^2 > 700
    bad = {k: v.copy() for k, v in cf.items()}
    bad["rhoinv"][1, 2] = bad["rhoinv"][2, 1] = -1.0
    with pytest.raises(pkg.B200MDError, match="must be positive"):
        ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, bad, g_ewald=0.28)
    cfl = pkg.pair_coeffs(pkg.PAIR_BUCK_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"])
    with pytest.raises(pkg.B200MDError, match="g_ewald_6"):
        ctx.pair_setup(pkg.PAIR_BUCK_LONG_COUL_LONG, 2, cfl, g_ewald=0.28, g_ewald_6=2.5, ewald_order=(1 << 1) | (1 << 6))
    ctx.pair_setup(pkg.PAIR_BUCK_COUL_LONG, 2, cf, g_ewald=0.28)
    # host-supplied list: a type outside 1..ntypes, an entry past nall
    n = len(s["x"])
    nn = np.ones(n, np.int32)
    off = np.arange(n, dtype=np.int64)
    ent = ((np.arange(n) + 1) % n).astype(np.int32)
    t_bad = s["type"].copy()
    t_bad[5] = 3
    with pytest.raises(pkg.B200MDError, match="type"):
        ctx.pair_eval_host(0, 0, n, s["x"], t_bad, s["q"], nn, off, ent)
    e_bad = ent.copy()
    e_bad[7] = n + 3
    with pytest.raises(pkg.B200MDError, match="past nall"):
        ctx.pair_eval_host(0, 0, n, s["x"], s["type"], s["q"], nn, off, e_bad)
    f, ev = ctx.pair_eval_host(0, 0, n, s["x"], s["type"], s["q"], nn, off, ent)
    assert np.isfinite(f).all()
    # stale k-space state
    ctx.pppm_setup(24, 24, 27, 5, 0.28)
    ctx.neigh_setup(0.2)                      # smaller skin: the halo still covers it, the state stays
    ctx.setup_forces(0, 0)
    ctx.neigh_setup(1.0)                      # larger skin: the state is dropped
    with pytest.raises(pkg.B200MDError, match="before b200md_pppm_setup"):
        ctx.pppm_compute(0, 0)
    # an empty system exports offsets[0] = 0
    ctx.close()
