"""Host layer (lammps-buck-intel_b200/host): the reference's class surface + the input-script driver.
CPU part: parsing, lattice / read_data / replicate, velocity, init_one products and PPPM sizing via `lmp_b200
-dry-run` (no device).  GPU part (-m gpu): the scripts run end to end through the classes and the C ABI and are
compared with the CPU oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

import scripts


def _run(pkg, args, cwd=None):
    lmp = pkg.build_host()
    r = subprocess.run([lmp] + args, capture_output=True, text=True, cwd=cwd, timeout=600)
    return r


def _summary(out):
    line = [l for l in out.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def _thermo(out):
    rows, on = [], False
    for l in out.splitlines():
        w = l.split()
        if w[:2] == ["Step", "Temp"]:
            on = True
            continue
        if on:
            try:
                rows.append([float(v) for v in w])
            except ValueError:
                on = False
    return np.array(rows)


def test_dry_run_in_buck(pkg, W, tmp_path):
    p = scripts.write(tmp_path, "in.buck", scripts.IN_BUCK.format(n=20, steps=100, thermo=0))
    r = _run(pkg, ["-in", p, "-sf", "intel", "-dry-run"])
    assert r.returncode == 0, r.stdout + r.stderr
    s = _summary(r.stdout)
    assert s["natoms"] == 32000 and s["pair_style"] == "buck" and s["cutforce"] == 2.5
    assert (s["every"], s["delay"], s["check"]) == (20, 0, 0)
    assert s["temperature"] == pytest.approx(1.44, rel=1e-12)
    assert np.allclose(s["box"], 20 * (4 / 0.8442) ** (1 / 3))
    # -var overrides the index variable (the script's own scaling mechanism)
    r = _run(pkg, ["-in", p, "-sf", "intel", "-dry-run", "-var", "x", "2"])
    assert _summary(r.stdout)["natoms"] == 64000


def test_dry_run_coul_long_sizing_matches_python(pkg, W, tmp_path):
    """PPPM::set_grid_global / adjust_gewald in C++ (host/pppm_intel.cpp) == the Python restatement used by bench.py"""
    for acc, r in ((1e-4, 2), (1e-6, 2), (1e-4, 3)):
        txt = scripts.IN_BUCK_COUL_LONG.format(r=r, kspace="pppm %g" % acc, pair_modify="", steps=1, thermo=0)
        p = scripts.write(tmp_path, "in.bcl", txt, W)
        out = _run(pkg, ["-in", p, "-sf", "intel", "-dry-run"])
        assert out.returncode == 0, out.stdout + out.stderr
        s = _summary(out.stdout)
        sysd = W.aC_system(r, jitter=0.0)
        u = W.UNITS["metal"]
        grid, g = pkg.pppm_init(acc, u["qqrd2e"], sysd["q"], len(sysd["x"]), 12.0, sysd["boxhi"] - sysd["boxlo"])
        assert tuple(s["grid"]) == tuple(grid)
        assert s["g_ewald"] == pytest.approx(g, rel=1e-9)
        assert s["natoms"] == 1200 * r ** 3
    # the shipped script says `kspace_style ewald 1e-6`: served by the mesh solver, with a warning
    txt = scripts.IN_BUCK_COUL_LONG.format(r=2, kspace="ewald 1e-6", pair_modify="", steps=1, thermo=0)
    out = _run(pkg, ["-in", scripts.write(tmp_path, "in.bcl", txt, W), "-sf", "intel", "-dry-run"])
    assert "ewald is not provided" in out.stderr and tuple(_summary(out.stdout)["grid"]) == (96, 96, 108)


def test_dry_run_slab_sizing_and_boundary_checks(pkg, W, tmp_path):
    """`boundary p p f` + `kspace_modify slab 3.0`: the z mesh and the error balance use zprd * 3 (PPPM::set_grid_global
    with zprd_slab), the same numbers as the Python restatement; wrong boundaries are the stock errors"""
    base = scripts.IN_BUCK_COUL_LONG.format(r=2, kspace="pppm 1e-4\nkspace_modify slab 3.0", pair_modify="", steps=1,
                                            thermo=0)
    txt = "boundary p p f\n" + base
    out = _run(pkg, ["-in", scripts.write(tmp_path, "in.slab", txt, W), "-sf", "intel", "-dry-run"])
    assert out.returncode == 0, out.stdout + out.stderr
    s = _summary(out.stdout)
    sysd = W.aC_system(2, jitter=0.0)
    u = W.UNITS["metal"]
    prd = sysd["boxhi"] - sysd["boxlo"]
    grid, g = pkg.pppm_init(1e-4, u["qqrd2e"], sysd["q"], len(sysd["x"]), 12.0, prd, slab=3.0)
    grid0, _ = pkg.pppm_init(1e-4, u["qqrd2e"], sysd["q"], len(sysd["x"]), 12.0, prd)
    assert tuple(s["grid"]) == tuple(grid) and grid[2] > 2 * grid0[2] and grid[:2] == grid0[:2]
    assert s["g_ewald"] == pytest.approx(g, rel=1e-9)
    out = _run(pkg, ["-in", scripts.write(tmp_path, "in.slab2", base, W), "-sf", "intel", "-dry-run"])
    assert out.returncode != 0 and "Incorrect boundaries with slab PPPM" in out.stdout + out.stderr
    noslab = txt.replace("kspace_modify slab 3.0\n", "")
    out = _run(pkg, ["-in", scripts.write(tmp_path, "in.slab3", noslab, W), "-sf", "intel", "-dry-run"])
    assert out.returncode != 0 and "Cannot use nonperiodic boundaries with PPPM" in out.stdout + out.stderr


def test_pppm_disp_function_selection(pkg, W, tmp_path):
    """PPPMDisp::init [UPSTREAM] as restated in PPPMDispIntel::init: order-6 dispersion is served by function[1]
    (geometric), function[2] (arithmetic: the pair style's ewald_mix) or function[3] (`kspace_modify mix/disp none`);
    `mix/disp geom` forces the geometric grid; the Buckingham style has no mixing rule to offer: geometric"""
    import json
    cases = [("", [1, 0, 1, 0], "arithmetic"), ("mix/disp none", [1, 0, 0, 1], "no mixing rule"),
             ("mix/disp geom", [1, 1, 0, 0], "geometric"), ("mix/disp pair", [1, 0, 1, 0], "arithmetic")]
    for mixdisp, fn, name in cases:
        txt = scripts.IN_LJ_DISP_MIX.format(mixdisp=mixdisp, steps=1, thermo=1)
        r = _run(pkg, ["-in", scripts.write(tmp_path, "in.ljmix", txt, W), "-sf", "intel", "-dry-run"])
        assert r.returncode == 0, r.stdout + r.stderr
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert d["disp_functions"] == fn and d["dispersion_grid"] == name and d["grid_6"] == [30, 30, 32], d
    # pair_modify mix geometric on the LJ style, and the Buckingham style of config 5
    txt = scripts.IN_LJ_DISP_MIX.format(mixdisp="", steps=1, thermo=1).replace("mix arithmetic", "mix geometric")
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.ljmix", txt, W), "-sf", "intel", "-dry-run"])
    assert json.loads(r.stdout.strip().splitlines()[-1])["disp_functions"] == [1, 1, 0, 0]
    txt = scripts.IN_BUCK_DISP.format(n=4, g6=0.9, m=30, steps=1, thermo=1, A=3000.0)
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.disp", txt), "-sf", "intel", "-dry-run"])
    assert json.loads(r.stdout.strip().splitlines()[-1])["disp_functions"] == [0, 1, 0, 0]
    # an unknown mix/disp keyword is an illegal kspace_modify
    txt = scripts.IN_LJ_DISP_MIX.format(mixdisp="mix/disp harmonic", steps=1, thermo=1)
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.ljmix", txt, W), "-sf", "intel", "-dry-run"])
    assert r.returncode != 0 and "Illegal kspace_modify command" in r.stdout + r.stderr


def test_dispersion_grid_components(pkg, orc):
    """csrc/pppm.cu disp_components (host side of b200md_pppm_setup): the signed self-coupled components a dispersion
    grid is run as reproduce the r^-6 coefficient of every type pair, sum_m sign_m W_m[i] W_m[j] = C_ij -
    geometric (one), arithmetic (seven: the coupled grids a_k, a_{6-k} rotated into p, m = (a_k +- a_{6-k}) / sqrt 2),
    none (cyclic-Jacobi eigen-split, checked against numpy's eigh)"""
    import ctypes as C
    lib = pkg.load()
    lib.b200md_debug_disp_components.restype = C.c_int
    dp = C.POINTER(C.c_double)

    def comps(mix, T, B):
        B = np.ascontiguousarray(B, dtype=np.float64)
        W = np.zeros(16 * (T + 1)); sg = np.zeros(16)
        n = lib.b200md_debug_disp_components(C.c_int(mix), C.c_int(T), B.ctypes.data_as(dp), W.ctypes.data_as(dp),
                                             sg.ctypes.data_as(dp))
        assert n > 0, n
        return W[:n * (T + 1)].reshape(n, T + 1), sg[:n]

    rng = np.random.default_rng(5)
    for T in (1, 2, 3, 5, 8):
        eps = np.concatenate([[0.0], rng.uniform(0.05, 2.0, T)])
        sig = np.concatenate([[0.0], rng.uniform(1.5, 4.0, T)])
        Bg = 2.0 * np.sqrt(eps) * sig ** 3
        W, sg = comps(1, T, Bg)
        assert W.shape == (1, T + 1) and np.array_equal(W[0], Bg) and sg[0] == 1.0
        B7 = orc.disp_B_arithmetic(eps, sig)
        W, sg = comps(2, T, B7)
        assert len(sg) == 7 and list(sg) == [1, 1, 1, 1, -1, -1, -1]
        Cl = 4.0 * np.sqrt(np.outer(eps, eps)) * ((sig[:, None] + sig[None, :]) / 2.0) ** 6
        Cw = np.einsum("m,mi,mj->ij", sg, W, W)
        assert np.allclose(Cw[1:, 1:], Cl[1:, 1:], rtol=1e-12, atol=1e-12 * Cl.max())
        # no mixing rule: a random symmetric (indefinite) matrix
        A = rng.normal(size=(T, T)); A = A + A.T
        Cn = np.zeros((T + 1, T + 1)); Cn[1:, 1:] = A
        W, sg = comps(3, T, Cn)
        Cw = np.einsum("m,mi,mj->ij", sg, W, W)
        assert np.allclose(Cw, Cn, rtol=0, atol=1e-12 * np.abs(A).max())
        lam = np.linalg.eigvalsh(A)
        assert sorted(np.round(sg * (W * W).sum(1), 9)) == sorted(np.round(lam, 9))   # |W_m|^2 sign_m = eigenvalue m
    # rank-deficient: one eigen-grid; asymmetric input is refused
    b = np.array([0.0, 9.0, 13.2])
    W, sg = comps(3, 2, np.outer(b, b))
    assert len(sg) == 1 and np.allclose(np.abs(W[0]), b, rtol=1e-12)
    bad = np.outer(b, b); bad[1, 2] += 1.0
    assert lib.b200md_debug_disp_components(C.c_int(3), C.c_int(2), bad.ctypes.data_as(dp), np.zeros(64).ctypes.data_as(dp),
                                            np.zeros(16).ctypes.data_as(dp)) < 0


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5, 6, 7])
def test_tiled_make_rho_plan(pkg, order):
    """host-side plan of the tiled charge assignment (csrc/pppm.cu: rho_lane_map, cover_table): every stencil-face
    point is owned by exactly one lane, the 16 lanes of a half warp hit 16 different 8-byte bank pairs, and the
    cover table of a dimension lists, for every grid point, exactly the tile blocks that contain it"""
    import ctypes as C
    lib = pkg.load()
    E = 8 + order - 1
    nlower, nupper = -((order - 1) // 2), order // 2
    for n in list(range(max(2 * order, 2), 100)) + [125, 128, 250, 270, 486]:
        pitch = C.c_int()
        lanes = (C.c_int * 64)()
        cover = (C.c_int * (4 * n))()
        assert lib.b200md_debug_rho_plan(order, n, C.byref(pitch), lanes, cover) == 0
        lanes = np.array(lanes[:])
        pts = lanes[lanes >= 0]
        assert sorted(pts) == list(range(order * order)) and pitch.value >= E
        for h in range(4):
            half = lanes[16 * h:16 * h + 16]
            half = half[half >= 0]
            res = ((half // order) * pitch.value + half % order) % 16
            assert len(set(res)) == len(res), "bank conflict in half warp %d: %s" % (h, res)
        cov = np.array(cover[:]).reshape(n, 4)
        nt = (n + 7) // 8
        for g in range(n):
            got = sorted((e >> 4, e & 15) for e in cov[g] if e >= 0)
            want = []
            for t in range(nt):
                lo, hi = 8 * t, min(8 * t + 8, n) - 1
                for shift in (-n, 0, n):          # the tile and its periodic images
                    if lo + shift + nlower <= g <= hi + shift + nupper:
                        want.append((t, g - (lo + shift) - nlower))
            assert got == sorted(want), (n, g, got, want)
            assert all(0 <= loc < E for _, loc in got)


REF_EXAMPLES = "/root/reference/examples"


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("script,natoms,style,cut", [("in.buck", 32000, "buck", 2.5), ("in.buck_big", 192000, "buck", 5.0),
                                                     ("in.buck_coul_cut", 76800, "buck/coul/cut", 10.0),
                                                     ("in.buck_coul_long", 9600, "buck/coul/long", 12.0)])
def test_shipped_example_scripts_parse_verbatim(pkg, script, natoms, style, cut):
    """the reference's own examples/in.buck* are accepted as they are (CPU-only check, in the container that has the
    reference mounted; the GPU tests run the same semantics from tests/scripts.py)"""
    r = subprocess.run([os.path.join(pkg.HERE, "lmp_b200"), "-in", script, "-sf", "intel", "-dry-run"], cwd=REF_EXAMPLES,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    s = _summary(r.stdout)
    assert (s["natoms"], s["pair_style"], s["cutforce"]) == (natoms, style, cut)


def test_styles_are_registered_like_the_reference_registers_them(pkg):
    """the headers of host/ carry the `#ifdef PAIR_CLASS  PairStyle(key,Class)  #else ... #endif` blocks of the reference's
    headers (pair_buck_intel.h:18-22 and siblings, pppm_intel.h:18-22, pppm_disp_intel.h:18-22) and the driver fills its
    style maps by including them with PAIR_CLASS / KSPACE_CLASS / FIX_CLASS defined, as stock Force / Modify do: the
    registered names (and, where the reference is mounted, the key -> class pairs) are the reference's"""
    import re
    r = _run(pkg, ["-styles"])
    assert r.returncode == 0
    got = sorted(tuple(l.split()) for l in r.stdout.splitlines())
    assert got == sorted([("pair", "buck/intel"), ("pair", "buck/coul/cut/intel"), ("pair", "buck/coul/long/intel"),
                          ("pair", "buck/long/coul/long/intel"), ("pair", "lj/long/coul/long/intel"),
                          ("kspace", "pppm/intel"), ("kspace", "pppm/disp/intel"), ("fix", "nve/intel")])
    pat = re.compile(r"^(Pair|KSpace|Fix)Style\(([^,]+),(\w+)\)", re.M)
    host = os.path.join(pkg.HERE, "host")
    mine = {m.groups() for f in os.listdir(host) if f.endswith(".h") for m in pat.finditer(open(os.path.join(host, f)).read())}
    assert {("%s" % k.lower(), key) for k, key, _ in mine} == set(got)
    if os.path.isdir("/root/reference"):
        ref = {m.groups() for f in os.listdir("/root/reference") if f.endswith(".h")
               for m in pat.finditer(open(os.path.join("/root/reference", f)).read())}
        assert len(ref) == 7 and ref <= mine          # the reference ships no header for fix nve/intel


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("unit,symbols", [
    ("pair_buck_intel.cpp", ["PairBuckIntel::init_style()", "PairBuckIntel::compute(int, int)", "b200md_pair_setup",
                             "b200md_pair_compute"]),
    ("pair_buck_coul_cut_intel.cpp", ["PairBuckCoulCutIntel::init_style()", "PairBuckCoulCutIntel::compute(int, int)",
                                      "b200md_pair_setup", "b200md_pair_compute"]),
    ("pair_buck_coul_long_intel.cpp", ["PairBuckCoulLongIntel::init_style()", "PairBuckCoulLongIntel::compute(int, int)",
                                       "b200md_pair_setup", "b200md_pair_compute"]),
    ("pair_buck_long_coul_long_intel.cpp", ["PairBuckLongCoulLongIntel::init_style()",
                                            "PairBuckLongCoulLongIntel::compute(int, int)", "b200md_pair_setup",
                                            "b200md_pair_compute"]),
    ("pair_lj_long_coul_long_intel.cpp", ["PairLJLongCoulLongIntel::init_style()",
                                          "PairLJLongCoulLongIntel::compute(int, int)", "b200md_pair_setup",
                                          "b200md_pair_compute"]),
    ("pppm_intel.cpp", ["PPPMIntel::init()", "PPPMIntel::compute(int, int)", "PPPMIntel::brick2fft()", "b200md_pppm_setup",
                        "b200md_pppm_compute"]),
    ("pppm_disp_intel.cpp", ["PPPMDispIntel::init()", "PPPMDispIntel::compute(int, int)", "b200md_pppm_setup",
                             "b200md_pppm_compute"])])
def test_integration_binding_compiles_against_the_reference_header(pkg, tmp_path, unit, symbols):
    """lammps-buck-intel_b200/integration/: the translation units a maintainer puts in place of the reference's five
    pair_*_intel.cpp, pppm_intel.cpp and pppm_disp_intel.cpp.  They implement the classes that the reference's OWN headers declare (included
    unchanged from /root/reference) through the C ABI, against the
    stand-ins of the stock LAMMPS headers that the reference's own sources compile against (oracle/ref_shim): g++ -Wall
    accepts them, every member the header declares is defined, and the only undefined b200md symbols are C-ABI entries
    of include/b200md.h"""
    root = os.path.dirname(pkg.HERE)
    obj = os.path.join(str(tmp_path), unit + ".o")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-Wall", "-Werror", "-Wno-unknown-pragmas", "-DINTEL_VMASK", "-DINTEL_ALLOW_TABLE",
           "-I", os.path.join(root, "oracle", "ref_shim"), "-I", "/root/reference", "-I", os.path.join(root, "include"),
           "-I", os.path.join(pkg.HERE, "integration"), "-c", os.path.join(pkg.HERE, "integration", unit), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    nm = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    defined = [l for l in nm.splitlines() if " T " in l]
    undefined = [l.split()[-1] for l in nm.splitlines() if " U " in l and "b200" in l]
    for sym in symbols:
        assert any(sym in l for l in (defined if "::" in sym else nm.splitlines())), sym
    header = open(os.path.join(root, "include", "b200md.h")).read()
    for u in undefined:
        assert u.startswith("b200md_") and (u + "(") in header or u.startswith("LAMMPS_NS::b200_"), u


def test_driver_errors(pkg, W, tmp_path):
    bad = scripts.IN_BUCK.format(n=4, steps=1, thermo=0).replace("pair_coeff 1 1 1.0 0.2 -0.8", "")
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.bad", bad), "-sf", "intel", "-dry-run"])
    assert r.returncode == 1 and "All pair coeffs are not set" in r.stdout
    nosf = scripts.IN_BUCK.format(n=4, steps=1, thermo=0)
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.nosf", nosf), "-dry-run"])
    assert r.returncode == 1 and "only the /intel styles are provided" in r.stdout
    cutc = scripts.IN_BUCK_DISP.format(n=4, g6=0.9, m=16, steps=1, thermo=0, A=3000.0).replace("long off", "long cut")
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.cut", cutc), "-sf", "intel", "-dry-run"])
    assert r.returncode == 1 and "Coulomb cut not supported" in r.stdout


@pytest.mark.gpu
def test_in_buck_runs_like_the_oracle(pkg, W, orc, tmp_path):
    """in.buck at 8^3 cells: step-0 E_pair / pressure equal the oracle on the same lattice; 100 NVE steps with the
    script's `every 20 check no` cadence conserve energy and do exactly 5 rebuilds"""
    p = scripts.write(tmp_path, "in.buck", scripts.IN_BUCK.format(n=8, steps=100, thermo=20))
    r = _run(pkg, ["-in", p, "-sf", "intel", "-pk", "intel", "0", "mode", "double"])
    assert r.returncode == 0, r.stdout + r.stderr
    th = _thermo(r.stdout)
    assert th.shape[0] == 6 and th[0, 0] == 0 and th[-1, 0] == 100
    s = W.fcc_system(8, 8, 8, jitter=0.0)
    co = W.coeffs_in_buck(2.5)
    P = orc.Params(orc.BUCK, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], None, s["boxlo"], s["boxhi"], 0.3)
    n = len(s["x"])
    assert th[0, 1] == pytest.approx(1.44, rel=1e-9)
    assert th[0, 2] == pytest.approx(ev[0], rel=1e-10)
    vol = np.prod(s["boxhi"] - s["boxlo"])
    press = ((3 * n - 3) * 1.44 + ev[2] + ev[3] + ev[4]) / 3.0 / vol
    assert th[0, 5] == pytest.approx(press, rel=1e-8)
    drift = np.abs(th[:, 4] - th[0, 4]).max() / n
    assert drift < 3e-3, "total energy per atom drifts by %g" % drift   # un-shifted cut-off at 2.5 sigma, T* = 1.44
    assert "Neighbor list builds = 5" in r.stdout


@pytest.mark.gpu
def test_fix_nve_on_a_group_freezes_the_rest(pkg, W, tmp_path):
    """`group mobile id 1:1000` + `fix 1 mobile nve`: FixNVEIntel's mask & groupbit branch (fix_nve_intel.cpp:88-97,
    173-190) — atoms outside the group end the run exactly where create_atoms put them, the others have moved"""
    dump0, dump1 = str(tmp_path / "x0.xyz"), str(tmp_path / "x1.xyz")
    txt = scripts.IN_BUCK.format(n=8, steps=40, thermo=0)
    txt = txt.replace("fix 1 all nve", "group mobile id 1:1000\nfix 1 mobile nve")
    txt = txt.replace("run 40", "write_dump all xyz %s\nrun 40\nwrite_dump all xyz %s" % (dump0, dump1))
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.group", txt), "-sf", "intel"])
    assert r.returncode == 0, r.stdout + r.stderr
    x0 = np.loadtxt(dump0, skiprows=2)[:, 1:]
    x1 = np.loadtxt(dump1, skiprows=2)[:, 1:]
    assert x0.shape == (2048, 3)
    assert np.array_equal(x1[1000:], x0[1000:])
    assert np.abs(x1[:1000] - x0[:1000]).max() > 1e-3
    bad = txt.replace("fix 1 mobile nve", "fix 1 nobody nve")
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.nogroup", bad), "-sf", "intel", "-dry-run"])
    assert r.returncode != 0 and "Could not find fix group ID" in r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("table", [0, 12])
def test_in_buck_coul_long_runs_like_the_oracle(pkg, W, orc, tmp_path, table):
    """in.buck_coul_long (data.aC x 2^3) with pppm 1e-4, analytic erfc and the stock Coulomb tables: step-0 energies
    against the oracle (pair + PPPM) on the same inputs; resident and host-step deployments agree"""
    txt = scripts.IN_BUCK_COUL_LONG.format(r=2, kspace="pppm 1e-4", pair_modify="pair_modify table %d" % table,
                                           steps=10, thermo=5)
    p = scripts.write(tmp_path, "in.bcl", txt, W)
    r = _run(pkg, ["-in", p, "-sf", "intel"])
    assert r.returncode == 0, r.stdout + r.stderr
    th = _thermo(r.stdout)
    s = W.aC_system(2, jitter=0.0)
    u = W.UNITS["metal"]
    grid, g = pkg.pppm_init(1e-4, u["qqrd2e"], s["q"], len(s["x"]), 12.0, s["boxhi"] - s["boxlo"])
    co = W.coeffs_aC(12.0, 12.0)
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=g)
    if table:
        ct = pkg.init_coul_tables(12.0, g, u["qqrd2e"])
        P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    pp = orc.PPPM(*grid, 5, g, s["boxlo"], s["boxhi"], u["qqrd2e"])
    fk, ek, vk = pp.compute(s["x"], s["q"])
    assert th[0, 2] == pytest.approx(ev[0] + ev[1] + ek, rel=1e-9)
    assert th[0, 1] == pytest.approx(300.0, rel=1e-9)
    n = len(s["x"])
    # Verlet fluctuation of the stiff Si-O modes at dt = 1 fs is ~(w dt)^2/8 of the kinetic energy (0.04 eV/atom)
    assert np.abs(th[:, 4] - th[0, 4]).max() / n < 1e-3
    # plug-in deployment (host owns x and f, PCIe every step) walks the same trajectory
    r2 = _run(pkg, ["-in", p, "-sf", "intel", "-host-step"])
    assert r2.returncode == 0, r2.stdout + r2.stderr
    th2 = _thermo(r2.stdout)
    assert np.allclose(th2[:, 2], th[:, 2], rtol=1e-9, atol=0)


@pytest.mark.gpu
def test_buck_long_coul_long_with_pppm_disp(pkg, W, orc, tmp_path):
    """BASELINE config 5 variant: buck/long/coul/long long off + pppm/disp (geometric grid) through the driver:
    step-0 E_pair against the oracle's pair + dispersion-PPPM, energy conservation over 20 steps"""
    n, g6, m = 6, 0.9, 30
    txt = scripts.IN_BUCK_DISP.format(n=n, g6=g6, m=m, steps=20, thermo=10, A=3000.0)   # a real repulsive wall
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.disp", txt), "-sf", "intel"])
    assert r.returncode == 0, r.stdout + r.stderr
    th = _thermo(r.stdout)
    s = W.fcc_system(n, n, n, jitter=0.0)
    A = np.zeros((2, 2)); rho = np.ones((2, 2)); C = np.zeros((2, 2))
    A[1, 1], rho[1, 1], C[1, 1] = 3000.0, 0.2, 0.8
    P = orc.Params(orc.BUCK_LONG_COUL_LONG, 1, A, rho, C, np.full((2, 2), 5.0), np.full((2, 2), 5.0), g_ewald_6=g6,
                   order6=1)
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], None, s["boxlo"], s["boxhi"], 0.3)
    pp = orc.PPPM.dispersion(m, m, m, 5, g6, s["boxlo"], s["boxhi"])
    B = np.array([0.0, np.sqrt(0.8)])
    fk, ek, vk = pp.compute(s["x"], B[s["type"]])
    assert th[0, 2] == pytest.approx(ev[0] + ek, rel=1e-9)
    assert np.abs(th[:, 4] - th[0, 4]).max() < 5e-4 * abs(th[0, 4])   # stiff, strongly compressed system


@pytest.mark.gpu
def test_lj_long_coul_long_with_pppm_disp_arithmetic_and_none(pkg, W, orc, tmp_path):
    """lj/long/coul/long long long with `pair_modify mix arithmetic` + pppm/disp through the driver: PPPMDispIntel::init
    picks function[2] (ewald_mix = ARITHMETIC), `kspace_modify mix/disp none` function[3]; step-0 E_pair against the
    oracle's pair + Coulomb PPPM + seven-grid dispersion PPPM.  Both rules see the same C_ij, so their energies agree."""
    e_pair = {}
    for rule, mixdisp in (("arithmetic", ""), ("none", "mix/disp none")):
        txt = scripts.IN_LJ_DISP_MIX.format(mixdisp=mixdisp, steps=4, thermo=2)
        r = _run(pkg, ["-in", scripts.write(tmp_path, "in.ljmix", txt, W), "-sf", "intel"])
        assert r.returncode == 0, r.stdout + r.stderr
        e_pair[rule] = _thermo(r.stdout)[0, 2]
    s = W.aC_system(1, jitter=0.0)
    u = W.UNITS["metal"]
    eps = np.array([0.0, 0.008, 0.021]); sig = np.array([0.0, 2.9, 3.3])
    e_ij = np.sqrt(np.outer(eps, eps)); s_ij = (sig[:, None] + sig[None, :]) / 2.0
    P = orc.Params(orc.LJ_LONG_COUL_LONG, 2, e_ij, s_ij, np.zeros((3, 3)), np.full((3, 3), 9.0), np.full((3, 3), 9.0),
                   qqrd2e=u["qqrd2e"], g_ewald=0.28, g_ewald_6=0.31, order1=1, order6=1)   # A = epsilon, rho = sigma
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    fc, ec, vc = orc.PPPM(24, 24, 27, 5, 0.28, s["boxlo"], s["boxhi"], u["qqrd2e"]).compute(s["x"], s["q"])
    B7 = orc.disp_B_arithmetic(eps, sig)
    fd, ed, vd = orc.PPPM.dispersion(30, 30, 32, 5, 0.31, s["boxlo"], s["boxhi"]).compute_arith(s["x"], B7[s["type"]])
    want = ev[0] + ev[1] + ec + ed
    assert e_pair["arithmetic"] == pytest.approx(want, rel=1e-9)
    assert e_pair["none"] == pytest.approx(want, rel=1e-9)


@pytest.mark.gpu
def test_in_buck_coul_cut_and_in_buck_big_run_like_the_oracle(pkg, W, orc, tmp_path):
    """the other two shipped scripts: in.buck_coul_cut (data.aC x 2^3 here, 4^3 as shipped) and in.buck_big
    (12x12x12 cells here): step-0 energy and pressure against the oracle, `delay 5 every 1 check yes` cadence"""
    txt = scripts.IN_BUCK_COUL_CUT.format(r=2, steps=10, thermo=5)
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.bcc", txt, W), "-sf", "intel"])
    assert r.returncode == 0, r.stdout + r.stderr
    th = _thermo(r.stdout)
    s = W.aC_system(2, jitter=0.0)
    u = W.UNITS["metal"]
    co = W.coeffs_aC(10.0, 10.0)
    P = orc.Params(orc.BUCK_COUL_CUT, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"])
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 0.3)
    assert th[0, 2] == pytest.approx(ev[0] + ev[1], rel=1e-10)
    n = len(s["x"])
    vol = np.prod(s["boxhi"] - s["boxlo"])
    press = ((3 * n - 3) * u["boltz"] * 300.0 + ev[2] + ev[3] + ev[4]) / 3.0 / vol * 1.6021765e6
    assert th[0, 5] == pytest.approx(press, rel=1e-8)
    # in.buck_big: delay 5 => no rebuild during the first 5 steps whatever the atoms do
    txt = scripts.IN_BUCK_BIG.format(nx=12, ny=12, nz=12, steps=5, thermo=5)
    r = _run(pkg, ["-in", scripts.write(tmp_path, "in.big", txt), "-sf", "intel"])
    assert r.returncode == 0, r.stdout + r.stderr
    th = _thermo(r.stdout)
    s = W.fcc_system(12, 12, 12, jitter=0.0)
    co = W.coeffs_in_buck(5.0)
    P = orc.Params(orc.BUCK, 1, co["A"], co["rho"], co["C"], co["cut_lj"])
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], None, s["boxlo"], s["boxhi"], 0.3)
    assert th[0, 2] == pytest.approx(ev[0], rel=1e-10)
    assert "Neighbor list builds = 0" in r.stdout
    # continuum estimate 4/3 pi (5.3)^3 * 0.8442 = 526.5; the perfect fcc lattice has 530 within 5.3
    assert "Ave neighs/atom = 530" in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="the reference tree is only mounted in the build container")
def test_read_data_atom_style_full_data_spce(pkg, W, tmp_path):
    """`atom_style full` + `read_data data.spce` (examples/in.spce:3-6): the molecule column is skipped, charges and
    positions are read, Bonds / Angles sections are ignored; PPPM sizing of the replicated box equals the Python one"""
    txt = ("units real\natom_style full\nread_data %s/data.spce\nreplicate 2 2 2\n"
           "pair_style buck/coul/long 8.8\npair_coeff * * 0.0 1.0 0.0\nkspace_style pppm 1.0e-4\n"
           "neighbor 2.0 bin\nneigh_modify every 1 delay 10 check yes\nfix 1 all nve\nrun 0\n" % REF_EXAMPLES)
    p = scripts.write(tmp_path, "in.spce_pppm", txt)
    r = _run(pkg, ["-in", p, "-sf", "intel", "-dry-run"])
    assert r.returncode == 0, r.stdout + r.stderr
    s = _summary(r.stdout)
    sysd = W.spce_system(2)
    u = W.UNITS["real"]
    assert s["natoms"] == 36000 == len(sysd["x"])
    assert np.allclose(s["box"], sysd["boxhi"] - sysd["boxlo"], rtol=1e-12)
    grid, g = pkg.pppm_init(1e-4, u["qqrd2e"], sysd["q"], 36000, 8.8, sysd["boxhi"] - sysd["boxlo"])
    assert tuple(s["grid"]) == tuple(grid) and s["g_ewald"] == pytest.approx(g, rel=1e-9)


def _write_data_spce(W, tmp_path):
    return scripts.write_data_spce(os.path.join(str(tmp_path), "data.spce"))


def test_dry_run_in_spce_nve(pkg, W, tmp_path):
    """the in.spce commands on the pair / k-space path are accepted: atom_style full with Bonds, lj/cut/coul/long mapped
    onto lj/long/coul/long cut long, special_bonds, the bonded-style commands (ignored), PPPM sizing"""
    data = _write_data_spce(W, tmp_path)
    p = scripts.write(tmp_path, "in.spce_nve", scripts.IN_SPCE_NVE.format(data=data, r=2, pair_modify="", thermo=0, steps=0))
    r = _run(pkg, ["-in", p, "-sf", "intel", "-dry-run"])
    assert r.returncode == 0, r.stdout + r.stderr
    s = _summary(r.stdout)
    sysd = W.spce_system(2)
    u = W.UNITS["real"]
    grid, g = pkg.pppm_init(1e-4, u["qqrd2e"], sysd["q"], 36000, 8.8, sysd["boxhi"] - sysd["boxlo"])
    assert s["natoms"] == 36000 and s["pair_style"] == "lj/cut/coul/long" and s["cutforce"] == 8.8
    assert tuple(s["grid"]) == tuple(grid) and s["g_ewald"] == pytest.approx(g, rel=1e-9)
    assert (s["every"], s["delay"], s["check"]) == (1, 10, 1) and s["dt"] == 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("table", [0, 12])
def test_in_spce_nve_runs_like_the_oracle(pkg, W, orc, tmp_path, table):
    """in.spce's non-bonded + k-space force through the driver and the host classes (PairLJLongCoulLongIntel, PPPMIntel,
    special bonds from the Bonds section): step-0 E_pair against the oracle on the same system"""
    import util
    data = _write_data_spce(W, tmp_path)
    txt = scripts.IN_SPCE_NVE.format(data=data, r=1, pair_modify="pair_modify table %d" % table, thermo=1, steps=2)
    p = scripts.write(tmp_path, "in.spce_nve", txt)
    r = _run(pkg, ["-in", p, "-sf", "intel"])
    assert r.returncode == 0, r.stdout + r.stderr
    th = _thermo(r.stdout)
    s = W.spce_system(1)
    # the data file is written with 5 decimals: read the same numbers
    s["x"] = W.wrap(np.round(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_spce.npz"))["x"], 5),
                    s["boxlo"], s["boxhi"])
    n = len(s["x"])
    u = W.UNITS["real"]
    co = W.coeffs_spce()
    grid, ge = pkg.pppm_init(1e-4, u["qqrd2e"], s["q"], n, 8.8, s["boxhi"] - s["boxlo"])
    P = orc.Params(orc.LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=ge, order1=1, special_lj=(1, 0.0, 0.0, 0.5), special_coul=(1, 0.0, 0.0, 0.5))
    if table:
        ct = pkg.init_coul_tables(8.8, ge, u["qqrd2e"])
        P.set_coul_tables(ct[0], 12, ct[1], ct[2], ct[3])
        dt = pkg.init_disp_tables(6.8, 0.3)
        P.set_disp_tables(dt[0], 0, dt[1], dt[2], dt[3])
    cm = P.cutmax() + 2.0
    xa, ta, qa, src, shift = orc.make_ghosts(s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], cm)
    hn, hoff, hent = orc.neigh_half_bin(n, xa, ta, 2, P.cutneighsq(2.0), s["boxlo"], s["boxhi"], cm, 0)
    hent = util.water_special_bits(n, hn, hent, src, s["mol"], s["type"])
    fo, evo = orc.pair_eval(P, 0, 1, 0, n, xa, ta, qa, hn, hoff, hent, newton=1)
    fk, ek, vk = orc.PPPM(*grid, 5, ge, s["boxlo"], s["boxhi"], u["qqrd2e"]).compute(s["x"], s["q"])
    assert th[0, 2] == pytest.approx(evo[0] + evo[1] + ek, rel=1e-8)
    assert th[0, 1] == pytest.approx(300.0, rel=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("table", [0, 12])
def test_in_hexane_nve_runs_like_the_oracle(pkg, W, orc, tmp_path, table):
    """in.hexane's non-bonded + k-space force through the driver and the host classes (PairLJLongCoulLongIntel with
    `long off`, PPPMDispIntel sizing g_ewald_6 and the dispersion mesh from `force/disp/real`, `force/disp/kspace`):
    step-0 E_pair against the oracle at the g_ewald_6 / mesh the dry run reports; `table/disp 12` is the stock default"""
    data = scripts.write_data_hexane(os.path.join(str(tmp_path), "data.hexane"))
    txt = scripts.IN_HEXANE_NVE.format(data=data, kspace_modify="", pair_modify="pair_modify table/disp %d" % table,
                                       thermo=1, steps=2, dt=1.0e-4)
    p = scripts.write(tmp_path, "in.hexane_nve", txt)
    d = _summary(_run(pkg, ["-in", p, "-sf", "intel", "-dry-run"]).stdout)
    g6, grid6 = d["g_ewald_6"], tuple(d["grid_6"])
    assert grid6 == (50, 24, 20) and g6 == pytest.approx(0.3044751226, rel=1e-9)
    r = _run(pkg, ["-in", p, "-sf", "intel"])
    assert r.returncode == 0, r.stdout + r.stderr
    th = _thermo(r.stdout)
    s = W.hexane_system()
    co = W.coeffs_hexane()
    P = orc.Params(orc.LJ_LONG_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], g_ewald_6=g6,
                   order1=0, order6=1)
    if table:
        dt = pkg.init_disp_tables(9.8, g6)
        P.set_disp_tables(dt[0], 12, dt[1], dt[2], dt[3])
    f, ev, _ = orc.pair_forces_periodic(P, 0, s["x"], s["type"], s["q"], s["boxlo"], s["boxhi"], 2.0)
    B = np.sqrt(4.0 * np.array([0.0, co["A"][1, 1], co["A"][2, 2]]) * 3.97 ** 6)
    fk, ek, vk = orc.PPPM.dispersion(*grid6, 5, g6, s["boxlo"], s["boxhi"]).compute(s["x"], B[s["type"]])
    assert th[0, 2] == pytest.approx(ev[0] + ek, rel=1e-9)
    assert th[0, 1] == pytest.approx(d["temperature"], rel=1e-7)      # the velocities of the data file (8 printed digits)
