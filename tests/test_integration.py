"""The drop-in translation units (lammps-buck-intel_b200/integration/*.cpp: the classes of the reference's own headers
implemented through the C ABI) linked into one library behind a harness of stand-in LAMMPS objects
(tests/integration_harness.cpp, oracle/_ref/libinteg.so).  CPU: the chain compiles, links, loads and — with no B200 in
the machine — fails loudly with the library's message coming back through error->all.  GPU: tests/integration_check.py
compares the forces with the oracle; it could not be run on a B200 before this round's GPU budget ran out, so it runs in
a child process and is marked xfail(strict=False): a failure there is reported, not fatal."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _have_lib():
    import integ
    try:
        return integ.build() is not None
    except RuntimeError:
        return False


def test_binding_chain_links_loads_and_fails_loudly_without_a_gpu(pkg, W, orc):
    import integ
    if not _have_lib():
        pytest.skip("oracle/_ref/libinteg.so: neither prebuilt nor buildable here")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: see test_integration_units_on_the_gpu")
    # every unit of integration/ is in the library, and its only unresolved b200md_* symbols come from libb200md.so
    nm = subprocess.run(["nm", "-C", "-D", integ.LIB], capture_output=True, text=True).stdout
    for cls in ("PairBuckIntel", "PairBuckCoulCutIntel", "PairBuckCoulLongIntel", "PairBuckLongCoulLongIntel",
                "PairLJLongCoulLongIntel"):
        assert " T LAMMPS_NS::%s::compute(int, int)" % cls in nm and " T LAMMPS_NS::%s::init_style()" % cls in nm, cls
    assert " T LAMMPS_NS::PPPMIntel::compute(int, int)" in nm and " T LAMMPS_NS::b200_positions_to_device" in nm
    assert " T LAMMPS_NS::PPPMDispIntel::compute(int, int)" in nm and " T LAMMPS_NS::PPPMDispIntel::init()" in nm
    und = subprocess.run(["ldd", "-r", integ.LIB], capture_output=True, text=True)
    assert "undefined symbol" not in und.stdout + und.stderr
    s = W.aC_system(1)
    u = W.UNITS["metal"]
    co = W.coeffs_aC(8.0, 8.0)
    P = orc.Params(orc.BUCK_COUL_LONG, 2, co["A"], co["rho"], co["C"], co["cut_lj"], co["cut_coul"], qqrd2e=u["qqrd2e"],
                   g_ewald=0.3)
    for grid in ((24, 24, 27), None):     # PPPMIntel::init or PairBuckCoulLongIntel::init_style makes the first C-ABI call
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            integ.buck_coul_long(P, 0, s, co["A"], co["rho"], co["C"], co["cut_lj"], 8.0, 0.3, grid=grid)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="written after this round's GPU budget was spent: never run on a B200 yet")
def test_integration_units_on_the_gpu():
    if not _have_lib():
        pytest.skip("oracle/_ref/libinteg.so: neither prebuilt nor buildable here")
    r = subprocess.run([sys.executable, os.path.join(HERE, "integration_check.py")], capture_output=True, text=True,
                       timeout=120, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "INTEGRATION OK" in r.stdout
