import os

import numpy as np
import pytest

REF = "/root/reference/examples/data.aC"


def test_fcc_counts(W):
    s = W.fcc_system(20, 20, 20)
    assert len(s["x"]) == 32000                       # examples/in.buck: 32 000 atoms
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    assert np.allclose(s["boxhi"], 20 * a)
    assert ((s["x"] >= s["boxlo"]) & (s["x"] < s["boxhi"])).all()
    assert np.abs((s["v"] * s["mass"][s["type"]][:, None]).sum(0)).max() < 1e-9


def test_aC_counts_and_neutrality(W):
    s = W.aC_system(2)
    assert len(s["x"]) == 9600                        # examples/in.buck_coul_long: data.aC x 2^3
    assert abs(s["q"].sum()) < 1e-9
    assert (s["type"] == 1).sum() * 2 == (s["type"] == 2).sum()


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_data_aC_regenerated_matches_shipped_file(W):
    x, t, q, lo, hi = W.data_aC()
    ref = np.loadtxt(REF, skiprows=16)
    assert ref.shape == (1200, 6)
    assert np.abs(ref[:, 3:6] - x).max() < 2e-5       # the file is rounded to 5 decimals
    assert (ref[:, 1].astype(int) == t).all() and np.abs(ref[:, 2] - q).max() == 0.0
    hdr = open(REF).read().split("\n")[5:8]
    assert float(hdr[0].split()[1]) == hi[0] and float(hdr[2].split()[1]) == hi[2]


@pytest.mark.skipif(not os.path.exists("/root/reference/examples/data.spce"), reason="reference tree not mounted")
def test_data_spce_fixture_is_the_reference_file():
    """tests/golden/data_spce.npz is examples/data.spce (positions, charges, types, box), re-parsed here"""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_data_spce", os.path.join(here, "golden", "make_data_spce.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    d = m.parse("/root/reference/examples/data.spce")
    g = np.load(os.path.join(here, "golden", "data_spce.npz"))
    for k in ("x", "q", "type", "mol", "boxlo", "boxhi", "mass"):
        assert np.array_equal(d[k], g[k]), k
    assert len(g["x"]) == 4500 and abs(g["q"].sum()) < 1e-9


def test_spce_system_replicates_the_fixture(W):
    s = W.spce_system(2)
    assert len(s["x"]) == 36000 and s["units"] == "real"
    assert np.all(s["x"] >= s["boxlo"]) and np.all(s["x"] < s["boxhi"])
    assert abs(s["q"].sum()) < 1e-8


@pytest.mark.skipif(not os.path.exists("/root/reference/examples/equilibrated_data.hexane"), reason="reference tree not mounted")
def test_data_hexane_fixture_is_the_reference_file():
    """tests/golden/data_hexane.npz is examples/equilibrated_data.hexane (positions, velocities, types, molecule ids, box,
    masses), re-parsed here"""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_data_hexane", os.path.join(here, "golden", "make_data_hexane.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    src = "/root/reference/examples/equilibrated_data.hexane"
    d = m.parse(src)
    d["v"] = m.velocities(src, len(d["x"]))
    g = np.load(os.path.join(here, "golden", "data_hexane.npz"))
    for k in ("x", "v", "type", "mol", "boxlo", "boxhi", "mass"):
        assert np.array_equal(d[k], g[k]), k
    assert len(g["x"]) == 6000 and not d["q"].any()


def test_hexane_system(W):
    """1 000 hexane molecules of six united atoms: CH3 (type 1) at both ends, four CH2 (type 2) between; wrapped into the box"""
    s = W.hexane_system()
    assert len(s["x"]) == 6000 and s["units"] == "real" and s["ntypes"] == 2
    assert np.all(s["x"] >= s["boxlo"]) and np.all(s["x"] < s["boxhi"])
    assert np.bincount(s["type"])[1:].tolist() == [2000, 4000]
    mol, cnt = np.unique(s["mol"], return_counts=True)
    assert len(mol) == 1000 and np.all(cnt == 6)
    for m in mol[:50]:
        assert np.bincount(s["type"][s["mol"] == m], minlength=3)[1:].tolist() == [2, 4]
    co = W.coeffs_hexane()
    assert co["A"][1, 2] == pytest.approx(np.sqrt(0.1744742 * 0.1147228)) and np.all(co["rho"][1:, 1:] == 3.97)
