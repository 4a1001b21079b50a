"""examples/equilibrated_data.hexane of the reference -> tests/golden/data_hexane.npz.

    python tests/golden/make_data_hexane.py [/root/reference/examples/equilibrated_data.hexane]

The input of examples/in.hexane (`lj/long/coul/long long off 9.8` + `pppm/disp 1.0e-4`: the script the reference's
pair_lj_long_coul_long_intel / pppm_disp_intel pair exists for) is not available on the GPU box, so it travels as this
fixture: 6 000 united-atom sites (1 000 hexane molecules, CH3 = type 1, CH2 = type 2, no charges, no Bonds section),
positions as in the file (read_data wraps them), velocities, molecule ids, box and masses; atoms in id order.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_data_spce import parse  # noqa: E402


def velocities(path, natoms):
    with open(path) as fh:
        lines = fh.read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.split()[:1] == ["Velocities"]) + 2
    rows = sorted((l.split() for l in lines[i:i + natoms]), key=lambda r: int(r[0]))
    assert [int(r[0]) for r in rows] == list(range(1, natoms + 1))
    return np.array([[float(c) for c in r[1:4]] for r in rows])


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples/equilibrated_data.hexane"
    d = parse(src)
    d["v"] = velocities(src, len(d["x"]))
    assert not d["q"].any()
    del d["q"]
    np.savez_compressed(os.path.join(HERE, "data_hexane.npz"), **d)
    print("atoms", len(d["x"]), "molecules", len(set(d["mol"])), "types", np.bincount(d["type"])[1:], "box", d["boxlo"], d["boxhi"])
