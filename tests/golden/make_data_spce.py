"""examples/data.spce of the reference -> tests/golden/data_spce.npz (positions, charges, types, molecule ids, box).

    python tests/golden/make_data_spce.py [/root/reference/examples/data.spce]

/root/reference does not exist on the GPU box, so the input of BASELINE config 4 (in.spce: `read_data data.spce`,
`replicate 4 4 4`, `kspace_style pppm 1.0e-4`) travels as this small fixture.  Only what the PPPM path reads is kept:
`atom_style full` lines `id mol type q x y z [ix iy iz]` -> x (as in the file: read_data wraps them), q, type, mol, and the
box bounds; bonds / angles are not part of the k-space path.  Atoms are stored in id order.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def parse(path):
    with open(path) as fh:
        lines = fh.read().splitlines()
    natoms = None
    lo, hi = np.zeros(3), np.zeros(3)
    masses = {}
    i = 0
    while i < len(lines):
        t = lines[i].split()
        if len(t) >= 2 and t[1] == "atoms":
            natoms = int(t[0])
        for d, key in enumerate(("xlo", "ylo", "zlo")):
            if len(t) >= 4 and t[2] == key:
                lo[d], hi[d] = float(t[0]), float(t[1])
        if t[:1] == ["Masses"]:
            i += 2
            while i < len(lines) and lines[i].strip():
                a, m = lines[i].split()[:2]
                masses[int(a)] = float(m)
                i += 1
            continue
        if t[:1] == ["Atoms"]:
            i += 2
            rows = []
            while i < len(lines) and lines[i].strip():
                rows.append(lines[i].split())
                i += 1
            rows.sort(key=lambda r: int(r[0]))
            ids = np.array([int(r[0]) for r in rows])
            mol = np.array([int(r[1]) for r in rows], np.int32)
            typ = np.array([int(r[2]) for r in rows], np.int32)
            q = np.array([float(r[3]) for r in rows])
            x = np.array([[float(r[4]), float(r[5]), float(r[6])] for r in rows])
            assert len(rows) == natoms and np.array_equal(ids, np.arange(1, natoms + 1))
            mass = np.zeros(max(masses) + 1)
            for a, m in masses.items():
                mass[a] = m
            return dict(x=x, q=q, type=typ, mol=mol, boxlo=lo, boxhi=hi, mass=mass)
        i += 1
    raise ValueError("no Atoms section in " + path)


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples/data.spce"
    d = parse(src)
    np.savez_compressed(os.path.join(HERE, "data_spce.npz"), **d)
    print("atoms", len(d["x"]), "net charge %.3e" % d["q"].sum(), "box", d["boxlo"], d["boxhi"])
