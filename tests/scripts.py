"""Input scripts with the semantics of the reference's examples (in.buck, in.buck_big, in.buck_coul_cut,
in.buck_coul_long — SURVEY §6.2), written out for the lmp_b200 driver; data.aC is regenerated from its 12-atom basis
(workloads.data_aC), so nothing under /root/reference is needed at run time."""
import os

IN_BUCK = """# Buckingham melt (examples/in.buck semantics)
variable x index 1
variable y index 1
variable z index 1
variable xx equal {n}*$x
variable yy equal {n}*$y
variable zz equal {n}*$z
units lj
atom_style atomic
lattice fcc 0.8442
region box block 0 ${{xx}} 0 ${{yy}} 0 ${{zz}}
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
pair_style buck 2.5
pair_coeff 1 1 1.0 0.2 -0.8
neighbor 0.3 bin
neigh_modify delay 0 every 20 check no
fix 1 all nve
thermo {thermo}
run {steps}
"""

IN_BUCK_COUL_CUT = """units metal
atom_style charge
read_data data.aC
replicate {r} {r} {r}
pair_style buck/coul/cut 10.0
pair_coeff 2 2 1388.77 .3623188 175.0
pair_coeff 1 2 18003 .2052124 133.5381
pair_coeff 1 1 0 .1 0
neighbor 0.3 bin
neigh_modify delay 0 every 1 check yes
velocity all create 300.0 1281937
fix 1 all nve
thermo {thermo}
run {steps}
"""

IN_BUCK_COUL_LONG = """units metal
atom_style charge
read_data data.aC
replicate {r} {r} {r}
pair_style buck/coul/long 12.0
pair_coeff 2 2 1388.77 .3623188 175.0
pair_coeff 1 2 18003 .2052124 133.5381
pair_coeff 1 1 0 .1 0
{pair_modify}
kspace_style {kspace}
neighbor 0.3 bin
neigh_modify delay 0 every 1 check yes
velocity all create 300.0 1281937
fix 1 all nve
thermo {thermo}
run {steps}
"""

IN_BUCK_DISP = """# in.buck_big variant of BASELINE config 5: long-range dispersion + pppm/disp
units lj
atom_style atomic
lattice fcc 0.8442
region box block 0 {n} 0 {n} 0 {n}
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
pair_style buck/long/coul/long long off 5.0
pair_coeff 1 1 {A} 0.2 0.8
pair_modify table/disp 0
kspace_style pppm/disp 1e-4
kspace_modify gewald/disp {g6} mesh/disp {m} {m} {m} order/disp 5
neighbor 0.3 bin
neigh_modify delay 5 every 1
fix 1 all nve
thermo {thermo}
run {steps}
"""


# lj/long/coul/long long long + pppm/disp with arithmetic mixing (PPPMDisp function[0] + function[2]) or, with
# `kspace_modify mix/disp none`, the no-mixing-rule branch (function[3]) on data.aC's two atom types
IN_LJ_DISP_MIX = """units metal
atom_style charge
read_data data.aC
pair_style lj/long/coul/long long long 9.0
pair_coeff 1 1 0.008 2.9
pair_coeff 2 2 0.021 3.3
pair_modify mix arithmetic table 0 table/disp 0
kspace_style pppm/disp 1e-4
kspace_modify mesh 24 24 27 gewald 0.28 mesh/disp 30 30 32 gewald/disp 0.31 order/disp 5 {mixdisp}
neighbor 0.3 bin
neigh_modify delay 0 every 1 check yes
velocity all create 300.0 1281937
fix 1 all nve
thermo {thermo}
run {steps}
"""


def write_data_aC(W, path):
    x, t, q, lo, hi = W.data_aC()
    with open(path, "w") as fh:
        fh.write(" regenerated alpha-cristobalite cell\n\n %d atoms\n 2 atom types\n\n" % len(x))
        for d, nm in enumerate("xyz"):
            fh.write(" %.16g %.16g %slo %shi\n" % (lo[d], hi[d], nm, nm))
        fh.write("\n Masses\n\n 1 28.0855\n 2 15.9999\n\n Atoms\n\n")
        for i in range(len(x)):
            fh.write(" %d %d %.16g %.16g %.16g %.16g\n" % (i + 1, t[i], q[i], x[i, 0], x[i, 1], x[i, 2]))


def write(tmpdir, name, text, W=None):
    p = os.path.join(str(tmpdir), name)
    with open(p, "w") as fh:
        fh.write(text)
    if W is not None:
        write_data_aC(W, os.path.join(str(tmpdir), "data.aC"))
    return p

IN_BUCK_BIG = """# examples/in.buck_big semantics (192 000 atoms as shipped; n scales the box for the tests)
units lj
atom_style atomic
lattice fcc 0.8442
region box block 0 {nx} 0 {ny} 0 {nz}
create_box 1 box
create_atoms 1 box
mass 1 1.0
velocity all create 1.44 87287 loop geom
pair_style buck 5.0
pair_coeff 1 1 0.8 0.2 -0.8
neighbor 0.3 bin
neigh_modify delay 5 every 1
fix 1 all nve
thermo {thermo}
run {steps}
"""

# examples/in.spce with the parts that are on the pair / k-space path: the real data.spce (atom_style full, Bonds read
# for the special lists), lj/cut/coul/long 6.8 8.8, pppm 1.0e-4, special_bonds lj/coul 0.0 0.0 0.5, the script's
# neighbour settings and 2 fs step.  SHAKE, the bonded terms and NVT of the original are outside that path: fix nve.
IN_SPCE_NVE = """units real
atom_style full
read_data {data}
replicate {r} {r} {r}
pair_style lj/cut/coul/long 6.8 8.8
kspace_style pppm 1.0e-4
pair_coeff 1 1 0.15535 3.166
pair_coeff * 2 0.0000 0.0000
bond_style harmonic
angle_style harmonic
dihedral_style none
improper_style none
bond_coeff 1 1000.00 1.000
angle_coeff 1 100.0 109.47
special_bonds lj/coul 0.0 0.0 0.5
{pair_modify}
neighbor 2.0 bin
neigh_modify every 1 delay 10 check yes
fix 1 all nve
velocity all create 300 432567 dist uniform
timestep 2.0
thermo {thermo}
run {steps}
"""


# examples/in.hexane with the parts that are on the pair / k-space path: the real equilibrated_data.hexane (atom_style
# full, no charges, no bonds), lj/long/coul/long long off 9.8, pppm/disp 1.0e-4 with the script's own accuracy split
# (`kspace_modify force/disp/real 0.0001`, `force/disp/kspace 0.002`: g_ewald_6 and the dispersion mesh are sized from
# them), the script's pair_coeff lines (1-2 by geometric mixing) and neighbour settings.  `fix rigid/small molecule` and
# the image dump of the original are outside that path: fix nve — and, with nothing holding the overlapping united
# atoms of a molecule apart from each other's r^-12 wall (4e5 kcal/mol/A), a time step far below the script's 2 fs.
IN_HEXANE_NVE = """units real
atom_style full
read_data {data}
pair_style lj/long/coul/long long off 9.8
kspace_style pppm/disp 1.0e-4
kspace_modify force/disp/real 0.0001
kspace_modify force/disp/kspace 0.002
{kspace_modify}
pair_coeff 1 1 0.1744742 3.97
pair_coeff 2 2 0.1147228 3.97
{pair_modify}
neighbor 2.0 bin
neigh_modify every 1 delay 10 check yes
fix 1 all nve
timestep {dt}
thermo_style one
thermo {thermo}
run {steps}
"""


def write_data_hexane(path):
    """equilibrated_data.hexane for the driver, regenerated from the committed fixture with the file's own layout
    (`Atoms # full`: id mol type q x y z, then Velocities) and full precision"""
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_hexane.npz"))
    n = len(d["x"])
    with open(path, "w") as fh:
        fh.write("LAMMPS data file regenerated from tests/golden/data_hexane.npz\n\n%d atoms\n2 atom types\n\n" % n)
        for k, c in enumerate("xyz"):
            fh.write("%.16e %.16e %slo %shi\n" % (d["boxlo"][k], d["boxhi"][k], c, c))
        fh.write("\nMasses\n\n1 %.16g\n2 %.16g\n\nAtoms # full\n\n" % (d["mass"][1], d["mass"][2]))
        for i in range(n):
            fh.write("%d %d %d 0.0 %.16e %.16e %.16e 0 0 0\n" % (i + 1, d["mol"][i], d["type"][i], *d["x"][i]))
        fh.write("\nVelocities\n\n")
        for i in range(n):
            fh.write("%d %.16e %.16e %.16e\n" % (i + 1, *d["v"][i]))
    return path


def write_data_spce(path):
    """data.spce for the driver, regenerated from the committed fixture (atom_style full: id mol type q x y z) with the
    O-H bonds of every molecule; the reference file itself is not available on the GPU box"""
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_spce.npz"))
    n = len(d["x"])
    with open(path, "w") as fh:
        fh.write("LAMMPS Atom File\n\n%d atoms\n%d bonds\n\n2 atom types\n1 bond types\n\n" % (n, 2 * (n // 3)))
        for k, c in enumerate("xyz"):
            fh.write("%.5f %.5f %slo %shi\n" % (d["boxlo"][k], d["boxhi"][k], c, c))
        fh.write("\nMasses\n\n1 %.4f\n2 %.5f\n\nAtoms\n\n" % (d["mass"][1], d["mass"][2]))
        for i in range(n):
            fh.write("%d %d %d %.4f %.5f %.5f %.5f 0 0 0\n" % (i + 1, d["mol"][i], d["type"][i], d["q"][i], *d["x"][i]))
        fh.write("\nBonds\n\n")
        b = 1
        for o in range(0, n, 3):
            for h in (1, 2):
                fh.write("%d 1 %d %d\n" % (b, o + 1, o + 1 + h))
                b += 1
    return path
