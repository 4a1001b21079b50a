"""Seeded synthetic inputs of the shapes BASELINE.json names.  numpy only.

`/root/reference` is not available on the GPU box, so `data.aC` (examples/data.aC, the alpha-cristobalite
SiO2 cell used by in.buck_coul_cut / in.buck_coul_long) is regenerated from its 12-atom basis; a CPU test
(tests/test_workloads.py) checks the regenerated coordinates against the shipped file when it is present.
"""
import numpy as np

# units
UNITS = {
    "lj": dict(qqrd2e=1.0, ftm2v=1.0, boltz=1.0, dt=0.005, mvv2e=1.0),
    "metal": dict(qqrd2e=14.399645, ftm2v=1.0 / 1.0364269e-4, boltz=8.617343e-5, dt=0.001, mvv2e=1.0364269e-4),
    "real": dict(qqrd2e=332.06371, ftm2v=1.0 / 48.88821291 / 48.88821291, boltz=0.0019872067, dt=1.0,
                 mvv2e=48.88821291 * 48.88821291),
}

# examples/data.aC: basis of one tetragonal cell (a = 25.15832/5, c = 28.020256/4), type, charge, xyz
_AC_A = 25.15832 / 5.0
_AC_C = 28.020256 / 4.0
_AC_BASIS = [
    (1, 2.96653, 1.50970, 1.50970, 0.00000), (1, 2.96653, -1.50970, -1.50970, 3.50253),
    (1, 2.96653, 1.00613, 4.02553, 1.75127), (1, 2.96653, 4.02553, 1.00613, 5.25380),
    (2, -1.483265, 1.20639, 0.51947, 1.24998), (2, -1.483265, -1.20639, -0.51947, 4.75252),
    (2, -1.483265, 1.99636, 3.72222, 3.00125), (2, -1.483265, 3.03530, 1.30944, 6.50378),
    (2, -1.483265, 0.51947, 1.20639, -1.24998), (2, -1.483265, -0.51947, -1.20639, 2.25255),
    (2, -1.483265, 1.30944, 3.03530, 0.50128), (2, -1.483265, 3.72222, 1.99636, 4.00381),
]
AC_MASS = np.array([0.0, 28.0855, 15.9999])


def data_aC():
    """The 1200-atom data.aC system: (x[n,3], type[n], q[n], boxlo, boxhi), atoms NOT yet wrapped."""
    xs, ts, qs = [], [], []
    for ix in range(5):
        for iy in range(5):
            for iz in range(4):
                for (t, q, bx, by, bz) in _AC_BASIS:
                    xs.append((bx + ix * _AC_A, by + iy * _AC_A, bz + iz * _AC_C))
                    ts.append(t)
                    qs.append(q)
    boxlo = np.zeros(3)
    boxhi = np.array([25.15832, 25.15832, 28.020256])
    return np.array(xs), np.array(ts, np.int32), np.array(qs), boxlo, boxhi


def wrap(x, boxlo, boxhi):
    """read_data remaps atoms into the periodic box (data.aC:18 has negative coordinates)."""
    prd = boxhi - boxlo
    x = x.copy()
    for d in range(3):
        lo = x[:, d] < boxlo[d]
        x[lo, d] += prd[d]
        hi = x[:, d] >= boxhi[d]
        x[hi, d] -= prd[d]
        x[:, d] = np.maximum(x[:, d], boxlo[d])
    return x


def replicate(x, type_, q, boxlo, boxhi, nx, ny, nz):
    """`replicate nx ny nz`: images ordered k (z) outer, j, i (x) inner, atoms inner-most."""
    prd = boxhi - boxlo
    n = len(x)
    reps = nx * ny * nz
    xo = np.empty((n * reps, 3))
    m = 0
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                xo[m:m + n] = x + np.array([i * prd[0], j * prd[1], k * prd[2]])
                m += n
    to = np.tile(type_, reps)
    qo = np.tile(q, reps) if q is not None else None
    hi = boxlo + prd * np.array([nx, ny, nz])
    return xo, to, qo, boxlo.copy(), hi


def aC_system(rep, seed=1281937, temperature=300.0, jitter=0.02):
    """S3: data.aC wrapped, replicated rep^3 (rep may be a 3-tuple); small seeded displacement (a perfect
    crystal has vanishing net forces, useless for parity); velocities Gaussian at `temperature` K."""
    r = (rep, rep, rep) if np.isscalar(rep) else tuple(rep)
    x, t, q, lo, hi = data_aC()
    x = wrap(x, lo, hi)
    x, t, q, lo, hi = replicate(x, t, q, lo, hi, *r)
    rng = np.random.default_rng(seed)
    x = wrap(x + rng.uniform(-jitter, jitter, x.shape), lo, hi)
    v = velocities(rng, t, AC_MASS, temperature, UNITS["metal"])
    return dict(x=x, v=v, type=t, q=q, boxlo=lo, boxhi=hi, mass=AC_MASS.copy(), ntypes=2, units="metal")


def fcc_system(nx, ny, nz, rho_star=0.8442, seed=87287, jitter=0.05, temperature=1.44):
    """S1/S2: `lattice fcc rho*` in lj units, region block 0 nx 0 ny 0 nz, create_atoms (loops k,j,i,basis),
    + uniform displacement +-jitter*a per coordinate; velocities at T* = 1.44 with zero momentum."""
    a = (4.0 / rho_star) ** (1.0 / 3.0)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    cells = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(float)
    x = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) * a
    lo = np.zeros(3)
    hi = np.array([nx, ny, nz]) * a
    rng = np.random.default_rng(seed)
    x = wrap(x + rng.uniform(-jitter * a, jitter * a, x.shape), lo, hi)
    t = np.ones(len(x), np.int32)
    mass = np.array([0.0, 1.0])
    v = velocities(rng, t, mass, temperature, UNITS["lj"])
    return dict(x=x, v=v, type=t, q=None, boxlo=lo, boxhi=hi, mass=mass, ntypes=1, units="lj")


def spce_system(rep, seed=432567, temperature=300.0):
    """S4: examples/data.spce (4 500 atoms, SPC/E water; tests/golden/data_spce.npz, made by
    tests/golden/make_data_spce.py from the reference's file) wrapped into the box and replicated — the PPPM input of
    in.spce (`read_data data.spce`, `replicate 4 4 4`; positions and charges only: bonds/SHAKE are not on the path)."""
    import os
    r = (rep, rep, rep) if np.isscalar(rep) else tuple(rep)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data_spce.npz")
    d = np.load(path)
    lo, hi = d["boxlo"].copy(), d["boxhi"].copy()
    x = wrap(d["x"], lo, hi)
    x, t, q, lo, hi = replicate(x, d["type"], d["q"], lo, hi, *r)
    rng = np.random.default_rng(seed)
    v = velocities(rng, t, d["mass"], temperature, UNITS["real"])
    nrep = r[0] * r[1] * r[2]
    mol = (np.tile(d["mol"], nrep) + np.repeat(np.arange(nrep), len(d["mol"])) * int(d["mol"].max())).astype(np.int32)
    return dict(x=x, v=v, type=t, q=q, boxlo=lo, boxhi=hi, mass=d["mass"].copy(), ntypes=2, units="real", mol=mol)


def coeffs_spce(cut_lj=6.8, cut_coul=8.8):
    """examples/in.spce:7-11: lj/cut/coul/long 6.8 8.8, pair_coeff 1 1 0.15535 3.166, * 2 0 0 (epsilon as A, sigma as rho)"""
    eps = np.zeros((3, 3)); sig = np.ones((3, 3))
    eps[1, 1], sig[1, 1] = 0.15535, 3.166
    sig[1, 2] = sig[2, 1] = sig[2, 2] = 1.0
    return dict(A=eps, rho=sig, C=np.zeros((3, 3)), cut_lj=np.full((3, 3), float(cut_lj)), cut_coul=np.full((3, 3), float(cut_coul)))


def hexane_system():
    """examples/equilibrated_data.hexane (6 000 united-atom sites, 1 000 molecules, no charges; tests/golden/
    data_hexane.npz, made by tests/golden/make_data_hexane.py from the reference's file) wrapped into the box, with the
    file's own velocities — the input of in.hexane (`lj/long/coul/long long off 9.8` + `pppm/disp 1.0e-4`)."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data_hexane.npz")
    d = np.load(path)
    lo, hi = d["boxlo"].copy(), d["boxhi"].copy()
    return dict(x=wrap(d["x"], lo, hi), v=d["v"].copy(), type=d["type"].copy(), q=np.zeros(len(d["x"])), boxlo=lo, boxhi=hi,
                mass=d["mass"].copy(), ntypes=2, units="real", mol=d["mol"].copy())


def coeffs_hexane(cut=9.8):
    """examples/in.hexane:9,19-20: lj/long/coul/long long off 9.8; pair_coeff 1 1 0.1744742 3.97, 2 2 0.1147228 3.97, the 1-2
    pair by the style's default geometric mixing (epsilon as A, sigma as rho)"""
    eps = np.zeros((3, 3)); sig = np.ones((3, 3))
    eps[1, 1], eps[2, 2] = 0.1744742, 0.1147228
    eps[1, 2] = eps[2, 1] = np.sqrt(eps[1, 1] * eps[2, 2])
    sig[1:, 1:] = 3.97
    return dict(A=eps, rho=sig, C=np.zeros((3, 3)), cut_lj=np.full((3, 3), float(cut)), cut_coul=np.zeros((3, 3)))


def water_like_system(nmol_side, seed=4711, box=35.5):
    """S4 stand-in for data.spce (positions+charges only, PPPM-only workload): rigid SPC/E-geometry
    molecules (O -0.8472, H +0.4236, r_OH = 1, angle 109.47) on a jittered simple-cubic lattice with random
    orientations, at the data.spce density for nmol_side = 11 or 12 in a 35.5 A box."""
    rng = np.random.default_rng(seed)
    h = box / nmol_side
    g = np.arange(nmol_side)
    cx, cy, cz = np.meshgrid(g, g, g, indexing="ij")
    cen = (np.stack([cx.ravel(), cy.ravel(), cz.ravel()], 1) + 0.5) * h
    cen += rng.uniform(-0.25 * h, 0.25 * h, cen.shape)
    nm = len(cen)
    ang = np.deg2rad(109.47)
    local = np.array([[0, 0, 0], [np.sin(ang / 2), np.cos(ang / 2), 0], [-np.sin(ang / 2), np.cos(ang / 2), 0]])
    qn = rng.normal(size=(nm, 4))
    qn /= np.linalg.norm(qn, axis=1)[:, None]
    w, a, b, c = qn.T
    R = np.stack([np.stack([1 - 2 * (b * b + c * c), 2 * (a * b - c * w), 2 * (a * c + b * w)], 1),
                  np.stack([2 * (a * b + c * w), 1 - 2 * (a * a + c * c), 2 * (b * c - a * w)], 1),
                  np.stack([2 * (a * c - b * w), 2 * (b * c + a * w), 1 - 2 * (a * a + b * b)], 1)], 1)
    x = (cen[:, None, :] + np.einsum("mij,kj->mki", R, local)).reshape(-1, 3)
    t = np.tile(np.array([1, 2, 2], np.int32), nm)
    q = np.tile(np.array([-0.8472, 0.4236, 0.4236]), nm)
    lo = np.zeros(3)
    hi = np.full(3, float(box))
    x = wrap(x, lo, hi)
    mass = np.array([0.0, 15.9994, 1.00794])
    return dict(x=x, v=np.zeros_like(x), type=t, q=q, boxlo=lo, boxhi=hi, mass=mass, ntypes=2, units="real")


def velocities(rng, type_, mass, temperature, units):
    """Gaussian velocities, zero total momentum, rescaled to `temperature` with dof = 3N-3.  (The exact
    `velocity ... loop geom` RanPark stream is a SURVEY §8f 'next' item; this is a seeded stand-in.)"""
    m = mass[type_][:, None]
    v = rng.normal(size=(len(type_), 3)) / np.sqrt(m)
    v -= (m * v).sum(0) / m.sum()
    ke = 0.5 * units["mvv2e"] * (m * v * v).sum()
    dof = 3 * len(type_) - 3
    t_now = 2.0 * ke / (dof * units["boltz"])
    return v * np.sqrt(temperature / t_now)


# pair coefficients of the shipped scripts ---------------------------------------------------------

def coeffs_in_buck(cut=2.5):
    """examples/in.buck:22-23 (cut 2.5) / in.buck_big:12-13 (cut 5.0, A = 0.8)"""
    A = np.zeros((2, 2)); rho = np.ones((2, 2)); C = np.zeros((2, 2))
    A[1, 1] = 1.0 if cut == 2.5 else 0.8
    rho[1, 1] = 0.2
    C[1, 1] = -0.8
    return dict(A=A, rho=rho, C=C, cut_lj=np.full((2, 2), cut))


def coeffs_aC(cut_lj, cut_coul=None):
    """examples/in.buck_coul_cut:8-11 and in.buck_coul_long:8-11"""
    A = np.zeros((3, 3)); rho = np.ones((3, 3)); C = np.zeros((3, 3))
    A[2, 2], rho[2, 2], C[2, 2] = 1388.77, 0.3623188, 175.0
    A[1, 2] = A[2, 1] = 18003.0
    rho[1, 2] = rho[2, 1] = 0.2052124
    C[1, 2] = C[2, 1] = 133.5381
    A[1, 1], rho[1, 1], C[1, 1] = 0.0, 0.1, 0.0
    cc = cut_lj if cut_coul is None else cut_coul
    return dict(A=A, rho=rho, C=C, cut_lj=np.full((3, 3), float(cut_lj)), cut_coul=np.full((3, 3), float(cc)))
