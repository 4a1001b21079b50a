"""ctypes wrapper of the CPU oracle (oracle/oracle.h).  TEST INFRASTRUCTURE ONLY.

Builds oracle/_build/liboracle.so on demand (rebuilt when the host CPU differs from the one it was
built on, since it is compiled with -march=native).  Nothing in the product imports this module.
"""
import ctypes as C
import hashlib
import os
import platform
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ODIR, "_build", "liboracle.so")
STAMP = os.path.join(ODIR, "_build", "host.txt")

BUCK, BUCK_COUL_CUT, BUCK_COUL_LONG, BUCK_LONG_COUL_LONG = 0, 1, 2, 3
LJ_LONG_COUL_LONG = 4   # A = epsilon, rho = sigma
DOUBLE, MIXED = 0, 1
SBBITS = 30
NEIGHMASK = 0x3FFFFFFF


def _host_id():
    try:
        with open("/proc/cpuinfo") as fh:
            flags = [l for l in fh if l.startswith(("model name", "flags"))][:2]
    except OSError:
        flags = [platform.processor()]
    return hashlib.sha1("".join(flags).encode()).hexdigest()


def build(force=False):
    srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.endswith((".cpp", ".h")) or f == "Makefile"]
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs)
    hid = _host_id()
    if os.path.exists(STAMP):
        with open(STAMP) as fh:
            stale = stale or fh.read().strip() != hid
    else:
        stale = True
    if stale or force:
        subprocess.run(["make", "-C", ODIR, "clean"], check=True, capture_output=True)
        r = subprocess.run(["make", "-C", ODIR], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
        with open(STAMP, "w") as fh:
            fh.write(hid)
    return LIB


_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_long)


class PairParams(C.Structure):
    _fields_ = [
        ("ntypes", C.c_int),
        ("cutsq", dp), ("cut_ljsq", dp), ("cut_coulsq", dp), ("buck1", dp), ("buck2", dp),
        ("rhoinv", dp), ("a", dp), ("c", dp), ("offset", dp),
        ("special_lj", C.c_double * 4), ("special_coul", C.c_double * 4),
        ("qqrd2e", C.c_double), ("g_ewald", C.c_double), ("g_ewald_6", C.c_double),
        ("order1", C.c_int), ("order6", C.c_int),
        ("ncoultablebits", C.c_int), ("ncoulmask", C.c_int), ("ncoulshiftbits", C.c_int),
        ("tabinnersq", C.c_double),
        ("rtable", dp), ("drtable", dp), ("ftable", dp), ("dftable", dp),
        ("etable", dp), ("detable", dp), ("ctable", dp), ("dctable", dp),
        ("ndisptablebits", C.c_int), ("ndispmask", C.c_int), ("ndispshiftbits", C.c_int),
        ("tabinnerdispsq", C.c_double),
        ("rdisptable", dp), ("drdisptable", dp), ("fdisptable", dp), ("dfdisptable", dp),
        ("edisptable", dp), ("dedisptable", dp),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_neigh_half_bin.restype = C.c_long
        _lib.orc_neigh_full_brute.restype = C.c_long
        _lib.orc_pppm_create.restype = C.c_void_p
        _lib.orc_pppm_nfft.restype = C.c_long
        for n in ("orc_pppm_greensfn", "orc_pppm_density_fft", "orc_pppm_field", "orc_pppm_sf_coeff"):
            getattr(_lib, n).restype = dp
        if hasattr(_lib, "orc_md_create"):
            _lib.orc_md_create.restype = C.c_void_p
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def _l(a):
    return a.ctypes.data_as(lp)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Params:
    """Owns the numpy arrays behind an orc_pair_params struct."""

    def __init__(self, style, ntypes, A, rho, Cc, cut_lj, cut_coul=None, offset_flag=0, qqrd2e=1.0,
                 g_ewald=0.0, g_ewald_6=0.0, order1=0, order6=0, special_lj=(1, 0, 0, 0),
                 special_coul=(1, 0, 0, 0)):
        tp1 = ntypes + 1
        self.style, self.ntypes = style, ntypes

        def full(v):
            v = np.asarray(v, dtype=np.float64)
            if v.ndim == 0:
                v = np.full((tp1, tp1), float(v))
            assert v.shape == (tp1, tp1)
            return f64(v)

        self.A, self.rho, self.Cc, self.cut_lj = full(A), full(rho), full(Cc), full(cut_lj)
        self.cut_coul = full(cut_coul) if cut_coul is not None else None
        self.arr = {k: np.zeros((tp1, tp1)) for k in
                    ("cutsq", "cut_ljsq", "cut_coulsq", "buck1", "buck2", "rhoinv", "a", "c", "offset")}
        self.p = PairParams()
        self.p.ntypes = ntypes
        for k, v in self.arr.items():
            setattr(self.p, k, _d(v))
        rho_safe = self.rho.copy()
        rho_safe[rho_safe == 0] = 1.0
        lib().orc_pair_init(C.c_int(style), C.c_int(ntypes), _d(self.A), _d(rho_safe), _d(self.Cc),
                            _d(self.cut_lj), _d(self.cut_coul) if self.cut_coul is not None else None,
                            C.c_int(offset_flag), C.byref(self.p))
        for i in range(4):
            self.p.special_lj[i] = special_lj[i]
            self.p.special_coul[i] = special_coul[i]
        self.p.qqrd2e, self.p.g_ewald, self.p.g_ewald_6 = qqrd2e, g_ewald, g_ewald_6
        self.p.order1, self.p.order6 = order1, order6
        self.tables = None

    def cutmax(self):
        return float(np.sqrt(self.arr["cutsq"].max()))

    def cutneighsq(self, skin):
        c = np.sqrt(self.arr["cutsq"]) + skin
        out = c * c
        out[0, :] = 0
        out[:, 0] = 0
        return f64(out)

    def make_coul_tables(self, cut_coul, tabinner=np.sqrt(2.0), nbits=12):
        n = 1 << nbits
        t = {k: np.zeros(n) for k in ("r", "dr", "f", "df", "e", "de", "c", "dc")}
        mask, shift, inner = C.c_int(), C.c_int(), C.c_double()
        lib().orc_init_coul_tables(C.c_double(cut_coul), C.c_double(tabinner), C.c_int(nbits),
                                   C.c_double(self.p.g_ewald), C.c_double(self.p.qqrd2e),
                                   _d(t["r"]), _d(t["dr"]), _d(t["f"]), _d(t["df"]), _d(t["e"]),
                                   _d(t["de"]), _d(t["c"]), _d(t["dc"]), C.byref(mask), C.byref(shift),
                                   C.byref(inner))
        self.set_coul_tables(t, nbits, mask.value, shift.value, inner.value)
        return t

    def set_coul_tables(self, t, nbits, mask, shift, tabinnersq):
        self.tables = t
        self.p.ncoultablebits, self.p.ncoulmask, self.p.ncoulshiftbits = nbits, mask, shift
        self.p.tabinnersq = tabinnersq
        self.p.rtable, self.p.drtable, self.p.ftable, self.p.dftable = _d(t["r"]), _d(t["dr"]), _d(t["f"]), _d(t["df"])
        self.p.etable, self.p.detable, self.p.ctable, self.p.dctable = _d(t["e"]), _d(t["de"]), _d(t["c"]), _d(t["dc"])


    def set_disp_tables(self, t, nbits, mask, shift, tabinnerdispsq):
        self.dtables = t
        self.p.ndisptablebits, self.p.ndispmask, self.p.ndispshiftbits = nbits, mask, shift
        self.p.tabinnerdispsq = tabinnerdispsq
        self.p.rdisptable, self.p.drdisptable, self.p.fdisptable = _d(t["r"]), _d(t["dr"]), _d(t["f"])
        self.p.dfdisptable, self.p.edisptable, self.p.dedisptable = _d(t["df"]), _d(t["e"]), _d(t["de"])


def make_ghosts(x, type_, q, boxlo, boxhi, cutghost, periodic=(1, 1, 1)):
    """Returns (x_all, type_all, q_all, src, shift) with ghosts appended."""
    n = len(x)
    prd = np.asarray(boxhi, float) - np.asarray(boxlo, float)
    frac = np.prod(1.0 + 2.0 * cutghost / prd) - 1.0
    cap = int(n * (1.0 + 1.5 * frac)) + 4096
    while True:
        xa = np.zeros((cap, 3)); xa[:n] = x
        ta = np.zeros(cap, np.int32); ta[:n] = type_
        qa = np.zeros(cap); qa[:n] = q if q is not None else 0.0
        src = np.zeros(cap, np.int32)
        shift = np.zeros((cap, 3), np.int32)
        ng = lib().orc_make_ghosts(C.c_int(n), _d(xa), _i(ta), _d(qa), _d(f64(boxlo)), _d(f64(boxhi)),
                                   _i(i32(periodic)), C.c_double(cutghost), C.c_int(cap), _i(src), _i(shift))
        if ng >= 0:
            break
        cap *= 2
    na = n + ng
    return xa[:na].copy(), ta[:na].copy(), qa[:na].copy(), src[:ng].copy(), shift[:ng].copy()


def _neigh(fn, nlocal, x, type_, ntypes, cutneighsq, prec, extra=()):
    nall = len(x)
    cap = max(1024, nlocal * 64)
    x = f64(x); type_ = i32(type_); cutneighsq = f64(cutneighsq)
    while True:
        numneigh = np.zeros(nlocal, np.int32)
        offsets = np.zeros(nlocal + 1, np.int64)
        entries = np.zeros(cap, np.int32)
        args = [C.c_int(nlocal), C.c_int(nall), _d(x), _i(type_), C.c_int(ntypes), _d(cutneighsq)]
        args += list(extra) + [C.c_int(prec), _i(numneigh), _l(offsets), _i(entries), C.c_long(cap)]
        tot = fn(*args)
        if tot >= 0:
            return numneigh, offsets, entries[:tot].copy()
        cap = int(offsets[nlocal]) + 16


def neigh_half_bin(nlocal, x, type_, ntypes, cutneighsq, boxlo, boxhi, cutneighmax, prec=DOUBLE):
    bl, bh = f64(boxlo), f64(boxhi)
    return _neigh(lib().orc_neigh_half_bin, nlocal, x, type_, ntypes, cutneighsq, prec,
                  extra=(_d(bl), _d(bh), C.c_double(cutneighmax)))


def neigh_full_brute(nlocal, x, type_, ntypes, cutneighsq, prec=DOUBLE):
    return _neigh(lib().orc_neigh_full_brute, nlocal, x, type_, ntypes, cutneighsq, prec)


def pair_eval(params, prec, eflag, vflag, nlocal, x, type_, q, numneigh, offsets, entries, newton=1,
              eatom=0, nthreads=0):
    nall = len(x)
    x = f64(x); type_ = i32(type_)
    q = f64(q) if q is not None else np.zeros(nall)
    f = np.zeros((nall, 4))
    ev = np.zeros(8)
    lib().orc_pair_eval(C.c_int(params.style), C.c_int(prec), C.c_int(eflag), C.c_int(vflag),
                        C.c_int(eatom), C.c_int(newton), C.c_int(nlocal), C.c_int(nall), _d(x),
                        _i(type_), _d(q), _i(i32(numneigh)), _l(np.ascontiguousarray(offsets, np.int64)),
                        _i(i32(entries)), C.byref(params.p), _d(f), _d(ev), C.c_int(nthreads))
    return f, ev


def reverse_comm(nlocal, src, f):
    f = f64(f)
    lib().orc_reverse_comm(C.c_int(nlocal), C.c_int(len(src)), _i(i32(src)), _d(f))
    return f


def pair_forces_periodic(params, prec, x, type_, q, boxlo, boxhi, skin, eflag=1, vflag=1, nthreads=0,
                         eatom=0):
    """Reference-shaped evaluation of a periodic system: ghosts -> half list (newton on) -> eval ->
    reverse comm.  Returns (f[nlocal,4], ev[8], aux)."""
    n = len(x)
    cutneighmax = params.cutmax() + skin
    xa, ta, qa, src, shift = make_ghosts(x, type_, q, boxlo, boxhi, cutneighmax)
    cns = params.cutneighsq(skin)
    nn, off, ent = neigh_half_bin(n, xa, ta, params.ntypes, cns, boxlo, boxhi, cutneighmax, prec)
    f, ev = pair_eval(params, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=1, eatom=eatom,
                      nthreads=nthreads)
    f = reverse_comm(n, src, f)
    return f[:n], ev, dict(x=xa, type=ta, q=qa, src=src, shift=shift, numneigh=nn, offsets=off, entries=ent)


def pair_forces_periodic_newtoff(params, prec, x, type_, q, boxlo, boxhi, skin, eflag=1, vflag=1, nthreads=0, eatom=0):
    """The reference's NEWTON_PAIR = 0 configuration (eval<EVFLAG,EFLAG,0>, pair_buck_intel.cpp:109-110): half list
    with newton off = owned-owned pairs once (j > i), owned-ghost pairs from each side, f[j] only for owned j.
    This is the configuration the device path mirrors (each side of a cross-boundary pair uses its own image)."""
    n = len(x)
    cutneighmax = params.cutmax() + skin
    xa, ta, qa, src, shift = make_ghosts(x, type_, q, boxlo, boxhi, cutneighmax)
    fn, foff, fent = neigh_full_brute(n, xa, ta, params.ntypes, params.cutneighsq(skin), prec)
    i = np.repeat(np.arange(n), fn)
    j = fent & NEIGHMASK
    keep = (j >= n) | (j > i)
    nn = np.bincount(i[keep], minlength=n).astype(np.int32)
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(nn)
    ent = fent[keep]
    f, ev = pair_eval(params, prec, eflag, vflag, n, xa, ta, qa, nn, off, ent, newton=0, eatom=eatom, nthreads=nthreads)
    return f[:n], ev


class PPPM:
    def __init__(self, nx, ny, nz, order, g_ewald, boxlo, boxhi, qqrd2e, diff_ad=0, prec=DOUBLE, slab=1.0):
        if slab > 1.0:
            lib().orc_pppm_create_slab.restype = C.c_void_p
            self.h = lib().orc_pppm_create_slab(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(order),
                                                C.c_double(g_ewald), C.c_int(diff_ad), _d(f64(boxlo)),
                                                _d(f64(boxhi)), C.c_double(qqrd2e), C.c_int(prec), C.c_double(slab))
        else:
            self.h = lib().orc_pppm_create(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(order),
                                           C.c_double(g_ewald), C.c_int(diff_ad), _d(f64(boxlo)),
                                           _d(f64(boxhi)), C.c_double(qqrd2e), C.c_int(prec))
        if not self.h:
            raise ValueError("PPPM order not supported")
        self.h = C.c_void_p(self.h)
        self.grid = (nx, ny, nz)
        self.order = order
        self.nfft = nx * ny * nz

    @classmethod
    def dispersion(cls, nx, ny, nz, order, g_ewald_6, boxlo, boxhi, prec=DOUBLE, diff_ad=0):
        """PPPMDispIntel 'g' grid (geometric mixing): compute(x, w) takes w[i] = B[type[i]]"""
        self = cls.__new__(cls)
        lib().orc_pppm_create_disp_ad.restype = C.c_void_p
        h = lib().orc_pppm_create_disp_ad(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(order), C.c_double(g_ewald_6),
                                          C.c_int(diff_ad), _d(f64(boxlo)), _d(f64(boxhi)), C.c_int(prec))
        if not h:
            raise ValueError("PPPM order not supported")
        self.h = C.c_void_p(h)
        self.grid = (nx, ny, nz)
        self.order = order
        self.nfft = nx * ny * nz
        return self

    @classmethod
    def triclinic(cls, nx, ny, nz, order, g_ewald, boxlo, boxhi, tilt, qqrd2e, prec=DOUBLE):
        """PPPMIntel::compute on a triclinic box, tilt = (xy, xz, yz); compute(x, q) takes box coordinates"""
        self = cls.__new__(cls)
        lib().orc_pppm_create_tri.restype = C.c_void_p
        h = lib().orc_pppm_create_tri(C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(order), C.c_double(g_ewald),
                                      _d(f64(boxlo)), _d(f64(boxhi)), C.c_double(tilt[0]), C.c_double(tilt[1]),
                                      C.c_double(tilt[2]), C.c_double(qqrd2e), C.c_int(prec))
        if not h:
            raise ValueError("PPPM order not supported")
        self.h = C.c_void_p(h)
        self.grid = (nx, ny, nz)
        self.order = order
        self.nfft = nx * ny * nz
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_pppm_destroy(self.h)
            self.h = None

    def compute(self, x, q, eflag=1, vflag=1, nthreads=0):
        n = len(x)
        f = np.zeros((n, 3))
        e = C.c_double(0.0)
        v = np.zeros(6)
        lib().orc_pppm_compute(self.h, C.c_int(n), _d(f64(x)), _d(f64(q)), C.c_int(eflag), C.c_int(vflag),
                               _d(f), C.byref(e), _d(v), C.c_int(nthreads))
        self._n = n
        return f, e.value, v

    def compute_arith(self, x, w7, eflag=1, vflag=1, nthreads=0):
        """function[2], arithmetic mixing: w7[n,7] = B[7 type + k] (disp_B_arithmetic(...)[type])"""
        n = len(x)
        f = np.zeros((n, 3))
        e = C.c_double(0.0)
        v = np.zeros(6)
        w7 = f64(w7)
        assert w7.shape == (n, 7)
        lib().orc_pppm_compute_arith(self.h, C.c_int(n), _d(f64(x)), _d(w7), C.c_int(eflag), C.c_int(vflag),
                                     _d(f), C.byref(e), _d(v), C.c_int(nthreads))
        return f, e.value, v

    def compute_none(self, x, wn, lam, eflag=1, vflag=1, nthreads=0):
        """function[3], no mixing rule: wn[n,nsplit] = B[nsplit type + k], lam[nsplit] eigenvalues"""
        n = len(x)
        f = np.zeros((n, 3))
        e = C.c_double(0.0)
        v = np.zeros(6)
        wn = f64(wn)
        lam = f64(lam)
        assert wn.shape == (n, len(lam))
        lib().orc_pppm_compute_none(self.h, C.c_int(n), _d(f64(x)), C.c_int(len(lam)), _d(wn), _d(lam), C.c_int(eflag),
                                    C.c_int(vflag), _d(f), C.byref(e), _d(v), C.c_int(nthreads))
        return f, e.value, v

    def peratom(self, eatom=True, vatom=True):
        """per-atom tallies of the last compute (eflag & 2 / vflag & 4)"""
        e = np.zeros(self._n) if eatom else None
        v = np.zeros((self._n, 6)) if vatom else None
        lib().orc_pppm_peratom(self.h, None if e is None else _d(e), None if v is None else _d(v))
        return e, v

    def _arr(self, ptr, n):
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    def greensfn(self):
        return self._arr(lib().orc_pppm_greensfn(self.h), self.nfft)

    def density(self):
        return self._arr(lib().orc_pppm_density_fft(self.h), self.nfft)

    def field(self, d):
        return self._arr(lib().orc_pppm_field(self.h, C.c_int(d)), self.nfft)

    def sf_coeff(self):
        return self._arr(lib().orc_pppm_sf_coeff(self.h), 6)

    def rho_coeff(self):
        a = np.zeros(self.order * self.order); b = np.zeros(self.order * self.order)
        lib().orc_pppm_rho_coeff(self.h, _d(a), _d(b))
        return a.reshape(self.order, self.order), b.reshape(self.order, self.order)


def disp_B_arithmetic(epsilon, sigma):
    """PPPMDisp::init_coeffs, function[2] [UPSTREAM, restated]: B[7 i + k] = sqrt(eps_i) / 4 * sqrt(binom(6,k)) *
    sigma_i^k for type i = 0..ntypes (epsilon[0], sigma[0] unused), so that sum_k B_i[k] B_j[6-k] = 4 sqrt(eps_i eps_j)
    ((sigma_i + sigma_j) / 2)^6 = lj4_ij of pair lj/long/coul/long with `pair_modify mix arithmetic`.  The prefactor is
    fixed by that identity (the r^-6 lattice sum must not depend on g_ewald_6, tests/test_oracle_kat.py)."""
    eps = f64(epsilon)
    sig = f64(sigma)
    c = np.sqrt([1.0, 6.0, 15.0, 20.0, 15.0, 6.0, 1.0])
    B = np.zeros((len(eps), 7))
    for i in range(len(eps)):
        B[i] = np.sqrt(eps[i]) / 4.0 * c * sig[i] ** np.arange(7)
    return B


def disp_B_none(Cij):
    """PPPMDisp::init_coeffs, function[3] [UPSTREAM, restated]: eigen-decomposition of the symmetric r^-6 coefficient
    matrix of types 1..ntypes, C = V diag(lam) V^T; returns (B[ntypes+1, nsplit] with B[i, k] = V[i-1, k], lam[nsplit]);
    eigenvalues below 1e-12 max|lam| are dropped."""
    Cij = f64(Cij)
    lam, V = np.linalg.eigh(Cij[1:, 1:])
    keep = np.abs(lam) > 1e-12 * np.abs(lam).max()
    lam, V = lam[keep], V[:, keep]
    B = np.zeros((Cij.shape[0], len(lam)))
    B[1:] = V
    return B, lam


def pppm_size(accuracy_relative, qqrd2e, qsqsum, natoms, cutoff, prd, order=5, grid=(0, 0, 0), g_ewald=0.0,
              two_charge_force=None):
    g = i32(list(grid))
    ge = C.c_double(g_ewald)
    tcf = qqrd2e if two_charge_force is None else two_charge_force
    lib().orc_pppm_size(C.c_double(accuracy_relative), C.c_double(tcf), C.c_double(qqrd2e),
                        C.c_double(qsqsum), C.c_long(natoms), C.c_double(cutoff), _d(f64(prd)),
                        C.c_int(order), C.c_int(0), _i(g), C.byref(ge))
    return tuple(int(v) for v in g), ge.value


def fft3d(a, direction):
    """a: complex128 [nz,ny,nx]"""
    a = np.ascontiguousarray(a, dtype=np.complex128).copy()
    nz, ny, nx = a.shape
    lib().orc_fft3d(a.ctypes.data_as(dp), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(direction), C.c_int(0))
    return a


def ewald_recip(x, q, boxlo, boxhi, g_ewald, kmax, qqrd2e):
    n = len(x)
    f = np.zeros((n, 3))
    e = C.c_double(0.0)
    v = np.zeros(6)
    lib().orc_ewald_recip(C.c_int(n), _d(f64(x)), _d(f64(q)), _d(f64(boxlo)), _d(f64(boxhi)),
                          C.c_double(g_ewald), C.c_int(kmax), C.c_double(qqrd2e), _d(f), C.byref(e), _d(v))
    return f, e.value, v


def ewald_recip_tri(x, q, boxlo, boxhi, tilt, g_ewald, kmax, qqrd2e):
    n = len(x)
    f = np.zeros((n, 3))
    e = C.c_double(0.0)
    v = np.zeros(6)
    lib().orc_ewald_recip_tri(C.c_int(n), _d(f64(x)), _d(f64(q)), _d(f64(boxlo)), _d(f64(boxhi)), C.c_double(tilt[0]),
                              C.c_double(tilt[1]), C.c_double(tilt[2]), C.c_double(g_ewald), C.c_int(kmax),
                              C.c_double(qqrd2e), _d(f), C.byref(e), _d(v))
    return f, e.value, v


def nve_dtfm(type_, mass, dt, ftm2v):
    n = len(type_)
    out = np.zeros(3 * n)
    lib().orc_nve_dtfm(C.c_int(n), _i(i32(type_)), _d(f64(mass)), C.c_double(dt), C.c_double(ftm2v), _d(out))
    return out


def nve_dtfm_group(type_, mass, dt, ftm2v, ingroup=None, rmass=None):
    n = len(type_)
    out = np.zeros(3 * n)
    g = None if ingroup is None else i32(ingroup)
    m = None if rmass is None else f64(rmass)
    lib().orc_nve_dtfm_group(C.c_int(n), _i(i32(type_)), _d(f64(mass)), None if m is None else _d(m),
                             None if g is None else _i(g), C.c_double(dt), C.c_double(ftm2v), _d(out))
    return out


def nve_initial_group(x, v, f, dtfm, dtv):
    x = f64(x).copy(); v = f64(v).copy()
    lib().orc_nve_initial_group(C.c_int(len(x)), _d(x), _d(v), _d(f64(f)), _d(f64(dtfm)), C.c_double(dtv))
    return x, v


def nve_initial(x, v, f, dtfm, dtv):
    x = f64(x).copy(); v = f64(v).copy()
    lib().orc_nve_initial(C.c_int(len(x)), _d(x), _d(v), _d(f64(f)), _d(f64(dtfm)), C.c_double(dtv))
    return x, v


def nve_final(v, f, dtfm):
    v = f64(v).copy()
    lib().orc_nve_final(C.c_int(len(v)), _d(v), _d(f64(f)), _d(f64(dtfm)))
    return v


class MD:
    """whole-timestep CPU loop (oracle/md.cpp) — CPU baseline and trajectory parity"""

    def __init__(self, system, params, prec=DOUBLE, skin=0.3, every=1, delay=0, check=1, dt=0.001, ftm2v=1.0, pppm=None):
        s = system
        n = len(s["x"])
        self.n = n
        self.params, self.pppm = params, pppm
        q = f64(s["q"]) if s.get("q") is not None else None
        lib().orc_md_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().orc_md_create(
            C.c_int(n), _d(f64(s["x"])), _d(f64(s["v"])) if s.get("v") is not None else None,
            _d(q) if q is not None else None, _i(i32(s["type"])), C.c_int(s["ntypes"]), _d(f64(s["mass"])),
            _d(f64(s["boxlo"])), _d(f64(s["boxhi"])), C.c_int(params.style), C.c_int(prec), C.byref(params.p),
            C.c_double(skin), C.c_int(every), C.c_int(delay), C.c_int(check), C.c_double(dt), C.c_double(ftm2v),
            pppm.h if pppm is not None else None))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_md_destroy(self.h)
            self.h = None

    def run(self, nsteps, nthreads=0):
        timers = np.zeros(8)
        nb = C.c_int(0)
        lib().orc_md_run(self.h, C.c_int(nsteps), C.c_int(nthreads), _d(timers), C.byref(nb))
        if lib().orc_md_lost(self.h):
            raise RuntimeError("oracle MD: an atom left the box by more than one period (the model blew up)")
        return dict(neigh=timers[0], pair=timers[1], kspace=timers[2], nve=timers[3], comm=timers[4], nbuilds=nb.value)

    def get(self):
        x = np.zeros((self.n, 3)); v = np.zeros((self.n, 3)); f = np.zeros((self.n, 3))
        lib().orc_md_get(self.h, _d(x), _d(v), _d(f))
        return x, v, f

    def energy(self, nthreads=0):
        ev = np.zeros(8)
        ek = C.c_double(0.0); ke = C.c_double(0.0)
        lib().orc_md_energy(self.h, C.c_int(nthreads), _d(ev), C.byref(ek), C.byref(ke))
        return ev, ek.value, ke.value
