"""b200md — Python (ctypes) binding of the C ABI in include/b200md.h.

This is the harness-side mirror of the reference's plug-in surface: thin wrappers named after the reference
classes (`PairBuck*Intel.compute`, `PPPMIntel.compute`, `FixNVEIntel`) over the C entry points.  It holds no
compute: every number comes from libb200md.so (hand-written sm_100a kernels).  There is no CPU fallback —
`load()` raises if the library is missing, and `Context()` raises if no B200 is visible.

The directory name contains a hyphen, so import it through `__graft_entry__.load_package()` (or
tests/conftest.py), which registers it as the module `lammps_buck_intel_b200`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# B200MD_LIB selects an alternative build of the same library (kernel-tuning experiments: scratch/variants.py)
LIBPATH = os.environ.get("B200MD_LIB") or os.path.join(HERE, "libb200md.so")
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(ROOT, "include", "b200md.h")
HEADER_TESTING = os.path.join(ROOT, "include", "b200md_testing.h")

PREC_DOUBLE, PREC_MIXED = 0, 1
PAIR_BUCK, PAIR_BUCK_COUL_CUT, PAIR_BUCK_COUL_LONG, PAIR_BUCK_LONG_COUL_LONG = 0, 1, 2, 3
PAIR_LJ_LONG_COUL_LONG = 4   # pair_coeffs: A = epsilon, rho = sigma
SBBITS = 30
NEIGHMASK = 0x3FFFFFFF

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# the float (mixed-mode) pair kernels replay the reference's un-fused IEEE arithmetic (csrc/pair_kernel.cuh)
PER_FILE_FLAGS = {"pair_mixed.cu": ["-fmad=false"]}

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_long)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, verbose=False):
    """nvcc -> lammps-buck-intel_b200/libb200md.so (in-tree, so it travels to the GPU box).
    Each .cu is compiled to build/<name>.o (in parallel, only when stale), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [HEADER, HEADER_TESTING]
    hdr_m = max(os.path.getmtime(h) for h in hdrs)
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + extra_compile_flags()
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            jobs.append([nvcc] + cflags + PER_FILE_FLAGS.get(os.path.basename(src), []) + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIBPATH) or any(os.path.getmtime(o) > os.path.getmtime(LIBPATH) for o in objs):
        run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIBPATH] + objs + extra_link_flags())
    return LIBPATH


HOST = os.path.join(HERE, "host")
LMP = os.path.join(HERE, "lmp_b200")


def build_host(force=False, verbose=False):
    """g++ -> lammps-buck-intel_b200/lmp_b200: the C++ host classes (host/*.cpp, the reference's class surface) and the
    input-script driver, linked against libb200md.so."""
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [HEADER, LIBPATH]
    if not force and os.path.exists(LMP) and os.path.getmtime(LMP) >= max(os.path.getmtime(d) for d in deps):
        return LMP
    cmd = ["g++", "-O2", "-fopenmp", "-std=c++17", "-Wall", "-o", LMP] + srcs + \
          ["-L", HERE, "-l:libb200md.so", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n" + r.stdout + r.stderr)
    return LMP


def nccl_paths():
    """(include dir, lib dir) of NCCL: the torch-bundled copy when present (so one libnccl.so.2 serves both torch's
    bootstrap and this library in the same process), else the system one."""
    import sysconfig
    sp = sysconfig.get_paths()["purelib"]
    inc, lib = os.path.join(sp, "nvidia", "nccl", "include"), os.path.join(sp, "nvidia", "nccl", "lib")
    if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(os.path.join(lib, "libnccl.so.2")):
        return inc, lib
    return "/usr/include", "/usr/lib/x86_64-linux-gnu"


def extra_compile_flags():
    return ["-I", nccl_paths()[0]]


def extra_link_flags():
    lib = nccl_paths()[1]
    return ["-L", lib, "-l:libnccl.so.2", "-Xlinker", "-rpath", "-Xlinker", lib]


class PairParams(C.Structure):
    _fields_ = [
        ("style", C.c_int), ("ntypes", C.c_int),
        ("cutsq", dp), ("cut_ljsq", dp), ("cut_coulsq", dp),
        ("buck1", dp), ("buck2", dp), ("rhoinv", dp), ("a", dp), ("c", dp), ("offset", dp),
        ("special_lj", C.c_double * 4), ("special_coul", C.c_double * 4),
        ("g_ewald", C.c_double), ("g_ewald_6", C.c_double), ("ewald_order", C.c_int),
        ("ncoultablebits", C.c_int), ("ncoulmask", C.c_int), ("ncoulshiftbits", C.c_int),
        ("tabinnersq", C.c_double),
        ("rtable", dp), ("drtable", dp), ("ftable", dp), ("dftable", dp),
        ("etable", dp), ("detable", dp), ("ctable", dp), ("dctable", dp),
        ("ndisptablebits", C.c_int), ("ndispmask", C.c_int), ("ndispshiftbits", C.c_int),
        ("tabinnerdispsq", C.c_double),
        ("rdisptable", dp), ("drdisptable", dp), ("fdisptable", dp), ("dfdisptable", dp),
        ("edisptable", dp), ("dedisptable", dp),
    ]


class PppmParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("order", C.c_int),
                ("g_ewald", C.c_double), ("differentiation", C.c_int), ("scale", C.c_double),
                ("dispersion", C.c_int), ("B", dp), ("slab_volfactor", C.c_double)]


_lib = None


def load():
    """Load libb200md.so; fail loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise RuntimeError("libb200md.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIBPATH, mode=C.RTLD_GLOBAL)
    lib.b200md_last_error.restype = C.c_char_p
    lib.b200md_last_error.argtypes = [C.c_void_p]
    lib.b200md_timer_name.restype = C.c_char_p
    lib.b200md_launch_count.restype = C.c_long
    lib.b200md_launch_count.argtypes = [C.c_void_p]
    lib.b200md_ctx_destroy.argtypes = [C.c_void_p]
    lib.b200md_ctx_destroy.restype = None
    _lib = lib
    return lib


class B200MDError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200md error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def _d(a):
    return a.ctypes.data_as(dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(ip) if a is not None else None


def _l(a):
    return a.ctypes.data_as(lp) if a is not None else None


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


# --------------------------------------------------------------------------------------------------
# host-side parameter logic of the stock base classes (PairBuck*::init_one, Pair::init_tables)

def pair_coeffs(style, ntypes, A, rho, Cc, cut_lj, cut_coul=None, offset_flag=0):
    """init_one for every type pair -> dict of (ntypes+1)^2 arrays the C ABI takes."""
    tp1 = ntypes + 1

    def full(v):
        v = np.asarray(v, dtype=np.float64)
        return np.full((tp1, tp1), float(v)) if v.ndim == 0 else v.reshape(tp1, tp1).copy()

    A, rho, Cc, cut_lj = full(A), full(rho), full(Cc), full(cut_lj)
    cut_coul = full(cut_coul) if cut_coul is not None else np.zeros((tp1, tp1))
    rho = np.where(rho == 0.0, 1.0, rho)
    cl = np.where(cut_lj > 0, cut_lj, 1.0)
    if style == PAIR_LJ_LONG_COUL_LONG:
        # PairLJLongCoulLong::init_one [UPSTREAM]: A = epsilon, rho = sigma; buck1, buck2, a, c carry lj1, lj2, lj3, lj4
        # (pair_lj_long_coul_long_intel.cpp:831-834)
        out = dict(buck1=48.0 * A * rho ** 12.0, buck2=24.0 * A * rho ** 6.0, a=4.0 * A * rho ** 12.0,
                   c=4.0 * A * rho ** 6.0, rhoinv=np.zeros((tp1, tp1)))
        ratio = rho / cl
        out["offset"] = np.where(cut_lj > 0, 4.0 * A * (ratio ** 12.0 - ratio ** 6.0), 0.0) if offset_flag else \
            np.zeros((tp1, tp1))
    else:
        out = dict(a=A, c=Cc, rhoinv=1.0 / rho, buck1=A / rho, buck2=6.0 * Cc)
        out["offset"] = (A * np.exp(-cl / rho) - Cc / cl ** 6) if offset_flag else np.zeros((tp1, tp1))
    out["cut_ljsq"] = cut_lj * cut_lj
    out["cut_coulsq"] = cut_coul * cut_coul
    cut = cut_lj if style == PAIR_BUCK else np.maximum(cut_lj, cut_coul)
    out["cutsq"] = cut * cut
    for k in out:
        out[k][0, :] = 0.0
        out[k][:, 0] = 0.0
        out[k] = np.ascontiguousarray(out[k])
    return out


def init_coul_tables(cut_coul, g_ewald, qqrd2e, nbits=12, tabinner=np.sqrt(2.0)):
    """Pair::init_bitmap + Pair::init_tables (no rRESPA, no MSM) -> (tables dict, mask, shift, tabinnersq)."""
    from math import erfc as _erfc, exp as _exp   # libm, as the C++ host classes: the two builders are bit-identical
    EWALD_F = 1.12837917
    inner, outer = float(tabinner), float(cut_coul)
    nlowermin = 1
    while not (2.0 ** nlowermin <= inner * inner < 2.0 ** (nlowermin + 1)):
        nlowermin += 1 if 2.0 ** nlowermin <= inner * inner else -1
    nexpbits = 0
    required = outer * outer / 2.0 ** nlowermin
    available = 2.0
    while available < required:
        nexpbits += 1
        available = 2.0 ** (2.0 ** nexpbits)
    nmantbits = nbits - nexpbits
    nshiftbits = 24 - (nmantbits + 1)
    nmask = (1 << (nbits + nshiftbits)) - 1
    f2i = lambda f: int(np.float32(f).view(np.int32))
    i2f = lambda i: float(np.int32(i).view(np.float32))
    maskhi = f2i(outer * outer) & ~nmask
    masklo = f2i(inner * inner) & ~nmask
    ntable = 1 << nbits
    tabinnersq = inner * inner
    t = {k: np.zeros(ntable) for k in ("r", "dr", "f", "df", "e", "de", "c", "dc")}
    minrsq = i2f(maskhi)
    for i in range(ntable):
        bits = (i << nshiftbits) | masklo
        if i2f(bits) < tabinnersq:
            bits = (i << nshiftbits) | maskhi
        rsq = i2f(bits)
        r = float(np.sqrt(np.float32(rsq)))
        grij = g_ewald * r
        expm2 = _exp(-grij * grij)
        derfc = _erfc(grij)
        t["r"][i] = rsq
        t["c"][i] = qqrd2e / r
        t["f"][i] = qqrd2e / r * (derfc + EWALD_F * grij * expm2)
        t["e"][i] = qqrd2e / r * derfc
        minrsq = min(minrsq, rsq)
    tabinnersq = minrsq
    for a, b in (("dr", "r"), ("df", "f"), ("dc", "c"), ("de", "e")):
        nxt = np.roll(t[b], -1)
        t[a] = (1.0 / (nxt - t[b])) if a == "dr" else (nxt - t[b])
    itablemin = (f2i(minrsq) & nmask) >> nshiftbits
    itablemax = itablemin - 1 if itablemin != 0 else ntable - 1
    top = i2f((itablemax << nshiftbits) | maskhi)
    cut_coulsq = cut_coul * cut_coul
    if top < cut_coulsq:
        rsq = float(np.float32(cut_coulsq))
        r = float(np.sqrt(np.float32(rsq)))
        grij = g_ewald * r
        expm2 = _exp(-grij * grij)
        derfc = _erfc(grij)
        t["dr"][itablemax] = 1.0 / (rsq - t["r"][itablemax])
        t["df"][itablemax] = qqrd2e / r * (derfc + EWALD_F * grij * expm2) - t["f"][itablemax]
        t["dc"][itablemax] = qqrd2e / r - t["c"][itablemax]
        t["de"][itablemax] = qqrd2e / r * derfc - t["e"][itablemax]
    for k in t:
        t[k] = np.ascontiguousarray(t[k])
    return t, nmask, nshiftbits, tabinnersq


def _bitmap(inner, outer, nbits):
    nlowermin = 1
    while not (2.0 ** nlowermin <= inner * inner < 2.0 ** (nlowermin + 1)):
        nlowermin += 1 if 2.0 ** nlowermin <= inner * inner else -1
    nexpbits = 0
    required = outer * outer / 2.0 ** nlowermin
    available = 2.0
    while available < required:
        nexpbits += 1
        available = 2.0 ** (2.0 ** nexpbits)
    nshiftbits = 24 - (nbits - nexpbits + 1)
    nmask = (1 << (nbits + nshiftbits)) - 1
    f2i = lambda f: int(np.float32(f).view(np.int32))
    return f2i(inner * inner) & ~nmask, f2i(outer * outer) & ~nmask, nmask, nshiftbits


def init_disp_tables(cut_lj, g_ewald_6, nbits=12, tabinner=np.sqrt(2.0)):
    """Pair::init_tables_disp (dispersion lookup of buck/long/coul/long, pair_buck_long_coul_long_intel.cpp:433-454)
    -> (tables dict {r,dr,f,df,e,de}, mask, shift, tabinnerdispsq)"""
    i2f = lambda i: float(np.int32(i).view(np.float32))
    f2i = lambda f: int(np.float32(f).view(np.int32))
    masklo, maskhi, nmask, nshiftbits = _bitmap(float(tabinner), float(cut_lj), nbits)
    g2 = g_ewald_6 * g_ewald_6
    g6 = g2 ** 3
    g8 = g6 * g2
    ntable = 1 << nbits
    tabinnersq = float(tabinner) ** 2

    def ev(rsq):
        from math import exp as _exp
        x2 = g2 * rsq
        a2 = 1.0 / x2
        x2 = a2 * _exp(-x2)
        return g8 * (((6.0 * a2 + 6.0) * a2 + 3.0) * a2 + 1.0) * x2 * rsq, g6 * ((a2 + 1.0) * a2 + 0.5) * x2

    t = {k: np.zeros(ntable) for k in ("r", "dr", "f", "df", "e", "de")}
    minrsq = i2f(maskhi)
    for i in range(ntable):
        bits = (i << nshiftbits) | masklo
        if i2f(bits) < tabinnersq:
            bits = (i << nshiftbits) | maskhi
        rsq = i2f(bits)
        t["r"][i] = rsq
        t["f"][i], t["e"][i] = ev(rsq)
        minrsq = min(minrsq, rsq)
    for a, b in (("dr", "r"), ("df", "f"), ("de", "e")):
        nxt = np.roll(t[b], -1)
        t[a] = (1.0 / (nxt - t[b])) if a == "dr" else (nxt - t[b])
    itablemin = (f2i(minrsq) & nmask) >> nshiftbits
    itablemax = itablemin - 1 if itablemin != 0 else ntable - 1
    top = i2f((itablemax << nshiftbits) | maskhi)
    if top < cut_lj * cut_lj:
        rsq = float(np.float32(cut_lj * cut_lj))
        fo, eo = ev(rsq)
        t["dr"][itablemax] = 1.0 / (rsq - t["r"][itablemax])
        t["df"][itablemax] = fo - t["f"][itablemax]
        t["de"][itablemax] = eo - t["e"][itablemax]
    for k in t:
        t[k] = np.ascontiguousarray(t[k])
    return t, nmask, nshiftbits, minrsq


_ACONS = {
    1: [2.0 / 3.0],
    2: [1.0 / 50.0, 5.0 / 294.0],
    3: [1.0 / 588.0, 7.0 / 1440.0, 21.0 / 3872.0],
    4: [1.0 / 4320.0, 3.0 / 1936.0, 7601.0 / 2271360.0, 143.0 / 28800.0],
    5: [1.0 / 23232.0, 7601.0 / 13628160.0, 143.0 / 69120.0, 517231.0 / 106536960.0, 106640677.0 / 11737571328.0],
    6: [691.0 / 68140800.0, 13.0 / 57600.0, 47021.0 / 35512320.0, 9694607.0 / 2095994880.0,
        733191589.0 / 59609088000.0, 326190917.0 / 11700633600.0],
    7: [1.0 / 345600.0, 3617.0 / 35512320.0, 745739.0 / 838397952.0, 56399353.0 / 12773376000.0,
        25091609.0 / 1560084480.0, 1755948832039.0 / 36229939200000.0, 4887769399.0 / 37838389248.0],
}


def pppm_init(accuracy_relative, qqrd2e, q, natoms, cutoff, prd, order=5, mesh=None, gewald=None,
              two_charge_force=None, slab=1.0):
    """PPPM::init -> set_grid_global -> adjust_gewald for ik differentiation: returns
    ((nx,ny,nz), g_ewald).  `q` is the charge array (qsqsum = sum q^2) or qsqsum itself.
    mesh / gewald mimic `kspace_modify mesh` / `kspace_modify gewald`."""
    from math import exp, log, sqrt, pi
    qsqsum = float(np.sum(np.asarray(q, dtype=np.float64) ** 2)) if np.ndim(q) else float(q)
    tcf = qqrd2e if two_charge_force is None else two_charge_force
    accuracy = accuracy_relative * tcf
    q2 = qsqsum * qqrd2e
    xprd, yprd, zprd = (float(v) for v in prd)

    def est(h, p, g):
        hg = h * g
        ssum = sum(a * hg ** (2.0 * m) for m, a in enumerate(_ACONS[order]))
        return q2 * hg ** order * sqrt(g * p * sqrt(2.0 * pi) * ssum / natoms) / (p * p)

    g = gewald
    if g is None:
        g = accuracy * sqrt(natoms * cutoff * xprd * yprd * zprd) / (2.0 * q2)
        g = (1.35 - 0.15 * log(accuracy)) / cutoff if g >= 1.0 else sqrt(-log(g)) / cutoff
    if mesh is None:
        n = []
        for p in (xprd, yprd, zprd * slab):   # zprd_slab (kspace_modify slab)
            h = 1.0 / g
            k = int(p / h) + 1
            err = est(h, p, g)
            while err > accuracy:
                err = est(h, p, g)
                k += 1
                h = p / k
            n.append(k)
    else:
        n = list(mesh)

    def factorable(k):
        for f in (2, 3, 5):
            while k % f == 0:
                k //= f
        return k == 1

    n = [next(k for k in range(v, 100000) if factorable(k)) for v in n]
    hs = [xprd / n[0], yprd / n[1], zprd * slab / n[2]]
    if gewald is None:
        def nr_f(gg):
            df_r = 2.0 * q2 * exp(-gg * gg * cutoff * cutoff) / sqrt(natoms * cutoff * xprd * yprd * zprd)
            l = [est(hs[0], xprd, gg), est(hs[1], yprd, gg), est(hs[2], zprd * slab, gg)]
            return df_r - sqrt(l[0] ** 2 + l[1] ** 2 + l[2] ** 2) / sqrt(3.0)
        for _ in range(10000):
            f1, f2 = nr_f(g), nr_f(g + 1e-6)
            g -= f1 / ((f2 - f1) / 1e-6)
            if abs(nr_f(g)) < 1e-5:
                break
    return tuple(n), g


class Context:
    """One device context (= FixIntel + IntelBuffers + the resident atom state)."""

    def __init__(self, device=0, precision=PREC_DOUBLE):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.b200md_ctx_create(C.c_int(device), C.c_int(precision), C.byref(h))
        if rc != 0:
            raise B200MDError(rc, self.lib.b200md_last_error(None).decode())
        self.h = h
        self.precision = precision
        self.nlocal = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200md_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise B200MDError(rc, self.lib.b200md_last_error(self.h).decode())

    # ---- setup ---------------------------------------------------------------------------------
    def set_units(self, qqrd2e, ftm2v):
        self._ck(self.lib.b200md_set_units(self.h, C.c_double(qqrd2e), C.c_double(ftm2v)))

    def set_box(self, boxlo, boxhi, periodic=(1, 1, 1)):
        self._ck(self.lib.b200md_set_box(self.h, _d(f64(boxlo)), _d(f64(boxhi)), _i(i32(periodic))))

    def set_box_triclinic(self, boxlo, boxhi, tilt):
        """triclinic box, tilt = (xy, xz, yz): k-space solver only (pppm_setup / pppm_compute / pppm_compute_host)"""
        self._ck(self.lib.b200md_set_box_triclinic(self.h, _d(f64(boxlo)), _d(f64(boxhi)), C.c_double(tilt[0]),
                                                   C.c_double(tilt[1]), C.c_double(tilt[2])))

    def atoms_upload(self, x, type_, mass, v=None, q=None):
        x = f64(x); v = f64(v); q = f64(q); type_ = i32(type_); mass = f64(mass)
        self.nlocal = len(x)
        self._ck(self.lib.b200md_atoms_upload(self.h, C.c_int(len(x)), C.c_int(len(mass) - 1), _d(x), _d(v), _d(q),
                                              _i(type_), _d(mass)))

    def atoms_set_special(self, nspecial=None, special=None):
        """special bonds of a molecular system for device-built lists: nspecial [n,3] cumulative counts (1-2, 1-3, 1-4),
        special [n,maxspecial] partner upload indices; None clears"""
        if nspecial is None:
            self._ck(self.lib.b200md_atoms_set_special(self.h, C.c_int(0), None, None))
            return
        ns = i32(nspecial)
        sp = i32(special)
        self._ck(self.lib.b200md_atoms_set_special(self.h, C.c_int(sp.shape[1]), _i(ns), _i(sp)))

    def atoms_set_x(self, x):
        self._ck(self.lib.b200md_atoms_set_x(self.h, _d(f64(x))))

    def atoms_download(self, want=("x", "v", "f")):
        n = self.nlocal
        out = {k: np.zeros((n, 3)) for k in want if k in ("x", "v", "f")}
        if "eatom" in want:
            out["eatom"] = np.zeros(n)
        self._ck(self.lib.b200md_atoms_download(self.h, _d(out.get("x")), _d(out.get("v")), _d(out.get("f")),
                                                _d(out.get("eatom"))))
        return out

    def pair_setup(self, style, ntypes, coeffs, special_lj=(1, 0, 0, 0), special_coul=(1, 0, 0, 0), g_ewald=0.0,
                   g_ewald_6=0.0, ewald_order=0, coul_tables=None, disp_tables=None):
        p = PairParams()
        p.style, p.ntypes = style, ntypes
        keep = {k: f64(v) for k, v in coeffs.items()}
        for k in ("cutsq", "cut_ljsq", "cut_coulsq", "buck1", "buck2", "rhoinv", "a", "c", "offset"):
            setattr(p, k, _d(keep[k]))
        for i in range(4):
            p.special_lj[i] = special_lj[i]
            p.special_coul[i] = special_coul[i]
        p.g_ewald, p.g_ewald_6, p.ewald_order = g_ewald, g_ewald_6, ewald_order
        if coul_tables is not None:
            t, mask, shift, inner = coul_tables
            t = {k: f64(v) for k, v in t.items()}
            keep["ct"] = t
            p.ncoultablebits = int(np.log2(len(t["r"])))
            p.ncoulmask, p.ncoulshiftbits, p.tabinnersq = mask, shift, inner
            p.rtable, p.drtable, p.ftable, p.dftable = _d(t["r"]), _d(t["dr"]), _d(t["f"]), _d(t["df"])
            p.etable, p.detable, p.ctable, p.dctable = _d(t["e"]), _d(t["de"]), _d(t["c"]), _d(t["dc"])
        if disp_tables is not None:
            t, mask, shift, inner = disp_tables
            t = {k: f64(v) for k, v in t.items()}
            keep["dt"] = t
            p.ndisptablebits = int(np.log2(len(t["r"])))
            p.ndispmask, p.ndispshiftbits, p.tabinnerdispsq = mask, shift, inner
            p.rdisptable, p.drdisptable, p.fdisptable, p.dfdisptable = _d(t["r"]), _d(t["dr"]), _d(t["f"]), _d(t["df"])
            p.edisptable, p.dedisptable = _d(t["e"]), _d(t["de"])
        self._ck(self.lib.b200md_pair_setup(self.h, C.byref(p)))

    def neigh_setup(self, skin, every=1, delay=0, check=1):
        self._ck(self.lib.b200md_neigh_setup(self.h, C.c_double(skin), C.c_int(every), C.c_int(delay), C.c_int(check)))

    def neigh_build(self):
        self._ck(self.lib.b200md_neigh_build(self.h))

    def neigh_decide(self, ntimestep=0):
        r = C.c_int(0)
        self._ck(self.lib.b200md_neigh_decide(self.h, C.c_long(ntimestep), C.byref(r)))
        return r.value

    def neigh_stats(self):
        tot, ng, mx, nb = C.c_long(), C.c_int(), C.c_int(), C.c_long()
        self._ck(self.lib.b200md_neigh_stats(self.h, C.byref(tot), C.byref(ng), C.byref(mx), C.byref(nb)))
        return dict(total=tot.value, nghost=ng.value, max_numneigh=mx.value, nbuilds=nb.value)

    def neigh_download(self):
        st = self.neigh_stats()
        n = self.nlocal
        numneigh = np.zeros(n, np.int32)
        offsets = np.zeros(n + 1, np.int64)
        entries = np.zeros(max(st["total"], 1), np.int32)
        gsrc = np.zeros(max(st["nghost"], 1), np.int32)
        gshift = np.zeros((max(st["nghost"], 1), 3), np.int32)
        self._ck(self.lib.b200md_neigh_download(self.h, _i(numneigh), _l(offsets), _i(entries), _i(gsrc), _i(gshift)))
        return numneigh, offsets, entries[:st["total"]], gsrc[:st["nghost"]], gshift[:st["nghost"]]

    # ---- Pair*Intel::compute --------------------------------------------------------------------
    def pair_compute(self, eflag=0, vflag=0):
        ev = np.zeros(8)
        self._ck(self.lib.b200md_pair_compute(self.h, C.c_int(eflag), C.c_int(vflag), _d(ev)))
        return ev

    def pair_eval_host(self, eflag, vflag, nlocal, x, type_, q, numneigh, cnumneigh, firstneigh):
        x = f64(x); type_ = i32(type_); q = f64(q)
        f = np.zeros((nlocal, 4))
        ev = np.zeros(8)
        self._ck(self.lib.b200md_pair_eval_host(self.h, C.c_int(eflag), C.c_int(vflag), C.c_int(nlocal),
                                                C.c_int(len(x)), _d(x), _i(type_), _d(q), _i(i32(numneigh)),
                                                _l(np.ascontiguousarray(cnumneigh, np.int64)), _i(i32(firstneigh)),
                                                _d(f), _d(ev)))
        return f, ev

    # ---- PPPMIntel --------------------------------------------------------------------------------
    def pppm_setup(self, nx, ny, nz, order, g_ewald, differentiation=0, scale=1.0, dispersion=0, B=None, slab=0.0):
        p = PppmParams()
        p.nx, p.ny, p.nz, p.order, p.g_ewald = nx, ny, nz, order, g_ewald
        p.differentiation, p.scale, p.dispersion = differentiation, scale, dispersion
        p.slab_volfactor = slab
        Bk = f64(B)
        p.B = _d(Bk)
        self._ck(self.lib.b200md_pppm_setup(self.h, C.byref(p)))
        self.pppm_grid = (nx, ny, nz)

    def pppm_peratom(self, eatom=True, vatom=True):
        """per-atom k-space energy [n] / virial [n,6] of the last pppm_compute(eflag & 2, vflag & 4), upload order"""
        n = self.nlocal
        e = np.zeros(n) if eatom else None
        v = np.zeros((n, 6)) if vatom else None
        self._ck(self.lib.b200md_pppm_peratom(self.h, None if e is None else _d(e), None if v is None else _d(v)))
        return e, v

    def pppm_compute(self, eflag=0, vflag=0):
        e = C.c_double(0.0)
        v = np.zeros(6)
        self._ck(self.lib.b200md_pppm_compute(self.h, C.c_int(eflag), C.c_int(vflag), C.byref(e), _d(v)))
        return e.value, v

    def pppm_compute_host(self, x, q, eflag=0, vflag=0):
        x = f64(x); q = f64(q)
        f = np.zeros((len(x), 3))
        e = C.c_double(0.0)
        v = np.zeros(6)
        self._ck(self.lib.b200md_pppm_compute_host(self.h, C.c_int(eflag), C.c_int(vflag), C.c_int(len(x)), _d(x),
                                                   _d(q), _d(f), C.byref(e), _d(v)))
        return f, e.value, v

    def pppm_download(self):
        nx, ny, nz = self.pppm_grid
        n = nx * ny * nz
        out = dict(density=np.zeros(n), greensfn=np.zeros(n), fx=np.zeros(n), fy=np.zeros(n), fz=np.zeros(n),
                   sf_coeff=np.zeros(6))
        self._ck(self.lib.b200md_pppm_download(self.h, _d(out["density"]), _d(out["greensfn"]), _d(out["fx"]),
                                               _d(out["fy"]), _d(out["fz"]), _d(out["sf_coeff"])))
        return out

    def fft3d(self, a, direction):
        a = np.ascontiguousarray(a, dtype=np.complex128).copy()
        nz, ny, nx = a.shape
        self._ck(self.lib.b200md_fft3d_host(self.h, a.ctypes.data_as(dp), C.c_int(nx), C.c_int(ny), C.c_int(nz),
                                            C.c_int(direction)))
        return a

    # ---- FixNVEIntel / Verlet -----------------------------------------------------------------------
    def nve_setup(self, dt):
        self._ck(self.lib.b200md_nve_setup(self.h, C.c_double(dt)))

    def nve_set_group(self, ingroup=None, rmass=None):
        """fix nve on a sub-group (0/1 per atom) and / or with per-atom masses, upload order; before nve_setup"""
        g = None if ingroup is None else np.ascontiguousarray(ingroup, dtype=np.int32)
        m = None if rmass is None else np.ascontiguousarray(rmass, dtype=np.float64)
        self._keep = (g, m)
        self._ck(self.lib.b200md_nve_set_group(
            self.h, None if g is None else g.ctypes.data_as(C.POINTER(C.c_int)),
            None if m is None else m.ctypes.data_as(C.POINTER(C.c_double))))

    def nve_initial_integrate(self):
        self._ck(self.lib.b200md_nve_initial_integrate(self.h))

    def nve_final_integrate(self):
        self._ck(self.lib.b200md_nve_final_integrate(self.h))

    def setup_forces(self, eflag=1, vflag=1):
        th = np.zeros(16)
        self._ck(self.lib.b200md_setup_forces(self.h, C.c_int(eflag), C.c_int(vflag), _d(th)))
        return th

    def run(self, nsteps, thermo=False):
        th = np.zeros(16) if thermo else None
        self._ck(self.lib.b200md_run(self.h, C.c_long(nsteps), _d(th)))
        return th

    def run_timed(self, nsteps):
        ms = C.c_double(0.0)
        self._ck(self.lib.b200md_run_timed(self.h, C.c_long(nsteps), None, C.byref(ms)))
        return ms.value

    def step_host(self, x_in, x_out, f_out):
        """x_in/x_out/f_out: C-contiguous float64 [n,3] numpy arrays (pinned for speed) or None"""
        self._ck(self.lib.b200md_step_host(self.h, _d(x_in), _d(x_out), _d(f_out)))

    def step_host_ids(self, ids, x_out, f_out):
        """one step on any number of GPUs; ids int32 [cap], x_out / f_out float64 [cap,3] (pinned for speed) are filled
        in device order for the atoms this rank owns after the step; returns that count"""
        n = C.c_int(0)
        self._ck(self.lib.b200md_step_host_ids(self.h, C.c_int(len(ids)), C.byref(n), _i(ids), _d(x_out), _d(f_out)))
        return n.value

    # ---- multi-GPU ----------------------------------------------------------------------------------
    def comm_init(self, rank, nranks, unique_id=None):
        """unique_id: the 128 bytes from comm_unique_id() on rank 0, broadcast by the caller.  Call before
        atoms_upload / pppm_setup (global atom ids and the grid decomposition depend on it)."""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
        self._ck(self.lib.b200md_comm_init(self.h, C.c_int(rank), C.c_int(nranks), buf))
        self.rank, self.nranks = rank, nranks

    def comm_unique_id(self):
        buf = (C.c_ubyte * 128)()
        rc = self.lib.b200md_comm_unique_id(buf)
        if rc != 0:
            raise B200MDError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def comm_init_torch(self, dist, rank, nranks):
        """bootstrap through torch.distributed: rank 0 creates the ncclUniqueId, everyone gets it by broadcast"""
        obj = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        self.comm_init(rank, nranks, obj[0])

    def atoms_download_ids(self, want=("x", "v", "f")):
        """owned atoms of this rank in device order + their global ids (multi-GPU: atoms migrate between ranks)"""
        nl, ng = C.c_int(), C.c_int()
        self._ck(self.lib.b200md_atoms_count(self.h, C.byref(nl), C.byref(ng)))
        cap = nl.value + 16
        ids = np.zeros(cap, np.int32)
        out = {k: np.zeros((cap, 3)) for k in want}
        n = C.c_int()
        self._ck(self.lib.b200md_atoms_download_ids(self.h, C.c_int(cap), C.byref(n), _i(ids), _d(out.get("x")),
                                                    _d(out.get("v")), _d(out.get("f"))))
        res = {k: v[:n.value] for k, v in out.items()}
        res["ids"] = ids[:n.value]
        return res

    # ---- timers -------------------------------------------------------------------------------------
    def timers_enable(self, on=True):
        self._ck(self.lib.b200md_timers_enable(self.h, C.c_int(1 if on else 0)))

    def timers_reset(self):
        self._ck(self.lib.b200md_timers_reset(self.h))

    def timers(self):
        n = self.lib.b200md_timer_count()
        ms = np.zeros(n)
        calls = np.zeros(n, np.int64)
        self._ck(self.lib.b200md_timers_get(self.h, _d(ms), _l(calls), C.c_int(n)))
        return {self.lib.b200md_timer_name(C.c_int(i)).decode(): (float(ms[i]), int(calls[i])) for i in range(n)}

    def microbench(self, kind):
        v = C.c_double(0.0)
        self._ck(self.lib.b200md_microbench(self.h, C.c_int(kind), C.byref(v)))
        return v.value

    def launch_count(self):
        return int(self.lib.b200md_launch_count(self.h))


def make_context(system, precision=PREC_DOUBLE, device=0):
    """Context with units, box and atoms of a workloads.* system dict."""
    from . import workloads as W  # noqa: F401  (resolved through the registered package name)
    u = W.UNITS[system["units"]]
    ctx = Context(device, precision)
    ctx.set_units(u["qqrd2e"], u["ftm2v"])
    if "periodic" in system:
        ctx.set_box(system["boxlo"], system["boxhi"], system["periodic"])
    else:
        ctx.set_box(system["boxlo"], system["boxhi"])
    ctx.atoms_upload(system["x"], system["type"], system["mass"], v=system.get("v"), q=system.get("q"))
    return ctx
