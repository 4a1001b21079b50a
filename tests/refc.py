"""ctypes wrapper of oracle/_ref/libref.so — the reference's OWN translation units, compiled unchanged from
/root/reference against the stand-in headers of oracle/ref_shim/ (recipe: oracle/Makefile.ref).  TEST INFRASTRUCTURE ONLY.

build() compiles the library where /root/reference is present (this container); on the GPU box the prebuilt
oracle/_ref/libref.so travels with the snapshot and is only loaded.  available() says whether it can be used.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ODIR, "_ref", "libref.so")
REFDIR = os.environ.get("B200MD_REFERENCE", "/root/reference")

_lib = None


def build(force=False):
    """returns the library path, or None when neither the reference sources nor a prebuilt library exist"""
    if not os.path.isdir(REFDIR):
        return LIB if os.path.exists(LIB) else None
    deps = [os.path.join(ODIR, "ref_harness.cpp"), os.path.join(ODIR, "Makefile.ref"), os.path.join(ODIR, "oracle.h"),
            os.path.join(ODIR, "fft.cpp")]
    sh = os.path.join(ODIR, "ref_shim")
    deps += [os.path.join(sh, f) for f in os.listdir(sh)]
    deps += [os.path.join(REFDIR, f) for f in os.listdir(REFDIR) if f.endswith((".cpp", ".h"))]
    stale = force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)
    if stale:
        r = subprocess.run(["make", "-C", ODIR, "-f", "Makefile.ref", "REF=" + REFDIR], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB


def available():
    try:
        return build() is not None
    except RuntimeError:
        return False


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref/libref.so is missing and /root/reference is not present")
        _lib = C.CDLL(path)
    return _lib


def _err():
    return C.create_string_buffer(512)


def pair_eval(params, prec, eflag, vflag, nlocal, x, type_, q, numneigh, offsets, entries, newton=1, eatom=0,
              nthreads=1, skin=0.0, offset_flag=0):
    """PairBuck*Intel::compute of the reference on a packed list; same call shape as orc.pair_eval"""
    nall = len(x)
    x = orc.f64(x)
    type_ = orc.i32(type_)
    q = orc.f64(q) if q is not None else np.zeros(nall)
    f = np.zeros((nall, 4))
    ev = np.zeros(8)
    err = _err()
    cc = params.cut_coul if params.cut_coul is not None else np.zeros_like(params.cut_lj)
    rc = lib().ref_pair_eval(
        C.c_int(params.style), C.c_int(prec), C.c_int(eflag), C.c_int(vflag), C.c_int(eatom), C.c_int(newton),
        C.c_int(nlocal), C.c_int(nall), orc._d(x), orc._i(type_), orc._d(q), orc._i(orc.i32(numneigh)),
        orc._l(np.ascontiguousarray(offsets, np.int64)), orc._i(orc.i32(entries)), orc._d(params.A), orc._d(params.rho),
        orc._d(params.Cc), orc._d(params.cut_lj), orc._d(orc.f64(cc)), C.c_int(offset_flag), C.c_double(skin),
        C.byref(params.p), orc._d(f), orc._d(ev), C.c_int(nthreads), err, C.c_int(512))
    if rc:
        raise RuntimeError("ref_pair_eval: " + err.value.decode())
    return f, ev


def nve(which, x, v, f, type_, mass, dt, ftm2v, rmass=None, ingroup=None):
    """FixNVEIntel::initial_integrate (which=0) / final_integrate (which=1); returns (x, v)"""
    x = orc.f64(x).copy()
    v = orc.f64(v).copy()
    n = len(x)
    err = _err()
    m = orc.f64(mass)
    rc = lib().ref_nve(C.c_int(which), C.c_int(n), C.c_int(len(m) - 1), orc._d(x), orc._d(v), orc._d(orc.f64(f)),
                       orc._i(orc.i32(type_)), orc._d(m), None if rmass is None else orc._d(orc.f64(rmass)),
                       None if ingroup is None else orc._i(orc.i32(ingroup)), C.c_double(dt), C.c_double(ftm2v),
                       err, C.c_int(512))
    if rc:
        raise RuntimeError("ref_nve: " + err.value.decode())
    return x, v


class PppmState(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("order", C.c_int), ("diff_ad", C.c_int),
                ("nlower", C.c_int), ("nupper", C.c_int), ("lo_out", C.c_int * 3), ("hi_out", C.c_int * 3),
                ("shift", C.c_double), ("shiftone", C.c_double), ("g_ewald", C.c_double), ("qqrd2e", C.c_double),
                ("scale", C.c_double), ("volume", C.c_double), ("boxlo", C.c_double * 3), ("prd", C.c_double * 3),
                ("delinv", C.c_double * 3), ("delvolinv", C.c_double),
                ("greensfn", orc.dp), ("vg", orc.dp), ("fkx", orc.dp), ("fky", orc.dp), ("fkz", orc.dp),
                ("rho_coeff", orc.dp), ("drho_coeff", orc.dp), ("sf_coeff", C.c_double * 6)]


def pppm_compute(pp, x, q, prec=orc.DOUBLE, eflag=1, vflag=1, nthreads=1, want_grids=True):
    """PPPMIntel::compute of the reference on the base-class state of the oracle's PPPM object `pp`.
    Returns (f, energy, virial, density_fft, fields) like orc.PPPM.compute + density()/field()."""
    st = PppmState()
    orc.lib().orc_pppm_export(pp.h, C.byref(st))
    n = len(x)
    f = np.zeros((n, 3))
    e = C.c_double(0.0)
    v = np.zeros(6)
    nfft = pp.nfft
    dens = np.zeros(nfft) if want_grids else None
    fields = np.zeros((1 if st.diff_ad else 3, nfft)) if want_grids else None
    err = _err()
    rc = lib().ref_pppm_compute(C.byref(st), C.c_int(prec), C.c_int(n), orc._d(orc.f64(x)), orc._d(orc.f64(q)),
                                C.c_int(eflag), C.c_int(vflag), orc._d(f), C.byref(e), orc._d(v),
                                None if dens is None else orc._d(dens), None if fields is None else orc._d(fields),
                                C.c_int(nthreads), err, C.c_int(512))
    if rc:
        raise RuntimeError("ref_pppm_compute: " + err.value.decode())
    return f, e.value, v, dens, fields


def last_seconds():
    """wall time of the reference's compute() inside the last pair_eval / pppm_compute call (harness set-up excluded)"""
    f = lib().ref_last_seconds
    f.restype = C.c_double
    return float(f())
