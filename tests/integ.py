"""ctypes wrapper of oracle/_ref/libinteg.so — the drop-in translation units of lammps-buck-intel_b200/integration/
(the B200 bodies of the classes that the reference's own headers declare) behind tests/integration_harness.cpp.
TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile.ref (target `integ`) where /root/reference is present; on the GPU box
the prebuilt library travels with the snapshot."""
import ctypes as C
import os
import subprocess

import numpy as np

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ODIR, "_ref", "libinteg.so")
REFDIR = os.environ.get("B200MD_REFERENCE", "/root/reference")
_lib = None


def build():
    """returns the library path, or None when neither the reference headers nor a prebuilt library exist"""
    if not os.path.isdir(REFDIR):
        return LIB if os.path.exists(LIB) else None
    r = subprocess.run(["make", "-C", ODIR, "-f", "Makefile.ref", "REF=" + REFDIR, "integ"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref/libinteg.so build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref/libinteg.so is missing and /root/reference is not present")
        _lib = C.CDLL(path)
    return _lib


def buck_coul_long(params, prec, system, A, rho, Cc, cut_lj, cut_coul, skin, grid=None, order=5, diff_ad=0, eflag=1, vflag=1,
                   nsteps=1, dx=None):
    """PairBuckCoulLongIntel (+ PPPMIntel when grid is given) of integration/ on `system`; returns (f, ev, ek, vk)"""
    s = system
    x = orc.f64(s["x"])
    n = len(x)
    f = np.zeros((n, 3))
    ev = np.zeros(8)
    ek = C.c_double(0.0)
    vk = np.zeros(6)
    err = C.create_string_buffer(512)
    nx, ny, nz = grid if grid is not None else (0, 0, 0)
    dxa = None if dx is None else orc.f64(dx)
    rc = lib().integ_buck_coul_long(
        C.c_int(prec), C.c_int(n), orc._d(x), orc._i(orc.i32(s["type"])), orc._d(orc.f64(s["q"])), C.c_int(int(s["ntypes"])),
        orc._d(orc.f64(s["mass"])), orc._d(orc.f64(s["boxlo"])), orc._d(orc.f64(s["boxhi"])), C.c_double(skin),
        orc._d(orc.f64(A)), orc._d(orc.f64(rho)), orc._d(orc.f64(Cc)), orc._d(orc.f64(cut_lj)), C.c_double(cut_coul),
        C.byref(params.p), C.c_int(nx), C.c_int(ny), C.c_int(nz), C.c_int(order), C.c_int(diff_ad), C.c_int(eflag),
        C.c_int(vflag), C.c_int(nsteps), None if dxa is None else orc._d(dxa), orc._d(f), orc._d(ev), C.byref(ek),
        orc._d(vk), err, C.c_int(512))
    if rc:
        raise RuntimeError(err.value.decode())
    return f, ev, ek.value, vk
