"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/b200md.h declares (no compute calls without a GPU), and the product fails loudly — never falls back —
when no device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    """every function declared in include/*.h (the drop-in boundary b200md.h and the test-only b200md_testing.h)"""
    syms = set()
    inc = os.path.join(ROOT, "include")
    for h in sorted(os.listdir(inc)):
        if not h.endswith(".h"):
            continue
        src = open(os.path.join(inc, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms.update(re.findall(r"\b(b200md_[a-z0-9_]+)\s*\(", src))
    return sorted(syms)


def test_header_symbols_exported(pkg):
    pkg.build()
    lib = ctypes.CDLL(pkg.LIBPATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, "declared in include/b200md.h but not exported: %s" % missing


def test_no_oracle_in_product(pkg):
    """the product never links, imports or executes anything under oracle/"""
    import subprocess
    out = subprocess.run(["ldd", pkg.LIBPATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    pdir = os.path.dirname(pkg.LIBPATH)
    for dirpath, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cc")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "import orc" not in txt and "oracle/" not in txt, f


def test_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.B200MDError) as ei:
        pkg.Context(0, pkg.PREC_DOUBLE)
    assert ei.value.code in (-3, -2)
    assert "no CPU fallback" in str(ei.value) or "device" in str(ei.value)
