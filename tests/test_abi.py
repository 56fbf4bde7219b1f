"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/ipb200.h declares (no compute calls here)."""
import ctypes
import os
import re

from imageprocess_b200 import _lib, build, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ipb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ipb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    so = build.build()
    c = ctypes.CDLL(so)
    names = header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(c, n), f"{n} declared in ipb200.h but not exported"
    lib = _lib.Lib(so)
    assert lib.c.ipb_is_emulated() == 0
    ops.check_struct_sizes(lib)
    for n in lib.exported():
        assert n in names, f"{n} bound in _lib.py but not declared in ipb200.h"


def test_sass_is_sm100a():
    so = build.build()
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "imageprocess_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "tests.emu" not in txt, f
