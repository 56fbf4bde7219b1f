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


def test_plain_c_caller(tmp_path):
    """include/ipb200.h as a non-Python binding sees it (tests/c/abi_consumer.c): compiles as strict C99,
    every declared entry point links against the product library, and the host-side queries (struct
    sizes, workspace sizes, the error convention) answer without a GPU with the values the Python
    binding works with."""
    import subprocess
    import numpy as np
    so = build.build()
    src = os.path.join(ROOT, "tests", "c", "abi_consumer.c")
    text = open(src).read()
    for n in header_functions():
        assert f"(fn_t){n}" in text, f"{n} is declared in ipb200.h but tests/c/abi_consumer.c does not reference it"
    exe = str(tmp_path / "abi_consumer")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    src, "-o", exe, so, f"-Wl,-rpath,{os.path.dirname(so)}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = {ln.split()[0]: ln.split()[1:] for ln in out.strip().splitlines()}
    assert int(got["entry_points"][0]) == len(header_functions())
    assert got["emulated"] == ["0"]
    assert [int(v) for v in got["sizeof"]] == [dt.itemsize for dt in ops._SIZEOF]
    lib = _lib.Lib(so)
    assert int(got["version"][0]) == lib.c.ipb_version()

    def sizes(name, n_out, *args):
        buf = (ctypes.c_int64 * 16)()
        fn = getattr(lib.c, name)
        fn.restype = ctypes.c_int
        rc = fn(*args, buf)
        return [str(rc)] + [str(int(v)) for v in buf[:n_out]]
    wh = np.array([141, 135, 1008, 469, 33, 7], dtype=np.int32)
    whp = wh.ctypes.data_as(ctypes.c_void_p)
    assert got["hist_sizes"] == sizes("ipb_hist_sizes", 3, 5, 2048, 1)
    assert got["hist_select_sizes"] == sizes("ipb_hist_select_sizes", 5, 7, 11)
    assert got["roi_stats_fused_sizes"] == sizes("ipb_roi_stats_fused_sizes", 6, 24, 24, 330, 325, 148)
    assert got["fa_segment_sizes"] == sizes("ipb_fa_segment_sizes", 9, 3, whp, 1)
    assert got["region_dilate_sizes"] == sizes("ipb_region_dilate_sizes", 1, 3, whp)
    # what the header states about the plain cases: 65 536 uint32 bins and 4 uint64 moments per histogram job
    assert got["hist_sizes"][:3] == ["0", str(5 * 65536 * 4), str(5 * 4 * 8)]
    assert got["bad_arg"] == [str(-1), "message"]
