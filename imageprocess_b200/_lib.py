"""ctypes binding of libipb200.so (C ABI declared in include/ipb200.h).

The product loads exactly one library: ``imageprocess_b200/lib/libipb200.so`` built by
``__graft_entry__.build()`` with nvcc for sm_100a.  There is no CPU fallback: if the library
is missing, ``load()`` raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libipb200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_longlong
_f = ctypes.c_float
_d = ctypes.c_double

# name -> argtypes (all return int status unless listed in _RESTYPE)
_PROTOS = {
    "ipb_version": [],
    "ipb_is_emulated": [],
    "ipb_sizeof": [_i],
    "ipb_rasterize_rois": [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp],
    "ipb_hist_u16": [_vp, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp],
    "ipb_hist_quantiles": [_vp, _vp, _vp, _i, _vp, _vp],
    "ipb_scatter_qvalues": [_vp, _vp, _i, _vp, _vp],
    "ipb_fret_eps": [_vp, _i, _i, _i, _f, _vp, _vp],
    "ipb_fa_params": [_vp, _vp, _vp, _i, _i64, _f, _vp, _vp],
    "ipb_fret_pixels": [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "ipb_fa_segment": [_vp, _i, _i, _i64, _vp, _i, _i, _vp, _vp, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                       _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp],
    "ipb_region_stats": [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "ipb_roi_stats_fused": [_vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp],
}
_PROTOS.update({
    "ipb_region_dilate": [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "ipb_region_moments": [_vp, _i, _vp, _vp, _vp],
    "ipb_preview_u16": [_vp, _i64, _i, _vp, _vp, _vp],
    "ipb_crop_normalize": [_vp, _i, _i64, _vp, _i, _i, _vp, _f, _vp, _vp, _vp, _vp, _vp],
    "ipb_eps_from_stat": [_vp, _vp, _i, _f, _vp, _vp],
    "ipb_hist_planes": [_vp, _i, _i, _vp, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp],
    "ipb_hist_select": [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ipb_selftest_fdiv": [_vp, _vp, _i64, _vp, _vp],
    "ipb_gaussian_f32": [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    "ipb_gauss_combine": [_vp, _vp, _vp, _i64, _i, _f, _vp],
    "ipb_graymorph_u16": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "ipb_sub_u16": [_vp, _vp, _vp, _i64, _vp],
    "ipb_convert_planes": [_vp, _vp, _i64, _i, _vp],
    "ipb_fa_contour_cells": [_vp, _i, _i64, _vp, _vp, _vp, _vp],
    "ipb_hist_sizes": [_i, _i, _i, _vp],
    "ipb_hist_select_sizes": [_i, _i, _vp],
    "ipb_roi_stats_fused_sizes": [_i, _i, _i, _i, _i, _vp],
    "ipb_fa_segment_sizes": [_i, _vp, _i, _vp],
    "ipb_region_dilate_sizes": [_i, _vp, _vp],
})
_RESTYPE = {"ipb_last_error": ctypes.c_char_p}


class IpbError(RuntimeError):
    pass


class Lib:
    """Thin checked wrapper: ``lib.call('ipb_xxx', *args)`` raises IpbError on status < 0."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise IpbError(
                f"{path} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        self.path = path
        self.c = ctypes.CDLL(path)
        self.c.ipb_last_error.restype = ctypes.c_char_p
        self.c.ipb_last_error.argtypes = []
        for name, argtypes in _PROTOS.items():
            fn = getattr(self.c, name)          # AttributeError if the symbol is missing
            fn.argtypes = argtypes
            fn.restype = _i

    def call(self, name, *args):
        rc = getattr(self.c, name)(*args)
        if rc < 0:
            msg = self.c.ipb_last_error()
            raise IpbError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
        return rc

    def sizes(self, name, n_out, *args):
        """Calls a workspace-size query (ipb_*_sizes): returns its int64 results as a list."""
        out = (ctypes.c_longlong * n_out)()
        self.call(name, *args, ctypes.cast(out, ctypes.c_void_p))
        return [int(v) for v in out]

    def exported(self):
        return list(_PROTOS) + list(_RESTYPE)


_lib = None


def load() -> Lib:
    global _lib
    if _lib is None:
        from . import build
        if os.path.isdir(build.CSRC):                      # source tree present: the library must match it
            have = open(build.HASH).read().strip() if os.path.exists(build.HASH) else None
            if os.path.exists(LIB_PATH) and have != build.source_digest():
                raise IpbError(f"{LIB_PATH} was built from different kernel sources: rebuild it "
                               "(python -c 'import __graft_entry__ as g; g.build()')")
        _lib = Lib(LIB_PATH)
        if _lib.c.ipb_is_emulated():
            raise IpbError("refusing to run the product on an emulated build")
    return _lib
