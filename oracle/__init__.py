"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's per-pixel analysis path (gavyek/ImageProcess,
SURVEY.md section 8).  It exists to *check* the CUDA path; nothing under
``imageprocess_b200/`` may import it.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs.

Layout
  c/polygon_rules.c   the two third-party polygon rules (matplotlib, skimage), C
  shims.py            restated third-party call surface (matplotlib.path.Path,
                      skimage.draw/morphology/measure) on numpy/scipy
  port.py             numpy restatement of the reference's own numeric functions,
                      each citing the reference file:line it follows
  refimport.py        imports the UNMODIFIED reference modules from /root/reference
                      with stub GUI modules (only possible in the build container)
  gen_golden.py       writes tests/golden/* from the reference itself

Parity pinning (see DESIGN.md "Oracle"):
  * matplotlib rule + bg_correct + quantify_stats: pinned bit-exactly by the 29 ROI
    rows of the reference's shipped fluor_intensity_perROI.csv files.
  * skimage polygon rule: pinned bit-exactly by the two shipped roi/mask/S01_mask.tif.
  * port.py functions: pinned against the unmodified reference functions executed in
    the build container (tests/test_oracle_vs_reference.py, tests/golden/*.npz).
  * FA morphology chain (remove_small_objects / closing / label / regionprops), FRET,
    Nesprin2, MOR, cropper: NO reference artefact with inputs exists ("parity
    unpinned by a shipped golden"); they are pinned only through the reference's own
    control flow run on the shims here.
  * shims.find_contours / approximate_polygon (skimage.measure) and port.segment_inside_polygon
    (ROI drawer assist): restated from the published scikit-image sources; the port equals the
    UNMODIFIED roi_manual_drawer.segment_inside_polygon run on these shims, but no reference
    artefact pins the shims themselves: PARITY UNPINNED.
  * shims.threshold_otsu and the optional Gaussian / top-hat stages: not in the reference at
    all; their oracle is scipy.ndimage / the restated skimage rule: PARITY
    UNPINNED BY REFERENCE.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "c", "polygon_rules.c")
_SO = os.path.join(_HERE, "_build", "libipb_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc -O2 -ffp-contract=off)."""
    if force or (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def clib():
    global _lib
    if _lib is None:
        so = _SO if os.path.exists(_SO) and not os.path.exists(_SRC) else build()
        lib = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        u8 = ctypes.POINTER(ctypes.c_uint8)
        lib.ipbo_mpl_points_in_path.argtypes = [dp, ctypes.c_int, dp, ctypes.c_size_t, u8]
        lib.ipbo_mpl_points_in_path.restype = ctypes.c_int
        lib.ipbo_sk_polygon_mask.argtypes = [dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8]
        lib.ipbo_sk_polygon_mask.restype = ctypes.c_long
        _lib = lib
    return _lib
