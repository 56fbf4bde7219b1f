"""oracle/refimport.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imports the UNMODIFIED reference modules from /root/reference/src with stub modules
for the GUI / plotting / I/O packages that are absent in the build container
(tkinter, matplotlib, tifffile, skimage, pptx).  ``matplotlib.path`` and ``skimage``
resolve to oracle/shims.py so the reference's own control flow runs on the restated
third-party arithmetic.  Only usable where /root/reference exists (the build
container); the GPU box never runs this.
"""
import importlib.util
import os
import sys
import types

import numpy as np

from . import shims

REF_ROOT = os.environ.get("IPB_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isdir(REF_SRC)


class _Anything(types.ModuleType):
    """Module whose every attribute is a permissive dummy (class / callable)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        dummy = type(name, (), {"__init__": lambda self, *a, **k: None,
                                "__call__": lambda self, *a, **k: None,
                                "__getattr__": lambda self, n: (lambda *a, **k: None)})
        setattr(self, name, dummy)
        return dummy


def _tif_imread(path, key=None, **_):
    from PIL import Image
    with Image.open(path) as im:
        if key:
            im.seek(int(key))
        return np.array(im)


def _tif_imwrite(path, data, **_):
    from PIL import Image
    Image.fromarray(np.asarray(data)).save(path, format="TIFF")


def install_stubs():
    if "tkinter" not in sys.modules or not isinstance(sys.modules["tkinter"], _Anything):
        try:
            import tkinter  # noqa: F401  (real one present: keep it)
        except Exception:
            for name in ("tkinter", "tkinter.filedialog", "tkinter.messagebox", "tkinter.ttk",
                         "tkinter.scrolledtext"):
                sys.modules[name] = _Anything(name)
    mpl_names = ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.figure",
                 "matplotlib.patches", "matplotlib.backends", "matplotlib.backends.backend_tkagg",
                 "matplotlib.widgets", "matplotlib.patheffects",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.inset_locator",
                 "pptx", "pptx.util")
    for name in mpl_names:
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    mpath = types.ModuleType("matplotlib.path")
    mpath.Path = shims.Path
    sys.modules["matplotlib.path"] = mpath
    sys.modules["matplotlib"].path = mpath
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].rcParams = {}                       # roi_manual_drawer.py:22 assigns key maps at import
    tf = types.ModuleType("tifffile")
    tf.imread, tf.imwrite = _tif_imread, _tif_imwrite
    sys.modules.setdefault("tifffile", tf)
    sk = types.ModuleType("skimage")
    sk_draw = types.ModuleType("skimage.draw")
    sk_draw.polygon = shims.polygon
    sk_morph = types.ModuleType("skimage.morphology")
    for n in ("remove_small_objects", "disk", "binary_closing", "binary_dilation", "binary_erosion"):
        setattr(sk_morph, n, getattr(shims, n))
    sk_meas = types.ModuleType("skimage.measure")
    for n in ("label", "regionprops", "find_contours", "approximate_polygon"):
        setattr(sk_meas, n, getattr(shims, n))
    sk.draw, sk.morphology, sk.measure = sk_draw, sk_morph, sk_meas
    for extra in ("transform", "exposure", "filters"):            # display-only imports of the ROI drawer
        mod = _Anything(f"skimage.{extra}")
        setattr(sk, extra, mod)
        sys.modules.setdefault(f"skimage.{extra}", mod)
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.draw", sk_draw)
    sys.modules.setdefault("skimage.morphology", sk_morph)
    sys.modules.setdefault("skimage.measure", sk_meas)


_cache = {}

_FILES = {
    "Fluor_INT": "INT/Fluor_INT.py",
    "FA_Analyzer": "INT/FA_Analyzer.py",
    "fret_ratio_builder": "FRET/fret_ratio_builder.py",
    "Nesprin2_FRET_Builder": "FRET/Nesprin2_FRET_Builder.py",
    "MOR_by_ROI": "MOR_by_ROI.py",
    "roi_channel_cropper": "roi_channel_cropper.py",
    "roi_manual_drawer": "roi_manual_drawer.py",
}


def load(name: str):
    """Return the unmodified reference module ``name`` (one of _FILES)."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError(f"reference not present at {REF_SRC}")
    install_stubs()
    path = os.path.join(REF_SRC, _FILES[name])
    spec = importlib.util.spec_from_file_location(f"_ipb_ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = [path]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    _cache[name] = mod
    return mod
