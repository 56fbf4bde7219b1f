"""oracle/shims.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restated call surface of the two third-party packages the reference's hot path uses
but which are not installed here and whose source is not under /root/reference:

  matplotlib (requirements.txt:3, unpinned)  -> ``Path(...).contains_points``
  scikit-image (requirements.txt:8, unpinned) -> ``draw.polygon``,
      ``morphology.{remove_small_objects, disk, binary_closing}``,
      ``measure.{label, regionprops, find_contours}``

Every function names the published algorithm it restates and the reference call
site that consumes it.  numpy / scipy.ndimage are called directly (they are what
skimage itself wraps); the two polygon rules are in c/polygon_rules.c.
"""
import ctypes

import numpy as np
from scipy import ndimage as ndi

from . import clib

_DP = ctypes.POINTER(ctypes.c_double)
_U8 = ctypes.POINTER(ctypes.c_uint8)


# --------------------------------------------------------------------------- matplotlib
class Path:
    """matplotlib.path.Path restricted to what rasterize_polygon uses
    (src/INT/Fluor_INT.py:398-403): vertices only, no codes."""

    def __init__(self, vertices, codes=None):
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=np.float64))
        if self.vertices.ndim != 2 or self.vertices.shape[1] != 2:
            raise ValueError("'vertices' must be 2D with shape (N, 2)")
        self.codes = codes

    def contains_points(self, points, transform=None, radius=0.0):
        """matplotlib _path.h points_in_path / point_in_path_impl, radius 0."""
        if radius != 0.0 or transform is not None:
            raise NotImplementedError("oracle shim: radius 0, identity transform only")
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
        n = pts.shape[0]
        out = np.zeros(n, dtype=np.uint8)
        rc = clib().ipbo_mpl_points_in_path(
            self.vertices.ctypes.data_as(_DP), int(self.vertices.shape[0]),
            pts.ctypes.data_as(_DP), n, out.ctypes.data_as(_U8))
        if rc != 0:
            raise MemoryError("ipbo_mpl_points_in_path")
        return out.astype(bool)


# --------------------------------------------------------------------------- skimage.draw
def polygon(r, c, shape=None):
    """skimage.draw.polygon(r, c, shape) -> (rr, cc)   (FA_Analyzer.py:1014)."""
    r = np.ascontiguousarray(np.asarray(r, dtype=np.float64))
    c = np.ascontiguousarray(np.asarray(c, dtype=np.float64))
    if shape is None:
        shape = (int(np.ceil(r.max())) + 1, int(np.ceil(c.max())) + 1)
    H, W = int(shape[0]), int(shape[1])
    if H <= 0 or W <= 0 or r.size == 0:
        return np.zeros(0, dtype=np.intp), np.zeros(0, dtype=np.intp)
    out = np.zeros((H, W), dtype=np.uint8)
    clib().ipbo_sk_polygon_mask(r.ctypes.data_as(_DP), c.ctypes.data_as(_DP), int(r.size),
                                H, W, out.ctypes.data_as(_U8))
    rr, cc = np.nonzero(out)            # raster order == _polygon's append order
    return rr.astype(np.intp), cc.astype(np.intp)


# --------------------------------------------------------------------------- skimage.morphology
def disk(radius, dtype=np.uint8):
    """skimage.morphology.disk: (x^2 + y^2) <= r^2 on a (2r+1)^2 grid."""
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)


def remove_small_objects(ar, min_size=64, connectivity=1):
    """skimage.morphology.remove_small_objects for a bool image: label with
    ndi.generate_binary_structure(ndim, connectivity), bincount, drop components with
    size < min_size (float compare; FA_Analyzer.py:151 passes a float)."""
    out = ar.copy()
    if min_size == 0:
        return out
    if out.dtype == bool:
        footprint = ndi.generate_binary_structure(ar.ndim, connectivity)
        ccs = np.zeros_like(ar, dtype=np.int32)
        ndi.label(ar, footprint, output=ccs)
    else:
        ccs = out
    component_sizes = np.bincount(ccs.ravel())
    too_small = component_sizes < min_size
    out[too_small[ccs]] = 0
    return out


def binary_dilation(image, footprint=None):
    return ndi.binary_dilation(image, structure=footprint)


def binary_erosion(image, footprint=None):
    return ndi.binary_erosion(image, structure=footprint, border_value=True)


def binary_closing(image, footprint=None):
    """skimage.morphology.binary_closing: ndi dilation (outside = 0) then ndi erosion
    with border_value=True (outside = 1)  (FA_Analyzer.py:155-156)."""
    return binary_erosion(binary_dilation(image, footprint), footprint)


# --------------------------------------------------------------------------- skimage.measure
def label(label_image, background=None, return_num=False, connectivity=None):
    """skimage.measure.label for a 2-D bool image: background 0, full (8-)connectivity
    by default, labels 1..N in raster order of each component's first pixel."""
    a = np.asarray(label_image)
    if connectivity is None:
        connectivity = a.ndim
    st = ndi.generate_binary_structure(a.ndim, connectivity)
    lab, n = ndi.label(a != 0, structure=st)
    lab = lab.astype(np.int64)
    return (lab, n) if return_num else lab


class RegionProperties:
    def __init__(self, lab, label_image, intensity_image):
        self.label = int(lab)
        self._sl = ndi.find_objects((label_image == lab).astype(np.int8))[0]
        self._mask = label_image[self._sl] == lab
        self._int = None if intensity_image is None else intensity_image[self._sl]

    @property
    def area(self):
        return np.float64(np.sum(self._mask))          # skimage >= 0.20: float area

    @property
    def mean_intensity(self):
        return np.mean(self._int[self._mask], axis=0)

    intensity_mean = mean_intensity

    @property
    def coords(self):
        idx = np.argwhere(self._mask)
        return idx + np.array([s.start for s in self._sl])

    @property
    def centroid(self):
        return tuple(self.coords.mean(axis=0))

    @property
    def bbox(self):
        return tuple(s.start for s in self._sl) + tuple(s.stop for s in self._sl)


def regionprops(label_image, intensity_image=None, **_):
    """skimage.measure.regionprops: one entry per label present, ascending."""
    label_image = np.asarray(label_image)
    labs = np.unique(label_image)
    return [RegionProperties(l, label_image, intensity_image) for l in labs if l != 0]


def find_contours(image, level=0.5, **_):
    """Placeholder for skimage.measure.find_contours: FA_Analyzer.py:168 uses the
    result only for drawing and as an emptiness check (never empty for a non-empty
    region).  Returns the outer boundary pixels as one (N, 2) float array."""
    m = np.asarray(image) > level
    if not m.any():
        return []
    er = ndi.binary_erosion(m)
    return [np.argwhere(m & ~er).astype(np.float64)]
