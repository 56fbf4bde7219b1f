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


def _mc_fraction(a, b, level):
    return 0.0 if b == a else (level - a) / (b - a)


def _contour_segments(array, level, vertex_connect_high=False):
    """skimage/measure/_find_contours_cy.pyx _get_contour_segments (scikit-image is unpinned in
    requirements.txt:8; restated from the published source): marching squares over every 2 x 2
    cell in raster order, 0-2 oriented segments per cell, points interpolated linearly."""
    segs = []
    H, W = array.shape
    for r0 in range(H - 1):
        r1 = r0 + 1
        for c0 in range(W - 1):
            c1 = c0 + 1
            ul, ur, ll, lr = array[r0, c0], array[r0, c1], array[r1, c0], array[r1, c1]
            if np.isnan(ul) or np.isnan(ur) or np.isnan(ll) or np.isnan(lr):
                continue
            case = (1 if ul > level else 0) + (2 if ur > level else 0) + (4 if ll > level else 0) + (8 if lr > level else 0)
            if case in (0, 15):
                continue
            top = (float(r0), c0 + _mc_fraction(ul, ur, level))
            bottom = (float(r1), c0 + _mc_fraction(ll, lr, level))
            left = (r0 + _mc_fraction(ul, ll, level), float(c0))
            right = (r0 + _mc_fraction(ur, lr, level), float(c1))
            if case == 1:
                segs.append((top, left))
            elif case == 2:
                segs.append((right, top))
            elif case == 3:
                segs.append((right, left))
            elif case == 4:
                segs.append((left, bottom))
            elif case == 5:
                segs.append((top, bottom))
            elif case == 6:
                if vertex_connect_high:
                    segs.append((left, top)); segs.append((right, bottom))
                else:
                    segs.append((right, top)); segs.append((left, bottom))
            elif case == 7:
                segs.append((right, bottom))
            elif case == 8:
                segs.append((bottom, right))
            elif case == 9:
                if vertex_connect_high:
                    segs.append((top, right)); segs.append((bottom, left))
                else:
                    segs.append((top, left)); segs.append((bottom, right))
            elif case == 10:
                segs.append((bottom, top))
            elif case == 11:
                segs.append((bottom, left))
            elif case == 12:
                segs.append((left, right))
            elif case == 13:
                segs.append((top, right))
            elif case == 14:
                segs.append((left, top))
    return segs


def _assemble_contours(segments):
    """skimage/measure/_find_contours.py _assemble_contours: links oriented segments into
    polylines; contours are returned in the order in which their first segment appeared."""
    from collections import deque
    current_index = 0
    contours, starts, ends = {}, {}, {}
    for from_point, to_point in segments:
        if from_point == to_point:
            continue
        tail, tail_num = starts.pop(to_point, (None, None))
        head, head_num = ends.pop(from_point, (None, None))
        if tail is not None and head is not None:
            if tail is head:
                head.append(to_point)
            elif tail_num > head_num:
                head.extend(tail)
                contours.pop(tail_num, None)
                starts[head[0]] = (head, head_num)
                ends[head[-1]] = (head, head_num)
            else:
                tail.extendleft(reversed(head))
                starts.pop(head[0], None)
                contours.pop(head_num, None)
                starts[tail[0]] = (tail, tail_num)
                ends[tail[-1]] = (tail, tail_num)
        elif tail is None and head is None:
            new_contour = deque((from_point, to_point))
            contours[current_index] = new_contour
            starts[from_point] = (new_contour, current_index)
            ends[to_point] = (new_contour, current_index)
            current_index += 1
        elif head is None:
            tail.appendleft(from_point)
            starts[from_point] = (tail, tail_num)
        else:
            head.append(to_point)
            ends[to_point] = (head, head_num)
    return [np.array(c) for _, c in sorted(contours.items())]


def find_contours(image, level=0.5, fully_connected="low", positive_orientation="low", **_):
    """skimage.measure.find_contours (FA_Analyzer.py:168: find_contours(labeled_img == k, 0.5)):
    marching squares + assembly, float64 (row, col) points; no padding, so outlines of regions that
    touch the array border stay open.  Parity unpinned by a reference artefact (scikit-image absent)."""
    arr = np.asarray(image)
    if arr.dtype == bool or arr.dtype.kind in "iu":
        arr = arr.astype(np.float64)
    if arr.ndim != 2 or arr.shape[0] < 2 or arr.shape[1] < 2:
        raise ValueError("Input array must be at least 2x2.")
    segs = _contour_segments(arr.astype(np.float64), float(level), fully_connected == "high")
    contours = _assemble_contours(segs)
    if positive_orientation == "high":
        contours = [c[::-1] for c in contours]
    return contours


def threshold_otsu(image):
    """skimage.filters.threshold_otsu(image) for an integer image (scikit-image unpinned and absent;
    restated from skimage/filters/thresholding.py + exposure.histogram): integer images get one
    bin per value from image.min() to image.max() (np.bincount), then
        weight1 = cumsum(counts); weight2 = cumsum(counts[::-1])[::-1]
        mean1 = cumsum(counts * centers) / weight1; mean2 = (cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
        variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
        threshold = centers[argmax(variance12)]
    A single-valued image returns that value.  The reference does NOT call this (optional stage)."""
    a = np.asarray(image)
    if a.min() == a.max():
        return a.flat[0]
    off = int(a.min())
    counts = np.bincount((a.ravel().astype(np.int64) - off)).astype(np.float64)
    centers = np.arange(off, off + counts.shape[0], dtype=np.float64)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    mean1 = np.cumsum(counts * centers) / weight1
    mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return centers[int(np.argmax(variance12))]


def approximate_polygon(coords, tolerance):
    """skimage.measure.approximate_polygon (skimage/measure/_polygon.py, Douglas-Peucker with an
    explicit stack); roi_manual_drawer.py:407.  Restated from the published source."""
    if tolerance <= 0:
        return coords
    chain = np.zeros(coords.shape[0], 'bool')
    dists = np.zeros(coords.shape[0])
    chain[0] = True
    chain[-1] = True
    pos_stack = [(0, chain.shape[0] - 1)]
    end_of_chain = False
    while not end_of_chain:
        start, end = pos_stack.pop()
        r0, c0 = coords[start, :]
        r1, c1 = coords[end, :]
        dr = r1 - r0
        dc = c1 - c0
        segment_angle = -np.arctan2(dr, dc)
        segment_dist = c0 * np.sin(segment_angle) + r0 * np.cos(segment_angle)
        segment_coords = coords[start + 1:end, :]
        segment_dists = dists[start + 1:end]
        dr0 = segment_coords[:, 0] - r0
        dc0 = segment_coords[:, 1] - c0
        dr1 = segment_coords[:, 0] - r1
        dc1 = segment_coords[:, 1] - c1
        projected_lengths0 = dr0 * dr + dc0 * dc
        projected_lengths1 = -dr1 * dr - dc1 * dc
        perp = np.logical_and(projected_lengths0 > 0, projected_lengths1 > 0)
        eucl = np.logical_not(perp)
        segment_dists[perp] = np.abs(segment_coords[perp, 0] * np.cos(segment_angle)
                                     + segment_coords[perp, 1] * np.sin(segment_angle) - segment_dist)
        segment_dists[eucl] = np.minimum(np.sqrt(dc0[eucl] ** 2 + dr0[eucl] ** 2),
                                         np.sqrt(dc1[eucl] ** 2 + dr1[eucl] ** 2))
        if np.any(segment_dists > tolerance):
            new_end = start + np.argmax(segment_dists) + 1
            pos_stack.append((new_end, end))
            pos_stack.append((start, new_end))
            chain[new_end] = True
        if len(pos_stack) == 0:
            end_of_chain = True
    return coords[chain, :]
