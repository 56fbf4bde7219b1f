"""Host-side ROI geometry: the three crop conventions and the rasteriser's rect tables
(SURVEY.md 8(a) a3).  Pure integer / float64 bookkeeping on a few hundred vertices; the
per-pixel work happens on the device."""
import math

import numpy as np

RULE_MPL = 0   # matplotlib.path.Path.contains_points (rasterize_polygon)
RULE_SK = 1    # skimage.draw.polygon


class RoiSpec:
    """One polygon to rasterise in a local grid.

    verts  (V,2) float64 local (x, y) -- already shifted exactly as the reference shifts
    grid   (w, h) size of the local grid (the frame, or a crop)
    org    (ox, oy) frame position of local (0, 0)
    frame  frame index in the batch
    erect  (x0, y0, x1, y1) half-open rect where the rule is evaluated
    srect  (x0, y0, x1, y1) half-open rect the stored mask covers (erect inside srect)
    """
    __slots__ = ("verts", "grid", "org", "frame", "erect", "srect")

    def __init__(self, verts, grid, org, frame, erect, srect):
        self.verts, self.grid, self.org, self.frame = verts, grid, org, frame
        self.erect, self.srect = erect, srect


def _clip_rect(x0, y0, x1, y1, w, h):
    x0, y0 = max(0, min(x0, w)), max(0, min(y0, h))
    x1, y1 = max(x0, min(x1, w)), max(y0, min(y1, h))
    return (x0, y0, x1, y1)


def mpl_spec(poly, grid_wh, org=(0, 0), frame=0, store_full=False, pad=0):
    """Spec for rasterize_polygon(poly, (h, w)) -- reference Fluor_INT.py:398-403.  The rule
    is evaluated on the vertex bbox padded by one pixel; outside it no pixel can be inside.
    pad > 0 stores the mask in a rect grown by `pad` pixels (room for later dilations)."""
    P = np.ascontiguousarray(np.asarray(poly, dtype=np.float64))
    w, h = int(grid_wh[0]), int(grid_wh[1])
    if P.ndim != 2 or P.shape[0] < 3 or not np.isfinite(P).all():
        e = (0, 0, 0, 0)
        return RoiSpec(P.reshape(-1, 2), (w, h), org, frame, e, (0, 0, w, h) if store_full else e)
    xmin, xmax = P[:, 0].min(), P[:, 0].max()
    ymin, ymax = P[:, 1].min(), P[:, 1].max()
    e = _clip_rect(math.floor(xmin) - 1, math.floor(ymin), math.ceil(xmax) + 2,
                   math.ceil(ymax) + 1, w, h)
    sr = e
    if store_full:
        sr = (0, 0, w, h)
    elif pad > 0 and e[2] > e[0] and e[3] > e[1]:
        sr = _clip_rect(e[0] - pad, e[1] - pad, e[2] + pad, e[3] + pad, w, h)
    return RoiSpec(P, (w, h), org, frame, e, sr)


def sk_spec(r, c, shape_hw, org=(0, 0), frame=0):
    """Spec for skimage.draw.polygon(r, c, shape) -- loop ranges of skimage _draw.pyx _polygon:
    minr = int(max(0, r.min())), maxr = min(shape[0]-1, int(ceil(r.max()))) (same for c).
    The stored mask covers the whole (crop) shape."""
    r = np.asarray(r, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    h, w = int(shape_hw[0]), int(shape_hw[1])
    P = np.ascontiguousarray(np.stack([c, r], axis=1))
    if P.shape[0] < 1 or h <= 0 or w <= 0 or not np.isfinite(P).all():
        return RoiSpec(P, (max(w, 0), max(h, 0)), org, frame, (0, 0, 0, 0), (0, 0, max(w, 0), max(h, 0)))
    minr = int(max(0, r.min()))
    maxr = min(h - 1, int(math.ceil(r.max())))
    minc = int(max(0, c.min()))
    maxc = min(w - 1, int(math.ceil(c.max())))
    e = _clip_rect(minc, minr, maxc + 1, maxr + 1, w, h)
    return RoiSpec(P, (w, h), org, frame, e, (0, 0, w, h))


def fa_crop_rect(poly, img_shape, pad=5):
    """FA_Analyzer.py:998-1003: floor/ceil bbox + pad, clipped, half-open slices.
    Returns (x_min, x_max, y_min, y_max)."""
    xs, ys = poly[:, 0], poly[:, 1]
    x_min, x_max = int(np.floor(xs.min())), int(np.ceil(xs.max()))
    y_min, y_max = int(np.floor(ys.min())), int(np.ceil(ys.max()))
    x_min = max(0, x_min - pad)
    x_max = min(img_shape[1], x_max + pad)
    y_min = max(0, y_min - pad)
    y_max = min(img_shape[0], y_max + pad)
    return x_min, x_max, y_min, y_max


def fa_spec(poly, img_shape, frame=0):
    """Crop + skimage mask spec of FA_Analyzer.py:997-1015.  Returns (spec, rect) or
    (None, rect) for an empty crop."""
    poly = np.array(poly, dtype=np.float64)
    x_min, x_max, y_min, y_max = fa_crop_rect(poly, img_shape)
    if x_min >= x_max or y_min >= y_max:
        return None, (x_min, x_max, y_min, y_max)
    pc = poly.copy()
    pc[:, 0] -= x_min
    pc[:, 1] -= y_min
    spec = sk_spec(pc[:, 1], pc[:, 0], (y_max - y_min, x_max - x_min), org=(x_min, y_min), frame=frame)
    return spec, (x_min, x_max, y_min, y_max)


def trunc_crop_rect(P, W, H, pad_ratio=0.05):
    """int() truncation + pad max(10, ratio*max(W,H)), INCLUSIVE ends
    (roi_channel_cropper.py:885-893; Fluor_INT.py:1028-1041; fret_ratio_builder.py:517-522).
    Returns (x0, x1, y0, y1) inclusive."""
    P = np.asarray(P)
    minx, maxx = P[:, 0].min(), P[:, 0].max()
    miny, maxy = P[:, 1].min(), P[:, 1].max()
    pad = max(10, int(pad_ratio * max(W, H)))
    return (max(int(minx) - pad, 0), min(int(maxx) + pad, W - 1),
            max(int(miny) - pad, 0), min(int(maxy) + pad, H - 1))


class RoiTable:
    """Flattened (CSR) host tables for a list of RoiSpec, ready to upload."""

    def __init__(self, specs):
        n = len(specs)
        self.n = n
        self.specs = specs
        self.vert_off = np.zeros(n + 1, dtype=np.int32)
        for i, s in enumerate(specs):
            self.vert_off[i + 1] = self.vert_off[i] + s.verts.shape[0]
        self.verts = (np.concatenate([s.verts for s in specs], axis=0) if n and self.vert_off[-1] > 0
                      else np.zeros((0, 2), dtype=np.float64)).astype(np.float64)
        self.erect = np.array([s.erect for s in specs], dtype=np.int32).reshape(n, 4)
        self.srect = np.array([s.srect for s in specs], dtype=np.int32).reshape(n, 4)
        self.org = np.array([s.org for s in specs], dtype=np.int32).reshape(n, 2)
        self.frame = np.array([s.frame for s in specs], dtype=np.int32)
        sw = self.srect[:, 2] - self.srect[:, 0]
        sh = self.srect[:, 3] - self.srect[:, 1]
        self.wpr = ((sw + 31) // 32).astype(np.int32)
        self.rows = sh.astype(np.int32)
        self.mask_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(self.wpr.astype(np.int64) * sh.astype(np.int64), out=self.mask_off[1:])
        self.max_rows = int(sh.max()) if n else 0
        self.max_wpr = int(self.wpr.max()) if n else 0
        self.total_words = int(self.mask_off[-1])


# ---------------------------------------------------------------------- vectorised batch tables
class BatchGeometry:
    """All ROI tables of a batch of frames, built with vectorised numpy (no per-polygon
    Python loop).  Results are identical to mpl_spec / fa_spec applied per polygon."""
    pass


def flatten_polys(polys_per_frame):
    """-> verts (V,2) f64, off (N+1,) int64, frame (N,) int32, roi (N,) int32 (1-based)."""
    polys, frame, roi = [], [], []
    for f, pl in enumerate(polys_per_frame):
        for i, P in enumerate(pl or ()):
            polys.append(np.asarray(P, dtype=np.float64).reshape(-1, 2))
            frame.append(f)
            roi.append(i + 1)
    n = len(polys)
    cnt = np.fromiter((p.shape[0] for p in polys), dtype=np.int64, count=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=off[1:])
    verts = np.concatenate(polys, axis=0) if n else np.zeros((0, 2), dtype=np.float64)
    return verts, off, np.asarray(frame, dtype=np.int32), np.asarray(roi, dtype=np.int32)


def _seg_minmax(a, off):
    if off.shape[0] <= 1:
        z = np.zeros(0, dtype=np.float64)
        return z, z
    idx = off[:-1]
    return np.minimum.reduceat(a, idx), np.maximum.reduceat(a, idx)


def _mask_layout(srect):
    sw = (srect[:, 2] - srect[:, 0]).astype(np.int64)
    sh = (srect[:, 3] - srect[:, 1]).astype(np.int64)
    wpr = (sw + 31) // 32
    mask_off = np.zeros(srect.shape[0] + 1, dtype=np.int64)
    np.cumsum(wpr * sh, out=mask_off[1:])
    return wpr.astype(np.int32), sh.astype(np.int32), mask_off


def mpl_tables(verts, off, W, H):
    """Vectorised mpl_spec: erect == srect == vertex bbox padded by one pixel in x, clipped."""
    n = off.shape[0] - 1
    cnt = np.diff(off)
    xmin, xmax = _seg_minmax(verts[:, 0], off)
    ymin, ymax = _seg_minmax(verts[:, 1], off)
    ok = (cnt >= 3) & np.isfinite(xmin) & np.isfinite(xmax) & np.isfinite(ymin) & np.isfinite(ymax)
    big = 1 << 30
    def cl(v, hi):
        return np.clip(np.where(ok, v, 0.0), -big, big).astype(np.int64).clip(0, hi)
    x0 = cl(np.floor(xmin) - 1, W)
    y0 = cl(np.floor(ymin), H)
    x1 = np.maximum(cl(np.ceil(xmax) + 2, W), x0)
    y1 = np.maximum(cl(np.ceil(ymax) + 1, H), y0)
    rect = np.stack([x0, y0, x1, y1], axis=1).astype(np.int32)
    rect[~ok] = 0
    return rect


def fa_tables(verts, off, W, H, pad=5):
    """Vectorised fa_spec: crop rects (FA_Analyzer.py:998-1003), crop-local vertices and the
    skimage loop ranges.  Returns dict(local_verts, erect, srect, org, crop_rect)."""
    n = off.shape[0] - 1
    cnt = np.diff(off)
    xmin, xmax = _seg_minmax(verts[:, 0], off)
    ymin, ymax = _seg_minmax(verts[:, 1], off)
    fin = np.isfinite(xmin) & np.isfinite(xmax) & np.isfinite(ymin) & np.isfinite(ymax) & (cnt >= 1)
    sx = lambda v: np.where(fin, v, 0.0)
    x_min = np.maximum(0, np.floor(sx(xmin)).astype(np.int64) - pad)
    x_max = np.minimum(W, np.ceil(sx(xmax)).astype(np.int64) + pad)
    y_min = np.maximum(0, np.floor(sx(ymin)).astype(np.int64) - pad)
    y_max = np.minimum(H, np.ceil(sx(ymax)).astype(np.int64) + pad)
    ok = fin & (x_min < x_max) & (y_min < y_max)
    w = np.where(ok, x_max - x_min, 0)
    h = np.where(ok, y_max - y_min, 0)
    local = verts.copy()
    local[:, 0] -= np.repeat(x_min, cnt)
    local[:, 1] -= np.repeat(y_min, cnt)
    cmin, cmax = _seg_minmax(local[:, 0], off)
    rmin, rmax = _seg_minmax(local[:, 1], off)
    z = lambda v: np.where(ok, v, 0.0)
    minr = np.maximum(0.0, z(rmin)).astype(np.int64)
    maxr = np.minimum(h - 1, np.ceil(z(rmax)).astype(np.int64))
    minc = np.maximum(0.0, z(cmin)).astype(np.int64)
    maxc = np.minimum(w - 1, np.ceil(z(cmax)).astype(np.int64))
    ex0 = np.clip(minc, 0, w)
    ey0 = np.clip(minr, 0, h)
    ex1 = np.maximum(np.clip(maxc + 1, 0, w), ex0)
    ey1 = np.maximum(np.clip(maxr + 1, 0, h), ey0)
    erect = np.stack([ex0, ey0, ex1, ey1], axis=1).astype(np.int32)
    erect[~ok] = 0
    srect = np.stack([np.zeros(n, np.int64), np.zeros(n, np.int64), w, h], axis=1).astype(np.int32)
    org = np.stack([np.where(ok, x_min, 0), np.where(ok, y_min, 0)], axis=1).astype(np.int32)
    crop_rect = np.stack([x_min, x_max, y_min, y_max], axis=1)
    return {"local_verts": local, "erect": erect, "srect": srect, "org": org,
            "crop_rect": crop_rect, "ok": ok}
