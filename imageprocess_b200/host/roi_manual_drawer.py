"""Device-backed assists of the interactive ROI drawer.  The drawer itself
(matplotlib widgets, ROI JSON / mask / zip writers) stays host code and is out of scope; these are
the two numeric helpers it calls while the user works:

    segment_inside_polygon(img, poly, thr_param, min_area, tolerance, mode) -> (thr, None, polygon)
                                                                    roi_manual_drawer.py:337-418
    render_pipeline(img, ...)  (the body of ROIEditor._render_pipeline)  roi_manual_drawer.py:870-876

On the device: the polygon's mask in its bounding box (matplotlib rule), the statistics of the
pixels under it (exact percentile / mean / std), threshold & mask, 4-connected labelling with
component sizes, the labelling of the complement for the hole fill, the marching-squares cells of
the filled mask, and the Gaussian filters (TMA-tiled).  On the host: argmax over the handful of
component sizes, the contour linking, polygon areas and the Douglas-Peucker simplification
(O(contour points), restated from skimage.measure.approximate_polygon).
"""
import numpy as np

from .. import contours as ct
from .. import filters, geometry as geo, ops
from ..ops import COMP, CROP, QK_PCT, REGION, SRC_U16, STAT_JOB, q32_of
from . import common


def _engine():
    import imageprocess_b200 as ipb
    return ipb.engine()


def polygon_area(xy):
    x, y = xy[:, 0], xy[:, 1]
    return 0.5 * np.abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def approximate_polygon(coords, tolerance):
    """skimage.measure.approximate_polygon (Douglas-Peucker, iterative with an explicit stack)."""
    if tolerance <= 0:
        return coords
    chain = np.zeros(coords.shape[0], "bool")
    dists = np.zeros(coords.shape[0])
    chain[0] = True
    chain[-1] = True
    pos_stack = [(0, chain.shape[0] - 1)]
    end_of_chain = False
    while not end_of_chain:
        start, end = pos_stack.pop()
        r0, c0 = coords[start, :]
        r1, c1 = coords[end, :]
        dr, dc = r1 - r0, c1 - c0
        segment_angle = -np.arctan2(dr, dc)
        segment_dist = c0 * np.sin(segment_angle) + r0 * np.cos(segment_angle)
        segment_coords = coords[start + 1:end, :]
        segment_dists = dists[start + 1:end]
        dr0, dc0 = segment_coords[:, 0] - r0, segment_coords[:, 1] - c0
        dr1, dc1 = segment_coords[:, 0] - r1, segment_coords[:, 1] - c1
        projected_lengths0 = dr0 * dr + dc0 * dc
        projected_lengths1 = -dr1 * dr - dc1 * dc
        perp = np.logical_and(projected_lengths0 > 0, projected_lengths1 > 0)
        eucl = np.logical_not(perp)
        segment_dists[perp] = np.abs(segment_coords[perp, 0] * np.cos(segment_angle)
                                     + segment_coords[perp, 1] * np.sin(segment_angle) - segment_dist)
        segment_dists[eucl] = np.minimum(np.sqrt(dc0[eucl] ** 2 + dr0[eucl] ** 2), np.sqrt(dc1[eucl] ** 2 + dr1[eucl] ** 2))
        if np.any(segment_dists > tolerance):
            new_end = start + np.argmax(segment_dists) + 1
            pos_stack.append((new_end, end))
            pos_stack.append((start, new_end))
            chain[new_end] = True
        if len(pos_stack) == 0:
            end_of_chain = True
    return coords[chain, :]


def _label_crop(eng, plane_dev, h, w, mask_pool, thr, conn):
    """(labels int32 [h][w], component areas) of (plane > thr) & mask for one crop on the device."""
    mem = eng.mem
    wpr = (w + 31) // 32
    crops = np.zeros(1, dtype=CROP)
    crops["w"], crops["h"], crops["wpr"] = w, h, wpr
    crop_wh = np.array([w, h], dtype=np.int32)                   # named: a temporary would be freed before the call runs
    sizes = eng.lib.sizes("ipb_fa_segment_sizes", 9, 1, crop_wh.ctypes.data, 1)
    bufs = [mem.empty(sizes[0], np.uint8) for _ in range(4)]
    L, cs = mem.empty(sizes[1], np.uint8), mem.empty(sizes[1], np.uint8)
    rr, rb = mem.empty(sizes[2], np.uint8), mem.empty(sizes[2], np.uint8)
    cc, comp_off = mem.empty(sizes[3], np.uint8), mem.zeros(2, np.int32)
    comps = mem.empty(sizes[5], COMP)
    labels = mem.empty(h * w, np.int32)
    fa_params = mem.from_host(np.array([[0, 0, 0, thr]], dtype=np.float32))
    d_crops = mem.from_host(crops)
    eng.call("ipb_fa_segment", d_crops.ptr, 1, h, h, plane_dev.ptr, h, w, fa_params.ptr, mask_pool.ptr, 0.0, 0,
             bufs[0].ptr, bufs[1].ptr, L.ptr, cs.ptr, bufs[2].ptr, rr.ptr, rb.ptr, cc.ptr, bufs[3].ptr, comp_off.ptr,
             comps.ptr, sizes[5], labels.ptr, 2, None, int(conn), mem.stream)
    n = int(comp_off.host()[1])
    return labels, comps.host()[:n]["area"].astype(np.int64), d_crops


def _full_mask_pool(eng, h, w):
    wpr = (w + 31) // 32
    rowbits = np.zeros(wpr * 32, dtype=np.uint8)
    rowbits[:w] = 1
    row = np.packbits(rowbits, bitorder="little").view(np.uint32)
    return eng.mem.from_host(np.tile(row, h))


def segment_inside_polygon(img, poly, thr_param=90.0, min_area=40, tolerance=1.0, mode="percentile", eng=None):
    """roi_manual_drawer.segment_inside_polygon (:337-418): threshold-assisted auto-contour inside a
    hand-drawn polygon.  Returns (thr, None, best polygon as (N, 2) x/y array) or the reference's
    None placeholders."""
    eng = eng or _engine()
    mem = eng.mem
    img = np.asarray(img)
    H, W = img.shape[:2]
    poly_arr = np.asarray(poly, dtype=float)
    min_x, max_x = int(np.floor(np.min(poly_arr[:, 0]))), int(np.ceil(np.max(poly_arr[:, 0])))
    min_y, max_y = int(np.floor(np.min(poly_arr[:, 1]))), int(np.ceil(np.max(poly_arr[:, 1])))
    min_x, max_x = max(0, min_x), min(W, max_x)
    min_y, max_y = max(0, min_y), min(H, max_y)
    if max_x <= min_x or max_y <= min_y:
        return None, None, None
    sub = common.as_u16_plane(img[min_y:max_y, min_x:max_x], "image")
    sh, sw = sub.shape
    d_sub = mem.from_host(sub.reshape(1, sh, sw))
    # the polygon's mask over the slice (matplotlib rule, local coordinates)
    rm = eng.rasterize(geo.RULE_MPL, [geo.mpl_spec(poly_arr - [min_x, min_y], (sw, sh), store_full=True)], (sh, sw), 1,
                       want_union=False)
    reg = ops.regions_from_masks(rm)
    # statistics of the pixels under the mask
    sj = np.zeros(1, dtype=STAT_JOB)
    sj["src"], sj["n_views"] = SRC_U16, 1
    sj["bidx"] = -1
    bnd = mode.lower() == "bnd"
    sj["qkind"][0, 0], sj["q32"][0, 0] = QK_PCT, q32_of(90.0 if bnd else float(thr_param))
    so = eng.region_stats(reg, sj, rm.pool, sh, sw, planes=d_sub).host()[0]
    n = int(so["n"])
    if n == 0:
        return None, None, None
    thr_param = float(thr_param)
    if bnd:
        m = float(np.float32(so["sum"] / n))
        s = float(np.float32(np.sqrt(max(so["ssd"] / n, 0.0))))
        thr = float(so["q"][0]) if (s <= 0 or not np.isfinite(s)) else m + thr_param * s
    else:
        thr = float(so["q"][0])
    # (sub >= thr) & inside, 4-connected components (scipy.ndimage.label's default), their sizes:
    # for integer pixels v >= thr  <=>  v > the float32 just below thr
    t32 = np.float32(thr)
    if float(t32) < thr:                       # thr is not a float32: v >= thr <=> v > float32 below-or-equal thr
        thr_gt = t32
    else:
        thr_gt = np.nextafter(t32, np.float32(-np.inf))
    labels, areas, _ = _label_crop(eng, d_sub, sh, sw, rm.pool, thr_gt, 4)
    if areas.size == 0:
        return thr, None, None
    k = int(np.argmax(areas)) + 1
    lab = labels.host().reshape(sh, sw)
    mask_sub = lab == k
    # binary_fill_holes: background components (4-connected) that do not touch the border are holes
    inv = mem.from_host((~mask_sub).astype(np.uint16).reshape(1, sh, sw))
    blab, _, _ = _label_crop(eng, inv, sh, sw, _full_mask_pool(eng, sh, sw), np.float32(0.5), 4)
    bl = blab.host().reshape(sh, sw)
    border = np.unique(np.concatenate([bl[0], bl[-1], bl[:, 0], bl[:, -1]]))
    holes = (bl > 0) & ~np.isin(bl, border)
    filled = mask_sub | holes
    # contours of the filled mask from one pass over it
    if sh < 2 or sw < 2:
        raise ValueError("Input array must be at least 2x2.")
    crops = np.zeros(1, dtype=CROP)
    crops["w"], crops["h"], crops["wpr"] = sw, sh, (sw + 31) // 32
    d_fill = mem.from_host(filled.astype(np.int32).reshape(-1))
    rec, rec_n = mem.empty((sh * sw, 2), np.uint32), mem.empty(1, np.uint32)
    d_crops = mem.from_host(crops)                               # named: a temporary would be freed before the call runs
    eng.call("ipb_fa_contour_cells", d_crops.ptr, 1, sh * sw, d_fill.ptr, rec.ptr, rec_n.ptr, mem.stream)
    cont = ct.contours_of_crop(rec.host(), int(rec_n.host()[0]), sw).get(1, [])
    if not cont:
        return thr, None, None
    polys = []
    for c in cont:
        xy = np.c_[c[:, 1] + min_x, c[:, 0] + min_y]
        area = polygon_area(xy)
        if area >= float(min_area):
            xy_s = approximate_polygon(xy, tolerance=float(tolerance))
            if len(xy_s) >= 3:
                polys.append((area, xy_s))
    if not polys:
        return thr, None, None
    return thr, None, max(polys, key=lambda t: t[0])[1]


def render_pipeline(img, use_bandpass=False, sigma_small=1.2, sigma_large=9.0, use_unsharp=False, unsharp_amount=0.7,
                    unsharp_radius=2.0, eng=None):
    """ROIEditor._render_pipeline (:870-876) for a float32 image: band-pass and unsharp mask through
    the device Gaussian (bit-identical with scipy.ndimage.gaussian_filter)."""
    eng = eng or _engine()
    im = np.ascontiguousarray(img, dtype=np.float32)
    if not (use_bandpass or use_unsharp):
        return im
    out = filters.render_pipeline(eng, eng.mem.from_host(im), use_bandpass, sigma_small, sigma_large, use_unsharp,
                                  unsharp_amount, unsharp_radius)
    return out.host()
