"""Per-ROI shape, preview and crop products on the device:

  morphology_batch   MOR_by_ROI.morphology_from_polygon (reference src/MOR_by_ROI.py:211-241):
                     raster area + exact integer first / second moments on the device; the
                     O(vertices) polygon math (perimeter, hull, eigen-decomposition of a 2x2
                     matrix) stays host numpy, as in the reference
  preview_u16_batch  auto_minmax + 16-bit preview (INT/Fluor_INT.py:540-548,930-943;
                     FRET/fret_ratio_builder.py:364-369,479-483)
  cropper_batch      roi_channel_cropper.run_crop numeric body (roi_channel_cropper.py:884-968)
"""
import math

import numpy as np

from . import geometry as geo
from . import ops
from .device import DevBuf
from .ops import CROP_JOB, QK_PCT, REGION, SRC_F32, SRC_U16, STAT_JOB, q32_of


# ---------------------------------------------------------------------- host polygon math
def polygon_perimeter(poly):
    """MOR_by_ROI.py:166-170."""
    P = np.asarray(poly, dtype=float)
    dif = P[(np.arange(len(P)) + 1) % len(P)] - P
    return float(np.sqrt((dif ** 2).sum(axis=1)).sum())


def shoelace_area(poly):
    """MOR_by_ROI.py:172-175."""
    P = np.asarray(poly, dtype=float)
    x, y = P[:, 0], P[:, 1]
    return float(0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


def convex_hull(points):
    """MOR_by_ROI.py:177-191: Andrew's monotone chain on the unique, lexicographically sorted
    vertices; collinear points are dropped (cross <= 0)."""
    pts = np.unique(points, axis=0)
    pts = pts[np.lexsort((pts[:, 1], pts[:, 0]))]
    if len(pts) <= 1:
        return pts

    def turn(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    def chain(seq):
        out = []
        for q in seq:
            while len(out) >= 2 and turn(out[-2], out[-1], q) <= 0:
                out.pop()
            out.append(tuple(q))
        return out
    lower, upper = chain(pts), chain(pts[::-1])
    return np.array(lower[:-1] + upper[:-1], dtype=float)


def _cov_from_moments(n, sx, sy, sxx, syy, sxy):
    """np.cov (ddof 1) of the pixel coordinates from exact integer sums (Python ints)."""
    n, sx, sy, sxx, syy, sxy = (int(v) for v in (n, sx, sy, sxx, syy, sxy))
    den = n * (n - 1)
    return np.array([[(n * sxx - sx * sx) / den, (n * sxy - sx * sy) / den],
                     [(n * sxy - sx * sy) / den, (n * syy - sy * sy) / den]], dtype=np.float64)


def morphology_batch(eng, polys, shape, px_um):
    """List of the dicts morphology_from_polygon returns, one per polygon."""
    H, W = int(shape[0]), int(shape[1])
    specs = [geo.mpl_spec(P, (W, H)) for P in polys]
    rm = eng.rasterize(geo.RULE_MPL, specs, (H, W), 1, want_union=False)
    mom = eng.region_moments(ops.regions_from_masks(rm), rm.pool)
    nan = np.nan
    out = []
    for P, m in zip(polys, mom):
        n = int(m[0])
        area_px = float(n)
        if area_px == 0:
            out.append({"area_px": 0, "area_um2": 0, "perimeter_px": nan, "perimeter_um": nan,
                        "circularity": nan, "roundness": nan, "solidity": nan, "major_um": nan,
                        "minor_um": nan, "aspect_ratio": nan, "orientation_deg": nan,
                        "centroid_x": nan, "centroid_y": nan})
            continue
        area_um2 = area_px * (px_um ** 2)
        perimeter_px = float(polygon_perimeter(P))
        hull = convex_hull(np.asarray(P, dtype=float))
        if hull.shape[0] >= 3:
            area_hull_px = shoelace_area(hull)
            solidity = float(area_px / area_hull_px) if area_hull_px > 0 else nan
        else:
            solidity = nan
        if n < 2:
            # np.cov of a single point is non-finite; the reference then fails to unpack the
            # centroid (MOR_by_ROI.py:203-204,233) -- same exception here
            raise TypeError("cannot unpack non-iterable float object")
        xc, yc = int(m[1]) / n, int(m[2]) / n
        cov = _cov_from_moments(*m)
        w, v = np.linalg.eigh(cov)
        lam1, lam2 = w[1], w[0]
        angle = math.degrees(math.atan2(v[1, 1], v[0, 1]))
        major_um = 4.0 * math.sqrt(max(lam1, 0.0)) * px_um
        minor_um = 4.0 * math.sqrt(max(lam2, 0.0)) * px_um
        ok = np.isfinite(major_um) and np.isfinite(minor_um)
        out.append({
            "area_px": area_px, "area_um2": area_um2, "perimeter_px": perimeter_px,
            "perimeter_um": perimeter_px * px_um,
            "circularity": float(4.0 * math.pi * area_px / (perimeter_px ** 2)) if perimeter_px > 0 else nan,
            "roundness": float(4.0 * area_um2 / (math.pi * (major_um ** 2))) if (np.isfinite(major_um) and major_um > 0) else nan,
            "solidity": solidity, "major_um": major_um, "minor_um": minor_um,
            "aspect_ratio": float(major_um / minor_um) if (ok and minor_um > 0) else nan,
            "orientation_deg": angle, "centroid_x": float(xc), "centroid_y": float(yc)})
    return out


# ---------------------------------------------------------------------- full rect masks
def _ones_pool(rects_wh):
    """All-ones bit rows for rects [(w, h), ...]: (pool uint32, mask_off, wpr)."""
    offs, wprs, chunks = [0], [], []
    for w, h in rects_wh:
        wpr = (w + 31) // 32
        rows = np.full((h, wpr), 0xFFFFFFFF, dtype=np.uint32)
        if w & 31:
            rows[:, -1] = (1 << (w & 31)) - 1
        chunks.append(rows.reshape(-1))
        wprs.append(wpr)
        offs.append(offs[-1] + h * wpr)
    pool = np.concatenate(chunks) if chunks else np.zeros(1, np.uint32)
    return pool, np.asarray(offs[:-1], dtype=np.int64), np.asarray(wprs, dtype=np.int32)


# ---------------------------------------------------------------------- previews
def preview_u16_batch(eng, images, p_lo=1.0, p_hi=99.0):
    """images: host float32 [K][H][W] (or a DevBuf of that shape).  Returns a list of uint16
    [H][W] previews; None where the image has no finite value (the reference skips it)."""
    mem = eng.mem
    dev = images if isinstance(images, DevBuf) else mem.from_host(np.ascontiguousarray(images, dtype=np.float32))
    K, H, W = dev.shape
    pool, moff, wprs = _ones_pool([(W, H)])
    reg = np.zeros(K, dtype=REGION)
    reg["w"], reg["h"], reg["wpr"], reg["frame"] = W, H, wprs[0], np.arange(K)
    jobs = np.zeros(K, dtype=STAT_JOB)
    jobs["region"], jobs["src"], jobs["plane"], jobs["n_views"] = np.arange(K), SRC_F32, np.arange(K), 1
    jobs["bidx"] = -1
    jobs["qkind"] = (QK_PCT, 0, QK_PCT)
    jobs["q32"] = (q32_of(p_lo), 0.0, q32_of(p_hi))
    jobs["out"][:, 0] = np.arange(K)
    so = eng.region_stats(reg, jobs, mem.from_host(pool), H, W, images=dev).host()
    lohi = np.zeros((K, 3), dtype=np.float32)
    valid = so["n"] > 0
    for k in range(K):
        if not valid[k]:
            lohi[k] = (0.0, 1.0, 1.0)
            continue
        lo, hi = np.float32(so["q"][k, 0]), np.float32(so["q"][k, 2])
        if hi <= lo:
            hi = lo + 1e-6                      # np.float32 + python float stays float32 (NEP 50)
        lo, hi = float(lo), float(hi)
        lohi[k] = (lo, hi, np.float32(hi - lo + 1e-12))
    out = eng.preview_u16(dev, H * W, K, lohi).host().reshape(K, H, W)
    return [out[k] if valid[k] else None for k in range(K)]


# ---------------------------------------------------------------------- cropper
def cropper_batch(eng, raw, polys, low_cut, high_cut, gamma, mask_outside=True, pad_ratio=0.05):
    """raw: host uint16 (H, W) (the selected channel).  Per polygon: dict(norm_gamma float32,
    out16 uint16, raw_out, rect (x0, x1, y0, y1 inclusive), mask, lo, hi) or None where the
    reference 'continue's (roi_channel_cropper.py:905-921)."""
    mem = eng.mem
    raw = np.ascontiguousarray(raw)
    assert raw.dtype == np.uint16 and raw.ndim == 2
    H, W = raw.shape
    n = len(polys)
    if n == 0:
        return []
    rects, specs = [], []
    for P in polys:
        P = np.asarray(P, dtype=np.float64)
        x0, x1, y0, y1 = geo.trunc_crop_rect(P, W, H, pad_ratio)
        rects.append((x0, x1, y0, y1))
        P2 = P.copy()
        P2[:, 0] -= x0
        P2[:, 1] -= y0
        specs.append(geo.mpl_spec(P2, (x1 - x0 + 1, y1 - y0 + 1), org=(x0, y0), store_full=True))
    planes = mem.from_host(raw.reshape(1, H, W))
    rm = eng.rasterize(geo.RULE_MPL, specs, (H, W), 1, want_union=False)
    mreg = ops.regions_from_masks(rm)
    wh = [(x1 - x0 + 1, y1 - y0 + 1) for x0, x1, y0, y1 in rects]
    pool, moff, wprs = _ones_pool(wh)
    reg = np.zeros(n, dtype=REGION)
    reg["mask_off"] = moff
    reg["x0"], reg["y0"] = [r[0] for r in rects], [r[2] for r in rects]
    reg["w"], reg["h"], reg["wpr"] = [w for w, _ in wh], [h for _, h in wh], wprs
    jobs = np.zeros(n, dtype=STAT_JOB)
    jobs["region"], jobs["src"], jobs["plane"], jobs["n_views"] = np.arange(n), SRC_U16, 0, 1
    jobs["bidx"] = -1
    jobs["qkind"] = (QK_PCT, 0, QK_PCT)
    jobs["q32"] = (q32_of(low_cut), 0.0, q32_of(100.0 - high_cut))
    jobs["out"][:, 0] = np.arange(n)
    so = eng.region_stats(reg, jobs, mem.from_host(pool), H, W, planes=planes).host()
    params = np.zeros((n, 2), dtype=np.float32)
    lohi, ok = [], []
    for k in range(n):
        lo, hi = np.float32(so["q"][k, 0]), np.float32(so["q"][k, 2])
        good = so["n"][k] > 0
        if good and ((not np.isfinite(lo)) or (not np.isfinite(hi)) or (hi <= lo)):
            lo, hi = float(so["vmin"][k]), float(so["vmax"][k])
        if good and ((not np.isfinite(lo)) or (not np.isfinite(hi)) or (hi <= lo)):
            good = False
        ok.append(bool(good))
        lohi.append((lo, hi))
        if good:
            params[k] = (np.float32(lo), np.float32(np.float32(hi) - np.float32(lo)) if isinstance(lo, np.float32)
                         else np.float32(hi - lo))
        else:
            params[k] = (0.0, 1.0)
    cj = np.zeros(n, dtype=CROP_JOB)
    px = np.array([w * h for w, h in wh], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(px)])
    cj["plane"] = 0
    cj["x0"], cj["y0"] = reg["x0"], reg["y0"]
    cj["w"], cj["h"] = reg["w"], reg["h"]
    cj["region"] = np.arange(n) if mask_outside else -1
    cj["out_off"] = off[:-1]
    d_cj, d_par, d_mreg = mem.from_host(cj), mem.from_host(params.reshape(-1)), mem.from_host(mreg)
    o_norm = mem.empty(int(off[-1]), np.float32)
    o_16 = mem.empty(int(off[-1]), np.uint16)
    eng.call("ipb_crop_normalize", d_cj.ptr, n, int(px.max()), planes.ptr, H, W, d_par.ptr,
             float(np.float32(1.0 / float(gamma))), d_mreg.ptr, rm.pool.ptr, o_norm.ptr, o_16.ptr, mem.stream)
    h_norm, h_16 = o_norm.host(), o_16.host()
    out = []
    for k in range(n):
        if not ok[k]:
            out.append(None)
            continue
        x0, x1, y0, y1 = rects[k]
        w, h = wh[k]
        mask = rm.mask_host(k)
        raw_out = raw[y0:y1 + 1, x0:x1 + 1].copy()
        if mask_outside:
            raw_out[~mask] = 0
        out.append({"norm_gamma": h_norm[off[k]: off[k + 1]].reshape(h, w),
                    "out16": h_16[off[k]: off[k + 1]].reshape(h, w), "raw_out": raw_out,
                    "rect": rects[k], "mask": mask, "lo": lohi[k][0], "hi": lohi[k][1]})
    return out
