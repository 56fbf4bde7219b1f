"""oracle/port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's numeric functions on the hot path
(SURVEY.md 8(a)).  The reference is Python that cannot be imported on the GPU box
(no /root/reference there, and matplotlib / scikit-image / tifffile / tkinter are not
installed anywhere), so the arithmetic is restated here with the same numpy / scipy
calls in the same order -- that is what makes the floating-point results identical.
Each function cites the reference lines it follows; tests/test_oracle_vs_reference.py
runs every one of them against the unmodified reference function in the build
container, and tests/test_oracle_golden.py against the reference's shipped outputs.

All paths below are relative to /root/reference/src/.
"""
import math

import numpy as np
from scipy.ndimage import binary_dilation, distance_transform_edt

from . import shims


# ============================================================ shared: ROI raster (a1)
def rasterize_polygon(poly, shape):
    """INT/Fluor_INT.py:398-403 (same body: FRET/fret_ratio_builder.py:292-296,
    FRET/Nesprin2_FRET_Builder.py:388-393, MOR_by_ROI.py:160-164,
    roi_channel_cropper.py:286-291)."""
    H, W = shape
    yy, xx = np.mgrid[0:H, 0:W]
    pts = np.vstack((xx.ravel(), yy.ravel())).T
    return shims.Path(np.asarray(poly, dtype=float)).contains_points(pts).reshape(H, W)


def valid_polys(rois):
    """ROI JSON -> list of (V,2) float arrays, polygons with < 3 points dropped
    (INT/Fluor_INT.py:417-422)."""
    out = []
    for poly in rois:
        P = np.asarray(poly, dtype=float)
        if P.shape[0] >= 3:
            out.append(P)
    return out


# ============================================================ Fluor_INT (a4, a5, a15)
def _hist_mode_level(vals, p):
    """'hist-mode' branch shared by the three bg_value copies
    (INT/Fluor_INT.py:474-483)."""
    hist, bins = np.histogram(vals, bins=2048)
    if hist.sum() <= 0:
        return float(np.percentile(vals, p))
    cdf = np.cumsum(hist).astype(float)
    cdf /= cdf[-1]
    idx = int(np.searchsorted(cdf, float(p) / 100.0, side="left"))
    if idx >= len(bins) - 1:
        return float(bins[-1])
    return float(0.5 * (bins[idx] + bins[idx + 1]))


def int_bg_value(img2d, mode="percentile", p=1.0, scope_mask=None, stride=4):
    """INT/Fluor_INT.py:464-485."""
    vals = img2d.ravel() if scope_mask is None else img2d[scope_mask]
    if vals.size == 0:
        return 0.0
    if stride and stride > 1:
        vals = vals[::int(stride)]
        if vals.size == 0:
            return 0.0
    if mode == "percentile":
        return float(np.percentile(vals, p))
    if mode == "hist-mode":
        return _hist_mode_level(vals, p)
    return 0.0


def int_bg_correct(img2d, mode="percentile", p=1.0, scope_mask=None, clip_neg=True, stride=4):
    """INT/Fluor_INT.py:487-492."""
    B = int_bg_value(img2d, mode=mode, p=p, scope_mask=scope_mask, stride=stride)
    J = img2d - B
    if clip_neg:
        J[J < 0] = 0.0
    return J, B


def quantify_stats(vals):
    """INT/Fluor_INT.py:494-507."""
    vals = vals[np.isfinite(vals)]
    if vals.size == 0:
        nan = np.nan
        return dict(mean=nan, median=nan, std=nan, p5=nan, p95=nan, vmin=nan, vmax=nan,
                    vsum=nan, npx=0)
    return dict(mean=float(np.mean(vals)), median=float(np.median(vals)),
                std=float(np.std(vals)), p5=float(np.percentile(vals, 5)),
                p95=float(np.percentile(vals, 95)), vmin=float(np.min(vals)),
                vmax=float(np.max(vals)), vsum=float(np.sum(vals)), npx=int(vals.size))


def quantify_per_roi_multi(images_dict, polys=None, union_mask=None):
    """INT/Fluor_INT.py:509-538."""
    any_img = next(iter(images_dict.values()))
    H, W = any_img.shape
    if polys is not None:
        todo = [(i, rasterize_polygon(poly, (H, W))) for i, poly in enumerate(polys, 1)]
    elif union_mask is not None:
        todo = [(1, union_mask.astype(bool, copy=False))]
    else:
        todo = [(0, np.ones_like(any_img, dtype=bool))]
    rows = []
    for i, m in todo:
        row = {"roi": i, "area_px": int(m.sum())}
        for ch, img in sorted(images_dict.items()):
            for k, v in quantify_stats(img[m]).items():
                row[f"ch{ch}_{k}"] = v
        rows.append(row)
    return rows


def auto_minmax(vals, p_lo=1.0, p_hi=99.0):
    """INT/Fluor_INT.py:540-548 (copies: fret_ratio_builder.py:364-369,
    Nesprin2_FRET_Builder.py:478-486)."""
    vals = vals[np.isfinite(vals)]
    if vals.size == 0:
        return 0.0, 1.0
    lo = np.percentile(vals, p_lo)
    hi = np.percentile(vals, p_hi)
    if hi <= lo:
        hi = lo + 1e-6
    return float(lo), float(hi)


def preview_u16(img, p_lo=1.0, p_hi=99.0):
    """16-bit preview: INT/Fluor_INT.py:930-943, FRET/fret_ratio_builder.py:479-483.
    Returns None when no finite value exists (the reference then skips / writes zeros)."""
    vals = img[np.isfinite(img)]
    if vals.size == 0:
        return None
    lo, hi = auto_minmax(vals, p_lo, p_hi)
    clip_ = np.clip(img, lo, hi)
    norm = (clip_ - lo) / (hi - lo + 1e-12)
    with np.errstate(invalid="ignore"):
        return (norm * 65535).astype(np.uint16)


def int_process_key(imgs_raw, polys, union_mask, task):
    """Numeric core of _process_key_task, INT/Fluor_INT.py:839-870 (+908-943 for the
    TIFF products).  imgs_raw: {ch: f32 HxW}.  Returns (per_roi rows, bg_used,
    imgs_bc)."""
    any_img = next(iter(imgs_raw.values()))
    H, W = any_img.shape
    scope_mask = None
    if task["bg_scope"] == "roi_union":
        if polys is not None:
            u = np.zeros((H, W), dtype=bool)
            for P in polys:
                u |= rasterize_polygon(P, (H, W))
            scope_mask = u
        elif union_mask is not None:
            scope_mask = union_mask
    imgs_bc, bg_used = {}, {}
    p_glob = float(task["percentile"])
    for ch, img in imgs_raw.items():
        pp = float(task["ch_p_map"].get(ch, p_glob)) if task.get("per_channel_p") else p_glob
        bc, B = int_bg_correct(img, mode=task["bg_mode"], p=pp, scope_mask=scope_mask,
                               clip_neg=task["clip_neg"], stride=int(task["bg_stride"]))
        imgs_bc[ch] = bc
        bg_used[ch] = {"bg": float(B), "p": float(pp)}
    per_roi = quantify_per_roi_multi(imgs_bc, polys=polys, union_mask=union_mask)
    return per_roi, bg_used, imgs_bc


# ============================================================ FA_Analyzer (a2, a3, a6, a7)
def fa_global_stats(img):
    """INT/FA_Analyzer.py:984-987 (same: 623-626)."""
    img_float = img.astype(np.float32)
    sample = img_float[::10, ::10]
    bg_val = np.percentile(sample, 1.0)
    return (np.nanmean(img_float), np.nanstd(img_float), bg_val)


def fa_crop_rect(roi_poly, img_shape, pad=5):
    """INT/FA_Analyzer.py:998-1003: floor/ceil bbox + pad, clipped, half-open."""
    xs = roi_poly[:, 0]
    ys = roi_poly[:, 1]
    x_min, x_max = int(np.floor(xs.min())), int(np.ceil(xs.max()))
    y_min, y_max = int(np.floor(ys.min())), int(np.ceil(ys.max()))
    x_min = max(0, x_min - pad)
    x_max = min(img_shape[1], x_max + pad)
    y_min = max(0, y_min - pad)
    y_max = min(img_shape[0], y_max + pad)
    return x_min, x_max, y_min, y_max


def fa_crop_and_mask(img, roi_poly):
    """INT/FA_Analyzer.py:1005-1015."""
    x_min, x_max, y_min, y_max = fa_crop_rect(roi_poly, img.shape)
    if x_min >= x_max or y_min >= y_max:
        img_crop = np.array([])
    else:
        img_crop = img[y_min:y_max, x_min:x_max]
    poly_crop = roi_poly.copy()
    poly_crop[:, 0] -= x_min
    poly_crop[:, 1] -= y_min
    mask_crop = np.zeros(img_crop.shape, dtype=bool)
    rr, cc = shims.polygon(poly_crop[:, 1], poly_crop[:, 0], img_crop.shape)
    mask_crop[rr, cc] = True
    return img_crop, mask_crop, (x_min, x_max, y_min, y_max)


def analyze_fa_crop(image_crop, roi_mask_crop, config, global_stats, with_contours=True):
    """INT/FA_Analyzer.py:123-195.  with_contours=False skips the per-adhesion
    find_contours call (drawing only; never empty for a non-empty region)."""
    if image_crop.size == 0 or image_crop.shape[0] == 0 or image_crop.shape[1] == 0:
        return ({'OK': [], 'Large': [], 'Small': []}, 0.0,
                np.zeros_like(image_crop, dtype=bool), np.zeros_like(image_crop, dtype=int))
    img_float = image_crop.astype(np.float32)
    bg_val_passed = None
    if len(global_stats) == 3:
        m, s, bg_val_passed = global_stats
    else:
        m, s = global_stats
    bg_val = bg_val_passed if bg_val_passed is not None else np.percentile(img_float, 1.0)
    threshold_val = m + config['alpha'] * s
    bw = img_float > threshold_val
    bw = bw & roi_mask_crop
    min_px = config['min_px']
    if min_px > 0:
        bw = shims.remove_small_objects(bw, min_size=min_px)
    close_rad = config['close_radius']
    if close_rad > 0:
        bw = shims.binary_closing(bw, shims.disk(close_rad))
    labeled_img = shims.label(bw)
    props = shims.regionprops(labeled_img, intensity_image=img_float)
    max_px = config['max_px']
    subtract_bg = config.get('subtract_bg', True)
    results = {'OK': [], 'Large': [], 'Small': []}
    for prop in props:
        area = prop.area
        contour = None
        if with_contours:
            contours = shims.find_contours(labeled_img == prop.label, 0.5)
            if not contours:
                continue
            contour = contours[0]
        category = 'OK'
        if area < min_px:
            category = 'Small'
        elif area > max_px:
            category = 'Large'
        mean_raw = prop.mean_intensity
        mean_corr = max(0, mean_raw - bg_val) if subtract_bg else mean_raw
        results[category].append({
            'label': prop.label, 'area': area, 'contour': contour, 'centroid': prop.centroid,
            'mean_int_raw': mean_raw, 'mean_int_corr': mean_corr,
            'int_den_raw': mean_raw * area, 'int_den_corr': mean_corr * area,
            'bg_level': bg_val})
    return results, threshold_val, bw, labeled_img


def fa_um_to_px_config(params, px_size):
    """INT/FA_Analyzer.py:527-535."""
    return {'alpha': params['alpha'], 'min_px': params['min_area_um'] / (px_size ** 2),
            'max_px': params['max_area_um'] / (px_size ** 2),
            'close_radius': params['close_radius'],
            'subtract_bg': params.get('subtract_bg', True)}


def fa_batch_rows(img, rois, params, px_size, s_tag="S01", save_ok_only=True,
                  with_contours=True, stats=None):
    """Per-file body of _run_batch_process, INT/FA_Analyzer.py:984-1039.  `stats` overrides
    the global (mean, std, bg) tuple -- used by parity tests when the float32 threshold built
    from exact integer moments differs from numpy's pairwise float32 one by an ulp."""
    config = fa_um_to_px_config(params, px_size)
    if stats is None:
        stats = fa_global_stats(img)
    rows = []
    for i, roi_poly in enumerate(rois):
        roi_poly = np.array(roi_poly, dtype=float)
        img_crop, mask_crop, _ = fa_crop_and_mask(img, roi_poly)
        res, th_val, _, _ = analyze_fa_crop(img_crop, mask_crop, config, stats,
                                            with_contours=with_contours)
        for cat, items in res.items():
            if save_ok_only and cat != 'OK':
                continue
            for item in items:
                rows.append({
                    'File': s_tag, 'Cell_ID': i + 1, 'Category': cat,
                    'Area_px': item['area'], 'Area_um2': item['area'] * (px_size ** 2),
                    'Mean_Intensity_Raw': item['mean_int_raw'],
                    'Mean_Intensity_Corr': item['mean_int_corr'],
                    'Int_Density_Raw': item['int_den_raw'],
                    'Int_Density_Corr': item['int_den_corr'],
                    'Background_Level': item['bg_level'],
                    'Used_Alpha': params['alpha'], 'Global_Threshold': th_val,
                    'Min_Area_Setting': params['min_area_um'],
                    'Max_Area_Setting': params['max_area_um'],
                    'Close_Radius_Setting': params['close_radius'],
                    'Subtract_BG_Setting': params.get('subtract_bg', True)})
    return rows


# ============================================================ fret_ratio_builder (a4, a8, a9, a14)
def fret_bg_value(img2d, mode="percentile", p=1.0, scope_mask=None):
    """FRET/fret_ratio_builder.py:314-330 (no stride)."""
    vals = img2d.ravel() if scope_mask is None else img2d[scope_mask]
    if vals.size == 0:
        return 0.0
    if mode == "percentile":
        return float(np.percentile(vals, p))
    if mode == "hist-mode":
        return _hist_mode_level(vals, p)
    return 0.0


def fret_bg_correct(img2d, mode="percentile", p=1.0, scope_mask=None, clip_neg=True):
    """FRET/fret_ratio_builder.py:332-336."""
    B = fret_bg_value(img2d, mode=mode, p=p, scope_mask=scope_mask)
    J = img2d - B
    if clip_neg:
        J[J < 0] = 0.0
    return J, B


def pick_epsilon(denom_vals, eps_abs=5.0, p_floor=1.0):
    """FRET/fret_ratio_builder.py:338-340."""
    if denom_vals.size == 0:
        return float(eps_abs)
    return float(max(eps_abs, np.percentile(denom_vals, p_floor)))


def fret_quantify_per_roi(R, polys, extra_imgs=None):
    """FRET/fret_ratio_builder.py:342-362."""
    rows = []
    H, W = R.shape
    for i, poly in enumerate(polys, 1):
        m = rasterize_polygon(poly, (H, W))
        vals = R[m]
        vals = vals[np.isfinite(vals)]
        row = {"roi": i, "area_px": int(m.sum())}
        if vals.size == 0:
            for k in ("mean", "median", "std", "p5", "p95"):
                row[f"ratio_{k}"] = np.nan
        else:
            row.update({"ratio_mean": float(np.mean(vals)), "ratio_median": float(np.median(vals)),
                        "ratio_std": float(np.std(vals)),
                        "ratio_p5": float(np.percentile(vals, 5)),
                        "ratio_p95": float(np.percentile(vals, 95))})
        if extra_imgs:
            for name, img in extra_imgs.items():
                iv = img[m].astype(np.float32)
                row[f"{name}_mean"] = float(np.mean(iv)) if iv.size else np.nan
                row[f"{name}_median"] = float(np.median(iv)) if iv.size else np.nan
        rows.append(row)
    return rows


def fret_process_pair(D, A, polys, p):
    """Numeric body of process_one_stage for one (S,t) pair,
    FRET/fret_ratio_builder.py:454-474,493-507.  D, A: f32 HxW.  Returns dict with
    Db, Ab, eps, R_full, R_roi (or None), union (or None), rows."""
    H, W = D.shape
    union = None
    if polys:
        union = np.zeros((H, W), dtype=bool)
        for P in polys:
            union |= rasterize_polygon(P, (H, W))
    scope_mask = union if (p["bg_scope"] == "roi_union" and union is not None) else None
    per_ch = bool(p["per_channel_p"])
    d_p = float(p["donor_p"]) if per_ch else float(p["percentile"])
    a_p = float(p["fret_p"]) if per_ch else float(p["percentile"])
    clip_neg = bool(p["clip_neg"])
    Dbc, Db = fret_bg_correct(D, mode=p["bg_mode"], p=d_p, scope_mask=scope_mask, clip_neg=clip_neg)
    Abc, Ab = fret_bg_correct(A, mode=p["bg_mode"], p=a_p, scope_mask=scope_mask, clip_neg=clip_neg)
    if p["ratio_mode"] == "FRET/Donor":
        numer, denom = Abc, Dbc
    else:
        numer, denom = Dbc, Abc
    denom_vals = denom[scope_mask] if scope_mask is not None else denom.ravel()
    eps = pick_epsilon(denom_vals, eps_abs=5.0, p_floor=p["eps_percentile"])
    R_full = (numer + eps) / (denom + eps)
    out = {"Db": Db, "Ab": Ab, "eps": eps, "R_full": R_full, "union": union, "R_roi": None,
           "rows": [], "Dbc": Dbc, "Abc": Abc}
    if polys:
        R_roi = R_full.copy()
        if union is not None:
            R_roi[~union] = np.nan
        out["R_roi"] = R_roi
        out["rows"] = fret_quantify_per_roi(R_full, polys, extra_imgs={"donor": Dbc, "yfret": Abc})
    return out


# ============================================================ Nesprin2 (a10-a13)
def make_inside_rim_mask(union_mask, rim_px):
    """FRET/Nesprin2_FRET_Builder.py:409-414."""
    if rim_px <= 0:
        return union_mask.copy()
    dist_in = distance_transform_edt(union_mask)
    return (dist_in > 0) & (dist_in <= rim_px)


def annulus_mask_from_poly(poly, shape, inner_px, outer_px):
    """FRET/Nesprin2_FRET_Builder.py:416-427."""
    H, W = shape
    base = rasterize_polygon(poly, (H, W))
    if inner_px < 1:
        inner_px = 1
    if outer_px <= inner_px:
        outer_px = inner_px + 1
    se_out = np.ones((2 * outer_px + 1, 2 * outer_px + 1), dtype=bool)
    se_in = np.ones((2 * inner_px + 1, 2 * inner_px + 1), dtype=bool)
    return binary_dilation(base, structure=se_out) & (~binary_dilation(base, structure=se_in))


def n2_bg_value(img2d, mode="percentile", p=1.0, scope_mask=None):
    """FRET/Nesprin2_FRET_Builder.py:432-451 (drops non-finite first)."""
    vals = img2d.ravel() if scope_mask is None else img2d[scope_mask]
    if vals.size == 0:
        return 0.0
    vals = vals[np.isfinite(vals)]
    if vals.size == 0:
        return 0.0
    if mode == "percentile":
        return float(np.percentile(vals, p))
    if mode == "hist-mode":
        return _hist_mode_level(vals, p)
    return 0.0


def n2_bg_correct(img2d, mode="percentile", p=1.0, scope_mask=None, clip_neg=True):
    """FRET/Nesprin2_FRET_Builder.py:453-458."""
    B = n2_bg_value(img2d, mode=mode, p=p, scope_mask=scope_mask)
    J = img2d - B
    if clip_neg:
        J[J < 0] = 0.0
    return J, B


def spectral_correct(yfret, donor, acceptor_only=None, alpha=0.0, beta=0.0, g_factor=1.0):
    """FRET/Nesprin2_FRET_Builder.py:460-468."""
    yf = yfret.astype(np.float32, copy=False)
    d = donor.astype(np.float32, copy=False)
    if acceptor_only is not None:
        ao = acceptor_only.astype(np.float32, copy=False)
        yf_corr = yf - alpha * d - beta * ao
    else:
        yf_corr = yf - alpha * d
    return d, yf_corr * float(g_factor)


def n2_pick_epsilon(denom_vals, eps_abs=5.0, p_floor=1.0):
    """FRET/Nesprin2_FRET_Builder.py:470-476."""
    if denom_vals.size == 0:
        return float(eps_abs)
    denom_vals = denom_vals[np.isfinite(denom_vals)]
    if denom_vals.size == 0:
        return float(eps_abs)
    return float(max(eps_abs, np.percentile(denom_vals, p_floor)))


def n2_process_pair(D, A, polys, p, Aonly=None):
    """Numeric body of run_pipeline's per-pair loop,
    FRET/Nesprin2_FRET_Builder.py:1415-1421,1445-1581.  D, A (, Aonly): f32 HxW.
    Returns dict(R_full, R_alt, rim_mask, union, eps, rows).  Row fields follow
    1537-1581 minus the constant echo columns; defect 2 of SURVEY.md 8(a) (ratio_mode
    compared with 'DoverF') is kept, defect 1 ('time' holds a function) is not
    reproducible in a table and is left out."""
    px_um = float(p["px_um"])
    rim_px = max(1, int(round(float(p["rim_um"]) / px_um)))
    ann_on = bool(p["annulus_on"])
    ann_in_px = max(1, int(round(float(p["ann_in_um"]) / px_um))) if ann_on else 0
    ann_out_px = max(ann_in_px + 1, int(round(float(p["ann_out_um"]) / px_um))) if ann_on else 0
    scope = p["bg_scope"]
    clip_neg = bool(p["clip_neg"])
    clip_on, clip_max = bool(p["clip_ratio_on"]), float(p["clip_ratio_max"])
    if bool(p["sat_filter_on"]):
        sat_thr = float(p["sat_threshold"])
        mask_sat = (D >= sat_thr) | (A >= sat_thr)
        if np.any(mask_sat):
            D = D.astype(np.float32, copy=True)
            A = A.astype(np.float32, copy=True)
            D[mask_sat] = np.nan
            A[mask_sat] = np.nan
    H, W = D.shape
    union = np.zeros((H, W), dtype=bool)
    for P in polys:
        union |= rasterize_polygon(P, (H, W))
    scope_mask = None if scope == "full" else union
    per_ch = bool(p["per_channel_p"])
    p_glob = float(p["percentile"])
    d_p = float(p["donor_p"]) if per_ch else p_glob
    a_p = float(p["fret_p"]) if per_ch else p_glob
    Dbc, _ = n2_bg_correct(D, mode=p["bg_mode"], p=d_p, scope_mask=scope_mask, clip_neg=clip_neg)
    Abc, _ = n2_bg_correct(A, mode=p["bg_mode"], p=a_p, scope_mask=scope_mask, clip_neg=clip_neg)
    Aonly_bc = None
    if Aonly is not None:
        Aonly_bc, _ = n2_bg_correct(Aonly, mode=p["bg_mode"], p=p_glob, scope_mask=scope_mask,
                                    clip_neg=clip_neg)
    if bool(p["use_spectral"]):
        Dcorr, Acorr = spectral_correct(Abc, Dbc, acceptor_only=Aonly_bc, alpha=float(p["alpha"]),
                                        beta=float(p["beta"]), g_factor=float(p["g_factor"]))
    else:
        Dcorr, Acorr = Dbc, Abc
    fd = p["ratio_mode"] == "FRET/Donor"
    eps = n2_pick_epsilon(Dcorr[union] if fd else Acorr[union], eps_abs=5.0,
                          p_floor=float(p["eps_percentile"]))
    if fd:
        numer, denom, numer_alt, denom_alt = Acorr, Dcorr, Dcorr, Acorr
    else:
        numer, denom, numer_alt, denom_alt = Dcorr, Acorr, Acorr, Dcorr
    R_full = (numer + eps) / (denom + eps)
    R_alt = (numer_alt + eps) / (denom_alt + eps)
    if clip_on:
        R_full = np.where(R_full > clip_max, np.nan, R_full)
        R_alt = np.where(R_alt > clip_max, np.nan, R_alt)
    rim_mask = make_inside_rim_mask(union, rim_px)
    rows = []
    for i, P in enumerate(polys, start=1):
        roi_mask = rasterize_polygon(P, (H, W)) & rim_mask
        R_roi, R_roi_alt = R_full, R_alt
        if scope == "annulus" or ann_on:
            ann = annulus_mask_from_poly(P, (H, W), inner_px=ann_in_px, outer_px=ann_out_px)

            def _med(img):
                return np.nanmedian(img[ann]) if np.isfinite(img[ann]).any() else 0.0
            bg_n, bg_d, bg_na, bg_da = _med(numer), _med(denom), _med(numer_alt), _med(denom_alt)

            def _eff(img, b):
                return np.maximum(img - b, 0.0) if clip_neg else (img - b)
            R_roi = (_eff(numer, bg_n) + eps) / (_eff(denom, bg_d) + eps)
            R_roi_alt = (_eff(numer_alt, bg_na) + eps) / (_eff(denom_alt, bg_da) + eps)
            if clip_on:
                R_roi = np.where(R_roi > clip_max, np.nan, R_roi)
                R_roi_alt = np.where(R_roi_alt > clip_max, np.nan, R_roi_alt)
        vals = R_roi[roi_mask]
        vals = vals[np.isfinite(vals)]
        vals_alt = R_roi_alt[roi_mask]
        vals_alt = vals_alt[np.isfinite(vals_alt)]
        with np.errstate(invalid="ignore"), np.testing.suppress_warnings() as sup:
            sup.filter(RuntimeWarning)
            mean_main = float(np.nanmean(vals)) if vals.size else np.nan
            mean_alt = float(np.nanmean(vals_alt)) if vals_alt.size else np.nan
        # defect 2: ratio_mode is never "DoverF", so FoverD_mean <- primary, DoverF_mean <- alt
        row = {"roi": i, "area_px": int(roi_mask.sum()),
               "ratio_FoverD_mean": mean_main, "ratio_DoverF_mean": mean_alt, "eps": eps}
        if vals.size == 0:
            row.update({"ratio_mean": np.nan, "ratio_median": np.nan, "ratio_std": np.nan,
                        "ratio_p5": np.nan, "ratio_p95": np.nan,
                        "donor_mean": np.nan, "fret_mean": np.nan})
        else:
            with np.testing.suppress_warnings() as sup:
                sup.filter(RuntimeWarning)
                row.update({"ratio_mean": float(np.mean(vals)),
                            "ratio_median": float(np.median(vals)),
                            "ratio_std": float(np.std(vals)),
                            "ratio_p5": float(np.percentile(vals, 5)),
                            "ratio_p95": float(np.percentile(vals, 95)),
                            "donor_mean": float(np.nanmean(Dcorr[roi_mask])),
                            "fret_mean": float(np.nanmean(Acorr[roi_mask]))})
        rows.append(row)
    return {"R_full": R_full, "R_alt": R_alt, "rim_mask": rim_mask, "union": union,
            "eps": eps, "rows": rows, "Dcorr": Dcorr, "Acorr": Acorr}


# ============================================================ MOR_by_ROI (a17)
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
    """MOR_by_ROI.py:177-191 (monotone chain)."""
    pts = np.unique(points, axis=0)
    pts = pts[np.lexsort((pts[:, 1], pts[:, 0]))]
    if len(pts) <= 1:
        return pts

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lower, upper = [], []
    for q in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], q) <= 0:
            lower.pop()
        lower.append(tuple(q))
    for q in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], q) <= 0:
            upper.pop()
        upper.append(tuple(q))
    return np.array(lower[:-1] + upper[:-1], dtype=float)


def second_moments(mask):
    """MOR_by_ROI.py:193-199."""
    ys, xs = np.nonzero(mask)
    if xs.size == 0:
        return (np.nan, np.nan), np.array([[np.nan, np.nan], [np.nan, np.nan]])
    xc, yc = xs.mean(), ys.mean()
    cov = np.cov(np.vstack([xs - xc, ys - yc]))
    return (yc, xc), cov


def major_minor_axes_um(mask, px_um):
    """MOR_by_ROI.py:201-209."""
    (yc, xc), cov = second_moments(mask)
    if not np.isfinite(cov).all():
        return np.nan, np.nan, np.nan, np.nan
    w, v = np.linalg.eigh(cov)
    lam1, lam2 = w[1], w[0]
    angle = math.degrees(math.atan2(v[1, 1], v[0, 1]))
    a = 4.0 * math.sqrt(max(lam1, 0.0))
    b = 4.0 * math.sqrt(max(lam2, 0.0))
    return a * px_um, b * px_um, angle, (yc, xc)


def morphology_from_polygon(poly, shape, px_um):
    """MOR_by_ROI.py:211-241."""
    H, W = shape
    mask = rasterize_polygon(poly, (H, W))
    area_px = float(mask.sum())
    nan = np.nan
    if area_px == 0:
        return {"area_px": 0, "area_um2": 0, "perimeter_px": nan, "perimeter_um": nan,
                "circularity": nan, "roundness": nan, "solidity": nan, "major_um": nan,
                "minor_um": nan, "aspect_ratio": nan, "orientation_deg": nan,
                "centroid_x": nan, "centroid_y": nan}
    area_um2 = area_px * (px_um ** 2)
    perimeter_px = float(polygon_perimeter(poly))
    hull = convex_hull(np.asarray(poly, dtype=float))
    if hull.shape[0] >= 3:
        area_hull_px = shoelace_area(hull)
        solidity = float(area_px / area_hull_px) if area_hull_px > 0 else nan
    else:
        solidity = nan
    major_um, minor_um, orientation_deg, (cy, cx) = major_minor_axes_um(mask, px_um)
    ok = np.isfinite(major_um) and np.isfinite(minor_um)
    return {"area_px": area_px, "area_um2": area_um2, "perimeter_px": perimeter_px,
            "perimeter_um": perimeter_px * px_um,
            "circularity": float(4.0 * math.pi * area_px / (perimeter_px ** 2)) if perimeter_px > 0 else nan,
            "roundness": float(4.0 * area_um2 / (math.pi * (major_um ** 2))) if (np.isfinite(major_um) and major_um > 0) else nan,
            "solidity": solidity, "major_um": major_um, "minor_um": minor_um,
            "aspect_ratio": float(major_um / minor_um) if (ok and minor_um > 0) else nan,
            "orientation_deg": orientation_deg, "centroid_x": float(cx), "centroid_y": float(cy)}


# ============================================================ roi_channel_cropper (a3, a16)
def trunc_crop_rect(P, W, H, pad_ratio=0.05):
    """int() truncation + pad max(10, ratio*max(W,H)), inclusive ends:
    roi_channel_cropper.py:885-893 (same: INT/Fluor_INT.py:1028-1041,
    fret_ratio_builder.py:517-522, Nesprin2_FRET_Builder.py:1586-1594)."""
    P = np.asarray(P)
    minx, maxx = P[:, 0].min(), P[:, 0].max()
    miny, maxy = P[:, 1].min(), P[:, 1].max()
    pad = max(10, int(pad_ratio * max(W, H)))
    x0 = max(int(minx) - pad, 0)
    x1 = min(int(maxx) + pad, W - 1)
    y0 = max(int(miny) - pad, 0)
    y1 = min(int(maxy) + pad, H - 1)
    return x0, x1, y0, y1


def cropper_normalize(img, raw_full, P, low_cut, high_cut, gamma, mask_outside=True,
                      pad_ratio=0.05):
    """Numeric body of run_crop for one ROI, roi_channel_cropper.py:884-968.
    Returns dict(norm_gamma f32, out16 u16, raw_out, rect) or None where the
    reference 'continue's."""
    H, W = img.shape
    P = np.asarray(P)
    x0, x1, y0, y1 = trunc_crop_rect(P, W, H, pad_ratio)
    crop_f32 = img[y0:y1 + 1, x0:x1 + 1].copy()
    crop_raw = raw_full[y0:y1 + 1, x0:x1 + 1].copy()
    P2 = P.copy()
    P2[:, 0] -= x0
    P2[:, 1] -= y0
    local_mask = rasterize_polygon(P2, crop_f32.shape)
    vals = crop_f32[np.isfinite(crop_f32)]
    if vals.size == 0:
        return None
    lo = np.percentile(vals, low_cut)
    hi = np.percentile(vals, 100.0 - high_cut)
    if (not np.isfinite(lo)) or (not np.isfinite(hi)) or (hi <= lo):
        lo = float(np.nanmin(vals))
        hi = float(np.nanmax(vals))
    if (not np.isfinite(lo)) or (not np.isfinite(hi)) or (hi <= lo):
        return None
    norm = np.clip((crop_f32 - lo) / (hi - lo), 0.0, 1.0)
    if mask_outside:
        norm = norm * local_mask.astype(np.float32)
    norm_gamma = np.power(norm, 1.0 / float(gamma))
    out16 = (np.clip(norm_gamma, 0, 1) * 65535).astype(np.uint16)
    raw_out = crop_raw.copy()
    if mask_outside:
        raw_out[~local_mask] = 0
    return {"norm_gamma": norm_gamma, "out16": out16, "raw_out": raw_out,
            "rect": (x0, x1, y0, y1), "mask": local_mask, "lo": lo, "hi": hi}


# ---------------------------------------------------------------- ROI drawer assist (roi_manual_drawer.py:337-418)
def segment_inside_polygon(img, poly, thr_param=90.0, min_area=40, tolerance=1.0, mode="percentile"):
    """roi_manual_drawer.segment_inside_polygon: bounding-box slice, matplotlib mask, percentile or
    mean + k * std threshold, largest 4-connected component, hole fill, contours, polygon area,
    Douglas-Peucker; returns (thr, None, best polygon)."""
    from scipy import ndimage as ndi
    H, W = img.shape[:2]
    poly_arr = np.asarray(poly)
    min_x = int(np.floor(np.min(poly_arr[:, 0]))); max_x = int(np.ceil(np.max(poly_arr[:, 0])))
    min_y = int(np.floor(np.min(poly_arr[:, 1]))); max_y = int(np.ceil(np.max(poly_arr[:, 1])))
    min_x = max(0, min_x); max_x = min(W, max_x)
    min_y = max(0, min_y); max_y = min(H, max_y)
    if max_x <= min_x or max_y <= min_y:
        return None, None, None
    sub_img = img[min_y:max_y, min_x:max_x]
    sh, sw = sub_img.shape
    inside_sub = rasterize_polygon(poly_arr - [min_x, min_y], (sh, sw))
    vals = sub_img[inside_sub]
    if vals.size == 0:
        return None, None, None
    thr_param = float(thr_param)
    if mode.lower() == "bnd":
        m = float(np.nanmean(vals)); s = float(np.nanstd(vals))
        thr = float(np.percentile(vals, 90.0)) if (s <= 0) or (not np.isfinite(s)) else m + thr_param * s
    else:
        thr = float(np.percentile(vals, thr_param))
    cand_sub = (sub_img >= thr) & inside_sub
    lab, n = ndi.label(cand_sub)
    if n == 0:
        return thr, None, None
    sizes = ndi.sum(cand_sub, lab, index=np.arange(1, n + 1))
    k = int(np.argmax(sizes)) + 1
    mask_sub = ndi.binary_fill_holes(lab == k)
    contours = shims.find_contours(mask_sub.astype(float), 0.5)
    if not contours:
        return thr, None, None
    polys = []
    for c in contours:
        xy = np.c_[c[:, 1] + min_x, c[:, 0] + min_y]
        x, y = xy[:, 0], xy[:, 1]
        area = 0.5 * np.abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        if area >= float(min_area):
            xy_s = shims.approximate_polygon(xy, tolerance=float(tolerance))
            if len(xy_s) >= 3:
                polys.append((area, xy_s))
    if not polys:
        return thr, None, None
    return thr, None, max(polys, key=lambda t: t[0])[1]
