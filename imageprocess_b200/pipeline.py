"""Batched frame pipelines: the hot path of SURVEY.md section 8 for F frames at a time.

Each function takes device-resident uint16 planes ``[F][C][H][W]`` plus the per-frame ROI
polygons and the reference's own parameter dict, launches the kernels on the current stream
without any host synchronisation in between, and reads back only the small per-ROI tables
at the end.  Row dicts use the reference's column names so the host mirrors
(Fluor_INT.py, fret_ratio_builder.py, ...) can hand them to the unchanged pandas writers.
"""
import math

import numpy as np

from . import geometry as geo
from . import ops
from .ops import (FP_BA, FP_BD, FP_STRIDE, HIST_JOB, PAT_FULL, PAT_MASKED, PAT_MASKED_STRIDE,
                  PAT_STRIDE1D, PAT_STRIDE2D, Q_JOB, QK_MEDIAN, QK_PCT, SRC_F32, SRC_U16, STAT_JOB,
                  q32_of)


def _f32(x):
    return float(np.float32(x))


def _stat_common(o):
    """mean / std from the exact accumulators, rounded to float32 like the reference's
    np.mean / np.std of a float32 selection."""
    n = int(o["n"])
    if n == 0:
        return n, math.nan, math.nan
    mean = float(o["sum"]) / n
    return n, _f32(mean), _f32(math.sqrt(max(float(o["ssd"]) / n, 0.0)))


def hist_mode_level(counts, p):
    """'hist-mode' background (reference Fluor_INT.py:474-483) from an exact integer
    histogram: np.histogram of the distinct values weighted by their counts places every
    value in the same bin as np.histogram of the full sample (same float32 edges)."""
    vals = np.flatnonzero(counts)
    if vals.size == 0:
        return 0.0
    w = counts[vals].astype(np.float64)
    v32 = vals.astype(np.float32)
    hist, bins = np.histogram(v32, bins=2048, weights=w)
    if hist.sum() <= 0:
        return None
    cdf = np.cumsum(hist).astype(float)
    cdf /= cdf[-1]
    idx = int(np.searchsorted(cdf, float(p) / 100.0, side="left"))
    if idx >= len(bins) - 1:
        return float(bins[-1])
    return float(0.5 * (bins[idx] + bins[idx + 1]))


# ====================================================================== ROI intensity
def intensity_batch(eng, planes, shape, polys_per_frame, task, ch_names=None):
    """Fluor_INT per-ROI quantification for F frames (reference
    src/INT/Fluor_INT.py:839-870: bg scope -> bg_correct per channel ->
    quantify_per_roi_multi).

    planes            DevBuf uint16 [F][C][H][W]
    polys_per_frame   list (len F) of lists of (V,2) float polygons (>= 3 vertices)
    task              reference task dict keys: bg_scope, bg_mode, percentile, per_channel_p,
                      ch_p_map, clip_neg, bg_stride
    ch_names          channel numbers for column names (default 1..C)
    Returns (rows_per_frame, bg_used_per_frame, masks)
    """
    F, C, H, W = shape
    ch_names = list(ch_names) if ch_names is not None else list(range(1, C + 1))
    specs, owner = [], []
    for f, polys in enumerate(polys_per_frame):
        for i, P in enumerate(polys):
            specs.append(geo.mpl_spec(P, (W, H), frame=f))
            owner.append((f, i + 1))
    scope_union = task["bg_scope"] == "roi_union"
    rm = eng.rasterize(geo.RULE_MPL, specs, (H, W), F, want_union=True)
    stride = int(task["bg_stride"]) if task.get("bg_stride") else 1
    p_glob = float(task["percentile"])
    pcs = [float(task["ch_p_map"].get(ch, p_glob)) if task.get("per_channel_p") else p_glob
           for ch in ch_names]
    jobs = np.zeros(F * C, dtype=HIST_JOB)
    for f in range(F):
        masked = scope_union and len(polys_per_frame[f]) > 0
        for c in range(C):
            j = jobs[f * C + c]
            j["plane"] = f * C + c
            j["mask_frame"] = f
            j["k"] = stride
            if masked:
                j["pattern"] = PAT_MASKED_STRIDE if stride > 1 else PAT_MASKED
            else:
                j["pattern"] = PAT_STRIDE1D if stride > 1 else PAT_FULL
    hres = eng.hist(planes, H, W, jobs, rm.union, rm.union_wpr)
    bvals = eng.mem.zeros(F * C, np.float32)
    mode = task["bg_mode"]
    if mode == "percentile":
        qj = np.zeros(F * C, dtype=Q_JOB)
        qj["hist"] = np.arange(F * C)
        qj["q32"] = np.tile(np.array([q32_of(p) for p in pcs], dtype=np.float32), F)
        qout = eng.quantiles(hres, qj)
        eng.scatter_qvalues(qout, np.arange(F * C), bvals)
    elif mode == "hist-mode":
        hh = hres.hist.host()
        b = np.zeros(F * C, dtype=np.float32)
        for idx in range(F * C):
            lvl = hist_mode_level(hh[idx], pcs[idx % C])
            b[idx] = 0.0 if lvl is None else lvl
        bvals = eng.mem.from_host(b)
    # any other mode: B = 0.0 (reference bg_value's final else)
    reg = ops.regions_from_masks(rm)
    sj = np.zeros(len(specs) * C, dtype=STAT_JOB)
    for r, (f, _) in enumerate(owner):
        for c in range(C):
            j = sj[r * C + c]
            j["region"] = r
            j["src"] = SRC_U16
            j["plane"] = f * C + c
            j["bidx"] = f * C + c
            j["clip_neg"] = int(bool(task["clip_neg"]))
            j["qkind"] = (QK_PCT, QK_MEDIAN, QK_PCT)
            j["q32"] = (q32_of(5), 0.0, q32_of(95))
    sout = eng.region_stats(reg, sj, rm.pool, H, W, planes=planes, bvals=bvals)
    so = sout.host()
    area = rm.area.host()
    bh = bvals.host()
    rows_per_frame = [[] for _ in range(F)]
    for r, (f, roi_i) in enumerate(owner):
        row = {"roi": roi_i, "area_px": int(area[r])}
        for c, ch in enumerate(ch_names):
            o = so[r * C + c]
            n, mean, std = _stat_common(o)
            if n == 0:
                st = dict(mean=math.nan, median=math.nan, std=math.nan, p5=math.nan, p95=math.nan,
                          vmin=math.nan, vmax=math.nan, vsum=math.nan, npx=0)
            else:
                st = dict(mean=mean, median=float(o["q"][1]), std=std, p5=float(o["q"][0]),
                          p95=float(o["q"][2]), vmin=float(o["vmin"]), vmax=float(o["vmax"]),
                          vsum=_f32(o["sum"]), npx=n)
            for k, v in st.items():
                row[f"ch{ch}_{k}"] = v
        rows_per_frame[f].append(row)
    bg_used = [{ch: {"bg": float(bh[f * C + c]), "p": float(pcs[c])} for c, ch in enumerate(ch_names)}
               for f in range(F)]
    return rows_per_frame, bg_used, rm


# ====================================================================== general FRET
def fret_cfg(p, n_ch=2, donor_ch=0, acc_ch=1):
    cfg = np.zeros(1, dtype=ops.FRET_CFG)
    cfg["numer_is_acceptor"] = int(p["ratio_mode"] == "FRET/Donor")
    cfg["clip_neg"] = int(bool(p["clip_neg"]))
    cfg["g_factor"] = 1.0
    cfg["donor_ch"], cfg["acc_ch"], cfg["aonly_ch"], cfg["n_ch"] = donor_ch, acc_ch, -1, n_ch
    return cfg


def fret_batch(eng, planes, shape, polys_per_frame, p, donor_ch=0, acc_ch=1, want_roi_image=False):
    """fret_ratio_builder.process_one_stage numeric body for F (donor, acceptor) pairs
    (reference src/FRET/fret_ratio_builder.py:454-474, 493-507).

    Returns dict: rows_per_frame, fparams (host [F][4] = Bd, Ba, eps, -), R (DevBuf f32
    [F][H][W]), R_roi (DevBuf or None), masks."""
    F, C, H, W = shape
    specs, owner = [], []
    for f, polys in enumerate(polys_per_frame):
        for i, P in enumerate(polys or []):
            specs.append(geo.mpl_spec(P, (W, H), frame=f))
            owner.append((f, i + 1))
    rm = eng.rasterize(geo.RULE_MPL, specs, (H, W), F, want_union=True)
    per_ch = bool(p["per_channel_p"])
    d_p = float(p["donor_p"]) if per_ch else float(p["percentile"])
    a_p = float(p["fret_p"]) if per_ch else float(p["percentile"])
    scope_union = p["bg_scope"] == "roi_union"
    jobs = np.zeros(2 * F, dtype=HIST_JOB)
    for f in range(F):
        masked = scope_union and bool(polys_per_frame[f])
        for s, ch in enumerate((donor_ch, acc_ch)):
            j = jobs[2 * f + s]
            j["plane"] = f * C + ch
            j["mask_frame"] = f
            j["pattern"] = PAT_MASKED if masked else PAT_FULL
    hres = eng.hist(planes, H, W, jobs, rm.union, rm.union_wpr)
    fparams = eng.mem.zeros((F, FP_STRIDE), np.float32)
    numer_is_acc = p["ratio_mode"] == "FRET/Donor"
    den_slot = FP_BD if numer_is_acc else FP_BA          # denominator: donor for F/D, acceptor for D/F
    mode = p["bg_mode"]
    qj = np.zeros(2 * F, dtype=Q_JOB)
    qe = np.zeros(F, dtype=Q_JOB)
    for f in range(F):
        qj[2 * f + 0] = (2 * f + 0, q32_of(d_p), (0, 0))
        qj[2 * f + 1] = (2 * f + 1, q32_of(a_p), (0, 0))
        qe[f] = (2 * f + (0 if numer_is_acc else 1), q32_of(p["eps_percentile"]), (0, 0))
    if mode == "percentile":
        qout = eng.quantiles(hres, qj)
        dst = np.empty(2 * F, dtype=np.int32)
        dst[0::2] = np.arange(F) * FP_STRIDE + FP_BD
        dst[1::2] = np.arange(F) * FP_STRIDE + FP_BA
        eng.scatter_qvalues(qout, dst, fparams)
    elif mode == "hist-mode":
        hh = hres.hist.host()
        fp = np.zeros((F, FP_STRIDE), dtype=np.float32)
        for f in range(F):
            for s, (slot, pp) in enumerate(((FP_BD, d_p), (FP_BA, a_p))):
                lvl = hist_mode_level(hh[2 * f + s], pp)
                fp[f, slot] = 0.0 if lvl is None else lvl
        fparams = eng.mem.from_host(fp)
    # eps: percentile of the bg-corrected denominator over the same scope
    qeps = eng.quantiles(hres, qe)
    eng.fret_eps(qeps, F, den_slot, bool(p["clip_neg"]), fparams)
    cfg = fret_cfg(p, C, donor_ch, acc_ch)
    R = eng.mem.empty((F, H, W), np.float32)
    Rroi = eng.mem.empty((F, H, W), np.float32) if want_roi_image else None
    eng.fret_pixels(planes, F, H, W, cfg, fparams, rm.union, rm.union_wpr, R=R, Rroi=Rroi)
    rows_per_frame = [[] for _ in range(F)]
    if specs:
        reg = ops.regions_from_masks(rm)
        sj = np.zeros(3 * len(specs), dtype=STAT_JOB)
        for r, (f, _) in enumerate(owner):
            sj[3 * r + 0] = (r, SRC_F32, f, -1, 0, (QK_PCT, QK_MEDIAN, QK_PCT), (q32_of(5), 0.0, q32_of(95)), 0)
            sj[3 * r + 1] = (r, SRC_U16, f * C + donor_ch, f * FP_STRIDE + FP_BD, int(bool(p["clip_neg"])),
                             (0, QK_MEDIAN, 0), (0.0, 0.0, 0.0), 0)
            sj[3 * r + 2] = (r, SRC_U16, f * C + acc_ch, f * FP_STRIDE + FP_BA, int(bool(p["clip_neg"])),
                             (0, QK_MEDIAN, 0), (0.0, 0.0, 0.0), 0)
        sout = eng.region_stats(reg, sj, rm.pool, H, W, planes=planes, images=R, bvals=fparams)
        so = sout.host()
        area = rm.area.host()
        for r, (f, roi_i) in enumerate(owner):
            o = so[3 * r]
            n, mean, std = _stat_common(o)
            row = {"roi": roi_i, "area_px": int(area[r])}
            if n == 0:
                row.update({f"ratio_{k}": math.nan for k in ("mean", "median", "std", "p5", "p95")})
            else:
                row.update({"ratio_mean": mean, "ratio_median": float(o["q"][1]), "ratio_std": std,
                            "ratio_p5": float(o["q"][0]), "ratio_p95": float(o["q"][2])})
            for name, oo in (("donor", so[3 * r + 1]), ("yfret", so[3 * r + 2])):
                nn, mm, _ = _stat_common(oo)
                row[f"{name}_mean"] = mm if nn else math.nan
                row[f"{name}_median"] = float(oo["q"][1]) if nn else math.nan
            rows_per_frame[f].append(row)
    return {"rows_per_frame": rows_per_frame, "fparams": fparams, "R": R, "R_roi": Rroi, "masks": rm}


# ====================================================================== focal adhesions
FA_CATS = ("OK", "Large", "Small")


def fa_um_to_px_config(params, px_size):
    """FA_Analyzer.py:527-535."""
    return {"alpha": params["alpha"], "min_px": params["min_area_um"] / (px_size ** 2),
            "max_px": params["max_area_um"] / (px_size ** 2),
            "close_radius": params["close_radius"], "subtract_bg": params.get("subtract_bg", True)}


def fa_batch(eng, planes, shape, polys_per_frame, params, px_size, channel=0, save_ok_only=True,
             want_labels=False, config=None):
    """FA_Analyzer batch body for F frames (reference src/INT/FA_Analyzer.py:984-1039):
    global stats -> per-ROI crop + skimage mask -> analyze_fa_crop -> CSV rows.

    Returns dict(rows_per_frame, stats (host [F][4] = mean, std, bg, thr), result (FaResult),
    items_per_crop, d2h_bytes)."""
    F, C, H, W = shape
    config = config or fa_um_to_px_config(params, px_size)
    specs, owner, rects = [], [], []
    for f, polys in enumerate(polys_per_frame):
        for i, P in enumerate(polys or []):
            spec, rect = geo.fa_spec(P, (H, W), frame=f)
            if spec is None:
                continue                      # empty crop: reference returns empty results
            specs.append(spec)
            owner.append((f, i + 1))
            rects.append(rect)
    rm = eng.rasterize(geo.RULE_SK, specs, (H, W), F, want_union=False)
    # global stats: exact moments of the whole plane + percentile(img[::10, ::10], 1.0)
    jobs = np.zeros(F, dtype=HIST_JOB)
    jobs["plane"] = np.arange(F) * C + channel
    jobs["pattern"] = PAT_STRIDE2D
    jobs["k"] = 10
    jobs["moments"] = 1
    hres = eng.hist(planes, H, W, jobs)
    qj = np.zeros(F, dtype=Q_JOB)
    qj["hist"] = np.arange(F)
    qj["q32"] = q32_of(1.0)
    qout = eng.quantiles(hres, qj)
    fa_params = eng.mem.empty((F, 4), np.float32)
    eng.fa_params(hres, np.arange(F), qout, F, H * W, config["alpha"], fa_params)
    rows_per_frame = [[] for _ in range(F)]
    if not specs:
        return {"rows_per_frame": rows_per_frame, "stats": fa_params.host(), "result": None,
                "items_per_crop": [], "d2h_bytes": 16 * F}
    crops, total_px, total_rows = ops.crops_from_masks(rm, np.array([f * C + channel for f, _ in owner]))
    res = eng.fa_segment(rm, crops, total_px, total_rows, planes, H, W, fa_params,
                         config["min_px"] if config["min_px"] > 0 else 0.0,
                         int(config["close_radius"]) if config["close_radius"] > 0 else 0,
                         want_labels=want_labels)
    comp_off = res.comp_off.host()
    total = int(comp_off[-1])
    if total > res.cap:
        raise RuntimeError("fa_segment: component table overflow")
    comps = res.comps.host()[:total]
    stats = fa_params.host()
    items_per_crop = fa_items(comps, comp_off, owner, stats, config)
    for k, (f, cell_id) in enumerate(owner):
        th_val = np.float32(stats[f, 3])
        for cat in FA_CATS:
            if save_ok_only and cat != "OK":
                continue
            for it in items_per_crop[k][cat]:
                rows_per_frame[f].append({
                    "Cell_ID": cell_id, "Category": cat, "Area_px": it["area"],
                    "Area_um2": it["area"] * (px_size ** 2),
                    "Mean_Intensity_Raw": it["mean_int_raw"], "Mean_Intensity_Corr": it["mean_int_corr"],
                    "Int_Density_Raw": it["int_den_raw"], "Int_Density_Corr": it["int_den_corr"],
                    "Background_Level": it["bg_level"], "Used_Alpha": params["alpha"],
                    "Global_Threshold": th_val, "Min_Area_Setting": params["min_area_um"],
                    "Max_Area_Setting": params["max_area_um"],
                    "Close_Radius_Setting": params["close_radius"],
                    "Subtract_BG_Setting": params.get("subtract_bg", True)})
    return {"rows_per_frame": rows_per_frame, "stats": stats, "result": res,
            "items_per_crop": items_per_crop, "owner": owner, "rects": rects,
            "d2h_bytes": int(comps.nbytes + comp_off.nbytes + stats.nbytes)}


def fa_items(comps, comp_off, owner, stats, config):
    """Per-adhesion dicts in the reference's format (FA_Analyzer.py:166-193) from the exact
    integer component table.  dtypes follow the reference: area np.float64, mean np.float32,
    integrated densities float64, centroid float64 (row, col) in crop coordinates."""
    out = []
    min_px, max_px = config["min_px"], config["max_px"]
    subtract_bg = config.get("subtract_bg", True)
    for k, (f, _) in enumerate(owner):
        res = {"OK": [], "Large": [], "Small": []}
        bg_val = np.float32(stats[f, 2])
        seg = comps[comp_off[k]: comp_off[k + 1]]
        for lab, c in enumerate(seg, 1):
            area = np.float64(c["area"])
            mean_raw = np.float32(float(c["sum_i"]) / float(c["area"]))
            category = "OK"
            if area < min_px:
                category = "Small"
            elif area > max_px:
                category = "Large"
            mean_corr = max(0, mean_raw - bg_val) if subtract_bg else mean_raw
            res[category].append({
                "label": lab, "area": area, "contour": None,
                "centroid": (float(c["sum_y"]) / float(c["area"]), float(c["sum_x"]) / float(c["area"])),
                "mean_int_raw": mean_raw, "mean_int_corr": mean_corr,
                "int_den_raw": mean_raw * area, "int_den_corr": mean_corr * area,
                "bg_level": bg_val})
        out.append(res)
    return out
