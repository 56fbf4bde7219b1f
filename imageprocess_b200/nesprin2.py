"""Nesprin2 FRET numeric body for a batch of F (donor, FRET[, acceptor-only]) frames resident in
HBM: reference src/FRET/Nesprin2_FRET_Builder.py:1415-1421,1445-1581 (run_pipeline's per-pair
loop), SURVEY.md 8(a) a4, a8-a14.

Device work per batch (no per-ROI host loop):
  rasterise ROIs + union            ipb_rasterize_rois          (a1)
  saturation-aware backgrounds      ipb_hist_u16 (excl plane)    (a4, a11)  -> Bd, Ba, Bao
  epsilon on the union-masked denominator
      monotone channel              histogram order statistics   (a8)
      spectrally corrected FRET     ipb_fret_pixels -> float32 region statistics -> ipb_eps_from_stat
  saturation / bg / spectral / ratio (+ inverse) / ratio clip
                                    ipb_fret_pixels              (a9-a11)   -> R, Ralt, Dcorr, Acorr
  inner rim 0 < EDT <= rim_px       ipb_region_dilate (ball)     (a12)
  per-ROI statistics on roi & rim   ipb_region_stats (AND plane) (a14)
  optional annulus background       ipb_region_dilate (squares) + medians + ratio re-derivation (a13)
"""
import math

import numpy as np

from . import geometry as geo
from . import ops
from .batch import hist_mode_level
from .ops import (FP_BA, FP_BAO, FP_BD, FP_STRIDE, FRET_CFG, HIST_JOB, PAT_FULL, PAT_MASKED, Q_JOB, Q_OUT,
                  QK_MEDIAN, QK_PCT, REGION, SRC_F32, SRC_RATIO, STAT_JOB, q32_of)


def n2_px_params(p):
    """rim / annulus radii in pixels: Nesprin2_FRET_Builder.py:1386-1390."""
    px_um = float(p["px_um"])
    rim_px = max(1, int(round(float(p["rim_um"]) / px_um)))
    ann_on = bool(p["annulus_on"])
    ann_in = max(1, int(round(float(p["ann_in_um"]) / px_um))) if ann_on else 0
    ann_out = max(ann_in + 1, int(round(float(p["ann_out_um"]) / px_um))) if ann_on else 0
    return rim_px, ann_on, ann_in, ann_out


def n2_fret_cfg(p, n_ch, donor_ch, acc_ch, aonly_ch):
    cfg = np.zeros(1, dtype=FRET_CFG)
    cfg["numer_is_acceptor"] = int(p["ratio_mode"] == "FRET/Donor")
    cfg["clip_neg"] = int(bool(p["clip_neg"]))
    cfg["sat_on"] = int(bool(p["sat_filter_on"]))
    cfg["sat_thr"] = np.float32(p["sat_threshold"])
    cfg["use_spectral"] = int(bool(p["use_spectral"]))
    cfg["alpha"], cfg["beta"] = np.float32(p["alpha"]), np.float32(p["beta"])
    cfg["g_factor"] = np.float32(float(p["g_factor"]))
    cfg["clip_on"] = int(bool(p["clip_ratio_on"]))
    cfg["clip_max"] = np.float32(p["clip_ratio_max"])
    cfg["donor_ch"], cfg["acc_ch"], cfg["n_ch"] = donor_ch, acc_ch, n_ch
    cfg["aonly_ch"] = aonly_ch if (aonly_ch is not None and bool(p["use_spectral"])) else -1
    return cfg


def nesprin2_batch(eng, planes, shape, polys_per_frame, p, donor_ch=0, acc_ch=1, aonly_ch=None):
    """Returns dict(rows_per_frame, eps[F], images (device float32 [4][F][H][W]: R, Ralt, Dcorr,
    Acorr), rim (host bool getter), union).  Frames without ROIs produce no rows (the reference
    skips them, Nesprin2_FRET_Builder.py:1441-1443)."""
    mem = eng.mem
    F, C, H, W = (int(s) for s in shape)
    rim_px, ann_on, ann_in, ann_out = n2_px_params(p)
    scope = p["bg_scope"]
    use_ann = ann_on or scope == "annulus"
    if use_ann:
        # the radii annulus_mask_from_poly really dilates by (Nesprin2_FRET_Builder.py:419-422): with the
        # option off and scope == "annulus" the reference passes 0, 0, which it clamps to 1 and 2
        if not ann_on:
            ann_in, ann_out = 0, 0
        ann_in = max(ann_in, 1)
        ann_out = ann_out if ann_out > ann_in else ann_in + 1
    fd = p["ratio_mode"] == "FRET/Donor"
    spectral = bool(p["use_spectral"])
    clip_neg = bool(p["clip_neg"])
    sat_on = bool(p["sat_filter_on"])
    sat_min = 0
    if sat_on:
        sat_min = int(min(max(math.ceil(float(p["sat_threshold"])), 1), 65536))
    per_ch = bool(p["per_channel_p"])
    p_glob = float(p["percentile"])
    d_p = float(p["donor_p"]) if per_ch else p_glob
    a_p = float(p["fret_p"]) if per_ch else p_glob
    pct = p["bg_mode"] == "percentile"
    wpr = (W + 31) // 32

    # ---- rasterise (ROI masks padded by the annulus radius so the dilations stay in-rect)
    pad = ann_out if use_ann else 0
    specs, owner = [], []
    for f, polys in enumerate(polys_per_frame):
        for i, P in enumerate(polys or ()):
            specs.append(geo.mpl_spec(P, (W, H), frame=f, pad=pad))
            owner.append((f, i + 1))
    NR = len(specs)
    has_rois = np.zeros(F, dtype=bool)
    for f, _ in owner:
        has_rois[f] = True
    rm = eng.rasterize(geo.RULE_MPL, specs, (H, W), F, want_union=True)
    reg = ops.regions_from_masks(rm)
    union = rm.union

    # ---- histograms: backgrounds (+ epsilon when the denominator is a monotone channel)
    masked = scope != "full"
    pat = np.where(masked & has_rois, PAT_MASKED, PAT_FULL)
    jobs = []

    def hjob(ch, pattern, excl_ch):
        j = np.zeros(F, dtype=HIST_JOB)
        j["plane"] = np.arange(F) * C + ch
        j["pattern"] = pattern
        j["mask_frame"] = np.arange(F)
        if sat_on and excl_ch is not None:
            j["excl_plane1"] = np.arange(F) * C + excl_ch + 1
            j["sat_min"] = sat_min
        jobs.append(j)
        return (len(jobs) - 1) * F

    h_d = hjob(donor_ch, pat, acc_ch)
    h_a = hjob(acc_ch, pat, donor_ch)
    have_ao = spectral and aonly_ch is not None
    h_ao = hjob(aonly_ch, pat, None) if have_ao else None
    eps_hist = fd or not spectral                   # denominator = Dbc, or Abc without spectral correction
    den_ch, den_other = (donor_ch, acc_ch) if fd else (acc_ch, donor_ch)
    h_eps = hjob(den_ch, np.where(has_rois, PAT_MASKED, PAT_FULL), den_other) if eps_hist else None
    hres = eng.hist(planes, H, W, np.concatenate(jobs), union=union, union_wpr=wpr)
    qj = []

    def qjob(h0, pp):
        q = np.zeros(F, dtype=Q_JOB)
        q["hist"] = h0 + np.arange(F)
        q["q32"] = q32_of(pp)
        qj.append(q)
        return (len(qj) - 1) * F

    q_d, q_a = qjob(h_d, d_p), qjob(h_a, a_p)
    q_ao = qjob(h_ao, p_glob) if have_ao else None
    q_eps = qjob(h_eps, float(p["eps_percentile"])) if eps_hist else None
    qout = eng.quantiles(hres, np.concatenate(qj))
    fparams = mem.zeros((F, FP_STRIDE), np.float32)
    if pct:
        idx = np.full(len(qj) * F, -1, dtype=np.int32)
        idx[q_d: q_d + F] = np.arange(F) * FP_STRIDE + FP_BD
        idx[q_a: q_a + F] = np.arange(F) * FP_STRIDE + FP_BA
        if have_ao:
            idx[q_ao: q_ao + F] = np.arange(F) * FP_STRIDE + FP_BAO
        eng.scatter_qvalues(qout, idx, fparams)
    elif p["bg_mode"] == "hist-mode":
        hh = hres.hist.host()
        fp = np.zeros((F, FP_STRIDE), dtype=np.float32)
        for f in range(F):
            for h0, slot, pp in ((h_d, FP_BD, d_p), (h_a, FP_BA, a_p)) + (((h_ao, FP_BAO, p_glob),) if have_ao else ()):
                lvl = hist_mode_level(hh[h0 + f], pp)
                fp[f, slot] = 0.0 if lvl is None else lvl
        fparams = mem.from_host(fp)
    cfg = n2_fret_cfg(p, C, donor_ch, acc_ch, aonly_ch)
    images = mem.empty((4, F, H, W), np.float32)
    plane_bytes = 4 * F * H * W
    img_ptr = lambda k: images.ptr + k * plane_bytes
    IMG_R, IMG_RALT, IMG_D, IMG_A = 0, 1, 2, 3

    def run_pixels(only_corr):
        eng.call("ipb_fret_pixels", planes.ptr, F, H, W, cfg.ctypes.data, fparams.ptr, None, wpr, None,
                 None if only_corr else img_ptr(IMG_R), None if only_corr else img_ptr(IMG_RALT), None,
                 img_ptr(IMG_D), img_ptr(IMG_A), None, None, 0, mem.stream)

    # ---- epsilon
    ureg = np.zeros(F, dtype=REGION)                       # the union plane of every frame as a region
    ureg["mask_off"] = np.arange(F, dtype=np.int64) * H * wpr
    ureg["w"], ureg["h"], ureg["wpr"], ureg["frame"] = W, H, wpr, np.arange(F)
    if eps_hist:
        eng.call("ipb_fret_eps", qout.ptr + Q_OUT.itemsize * q_eps, F, FP_BD if fd else FP_BA, int(clip_neg), 5.0,
                 fparams.ptr, mem.stream)
    else:
        run_pixels(True)                                   # Dcorr / Acorr with the backgrounds only
        ej = np.zeros(F, dtype=STAT_JOB)
        ej["region"], ej["src"], ej["n_views"] = np.arange(F), SRC_F32, 1
        ej["plane"] = IMG_A * F + np.arange(F)             # denominator = corrected FRET channel
        ej["bidx"] = -1
        ej["qkind"] = (QK_PCT, 0, 0)
        ej["q32"] = (q32_of(float(p["eps_percentile"])), 0.0, 0.0)
        ej["out"][:, 0] = np.arange(F)
        so = eng.region_stats(ureg, ej, union, H, W, images=images)
        row = np.where(has_rois, np.arange(F), -1).astype(np.int32)
        d_row = mem.from_host(row)
        eng.call("ipb_eps_from_stat", so.ptr, d_row.ptr, F, 5.0, fparams.ptr, mem.stream)
    run_pixels(False)

    # ---- inner rim of the union, per frame
    gm, R = ops.ball_gmax(ops.rim_d2max(rim_px))
    rim = eng.region_dilate(ureg, union, gm, R, invert=True, and_pool=union)

    # ---- optional annulus background per ROI -> per-ROI ratio re-derivation parameters
    ratio_params = None
    ring = None
    if use_ann and NR:
        inner_px, outer_px = ann_in, ann_out          # already clamped; the stored rects are padded by outer_px
        gi, Ri = ops.square_gmax(inner_px)
        go, Ro = ops.square_gmax(outer_px)
        inner = eng.region_dilate(reg, rm.pool, gi, Ri)
        ring = eng.region_dilate(reg, rm.pool, go, Ro, andnot_pool=inner)
        frame_of = reg["frame"]
        mj = np.zeros((NR, 2), dtype=STAT_JOB)
        for k, img in enumerate((IMG_D, IMG_A)):
            s = mj[:, k]
            s["region"], s["src"], s["n_views"] = np.arange(NR), SRC_F32, 1
            s["plane"] = img * F + frame_of
            s["bidx"] = -1
            s["qkind"] = (QK_MEDIAN, 0, 0)
            s["out"][:, 0] = np.arange(NR) * 2 + k
        med = eng.region_stats(reg, mj.reshape(-1), ring, H, W, images=images).host().reshape(NR, 2)
        bg_D = np.where(med["n"][:, 0] > 0, med["q"][:, 0, 0], np.float32(0)).astype(np.float32)
        bg_A = np.where(med["n"][:, 1] > 0, med["q"][:, 1, 0], np.float32(0)).astype(np.float32)
        eps_h = fparams.host()[:, 2]
        # per ROI: {bg_numer, bg_denom, eps, clip_neg, clip_on, clip_max} for main and inverse ratio
        rp = np.zeros((NR, 2, 6), dtype=np.float32)
        bn, bd = (bg_A, bg_D) if fd else (bg_D, bg_A)
        rp[:, 0, 0], rp[:, 0, 1] = bn, bd
        rp[:, 1, 0], rp[:, 1, 1] = bd, bn
        rp[:, :, 2] = eps_h[frame_of][:, None]
        rp[:, :, 3] = float(clip_neg)
        rp[:, :, 4] = float(bool(p["clip_ratio_on"]))
        rp[:, :, 5] = np.float32(p["clip_ratio_max"])
        ratio_params = mem.from_host(rp.reshape(-1))

    # ---- per-ROI statistics on roi & rim
    rows_per_frame = [[] for _ in range(F)]
    if NR:
        reg_and = reg.copy()
        reg_and["use_and"], reg_and["and_plane"] = 1, reg["frame"]
        frame_of = reg["frame"]
        sj = np.zeros((NR, 4), dtype=STAT_JOB)
        num_img, den_img = (IMG_A, IMG_D) if fd else (IMG_D, IMG_A)
        for k in range(4):
            s = sj[:, k]
            s["region"], s["n_views"] = np.arange(NR), 1
            s["bidx"] = -1
            s["out"][:, 0] = np.arange(NR) * 4 + k
        for k, img in ((0, IMG_R), (1, IMG_RALT)):
            s = sj[:, k]
            if ratio_params is None:
                s["src"], s["plane"] = SRC_F32, img * F + frame_of
            else:                                       # ratio re-derived from the corrected channels
                s["src"] = SRC_RATIO
                s["plane"] = (num_img if k == 0 else den_img) * F + frame_of
                s["clip_neg"][:, 0] = (den_img if k == 0 else num_img) * F + frame_of     # second image
                s["bidx"][:, 0] = (np.arange(NR) * 2 + k) * 6
        sj[:, 0]["qkind"] = (QK_PCT, QK_MEDIAN, QK_PCT)
        sj[:, 0]["q32"] = (q32_of(5), 0.0, q32_of(95))
        for k, img in ((2, IMG_D), (3, IMG_A)):
            sj[:, k]["src"], sj[:, k]["plane"] = SRC_F32, img * F + frame_of
        so = eng.region_stats(reg_and, sj.reshape(-1), rm.pool, H, W, images=images, bvals=ratio_params,
                              and_bits=rim, and_wpr=wpr).host().reshape(NR, 4)
        eps_h = fparams.host()[:, 2]
        for r, (f, roi) in enumerate(owner):
            o = so[r, 0]
            n = int(o["n"])
            mean_main = float(np.float32(o["sum"] / n)) if n else math.nan
            na = int(so[r, 1]["n"])
            mean_alt = float(np.float32(so[r, 1]["sum"] / na)) if na else math.nan
            row = {"roi": roi, "area_px": int(o["area"]), "ratio_FoverD_mean": mean_main,
                   "ratio_DoverF_mean": mean_alt, "eps": float(eps_h[f])}
            if n == 0:
                row.update({k: math.nan for k in ("ratio_mean", "ratio_median", "ratio_std", "ratio_p5",
                                                  "ratio_p95", "donor_mean", "fret_mean")})
            else:
                nd, nf = int(so[r, 2]["n"]), int(so[r, 3]["n"])
                row.update({"ratio_mean": mean_main, "ratio_median": float(o["q"][1]),
                            "ratio_std": float(np.float32(math.sqrt(max(float(o["ssd"]) / n, 0.0)))),
                            "ratio_p5": float(o["q"][0]), "ratio_p95": float(o["q"][2]),
                            "donor_mean": float(np.float32(so[r, 2]["sum"] / nd)) if nd else math.nan,
                            "fret_mean": float(np.float32(so[r, 3]["sum"] / nf)) if nf else math.nan})
            rows_per_frame[f].append(row)
    return {"rows_per_frame": rows_per_frame, "eps": fparams.host()[:, 2].copy(), "fparams": fparams.host(),
            "images": images, "rim": rim, "union": rm, "has_rois": has_rois, "masks": rm,
            "ring": ring, "regions": reg}


def bits_to_bool(words, H, W):
    """[..., H, wpr] uint32 bit planes -> bool [..., H, W]."""
    b = np.unpackbits(np.ascontiguousarray(words).view(np.uint8), axis=-1, bitorder="little")
    return b[..., :W].astype(bool)
