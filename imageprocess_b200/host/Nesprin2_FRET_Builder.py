"""Mirror of src/FRET/Nesprin2_FRET_Builder.py's run_pipeline:

    run_pipeline(p) -> None (writes files)                        Nesprin2_FRET_Builder.py:1331-1736

`p` is the dict gui_get_params returns (keys at :674-698).  The per-pair loop body (:1407-1582)
runs on the device through imageprocess_b200.nesprin2.nesprin2_batch.  Rows carry the
reference's columns (:1537-1581).  Two reference defects on this path are handled explicitly
(SURVEY.md 8(a) "Reference defects"): #2 (ratio_mode compared with "DoverF") is reproduced, so
ratio_FoverD_mean always holds the primary ratio's mean; #1 ("time" holds the translation
function when timelapse is on, which breaks the reference's own save_xls) is NOT reproduced:
"time" holds the t-code.
"""
import os

import numpy as np

from .. import nesprin2
from . import common
from ._fretnames import build_pairs_by_channel, load_roi_polys  # noqa: F401
from ._fretnames import parse_tokens_delimited as parse_tokens  # noqa: F401  (this script's own grammar, :292-307)
from .common import ensure_dir, list_tifs

DEFAULT_P = {
    "timelapse": False, "donor_ch": 2, "fret_ch": 3, "intensity_ch": 1, "ratio_mode": "FRET/Donor",
    "bg_scope": "full", "bg_mode": "percentile", "percentile": 1.0, "per_channel_p": False, "donor_p": 1.0,
    "fret_p": 1.0, "clip_neg": True, "eps_percentile": 1.0, "px_um": 0.223, "rim_um": 1.12, "annulus_on": False,
    "ann_in_um": 1.2, "ann_out_um": 2.5, "use_spectral": False, "alpha": 0.0, "beta": 0.0, "g_factor": 1.0,
    "aonly_ch": None, "out_xls": True, "out_tif": True, "out_png": False, "subset_on": False, "subset_stage": None,
    "subset_time": None, "sat_filter_on": True, "sat_threshold": 65535.0, "clip_ratio_on": True,
    "clip_ratio_max": 20.0, "out_root": ""}                            # Nesprin2_FRET_Builder.py:674-698


def _engine():
    import imageprocess_b200 as ipb
    return ipb.engine()


def swap_ch(path, ch_from, ch_to):
    """Sibling file of another channel: replaces the trailing channel token."""
    d, b = os.path.split(path)
    stem, ext = os.path.splitext(b)
    import re
    new = re.sub(rf"([_-](?:ch|c)?){int(ch_from)}$", rf"\g<1>{int(ch_to)}", stem, flags=re.IGNORECASE)
    return os.path.join(d, new + ext)


def run_pipeline(p, eng=None, log=print, frames_per_batch=16):
    eng = eng or _engine()
    p = {**DEFAULT_P, **p}
    img_dir, roi_dir = p["img_dir"], p["roi_dir"]
    out_root = (p.get("out_root") or "").strip() or os.path.join(img_dir, "RES")
    timelapse = bool(p["timelapse"])
    donor_ch, fret_ch = int(p["donor_ch"]), int(p["fret_ch"])
    pairs, _ = build_pairs_by_channel(list_tifs(img_dir), timelapse, donor_ch, fret_ch, parse=parse_tokens)
    log(f"[info] pairs to process: {len(pairs)}")
    if not pairs:
        log("no matching (donor, fret) channel pairs")
        return []
    if p["subset_on"] and p["subset_stage"] is not None:
        s_code = f"S{int(p['subset_stage']):02d}"
        if timelapse and p.get("subset_time") is not None:
            t_code = f"t{int(p['subset_time']):02d}"
            pairs = [pp for pp in pairs if pp[0] == (s_code, t_code)]
        else:
            pairs = [pp for pp in pairs if pp[0][0] == s_code]
        if not pairs:
            log("[subset] no pair matches")
            return []
    res_root = ensure_dir(out_root)
    tif_full = ensure_dir(os.path.join(res_root, "TIF", "ratio32_full")) if p["out_tif"] else None
    tif_rim = ensure_dir(os.path.join(res_root, "TIF", "ratio32_rim")) if p["out_tif"] else None
    fd = p["ratio_mode"] == "FRET/Donor"
    suffix = "FoverD" if fd else "DoverF"
    aonly = p.get("aonly_ch")
    aonly = int(aonly) if aonly not in (None, "", "None") else None
    per_ch = bool(p["per_channel_p"])
    d_p = float(p["donor_p"]) if per_ch else float(p["percentile"])
    a_p = float(p["fret_p"]) if per_ch else float(p["percentile"])

    items = []
    for key, dpath, apath in pairs:
        s, t_code = key
        tag = f"{s}_{t_code}" if (timelapse and t_code is not None) else s
        log(f"[process] {tag} ...")
        planes = [common.as_u16_plane(common.read_image_raw(dpath), f"{tag} donor"),
                  common.as_u16_plane(common.read_image_raw(apath), f"{tag} fret")]
        has_ao = False
        if aonly is not None:
            cand = swap_ch(dpath, donor_ch, aonly)
            if not os.path.exists(cand):
                cand = swap_ch(apath, fret_ch, aonly)
            if os.path.exists(cand):
                planes.append(common.as_u16_plane(common.read_image_raw(cand), f"{tag} acceptor-only"))
                has_ao = True
        polys = load_roi_polys(roi_dir, s, t_code, timelapse)
        if not polys:
            log(f"[warn] {tag}: no ROI - skipped")
            continue
        items.append((s, t_code, tag, np.stack(planes), polys, has_ao))
    rows_all = []
    groups = {}
    for it in items:
        groups.setdefault((it[3].shape, it[5]), []).append(it)
    for (shape, has_ao), group in groups.items():
        for b0 in range(0, len(group), frames_per_batch):
            chunk = group[b0: b0 + frames_per_batch]
            planes = np.stack([it[3] for it in chunk])
            out = nesprin2.nesprin2_batch(eng, eng.mem.from_host(planes), planes.shape, [it[4] for it in chunk], p,
                                          donor_ch=0, acc_ch=1, aonly_ch=2 if has_ao else None)
            if p["out_tif"]:
                imgs = out["images"].host()
                F, _, H, W = planes.shape
                rim = nesprin2.bits_to_bool(out["rim"].host().reshape(F, H, (W + 31) // 32), H, W)
            for f, (s, t_code, tag, _, polys, _) in enumerate(chunk):
                if p["out_tif"]:
                    common.write_tiff(os.path.join(tif_full, f"{tag}_ratio_{suffix}.tif"), imgs[0, f])
                    common.write_tiff(os.path.join(tif_rim, f"{tag}_ratio_{suffix}_rim.tif"),
                                      np.where(rim[f], imgs[0, f], np.float32(np.nan)).astype(np.float32))
                for r in out["rows_per_frame"][f]:
                    row = {"stage": s, "time": (t_code if timelapse else None)}
                    row.update(r)
                    row.update({"p": float(p["percentile"]), "donor_p": d_p, "fret_p": a_p, "ratio_mode": p["ratio_mode"],
                                "bg_scope": p["bg_scope"], "bg_mode": p["bg_mode"], "clip_neg": bool(p["clip_neg"]),
                                "sat_filter_on": bool(p["sat_filter_on"]), "sat_threshold": float(p["sat_threshold"]),
                                "clip_ratio_on": bool(p["clip_ratio_on"]), "clip_ratio_max": float(p["clip_ratio_max"])})
                    rows_all.append(row)
    if p.get("out_png"):
        log("[SKIP-PNG] figure rendering is host matplotlib code outside the device path")
    if p["out_xls"]:
        save_xls(rows_all, ensure_dir(os.path.join(res_root, "xls")), timelapse, log=log)
    return rows_all


KEEP_COLS = ["stage", "time", "roi", "area_px", "ratio_mode", "ratio_mean", "ratio_median", "ratio_std", "ratio_p5",
             "ratio_p95", "ratio_FoverD_mean", "ratio_DoverF_mean", "donor_mean", "fret_mean", "eps", "p", "donor_p",
             "fret_p", "bg_scope", "bg_mode", "clip_neg", "sat_filter_on", "sat_threshold", "clip_ratio_on",
             "clip_ratio_max"]


def per_roi_frame(rows_all, timelapse):
    """The table save_xls writes (Nesprin2_FRET_Builder.py:1292-1306): the 25 kept columns in the
    reference's order, then stage_idx, time_idx, roi_lab."""
    import re
    import pandas as pd
    df = pd.DataFrame(rows_all)
    if df.empty:
        return None
    df = df[[c for c in KEEP_COLS if c in df.columns]].copy()
    df["stage_idx"] = [int(re.search(r"S(\d+)", s).group(1)) for s in df["stage"]]
    df["time_idx"] = [int(re.search(r"t(\d+)", tt).group(1)) for tt in df["time"]] if timelapse else 0
    df["roi_lab"] = ["s%dc%d" % (si, r) for si, r in zip(df["stage_idx"], df["roi"])]
    return df


def save_xls(rows_all, xls_dir, timelapse, log=print):
    """nesprin2_fret_perROI.csv first, then the .xlsx with the two time matrices when openpyxl is
    installed (Nesprin2_FRET_Builder.py:1287-1326)."""
    df = per_roi_frame(rows_all, timelapse)
    if df is None:
        log("[warn] no ROI: no metric table")
        return None
    df.to_csv(os.path.join(xls_dir, "nesprin2_fret_perROI.csv"), index=False)
    log("[saved] xls/nesprin2_fret_perROI.csv (CSV)")
    try:
        import openpyxl  # noqa: F401
        import pandas as pd
        with pd.ExcelWriter(os.path.join(xls_dir, "nesprin2_fret_perROI.xlsx"), engine="openpyxl") as w:
            df.to_excel(w, index=False, sheet_name="per_ROI")
            for what in ("mean", "median"):
                df.pivot(index="time_idx", columns="roi_lab", values=f"ratio_{what}").sort_index() \
                  .to_excel(w, sheet_name=f"ratio_{what}_matrix")
        log("[saved] xls/nesprin2_fret_perROI.xlsx (XLSX)")
    except ModuleNotFoundError:
        log("[warn] openpyxl not installed: XLSX skipped, CSV only")
    return df
