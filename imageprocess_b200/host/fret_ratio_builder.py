"""Mirror of src/FRET/fret_ratio_builder.py's stage worker:

    process_one_stage(stage_key, pairs_for_stage, p, paths) -> (stage_key, rows, logs)   :429-552

All (stage, time) pairs of the stage that share an image shape go through ONE
batch.FrameBatchJob("fret") (rasterisation, backgrounds, epsilon, ratio image, per-ROI tables);
ratio TIFFs and their 16-bit previews are produced from the device-resident ratio images.
PNG rendering (matplotlib) is outside the device path and is skipped with a log line.
"""
import os

import numpy as np

from .. import batch, roi_ops
from . import common
from ._fretnames import build_pairs_by_channel, load_roi_polys, parse_tokens  # noqa: F401
from .common import ensure_dir, list_tifs

LANG_CURRENT = "en"

DEFAULT_P = {"timelapse": False, "donor_ch": 1, "fret_ch": 2, "ratio_mode": "FRET/Donor", "bg_scope": "full",
             "bg_mode": "percentile", "percentile": 1.0, "per_channel_p": False, "donor_p": 1.0, "fret_p": 1.0,
             "clip_neg": True, "eps_percentile": 1.0, "out_xls": True, "out_tif": True, "out_png": False,
             "save_full": False, "save_crop": False, "mask_outside": True}


def _engine():
    import imageprocess_b200 as ipb
    return ipb.engine()


def process_one_stage(stage_key, pairs_for_stage, p, paths, eng=None, frames_per_batch=32, timing=None):
    import time
    t_scan = time.perf_counter()
    eng = eng or _engine()
    logs = [f"[Stage {stage_key}] start"]
    RES_ROOT, RAT32, RAT16, RROI32, RROI16, PNG_FULL, PNG_CROP = paths
    timelapse = bool(p["timelapse"])
    out_tif = bool(p["out_tif"])
    suffix = "FoverD" if p["ratio_mode"] == "FRET/Donor" else "DoverF"
    per_ch = bool(p["per_channel_p"])
    d_p = float(p["donor_p"]) if per_ch else float(p["percentile"])
    a_p = float(p["fret_p"]) if per_ch else float(p["percentile"])
    from .stream import FrameStream
    # (stage, time) pairs are grouped on metadata only (paths, shape from the TIFF header) and decoded
    # per batch by the stream's thread pool; one FrameBatchJob serves every batch of a group
    items, by_shape = [], {}
    for (s, t_code), dpath, apath in pairs_for_stage:
        stid = f"{s}_{t_code}" if (timelapse and t_code is not None) else s
        logs.append(f"  - Processing: {stid}")
        polys = load_roi_polys(p["roi_dir"], s, t_code, timelapse=timelapse)
        if not polys:
            logs.append(f"    [Warn] ROI missing: {stid}.json ? skip ROI-based outputs")
        try:
            shape = common.image_shape(dpath)
        except Exception as e:
            logs.append(f"    [Error] {stid}: {e}")
            continue
        by_shape.setdefault(shape, []).append((s, t_code, stid, dpath, apath, polys))
    rows_stage = []
    if timing is not None:
        timing["scan_s"] = timing.get("scan_s", 0.0) + time.perf_counter() - t_scan      # ROI JSONs + TIFF headers
    t_rows = 0.0

    def load(it, out=None):
        if out is not None and common.read_plane_into(it[3], out[0]) and common.read_plane_into(it[4], out[1]):
            return None                                          # plain uint16 TIFFs: file -> pinned buffer
        return np.stack([common.as_u16_plane(common.read_image_raw(it[3]), f"{it[2]} donor"),
                         common.as_u16_plane(common.read_image_raw(it[4]), f"{it[2]} fret")])

    for (H, W), group in by_shape.items():
        make_job = lambda shape: batch.FrameBatchJob(eng, shape, stages=("fret",), fret_p=p, want_roi_image=out_tif)
        # ratio images of a result live in the job's (single) device buffers: with TIFF output every
        # batch is collected before the next one is submitted
        stream = FrameStream(eng, (2, H, W), make_job, frames_per_batch=min(frames_per_batch, len(group)),
                             decode_threads=int(p.get("n_workers", 8) or 8), lag=0 if out_tif else 2)
        try:
            for pos, res in stream.run(group, load, lambda it: it[5] or []):
                t_r = time.perf_counter()
                F = stream.F
                rows_pf = batch.rows_fret(res, F)
                if out_tif:
                    R = res.R.host()
                    prev = roi_ops.preview_u16_batch(eng, res.R, 1.0, 99.0)
                    Rroi = res.R_roi.host()
                    prev_roi = roi_ops.preview_u16_batch(eng, res.R_roi, 1.0, 99.0)
                for f, k in enumerate(pos):
                    s, t_code, stid, _, _, polys = group[k]
                    if k in stream.errors:
                        logs.append(f"    [Error] {stid}: {stream.errors[k]}")
                        continue
                    if out_tif:
                        common.write_tiff(os.path.join(RAT32, f"{stid}_ratio_{suffix}.tif"), R[f])
                        pv = prev[f] if prev[f] is not None else np.zeros(R[f].shape, np.uint16)
                        common.write_tiff(os.path.join(RAT16, f"{stid}_ratio_{suffix}_preview.tif"), pv)
                    if not polys:
                        continue
                    if out_tif:
                        common.write_tiff(os.path.join(RROI32, f"{stid}_ratio_{suffix}.tif"), Rroi[f])
                        pv = prev_roi[f] if prev_roi[f] is not None else np.zeros(R[f].shape, np.uint16)
                        common.write_tiff(os.path.join(RROI16, f"{stid}_ratio_{suffix}_preview.tif"), pv)
                    eps = float(res.fret_params[f, 2])
                    for r in rows_pf[f]:
                        r.update({"stage": s, "time": (t_code if timelapse else None), "eps": eps, "p": p["percentile"],
                                  "donor_p": d_p, "fret_p": a_p, "ratio_mode": p["ratio_mode"],
                                  "bg_scope": p["bg_scope"], "bg_mode": p["bg_mode"], "clip_neg": p["clip_neg"],
                                  "eps_p": p["eps_percentile"]})
                    rows_stage.extend(rows_pf[f])
                t_rows += time.perf_counter() - t_r
        finally:
            stream.close()
            if timing is not None:
                for k, v in stream.timing.items():
                    timing[k] = timing.get(k, 0) + v
                timing["rows_s"] = timing.get("rows_s", 0.0) + t_rows
    if p.get("out_png"):
        logs.append("  [SKIP-PNG] figure rendering is host matplotlib code outside the device path")
    logs.append(f"[Stage {stage_key}] end (total {len(pairs_for_stage)} time/files)")
    return stage_key, rows_stage, logs


def run_headless(img_dir, roi_dir, out_root=None, p=None, eng=None, log=print, frames_per_batch=32, timing=None):
    """_pipeline_thread without Tk (fret_ratio_builder.py:892-1011): pair files, group by stage,
    process, write RES/xls/fret_ratio_perROI.csv."""
    p = {**DEFAULT_P, **(p or {}), "img_dir": img_dir, "roi_dir": roi_dir}
    res_root = ensure_dir(out_root or os.path.join(img_dir, "RES"))
    tif = os.path.join(res_root, "TIF")
    paths = (res_root,) + tuple(ensure_dir(os.path.join(tif, d)) if p["out_tif"] else None
                                for d in ("ratio32", "ratio16_preview", "ratio32_roi", "ratio16_roi_preview")) + (None, None)
    pairs, _ = build_pairs_by_channel(list_tifs(img_dir), bool(p["timelapse"]), int(p["donor_ch"]), int(p["fret_ch"]))
    stages = {}
    for pr in pairs:
        stages.setdefault(pr[0][0], []).append(pr)
    rows_all = []
    for stage_key, prs in stages.items():
        _, rows, logs = process_one_stage(stage_key, prs, p, paths, eng=eng, frames_per_batch=frames_per_batch, timing=timing)
        rows_all.extend(rows)
        for line in logs:
            log(line)
    if rows_all and p["out_xls"]:
        import time
        t0 = time.perf_counter()
        save_tables(rows_all, bool(p["timelapse"]), ensure_dir(os.path.join(res_root, "xls")), log=log)
        if timing is not None:
            timing["save_s"] = timing.get("save_s", 0.0) + time.perf_counter() - t0
    elif p["out_xls"]:
        log("[Warn] No ROI \u2192 metric table not generated.")
    return rows_all


CSV_COLS = ["stage", "time", "roi", "area_px", "ratio_mean", "ratio_median", "ratio_std", "ratio_p5", "ratio_p95",
            "donor_mean", "donor_median", "yfret_mean", "yfret_median", "eps", "p", "ratio_mode", "bg_mode"]


def per_roi_frame(rows_all, timelapse):
    """The per_ROI table of _pipeline_thread (fret_ratio_builder.py:980-990): the reference's 17
    columns in its order, then time_idx, stage_idx, roi_lab."""
    import re
    import pandas as pd
    df = pd.DataFrame(rows_all)
    df = df[[c for c in CSV_COLS if c in df.columns]].copy()
    df["time_idx"] = [int(re.search(r"t(\d+)", tt).group(1)) for tt in df["time"]] if timelapse else 0
    df["stage_idx"] = [int(re.search(r"S(\d+)", s).group(1)) for s in df["stage"]]
    df["roi_lab"] = ["s%dc%d" % (si, r) for si, r in zip(df["stage_idx"], df["roi"])]
    return df


def save_tables(rows_all, timelapse, xls_dir, log=print):
    """fret_ratio_perROI.csv, and the .xlsx with the two time matrices when an Excel engine is
    installed (fret_ratio_builder.py:991-1011)."""
    df = per_roi_frame(rows_all, timelapse)
    engine = None
    for name in ("xlsxwriter", "openpyxl"):
        try:
            __import__(name)
            engine = name
            break
        except ImportError:
            pass
    if engine:
        import pandas as pd
        with pd.ExcelWriter(os.path.join(xls_dir, "fret_ratio_perROI.xlsx"), engine=engine) as w:
            df.to_excel(w, index=False, sheet_name="per_ROI")
            for what in ("mean", "median"):
                df.pivot(index="time_idx", columns="roi_lab", values=f"ratio_{what}").sort_index() \
                  .to_excel(w, sheet_name=f"ratio_{what}_matrix")
        log("[Saved] xls/fret_ratio_perROI.xlsx")
    df.to_csv(os.path.join(xls_dir, "fret_ratio_perROI.csv"), index=False)
    log("[Saved] xls/fret_ratio_perROI.csv")
    return df
