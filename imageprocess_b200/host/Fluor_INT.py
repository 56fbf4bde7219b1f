"""Mirror of the reference's src/INT/Fluor_INT.py worker boundary:

    _process_key_task(task) -> {"rows": [dict], "steps": int, "logs": [str]}     Fluor_INT.py:795

`task` is the dict the reference's _run_pipeline builds (keys at Fluor_INT.py:2144-2206); rows
carry exactly the reference's columns (Fluor_INT.py:516-520,876-891).  The per-pixel work --
ROI rasterisation, backgrounds, per-ROI statistics, bg-corrected images and their 16-bit
previews -- runs on the device through imageprocess_b200.batch / roi_ops.  Like the
reference's worker it never raises: errors come back as a log line (Fluor_INT.py:1139-1143).
`process_key_tasks` is the batched form (many (stage, time) keys per launch sequence) that
`run_headless` uses instead of the reference's process pool.
"""
import os
import re

import numpy as np

from .. import batch, geometry as geo, ops
from . import common
from .common import ensure_dir, fmt_stage, fmt_time, list_tifs  # noqa: F401  (reference names)

LANG_CURRENT = "ko"
_MSG = {
    "log_no_ch": ("[SKIP] {stid} — 채널 없음", "[SKIP] {stid} — no channel"),
    "log_no_roi": ("[SKIP] {stid} — ROI 없음", "[SKIP] {stid} — no ROI"),
    "log_done_quant": ("[DONE-QUANT] {stid} ROI={roi_count}", "[DONE-QUANT] {stid} ROI={roi_count}"),
}


def t(key, default=None):
    ko, en = _MSG.get(key, (default, default))
    return en if LANG_CURRENT == "en" else ko


# ---------------------------------------------------------------------- file-name grammar (T2)
def parse_tokens(basename, timelapse):
    """(stage, time, channel) of `S<stage>[_t<time>]_<ch>.tif` style names: stage / time are the
    S<n> / t<n> tokens between separators, the channel is a ch<n> / c<n> token or else the last
    all-digit token that is not the time token's digits (Fluor_INT.py:285-322)."""
    name = os.path.splitext(basename)[0]
    sep = r"(?:^|[_-])"
    end = r"(?=$|[_-])"
    ms = re.search(rf"(?i){sep}S(\d+){end}", name)
    s_num = int(ms.group(1)) if ms else None
    t_num, t_str = None, None
    if timelapse:
        mt = re.search(rf"(?i){sep}t(\d+){end}", name)
        if mt:
            t_str = mt.group(1)
            t_num = int(t_str)
    mc = re.search(rf"(?i){sep}(?:ch|c)(\d{{1,3}}){end}", name)
    ch = None
    if mc:
        ch = int(mc.group(1))
    else:
        nums = [tok for tok in re.split(r"[_-]", name) if tok.isdigit()]
        if timelapse and t_str is not None:
            nums = [n for n in nums if n != t_str]
        if nums:
            ch = int(nums[-1])
    return s_num, t_num, ch


def clean_base_for_save(basename, timelapse):
    s_num, t_num, _ = parse_tokens(basename, timelapse)
    if s_num is None:
        return re.sub(r"([_-])\d+$", "", os.path.splitext(basename)[0])
    if timelapse and t_num is not None:
        return f"{fmt_stage(s_num)}_{fmt_time(t_num)}"
    return fmt_stage(s_num)


def find_roi_basepath(roi_dir, basename, timelapse):
    """S01[_t00] first, legacy S1[_t0] second (Fluor_INT.py:333-346)."""
    s_num, t_num, _ = parse_tokens(basename, timelapse)
    cands = [os.path.join(roi_dir, clean_base_for_save(basename, timelapse))]
    if s_num is not None:
        legacy = f"S{int(s_num)}"
        if timelapse and t_num is not None:
            legacy = f"{legacy}_t{int(t_num)}"
        cands.append(os.path.join(roi_dir, legacy))
    for b in cands:
        if os.path.exists(b + ".json") or os.path.exists(b + ".png"):
            return b
    return cands[0]


def build_keymap(files, timelapse):
    """{(Sxx, txx | None): {channel: path}} ordered by (stage, time) (Fluor_INT.py:372-394)."""
    out = {}
    for p in files:
        s_num, t_num, ch = parse_tokens(os.path.basename(p), timelapse)
        if s_num is None or ch is None:
            continue
        key = (fmt_stage(s_num), fmt_time(t_num) if (timelapse and t_num is not None) else None)
        out.setdefault(key, {})[ch] = p

    def order(item):
        s, tt = item[0]
        return (int(re.search(r"\d+", s).group()) if s else -1, int(re.search(r"\d+", tt).group()) if tt else -1)
    return dict(sorted(out.items(), key=order))


def load_roi_polys_or_mask(roi_folder, s, t_code, timelapse, img_shape=None):
    """(polys, None) from the ROI JSON, else (None, union mask) from a PNG, else (None, None)
    (Fluor_INT.py:405-441)."""
    base = f"{s}_{t_code}" if (timelapse and t_code is not None) else s
    roi_base = find_roi_basepath(roi_folder, base, timelapse)
    if os.path.exists(roi_base + ".json"):
        polys = common.load_roi_json(roi_base + ".json")
        if polys:
            return polys, None
    if os.path.exists(roi_base + ".png"):
        from PIL import Image
        with Image.open(roi_base + ".png") as im:
            mask = np.array(im.convert("L")) > 0
        if img_shape is not None and mask.shape != tuple(img_shape):
            H, W = img_shape
            mask = mask[:H, :W]
            mask = np.pad(mask, ((0, H - mask.shape[0]), (0, W - mask.shape[1])), mode="constant")
        return None, mask
    return None, None


# ---------------------------------------------------------------------- device-backed numeric functions
def _engine():
    import imageprocess_b200 as ipb
    return ipb.engine()


def rasterize_polygon(poly, shape, eng=None):
    """Fluor_INT.rasterize_polygon (398-403) on the device: bool (H, W)."""
    eng = eng or _engine()
    H, W = int(shape[0]), int(shape[1])
    rm = eng.rasterize(geo.RULE_MPL, [geo.mpl_spec(poly, (W, H), store_full=True)], (H, W), 1, want_union=False)
    return rm.mask_host(0)


def _mask_bits(mask):
    H, W = mask.shape
    wpr = (W + 31) // 32
    padded = np.zeros((H, wpr * 32), dtype=np.uint8)
    padded[:, :W] = mask
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(H, wpr)


def _quantify_mask(eng, planes_dev, shape, chs, task, mask, roi_id):
    """union-mask PNG / whole-frame branch of quantify_per_roi_multi (Fluor_INT.py:522-538) plus
    the backgrounds of bg_correct: one region (the bitmap, or all pixels) per frame."""
    from ..ops import HIST_JOB, PAT_FULL, PAT_MASKED, PAT_MASKED_STRIDE, PAT_STRIDE1D, Q_JOB, QK_MEDIAN, QK_PCT, REGION, SRC_U16, STAT_JOB, q32_of
    mem = eng.mem
    F, C, H, W = shape
    assert F == 1
    wpr = (W + 31) // 32
    bits = _mask_bits(mask if mask is not None else np.ones((H, W), dtype=bool))
    d_bits = mem.from_host(bits.reshape(1, H, wpr))
    stride = int(task["bg_stride"]) if task.get("bg_stride") else 1
    scoped = task["bg_scope"] == "roi_union" and mask is not None
    jobs = np.zeros(C, dtype=HIST_JOB)
    jobs["plane"] = np.arange(C)
    jobs["k"] = stride
    jobs["pattern"] = ((PAT_MASKED_STRIDE if scoped else PAT_STRIDE1D) if stride > 1 else (PAT_MASKED if scoped else PAT_FULL))
    hres = eng.hist(planes_dev, H, W, jobs, union=d_bits, union_wpr=wpr)
    p_glob = float(task["percentile"])
    pps = [float(task["ch_p_map"].get(ch, p_glob)) if task["per_channel_p"] else p_glob for ch in chs]
    bgs = []
    if task["bg_mode"] == "percentile":
        qj = np.zeros(C, dtype=Q_JOB)
        qj["hist"] = np.arange(C)
        qj["q32"] = [q32_of(pp) for pp in pps]
        qo = eng.quantiles(hres, qj).host()
        bgs = [float(q["value"]) if q["n"] > 0 else 0.0 for q in qo]
    elif task["bg_mode"] == "hist-mode":
        hh = hres.hist.host()
        for ci, pp in enumerate(pps):
            lvl = batch.hist_mode_level(hh[ci], pp)
            bgs.append(0.0 if lvl is None else lvl)
    else:
        bgs = [0.0] * C
    reg = np.zeros(1, dtype=REGION)
    reg["w"], reg["h"], reg["wpr"] = W, H, wpr
    sj = np.zeros(C, dtype=STAT_JOB)
    sj["region"], sj["src"], sj["plane"], sj["n_views"] = 0, SRC_U16, np.arange(C), 1
    sj["bidx"][:, 0] = np.arange(C)
    sj["bidx"][:, 1] = -1
    sj["clip_neg"][:, 0] = int(bool(task["clip_neg"]))
    sj["qkind"] = (QK_PCT, QK_MEDIAN, QK_PCT)
    sj["q32"] = (q32_of(5), 0.0, q32_of(95))
    sj["out"][:, 0] = np.arange(C)
    so = eng.region_stats(reg, sj, d_bits, H, W, planes=planes_dev,
                          bvals=mem.from_host(np.asarray(bgs, dtype=np.float32))).host()
    res = batch.BatchResult()
    res.n_rois, res.roi, res.frame = 1, np.array([roi_id]), np.array([0])
    res.area = np.array([int(so[0]["area"])])
    res.int_stat = so.reshape(1, C)
    rows = batch.rows_intensity(res, 1, list(chs))[0]
    return rows, {ch: {"bg": float(np.float32(b)), "p": float(pp)} for ch, b, pp in zip(chs, bgs, pps)}


def _bg_corrected_images(eng, planes_dev, shape, bgs, clip_neg):
    """J = img - B, negatives clipped (bg_correct, Fluor_INT.py:487-492) for every channel of one
    frame, float32 on the device (the fused FRET pass computes exactly this per channel)."""
    from ..ops import FP_STRIDE, FRET_CFG
    mem = eng.mem
    F, C, H, W = shape
    out = mem.empty((C, H, W), np.float32)
    for ci in range(C):
        cfg = np.zeros(1, dtype=FRET_CFG)
        cfg["clip_neg"], cfg["g_factor"] = int(bool(clip_neg)), 1.0
        cfg["donor_ch"], cfg["acc_ch"], cfg["aonly_ch"], cfg["n_ch"] = ci, ci, -1, C
        fp = np.zeros((1, FP_STRIDE), dtype=np.float32)
        fp[0, 0] = fp[0, 1] = np.float32(bgs[ci])
        d_fp = mem.from_host(fp)                                 # named: a temporary would be freed before the call runs
        eng.call("ipb_fret_pixels", planes_dev.ptr, 1, H, W, cfg.ctypes.data, d_fp.ptr, None, 0,
                 None, None, None, None, out.ptr + 4 * ci * H * W, None, None, None, 0, mem.stream)
    return out


def _process_one(task, eng):
    global LANG_CURRENT
    if "lang" in task:
        LANG_CURRENT = task["lang"]
    s, t_code, stid = task["s"], task["t"], task["stid"]
    timelapse = task["timelapse"]
    raw = {}
    for ch in task["chs_to_quant"]:
        pth = task["chmap"].get(ch)
        if pth is not None:
            raw[ch] = common.read_image_raw(pth)
    if not raw:
        return {"rows": [], "steps": 1, "logs": [t("log_no_ch").format(stid=stid)]}
    chs = sorted(raw)
    H, W = next(iter(raw.values())).shape
    polys, union_mask = load_roi_polys_or_mask(task["roi_dir"], s, t_code, timelapse, img_shape=(H, W))
    if polys is None and union_mask is None and task["skip_no_roi"]:
        return {"rows": [], "steps": 1, "logs": [t("log_no_roi").format(stid=stid)]}
    planes = np.stack([common.as_u16_plane(raw[ch], f"{stid} ch{ch}") for ch in chs])[None]
    shape = planes.shape
    dev = eng.mem.from_host(planes)
    if polys is not None:
        job = batch.FrameBatchJob(eng, shape, stages=("int",), int_task=task, int_channels=list(range(len(chs))))
        job.ch_names = chs
        res = job.run(dev, [polys])
        per_roi = batch.rows_intensity(res, 1, chs)[0]
        bg_used = {ch: {"bg": float(res.int_bg[0, ci]), "p": float(res.int_p[ci])} for ci, ch in enumerate(chs)}
        union_for_mask = None
    else:
        per_roi, bg_used = _quantify_mask(eng, dev, shape, chs, task, union_mask, 1 if union_mask is not None else 0)
        union_for_mask = union_mask
    rows = []
    for r in per_roi:
        r.update({"stage": s, "time": t_code if timelapse else None, "bg_scope": task["bg_scope"],
                  "bg_mode": task["bg_mode"], "clip_neg": bool(task["clip_neg"]), "bg_stride": int(task["bg_stride"])})
        for ch in task["chs_to_quant"]:
            if ch in bg_used:
                r[f"ch{ch}_bg"] = bg_used[ch]["bg"]
                r[f"ch{ch}_p"] = bg_used[ch]["p"]
            r[f"ch{ch}_color"] = task["ch_color_map"].get(ch, "Grayscale")
        rows.append(r)
    logs = [t("log_done_quant").format(stid=stid, roi_count=len(per_roi))]
    steps = max(1, len(per_roi))
    if task.get("do_tif"):
        _write_tifs(task, eng, dev, shape, chs, bg_used, polys, union_for_mask, stid, logs)
    if task.get("do_png"):
        logs.append(f"[SKIP-PNG] {stid}: figure rendering is host matplotlib code outside the device path")
    return {"rows": rows, "steps": steps, "logs": logs}


def _write_tifs(task, eng, dev, shape, chs, bg_used, polys, union_mask, stid, logs):
    """bg-corrected float32 TIFF + 16-bit preview per channel (Fluor_INT.py:917-943)."""
    from .. import roi_ops
    F, C, H, W = shape
    bc = _bg_corrected_images(eng, dev, shape, [bg_used[ch]["bg"] for ch in chs], task["clip_neg"]).host()
    if task.get("tif_mask_outside"):
        if polys is not None:
            rm = eng.rasterize(geo.RULE_MPL, [geo.mpl_spec(P, (W, H)) for P in polys], (H, W), 1, want_union=True)
            union_mask = rm.union_host()[0]
        if union_mask is not None:
            bc = np.where(union_mask[None], bc, np.float32(0))
    previews = roi_ops.preview_u16_batch(eng, bc, float(task["auto_lo"]), float(task["auto_hi"]))
    for ci, ch in enumerate(chs):
        common.write_tiff(os.path.join(task["tif32_dir"], f"{stid}_ch{ch}_bgcorr.tif"), bc[ci].astype(np.float32))
        if previews[ci] is not None:
            common.write_tiff(os.path.join(task["tif16_dir"], f"{stid}_ch{ch}_bgcorr_preview.tif"), previews[ci])


def _process_key_task(task, eng=None):
    """One (stage, time) key; same contract as the reference's worker, including never raising."""
    try:
        return _process_one(task, eng or _engine())
    except Exception as e:                                   # Fluor_INT.py:1139-1143
        return {"rows": [], "steps": 1, "logs": [f"[ERROR][WORKER] {task.get('stid', '?')}: {e}"]}


def _rows_of_frame(task, rows, chs, res, f):
    out = []
    for r in rows:
        r.update({"stage": task["s"], "time": task["t"] if task["timelapse"] else None,
                  "bg_scope": task["bg_scope"], "bg_mode": task["bg_mode"],
                  "clip_neg": bool(task["clip_neg"]), "bg_stride": int(task["bg_stride"])})
        for ch in task["chs_to_quant"]:
            if ch in chs:
                r[f"ch{ch}_bg"] = float(res.int_bg[f, chs.index(ch)])
                r[f"ch{ch}_p"] = float(res.int_p[chs.index(ch)])
            r[f"ch{ch}_color"] = task["ch_color_map"].get(ch, "Grayscale")
        out.append(r)
    return out


def process_key_tasks(tasks, eng=None, frames_per_batch=32, decode_threads=8, timing=None):
    """Batched form of the reference's pool over (stage, time) keys (Fluor_INT.py:2211-2229): keys
    with ROI polygons, equal image shape, channel set and settings share ONE FrameBatchJob for the
    whole run, fed by stream.FrameStream (threaded decode -> pinned ring -> async upload, results
    two batches behind); everything else goes through _process_key_task.  Keys are grouped on
    metadata only (paths, shape from the TIFF header) and decoded per batch.  Returns the per-task
    results in task order."""
    import time
    from .stream import FrameStream
    eng = eng or _engine()
    results = [None] * len(tasks)
    groups = {}
    t_scan, t_rows = time.perf_counter(), 0.0
    for i, task in enumerate(tasks):
        try:
            chs = sorted(ch for ch in task["chs_to_quant"] if task["chmap"].get(ch) is not None)
            base = f"{task['s']}_{task['t']}" if (task["timelapse"] and task["t"] is not None) else task["s"]
            roi_base = find_roi_basepath(task["roi_dir"], base, task["timelapse"])
            polys = common.load_roi_json(roi_base + ".json") if os.path.exists(roi_base + ".json") else None
            if not chs or not polys or task.get("do_tif") or task.get("do_png"):
                raise LookupError
            shape = common.image_shape(task["chmap"][chs[0]])
            key = (shape, tuple(chs), task["bg_scope"], task["bg_mode"], int(task["bg_stride"]),
                   float(task["percentile"]), bool(task["per_channel_p"]), tuple(sorted(task["ch_p_map"].items())),
                   bool(task["clip_neg"]))
            groups.setdefault(key, []).append((i, [task["chmap"][ch] for ch in chs], polys))
        except Exception:
            results[i] = _process_key_task(task, eng)
    if timing is not None:
        timing["scan_s"] = timing.get("scan_s", 0.0) + time.perf_counter() - t_scan          # ROI JSONs + TIFF headers
    for key, items in groups.items():
        (H, W), chs = key[0], list(key[1])
        task0 = tasks[items[0][0]]

        def make_job(shape):
            job = batch.FrameBatchJob(eng, shape, stages=("int",), int_task=task0, int_channels=list(range(len(chs))))
            job.ch_names = chs
            return job

        def load(it, out=None):
            if out is not None and all(common.read_plane_into(p, out[ci]) for ci, p in enumerate(it[1])):
                return None                                      # plain uint16 TIFFs: file -> pinned buffer
            return np.stack([common.as_u16_plane(common.read_image_raw(p)) for p in it[1]])
        stream = FrameStream(eng, (len(chs), H, W), make_job, frames_per_batch=min(frames_per_batch, len(items)),
                             decode_threads=decode_threads)
        try:
            for pos, res in stream.run(items, load, lambda it: it[2]):
                t_r = time.perf_counter()
                rows_pf = batch.rows_intensity(res, stream.F, chs)
                for f, k in enumerate(pos):
                    i = items[k][0]
                    task = tasks[i]
                    if k in stream.errors:
                        results[i] = {"rows": [], "steps": 1, "logs": [f"[ERROR][WORKER] {task.get('stid', '?')}: {stream.errors[k]}"]}
                        continue
                    rows = _rows_of_frame(task, rows_pf[f], chs, res, f)
                    results[i] = {"rows": rows, "steps": max(1, len(rows)),
                                  "logs": [t("log_done_quant").format(stid=task["stid"], roi_count=len(rows))]}
                t_rows += time.perf_counter() - t_r
        except Exception as e:
            for i, _, _ in items:
                if results[i] is None:
                    results[i] = {"rows": [], "steps": 1, "logs": [f"[ERROR][WORKER] {tasks[i].get('stid', '?')}: {e}"]}
        finally:
            stream.close()
            if timing is not None:
                for k, v in stream.timing.items():
                    timing[k] = timing.get(k, 0) + v
                timing["rows_s"] = timing.get("rows_s", 0.0) + t_rows
                t_rows = 0.0
    return results


DEFAULT_CFG = {"timelapse": False, "channels_to_quant": [1, 2], "bg_scope": "full", "bg_mode": "percentile",
               "percentile": 1.0, "per_channel_p": False, "ch_p_map": {}, "clip_neg": True, "bg_stride": 4,
               "px_um": None, "out_tif": False, "out_png": False, "tif_mask_outside": False,
               "auto_clip_lo": 1.0, "auto_clip_hi": 99.0, "ch_color_map": {}, "out_xls": True}


def build_tasks(img_dir, roi_dir, out_root, cfg):
    """The task list of _run_pipeline (Fluor_INT.py:2141-2207), without the PNG-only keys."""
    cfg = {**DEFAULT_CFG, **cfg}
    timelapse = bool(cfg["timelapse"])
    keymap = build_keymap(list_tifs(img_dir), timelapse)
    tif32 = ensure_dir(os.path.join(out_root, "TIF", "bgcorr32")) if cfg["out_tif"] else None
    tif16 = ensure_dir(os.path.join(out_root, "TIF", "bgcorr16_preview")) if cfg["out_tif"] else None
    tasks = []
    for (s, t_code), chmap in keymap.items():
        stid = f"{s}_{t_code}" if (timelapse and t_code is not None) else s
        tasks.append({"s": s, "t": t_code, "stid": stid, "chmap": chmap, "chs_to_quant": cfg["channels_to_quant"],
                      "roi_dir": roi_dir, "timelapse": timelapse, "skip_no_roi": True,
                      "bg_scope": cfg["bg_scope"], "bg_mode": cfg["bg_mode"], "percentile": cfg["percentile"],
                      "per_channel_p": cfg["per_channel_p"], "ch_p_map": cfg["ch_p_map"], "clip_neg": cfg["clip_neg"],
                      "bg_stride": cfg["bg_stride"], "px_um": cfg["px_um"], "do_tif": cfg["out_tif"],
                      "do_png": cfg["out_png"], "tif_mask_outside": cfg["tif_mask_outside"], "tif32_dir": tif32,
                      "tif16_dir": tif16, "auto_lo": cfg["auto_clip_lo"], "auto_hi": cfg["auto_clip_hi"],
                      "ch_color_map": cfg["ch_color_map"], "lang": LANG_CURRENT})
    return tasks, keymap


BASE_COLS = ["stage", "time", "roi", "area_px", "bg_mode", "bg_scope", "clip_neg", "bg_stride"]


def per_roi_frame(rows_all):
    """The per_ROI table of save_excel (Fluor_INT.py:728-749): the eight fixed columns, every
    other column in natural order (ch2_bg, ch2_color, ..., ch10_*), then the four derived index /
    label columns.  Returns a pandas DataFrame (None when there are no rows)."""
    import pandas as pd
    df = pd.DataFrame(rows_all)
    if df.empty:
        return None
    for c in BASE_COLS:
        if c not in df.columns:
            df[c] = None
    dyn = sorted((c for c in df.columns if c not in BASE_COLS), key=common.natural_key)
    df = df[BASE_COLS + dyn].copy()
    df["stage_idx"] = [int(re.search(r"S(\d+)", s).group(1)) for s in df["stage"]]
    if df["time"].notna().any():
        df["time_idx"] = [int(re.search(r"t(\d+)", tt if isinstance(tt, str) else "t0").group(1)) for tt in df["time"]]
    else:
        df["time_idx"] = 0
    df["roi_lab"] = ["s%dc%d" % (si, r) for si, r in zip(df["stage_idx"], df["roi"])]
    df["roi_id"] = ["%s_roi%d" % (s, r) for s, r in zip(df["stage"], df["roi"])]
    return df


def save_excel(rows_all, keymap, xls_dir, log=print):
    """fluor_intensity_perROI.csv (+ .xlsx with the per-channel sheets / time matrices when an
    Excel engine is installed) exactly as the reference's save_excel lays them out
    (Fluor_INT.py:728-790)."""
    df = per_roi_frame(rows_all)
    if df is None:
        return None
    csv_path = os.path.join(xls_dir, "fluor_intensity_perROI.csv")
    try:
        import openpyxl  # noqa: F401
        import pandas as pd
        xlsx = os.path.join(xls_dir, "fluor_intensity_perROI.xlsx")
        with pd.ExcelWriter(xlsx, engine="openpyxl") as w:
            df.to_excel(w, index=False, sheet_name="per_ROI")
            chans = sorted({int(m.group(1)) for c in df.columns if (m := re.match(r"ch(\d+)_mean", c))})
            if not any(k[1] is not None for k in keymap):
                for ch in chans:
                    keep = [c for c in ["stage", "roi", "roi_id", "area_px"] if c in df.columns] + \
                           [c for c in df.columns if c.startswith(f"ch{ch}_")]
                    sub = df[keep].sort_values(["stage", "roi"])
                    sub.insert(0, "No.", range(1, len(sub) + 1))
                    sub.to_excel(w, index=False, sheet_name=f"ch{ch}")
            else:
                for ch in chans:
                    for what in ("mean", "median"):
                        df.pivot(index="time_idx", columns="roi_lab", values=f"ch{ch}_{what}").sort_index() \
                          .to_excel(w, sheet_name=f"ch{ch}_{what}_matrix")
    except ImportError:
        log("[INFO] no Excel engine (openpyxl) installed: wrote the CSV only")
    df.to_csv(csv_path, index=False)
    return csv_path


def run_headless(img_dir, roi_dir, out_root=None, cfg=None, eng=None, log=print, frames_per_batch=32, timing=None):
    """_run_pipeline without Tk: tasks -> device batches -> RES/xls/fluor_intensity_perROI.csv."""
    out_root = out_root or os.path.join(img_dir, "RES")
    tasks, keymap = build_tasks(img_dir, roi_dir, out_root, cfg or {})
    rows_all = []
    for res in process_key_tasks(tasks, eng=eng, frames_per_batch=frames_per_batch, timing=timing):
        rows_all.extend(res["rows"])
        for line in res.get("logs", []):
            log(line)
    if rows_all and (cfg or {}).get("out_xls", True):
        import time
        t0 = time.perf_counter()
        save_excel(rows_all, keymap, ensure_dir(os.path.join(out_root, "xls")), log=log)
        if timing is not None:
            timing["save_s"] = timing.get("save_s", 0.0) + time.perf_counter() - t0
    return rows_all
