"""Mirror of src/INT/FA_Analyzer.py's numeric boundary:

    load_image_safe(path) -> float32 2-D array | None                               FA_Analyzer.py:44-72
    analyze_fa_crop(image_crop, roi_mask_crop, config, global_stats)
        -> (results, threshold_val, bw, labeled_img)                                 FA_Analyzer.py:123-195
    run_batch(file_list, params, px_size, out_root, save_ok_only)                    FA_Analyzer.py:939-1044

analyze_fa_crop keeps the reference's call shape for the GUI's single-crop paths (live preview,
single-file run, export dialog): the crop and its mask are uploaded, the whole chain
(threshold & mask, 4-connected small-object removal, disk closing, 8-connected labelling in
raster order, per-adhesion sums) runs in one ipb_fa_segment call, and the results dict has the
reference's item fields and dtypes.  `contour` (skimage.measure.find_contours, used only to draw
outlines) is left None: SURVEY.md 8(f) item 2.  run_batch is the batch body: every image of
the folder that shares a shape goes through one batch.FrameBatchJob("fa").
"""
import glob
import json
import os

import numpy as np

from .. import batch, ops, pipeline
from ..ops import COMP, CROP
from . import common
from .common import ensure_dir


def _engine():
    import imageprocess_b200 as ipb
    return ipb.engine()


def load_image_safe(path):
    """2-D float32 image; (C,H,W) / (H,W,C) stacks give their first channel; None on failure."""
    try:
        from PIL import Image
        with Image.open(path) as im:
            img = np.array(im)
    except Exception:
        return None
    if img.ndim == 3:
        if img.shape[0] < img.shape[1] and img.shape[0] < img.shape[2]:
            img = img[0]
        elif img.shape[2] < img.shape[0] and img.shape[2] < img.shape[1]:
            img = img[:, :, 0]
    while img.ndim > 2:
        img = img[0]
    return img.astype(np.float32)


def global_stats(img):
    """(mean, std, bg) of FA_Analyzer.py:984-987 for one image, on the device (exact integer
    moments + the [::10, ::10] sample percentile)."""
    eng = _engine()
    plane = common.as_u16_plane(img)[None, None]
    job = batch.FrameBatchJob(eng, plane.shape, stages=("fa",), fa_params={"alpha": 0.0, "min_area_um": 0.0,
                              "max_area_um": 1.0, "close_radius": 0}, fa_px=1.0)
    res = job.run(eng.mem.from_host(plane), [[]])
    m, s, bg, _ = res.fa_stats[0]
    return np.float32(m), np.float32(s), np.float32(bg)


def convert_um_to_px_config(params, px_size):
    """FA_Analyzer.py:527-535."""
    return batch.fa_um_to_px_config(params, px_size)


def analyze_fa_crop(image_crop, roi_mask_crop, config, global_stats, eng=None, with_contours=True):
    """analyze_fa_crop (FA_Analyzer.py:123-195) on the device.  'contour' of every adhesion is
    find_contours(labeled_img == k, 0.5)[0] as in the reference (FA_Analyzer.py:168-170), produced
    from one pass over the label map (ipb_fa_contour_cells) instead of one scan per adhesion; an
    adhesion without any contour is dropped like the reference does (`if not contours: continue`)."""
    empty = {"OK": [], "Large": [], "Small": []}
    image_crop = np.asarray(image_crop)
    if image_crop.size == 0:                                     # FA_Analyzer.py:125-126
        return empty, 0, np.zeros(image_crop.shape, dtype=bool), np.zeros(image_crop.shape, dtype=int)
    eng = eng or _engine()
    mem = eng.mem
    m, s, bg = global_stats
    thr = np.float32(m) + config["alpha"] * np.float32(s) if isinstance(m, np.floating) else m + config["alpha"] * s
    h, w = image_crop.shape
    wpr = (w + 31) // 32
    plane = common.as_u16_plane(image_crop, "image_crop")
    padded = np.zeros((h, wpr * 32), dtype=np.uint8)
    padded[:, :w] = np.asarray(roi_mask_crop, dtype=bool)
    pool = mem.from_host(np.packbits(padded, axis=1, bitorder="little").view(np.uint32).reshape(-1))
    crops = np.zeros(1, dtype=CROP)
    crops["w"], crops["h"], crops["wpr"] = w, h, wpr
    fa_params = mem.from_host(np.array([[m, s, bg, thr]], dtype=np.float32))
    words = h * wpr
    bufs = [mem.empty(words, np.uint32) for _ in range(4)]
    L, cs = mem.empty(h * w, np.int32), mem.empty(h * w, np.uint32)
    rr, rb, cc = mem.empty(h, np.int32), mem.empty(h, np.int32), mem.empty(1, np.int32)
    comp_off = mem.zeros(2, np.int32)
    cap = ((h + 1) // 2) * ((w + 1) // 2)
    comps = mem.empty(cap, COMP)
    labels = mem.empty(h * w, np.int32)
    d_crops = mem.from_host(crops)
    min_px, rad = config["min_px"], config["close_radius"]
    d_plane = mem.from_host(plane.reshape(1, h, w))          # named: a temporary would be freed before the call runs
    eng.call("ipb_fa_segment", d_crops.ptr, 1, h, h, d_plane.ptr, h, w, fa_params.ptr,
             pool.ptr, float(min_px) if min_px > 0 else 0.0, int(rad) if rad > 0 else 0, bufs[0].ptr, bufs[1].ptr,
             L.ptr, cs.ptr, bufs[2].ptr, rr.ptr, rb.ptr, cc.ptr, bufs[3].ptr, comp_off.ptr, comps.ptr, cap,
             labels.ptr, 0, None, 8, mem.stream)
    off = comp_off.host()
    res = batch.BatchResult()
    res.n_rois, res.frame, res.roi = 1, np.array([0]), np.array([1])
    res.fa_comp_off = off
    res.fa_comps = comps.host()[: int(off[1])]
    res.fa_stats = np.array([[m, s, bg, thr]], dtype=np.float32)
    contours = None
    if with_contours and (h < 2 or w < 2) and int(off[1]) > 0:
        # find_contours(labeled_img == k, 0.5) of the reference (FA_Analyzer.py:168) refuses such crops
        raise ValueError("Input array must be at least 2x2.")
    if with_contours and h >= 2 and w >= 2:
        from .. import contours as ct
        rec = mem.empty((h * w, 2), np.uint32)
        rec_n = mem.empty(1, np.uint32)
        eng.call("ipb_fa_contour_cells", d_crops.ptr, 1, h * w, labels.ptr, rec.ptr, rec_n.ptr, mem.stream)
        contours = [ct.contours_of_crop(rec.host(), int(rec_n.host()[0]), w)]
    items = batch.fa_items(res, config, contours=contours)[0]
    bits = np.unpackbits(bufs[3].host().reshape(h, wpr).view(np.uint8), axis=1, bitorder="little")[:, :w].astype(bool)
    return items, thr, bits, labels.host().reshape(h, w)


def load_file_list(img_dir, roi_dir, target_ch="1"):
    """(img_path, json_path, s_tag) triples: *_<ch>.tif paired with <Sxx>.json (FA_Analyzer.py:537-566)."""
    out = []
    for img_path in sorted(glob.glob(os.path.join(img_dir, "*.tif")) + glob.glob(os.path.join(img_dir, "*.TIF"))):
        fname = os.path.basename(img_path)
        if f"_{target_ch}.tif" in fname or f"_{target_ch}.TIF" in fname:
            s_tag = fname.split("_")[0]
            json_path = os.path.join(roi_dir, f"{s_tag}.json")
            if os.path.exists(json_path) and (img_path, json_path, s_tag) not in out:
                out.append((img_path, json_path, s_tag))
    return out


def load_rois(json_path):
    """ROI list as the batch loop reads it (FA_Analyzer.py:989-995): list items are polygons."""
    with open(json_path, "r") as f:
        data = json.load(f)
    rois = []
    for item in data.get("rois", []):
        pts = item if isinstance(item, list) else item.get("rois", item)
        if pts:
            rois.append(np.array(pts, dtype=float))
    return rois


def run_batch(file_list, params, px_size, out_root, save_ok_only=True, eng=None, log=print, frames_per_batch=16):
    """_run_batch_process body: one <Sxx>_results.csv per image under individual_results/."""
    eng = eng or _engine()
    indiv = ensure_dir(os.path.join(out_root, "individual_results"))
    from .stream import FrameStream
    by_shape = {}
    for img_path, json_path, s_tag in file_list:
        log(f"Processing {s_tag}...")
        try:
            shape = common.image_shape(img_path)
        except Exception:
            log(f"  [Error] Failed to load image: {s_tag}")
            continue
        by_shape.setdefault(shape, []).append((s_tag, img_path, load_rois(json_path)))
    count = 0

    def load(it):
        img = load_image_safe(it[1])
        if img is None:
            raise IOError("unreadable image")
        return common.as_u16_plane(img, it[0])[None]

    cfg = convert_um_to_px_config(params, px_size)
    for (H, W), group in by_shape.items():
        make_job = lambda shape: batch.FrameBatchJob(eng, shape, stages=("fa",), fa_params=params, fa_px=px_size, fa_ch=0,
                                                     fa_config=cfg)
        stream = FrameStream(eng, (1, H, W), make_job, frames_per_batch=min(frames_per_batch, len(group)))
        try:
            for pos, res in stream.run(group, load, lambda it: it[2]):
                rows_pf = batch.rows_fa(res, cfg, params, px_size, stream.F, save_ok_only)
                for f, k in enumerate(pos):
                    s_tag = group[k][0]
                    if k in stream.errors:
                        log(f"  [Error] Failed to load image: {s_tag}")
                        continue
                    rows = [{"File": s_tag, **r} for r in rows_pf[f]]
                    if rows:
                        common.write_rows_csv(os.path.join(indiv, f"{s_tag}_results.csv"), rows)
                        count += 1
        finally:
            stream.close()
    log(f"Done. Processed {count} files.")
    return count
