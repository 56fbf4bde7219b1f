"""Mirror of src/MOR_by_ROI.py:

    morphology_from_polygon(poly, shape, px_um) -> dict of 13 values            MOR_by_ROI.py:211-241

The O(H*W) part (rasterisation, area, first / second pixel moments) runs on the device; the
polygon-only math is host numpy as in the reference.  `run_headless` is main() (:401-505)
without Tk / matplotlib: one batched device call per image, morphology_perROI.csv out.
"""
import os

from .. import roi_ops
from . import common
from ._fretnames import load_roi_polys, parse_tokens
from .common import ensure_dir, fmt_stage, fmt_time, list_tifs

COLUMNS = ["stage", "time", "roi", "img", "channel", "px_um", "area_px", "area_um2", "perimeter_px", "perimeter_um",
           "major_um", "minor_um", "aspect_ratio", "orientation_deg", "circularity", "roundness", "solidity",
           "centroid_x", "centroid_y"]                                            # MOR_by_ROI.py:501-505


def _engine():
    import imageprocess_b200 as ipb
    return ipb.engine()


def morphology_from_polygon(poly, shape, px_um, eng=None):
    return roi_ops.morphology_batch(eng or _engine(), [poly], shape, px_um)[0]


def run_headless(img_dir, roi_dir, sel_ch=1, px_um=0.223, timelapse=False, include_no_channel=False, eng=None,
                 log=print):
    eng = eng or _engine()
    rows = []
    for img_path in list_tifs(img_dir):
        base = os.path.basename(img_path)
        s_num, t_num, ch = parse_tokens(base, timelapse)
        if (ch is None and not include_no_channel) or (ch is not None and ch != sel_ch):
            continue
        if s_num is None:
            log(f"[skip] no stage token: {base}")
            continue
        S = fmt_stage(s_num)
        t_code = fmt_time(t_num) if (timelapse and t_num is not None) else None
        polys = load_roi_polys(roi_dir, S, t_code, timelapse)
        if not polys:
            log(f"[warn] ROI not found: {S if t_code is None else S + '_' + t_code}.json")
            continue
        H, W = common.read_image_raw(img_path).shape
        for i, met in enumerate(roi_ops.morphology_batch(eng, polys, (H, W), px_um), 1):
            met.update({"stage": S, "time": (t_code if timelapse else None), "roi": i, "px_um": px_um, "img": base,
                        "channel": sel_ch})
            rows.append(met)
    if rows:
        rows.sort(key=lambda r: (r["stage"], r["time"] is None, r["time"] or "", r["roi"]))
        out = ensure_dir(os.path.join(ensure_dir(os.path.join(img_dir, "RES_MOR")), "xls"))
        common.write_rows_csv(os.path.join(out, "morphology_perROI.csv"), rows, columns=COLUMNS)
    return rows
