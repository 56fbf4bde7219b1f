"""Mirror of src/roi_channel_cropper.py's compute body.  In the reference the
math sits inside the nested closure run_crop (:778-969, math at :884-968) and cannot be imported;
here it is a function:

    crop_rois(raw, polys, low_cut, high_cut, gamma, mask_outside) -> [dict | None]

per ROI: bbox + pad crop, percentile clip-normalisation, ROI mask, gamma, uint16 and masked raw
outputs -- all per-pixel work on the device (roi_ops.cropper_batch).  `run_headless` walks a
folder like run_crop does and writes the TIFF16 / raw-crop TIFFs (PNG rendering is matplotlib
host code and is skipped).
"""
import os

from .. import roi_ops
from . import common
from ._fretnames import load_roi_polys
from ._fretnames import parse_tokens_cropper as parse_tokens      # this script's own grammar (:211-252)
from .common import ensure_dir, fmt_stage, fmt_time, list_tifs


def _engine():
    import imageprocess_b200 as ipb
    return ipb.engine()


def crop_rois(raw, polys, low_cut=1.0, high_cut=1.0, gamma=1.0, mask_outside=True, pad_ratio=0.05, eng=None):
    return roi_ops.cropper_batch(eng or _engine(), common.as_u16_plane(raw), polys, low_cut, high_cut, gamma,
                                 mask_outside=mask_outside, pad_ratio=pad_ratio)


def run_headless(img_dir, roi_dir, ch_select=1, low_cut=1.0, high_cut=1.0, gamma=1.0, mask_outside=True,
                 timelapse=False, do_tif16=True, do_tif_raw=True, out_root=None, eng=None, log=print):
    eng = eng or _engine()
    res_root = ensure_dir(out_root or os.path.join(img_dir, "RES_CROP"))
    tif16_dir = ensure_dir(os.path.join(res_root, "TIF16")) if do_tif16 else None
    tif_dir = ensure_dir(os.path.join(res_root, "TIF")) if do_tif_raw else None
    done = []
    for img_path in list_tifs(img_dir):
        s_num, t_num, ch = parse_tokens(os.path.basename(img_path), timelapse)
        if s_num is None or ch != ch_select:
            continue
        S = fmt_stage(s_num)
        t_code = fmt_time(t_num) if (timelapse and t_num is not None) else None
        keytag = f"{S}_{t_code}" if t_code is not None else S
        polys = load_roi_polys(roi_dir, S, t_code, timelapse, legacy=False)
        if not polys:
            log(f"[warn] ROI not found: {keytag}.json")
            continue
        raw = common.read_image_raw(img_path)
        for i, res in enumerate(crop_rois(raw, polys, low_cut, high_cut, gamma, mask_outside, eng=eng), 1):
            if res is None:
                log(f"[warn] normalisation failed: {keytag}_roi{i}")
                continue
            if do_tif16:
                common.write_tiff(os.path.join(tif16_dir, f"{keytag}_roi{i}_ch{ch_select}.tif"), res["out16"])
            if do_tif_raw:
                common.write_tiff(os.path.join(tif_dir, f"{keytag}_roi{i}_ch{ch_select}.tif"), res["raw_out"])
            done.append((keytag, i, res["rect"]))
    return done
