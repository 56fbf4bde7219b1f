"""The complete per-frame hot path for a time-lapse batch (BASELINE.json config C4):
FRET ratio imaging + per-ROI intensity + focal-adhesion segmentation on F two-channel
uint16 frames resident in HBM.  Frames are independent, so multi-GPU runs shard them by
rank (pipeline.shard_frames) and no collective touches pixel data."""
import numpy as np

from . import pipeline


class TimelapseJob:
    def __init__(self, eng, shape, polys_per_frame, fret_p, int_task, fa_params=None, fa_px=0.112,
                 donor_ch=0, acc_ch=1, fa_ch=0):
        self.eng, self.shape = eng, tuple(shape)
        self.polys_per_frame = polys_per_frame
        self.fret_p, self.int_task, self.fa_params, self.fa_px = fret_p, int_task, fa_params, fa_px
        self.donor_ch, self.acc_ch, self.fa_ch = donor_ch, acc_ch, fa_ch
        self.stages = ["fret", "roi_intensity"] + (["fa"] if (fa_params and hasattr(pipeline, "fa_batch")) else [])
        self.n_roi_px = None

    def run(self, planes):
        eng, shape = self.eng, self.shape
        out = {}
        fr = pipeline.fret_batch(eng, planes, shape, self.polys_per_frame, self.fret_p,
                                 self.donor_ch, self.acc_ch)
        out["fret_rows"] = fr["rows_per_frame"]
        out["R"] = fr["R"]
        rows, bg, rm = pipeline.intensity_batch(eng, planes, shape, self.polys_per_frame, self.int_task)
        out["int_rows"], out["int_bg"] = rows, bg
        d2h = 0
        for rpf in (fr["rows_per_frame"], rows):
            d2h += sum(len(r) for r in rpf) * 3 * 56
        if "fa" in self.stages:
            fa = pipeline.fa_batch(eng, planes, shape, self.polys_per_frame, self.fa_params, self.fa_px,
                                   channel=self.fa_ch)
            out["fa_rows"] = fa["rows_per_frame"]
            d2h += fa["d2h_bytes"]
        if self.n_roi_px is None:
            self.n_roi_px = int(sum(r["area_px"] for rr in rows for r in rr))
        out["d2h_bytes"] = d2h
        return out

    def algorithmic_bytes(self, entry):
        """Compulsory bytes one launch of `entry` moves (DESIGN.md 'Kernels')."""
        F, C, H, W = self.shape
        px = F * H * W
        roi_px = self.n_roi_px or 0
        return {
            "ipb_hist_u16": 2 * px * C,                 # each sampled plane read once
            "ipb_fret_pixels": 8 * px,                  # 2 x uint16 in, float32 ratio out
            "ipb_region_stats": 4 * roi_px * 3,         # one value per ROI pixel per job
            "ipb_rasterize_rois": roi_px // 8 + 1,      # bit masks written
            "ipb_fa_segment": 2 * roi_px,               # crop pixels of the FA channel read once
        }.get(entry, 0)
