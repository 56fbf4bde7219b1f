"""Batched pipelines with the reference's row formats: thin fronts of batch.FrameBatchJob
(one stage each) used by the host mirrors (Fluor_INT.py, fret_ratio_builder.py,
FA_Analyzer.py) and the parity tests."""
import numpy as np

from . import batch
from .batch import fa_um_to_px_config, hist_mode_level  # noqa: F401  (re-exported)


def intensity_batch(eng, planes, shape, polys_per_frame, task, ch_names=None, int_channels=None):
    """Fluor_INT per-ROI quantification for F frames (reference src/INT/Fluor_INT.py:839-870).
    Returns (rows_per_frame, bg_used_per_frame, BatchResult)."""
    F, C, H, W = shape
    int_channels = list(int_channels) if int_channels is not None else list(range(C))
    ch_names = list(ch_names) if ch_names is not None else [c + 1 for c in int_channels]
    job = batch.FrameBatchJob(eng, shape, stages=("int",), int_task=task, int_channels=int_channels)
    job.ch_names = ch_names
    res = job.run(planes, polys_per_frame)
    rows = batch.rows_intensity(res, F, ch_names)
    bg_used = [{ch: {"bg": float(res.int_bg[f, ci]), "p": float(res.int_p[ci])}
                for ci, ch in enumerate(ch_names)} for f in range(F)]
    return rows, bg_used, res


def fret_batch(eng, planes, shape, polys_per_frame, p, donor_ch=0, acc_ch=1, want_roi_image=False):
    """fret_ratio_builder.process_one_stage numeric body for F pairs (reference
    src/FRET/fret_ratio_builder.py:454-474,493-507)."""
    F = shape[0]
    job = batch.FrameBatchJob(eng, shape, stages=("fret",), fret_p=p, donor_ch=donor_ch, acc_ch=acc_ch,
                              want_roi_image=want_roi_image)
    res = job.run(planes, polys_per_frame)
    return {"rows_per_frame": batch.rows_fret(res, F), "fparams": res.fret_params, "R": res.R,
            "R_roi": res.R_roi, "result": res}


class _FaView:
    def __init__(self, res):
        self.res = res

    def bw_host(self, i):
        c = self.res.fa_crops[i]
        words = self.res.fa_bw.host()[c["bit_off"]: c["bit_off"] + c["h"] * c["wpr"]].reshape(c["h"], c["wpr"])
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")
        return bits[:, :c["w"]].astype(bool)

    def labels_host(self, i):
        c = self.res.fa_crops[i]
        return self.res.fa_labels.host()[c["pix_off"]: c["pix_off"] + c["h"] * c["w"]].reshape(c["h"], c["w"])


def fa_batch(eng, planes, shape, polys_per_frame, params, px_size, channel=0, save_ok_only=True,
             want_labels=False, config=None, fa_path=0, want_contours=False, prefilter=None, threshold=None):
    """FA_Analyzer batch body for F frames (reference src/INT/FA_Analyzer.py:984-1039).

    Optional stages, OFF by default (the reference has none of them; BASELINE.json's north_star names
    them; their oracle is scipy.ndimage / the restated skimage rule): prefilter = ("tophat", size) or
    ("gaussian", sigma) filters the FA channel first (filters.prefilter_planes); threshold = "otsu"
    replaces mean + alpha * std by Otsu's threshold of the (filtered) frame."""
    F = shape[0]
    cfg = config or fa_um_to_px_config(params, px_size)
    if prefilter is not None:
        from . import filters
        planes = filters.prefilter_planes(eng, planes, shape, channel, prefilter)
        shape, channel = (F, 1, shape[2], shape[3]), 0
    job = batch.FrameBatchJob(eng, shape, stages=("fa",), fa_params=params, fa_px=px_size, fa_ch=channel,
                              want_labels=want_labels, fa_config=cfg, want_contours=want_contours)
    job.fa_threshold = threshold
    job.fa_path = fa_path
    res = job.run(planes, polys_per_frame)
    contours = batch.fa_contours(res) if want_contours else None
    owner = [(int(f), int(r)) for f, r in zip(res.frame, res.roi)]
    rects = [tuple(int(v) for v in rc) for rc in res.fa_rect]
    return {"rows_per_frame": batch.rows_fa(res, cfg, params, px_size, F, save_ok_only),
            "stats": res.fa_stats, "result": _FaView(res), "items_per_crop": batch.fa_items(res, cfg),
            "contours": contours,
            "owner": owner, "rects": rects, "d2h_bytes": res.d2h_bytes, "raw": res}
