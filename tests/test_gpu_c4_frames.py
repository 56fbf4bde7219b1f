"""BASELINE config C4 on the GPU at full size: 2048 x 2048 uint16 two-channel frames (24 cells,
60 adhesion blobs per cell) through ONE FrameBatchJob with all three stages -- the path bench.py
times (percentiles by sampled windows, shared-memory FA chain, unit walks, CUDA-graph replay).

* frame 0 against the oracle (the reference's own control flow; ~20 s of CPU): ratio image
  bit-exact, per-ROI ratio / intensity rows (order statistics exact, means within 1e-5), FA
  background exact, binary images and label maps of every cell crop bit-exact, per-adhesion rows.
* size-independent properties over all frames: a replayed step (CUDA graph) reproduces the eager
  step bit for bit; results of a frame do not depend on its position in the batch; no window miss.
"""
import numpy as np
import pytest

import bench
from imageprocess_b200 import batch, pipeline
from oracle import port
from tests import checks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import imageprocess_b200 as ipb
    return ipb.engine("cuda:0")


def _job(eng, shape):
    return batch.FrameBatchJob(eng, shape, stages=("fret", "int", "fa"), fret_p=bench.FRET_P, int_task=bench.INT_TASK,
                               fa_params=bench.FA_PARAMS, fa_px=bench.FA_PX, want_labels=True)


def _snap(res):
    return (res.fret_params.copy(), res.int_bg.copy(), res.fa_stats.copy(), res.fret_stat.copy(), res.int_stat.copy(),
            res.fa_comp_off.copy(), res.fa_comps.copy())


def _same(x, y):
    return all(np.array_equal(np.asarray(p).view(np.uint8), np.asarray(q).view(np.uint8)) for p, q in zip(x, y))


def test_c4_frames(eng):
    frames, polys = bench.make_frames(3, n_unique=2)             # frame 2 = frame 0 + per-pixel jitter
    F = frames.shape[0]
    job = _job(eng, frames.shape)
    dev = eng.mem.from_host(frames)
    res = job.run(dev, [polys] * F)
    pl = job._plans[next(iter(job._plans))]
    assert pl.pq_ok and job.window_misses == 0 and pl.NF == len(polys) * F and res.roi_fallbacks == 0, res.roi_fallback_why
    first = _snap(res)
    R0 = res.R.host()[0].copy()
    labels = res.fa_labels.host().copy()
    bw = res.fa_bw.host().copy()
    rows_f, rows_i = batch.rows_fret(res, F), batch.rows_intensity(res, F, [1, 2])
    rows_a = batch.rows_fa(res, job.fa_cfg, bench.FA_PARAMS, bench.FA_PX, F, save_ok_only=False)

    # ---- frame 0 against the oracle
    d, a = frames[0]
    D, A = d.astype(np.float32), a.astype(np.float32)
    want = port.fret_process_pair(D, A, polys, bench.FRET_P)
    assert np.array_equal(R0, want["R_full"], equal_nan=True)
    assert len(rows_f[0]) == len(want["rows"]) == len(polys)
    for g, w in zip(rows_f[0], want["rows"]):
        assert g["roi"] == w["roi"] and g["area_px"] == w["area_px"]
        for k in ("ratio_median", "ratio_p5", "ratio_p95", "donor_median", "yfret_median"):
            assert g[k] == w[k], (k, g[k], w[k])
        for k in ("ratio_mean", "ratio_std", "donor_mean", "yfret_mean"):
            assert checks.close(g[k], w[k]), (k, g[k], w[k])
    wrows, wbg, _ = port.int_process_key({1: D.copy(), 2: A.copy()}, polys, None, bench.INT_TASK)
    assert float(res.int_bg[0, 0]) == wbg[1]["bg"] and float(res.int_bg[0, 1]) == wbg[2]["bg"]
    checks.check_int_rows(rows_i[0], wrows, (1, 2))
    ref_stats = port.fa_global_stats(D)
    got = res.fa_stats[0]
    assert got[2] == ref_stats[2]                                  # background percentile: exact
    assert checks.close(float(got[0]), float(ref_stats[0]), 1e-6) and checks.close(float(got[1]), float(ref_stats[1]), 1e-6)
    stats = ref_stats
    if np.float32(got[3]) != ref_stats[0] + job.fa_cfg["alpha"] * ref_stats[1]:
        stats = (np.float32(got[0]), np.float32(got[1]), ref_stats[2])       # see checks.check_fa_batch
    view = pipeline._FaView(res)
    n_fa = 0
    for i, P in enumerate(polys):
        crop, mask, rect = port.fa_crop_and_mask(D, P.copy())
        _, thr, wbw, wlab = port.analyze_fa_crop(crop, mask, job.fa_cfg, stats, with_contours=False)
        assert np.array_equal(view.bw_host(i), wbw), i
        assert np.array_equal(view.labels_host(i), wlab), i
        n_fa += int(wlab.max())
    wfa = port.fa_batch_rows(D, polys, bench.FA_PARAMS, bench.FA_PX, save_ok_only=False, with_contours=False, stats=stats)
    assert len(rows_a[0]) == len(wfa) == n_fa and n_fa >= 500             # spread layout: the blobs of a cell stay separate adhesions
    for g, w in zip(rows_a[0], wfa):
        assert g["Cell_ID"] == w["Cell_ID"] and g["Category"] == w["Category"] and g["Area_px"] == w["Area_px"]
        assert checks.close(float(g["Mean_Intensity_Raw"]), float(w["Mean_Intensity_Raw"]))

    # ---- a replayed step reproduces the eager one bit for bit
    for _ in range(6):
        again = job.run(dev, [polys] * F)
    assert _same(_snap(again), first)
    assert np.array_equal(again.fa_labels.host(), labels) and np.array_equal(again.fa_bw.host(), bw)

    # ---- a frame's results do not depend on its position in the batch
    perm = [2, 0, 1]
    job2 = _job(eng, frames.shape)
    res2 = job2.run(eng.mem.from_host(np.ascontiguousarray(frames[perm])), [polys] * F)
    assert job2.window_misses == 0
    assert np.array_equal(res2.fret_params, res.fret_params[perm])
    assert np.array_equal(res2.int_bg, res.int_bg[perm]) and np.array_equal(res2.fa_stats, res.fa_stats[perm])
    rf2, ri2 = batch.rows_fret(res2, F), batch.rows_intensity(res2, F, [1, 2])
    ra2 = batch.rows_fa(res2, job2.fa_cfg, bench.FA_PARAMS, bench.FA_PX, F, save_ok_only=False)

    def same_rows(x, y):                                           # every measured value identical (NaN == NaN)
        if len(x) != len(y):
            return False
        for p, q in zip(x, y):
            if p.keys() != q.keys():
                return False
            for k in p:
                u, v = p[k], q[k]
                both_nan = isinstance(u, float) and isinstance(v, float) and u != u and v != v
                if not (both_nan or u == v):
                    return False
        return True
    for pos, f in enumerate(perm):
        assert same_rows(rf2[pos], rows_f[f]) and same_rows(ri2[pos], rows_i[f]) and same_rows(ra2[pos], rows_a[f]), (pos, f)
