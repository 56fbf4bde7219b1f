"""tests/fuzz/fuzz_combined.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

All three stages in ONE FrameBatchJob, the bench's configuration (shared ROI rasterisation under both rules, the FA
channel's moments riding on the percentile or the FRET pass, side branches joined before the tables leave): ratio image,
FRET rows, intensity rows, FA binary images / label maps / adhesion counts against the oracle on random scenes with
random parameters, 1-3 frames, ROI lists shared between frames or not, IPB_FRET_MOMENTS on and off.

    python tests/fuzz/fuzz_combined.py <first seed> <number of seeds>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, math, traceback
from imageprocess_b200.ops import Engine
from imageprocess_b200 import batch, pipeline
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close, small_scene
from oracle import port
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time(); stats = {"adhesions": 0, "straddles": 0}
for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    F = int(rng.integers(1, 4))
    H, W = int(rng.choice([96, 120, 136])), int(rng.choice([128, 168, 150]))
    scenes = [small_scene(int(rng.integers(0, 10000)), H=H, W=W, n_cells=int(rng.integers(1, 4)), blobs=int(rng.integers(3, 12))) for _ in range(F)]
    planes = np.stack([np.stack([d, a]) for d, a, _ in scenes])
    shared = rng.random() < 0.5
    polys = [scenes[0][2]] * F if shared else [sc[2] for sc in scenes]
    fret_p = {"bg_scope": str(rng.choice(["full", "roi_union"])), "bg_mode": "percentile", "percentile": float(rng.choice([1.0, 5.0])),
              "per_channel_p": bool(rng.integers(0, 2)), "donor_p": 1.0, "fret_p": 3.0, "clip_neg": bool(rng.integers(0, 2)),
              "eps_percentile": float(rng.choice([1.0, 3.0])), "ratio_mode": str(rng.choice(["Donor/FRET", "FRET/Donor"]))}
    task = {"bg_scope": str(rng.choice(["full", "roi_union"])), "bg_mode": str(rng.choice(["percentile", "hist-mode"])), "clip_neg": bool(rng.integers(0, 2)),
            "bg_stride": int(rng.choice([1, 4])), "percentile": float(rng.choice([1.0, 10.0])), "per_channel_p": False, "ch_p_map": {}}
    px = 0.112
    fa = {"alpha": float(rng.choice([1.0, 2.0, 3.0])), "min_area_um": float(rng.choice([0.0, 5.0, 12.5])) * px ** 2, "max_area_um": 300.0 * px ** 2,
          "close_radius": int(rng.integers(0, 4)), "subtract_bg": bool(rng.integers(0, 2))}
    fa_ch = int(rng.integers(0, 2))
    try:
        job = batch.FrameBatchJob(eng, planes.shape, stages=("fret", "int", "fa"), fret_p=fret_p, int_task=task, fa_params=fa, fa_px=px,
                                  fa_ch=fa_ch, want_labels=True)
        job.pq_min_px = 0 if rng.random() < 0.7 else job.pq_min_px
        job.fret_moments = bool(rng.integers(0, 2))
        job.fa_path = int(rng.choice([0, 1, 2, 3]))
        res = None
        for _ in range(int(rng.integers(1, 4))):                 # later steps reuse the plan (and replay on the GPU)
            res = job.run(eng.mem.from_host(planes), polys)
        rows_i, rows_f = batch.rows_intensity(res, F, [1, 2]), batch.rows_fret(res, F)
        R = res.R.host()
        cfg = pipeline.fa_um_to_px_config(fa, px)
        view = pipeline._FaView(res)
        k = 0
        for f in range(F):
            D, A = planes[f, 0].astype(np.float32), planes[f, 1].astype(np.float32)
            with np.errstate(all="ignore"):
                want = port.fret_process_pair(D, A, polys[f], fret_p)
                wrows, wbg, _ = port.int_process_key({1: D.copy(), 2: A.copy()}, polys[f], None, task)
            assert np.array_equal(R[f], want["R_full"], equal_nan=True), ("R", f)
            for g, w in zip(rows_f[f], want["rows"]):
                assert g["roi"] == w["roi"] and g["area_px"] == w["area_px"]
                for kk in ("ratio_median", "ratio_p5", "ratio_p95", "donor_median", "yfret_median"):
                    assert g[kk] == w[kk] or (math.isnan(g[kk]) and math.isnan(w[kk])), (f, kk, g[kk], w[kk])
            assert res.int_bg[f, 0] == wbg[1]["bg"] and res.int_bg[f, 1] == wbg[2]["bg"], ("int bg", f)
            for g, w in zip(rows_i[f], wrows):
                for ch in (1, 2):
                    for kk in ("median", "p5", "p95", "vmin", "vmax", "npx"):
                        assert g[f"ch{ch}_{kk}"] == w[f"ch{ch}_{kk}"], (f, ch, kk)
                    assert close(g[f"ch{ch}_mean"], w[f"ch{ch}_mean"]) or abs(g[f"ch{ch}_mean"] - w[f"ch{ch}_mean"]) <= 1e-5 * max(abs(w[f"ch{ch}_vmax"]), abs(w[f"ch{ch}_vmin"]))
            img = planes[f, fa_ch].astype(np.float32)
            ref_stats = port.fa_global_stats(img)
            got = res.fa_stats[f]
            assert got[2] == ref_stats[2], ("fa bg", f)
            st = ref_stats
            if np.float32(got[3]) != ref_stats[0] + cfg["alpha"] * ref_stats[1]:
                stats["straddles"] += 1
                st = (np.float32(got[0]), np.float32(got[1]), ref_stats[2])
            for i, P in enumerate(polys[f]):
                crop, mask, rect = port.fa_crop_and_mask(img, P.copy())
                _, thr, bw, lab = port.analyze_fa_crop(crop, mask, cfg, st, with_contours=False)
                assert np.array_equal(view.bw_host(k), bw), ("bw", f, i)
                assert np.array_equal(view.labels_host(k), lab), ("labels", f, i)
                stats["adhesions"] += int(lab.max())
                k += 1
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, (F, H, W), "fa_ch", fa_ch, fa, type(e).__name__, str(e)[:200], [(t.filename.split("/")[-1], t.lineno) for t in tb][-2:], flush=True)
print("done", seed0, n, "bad", bad, stats, round(time.time() - t0, 1), flush=True)
