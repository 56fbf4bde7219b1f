"""tests/fuzz/fuzz_pipeline.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

The host pipeline of a persistent FrameBatchJob: random sequences of submit() / collect() with up to two tickets in
flight, six ROI layouts (more than the four cached plans: evictions), two input buffers, frames without ROIs, layouts
that share ROI lists between frames.  Every collected step must equal what a FRESH job computes for the same inputs
(tables byte for byte, the ratio image bit for bit); the cached plans, output slots, staged downloads and repeat
logic are what is under test, not the kernels.

    python tests/fuzz/fuzz_pipeline.py <first seed> <number of seeds>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, traceback
from imageprocess_b200.ops import Engine
from imageprocess_b200 import batch
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import small_scene, FA_CASES
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time()
FRET_P = {"bg_scope": "full", "bg_mode": "percentile", "percentile": 1.0, "per_channel_p": False, "donor_p": 1.0, "fret_p": 1.0,
          "clip_neg": True, "eps_percentile": 1.0, "ratio_mode": "FRET/Donor"}
TASK = {"bg_scope": "full", "bg_mode": "percentile", "clip_neg": True, "bg_stride": 4, "percentile": 1.0, "per_channel_p": False,
        "ch_p_map": {}}


def snap(res, with_R=True):
    out = [res.fret_params.copy(), res.int_bg.copy(), res.fa_stats.copy(), res.fret_stat.copy(), res.int_stat.copy(),
           res.fa_comp_off.copy(), res.fa_comps.copy(), np.asarray(res.area).copy()]
    if with_R:
        out.append(res.R.host().copy())
    return out


def same(x, y):
    return len(x) == len(y) and all(np.asarray(p).shape == np.asarray(q).shape and
                                    np.array_equal(np.ascontiguousarray(p).view(np.uint8), np.ascontiguousarray(q).view(np.uint8))
                                    for p, q in zip(x, y))


for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    F, H, W = 3, 96, 128
    scenes = [small_scene(int(rng.integers(0, 1000)), H=H, W=W, n_cells=2, blobs=5) for _ in range(4)]
    bufs_np = [np.stack([np.stack([scenes[int(rng.integers(0, 4))][k] for k in (0, 1)]) for _ in range(F)]) for _ in range(2)]
    roi_sets = [sc[2] for sc in scenes] + [[], [scenes[0][2][0]]]
    layouts = []
    for _ in range(6):
        kind = int(rng.integers(0, 3))
        if kind == 0:
            L = [roi_sets[int(rng.integers(0, 6))]] * F                       # one list shared by all frames
        else:
            L = [roi_sets[int(rng.integers(0, 6))] for _ in range(F)]
        layouts.append(L)
    try:
        mk = lambda: batch.FrameBatchJob(eng, bufs_np[0].shape, stages=("fret", "int", "fa"), fret_p=FRET_P, int_task=TASK,
                                         fa_params=FA_CASES[0], fa_px=0.112)
        job = mk()
        job.pq_min_px = 0
        bufs = [eng.mem.from_host(b) for b in bufs_np]
        want_cache = {}
        pending = []
        n_steps = int(rng.integers(6, 16))
        for step in range(n_steps + 1):
            while pending and (step == n_steps or len(pending) > int(rng.integers(0, 3))):
                tk, key = pending.pop(0)
                res = job.collect(tk)
                got = snap(res, with_R=not pending)             # device images are per job, not per slot: only the newest step's is current
                if key not in want_cache:
                    fresh = mk(); fresh.pq_min_px = 0
                    want_cache[key] = snap(fresh.run(eng.mem.from_host(bufs_np[key[0]]), layouts[key[1]]))
                want = want_cache[key] if not pending else want_cache[key][:-1]
                assert same(got, want), ("step result differs from a fresh job", step, key)
            if step == n_steps:
                break
            b, l = int(rng.integers(0, 2)), int(rng.integers(0, 6))
            if rng.random() < 0.15:                              # new pixels in a buffer between steps
                bufs_np[b] = np.roll(bufs_np[b], int(rng.integers(1, 9)), axis=3).copy()
                eng.mem.upload_async(bufs[b], bufs_np[b].view(np.uint8).reshape(-1))
                want_cache = {k: v for k, v in want_cache.items() if k[0] != b}
                if any(key[0] == b for _, key in pending):       # steps in flight read the old pixels: drain first is the caller's job
                    raise RuntimeError("test bug: buffer rewritten with steps in flight")
            pending.append((job.submit(bufs[b], layouts[l]), (b, l)))
    except RuntimeError as e:
        if "test bug" in str(e):
            continue
        bad += 1; print("FAIL seed", seed, type(e).__name__, str(e)[:200], flush=True)
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, type(e).__name__, str(e)[:200], [(t.filename.split("/")[-1], t.lineno) for t in tb][-3:], flush=True)
print("done", seed0, n, "bad", bad, round(time.time() - t0, 1), flush=True)
