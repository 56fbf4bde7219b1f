"""tests/fuzz/fuzz_int_mask.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

Fluor_INT's union-mask / whole-frame branch (host/Fluor_INT._quantify_mask; reference Fluor_INT.py:522-538: a PNG
mask instead of ROI polygons, or no ROI at all) against oracle.port.int_process_key with ARBITRARY masks, adversarial
planes, every background mode / scope / stride.

    python tests/fuzz/fuzz_int_mask.py <first seed> <number of seeds>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, math, traceback
from imageprocess_b200.ops import Engine
from imageprocess_b200.host import Fluor_INT as mF
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close
from oracle import port
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time(); stats = {}


def plane(rng, H, W):
    kind = int(rng.integers(0, 6))
    if kind == 0: return rng.poisson(float(rng.choice([3, 40, 900, 20000])), (H, W)).clip(0, 65535).astype(np.uint16)
    if kind == 1: return np.full((H, W), int(rng.integers(0, 4000)), np.uint16)
    if kind == 2: return rng.choice(np.array([int(rng.integers(0, 65536)), int(rng.integers(0, 65536))], np.uint16), (H, W))
    if kind == 3: return rng.integers(0, 65536, (H, W)).astype(np.uint16)
    if kind == 4:
        p = rng.poisson(500, (H, W)).astype(np.uint16); p[rng.random((H, W)) < 0.3] = 65535; return p
    return (np.arange(H * W).reshape(H, W) % int(rng.integers(2, 5000))).astype(np.uint16)


for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(8, 130)), int(rng.choice([8 * int(rng.integers(1, 24)), int(rng.integers(5, 190))]))
    planes = np.stack([plane(rng, H, W), plane(rng, H, W)])[None]
    kind = int(rng.integers(0, 5))
    if kind == 0: mask = None
    elif kind == 1: mask = rng.random((H, W)) < float(rng.choice([0.02, 0.5, 0.95]))
    elif kind == 2: mask = np.zeros((H, W), bool); mask[H // 3:, : max(1, W // 2)] = True
    elif kind == 3: mask = np.zeros((H, W), bool); mask[int(rng.integers(0, H)), int(rng.integers(0, W))] = True
    else: mask = np.zeros((H, W), bool)
    task = {"bg_scope": str(rng.choice(["full", "roi_union"])), "bg_mode": str(rng.choice(["percentile", "percentile", "hist-mode", "none"])),
            "clip_neg": bool(rng.integers(0, 2)), "bg_stride": int(rng.choice([1, 2, 3, 4, 10])), "percentile": float(rng.choice([0.0, 1.0, 10.0, 50.0, 100.0])),
            "per_channel_p": bool(rng.integers(0, 2)), "ch_p_map": {3: 7.5}}
    raw = {2: planes[0, 0].astype(np.float32), 3: planes[0, 1].astype(np.float32)}
    try:
        with np.errstate(all="ignore"):
            want = port.int_process_key({k: v.copy() for k, v in raw.items()}, None, mask, task)
    except Exception as e:
        try:
            mF._quantify_mask(eng, eng.mem.from_host(planes), planes.shape, [2, 3], task, mask, 1)
            bad += 1; print("FAIL seed", seed, "oracle raises", type(e).__name__, str(e)[:80], "ours does not", flush=True)
        except Exception as e2:
            if type(e2) is not type(e): bad += 1; print("FAIL seed", seed, "different exceptions", type(e).__name__, type(e2).__name__, str(e2)[:100], flush=True)
            else: stats["both_raise"] = stats.get("both_raise", 0) + 1
        continue
    try:
        rows, bg = mF._quantify_mask(eng, eng.mem.from_host(planes), planes.shape, [2, 3], task, mask, 1)
        wrows, wbg = want[0], want[1]
        assert len(rows) == len(wrows), ("rows", len(rows), len(wrows))
        for ch in (2, 3):
            assert bg[ch]["bg"] == wbg[ch]["bg"] and bg[ch]["p"] == wbg[ch]["p"], ("bg", ch, bg[ch], wbg[ch])
        for g, w in zip(rows, wrows):
            assert g["area_px"] == w["area_px"], ("area", g["area_px"], w["area_px"])
            for ch in (2, 3):
                for k in ("median", "p5", "p95", "vmin", "vmax", "npx"):
                    a, b = g[f"ch{ch}_{k}"], w[f"ch{ch}_{k}"]
                    assert a == b or (math.isnan(a) and math.isnan(b)), (ch, k, a, b)
                scale = max(abs(w[f"ch{ch}_vmin"]), abs(w[f"ch{ch}_vmax"])) if w[f"ch{ch}_npx"] else 0.0
                for k in ("mean", "std"):
                    a, b = g[f"ch{ch}_{k}"], w[f"ch{ch}_{k}"]
                    assert (math.isnan(a) and math.isnan(b)) or close(a, b) or abs(a - b) <= 1e-5 * scale, (ch, k, a, b)
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, (H, W), "mask kind", kind, task, type(e).__name__, str(e)[:200], [(t.filename.split("/")[-1], t.lineno) for t in tb][-2:], flush=True)
print("done", seed0, n, "bad", bad, stats, round(time.time() - t0, 1), flush=True)
