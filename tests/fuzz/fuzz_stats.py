"""tests/fuzz/fuzz_stats.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

FRET + ROI-intensity in one FrameBatchJob against the oracle: adversarial value distributions (constant, two-valued, saturated, uniform, ramps), random ROIs (also off-frame), random scopes / strides / percentiles / hist-mode.  Found the narrow-ROI bug fixed in round 2 (tests/checks.py: check_narrow_rois).

    python tests/fuzz/fuzz_stats.py <first seed> <number of seeds> [big] [hist]     (prints one FAIL line per seed that differs)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, math
from imageprocess_b200.ops import Engine
from imageprocess_b200 import batch
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests import checks
from tests.checks import close, check_int_rows
from oracle import port
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2]); BIG = "big" in sys.argv[3:]; HIST = "hist" in sys.argv[3:]      # hist: hist-mode backgrounds in the FRET stage too
bad = 0; t0 = time.time(); stats = {"miss": 0, "fallback": 0}

def soft_close(a, b, scale):
    """1e-5 relative, or -- for sums that cancel (mean ~ 0 with clip_neg off) and for constant regions, where numpy's
    float32 pairwise sums carry the rounding noise of the VALUES -- 1e-5 of the values' scale.  Counted in stats."""
    if (math.isnan(a) and math.isnan(b)) or close(a, b):
        return True
    if abs(a - b) <= 1e-5 * scale:
        stats["scale_tol"] = stats.get("scale_tol", 0) + 1
        return True
    return False


def plane(rng, H, W):
    kind = int(rng.integers(0, 8))
    if kind == 0: return rng.poisson(float(rng.choice([3, 40, 900, 20000])), (H, W)).clip(0, 65535).astype(np.uint16)
    if kind == 1: return np.full((H, W), int(rng.integers(0, 65536)), np.uint16)
    if kind == 2: return rng.choice(np.array([int(rng.integers(0, 65536)), int(rng.integers(0, 65536))], np.uint16), (H, W))
    if kind == 3: return rng.integers(0, 65536, (H, W)).astype(np.uint16)
    if kind == 4:
        p = rng.poisson(500, (H, W)).astype(np.uint16); p[rng.random((H, W)) < 0.3] = 65535; return p
    if kind == 5: return (np.arange(H * W).reshape(H, W) % int(rng.integers(2, 5000))).astype(np.uint16)
    if kind == 6:
        p = rng.poisson(200, (H, W)).astype(np.uint16); p[: H // 2] += 30000; return p
    return (rng.poisson(100, (H, W)) * int(rng.integers(1, 600))).clip(0, 65535).astype(np.uint16)


def patchy(rng, H, W):
    """Blocks of different distributions: a ROI sees mixtures with heavy ties around its quantile ranks."""
    p = plane(rng, H, W)
    for _ in range(int(rng.integers(0, 4))):
        y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W)); y1, x1 = int(rng.integers(y0, H + 1)), int(rng.integers(x0, W + 1))
        p[y0:y1, x0:x1] = plane(rng, H, W)[y0:y1, x0:x1]
    return p

for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    if BIG:            # ROIs of thousands of pixels: the fused kernel's sampled-window path proper
        H = int(rng.integers(150, 300)); W = int(rng.choice([8 * int(rng.integers(20, 48)), int(rng.integers(150, 380))]))
    else:
        H = int(rng.integers(16, 120)); W = int(rng.choice([8 * int(rng.integers(3, 24)), int(rng.integers(17, 190))]))
    F = int(rng.integers(1, 3))
    planes = np.stack([np.stack([patchy(rng, H, W), patchy(rng, H, W)]) for _ in range(F)])
    polys = []
    for f in range(F):
        pl = []
        for k in range(int(rng.integers(0, 4))):
            nv = int(rng.integers(3, 8))
            P = np.stack([rng.uniform(-6, W + 6, nv), rng.uniform(-6, H + 6, nv)], axis=1)
            if rng.random() < 0.3: P = np.round(P * 2) / 2
            pl.append(P)
        if rng.random() < 0.3:
            pl.append(np.array([[0.0, 0.0], [W - 1.0, 0.0], [W - 1.0, H - 1.0], [0.0, H - 1.0]]))
        polys.append(pl)
    scope = str(rng.choice(["full", "roi_union"]))
    fret_p = {"bg_scope": scope, "bg_mode": str(rng.choice(["percentile", "percentile", "hist-mode"])) if HIST else "percentile", "percentile": float(rng.choice([0.0, 1.0, 5.0, 50.0, 99.0, 100.0])),
              "per_channel_p": bool(rng.integers(0, 2)), "donor_p": float(rng.choice([0.5, 1.0, 10.0])), "fret_p": float(rng.choice([1.0, 3.0, 30.0])),
              "clip_neg": bool(rng.integers(0, 2)), "eps_percentile": float(rng.choice([0.0, 1.0, 3.0, 50.0])),
              "ratio_mode": str(rng.choice(["Donor/FRET", "FRET/Donor"]))}
    task = {"bg_scope": str(rng.choice(["full", "roi_union"])), "bg_mode": str(rng.choice(["percentile", "percentile", "hist-mode"])),
            "clip_neg": bool(rng.integers(0, 2)), "bg_stride": int(rng.choice([1, 3, 4, 10])),
            "percentile": float(rng.choice([0.0, 1.0, 10.0, 50.0, 100.0])), "per_channel_p": False, "ch_p_map": {}}
    try:
        job = batch.FrameBatchJob(eng, planes.shape, stages=("fret", "int"), fret_p=fret_p, int_task=task, want_roi_image=True)
        if rng.random() < 0.7: job.pq_min_px = 0
        res = job.run(eng.mem.from_host(planes), polys)
        stats["miss"] += job.window_misses
        rows_i = batch.rows_intensity(res, F, [1, 2]); rows_f = batch.rows_fret(res, F)
        R = res.R.host()
        for f in range(F):
            D, A = planes[f, 0].astype(np.float32), planes[f, 1].astype(np.float32)
            with np.errstate(all="ignore"):
                want = port.fret_process_pair(D, A, polys[f], fret_p)
            assert np.array_equal(R[f], want["R_full"], equal_nan=True), ("R", f)
            assert len(rows_f[f]) == len(want["rows"])
            for g, w in zip(rows_f[f], want["rows"]):
                assert g["roi"] == w["roi"] and g["area_px"] == w["area_px"], (f, g["roi"])
                for k in ("ratio_median", "ratio_p5", "ratio_p95", "donor_median", "yfret_median"):
                    assert g[k] == w[k] or (math.isnan(g[k]) and math.isnan(w[k])), (f, k, g[k], w[k])
                for k in ("ratio_mean", "ratio_std", "donor_mean", "yfret_mean"):
                    pre = k.split("_")[0]
                    scale = abs(w[pre + "_median"]) + (abs(w["ratio_p5"]) + abs(w["ratio_p95"]) if pre == "ratio" else 0.0)
                    assert soft_close(g[k], w[k], scale), (f, k, g[k], w[k])
            if polys[f]:
                with np.errstate(all="ignore"):
                    wrows, wbg, _ = port.int_process_key({1: D.copy(), 2: A.copy()}, polys[f], None, task)
                assert res.int_bg[f, 0] == wbg[1]["bg"] and res.int_bg[f, 1] == wbg[2]["bg"], ("bg", f, res.int_bg[f], wbg)
                assert len(rows_i[f]) == len(wrows)
                for g, w in zip(rows_i[f], wrows):
                    assert g["roi"] == w["roi"] and g["area_px"] == w["area_px"]
                    for ch in (1, 2):
                        for k in ("median", "p5", "p95", "vmin", "vmax", "npx"):
                            a, b = g[f"ch{ch}_{k}"], w[f"ch{ch}_{k}"]
                            assert a == b or (math.isnan(a) and math.isnan(b)), (f, g["roi"], ch, k, a, b)
                        scale = max(abs(w[f"ch{ch}_vmin"]), abs(w[f"ch{ch}_vmax"])) if w[f"ch{ch}_npx"] else 0.0
                        for k in ("mean", "std"):
                            assert soft_close(g[f"ch{ch}_{k}"], w[f"ch{ch}_{k}"], scale), (f, g["roi"], ch, k, g[f"ch{ch}_{k}"], w[f"ch{ch}_{k}"])
                        assert soft_close(g[f"ch{ch}_vsum"], w[f"ch{ch}_vsum"], scale * max(w[f"ch{ch}_npx"], 1)), (f, g["roi"], ch, "vsum")
            else:
                assert rows_i[f] == []
    except Exception as e:
        if isinstance(e, ValueError) and "Too many bins" in str(e):
            # hist-mode on a (nearly) constant bright plane: np.histogram refuses in the reference as well;
            # both sides raise the same error (checked separately)
            stats["both_raise"] = stats.get("both_raise", 0) + 1
            continue
        bad += 1
        import traceback
        tb = traceback.extract_tb(e.__traceback__)[-1]
        print("FAIL seed", seed, H, W, F, fret_p, task, type(e).__name__, str(e)[:300], "at", tb.lineno, flush=True)
print("done", seed0, n, "bad", bad, stats, round(time.time() - t0, 1), flush=True)
