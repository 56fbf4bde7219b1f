"""tests/fuzz/fuzz_drawer.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

ROI-drawer assist segment_inside_polygon against the oracle.

    python tests/fuzz/fuzz_drawer.py <first seed> <number of seeds>     (prints one FAIL line per seed that differs)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, math, traceback
from imageprocess_b200.ops import Engine
from imageprocess_b200.host import roi_manual_drawer as rmd
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close
from oracle import port
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time()
for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(30, 140)), int(rng.integers(30, 180))
    img = rng.poisson(float(rng.choice([30, 300, 3000])), (H, W)).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for k in range(int(rng.integers(1, 4))):
        cx, cy, a, b = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(4, 50), rng.uniform(4, 40)
        cell = ((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0
        img[cell] += rng.poisson(float(rng.choice([200, 1500])), int(cell.sum())).astype(np.float32)
    if rng.random() < 0.4:
        y, x = int(rng.integers(0, H - 6)), int(rng.integers(0, W - 6)); img[y:y + 6, x:x + 9] = img.min()
    nv = int(rng.integers(3, 8))
    poly = np.stack([rng.uniform(-10, W + 10, nv), rng.uniform(-10, H + 10, nv)], axis=1)
    if rng.random() < 0.3: poly = np.round(poly * 2) / 2
    mode, par = [("percentile", 60.0), ("percentile", 90.0), ("bnd", 0.5), ("percentile", 5.0), ("bnd", 0.9)][int(rng.integers(0, 5))]
    min_area, tol = int(rng.choice([1, 10, 40, 400])), float(rng.choice([0.0, 0.5, 1.0, 3.0]))
    try:
        with np.errstate(all="ignore"):
            w = port.segment_inside_polygon(img, poly, thr_param=par, min_area=min_area, tolerance=tol, mode=mode)
    except Exception as e:
        try:
            rmd.segment_inside_polygon(img, poly, thr_param=par, min_area=min_area, tolerance=tol, mode=mode, eng=eng)
            bad += 1; print("FAIL seed", seed, "oracle raises", type(e).__name__, "ours does not", flush=True)
        except Exception as e2:
            if type(e2) is not type(e): bad += 1; print("FAIL seed", seed, "different exceptions", type(e).__name__, type(e2).__name__, str(e2)[:100], flush=True)
        continue
    try:
        g = rmd.segment_inside_polygon(img, poly, thr_param=par, min_area=min_area, tolerance=tol, mode=mode, eng=eng)
        assert (g[0] is None) == (w[0] is None), ("thr none", g[0], w[0])
        if w[0] is not None:
            assert close(g[0], w[0], 1e-6), ("thr", g[0], w[0])
        assert (g[2] is None) == (w[2] is None), ("poly none", g[2] is None, w[2] is None)
        if w[2] is not None:
            assert g[2].shape == w[2].shape and np.array_equal(g[2], w[2]), ("poly", g[2].shape, w[2].shape)
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, (H, W), mode, par, min_area, tol, type(e).__name__, str(e)[:200], [(t.filename.split("/")[-1], t.lineno) for t in tb][-2:], flush=True)
print("done", seed0, n, "bad", bad, round(time.time() - t0, 1), flush=True)
