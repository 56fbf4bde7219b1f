"""tests/fuzz/fuzz_fa.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

FA chain (threshold, size filter, closing, labelling, per-adhesion rows, outlines) on all three device paths against oracle.port.analyze_fa_crop: random crop sizes (widths at word boundaries), noise densities, ROI shapes, parameters.

    python tests/fuzz/fuzz_fa.py <first seed> <number of seeds> [big]     (prints one FAIL line per seed that differs)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, traceback
from imageprocess_b200.ops import Engine
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests import checks
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2]); BIG = len(sys.argv) > 3 and sys.argv[3] == "big"
bad = 0
t0 = time.time()
for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    if BIG:        # crops beyond the shared-memory kernel's limits (4096 words, 8192 runs, 512 adhesions): flagged, finished in global memory
        H = int(rng.integers(150, 420)); W = int(rng.choice([int(rng.integers(200, 560)), 256, 512, 257, 511]))
    else:
        H = int(rng.integers(8, 90)); W = int(rng.choice([int(rng.integers(8, 140)), 32, 64, 96, 33, 65, 31, 63, 128]))
    dens = float(rng.choice([0.02, 0.1, 0.3, 0.5, 0.7]))
    d = rng.poisson(500, (H, W)).astype(np.int64)
    m = rng.random((H, W)) < dens
    if rng.random() < 0.5:       # clumpier
        from scipy import ndimage as ndi
        m = ndi.binary_dilation(rng.random((H, W)) < dens / 4, iterations=int(rng.integers(1, 3)))
    d[m] += int(rng.integers(300, 3000))
    d = np.minimum(d, 65535).astype(np.uint16)
    a = rng.poisson(300, (H, W)).astype(np.uint16)
    polys = []
    for k in range(int(rng.integers(1, 4))):
        kind = int(rng.integers(0, 4))
        if kind == 0:   # whole frame, on pixel centres
            P = np.array([[0.0, 0.0], [W - 1.0, 0.0], [W - 1.0, H - 1.0], [0.0, H - 1.0]])
        elif kind == 1: # random rectangle with word-aligned or odd edges
            x0 = float(rng.integers(0, max(1, W - 4))); x1 = float(rng.integers(int(x0) + 2, W + 6))
            y0 = float(rng.integers(0, max(1, H - 4))); y1 = float(rng.integers(int(y0) + 2, H + 6))
            P = np.array([[x0 + 0.5, y0 + 0.5], [x1 + 0.5, y0 + 0.5], [x1 + 0.5, y1 + 0.5], [x0 + 0.5, y1 + 0.5]])
        elif kind == 2: # random polygon
            nv = int(rng.integers(3, 9))
            P = np.stack([rng.uniform(-4, W + 4, nv), rng.uniform(-4, H + 4, nv)], axis=1)
            if P[:, 0].max() < 1 or P[:, 1].max() < 1 or P[:, 0].min() > W - 2 or P[:, 1].min() > H - 2: continue
        else:           # thin sliver
            x0 = float(rng.integers(0, W - 2)); y0 = float(rng.integers(0, H - 2))
            P = np.array([[x0, y0], [x0 + float(rng.integers(1, 4)), y0], [x0 + 1.0, y0 + float(rng.integers(1, min(H, 40)))]])
        polys.append(P)
    if not polys: continue
    params = {"alpha": float(rng.choice([0.5, 1.0, 2.0])), "min_area_um": float(rng.choice([0.0, 2.5, 6.0, 12.5])) * 0.112 ** 2,
              "max_area_um": float(rng.choice([40.0, 400.0])) * 0.112 ** 2, "close_radius": int(rng.integers(0, 6)), "subtract_bg": bool(rng.integers(0, 2))}
    for path in (1, 2, 3):
        try:
            checks.check_fa_batch(eng, params, fa_path=path, frames=[(d, a, polys)], contour_stride=97 if BIG else 7)
        except Exception as e:
            bad += 1
            print("FAIL seed", seed, "path", path, H, W, params, type(e).__name__, str(e)[:200], flush=True)
print("done", seed0, n, "bad", bad, round(time.time() - t0, 1), flush=True)
