"""tests/fuzz/fuzz_filters.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

Gaussian filter (ROI drawer display pipeline, optional FA pre-filter) and grey erosion / dilation / white top-hat
(optional FA pre-filter) against scipy.ndimage, bit for bit, on random shapes (tile-aligned and odd), radii from
below one pixel to larger than the image, Otsu's threshold against the restated skimage rule.

    python tests/fuzz/fuzz_filters.py <first seed> <number of seeds>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, traceback
from scipy import ndimage as ndi
from imageprocess_b200.ops import Engine
from imageprocess_b200 import filters
from tests.emu.emu_backend import NumpyMem, emu_lib
from oracle import shims
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time()
for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    H = int(rng.integers(3, 150)); W = int(rng.choice([4 * int(rng.integers(1, 70)), int(rng.integers(3, 300)), 128, 256]))
    try:
        img = (rng.poisson(float(rng.choice([5, 400, 30000])), (H, W)) + 50 * np.sin(np.arange(W) / 7.0)[None, :]).astype(np.float32)
        sg = float(rng.choice([0.3, 0.4, 1.0, 1.2, 2.0, 5.5, 9.0, 33.0, rng.uniform(0.2, 20.0)]))
        got = filters.gaussian_filter(eng, eng.mem.from_host(img), sg).host()
        want = ndi.gaussian_filter(img, sg)
        assert got.dtype == want.dtype and np.array_equal(got, want), ("gauss", sg, float(np.abs(got - want).max()))
        u = rng.poisson(float(rng.choice([5, 500, 20000])), (2, H, W)).clip(0, 65535).astype(np.uint16)
        if rng.random() < 0.3: u[:, H // 3: H // 3 + 4, W // 4: W // 4 + 6] = 65535
        size = int(rng.choice([1, 3, 5, 9, 15, 31, 129, 2 * int(rng.integers(1, 40)) + 1]))
        d = eng.mem.from_host(u)
        assert np.array_equal(filters.grey_morph(eng, d, size, False).host()[0], ndi.grey_erosion(u[0], size=(size, size))), ("erosion", size)
        assert np.array_equal(filters.grey_morph(eng, d, size, True).host()[1], ndi.grey_dilation(u[1], size=(size, size))), ("dilation", size)
        th = filters.white_tophat(eng, d, size).host()
        for k in range(2):
            assert np.array_equal(th[k], ndi.white_tophat(u[k], size=(size, size))), ("tophat", size, k)
        got = filters.threshold_otsu(eng, d, H, W, [0, 1])
        assert got == [int(shims.threshold_otsu(p)) for p in u], ("otsu", got)
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, (H, W), type(e).__name__, str(e)[:200], [(t.filename.split("/")[-1], t.lineno) for t in tb][-2:], flush=True)
print("done", seed0, n, "bad", bad, round(time.time() - t0, 1), flush=True)
