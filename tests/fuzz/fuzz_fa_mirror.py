"""tests/fuzz/fuzz_fa_mirror.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

The FA boundary function itself, host/FA_Analyzer.analyze_fa_crop(image_crop, roi_mask_crop, config, global_stats)
(reference FA_Analyzer.py:123-195), with ARBITRARY masks (not rasterised polygons), float32 crops as the reference's
loader hands them over, empty / one-row / one-column crops, against oracle.port.analyze_fa_crop: threshold value, binary
image, label map, the three lists of per-adhesion dicts incl. contours.

    python tests/fuzz/fuzz_fa_mirror.py <first seed> <number of seeds> [ref]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, traceback
from scipy import ndimage as ndi
from imageprocess_b200.ops import Engine
from imageprocess_b200.host import FA_Analyzer as mFA
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close
from oracle import port
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
if "ref" in sys.argv[3:]:            # expected side = the UNMODIFIED reference function (build container only) instead of the oracle port
    from oracle import refimport
    refimport.install_stubs()
    _ref = refimport.load("FA_Analyzer")
    expected = lambda crop, mask, cfg, stats: _ref.analyze_fa_crop(crop, mask, cfg, stats)
else:
    expected = lambda crop, mask, cfg, stats: port.analyze_fa_crop(crop, mask, cfg, stats, with_contours=True)
bad = 0; t0 = time.time()
for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    h = int(rng.choice([1, 2, 3, int(rng.integers(4, 90))])); w = int(rng.choice([1, 2, 31, 32, 33, 64, int(rng.integers(3, 150))]))
    img = rng.poisson(500, (h, w)).astype(np.int64)
    blobs = ndi.binary_dilation(rng.random((h, w)) < float(rng.choice([0.01, 0.05, 0.2])), iterations=int(rng.integers(1, 3))) if h > 2 and w > 2 \
        else rng.random((h, w)) < 0.4
    img[blobs] += int(rng.integers(400, 4000))
    crop = np.minimum(img, 65535).astype(np.float32)
    kind = int(rng.integers(0, 4))
    if kind == 0: mask = np.ones((h, w), bool)
    elif kind == 1: mask = rng.random((h, w)) < 0.7
    elif kind == 2: mask = np.zeros((h, w), bool); mask[h // 4:, : max(1, 3 * w // 4)] = True
    else: mask = np.zeros((h, w), bool)
    full = rng.poisson(500, (max(h, 40), max(w, 40))).astype(np.float32)
    stats = port.fa_global_stats(full)
    cfg = {"alpha": float(rng.choice([0.5, 1.0, 2.0, 4.0])), "min_px": float(rng.choice([0.0, 2.5, 6.0, 12.5])), "max_px": float(rng.choice([40.0, 400.0])),
           "close_radius": int(rng.integers(0, 6)), "subtract_bg": bool(rng.integers(0, 2))}
    try:
        try:
            want = expected(crop, mask, cfg, stats)
        except ValueError as e:                  # find_contours refuses crops smaller than 2 x 2 that hold a region
            try:
                mFA.analyze_fa_crop(crop, mask, cfg, stats, eng=eng)
                raise AssertionError(f"the oracle raises ValueError({e}), the mirror does not")
            except ValueError as e2:
                assert str(e2) == str(e), (str(e), str(e2))
            continue
        got = mFA.analyze_fa_crop(crop, mask, cfg, stats, eng=eng)
        assert np.float32(got[1]) == np.float32(want[1]), ("thr", got[1], want[1])
        assert got[2].dtype == want[2].dtype and np.array_equal(got[2], want[2]), "bw"
        assert np.array_equal(got[3], want[3]), "labels"
        for cat in ("OK", "Large", "Small"):
            assert len(got[0][cat]) == len(want[0][cat]), (cat, len(got[0][cat]), len(want[0][cat]))
            for g, wv in zip(got[0][cat], want[0][cat]):
                assert g["label"] == wv["label"] and g["area"] == wv["area"] and g["centroid"] == wv["centroid"], (cat, g["label"])
                assert type(g["area"]) is type(wv["area"]) and type(g["mean_int_raw"]) is type(wv["mean_int_raw"])
                assert close(float(g["mean_int_raw"]), float(wv["mean_int_raw"])) and g["bg_level"] == wv["bg_level"]
                assert np.array_equal(np.asarray(g["contour"]), np.asarray(wv["contour"])), ("contour", cat, g["label"])
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, (h, w), kind, cfg, type(e).__name__, str(e)[:200], [(t.filename.split("/")[-1], t.lineno) for t in tb][-3:], flush=True)
print("done", seed0, n, "bad", bad, round(time.time() - t0, 1), flush=True)
