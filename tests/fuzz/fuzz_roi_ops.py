"""tests/fuzz/fuzz_roi_ops.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

MOR moments, 16-bit previews and the ROI cropper against the oracle on random polygons / images.

    python tests/fuzz/fuzz_roi_ops.py <first seed> <number of seeds>     (prints one FAIL line per seed that differs)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time, math
from imageprocess_b200.ops import Engine
from imageprocess_b200 import roi_ops
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close
from oracle import port
eng = Engine(emu_lib(), NumpyMem())
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time()
def rpoly(rng, H, W):
    kind = int(rng.integers(0, 4))
    if kind == 0:
        nv = int(rng.integers(3, 9)); return np.stack([rng.uniform(-6, W + 6, nv), rng.uniform(-6, H + 6, nv)], axis=1)
    if kind == 1:
        x0, y0 = float(rng.integers(0, W - 3)), float(rng.integers(0, H - 3)); w, h = float(rng.integers(1, 9)), float(rng.integers(1, 30))
        return np.array([[x0 - .5, y0 - .5], [x0 + w - .5, y0 - .5], [x0 + w - .5, y0 + h - .5], [x0 - .5, y0 + h - .5]])
    if kind == 2:
        cx, cy, r = rng.uniform(10, W - 10), rng.uniform(10, H - 10), rng.uniform(3, 30)
        t = np.sort(rng.uniform(0, 2 * np.pi, int(rng.integers(5, 14)))); return np.stack([cx + r * np.cos(t), cy + r * np.sin(t)], axis=1)
    return np.round(np.stack([rng.uniform(0, W, 5), rng.uniform(0, H, 5)], axis=1))
for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(24, 130)), int(rng.integers(24, 170))
    polys = [rpoly(rng, H, W) for _ in range(int(rng.integers(1, 5)))]
    try:
        # a17 MOR
        px_um = float(rng.choice([0.112, 1.0, 0.223]))
        try:
            got = roi_ops.morphology_batch(eng, polys, (H, W), px_um)
        except TypeError:
            raised = False
            for P in polys:
                try:
                    with np.errstate(all="ignore"): port.morphology_from_polygon(P, (H, W), px_um)
                except TypeError:
                    raised = True
            assert raised, "ours raises TypeError, the oracle does not"
            got = []
        for P, g in zip(polys if got else [], got):
            with np.errstate(all="ignore"):
                w = port.morphology_from_polygon(P, (H, W), px_um)
            assert g.keys() == w.keys()
            assert g["area_px"] == w["area_px"], ("area", g["area_px"], w["area_px"])
            for k in w:
                if k == "area_px": continue
                gv, wv = float(g[k]), float(w[k])
                if math.isnan(wv) or math.isnan(gv):
                    assert math.isnan(wv) and math.isnan(gv), ("nan", k, gv, wv); continue
                if k == "orientation_deg":
                    dd = abs(gv - wv) % 180.0
                    if min(dd, 180.0 - dd) >= 1e-6:
                        # isotropic regions: the eigenvectors are arbitrary
                        assert abs(float(w["major_axis_um"]) - float(w["minor_axis_um"])) <= 1e-9 * max(1.0, float(w["major_axis_um"])), (k, gv, wv)
                else:
                    assert close(gv, wv, 1e-9) or abs(gv - wv) < 1e-9, (k, gv, wv, P.tolist())
        # a15 preview + a16 cropper
        d = rng.poisson(float(rng.choice([50, 800, 20000])), (H, W)).clip(0, 65535).astype(np.uint16)
        if rng.random() < 0.2: d[:] = int(rng.integers(0, 65536))
        img = d.astype(np.float32) - np.float32(rng.choice([0.0, 97.0, 700.5]))
        if rng.random() < 0.5: img[img < 0] = 0
        R = (d.astype(np.float32) + 5) / (rng.poisson(300, (H, W)).astype(np.float32) + 5)
        R[rng.random(R.shape) < float(rng.choice([0.0, 0.01, 0.6]))] = np.nan
        lo, hi = float(rng.choice([0.0, 1.0, 5.0])), float(rng.choice([95.0, 99.0, 100.0]))
        got = roi_ops.preview_u16_batch(eng, np.stack([img, R]), lo, hi)
        for k, src in enumerate((img, R)):
            with np.errstate(all="ignore"):
                want = port.preview_u16(src, lo, hi)
            assert np.array_equal(got[k], want), ("preview", k, lo, hi)
        gamma, low, high, mo = float(rng.choice([1.0, 1.0, 2.2, 0.7])), float(rng.choice([0.0, 0.5, 1.0])), float(rng.choice([0.0, 1.0, 2.0])), bool(rng.integers(0, 2))
        outs = roi_ops.cropper_batch(eng, d, polys, low, high, gamma, mask_outside=mo)
        for P, g in zip(polys, outs):
            with np.errstate(all="ignore"):
                w = port.cropper_normalize(d.astype(np.float32), d, P, low, high, gamma, mask_outside=mo)
            assert (g is None) == (w is None), ("none", g is None, w is None, P.tolist())
            if w is None: continue
            assert g["rect"] == w["rect"], ("rect", g["rect"], w["rect"])
            assert np.array_equal(g["mask"], w["mask"]), "mask"
            assert np.array_equal(g["raw_out"], w["raw_out"]), "raw_out"
            assert np.allclose(g["norm_gamma"], w["norm_gamma"], rtol=2e-6, atol=1e-7), "norm_gamma"
            diff = np.abs(g["out16"].astype(np.int64) - w["out16"].astype(np.int64))
            assert diff.max() <= (0 if gamma == 1.0 else 1), ("out16", gamma, int(diff.max()))
    except Exception as e:
        bad += 1
        import traceback
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, (H, W), type(e).__name__, str(e)[:300], [(t.filename.split("/")[-1], t.lineno) for t in tb][-3:], flush=True)
print("done", seed0, n, "bad", bad, round(time.time() - t0, 1), flush=True)
