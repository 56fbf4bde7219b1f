"""BASELINE configs C2 and C3 (spectral) on the GPU at full size.

C2  2FocalAdhesion.bat: the reference ships the ROI JSONs of its FA sample (2200 x 3200 images,
    2-5 cell outlines of 62-540 vertices each) but not the images, so synthetic
    images are painted under EVERY shipped polygon and go through the batched FA chain; every
    crop is compared with the oracle's analyze_fa_crop: crop rect, skimage mask, binary image and
    label map bit-exact, per-adhesion rows in the reference's order.
C3  3FRET.bat, Nesprin2 builder with use_spectral=True, alpha=0.12, beta=0.05, g_factor=1.1 on a
    2048 x 2048 donor / FRET / acceptor-only triple with 24 ROIs: ratio images bit-exact incl. the
    NaN pattern, EDT rim mask exact, per-ROI rows; a second case adds the annulus background.
"""
import json
import math
import os

import numpy as np
import pytest

from imageprocess_b200 import nesprin2, pipeline, synth
from imageprocess_b200.nesprin2 import bits_to_bool
from oracle import port
from tests import checks, goldenio

pytestmark = pytest.mark.gpu

PX = 0.112
FA_PARAMS = {"alpha": 2.0, "min_area_um": 1.5, "max_area_um": 30.0, "close_radius": 1, "subtract_bg": True}


@pytest.fixture(scope="module")
def eng():
    import imageprocess_b200 as ipb
    return ipb.engine("cuda:0")


def test_c2_all_fixture_polygons_through_fa_batch(eng):
    rois = json.load(open(os.path.join(goldenio.GOLD, "fa_rois.json")))
    names = sorted(rois)
    H, W = rois[names[0]]["image_shape"]["height"], rois[names[0]]["image_shape"]["width"]
    polys_pf = [[np.asarray(P, dtype=float) for P in rois[n]["rois"]] for n in names]
    info = {}
    frames = np.stack([synth.fa_cells_frame(100 + k, H, W, polys, blobs_per_cell=40, info=info)
                       for k, polys in enumerate(polys_pf)])[:, None]
    assert frames.shape == (4, 1, 2200, 3200) and sum(len(p) for p in polys_pf) == 16
    out = pipeline.fa_batch(eng, eng.mem.from_host(frames), frames.shape, polys_pf, FA_PARAMS, PX, channel=0,
                            save_ok_only=False, want_labels=True)
    cfg = pipeline.fa_um_to_px_config(FA_PARAMS, PX)
    k = n_fa = straddles = 0
    for f, polys in enumerate(polys_pf):
        img = frames[f, 0].astype(np.float32)
        ref_stats = port.fa_global_stats(img)
        got = out["stats"][f]
        assert got[2] == ref_stats[2]
        assert checks.close(float(got[0]), float(ref_stats[0]), 1e-6) and checks.close(float(got[1]), float(ref_stats[1]), 1e-6)
        stats = ref_stats
        if np.float32(got[3]) != ref_stats[0] + cfg["alpha"] * ref_stats[1]:
            straddles += 1
            stats = (np.float32(got[0]), np.float32(got[1]), ref_stats[2])
        want_rows = port.fa_batch_rows(img, polys, FA_PARAMS, PX, s_tag="S", save_ok_only=False, with_contours=False,
                                       stats=stats)
        for i, P in enumerate(polys):
            crop, mask, rect = port.fa_crop_and_mask(img, P.copy())
            assert out["rects"][k] == rect and out["owner"][k] == (f, i + 1)
            _, thr, bw, lab = port.analyze_fa_crop(crop, mask, cfg, stats, with_contours=False)
            assert np.array_equal(out["result"].bw_host(k), bw), (f, i)
            assert np.array_equal(out["result"].labels_host(k), lab), (f, i)
            n_fa += int(lab.max())
            k += 1
        got_rows = out["rows_per_frame"][f]
        assert len(got_rows) == len(want_rows)
        for g, w in zip(got_rows, want_rows):
            assert g["Cell_ID"] == w["Cell_ID"] and g["Category"] == w["Category"] and g["Area_px"] == w["Area_px"]
            assert checks.close(float(g["Mean_Intensity_Raw"]), float(w["Mean_Intensity_Raw"]))
            assert g["Background_Level"] == w["Background_Level"]
    assert k == 16 and n_fa >= 200, n_fa
    print(f"C2: 16 cell crops, {n_fa} adhesions labelled identically ({info['blobs_placed']} blobs painted), "
          f"{straddles} of 4 frames with a float32 threshold one ulp from numpy's")


N2 = dict(checks.N2_BASE, use_spectral=True, alpha=0.12, beta=0.05, g_factor=1.1)


@pytest.mark.parametrize("case", [{"n_rois": 24}, {"n_rois": 6, "annulus_on": True}], ids=["rim24", "annulus6"])
def test_c3_spectral_full_size(eng, case):
    H = W = 2048
    d, a, polys = synth.fret_frame(seed=1234, H=H, W=W, n_cells=24, r_min=80, r_max=160)
    polys = polys[: case["n_rois"]]
    rng = np.random.default_rng(5)
    ao = (0.3 * a + rng.poisson(50, d.shape)).astype(np.uint16)
    planes = np.stack([d, a, ao])[None]
    p = dict(N2, **{k: v for k, v in case.items() if k != "n_rois"})
    out = nesprin2.nesprin2_batch(eng, eng.mem.from_host(planes), planes.shape, [polys], p, donor_ch=0, acc_ch=1, aonly_ch=2)
    imgs = out["images"].host()
    rim = bits_to_bool(out["rim"].host().reshape(1, H, (W + 31) // 32), H, W)
    want = port.n2_process_pair(d.astype(np.float32), a.astype(np.float32), polys, p, Aonly=ao.astype(np.float32))
    assert np.float32(out["eps"][0]) == np.float32(want["eps"])
    for k, name in enumerate(("R_full", "R_alt", "Dcorr", "Acorr")):
        assert np.array_equal(imgs[k, 0], want[name], equal_nan=True), name
    assert np.array_equal(rim[0], want["rim_mask"])
    assert len(out["rows_per_frame"][0]) == len(want["rows"]) == len(polys)
    for g, w in zip(out["rows_per_frame"][0], want["rows"]):
        assert g["roi"] == w["roi"] and g["area_px"] == w["area_px"]
        for key in ("ratio_median", "ratio_p5", "ratio_p95"):
            assert g[key] == w[key] or (math.isnan(g[key]) and math.isnan(w[key])), (key, g[key], w[key])
        for key in ("ratio_mean", "ratio_std", "ratio_FoverD_mean", "ratio_DoverF_mean", "donor_mean", "fret_mean"):
            assert checks.close(g[key], w[key]), (key, g[key], w[key])
