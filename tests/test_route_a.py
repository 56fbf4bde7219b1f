"""INTEGRATION.md route A executed: the UNMODIFIED reference modules (loaded from /root/reference
through oracle/refimport.py) get their worker functions rebound to the mirrors exactly as the
maintainer's snippet does, and the rebound module-level function is called with the reference's
own argument shapes.  Its result is compared with what the reference's ORIGINAL function returns
for the same call (the reference running on the restated third-party shims).  Build container
only (skipped where /root/reference is absent); the kernels run in the emulated build."""
import json
import math
import os

import numpy as np
import pytest

from oracle import refimport
from oracle.gen_golden import small_scene
from tests import goldenio
from tests.checks import close

pytestmark = pytest.mark.skipif(not refimport.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def emu_engine():
    import imageprocess_b200 as ipb
    from imageprocess_b200.ops import Engine
    from tests.emu.emu_backend import NumpyMem, emu_lib
    saved = ipb._engine
    ipb._engine = Engine(emu_lib(), NumpyMem())      # what ipb.engine() hands to the mirrors
    yield ipb._engine
    ipb._engine = saved


def _rebind(ref_mod, mirror, names):
    """The INTEGRATION.md route-A snippet: `name = _ipb.name` inside the reference module."""
    orig = {n: getattr(ref_mod, n) for n in names}
    for n in names:
        setattr(ref_mod, n, getattr(mirror, n))
    return orig


def _restore(ref_mod, orig):
    for n, f in orig.items():
        setattr(ref_mod, n, f)


def _write_rois(path, polys, shape):
    with open(path, "w") as f:
        json.dump({"name": "S01", "image_shape": {"height": shape[0], "width": shape[1]},
                   "rois": [np.asarray(P).tolist() for P in polys]}, f)


def test_fluor_int_worker_rebound(emu_engine, tmp_path):
    from imageprocess_b200.host import Fluor_INT as mirror, common
    ref = refimport.load("Fluor_INT")
    imgs, polys, rows, _ = goldenio.load_intensity("e1_P0")
    img_dir = str(tmp_path)
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    for ch in (2, 3):
        common.write_tiff(os.path.join(img_dir, f"S01_{ch}.TIF"), imgs[ch])
    _write_rois(os.path.join(roi_dir, "S01.json"), polys, imgs[2].shape)
    tasks, _ = mirror.build_tasks(img_dir, roi_dir, os.path.join(img_dir, "RES"),
                                  {"channels_to_quant": [2, 3], "ch_color_map": {2: "Green", 3: "Red"}})
    task = dict(tasks[0])
    task.update({"px_um": None, "lang": "en"})                    # keys the reference's task dict carries
    want = ref._process_key_task(dict(task))                     # the reference's own worker (numpy + shims)
    orig = _rebind(ref, mirror, ["_process_key_task", "rasterize_polygon"])
    try:
        got = ref._process_key_task(dict(task))                  # the rebound module-level name
        m = ref.rasterize_polygon(polys[0], imgs[2].shape)
    finally:
        _restore(ref, orig)
    assert np.array_equal(m, orig["rasterize_polygon"](polys[0], imgs[2].shape))
    assert want["rows"] and len(got["rows"]) == len(want["rows"]) == len(rows) and got["steps"] == want["steps"]
    for g, w in zip(got["rows"], want["rows"]):
        assert set(g) == set(w), set(g) ^ set(w)
        for k, wv in w.items():
            gv = g[k]
            if isinstance(wv, float) and k.endswith(("_mean", "_std", "_vsum")):
                assert close(gv, wv), (k, gv, wv)
            elif isinstance(wv, float) and math.isnan(wv):
                assert math.isnan(gv), k
            else:
                assert gv == wv, (k, gv, wv)


def test_fa_analyze_crop_rebound(emu_engine):
    from imageprocess_b200.host import FA_Analyzer as mirror
    ref = refimport.load("FA_Analyzer")
    d, _, polys = small_scene(21, H=120, W=168, n_cells=2, blobs=10)
    img = d.astype(np.float32)
    stats = (np.nanmean(img), np.nanstd(img), np.percentile(img[::10, ::10], 1.0))     # FA_Analyzer.py:984-987
    cfg = {"alpha": 2.0, "min_px": 12.5, "max_px": 300.0, "close_radius": 1, "subtract_bg": True}
    from oracle import port
    crops = [port.fa_crop_and_mask(img, P.copy())[:2] for P in polys]
    want = [ref.analyze_fa_crop(c, m, cfg, stats) for c, m in crops]
    orig = _rebind(ref, mirror, ["analyze_fa_crop", "load_image_safe"])
    try:
        got = [ref.analyze_fa_crop(c, m, cfg, stats) for c, m in crops]
    finally:
        _restore(ref, orig)
    n = 0
    for (gres, gthr, gbw, glab), (wres, wthr, wbw, wlab) in zip(got, want):
        assert gthr == wthr and np.array_equal(gbw, wbw) and np.array_equal(glab, wlab)
        for cat in ("OK", "Large", "Small"):
            assert len(gres[cat]) == len(wres[cat])
            for g, w in zip(gres[cat], wres[cat]):
                n += 1
                assert g["label"] == w["label"] and g["area"] == w["area"] and g["centroid"] == w["centroid"]
                assert close(float(g["mean_int_raw"]), float(w["mean_int_raw"]))
                assert g["bg_level"] == w["bg_level"]
    assert n > 3


def test_fret_stage_and_mor_rebound(emu_engine, tmp_path):
    from imageprocess_b200.host import MOR_by_ROI as mor_mirror, common, fret_ratio_builder as mirror
    ref = refimport.load("fret_ratio_builder")
    d, a, polys = small_scene(8, H=96, W=128, n_cells=2)
    img_dir = str(tmp_path)
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    common.write_tiff(os.path.join(img_dir, "S01_1.tif"), d)
    common.write_tiff(os.path.join(img_dir, "S01_2.tif"), a)
    _write_rois(os.path.join(roi_dir, "S01.json"), polys, d.shape)
    # the dict the reference's GUI assembles (fret_ratio_builder.py:567-589), file outputs off
    p = {"img_dir": img_dir, "roi_dir": roi_dir, "out_root": "", "timelapse": False, "ratio_mode": "Donor/FRET",
         "donor_ch": 1, "acceptor_ch": 2, "fret_ch": 2, "bg_scope": "full", "bg_mode": "percentile", "percentile": 1.0,
         "per_channel_p": False, "donor_p": 1.0, "fret_p": 1.0, "clip_neg": True, "eps_percentile": 1.0,
         "px_um": None, "out_xls": True, "out_tif": False, "out_png": False, "save_full": False, "save_crop": True,
         "mask_outside": True, "apply_cmap": True, "cmap_name": "jet", "show_colorbar": False, "png_dpi": 300,
         "add_scalebar": False, "scale_bar_um": 20.0, "cmin_txt": "", "cmax_txt": "", "fixed_crop": True,
         "crop_w": 500, "crop_h": 500, "subset_on": False, "subset_stage": "", "subset_time": "",
         "subset_roi": "", "n_workers": 1, "lang": "en"}
    pairs = [(("S01", None), os.path.join(img_dir, "S01_1.tif"), os.path.join(img_dir, "S01_2.tif"))]
    paths = (img_dir, None, None, None, None, None, None)
    _, want_rows, _ = ref.process_one_stage("S01", pairs, dict(p), paths)
    orig = _rebind(ref, mirror, ["process_one_stage"])
    try:
        key, got_rows, logs = ref.process_one_stage("S01", pairs, dict(p), paths)
    finally:
        _restore(ref, orig)
    assert key == "S01" and want_rows and len(got_rows) == len(want_rows)
    for g, w in zip(got_rows, want_rows):
        for k in ("stage", "time", "roi", "area_px", "ratio_median", "ratio_p5", "ratio_p95", "donor_median",
                  "yfret_median", "p", "ratio_mode", "bg_mode"):
            assert g[k] == w[k], (k, g[k], w[k])
        assert np.float32(g["eps"]) == np.float32(w["eps"])
        for k in ("ratio_mean", "ratio_std", "donor_mean", "yfret_mean"):
            assert close(g[k], w[k]), k
    # MOR_by_ROI.morphology_from_polygon
    refm = refimport.load("MOR_by_ROI")
    want = refm.morphology_from_polygon(polys[0], d.shape, 0.223)
    orig = _rebind(refm, mor_mirror, ["morphology_from_polygon"])
    try:
        got = refm.morphology_from_polygon(polys[0], d.shape, 0.223)
    finally:
        _restore(refm, orig)
    assert set(got) == set(want) and got["area_px"] == want["area_px"]
    for k, wv in want.items():
        assert close(float(got[k]), float(wv), 1e-9), k


@pytest.mark.parametrize("seed", [0, 3, 5])
def test_nesprin2_run_pipeline_against_reference(emu_engine, seed):
    """Nesprin2_FRET_Builder.run_pipeline(p) of the unmodified reference and of the mirror on the same random folder and
    parameter dict (tests/fuzz/fuzz_route_a_nesprin2.py: 600 such folders were run by hand): the two
    nesprin2_fret_perROI.csv agree cell for cell."""
    from tests.fuzz import fuzz_route_a_nesprin2 as fz
    assert fz.run_seed(seed, refimport.load("Nesprin2_FRET_Builder")) > 0
