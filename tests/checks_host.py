"""Check bodies for the host mirrors (imageprocess_b200/host/*): every function takes an Engine
and a tmp directory, builds the reference's folder layout on disk, runs the mirror's entry point
and compares tables / TIFFs with the oracle (and the shipped golden CSV for Fluor_INT)."""
import csv
import json
import math
import os

import numpy as np

from imageprocess_b200.host import (FA_Analyzer, Fluor_INT, MOR_by_ROI, Nesprin2_FRET_Builder, common,
                                    fret_ratio_builder, roi_channel_cropper)
from oracle import port
from oracle.gen_golden import small_scene
from tests import goldenio
from tests.checks import N2_BASE, close


def _write_rois(path, polys, shape):
    with open(path, "w") as f:
        json.dump({"name": os.path.basename(path)[:-5], "image_shape": {"height": shape[0], "width": shape[1]},
                   "rois": [np.asarray(P).tolist() for P in polys]}, f)


def compare_csv_text(got_path, want_path, tolerant=lambda col: False, rel=1e-5):
    """The written table against a reference-written one, TEXT for text: same header line, same
    number of lines, every field the same string -- except the columns `tolerant` names (float64
    sums against numpy's pairwise float32 ones), whose text must parse to a value within `rel` and
    be the shortest round-trip repr of that value, like every float pandas writes."""
    with open(got_path, newline="") as f:
        got = list(csv.reader(f))
    with open(want_path, newline="") as f:
        want = list(csv.reader(f))
    assert got[0] == want[0], (got[0], want[0])
    assert len(got) == len(want), (len(got), len(want))
    n_tol = 0
    for gr, wr in zip(got[1:], want[1:]):
        assert len(gr) == len(wr)
        for col, g, w in zip(got[0], gr, wr):
            if g == w:
                continue
            assert tolerant(col), (col, g, w)
            assert close(float(g), float(w), rel), (col, g, w)
            assert g == repr(float(g)), (col, g)
            n_tol += 1
    return n_tol


def check_tiff_roundtrip(eng, tmp):
    rng = np.random.default_rng(0)
    for arr in (rng.integers(0, 65535, (37, 53)).astype(np.uint16), rng.normal(size=(20, 31)).astype(np.float32),
                rng.integers(0, 255, (9, 7)).astype(np.uint8)):
        p = os.path.join(tmp, f"t_{arr.dtype}.tif")
        common.write_tiff(p, arr)
        back = common.read_image_raw(p)
        assert back.dtype == arr.dtype and np.array_equal(back, arr)
    assert common.read_2d(p).dtype == np.float32


def check_fluor_int_golden(eng, tmp, exp="e1_P0"):
    """1Intensity.bat config (BASELINE C1): the shipped experiment folder rebuilt on disk ->
    _process_key_task / run_headless -> the reference's own fluor_intensity_perROI.csv."""
    imgs, polys, rows, _ = goldenio.load_intensity(exp)
    img_dir = os.path.join(tmp, exp)
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    for ch in (2, 3):
        common.write_tiff(os.path.join(img_dir, f"S01_{ch}.TIF"), imgs[ch])
    common.write_tiff(os.path.join(img_dir, "S01_1.TIF"), imgs[2])           # a channel that is not quantified
    _write_rois(os.path.join(roi_dir, "S01.json"), polys, imgs[2].shape)
    cfg = {"channels_to_quant": [2, 3], "bg_stride": 4, "percentile": 1.0, "ch_color_map": {2: "Green", 3: "Red"}}
    tasks, keymap = Fluor_INT.build_tasks(img_dir, roi_dir, os.path.join(img_dir, "RES"), cfg)
    assert list(keymap) == [("S01", None)] and sorted(keymap[("S01", None)]) == [1, 2, 3]
    res = Fluor_INT._process_key_task(tasks[0], eng=eng)
    assert res["logs"] == [f"[DONE-QUANT] S01 ROI={len(rows)}"] and res["steps"] == len(rows)
    for g, e in zip(res["rows"], rows):
        assert g["roi"] == int(e["roi"]) and g["area_px"] == int(e["area_px"]) and g["stage"] == e["stage"] == "S01"
        assert g["time"] is None and g["bg_scope"] == e["bg_scope"] and g["bg_mode"] == e["bg_mode"]
        assert str(g["clip_neg"]) == e["clip_neg"] and g["bg_stride"] == int(e["bg_stride"])
        for ch in (2, 3):
            assert g[f"ch{ch}_bg"] == float(e[f"ch{ch}_bg"]) and g[f"ch{ch}_p"] == float(e[f"ch{ch}_p"])
            for k in ("median", "p5", "p95", "vmin", "vmax"):
                assert g[f"ch{ch}_{k}"] == float(e[f"ch{ch}_{k}"]), (ch, k)
            assert g[f"ch{ch}_npx"] == int(e[f"ch{ch}_npx"])
            for k in ("mean", "std", "vsum"):
                assert close(g[f"ch{ch}_{k}"], float(e[f"ch{ch}_{k}"])), (ch, k)
    # batched driver writes the CSV the .bat workflow expects
    got = Fluor_INT.run_headless(img_dir, roi_dir, cfg=cfg, eng=eng, log=lambda s: None)
    assert len(got) == len(rows)
    out_csv = os.path.join(img_dir, "RES", "xls", "fluor_intensity_perROI.csv")
    compare_csv_text(out_csv, os.path.join(goldenio.GOLD, "intensity", exp, "expected.csv"),
                     tolerant=lambda c: c.endswith(("_mean", "_std", "_vsum")))
    # worker never raises: a broken task comes back as a log line
    bad = dict(tasks[0])
    bad["chmap"] = {2: os.path.join(img_dir, "missing.tif")}
    r = Fluor_INT._process_key_task(bad, eng=eng)
    assert r["rows"] == [] and r["logs"][0].startswith("[ERROR][WORKER] S01")


def check_fluor_int_tifs_and_masks(eng, tmp):
    """do_tif outputs (bg-corrected float32 + 16-bit preview), timelapse names, a PNG union mask
    and a stage without any ROI."""
    from PIL import Image
    d, a, polys = small_scene(41, H=96, W=128, n_cells=2)
    img_dir = os.path.join(tmp, "int2")
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    for s, tt in (("S1", "t0"), ("S2", "t3"), ("S3", "t0")):
        common.write_tiff(os.path.join(img_dir, f"{s}_{tt}_c1.tif"), d)
        common.write_tiff(os.path.join(img_dir, f"{s}_{tt}_c2.tif"), a)
    _write_rois(os.path.join(roi_dir, "S1_t0.json"), polys, d.shape)          # legacy ROI name
    union = np.zeros(d.shape, bool)
    for P in polys:
        union |= port.rasterize_polygon(P, d.shape)
    Image.fromarray((union * 255).astype(np.uint8)).save(os.path.join(roi_dir, "S02_t03.png"))
    cfg = {"timelapse": True, "channels_to_quant": [1, 2], "out_tif": True, "tif_mask_outside": True,
           "bg_stride": 1, "bg_scope": "roi_union", "percentile": 3.0}
    tasks, keymap = Fluor_INT.build_tasks(img_dir, roi_dir, os.path.join(img_dir, "RES"), cfg)
    assert list(keymap) == [("S01", "t00"), ("S02", "t03"), ("S03", "t00")]
    task = {"bg_scope": "roi_union", "bg_mode": "percentile", "clip_neg": True, "bg_stride": 1,
            "percentile": 3.0, "per_channel_p": False, "ch_p_map": {}}
    raw = {1: d.astype(np.float32), 2: a.astype(np.float32)}
    # S01_t00: polygons
    r = Fluor_INT._process_key_task(tasks[0], eng=eng)
    want, wbg, bc = port.int_process_key({k: v.copy() for k, v in raw.items()}, polys, None, task)
    assert len(r["rows"]) == len(want) and all(x["time"] == "t00" for x in r["rows"])
    for g, w in zip(r["rows"], want):
        assert g["area_px"] == w["area_px"] and g["ch1_median"] == w["ch1_median"] and g["ch2_p95"] == w["ch2_p95"]
        assert g["ch1_bg"] == wbg[1]["bg"]
    for ch in (1, 2):
        got32 = common.read_image_raw(os.path.join(img_dir, "RES", "TIF", "bgcorr32", f"S01_t00_ch{ch}_bgcorr.tif"))
        masked = np.zeros_like(bc[ch])
        masked[union] = bc[ch][union]
        assert got32.dtype == np.float32 and np.array_equal(got32, masked)
        got16 = common.read_image_raw(os.path.join(img_dir, "RES", "TIF", "bgcorr16_preview", f"S01_t00_ch{ch}_bgcorr_preview.tif"))
        assert np.array_equal(got16, port.preview_u16(masked, 1.0, 99.0))
    # S02_t03: PNG union mask -> one row, roi 1
    r = Fluor_INT._process_key_task(tasks[1], eng=eng)
    want, wbg, _ = port.int_process_key({k: v.copy() for k, v in raw.items()}, None, union, task)
    assert len(r["rows"]) == 1 and r["rows"][0]["roi"] == 1 and r["rows"][0]["area_px"] == want[0]["area_px"]
    for k in ("ch1_median", "ch1_p5", "ch2_p95", "ch2_vmax", "ch1_npx"):
        assert r["rows"][0][k] == want[0][k], k
    assert close(r["rows"][0]["ch2_mean"], want[0]["ch2_mean"]) and r["rows"][0]["ch2_bg"] == wbg[2]["bg"]
    # S03_t00: nothing to measure
    r = Fluor_INT._process_key_task(tasks[2], eng=eng)
    assert r["rows"] == [] and "S03_t00" in r["logs"][0]


def check_fa_mirror(eng, tmp):
    px = 0.112
    params = {"alpha": 2.0, "min_area_um": 12.5 * px ** 2, "max_area_um": 300.0 * px ** 2, "close_radius": 1,
              "subtract_bg": True}
    cfg = FA_Analyzer.convert_um_to_px_config(params, px)
    d, a, polys = small_scene(21, H=120, W=168, n_cells=2, blobs=10)
    img = d.astype(np.float32)
    stats = port.fa_global_stats(img)
    # single-crop entry point, the GUI's call shape
    for P in polys:
        crop, mask, _ = port.fa_crop_and_mask(img, P.copy())
        wres, wthr, wbw, wlab = port.analyze_fa_crop(crop, mask, cfg, stats, with_contours=True)
        gres, gthr, gbw, glab = FA_Analyzer.analyze_fa_crop(crop, mask, cfg, stats, eng=eng)
        assert gthr == wthr and type(gthr) is type(wthr)
        assert np.array_equal(gbw, wbw) and np.array_equal(glab, wlab)
        for cat in ("OK", "Large", "Small"):
            assert len(gres[cat]) == len(wres[cat])
            for g, w in zip(gres[cat], wres[cat]):
                assert g["label"] == w["label"] and g["area"] == w["area"] and g["centroid"] == w["centroid"]
                assert close(float(g["mean_int_raw"]), float(w["mean_int_raw"]))
                assert g["contour"].dtype == np.float64 and np.array_equal(g["contour"], w["contour"])     # FA_Analyzer.py:168-170
    e = FA_Analyzer.analyze_fa_crop(np.array([]), np.zeros((0,), bool), cfg, stats, eng=eng)
    assert e[0] == {"OK": [], "Large": [], "Small": []} and e[1] == 0
    # batch body on a folder
    img_dir = os.path.join(tmp, "fa")
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    scenes = {"S01": (d, polys)}
    d2, _, polys2 = small_scene(22, H=120, W=168, n_cells=2, blobs=10)
    scenes["S02"] = (d2, polys2)
    for s_tag, (im, pl) in scenes.items():
        common.write_tiff(os.path.join(img_dir, f"{s_tag}_1.tif"), im)
        common.write_tiff(os.path.join(img_dir, f"{s_tag}_2.tif"), im)
        _write_rois(os.path.join(roi_dir, f"{s_tag}.json"), pl, im.shape)
    fl = FA_Analyzer.load_file_list(img_dir, roi_dir, "1")
    assert [x[2] for x in fl] == ["S01", "S02"]
    n = FA_Analyzer.run_batch(fl, params, px, os.path.join(img_dir, "out"), save_ok_only=False, eng=eng, log=lambda s: None)
    assert n == 2
    schema = json.load(open(os.path.join(goldenio.GOLD, "fa_csv_schema.json")))
    for s_tag, (im, pl) in scenes.items():
        with open(os.path.join(img_dir, "out", "individual_results", f"{s_tag}_results.csv"), newline="") as f:
            rd = csv.DictReader(f)
            back = list(rd)
            cols = rd.fieldnames
        assert cols == next(iter(schema.values()))["header"].split(","), cols
        fimg = im.astype(np.float32)
        st = port.fa_global_stats(fimg)
        want = port.fa_batch_rows(fimg, pl, params, px, s_tag=s_tag, save_ok_only=False, with_contours=False, stats=st)
        if len(back) != len(want) or np.float32(float(back[0]["Global_Threshold"])) != np.float32(want[0]["Global_Threshold"]):
            # float32 threshold one ulp from numpy's pairwise value: the oracle fed with the device's statistics
            from imageprocess_b200 import pipeline
            dev_st = pipeline.fa_batch(eng, eng.mem.from_host(im[None, None]), (1, 1) + im.shape, [pl], params, px)["stats"][0]
            assert close(float(dev_st[0]), float(st[0]), 1e-6) and close(float(dev_st[1]), float(st[1]), 1e-6) and dev_st[2] == st[2]
            st = (np.float32(dev_st[0]), np.float32(dev_st[1]), st[2])
            want = port.fa_batch_rows(fimg, pl, params, px, s_tag=s_tag, save_ok_only=False, with_contours=False, stats=st)
        assert len(back) == len(want), (s_tag, len(back), len(want))
        for g, w in zip(back, want):
            assert g["File"] == w["File"] and int(g["Cell_ID"]) == w["Cell_ID"] and g["Category"] == w["Category"]
            assert float(g["Area_px"]) == float(w["Area_px"])
            assert g["Area_px"] == str(w["Area_px"])                       # '300.0': float64 area, as in the golden CSV
            assert close(float(g["Mean_Intensity_Raw"]), float(w["Mean_Intensity_Raw"]))


def check_fret_mirror(eng, tmp):
    frames = [small_scene(s, H=96, W=128, n_cells=2) for s in (8, 9)]
    img_dir = os.path.join(tmp, "fret")
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    for k, (d, a, polys) in enumerate(frames):
        common.write_tiff(os.path.join(img_dir, f"S01_t{k:02d}_1.tif"), d)
        common.write_tiff(os.path.join(img_dir, f"S01_t{k:02d}_2.tif"), a)
        _write_rois(os.path.join(roi_dir, f"S01_t{k:02d}.json"), polys, d.shape)
    common.write_tiff(os.path.join(img_dir, "S02_t00_1.tif"), frames[0][0])       # stage without ROI file
    common.write_tiff(os.path.join(img_dir, "S02_t00_2.tif"), frames[0][1])
    p = {"timelapse": True, "donor_ch": 1, "fret_ch": 2, "ratio_mode": "FRET/Donor", "bg_scope": "roi_union",
         "percentile": 2.0, "eps_percentile": 1.5, "out_tif": True}
    rows = fret_ratio_builder.run_headless(img_dir, roi_dir, p=p, eng=eng, log=lambda s: None)
    pp = {**fret_ratio_builder.DEFAULT_P, **p}
    k0 = 0
    for k, (d, a, polys) in enumerate(frames):
        want = port.fret_process_pair(d.astype(np.float32), a.astype(np.float32), polys, pp)
        got = rows[k0: k0 + len(want["rows"])]
        k0 += len(want["rows"])
        for g, w in zip(got, want["rows"]):
            assert g["stage"] == "S01" and g["time"] == f"t{k:02d}" and g["roi"] == w["roi"] and g["area_px"] == w["area_px"]
            assert np.float32(g["eps"]) == np.float32(want["eps"])
            for key in ("ratio_median", "ratio_p5", "ratio_p95", "donor_median", "yfret_median"):
                assert g[key] == w[key], key
            assert close(g["ratio_mean"], w["ratio_mean"]) and close(g["ratio_std"], w["ratio_std"])
        base = os.path.join(img_dir, "RES", "TIF")
        R = common.read_image_raw(os.path.join(base, "ratio32", f"S01_t{k:02d}_ratio_FoverD.tif"))
        assert np.array_equal(R, want["R_full"], equal_nan=True)
        Rroi = common.read_image_raw(os.path.join(base, "ratio32_roi", f"S01_t{k:02d}_ratio_FoverD.tif"))
        assert np.array_equal(Rroi, want["R_roi"], equal_nan=True)
        pv = common.read_image_raw(os.path.join(base, "ratio16_preview", f"S01_t{k:02d}_ratio_FoverD_preview.tif"))
        assert np.array_equal(pv, port.preview_u16(want["R_full"], 1.0, 99.0))
        pv = common.read_image_raw(os.path.join(base, "ratio16_roi_preview", f"S01_t{k:02d}_ratio_FoverD_preview.tif"))
        assert np.array_equal(pv, port.preview_u16(want["R_roi"], 1.0, 99.0))
    assert k0 == len(rows)
    assert os.path.exists(os.path.join(img_dir, "RES", "TIF", "ratio32", "S02_t00_ratio_FoverD.tif"))
    with open(os.path.join(img_dir, "RES", "xls", "fret_ratio_perROI.csv"), newline="") as f:
        back = list(csv.reader(f))
    # the reference's 17 columns in its order + its three derived ones (fret_ratio_builder.py:981-990)
    assert back[0] == ["stage", "time", "roi", "area_px", "ratio_mean", "ratio_median", "ratio_std", "ratio_p5",
                       "ratio_p95", "donor_mean", "donor_median", "yfret_mean", "yfret_median", "eps", "p",
                       "ratio_mode", "bg_mode", "time_idx", "stage_idx", "roi_lab"]
    assert len(back) == 1 + len(rows)
    for line, r in zip(back[1:], rows):
        rec = dict(zip(back[0], line))
        assert rec["stage"] == r["stage"] and rec["time"] == r["time"] and int(rec["roi"]) == r["roi"]
        assert rec["ratio_median"] == repr(float(r["ratio_median"])) and rec["eps"] == repr(float(r["eps"]))
        assert rec["time_idx"] == str(int(r["time"][1:])) and rec["stage_idx"] == "1" and rec["roi_lab"] == f"s1c{r['roi']}"


def check_nesprin2_mirror(eng, tmp):
    frames = [small_scene(s, H=96, W=128, n_cells=2) for s in (11, 12)]
    img_dir = os.path.join(tmp, "n2")
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    rng = np.random.default_rng(3)
    planes = []
    for k, (d, a, polys) in enumerate(frames, 1):
        ao = (0.3 * a + rng.poisson(50, d.shape)).astype(np.uint16)
        d = d.copy()
        d[rng.random(d.shape) < 0.002] = 65535
        planes.append((d, a, ao))
        common.write_tiff(os.path.join(img_dir, f"S{k:02d}_2.tif"), d)
        common.write_tiff(os.path.join(img_dir, f"S{k:02d}_3.tif"), a)
        common.write_tiff(os.path.join(img_dir, f"S{k:02d}_4.tif"), ao)
        _write_rois(os.path.join(roi_dir, f"S{k:02d}.json"), polys, d.shape)
    p = {**N2_BASE, "img_dir": img_dir, "roi_dir": roi_dir, "use_spectral": True, "alpha": 0.1, "beta": 0.04,
         "g_factor": 1.05, "aonly_ch": 4, "donor_ch": 2, "fret_ch": 3, "out_tif": True}
    rows = Nesprin2_FRET_Builder.run_pipeline(p, eng=eng, log=lambda s: None)
    k0 = 0
    for k, ((d, a, ao), (_, _, polys)) in enumerate(zip(planes, frames), 1):
        want = port.n2_process_pair(d.astype(np.float32), a.astype(np.float32), polys, p, Aonly=ao.astype(np.float32))
        for w in want["rows"]:
            g = rows[k0]
            k0 += 1
            assert g["stage"] == f"S{k:02d}" and g["time"] is None and g["roi"] == w["roi"] and g["area_px"] == w["area_px"]
            assert np.float32(g["eps"]) == np.float32(w["eps"])
            for key in ("ratio_median", "ratio_p5", "ratio_p95"):
                assert g[key] == w[key] or (math.isnan(g[key]) and math.isnan(w[key]))
            for key in ("ratio_mean", "ratio_FoverD_mean", "ratio_DoverF_mean", "donor_mean", "fret_mean"):
                assert close(g[key], w[key]), key
        R = common.read_image_raw(os.path.join(img_dir, "RES", "TIF", "ratio32_full", f"S{k:02d}_ratio_FoverD.tif"))
        assert np.array_equal(R, want["R_full"], equal_nan=True)
    assert k0 == len(rows)
    with open(os.path.join(img_dir, "RES", "xls", "nesprin2_fret_perROI.csv"), newline="") as f:
        back = list(csv.reader(f))
    # save_xls: kept columns in the reference's order, then stage_idx, time_idx, roi_lab (Nesprin2_FRET_Builder.py:1292-1306)
    assert back[0] == Nesprin2_FRET_Builder.KEEP_COLS + ["stage_idx", "time_idx", "roi_lab"]
    assert len(back) == 1 + len(rows)
    rec = dict(zip(back[0], back[1]))
    assert rec["time"] == "" and rec["time_idx"] == "0" and rec["roi_lab"] == "s1c1" and rec["clip_neg"] in ("True", "False")


def check_mor_and_cropper_mirrors(eng, tmp):
    d, a, polys = small_scene(15, H=120, W=160, n_cells=2)
    img_dir = os.path.join(tmp, "mor")
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    common.write_tiff(os.path.join(img_dir, "S01_1.tif"), d)
    common.write_tiff(os.path.join(img_dir, "S01_2.tif"), a)
    _write_rois(os.path.join(roi_dir, "S01.json"), polys, d.shape)
    g = MOR_by_ROI.morphology_from_polygon(polys[0], d.shape, 0.223, eng=eng)
    w = port.morphology_from_polygon(polys[0], d.shape, 0.223)
    assert g["area_px"] == w["area_px"] and close(g["major_um"], w["major_um"], 1e-9)
    rows = MOR_by_ROI.run_headless(img_dir, roi_dir, sel_ch=2, px_um=0.223, eng=eng, log=lambda s: None)
    assert [r["roi"] for r in rows] == list(range(1, len(polys) + 1)) and rows[0]["img"] == "S01_2.tif"
    with open(os.path.join(img_dir, "RES_MOR", "xls", "morphology_perROI.csv"), newline="") as f:
        assert csv.DictReader(f).fieldnames == MOR_by_ROI.COLUMNS
    done = roi_channel_cropper.run_headless(img_dir, roi_dir, ch_select=1, low_cut=1.0, high_cut=1.0, gamma=1.0, eng=eng,
                                            log=lambda s: None)
    assert len(done) == len(polys)
    for (keytag, i, rect), P in zip(done, polys):
        wv = port.cropper_normalize(d.astype(np.float32), d, P, 1.0, 1.0, 1.0)
        assert rect == wv["rect"]
        t16 = common.read_image_raw(os.path.join(img_dir, "RES_CROP", "TIF16", f"{keytag}_roi{i}_ch1.tif"))
        assert np.array_equal(t16, wv["out16"])
        traw = common.read_image_raw(os.path.join(img_dir, "RES_CROP", "TIF", f"{keytag}_roi{i}_ch1.tif"))
        assert np.array_equal(traw, wv["raw_out"])


def check_stream_batches(eng, tmp):
    """Folder streaming (host/stream.py) over several batches: 5 (stage, time) keys with 2 frames per
    batch (three batches, the last one padded), one unreadable TIFF in the middle: every readable key
    gets exactly the oracle's rows, in task order; the broken key becomes a log line only."""
    frames = [small_scene(s, H=96, W=128, n_cells=2) for s in (61, 62, 63, 64, 65)]
    img_dir = os.path.join(tmp, "stream")
    roi_dir = os.path.join(img_dir, "roi")
    os.makedirs(roi_dir)
    for k, (d, a, polys) in enumerate(frames):
        common.write_tiff(os.path.join(img_dir, f"S01_t{k:02d}_1.tif"), d)
        common.write_tiff(os.path.join(img_dir, f"S01_t{k:02d}_2.tif"), a)
        _write_rois(os.path.join(roi_dir, f"S01_t{k:02d}.json"), polys, d.shape)
    with open(os.path.join(img_dir, "S01_t02_2.tif"), "r+b") as f:          # same header (shape), pixel data cut off
        f.truncate(4000)
    cfg = {"timelapse": True, "channels_to_quant": [1, 2], "bg_stride": 4, "percentile": 1.0}
    tasks, _ = Fluor_INT.build_tasks(img_dir, roi_dir, os.path.join(img_dir, "RES"), cfg)
    timing = {}
    res = Fluor_INT.process_key_tasks(tasks, eng=eng, frames_per_batch=2, decode_threads=3, timing=timing)
    assert len(res) == 5 and timing["batches"] == 3 and timing["frames"] == 5
    task = {"bg_scope": "full", "bg_mode": "percentile", "clip_neg": True, "bg_stride": 4, "percentile": 1.0,
            "per_channel_p": False, "ch_p_map": {}}
    for k, (d, a, polys) in enumerate(frames):
        if k == 2:
            assert res[k]["rows"] == [] and res[k]["logs"][0].startswith("[ERROR][WORKER] S01_t02")
            continue
        want, wbg, _ = port.int_process_key({1: d.astype(np.float32), 2: a.astype(np.float32)}, polys, None, task)
        got = res[k]["rows"]
        assert len(got) == len(want) and all(r["time"] == f"t{k:02d}" for r in got)
        for g, w in zip(got, want):
            assert g["roi"] == w["roi"] and g["area_px"] == w["area_px"] and g["ch1_bg"] == wbg[1]["bg"]
            for key in ("ch1_median", "ch1_p5", "ch2_p95", "ch2_vmax", "ch2_npx"):
                assert g[key] == w[key], (k, key)
            assert close(g["ch2_mean"], w["ch2_mean"]) and close(g["ch1_std"], w["ch1_std"])
    # the FRET entry point over the same folder (out_tif off: tickets lag two batches)
    rows = fret_ratio_builder.run_headless(img_dir, roi_dir, p={"timelapse": True, "out_tif": False}, eng=eng,
                                           log=lambda s: None, frames_per_batch=2)
    pp = {**fret_ratio_builder.DEFAULT_P, "timelapse": True}
    k0 = 0
    for k, (d, a, polys) in enumerate(frames):
        if k == 2:
            continue
        want = port.fret_process_pair(d.astype(np.float32), a.astype(np.float32), polys, pp)
        for g, w in zip(rows[k0: k0 + len(want["rows"])], want["rows"]):
            assert g["time"] == f"t{k:02d}" and g["roi"] == w["roi"] and g["ratio_median"] == w["ratio_median"]
            assert close(g["ratio_mean"], w["ratio_mean"])
        k0 += len(want["rows"])
    assert k0 == len(rows)


HOST_CHECKS = [check_tiff_roundtrip, check_fluor_int_golden, check_fluor_int_tifs_and_masks, check_fa_mirror,
               check_fret_mirror, check_nesprin2_mirror, check_mor_and_cropper_mirrors, check_stream_batches]


def check_fa_crop_arbitrary_masks(eng, tmp):
    """FA_Analyzer.analyze_fa_crop with masks that are not rasterised polygons (random pixels, an L-shaped block, an
    empty mask), crops one pixel high or wide (the reference's find_contours refuses those with a ValueError as soon
    as a region exists), repeated calls giving identical results (the image plane used to be uploaded through a
    temporary that was freed before the call ran)."""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(17)
    stats = port.fa_global_stats(rng.poisson(500, (64, 64)).astype(np.float32))
    for h, w, kind in ((40, 31, 0), (29, 33, 1), (61, 35, 2), (17, 64, 3), (2, 2, 0), (3, 70, 1)):
        img = rng.poisson(500, (h, w)).astype(np.int64)
        img[ndi.binary_dilation(rng.random((h, w)) < 0.05)] += 2500
        crop = img.astype(np.float32)
        mask = [np.ones((h, w), bool), rng.random((h, w)) < 0.7, np.zeros((h, w), bool), np.zeros((h, w), bool)][kind]
        if kind == 2:
            mask[h // 4:, : 3 * w // 4] = True
        cfg = {"alpha": 2.0, "min_px": float(rng.choice([0.0, 2.5, 6.0])), "max_px": 400.0, "close_radius": int(rng.integers(0, 4)),
               "subtract_bg": True}
        want = port.analyze_fa_crop(crop, mask, cfg, stats, with_contours=True)
        for rep in range(3):
            got = FA_Analyzer.analyze_fa_crop(crop, mask, cfg, stats, eng=eng)
            assert got[1] == want[1] and np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3]), (h, w, kind, rep)
            for cat in ("OK", "Large", "Small"):
                assert [(g["label"], g["area"], g["centroid"]) for g in got[0][cat]] == \
                       [(x["label"], x["area"], x["centroid"]) for x in want[0][cat]]
                assert all(np.array_equal(g["contour"], x["contour"]) for g, x in zip(got[0][cat], want[0][cat]))
    row = np.full((1, 40), 400.0, np.float32)
    row[0, 10:20] = 5000.0
    cfg = {"alpha": 2.0, "min_px": 0.0, "max_px": 400.0, "close_radius": 0, "subtract_bg": True}
    for crop in (row, row.T.copy()):
        m = np.ones(crop.shape, bool)
        for fn in (lambda: port.analyze_fa_crop(crop, m, cfg, stats, with_contours=True),
                   lambda: FA_Analyzer.analyze_fa_crop(crop, m, cfg, stats, eng=eng)):
            try:
                fn()
                raise AssertionError("a 1-pixel-thin crop with a region must raise like skimage's find_contours")
            except ValueError as e:
                assert str(e) == "Input array must be at least 2x2."
        # without a region above the threshold nothing is traced and nothing is raised
        dark = np.full(crop.shape, 400.0, np.float32)
        assert FA_Analyzer.analyze_fa_crop(dark, m, cfg, stats, eng=eng)[0] == port.analyze_fa_crop(dark, m, cfg, stats)[0]


# added after the round's last GPU session: the GPU tier runs it from tests/test_gpu_zz_late_checks.py
LATE_HOST_CHECKS = [check_fa_crop_arbitrary_masks]
