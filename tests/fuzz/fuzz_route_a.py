"""tests/fuzz/fuzz_route_a.py -- TEST INFRASTRUCTURE, run by hand, BUILD CONTAINER ONLY (needs /root/reference), emulated build.

The drop-in boundary against the UNMODIFIED reference workers (loaded through oracle/refimport.py on the restated
third-party shims): random small folders (TIFF pairs + ROI JSON, static and time-lapse names, polygons with fewer than
three points, ROIs off the frame, a stage without ROI file) and random worker settings go through
  * Fluor_INT._process_key_task(task)              (reference Fluor_INT.py:795) and
  * fret_ratio_builder.process_one_stage(...)      (reference fret_ratio_builder.py:429)
of the reference and of the mirrors; rows and log-relevant counts must agree (order statistics, areas, backgrounds,
epsilons exactly; means / standard deviations to 1e-5 of the values' scale).

    python tests/fuzz/fuzz_route_a.py <first seed> <number of seeds>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import json, math, shutil, tempfile, time, traceback
import numpy as np
from oracle import refimport
import imageprocess_b200 as ipb
from imageprocess_b200.ops import Engine
from imageprocess_b200.host import Fluor_INT as mF, fret_ratio_builder as mR, common
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close
ipb._engine = Engine(emu_lib(), NumpyMem())
refimport.install_stubs()
rF, rR = refimport.load("Fluor_INT"), refimport.load("fret_ratio_builder")
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time(); stats = {"int_rows": 0, "fret_rows": 0}


def same(g, w, scale_keys=()):
    assert set(g) == set(w), set(g) ^ set(w)
    for k, wv in w.items():
        gv = g[k]
        if isinstance(wv, float) and math.isnan(wv):
            assert isinstance(gv, float) and math.isnan(gv), (k, gv)
        elif isinstance(wv, float) and k.endswith(("_mean", "_std", "_vsum")):
            pre = k.rsplit("_", 1)[0]
            scale = max(abs(w.get(pre + "_vmin", 0.0) or 0.0), abs(w.get(pre + "_vmax", 0.0) or 0.0), abs(w.get(pre + "_median", 0.0) or 0.0),
                        abs(w.get(pre + "_p95", 0.0) or 0.0), abs(w.get(pre + "_p5", 0.0) or 0.0)) * (max(w.get(pre + "_npx", 1), 1) if k.endswith("_vsum") else 1)
            assert close(gv, wv) or abs(gv - wv) <= 1e-5 * scale, (k, gv, wv)
        elif k == "eps":
            assert np.float32(gv) == np.float32(wv), (k, gv, wv)
        else:
            assert gv == wv, (k, gv, wv)


for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    root = tempfile.mkdtemp(prefix="ipb_fuzz_a_")
    try:
        H, W = int(rng.integers(24, 100)), int(rng.choice([8 * int(rng.integers(3, 16)), int(rng.integers(25, 130))]))
        timelapse = bool(rng.integers(0, 2))
        roi_dir = os.path.join(root, "roi")
        os.makedirs(roi_dir)
        keys = [(s, t) for s in (1, 2) for t in ((0, 3) if timelapse else (None,))]
        for s, t in keys:
            stem = f"S{s:02d}" + (f"_t{t:02d}" if t is not None else "")
            for ch in (1, 2):
                img = rng.poisson(float(rng.choice([40, 800, 9000])), (H, W)).clip(0, 65535).astype(np.uint16)
                if rng.random() < 0.3: img[rng.random((H, W)) < 0.01] = 65535
                common.write_tiff(os.path.join(root, f"{stem}_{ch}.tif"), img)
            if s == 2 and t in (None, 3) and rng.random() < 0.5:
                continue                                                     # a key without ROI file
            polys = []
            for k in range(int(rng.integers(1, 4))):
                nv = int(rng.choice([2, 3, 4, 5, 7]))
                P = np.stack([rng.uniform(-8, W + 8, nv), rng.uniform(-8, H + 8, nv)], axis=1)
                if rng.random() < 0.4: P = np.round(P * 2) / 2
                polys.append(P.tolist())
            legacy = rng.random() < 0.3
            name = (f"S{s}" + (f"_t{t}" if t is not None else "")) if legacy else stem
            with open(os.path.join(roi_dir, name + ".json"), "w") as fh:
                json.dump({"name": stem, "image_shape": {"height": H, "width": W}, "rois": polys}, fh)
        cfg = {"channels_to_quant": [1, 2], "timelapse": timelapse, "bg_scope": str(rng.choice(["full", "roi_union"])),
               "bg_mode": str(rng.choice(["percentile", "percentile", "hist-mode"])), "percentile": float(rng.choice([0.0, 1.0, 10.0, 50.0])),
               "per_channel_p": bool(rng.integers(0, 2)), "ch_p_map": {2: 7.5}, "clip_neg": bool(rng.integers(0, 2)),
               "bg_stride": int(rng.choice([1, 3, 4])), "out_tif": False, "out_png": False}
        tasks, _ = mF.build_tasks(root, roi_dir, os.path.join(root, "RES"), cfg)
        for task in tasks:
            task = dict(task); task.update({"px_um": None, "lang": "en"})
            want = rF._process_key_task(dict(task))
            got = mF._process_key_task(dict(task))
            assert len(got["rows"]) == len(want["rows"]) and got["steps"] == want["steps"], ("INT", task["stid"], len(got["rows"]), len(want["rows"]), got["logs"], want["logs"])
            assert [l.split(":")[0] for l in got["logs"]] == [l.split(":")[0] for l in want["logs"]] or len(got["logs"]) == len(want["logs"]), (got["logs"], want["logs"])
            for g, w in zip(got["rows"], want["rows"]):
                same(g, w)
            stats["int_rows"] += len(want["rows"])
        p = {"img_dir": root, "roi_dir": roi_dir, "out_root": "", "timelapse": timelapse, "ratio_mode": str(rng.choice(["Donor/FRET", "FRET/Donor"])),
             "donor_ch": 1, "acceptor_ch": 2, "fret_ch": 2, "bg_scope": str(rng.choice(["full", "roi_union"])), "bg_mode": str(rng.choice(["percentile", "hist-mode"])),
             "percentile": float(rng.choice([1.0, 5.0, 50.0])), "per_channel_p": bool(rng.integers(0, 2)), "donor_p": 1.0, "fret_p": 3.0,
             "clip_neg": bool(rng.integers(0, 2)), "eps_percentile": float(rng.choice([0.0, 1.0, 5.0])), "px_um": None, "out_xls": True, "out_tif": False,
             "out_png": False, "save_full": False, "save_crop": True, "mask_outside": True, "apply_cmap": True, "cmap_name": "jet", "show_colorbar": False,
             "png_dpi": 300, "add_scalebar": False, "scale_bar_um": 20.0, "cmin_txt": "", "cmax_txt": "", "fixed_crop": True, "crop_w": 500, "crop_h": 500,
             "subset_on": False, "subset_stage": "", "subset_time": "", "subset_roi": "", "n_workers": 1, "lang": "en"}
        files = common.list_tifs(root) if hasattr(common, "list_tifs") else sorted(os.path.join(root, f) for f in os.listdir(root) if f.endswith(".tif"))
        pairs_all, _ = mR.build_pairs_by_channel(files, timelapse, 1, 2)
        for stage in ("S01", "S02"):
            pairs = [pr for pr in pairs_all if pr[0][0] == stage]
            paths = (root, None, None, None, None, None, None)
            with np.errstate(all="ignore"):
                _, want_rows, wlogs = rR.process_one_stage(stage, pairs, dict(p), paths)
            _, got_rows, glogs = mR.process_one_stage(stage, pairs, dict(p), paths)
            assert len(got_rows) == len(want_rows), ("FRET", stage, len(got_rows), len(want_rows), glogs, wlogs)
            for g, w in zip(got_rows, want_rows):
                same(g, w)
            stats["fret_rows"] += len(want_rows)
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, type(e).__name__, str(e)[:400], [(t.filename.split("/")[-1], t.lineno) for t in tb][-3:], flush=True)
    finally:
        shutil.rmtree(root, ignore_errors=True)
print("done", seed0, n, "bad", bad, stats, round(time.time() - t0, 1), flush=True)
