"""tests/fuzz/fuzz_folder.py -- TEST INFRASTRUCTURE, run by hand, BUILD CONTAINER ONLY (needs /root/reference), emulated build.

The folder-level entry point Fluor_INT.run_headless (threaded decode -> pinned ring -> ONE persistent job per shape group,
tickets collected two batches behind) against the UNMODIFIED reference worker called key by key in task order: folders with
MIXED image shapes (several groups), small frames_per_batch (many batches), a truncated TIFF, keys without ROI file.  The rows
must come back in the reference's order with the reference's values; an unreadable image costs its own key only.

    python tests/fuzz/fuzz_folder.py <first seed> <number of seeds>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import json, math, shutil, tempfile, time, traceback, io, contextlib
import numpy as np
from oracle import refimport
import imageprocess_b200 as ipb
from imageprocess_b200.ops import Engine
from imageprocess_b200.host import Fluor_INT as mF, common, fret_ratio_builder as mR
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close
ipb._engine = Engine(emu_lib(), NumpyMem())
refimport.install_stubs()
rF, rR = refimport.load("Fluor_INT"), refimport.load("fret_ratio_builder")
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0; t0 = time.time(); stats = {"rows": 0, "broken_keys": 0}
for seed in range(seed0, seed0 + n):
    rng = np.random.default_rng(seed)
    root = tempfile.mkdtemp(prefix="ipb_fuzz_f_")
    try:
        shapes = [(int(rng.integers(24, 80)), 8 * int(rng.integers(4, 12))) for _ in range(int(rng.integers(1, 4)))]
        roi_dir = os.path.join(root, "roi")
        os.makedirs(roi_dir)
        n_t = int(rng.integers(2, 7))
        broken = None
        for s in (1, 2, 3):
            H, W = shapes[int(rng.integers(0, len(shapes)))]
            for t in range(n_t):
                stem = f"S{s:02d}_t{t:02d}"
                for ch in (1, 2):
                    img = rng.poisson(float(rng.choice([40, 800, 9000])), (H, W)).clip(0, 65535).astype(np.uint16)
                    common.write_tiff(os.path.join(root, f"{stem}_{ch}.tif"), img)
                if rng.random() < 0.15:
                    continue
                polys = []
                for k in range(int(rng.integers(1, 4))):
                    nv = int(rng.choice([3, 4, 5, 7]))
                    polys.append(np.stack([rng.uniform(-4, W + 4, nv), rng.uniform(-4, H + 4, nv)], axis=1).tolist())
                with open(os.path.join(roi_dir, stem + ".json"), "w") as fh:
                    json.dump({"name": stem, "image_shape": {"height": H, "width": W}, "rois": polys}, fh)
        if rng.random() < 0.5:                                  # one truncated file
            broken = os.path.join(root, f"S{int(rng.integers(1, 4)):02d}_t{int(rng.integers(0, n_t)):02d}_{int(rng.integers(1, 3))}.tif")
            with open(broken, "r+b") as fh:
                fh.truncate(60)
        cfg = {"channels_to_quant": [1, 2], "timelapse": True, "bg_scope": str(rng.choice(["full", "roi_union"])),
               "bg_mode": str(rng.choice(["percentile", "hist-mode"])), "percentile": float(rng.choice([1.0, 10.0])), "per_channel_p": False, "ch_p_map": {},
               "clip_neg": bool(rng.integers(0, 2)), "bg_stride": int(rng.choice([1, 4])), "out_tif": False, "out_png": False, "out_xls": False}
        tasks, _ = mF.build_tasks(root, roi_dir, os.path.join(root, "RES"), cfg)
        want = []
        for task in tasks:
            task = dict(task); task.update({"px_um": None, "lang": "en"})
            with contextlib.redirect_stdout(io.StringIO()):
                r = rF._process_key_task(dict(task))
            want.extend(r["rows"])
            if not r["rows"] and any(l.startswith("[ERROR]") for l in r["logs"]):
                stats["broken_keys"] += 1
        logs = []
        got = mF.run_headless(root, roi_dir, out_root=os.path.join(root, "RES2"), cfg=cfg, log=logs.append,
                              frames_per_batch=int(rng.choice([1, 2, 3, 5])))
        assert [(g["stage"], g["time"], g["roi"]) for g in got] == [(w["stage"], w["time"], w["roi"]) for w in want], \
            ("row order / set", len(got), len(want), [l for l in logs if "ERROR" in l][:2])
        for g, w in zip(got, want):
            assert set(g) == set(w), set(g) ^ set(w)
            for k, wv in w.items():
                gv = g[k]
                if isinstance(wv, float) and math.isnan(wv):
                    assert math.isnan(gv), k
                elif isinstance(wv, float) and k.endswith(("_mean", "_std", "_vsum")):
                    pre = k.rsplit("_", 1)[0]
                    scale = max(abs(w[pre + "_vmin"]), abs(w[pre + "_vmax"])) * (max(w[pre + "_npx"], 1) if k.endswith("_vsum") else 1)
                    assert close(gv, wv) or abs(gv - wv) <= 1e-5 * scale, (k, gv, wv)
                else:
                    assert gv == wv, (k, gv, wv)
        stats["rows"] += len(want)
        # the same folder through fret_ratio_builder.process_one_stage, stage by stage, in small batches
        pf = {"img_dir": root, "roi_dir": roi_dir, "out_root": "", "timelapse": True, "ratio_mode": str(rng.choice(["Donor/FRET", "FRET/Donor"])),
              "donor_ch": 1, "acceptor_ch": 2, "fret_ch": 2, "bg_scope": str(rng.choice(["full", "roi_union"])), "bg_mode": "percentile",
              "percentile": 1.0, "per_channel_p": False, "donor_p": 1.0, "fret_p": 1.0, "clip_neg": bool(rng.integers(0, 2)), "eps_percentile": 1.0,
              "px_um": None, "out_xls": False, "out_tif": False, "out_png": False, "save_full": False, "save_crop": False, "mask_outside": True,
              "apply_cmap": True, "cmap_name": "jet", "show_colorbar": False, "png_dpi": 300, "add_scalebar": False, "scale_bar_um": 20.0,
              "cmin_txt": "", "cmax_txt": "", "fixed_crop": True, "crop_w": 500, "crop_h": 500, "subset_on": False, "subset_stage": "",
              "subset_time": "", "subset_roi": "", "n_workers": 1, "lang": "en"}
        pairs_all, _ = mR.build_pairs_by_channel(common.list_tifs(root), True, 1, 2)
        for stage in ("S01", "S02", "S03"):
            pairs = [pr for pr in pairs_all if pr[0][0] == stage]
            paths = (root, None, None, None, None, None, None)
            try:
                with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                    _, wrows, _ = rR.process_one_stage(stage, pairs, dict(pf), paths)
            except Exception as e:                      # the reference lets an unreadable image fail the whole stage
                try:
                    mR.process_one_stage(stage, pairs, dict(pf), paths, frames_per_batch=int(rng.choice([1, 2, 3, 5])))
                    stats["ref_stage_fails_ours_continues"] = stats.get("ref_stage_fails_ours_continues", 0) + 1
                except Exception:
                    stats["both_fail_stage"] = stats.get("both_fail_stage", 0) + 1
                continue
            _, grows, _ = mR.process_one_stage(stage, pairs, dict(pf), paths, frames_per_batch=int(rng.choice([1, 2, 3, 5])))
            assert [(g["stage"], g["time"], g["roi"]) for g in grows] == [(w["stage"], w["time"], w["roi"]) for w in wrows], ("FRET order", stage, len(grows), len(wrows))
            for g, w in zip(grows, wrows):
                for k, wv in w.items():
                    gv = g[k]
                    if isinstance(wv, float) and math.isnan(wv):
                        assert math.isnan(gv), k
                    elif k == "eps":
                        assert np.float32(gv) == np.float32(wv), k
                    elif isinstance(wv, float) and k.endswith(("_mean", "_std")):
                        pre = k.rsplit("_", 1)[0]
                        scale = abs(w[pre + "_median"]) + (abs(w["ratio_p5"]) + abs(w["ratio_p95"]) if pre == "ratio" else 0.0)
                        assert close(gv, wv) or abs(gv - wv) <= 1e-5 * scale, (k, gv, wv)
                    else:
                        assert gv == wv, (k, gv, wv)
            stats["fret_rows"] = stats.get("fret_rows", 0) + len(wrows)
    except Exception as e:
        bad += 1
        tb = traceback.extract_tb(e.__traceback__)
        print("FAIL seed", seed, type(e).__name__, str(e)[:400], [(t.filename.split("/")[-1], t.lineno) for t in tb][-3:], flush=True)
    finally:
        shutil.rmtree(root, ignore_errors=True)
print("done", seed0, n, "bad", bad, stats, round(time.time() - t0, 1), flush=True)
