"""tests/fuzz/fuzz_route_a_nesprin2.py -- TEST INFRASTRUCTURE, run by hand, BUILD CONTAINER ONLY (needs /root/reference).

Nesprin2_FRET_Builder.run_pipeline(p) of the UNMODIFIED reference (reference Nesprin2_FRET_Builder.py:1331; loaded through
oracle/refimport.py on the restated shims) and of the mirror on the same random folder (donor / FRET / acceptor-only
triples + ROI JSONs) with random parameter dicts; file outputs limited to the table: nesprin2_fret_perROI.csv of both
runs must agree cell for cell (text-equal except means / standard deviations, which agree to 1e-5 of the values' scale).
Static mode only: in time-lapse mode the reference stores a function in its "time" column (SURVEY.md 8(a) defect 1) and
its writer fails.

    python tests/fuzz/fuzz_route_a_nesprin2.py <first seed> <number of seeds>
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import csv, json, math, shutil, tempfile, time, traceback, io, contextlib
import numpy as np
from oracle import refimport
import imageprocess_b200 as ipb
from imageprocess_b200.ops import Engine
from imageprocess_b200.host import Nesprin2_FRET_Builder as mN, common
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests.checks import close
stats = {"rows": 0}
BASE = {"out_root": "", "timelapse": False, "donor_ch": 2, "fret_ch": 3, "intensity_ch": 1, "ratio_mode": "FRET/Donor", "bg_scope": "full",
        "bg_mode": "percentile", "percentile": 1.0, "per_channel_p": False, "donor_p": 1.0, "fret_p": 1.0, "clip_neg": True, "eps_percentile": 1.0,
        "px_um": 0.223, "rim_um": 1.12, "annulus_on": False, "ann_in_um": 1.2, "ann_out_um": 2.5, "use_spectral": False, "alpha": 0.0, "beta": 0.0,
        "g_factor": 1.0, "aonly_ch": None, "scale_preset": "auto", "fret_min": 0.0, "fret_max": 1.5, "cmap_name": "turbo", "show_colorbar": False,
        "add_scalebar": False, "scale_bar_um": 20.0, "out_xls": True, "out_tif": False, "out_png": False, "save_panel": False, "save_full": False,
        "save_crop": False, "save_crop_intensity": False, "mask_outside": True, "crop_fixed": True, "crop_w": 500, "crop_h": 500, "crop_vmin_txt": "",
        "crop_vmax_txt": "", "subset_on": False, "subset_stage": None, "subset_time": None, "sat_filter_on": True, "sat_threshold": 65535.0,
        "clip_ratio_on": True, "clip_ratio_max": 20.0}

def run_seed(seed, rN):
    """One random folder through both run_pipelines; returns the number of table rows compared (raises on a difference)."""
    rng = np.random.default_rng(seed)
    root = tempfile.mkdtemp(prefix="ipb_fuzz_n2_")
    try:
        return _run_seed(rng, root, rN)
    finally:
        shutil.rmtree(root, ignore_errors=True)


def _run_seed(rng, root, rN):
    H, W = int(rng.integers(40, 110)), int(rng.choice([8 * int(rng.integers(5, 16)), int(rng.integers(41, 130))]))
    roi_dir = os.path.join(root, "roi")
    os.makedirs(roi_dir)
    for s in (1, 2):
        stem = f"S{s:02d}"
        d = rng.poisson(float(rng.choice([300, 2000])), (H, W)).astype(np.int64)
        a = rng.poisson(float(rng.choice([200, 1500])), (H, W)).astype(np.int64)
        y, x = int(rng.integers(0, H - 10)), int(rng.integers(0, W - 10))
        d[y:y + 25, x:x + 30] += 2500; a[y:y + 25, x:x + 30] += 1800
        d = np.minimum(d, 65535).astype(np.uint16); a = np.minimum(a, 65535).astype(np.uint16)
        d[rng.random((H, W)) < 0.003] = 65535
        ao = (0.3 * a + rng.poisson(50, (H, W))).astype(np.uint16)
        for ch, im in ((1, ao), (2, d), (3, a), (4, ao)):
            common.write_tiff(os.path.join(root, f"{stem}_{ch}.tif"), im)
        polys = []
        for k in range(int(rng.integers(1, 4))):
            cx, cy, r = rng.uniform(8, W - 8), rng.uniform(8, H - 8), rng.uniform(4, 25)
            tt = np.sort(rng.uniform(0, 2 * np.pi, int(rng.integers(4, 12))))
            polys.append(np.stack([cx + r * np.cos(tt), cy + r * np.sin(tt)], axis=1).tolist())
        with open(os.path.join(roi_dir, stem + ".json"), "w") as fh:
            json.dump({"name": stem, "image_shape": {"height": H, "width": W}, "rois": polys}, fh)
    p = dict(BASE, img_dir=root, roi_dir=roi_dir)
    p.update({"ratio_mode": str(rng.choice(["FRET/Donor", "Donor/FRET"])), "bg_scope": str(rng.choice(["full", "roi_union", "annulus"])),
              "bg_mode": str(rng.choice(["percentile", "percentile", "hist-mode"])), "percentile": float(rng.choice([0.5, 1.0, 5.0])),
              "per_channel_p": bool(rng.integers(0, 2)), "donor_p": float(rng.choice([0.5, 1.0])), "fret_p": float(rng.choice([1.0, 3.0])),
              "clip_neg": bool(rng.integers(0, 2)), "eps_percentile": float(rng.choice([0.0, 1.0, 5.0])), "rim_um": float(rng.choice([0.3, 1.12, 2.0])),
              "annulus_on": bool(rng.integers(0, 2)), "ann_in_um": float(rng.choice([0.0, 0.5, 1.2])), "ann_out_um": float(rng.choice([1.4, 2.5])),
              "use_spectral": bool(rng.integers(0, 2)), "alpha": float(rng.choice([0.0, 0.12])), "beta": float(rng.choice([0.0, 0.05])),
              "g_factor": float(rng.choice([1.0, 1.1])), "aonly_ch": int(rng.choice([4, 4, 9])) if rng.random() < 0.6 else None,
              "sat_filter_on": bool(rng.integers(0, 2)), "sat_threshold": float(rng.choice([65535.0, 30000.0])),
              "clip_ratio_on": bool(rng.integers(0, 2)), "clip_ratio_max": float(rng.choice([3.0, 20.0]))})
    outs = {}
    for who, fn in (("ref", lambda q: rN.run_pipeline(q)), ("ours", lambda q: mN.run_pipeline(q, log=lambda s: None))):
        q = dict(p, out_root=os.path.join(root, "RES_" + who))
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            fn(q)
        path = os.path.join(root, "RES_" + who, "xls", "nesprin2_fret_perROI.csv")
        outs[who] = list(csv.reader(open(path, newline=""))) if os.path.exists(path) else None
    assert (outs["ref"] is None) == (outs["ours"] is None), ("one side wrote no table", outs["ref"] is None, outs["ours"] is None)
    if outs["ref"] is None:
        return 0
    assert outs["ours"][0] == outs["ref"][0], ("header", outs["ours"][0], outs["ref"][0])
    assert len(outs["ours"]) == len(outs["ref"]), ("rows", len(outs["ours"]), len(outs["ref"]))
    hdr = outs["ref"][0]
    for g, w in zip(outs["ours"][1:], outs["ref"][1:]):
        rowd = dict(zip(hdr, w))
        scale = max(abs(float(rowd[k])) for k in ("ratio_median", "ratio_p5", "ratio_p95") if rowd[k] not in ("", "nan")) if any(
            rowd[k] not in ("", "nan") for k in ("ratio_median", "ratio_p5", "ratio_p95")) else 0.0
        for k, gv, wv in zip(hdr, g, w):
            if gv == wv:
                continue
            if {gv, wv} == {"0.0", "-0.0"}:             # a tie between +0.0 and -0.0 at the wanted rank: equal values, numpy's
                stats["signed_zero"] = stats.get("signed_zero", 0) + 1      # partition picks either; only the text differs
                continue
            if k.endswith(("_mean", "_std")):
                sc = scale if k.startswith("ratio") else 1e9
                assert close(float(gv), float(wv)) or abs(float(gv) - float(wv)) <= 1e-5 * sc, (k, gv, wv)
            else:
                raise AssertionError((k, gv, wv, rowd["stage"], rowd["roi"]))
    return len(outs["ref"]) - 1


def main():
    ipb._engine = Engine(emu_lib(), NumpyMem())
    refimport.install_stubs()
    rN = refimport.load("Nesprin2_FRET_Builder")
    seed0, n = int(sys.argv[1]), int(sys.argv[2])
    bad, t0 = 0, time.time()
    for seed in range(seed0, seed0 + n):
        try:
            stats["rows"] += run_seed(seed, rN)
        except Exception as e:
            bad += 1
            tb = traceback.extract_tb(e.__traceback__)
            print("FAIL seed", seed, type(e).__name__, str(e)[:400], [(t.filename.split("/")[-1], t.lineno) for t in tb][-3:], flush=True)
    print("done", seed0, n, "bad", bad, stats, round(time.time() - t0, 1), flush=True)


if __name__ == "__main__":
    main()
