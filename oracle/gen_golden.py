"""oracle/gen_golden.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Regenerates tests/golden/ from the reference itself.  Run in the build container
(needs /root/reference):   python -m oracle.gen_golden

  tests/golden/intensity/<exp>/ch{2,3}.u16.xz  the shipped S01_{2,3}.TIF pixels (uint16,
                                               byte planes, lzma) -- inputs of the only
                                               complete golden the reference ships
  tests/golden/intensity/<exp>/rois.json       polygons of roi/S01.json
  tests/golden/intensity/<exp>/expected.csv    the reference's shipped
                                               RES/xls/fluor_intensity_perROI.csv
  tests/golden/intensity/<exp>/mask.bits.xz    roi/mask/S01_mask.tif (> 0), packed bits
  tests/golden/fa_rois.json                    polygons of the FA sample's four ROI JSONs
  tests/golden/fa_csv_schema.json              column names + first rows of the FA CSVs
  tests/golden/ref_vectors.npz                 outputs of the UNMODIFIED reference functions
                                               (oracle.refimport) on small seeded inputs
"""
import json
import lzma
import os
import shutil

import numpy as np
from PIL import Image

from . import refimport
from imageprocess_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
INT_BASE = os.path.join(refimport.REF_ROOT, "Testsamples", "1Flu_Intensity(BCC P0 and P1)", "ANA")
FA_BASE = os.path.join(refimport.REF_ROOT, "Testsamples", "2FA_BND_INT (251127 FA test sample)",
                       "Python", "ANA")


def pack_u16(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    planes = (a >> 8).astype(np.uint8).tobytes() + (a & 0xFF).astype(np.uint8).tobytes()
    return lzma.compress(planes, preset=6)


def intensity_fixtures():
    for exp in ("e1_P0", "e2_P1"):
        src = os.path.join(INT_BASE, exp)
        dst = os.path.join(GOLD, "intensity", exp)
        os.makedirs(dst, exist_ok=True)
        for ch in (2, 3):
            a = np.array(Image.open(os.path.join(src, f"S01_{ch}.TIF")))
            assert a.dtype == np.uint16 and a.shape == (1536, 2048)
            with open(os.path.join(dst, f"ch{ch}.u16.xz"), "wb") as f:
                f.write(pack_u16(a))
        with open(os.path.join(src, "roi", "S01.json"), "r", encoding="utf-8") as f:
            d = json.load(f)
        with open(os.path.join(dst, "rois.json"), "w") as f:
            json.dump({"image_shape": d["image_shape"], "rois": d["rois"]}, f)
        shutil.copyfile(os.path.join(src, "RES", "xls", "fluor_intensity_perROI.csv"),
                        os.path.join(dst, "expected.csv"))
        m = np.array(Image.open(os.path.join(src, "roi", "mask", "S01_mask.tif"))) > 0
        with open(os.path.join(dst, "mask.bits.xz"), "wb") as f:
            f.write(lzma.compress(np.packbits(m).tobytes(), preset=6))


def fa_fixtures():
    out = {}
    schema = {}
    for exp in ("e1", "e2"):
        for s in ("S01", "S02"):
            with open(os.path.join(FA_BASE, exp, "roi", f"{s}.json")) as f:
                d = json.load(f)
            out[f"{exp}/{s}"] = {"image_shape": d["image_shape"], "rois": d["rois"]}
            csv = os.path.join(FA_BASE, exp, "BND_FA", "individual_results", f"{s}_results.csv")
            with open(csv) as f:
                lines = f.read().splitlines()
            schema[f"{exp}/{s}"] = {"header": lines[0], "rows": lines[1:]}
    with open(os.path.join(GOLD, "fa_rois.json"), "w") as f:
        json.dump(out, f)
    with open(os.path.join(GOLD, "fa_csv_schema.json"), "w") as f:
        json.dump(schema, f)


def small_scene(seed, H=192, W=256, n_cells=3, blobs=9):
    d, a, polys = synth.fret_frame(seed=seed, H=H, W=W, n_cells=n_cells, r_min=22, r_max=40,
                                   blobs_per_cell=blobs, sat_frac=2e-4,
                                   blob_area=(14, 70))
    return d, a, polys


def ref_vectors():
    """Unmodified reference functions on small seeded inputs."""
    F = refimport.load("Fluor_INT")
    FA = refimport.load("FA_Analyzer")
    FR = refimport.load("fret_ratio_builder")
    N2 = refimport.load("Nesprin2_FRET_Builder")
    MOR = refimport.load("MOR_by_ROI")
    vec = {}
    d, a, polys = small_scene(7)
    H, W = d.shape
    vec["scene7_seed"] = np.array([7])
    # a1: matplotlib-rule masks
    vec["mpl_masks"] = np.packbits(np.stack([F.rasterize_polygon(P, (H, W)) for P in polys]))
    # a4/a5: bg_correct + quantify_stats
    img = d.astype(np.float32)
    bc, B = F.bg_correct(img, mode="percentile", p=1.0, scope_mask=None, clip_neg=True, stride=4)
    vec["int_bg_p1_s4"] = np.array([B])
    bc2, B2 = F.bg_correct(img, mode="hist-mode", p=1.0, scope_mask=None, clip_neg=True, stride=4)
    vec["int_bg_hist_s4"] = np.array([B2])
    rows = F.quantify_per_roi_multi({1: bc, 2: F.bg_correct(a.astype(np.float32))[0]}, polys=polys)
    keys = sorted(k for k in rows[0] if k != "roi")
    vec["int_rows_keys"] = np.array(keys)
    vec["int_rows"] = np.array([[r[k] for k in keys] for r in rows], dtype=np.float64)
    # a7: analyze_fa_crop on each ROI crop, reference crop convention
    img_f = d.astype(np.float32)
    stats = (np.nanmean(img_f), np.nanstd(img_f), np.percentile(img_f[::10, ::10], 1.0))
    vec["fa_stats"] = np.array(stats, dtype=np.float32)
    cfg = {'alpha': 2.0, 'min_px': 12.5, 'max_px': 400.0, 'close_radius': 1, 'subtract_bg': True}
    fa_tab = []
    for i, P in enumerate(polys):
        xs, ys = P[:, 0], P[:, 1]
        x0, x1 = max(0, int(np.floor(xs.min())) - 5), min(W, int(np.ceil(xs.max())) + 5)
        y0, y1 = max(0, int(np.floor(ys.min())) - 5), min(H, int(np.ceil(ys.max())) + 5)
        crop = img_f[y0:y1, x0:x1]
        pc = P.copy()
        pc[:, 0] -= x0
        pc[:, 1] -= y0
        mask = np.zeros(crop.shape, bool)
        rr, cc = FA.polygon(pc[:, 1], pc[:, 0], crop.shape)
        mask[rr, cc] = True
        res, thr, bw, lab = FA.analyze_fa_crop(crop, mask, cfg, stats)
        vec[f"fa_mask_{i}"] = np.packbits(mask)
        vec[f"fa_bw_{i}"] = np.packbits(bw)
        vec[f"fa_lab_{i}"] = lab.astype(np.int32)
        vec[f"fa_rect_{i}"] = np.array([x0, x1, y0, y1])
        for cat in ("OK", "Large", "Small"):
            for it in res[cat]:
                fa_tab.append([i, ("OK", "Large", "Small").index(cat), it["label"], it["area"],
                               float(it["mean_int_raw"]), float(it["mean_int_corr"]),
                               it["int_den_raw"], it["int_den_corr"], it["centroid"][0],
                               it["centroid"][1], float(thr)])
    vec["fa_table"] = np.array(fa_tab, dtype=np.float64)
    # a9/a14: general FRET numeric body (functions of the unmodified module)
    D, A = d.astype(np.float32), a.astype(np.float32)
    Dbc, Db = FR.bg_correct(D, mode="percentile", p=1.0, scope_mask=None, clip_neg=True)
    Abc, Ab = FR.bg_correct(A, mode="percentile", p=1.0, scope_mask=None, clip_neg=True)
    eps = FR.pick_epsilon(Abc.ravel(), eps_abs=5.0, p_floor=1.0)
    R = (Dbc + eps) / (Abc + eps)
    rows = FR.quantify_per_roi(R, polys, extra_imgs={"donor": Dbc, "yfret": Abc})
    keys = sorted(k for k in rows[0] if k != "roi")
    vec["fret_scalars"] = np.array([Db, Ab, eps])
    vec["fret_R"] = R.astype(np.float32)
    vec["fret_rows_keys"] = np.array(keys)
    vec["fret_rows"] = np.array([[r[k] for k in keys] for r in rows], dtype=np.float64)
    # a10-a12: Nesprin2 helpers
    u = np.zeros((H, W), bool)
    for P in polys:
        u |= N2.rasterize_polygon(P, (H, W))
    vec["n2_rim5"] = np.packbits(N2.make_inside_rim_mask(u, 5))
    vec["n2_ann"] = np.packbits(N2.annulus_mask_from_poly(polys[0], (H, W), 5, 11))
    dd, yy = N2.spectral_correct(Abc, Dbc, acceptor_only=None, alpha=0.12, beta=0.05, g_factor=1.1)
    vec["n2_spec"] = yy.astype(np.float32)
    # a17: morphology
    mor = [MOR.morphology_from_polygon(P, (H, W), 0.223) for P in polys]
    mkeys = sorted(mor[0])
    vec["mor_keys"] = np.array(mkeys)
    vec["mor_rows"] = np.array([[m[k] for k in mkeys] for m in mor], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "ref_vectors.npz"), **vec)


def main():
    if not refimport.available():
        raise SystemExit("reference not available")
    os.makedirs(GOLD, exist_ok=True)
    intensity_fixtures()
    fa_fixtures()
    ref_vectors()
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
