"""oracle/port.py against (a) the UNMODIFIED reference functions run here through
oracle/refimport.py (skipped where /root/reference is absent) and (b) the committed
vectors those functions produced (tests/golden/ref_vectors.npz, always run)."""
import numpy as np
import pytest

from oracle import port, refimport, shims
from oracle.gen_golden import small_scene
from tests import goldenio

needs_ref = pytest.mark.skipif(not refimport.available(), reason="/root/reference not present")


def _eq(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(a, b, equal_nan=True), np.nanmax(np.abs(a - b))


# ------------------------------------------------------------------ committed vectors
def test_vectors_mpl_and_intensity():
    v = goldenio.load_ref_vectors()
    d, a, polys = small_scene(int(v["scene7_seed"][0]))
    H, W = d.shape
    masks = np.stack([port.rasterize_polygon(P, (H, W)) for P in polys])
    assert np.array_equal(np.packbits(masks), v["mpl_masks"])
    img = d.astype(np.float32)
    bc, B = port.int_bg_correct(img, "percentile", 1.0, None, True, 4)
    assert B == v["int_bg_p1_s4"][0]
    _, B2 = port.int_bg_correct(img, "hist-mode", 1.0, None, True, 4)
    assert B2 == v["int_bg_hist_s4"][0]
    rows = port.quantify_per_roi_multi({1: bc, 2: port.int_bg_correct(a.astype(np.float32))[0]},
                                       polys=polys)
    keys = list(v["int_rows_keys"])
    _eq([[r[k] for k in keys] for r in rows], v["int_rows"])


def test_vectors_fa():
    v = goldenio.load_ref_vectors()
    d, a, polys = small_scene(int(v["scene7_seed"][0]))
    img = d.astype(np.float32)
    stats = port.fa_global_stats(img)
    assert np.array_equal(np.array(stats, dtype=np.float32), v["fa_stats"])
    cfg = {'alpha': 2.0, 'min_px': 12.5, 'max_px': 400.0, 'close_radius': 1, 'subtract_bg': True}
    tab = []
    for i, P in enumerate(polys):
        crop, mask, rect = port.fa_crop_and_mask(img, P.copy())
        assert tuple(v[f"fa_rect_{i}"]) == rect
        assert np.array_equal(np.packbits(mask), v[f"fa_mask_{i}"])
        res, thr, bw, lab = port.analyze_fa_crop(crop, mask, cfg, stats)
        assert np.array_equal(np.packbits(bw), v[f"fa_bw_{i}"])
        assert np.array_equal(lab.astype(np.int32), v[f"fa_lab_{i}"])
        for ci, cat in enumerate(("OK", "Large", "Small")):
            for it in res[cat]:
                tab.append([i, ci, it["label"], it["area"], float(it["mean_int_raw"]),
                            float(it["mean_int_corr"]), it["int_den_raw"], it["int_den_corr"],
                            it["centroid"][0], it["centroid"][1], float(thr)])
    assert len(tab) > 5
    _eq(tab, v["fa_table"])


def test_vectors_fret_n2_mor():
    v = goldenio.load_ref_vectors()
    d, a, polys = small_scene(int(v["scene7_seed"][0]))
    H, W = d.shape
    p = {"bg_scope": "full", "bg_mode": "percentile", "percentile": 1.0, "per_channel_p": False,
         "donor_p": 1.0, "fret_p": 1.0, "clip_neg": True, "eps_percentile": 1.0,
         "ratio_mode": "Donor/FRET"}
    out = port.fret_process_pair(d.astype(np.float32), a.astype(np.float32), polys, p)
    _eq([out["Db"], out["Ab"], out["eps"]], v["fret_scalars"])
    assert np.array_equal(out["R_full"], v["fret_R"])
    keys = list(v["fret_rows_keys"])
    _eq([[r[k] for k in keys] for r in out["rows"]], v["fret_rows"])
    assert np.array_equal(np.packbits(port.make_inside_rim_mask(out["union"], 5)), v["n2_rim5"])
    assert np.array_equal(np.packbits(port.annulus_mask_from_poly(polys[0], (H, W), 5, 11)),
                          v["n2_ann"])
    _, yy = port.spectral_correct(out["Abc"], out["Dbc"], None, 0.12, 0.05, 1.1)
    assert np.array_equal(yy, v["n2_spec"])
    mor = [port.morphology_from_polygon(P, (H, W), 0.223) for P in polys]
    mk = list(v["mor_keys"])
    _eq([[m[k] for k in mk] for m in mor], v["mor_rows"])


# ------------------------------------------------------------------ live reference
@needs_ref
@pytest.mark.parametrize("seed", [11, 12])
def test_live_fluor_int(seed):
    F = refimport.load("Fluor_INT")
    d, a, polys = small_scene(seed, H=160, W=200)
    rng = np.random.default_rng(seed)
    img = d.astype(np.float32)
    scope = rng.random(img.shape) < 0.3
    for mode in ("percentile", "hist-mode"):
        for sm in (None, scope):
            for stride in (1, 4, 7):
                r_bc, r_B = F.bg_correct(img.copy(), mode, 2.5, sm, True, stride)
                o_bc, o_B = port.int_bg_correct(img.copy(), mode, 2.5, sm, True, stride)
                assert r_B == o_B and np.array_equal(r_bc, o_bc)
    bc = {1: F.bg_correct(img)[0], 2: F.bg_correct(a.astype(np.float32))[0]}
    assert F.quantify_per_roi_multi(bc, polys=polys) == port.quantify_per_roi_multi(bc, polys=polys)
    assert F.auto_minmax(bc[1].ravel(), 1, 99) == port.auto_minmax(bc[1].ravel(), 1, 99)


@needs_ref
@pytest.mark.parametrize("seed", [21, 22])
def test_live_fa(seed):
    FA = refimport.load("FA_Analyzer")
    d, a, polys = small_scene(seed, H=200, W=240, blobs=10)
    img = d.astype(np.float32)
    stats = port.fa_global_stats(img)
    for cfg in ({'alpha': 2.0, 'min_px': 12.5, 'max_px': 300.0, 'close_radius': 1, 'subtract_bg': True},
                {'alpha': 1.0, 'min_px': 0, 'max_px': 50.0, 'close_radius': 0, 'subtract_bg': False},
                {'alpha': 3.0, 'min_px': 30.0, 'max_px': 5000.0, 'close_radius': 2, 'subtract_bg': True}):
        for P in polys:
            crop, mask, _ = port.fa_crop_and_mask(img, P.copy())
            r = FA.analyze_fa_crop(crop, mask, cfg, stats)
            o = port.analyze_fa_crop(crop, mask, cfg, stats)
            assert r[1] == o[1] and np.array_equal(r[2], o[2]) and np.array_equal(r[3], o[3])
            for cat in ("OK", "Large", "Small"):
                assert len(r[0][cat]) == len(o[0][cat])
                for x, y in zip(r[0][cat], o[0][cat]):
                    for k in ("label", "area", "centroid", "mean_int_raw", "mean_int_corr",
                              "int_den_raw", "int_den_corr", "bg_level"):
                        assert x[k] == y[k] and type(x[k]) is type(y[k]), k


@needs_ref
def test_live_fret_and_nesprin2():
    FR = refimport.load("fret_ratio_builder")
    N2 = refimport.load("Nesprin2_FRET_Builder")
    d, a, polys = small_scene(31, H=160, W=200)
    D, A = d.astype(np.float32), a.astype(np.float32)
    H, W = D.shape
    u = np.zeros((H, W), bool)
    for P in polys:
        u |= FR.rasterize_polygon(P, (H, W))
    for mode in ("percentile", "hist-mode"):
        for sm in (None, u):
            r = FR.bg_correct(D.copy(), mode, 1.0, sm, True)
            o = port.fret_bg_correct(D.copy(), mode, 1.0, sm, True)
            assert r[1] == o[1] and np.array_equal(r[0], o[0])
    Dn = D.copy()
    Dn[D >= 65535] = np.nan
    for sm in (None, u):
        r = N2.bg_correct(Dn.copy(), "percentile", 1.0, sm, True)
        o = port.n2_bg_correct(Dn.copy(), "percentile", 1.0, sm, True)
        assert r[1] == o[1] and np.array_equal(r[0], o[0], equal_nan=True)
    assert N2.pick_epsilon(Dn[u], 5.0, 1.0) == port.n2_pick_epsilon(Dn[u], 5.0, 1.0)
    assert np.array_equal(N2.make_inside_rim_mask(u, 5), port.make_inside_rim_mask(u, 5))
    assert np.array_equal(N2.annulus_mask_from_poly(polys[0], (H, W), 5, 11),
                          port.annulus_mask_from_poly(polys[0], (H, W), 5, 11))
    Rr = FR.quantify_per_roi((D + 5) / (A + 5), polys, extra_imgs={"donor": D, "yfret": A})
    Ro = port.fret_quantify_per_roi((D + 5) / (A + 5), polys, extra_imgs={"donor": D, "yfret": A})
    assert Rr == Ro


@needs_ref
def test_live_mor():
    MOR = refimport.load("MOR_by_ROI")
    d, a, polys = small_scene(41, H=160, W=200)
    for P in polys:
        r = MOR.morphology_from_polygon(P, d.shape, 0.223)
        o = port.morphology_from_polygon(P, d.shape, 0.223)
        assert r == o


def test_label_numbering_is_raster_order():
    """shims.label must number components by first pixel in raster order
    (skimage.measure.label)."""
    rng = np.random.default_rng(5)
    for _ in range(5):
        bw = rng.random((40, 53)) < 0.45
        lab = shims.label(bw)
        firsts = [np.flatnonzero(lab.ravel() == k)[0] for k in range(1, lab.max() + 1)]
        assert firsts == sorted(firsts)
        # 8-connectivity: diagonal neighbours share a label
        assert (lab[:-1, :-1][bw[:-1, :-1] & bw[1:, 1:]] == lab[1:, 1:][bw[:-1, :-1] & bw[1:, 1:]]).all()


@needs_ref
def test_segment_inside_polygon_matches_reference():
    """oracle/port.segment_inside_polygon against the unmodified roi_manual_drawer.segment_inside_polygon
    (run on the restated matplotlib / skimage shims)."""
    ref = refimport.load("roi_manual_drawer")
    rng = np.random.default_rng(31)
    H, W = 120, 150
    img = rng.poisson(300, (H, W)).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    cell = ((xx - 70) / 40.0) ** 2 + ((yy - 60) / 28.0) ** 2 <= 1.0
    img[cell] += rng.poisson(1500, int(cell.sum())).astype(np.float32)
    img[55:60, 66:74] = 300.0
    poly = np.array([[20.0, 15.5], [130.5, 12.0], [140.0, 95.0], [75.0, 112.5], [15.5, 85.0]])
    for mode, par in (("percentile", 70.0), ("bnd", 0.5)):
        w = ref.segment_inside_polygon(img, poly, thr_param=par, min_area=40, tolerance=1.0, mode=mode)
        g = port.segment_inside_polygon(img, poly, thr_param=par, min_area=40, tolerance=1.0, mode=mode)
        assert g[0] == w[0] and g[1] is None and w[1] is None
        assert np.array_equal(g[2], w[2]) and g[2].shape[0] >= 6


# ------------------------------------------------------------------ file-name grammars
def _adversarial_names():
    S = ["S1", "S01", "s2", "S003", "S12", "S1234", "Stage1", "S", "S1a", "XS1", "S1S2", "S01 "]
    T = ["", "_t0", "_t00", "_t003", "_T7", "_t12", "t3", "_time4", "_t", "_t1_t2", "_t1234", "-t5"]
    C = ["", "_1", "_2", "_ch1", "_ch02", "_c3", "_C4", "_CH5", "_10", "_ch", "-1", "_1_2", "_DAPI", "_c1_extra", "_ch1234",
         "_03_7"]
    E = [".tif", ".TIF", ".tiff", ".png", ""]
    names = [s + t + c + e for s in S for t in T for c in C for e in E]
    return names + ["S01_t000_1.tif", "S01_t000_2.tif", "S1_t0.json", "S01.json", "image.tif", "S01_1.tif.bak", "S01__1.tif",
                    "_S01_1.tif", "S01_t01_ch2_ch3.tif", "exp-S02-t07-c3.tif", "S02_t07_007.tif", "S02_t07_07.tif"]


@needs_ref
def test_name_grammars_match_reference():
    """Every mirror parses file names with ITS script's regex set (the five sets differ): the mirrors' functions
    against the unmodified reference functions on ~12 000 (name, time-lapse) pairs, exceptions included."""
    from imageprocess_b200.host import Fluor_INT as mF, _fretnames as mN
    ref = {n: refimport.load(n) for n in ("Fluor_INT", "fret_ratio_builder", "Nesprin2_FRET_Builder", "MOR_by_ROI",
                                          "roi_channel_cropper")}

    def cropper_ref(name, tl):
        s, t = ref["roi_channel_cropper"].parse_stage_time(name, tl)
        return s, t, ref["roi_channel_cropper"].detect_channel(name, timelapse=tl)

    def cropper_ours(name, tl):
        s, t, ch = mN.parse_tokens_cropper(name, tl)
        return (f"S{s:02d}" if s is not None else None), (f"t{t:02d}" if t is not None else None), ch

    pairs = [("Fluor_INT.parse_tokens", ref["Fluor_INT"].parse_tokens, mF.parse_tokens),
             ("Fluor_INT.clean_base_for_save", ref["Fluor_INT"].clean_base_for_save, mF.clean_base_for_save),
             ("fret_ratio_builder.parse_tokens", ref["fret_ratio_builder"].parse_tokens, mN.parse_tokens),
             ("MOR_by_ROI.parse_tokens", ref["MOR_by_ROI"].parse_tokens, mN.parse_tokens),
             ("Nesprin2_FRET_Builder.parse_tokens", ref["Nesprin2_FRET_Builder"].parse_tokens, mN.parse_tokens_delimited),
             ("roi_channel_cropper.parse_stage_time + detect_channel", cropper_ref, cropper_ours)]
    args = [(n, tl) for n in _adversarial_names() for tl in (False, True)]
    for label, fr, fm in pairs:
        diffs = []
        for a in args:
            try:
                r = ("ok", fr(*a))
            except Exception as e:          # noqa: BLE001  (the exception type is part of the behaviour)
                r = ("raises", type(e).__name__)
            try:
                m = ("ok", fm(*a))
            except Exception as e:          # noqa: BLE001
                m = ("raises", type(e).__name__)
            if r != m:
                diffs.append((a, r, m))
        assert not diffs, (label, len(diffs), diffs[:4])
    # and the mirrors are wired to their own grammar
    from imageprocess_b200.host import Nesprin2_FRET_Builder as mNe, roi_channel_cropper as mC, MOR_by_ROI as mM
    assert mNe.parse_tokens is mN.parse_tokens_delimited and mC.parse_tokens is mN.parse_tokens_cropper
    assert mM.parse_tokens is mN.parse_tokens


# ------------------------------------------------------------------ table writers
class _NoExcel:
    """Stands in for pandas.ExcelWriter where no Excel engine is installed: the reference's writers open it
    unconditionally (Fluor_INT.py:754) -- only their CSV output is compared here."""

    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@needs_ref
@pytest.mark.parametrize("timelapse", [False, True])
def test_table_writers_match_reference(tmp_path, monkeypatch, timelapse):
    """fluor_intensity_perROI.csv and nesprin2_fret_perROI.csv: the UNMODIFIED reference writers
    (Fluor_INT.save_excel :728-790, Nesprin2_FRET_Builder.save_xls :1287-1326) and the mirrors' writers get the same
    rows and must produce the same bytes (column set and order, derived columns, dtypes, float text)."""
    import pandas as pd
    from imageprocess_b200.host import Fluor_INT as mF, Nesprin2_FRET_Builder as mN
    monkeypatch.setattr(pd, "ExcelWriter", _NoExcel)
    monkeypatch.setattr(pd.DataFrame, "to_excel", lambda self, *a, **k: None)
    rF, rN = refimport.load("Fluor_INT"), refimport.load("Nesprin2_FRET_Builder")
    rng = np.random.default_rng(5)
    stages = ["S01", "S02", "S10"]
    times = ["t00", "t01", "t12"] if timelapse else [None]
    rows_i, rows_n, keymap = [], [], {}
    for s in stages:
        for tm in times:
            keymap[(s, tm)] = {1: "a.tif", 2: "b.tif", 10: "c.tif"}
            for roi in (1, 2, 11):
                r = {"stage": s, "time": tm, "roi": roi, "area_px": int(rng.integers(50, 5000)), "bg_mode": "percentile",
                     "bg_scope": "full", "clip_neg": True, "bg_stride": 4}
                for ch in (1, 2, 10):                        # natural_key order: ch1, ch2, ch10
                    for k in ("mean", "median", "std", "p5", "p95", "vmin", "vmax", "vsum"):
                        r[f"ch{ch}_{k}"] = float(np.float32(rng.uniform(0, 5000)))
                    r[f"ch{ch}_npx"] = int(rng.integers(50, 5000))
                    r[f"ch{ch}_bg"], r[f"ch{ch}_p"] = float(rng.integers(50, 500)), 1.0
                rows_i.append(r)
                n = {"stage": s, "time": tm, "roi": roi, "area_px": int(rng.integers(50, 5000)), "ratio_mode": "FRET/Donor"}
                for k in ("ratio_mean", "ratio_median", "ratio_std", "ratio_p5", "ratio_p95", "ratio_FoverD_mean",
                          "ratio_DoverF_mean", "donor_mean", "fret_mean"):
                    n[k] = float(np.float32(rng.uniform(0, 3)))
                n.update({"eps": 5.0, "p": 1.0, "donor_p": 1.0, "fret_p": 1.0, "bg_scope": "roi_union", "bg_mode": "percentile",
                          "clip_neg": True, "sat_filter_on": True, "sat_threshold": 65535.0, "clip_ratio_on": False,
                          "clip_ratio_max": 10.0, "extra_column_the_writer_drops": 1})
                rows_n.append(n)
    for name, ref_fn, our_fn, rows, csv in (
            ("Fluor_INT", lambda d: rF.save_excel([dict(r) for r in rows_i], keymap, d),
             lambda d: mF.save_excel([dict(r) for r in rows_i], keymap, d, log=lambda s: None), rows_i, "fluor_intensity_perROI.csv"),
            ("Nesprin2", lambda d: rN.save_xls([dict(r) for r in rows_n], d, timelapse),
             lambda d: mN.save_xls([dict(r) for r in rows_n], d, timelapse, log=lambda s: None), rows_n, "nesprin2_fret_perROI.csv")):
        d_ref, d_our = tmp_path / f"{name}_ref", tmp_path / f"{name}_our"
        d_ref.mkdir()
        d_our.mkdir()
        ref_fn(str(d_ref))
        our_fn(str(d_our))
        want, got = (d_ref / csv).read_bytes(), (d_our / csv).read_bytes()
        assert want.count(b"\n") == len(rows) + 1
        assert got == want, (name, got[:300], want[:300])
