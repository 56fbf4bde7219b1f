"""Pins the oracle against every golden the reference ships for the hot path
(SURVEY.md 8(c)): 29 ROI rows of fluor_intensity_perROI.csv (matplotlib rule + bg +
stats) and the two roi/mask/S01_mask.tif (skimage rule)."""
import numpy as np
import pytest

from oracle import port, shims
from tests import goldenio

TASK = {"bg_scope": "full", "bg_mode": "percentile", "clip_neg": True, "bg_stride": 4,
        "percentile": 1.0, "per_channel_p": False, "ch_p_map": {}}


@pytest.mark.parametrize("exp,n_roi", [("e1_P0", 18), ("e2_P1", 11)])
def test_intensity_csv_rows(exp, n_roi):
    imgs, polys, rows, _ = goldenio.load_intensity(exp)
    assert len(polys) == n_roi == len(rows)
    raw = {ch: a.astype(np.float32) for ch, a in imgs.items()}
    per_roi, bg_used, _ = port.int_process_key(raw, polys, None, TASK)
    for ch in (2, 3):
        assert bg_used[ch]["bg"] == float(rows[0][f"ch{ch}_bg"])
    for got, exp_row in zip(per_roi, rows):
        assert got["roi"] == int(exp_row["roi"])
        assert got["area_px"] == int(exp_row["area_px"])          # integer: bit-exact
        for ch in (2, 3):
            for k in goldenio.STAT_KEYS:
                g, e = got[f"ch{ch}_{k}"], float(exp_row[f"ch{ch}_{k}"])
                if k == "npx":
                    assert g == int(e)
                else:
                    assert g == pytest.approx(e, rel=1e-12, abs=0), (ch, k)


@pytest.mark.parametrize("exp", ["e1_P0", "e2_P1"])
def test_skimage_polygon_mask_tif(exp):
    imgs, polys, _, mask = goldenio.load_intensity(exp)
    H, W = mask.shape
    got = np.zeros((H, W), dtype=bool)
    for P in polys:                       # roi_manual_drawer.py:1332-1340
        rr, cc = shims.polygon(P[:, 1], P[:, 0], (H, W))
        got[rr, cc] = True
    assert int((got ^ mask).sum()) == 0


def test_rules_are_not_interchangeable():
    """matplotlib and skimage rules differ on e2_P1 (so both are needed)."""
    _, polys, _, mask = goldenio.load_intensity("e2_P1")
    H, W = mask.shape
    u = np.zeros((H, W), dtype=bool)
    for P in polys:
        u |= port.rasterize_polygon(P, (H, W))
    assert int((u ^ mask).sum()) > 0


def test_fa_csv_schema():
    import json, os
    sch = json.load(open(os.path.join(goldenio.GOLD, "fa_csv_schema.json")))
    hdr = ("File,Cell_ID,Category,Area_px,Area_um2,Mean_Intensity_Raw,Mean_Intensity_Corr,"
           "Int_Density_Raw,Int_Density_Corr,Background_Level,Used_Alpha,Global_Threshold,"
           "Min_Area_Setting,Max_Area_Setting,Close_Radius_Setting,Subtract_BG_Setting")
    n = 0
    for v in sch.values():
        assert v["header"] == hdr
        n += len(v["rows"])
    assert n == 156
    # dtype evidence: Int_Density_Raw == float64(float32 mean) * float64(area)
    first = sch["e1/S01"]["rows"][0].split(",")
    assert float(first[7]) == float(np.float32(first[5])) * float(first[3])
