"""Kernel-logic tests on the CPU: the product's rasteriser sources (ipb_raster.cuh) built
with the CUDA emulator (tests/emu) against the oracle.  The GPU tier (test_gpu_*.py) repeats
these through the real library."""
import numpy as np
import pytest

from imageprocess_b200 import geometry as geo
from imageprocess_b200.ops import Engine
from oracle import port, shims
from oracle.gen_golden import small_scene
from tests import goldenio
from tests.emu.emu_backend import NumpyMem, emu_lib


@pytest.fixture(scope="module")
def eng():
    return Engine(emu_lib(), NumpyMem())


def _check_mpl(eng, polys, H, W):
    specs = [geo.mpl_spec(P, (W, H)) for P in polys]
    rm = eng.rasterize(geo.RULE_MPL, specs, (H, W), 1, want_union=True)
    area = rm.area.host()
    union = np.zeros((H, W), bool)
    for i, P in enumerate(polys):
        want = port.rasterize_polygon(P, (H, W))
        x0, y0, x1, y1 = specs[i].srect
        got = np.zeros((H, W), bool)
        got[y0:y1, x0:x1] = rm.mask_host(i)
        assert int((got ^ want).sum()) == 0, i
        assert int(area[i]) == int(want.sum())
        union |= want
    assert np.array_equal(rm.union_host()[0], union)


def test_mpl_small_scene(eng):
    d, a, polys = small_scene(7)
    _check_mpl(eng, polys, *d.shape)


def test_mpl_random_polygons(eng):
    rng = np.random.default_rng(3)
    H, W = 70, 150
    polys = []
    for k in range(12):
        n = int(rng.integers(3, 12))
        P = rng.uniform(-10, 160, (n, 2))
        P[:, 1] = rng.uniform(-10, 80, n)
        if k % 3 == 0:
            P = np.round(P * 2) / 2        # .5 grid: vertices on pixel centres / edges
        if k % 4 == 1:
            P = np.round(P)                # integer vertices: ties with pixel centres
        polys.append(P)
    polys.append(np.array([[5.0, 5.0], [140.0, 5.0], [140.0, 60.0], [5.0, 60.0]]))   # long flat edges
    polys.append(np.array([[0.0, 10.0], [149.0, 10.5], [149.0, 12.0], [0.0, 11.0]]))  # long shallow edges
    polys.append(np.array([[10.0, 10.0], [20.0, 10.0], [20.0, 20.0], [10.0, 20.0], [10.0, 10.0]]))  # closed
    _check_mpl(eng, polys, H, W)


def test_sk_crops_match_oracle(eng):
    d, a, polys = small_scene(7)
    img = d.astype(np.float32)
    specs, wants = [], []
    for P in polys:
        spec, rect = geo.fa_spec(P, img.shape)
        crop, mask, rect2 = port.fa_crop_and_mask(img, P.copy())
        assert rect == rect2
        specs.append(spec)
        wants.append(mask)
    rm = eng.rasterize(geo.RULE_SK, specs, img.shape, 1, want_union=True)
    area = rm.area.host()
    for i, want in enumerate(wants):
        assert np.array_equal(rm.mask_host(i), want), i
        assert int(area[i]) == int(want.sum())


def test_sk_random_polygons(eng):
    rng = np.random.default_rng(9)
    H, W = 60, 90
    specs, wants = [], []
    for k in range(14):
        n = int(rng.integers(3, 10))
        P = np.stack([rng.uniform(-8, 98, n), rng.uniform(-8, 68, n)], axis=1)
        if k % 2 == 0:
            P = np.round(P * 2) / 2
        if k % 5 == 1:
            P = np.round(P)
        specs.append(geo.sk_spec(P[:, 1], P[:, 0], (H, W)))
        m = np.zeros((H, W), bool)
        rr, cc = shims.polygon(P[:, 1], P[:, 0], (H, W))
        m[rr, cc] = True
        wants.append(m)
    rm = eng.rasterize(geo.RULE_SK, specs, (H, W), 1, want_union=False)
    for i, want in enumerate(wants):
        assert int((rm.mask_host(i) ^ want).sum()) == 0, i


def test_fa_fixture_polygons_sk(eng):
    """The FA sample's 62-540 vertex ROI polygons (2200x3200) -- one of them, both rules."""
    shape, polys = goldenio.load_fa_rois()["e2/S02"]
    H, W = shape["height"], shape["width"]
    P = polys[0]
    spec, rect = geo.fa_spec(P, (H, W))
    rm = eng.rasterize(geo.RULE_SK, [spec], (H, W), 1, want_union=False)
    pc = P.copy()
    pc[:, 0] -= rect[0]
    pc[:, 1] -= rect[2]
    h, w = rect[3] - rect[2], rect[1] - rect[0]
    want = np.zeros((h, w), bool)
    rr, cc = shims.polygon(pc[:, 1], pc[:, 0], (h, w))
    want[rr, cc] = True
    assert int((rm.mask_host(0) ^ want).sum()) == 0
