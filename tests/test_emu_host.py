"""CPU tier of the host mirrors: the reference-shaped entry points driven end to end (TIFF + ROI
JSON folders in, CSV / TIFF out) with the kernels running in the emulated build."""
import pytest

from imageprocess_b200.ops import Engine
from tests import checks_host
from tests.emu.emu_backend import NumpyMem, emu_lib


@pytest.fixture(scope="module")
def eng():
    return Engine(emu_lib(), NumpyMem())


@pytest.mark.parametrize("fn", checks_host.HOST_CHECKS + checks_host.LATE_HOST_CHECKS, ids=lambda f: f.__name__)
def test_host(eng, fn, tmp_path):
    fn(eng, str(tmp_path))


@pytest.mark.parametrize("config", ["c1", "c2"])
def test_bench_fixture_workloads(eng, config):
    """`bench.py --config c1 | c2` (BASELINE configs 1 and 2) build and run one step on the emulated
    engine: the shipped intensity frames give the 18 + 11 ROI rows of the golden CSVs, the shipped FA
    outlines give 16 cell crops with adhesions."""
    import types

    import bench
    wl = bench.other_workload(types.SimpleNamespace(config=config, frames=64), eng)
    assert wl["step"]() > 0
    if config == "c1":
        assert wl["seen"]["roi_rows"] == 29 and wl["px"] == 2 * 1536 * 2048 and wl["h2d"] == 4 * wl["px"]
    else:
        assert wl["seen"]["cell_crops"] == 16 and wl["seen"]["adhesions"] >= 100 and wl["px"] == 4 * 2200 * 3200
    assert wl["step_e2e"]() > 0
