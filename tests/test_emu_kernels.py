"""CPU tier: the product's kernel sources built with the CUDA emulator (tests/emu) and
driven by the product's host code, checked against the oracle.  Logic only -- the parity
tests proper are tests/test_gpu_kernels.py on the B200."""
import pytest

from imageprocess_b200.ops import Engine
from tests import checks
from tests.emu.emu_backend import NumpyMem, emu_lib


@pytest.fixture(scope="module")
def eng():
    return Engine(emu_lib(), NumpyMem())


@pytest.mark.parametrize("fn", checks.RASTER_CHECKS + checks.LATE_CHECKS, ids=lambda f: f.__name__)
def test_raster(eng, fn):
    fn(eng)


@pytest.mark.parametrize("scope,stride,mode", checks.INTENSITY_CASES)
def test_intensity_batch(eng, scope, stride, mode):
    checks.check_intensity_batch(eng, scope, stride, mode)


@pytest.mark.parametrize("ratio_mode,scope,clip", checks.FRET_CASES)
def test_fret_batch(eng, ratio_mode, scope, clip):
    checks.check_fret_batch(eng, ratio_mode, scope, clip)


@pytest.mark.parametrize("fa_path", [1, 2, 3], ids=["fused-smem", "phases", "fused-global"])
@pytest.mark.parametrize("params", checks.FA_CASES, ids=lambda p: f"a{p['alpha']}_r{p['close_radius']}")
def test_fa_batch(eng, params, fa_path):
    checks.check_fa_batch(eng, params, fa_path=fa_path)


@pytest.mark.parametrize("exp", ["e1_P0", "e2_P1"])
def test_intensity_golden(eng, exp):
    checks.check_intensity_golden(eng, exp)


@pytest.mark.parametrize("case", checks.N2_CASES, ids=lambda c: "_".join(sorted(c)) or "default")
def test_nesprin2_batch(eng, case):
    checks.check_nesprin2_batch(eng, case)
