"""GPU tier (B200): parity of the CUDA path, called through the C ABI of libipb200.so,
against the oracle on seeded inputs and against the reference's shipped golden."""
import pytest

from tests import checks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import imageprocess_b200 as ipb
    return ipb.engine("cuda:0")


@pytest.mark.parametrize("fn", checks.RASTER_CHECKS, ids=lambda f: f.__name__)
def test_raster(eng, fn):
    fn(eng)


@pytest.mark.parametrize("scope,stride,mode", checks.INTENSITY_CASES)
def test_intensity_batch(eng, scope, stride, mode):
    checks.check_intensity_batch(eng, scope, stride, mode)


@pytest.mark.parametrize("ratio_mode,scope,clip", checks.FRET_CASES)
def test_fret_batch(eng, ratio_mode, scope, clip):
    checks.check_fret_batch(eng, ratio_mode, scope, clip)


@pytest.mark.parametrize("exp", ["e1_P0", "e2_P1"])
def test_intensity_golden(eng, exp):
    checks.check_intensity_golden(eng, exp)


@pytest.mark.parametrize("fa_path", [1, 2, 3], ids=["fused-smem", "phases", "fused-global"])
@pytest.mark.parametrize("params", checks.FA_CASES, ids=lambda p: f"a{p['alpha']}_r{p['close_radius']}")
def test_fa_batch(eng, params, fa_path):
    checks.check_fa_batch(eng, params, fa_path=fa_path)


@pytest.mark.parametrize("case", checks.N2_CASES, ids=lambda c: "_".join(sorted(c)) or "default")
def test_nesprin2_batch(eng, case):
    checks.check_nesprin2_batch(eng, case)
