"""GPU tier of the host mirrors: same checks as tests/test_emu_host.py through libipb200.so."""
import pytest

from tests import checks_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import imageprocess_b200 as ipb
    return ipb.engine("cuda:0")


@pytest.mark.parametrize("fn", checks_host.HOST_CHECKS, ids=lambda f: f.__name__)
def test_host(eng, fn, tmp_path):
    fn(eng, str(tmp_path))
