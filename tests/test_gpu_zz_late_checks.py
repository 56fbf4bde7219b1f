"""GPU tier, last file of the suite on purpose: parity checks added in round 2 after the round's last GPU session
(validated on the emulated build of the same kernel sources only).  `pytest -x` reaches them after every test that
has already run on a B200."""
import pytest

from tests import checks, checks_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import imageprocess_b200 as ipb
    return ipb.engine("cuda:0")


@pytest.mark.parametrize("fn", checks.LATE_CHECKS, ids=lambda f: f.__name__)
def test_late_check(eng, fn):
    fn(eng)


@pytest.mark.parametrize("fn", checks_host.LATE_HOST_CHECKS, ids=lambda f: f.__name__)
def test_late_host_check(eng, fn, tmp_path):
    fn(eng, str(tmp_path))
