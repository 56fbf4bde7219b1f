"""CPU tier of the host mirrors: the reference-shaped entry points driven end to end (TIFF + ROI
JSON folders in, CSV / TIFF out) with the kernels running in the emulated build."""
import pytest

from imageprocess_b200.ops import Engine
from tests import checks_host
from tests.emu.emu_backend import NumpyMem, emu_lib


@pytest.fixture(scope="module")
def eng():
    return Engine(emu_lib(), NumpyMem())


@pytest.mark.parametrize("fn", checks_host.HOST_CHECKS, ids=lambda f: f.__name__)
def test_host(eng, fn, tmp_path):
    fn(eng, str(tmp_path))
