"""Race smoke test on the CPU tier ("race detection": compute-sanitizer is not
available on the GPU pool).  The emulator runs the threads of a block one after the other between two
barriers / warp collectives, so a missing barrier between a producer and a consumer phase is invisible
whenever the producer happens to run first.  Here the same parity checks run with the threads of every
block taken in REVERSE order, and in a fresh RANDOM order on every scheduling sweep with a thread giving
up its turn after every atomic (tests/emu/cuda_emu.cpp: IPB_EMU_ORDER); the blocks of a grid run backwards
as well.  Correctly synchronised kernels give bit-identical results under every order, so the checks'
comparisons with the oracle must hold unchanged."""
import pytest

from imageprocess_b200.ops import Engine
from tests import checks
from tests.emu.emu_backend import NumpyMem, emu_lib

ORDERS = ["reverse", "random:5:preempt"]
FAST_CHECKS = [checks.check_edge_cases, checks.check_fa_overflow, checks.check_fa_wide_crop, checks.check_graph_replay,
               checks.check_hist_select_paths, checks.check_combined_batch_shared_rois,
               checks.check_region_stats_ties, checks.check_region_stats_two_views, checks.check_region_stats_streaming,
               checks.check_rim_mask, checks.check_square_dilation, checks.check_region_moments,
               checks.check_preview_and_crop, checks.check_gaussian_filters, checks.check_tophat_and_otsu,
               checks.check_segment_inside_polygon, checks.check_fa_row_refetch]


@pytest.fixture(scope="module")
def eng():
    return Engine(emu_lib(), NumpyMem())


@pytest.fixture(params=ORDERS)
def order(request, monkeypatch):
    monkeypatch.setenv("IPB_EMU_ORDER", request.param)      # read by the emulator at every launch
    return request.param


# the slow checks run under ONE of the two orders each (alternating), the quick ones under both
SLOW = {"check_combined_batch_shared_rois", "check_fa_overflow", "check_fa_wide_crop", "check_hist_select_paths",
        "check_edge_cases", "check_segment_inside_polygon"}
CASES = [(o, fn) for k, fn in enumerate(FAST_CHECKS) for i, o in enumerate(ORDERS)
         if fn.__name__ not in SLOW or i == k % 2]


@pytest.mark.parametrize("which,fn", CASES, ids=lambda v: v if isinstance(v, str) else v.__name__)
def test_checks_under_other_thread_orders(eng, monkeypatch, which, fn):
    monkeypatch.setenv("IPB_EMU_ORDER", which)
    fn(eng)


@pytest.mark.parametrize("fa_path", [1, 2, 3], ids=["fused-smem", "phases", "fused-global"])
def test_fa_chain_under_other_thread_orders(eng, order, fa_path):
    checks.check_fa_batch(eng, checks.FA_CASES[0], fa_path=fa_path)


@pytest.mark.parametrize("case", checks.N2_CASES[:2], ids=lambda c: "_".join(sorted(c)) or "default")
def test_nesprin2_under_other_thread_orders(eng, order, case):
    checks.check_nesprin2_batch(eng, case)


def test_fret_and_intensity_under_other_thread_orders(eng, order):
    checks.check_fret_batch(eng, *checks.FRET_CASES[1])
    checks.check_intensity_batch(eng, *checks.INTENSITY_CASES[3])


def test_unknown_order_is_refused():
    """A typo in IPB_EMU_ORDER must not silently fall back to the default order."""
    import subprocess
    import sys
    code = ("import os, numpy as np; os.environ['IPB_EMU_ORDER'] = 'sideways';"
            "from imageprocess_b200.ops import Engine; from tests.emu.emu_backend import NumpyMem, emu_lib;"
            "from tests import checks; checks.check_region_moments(Engine(emu_lib(), NumpyMem()))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "not understood" in r.stderr
