"""N > 1 path on the CPU: two gloo ranks shard a frame list, run the host-side table logic on
their block and gather the row tables to rank 0.  The per-pixel kernels are
not involved here (they are exercised per rank in the emu / GPU tiers)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from imageprocess_b200 import parallel


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            blocks = [parallel.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b[1] == blocks[i + 1][0] for i, b in enumerate(blocks[:-1]))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(n_frames, rank, world)
    rows = []
    for f in range(lo, hi):                       # per-frame "results" with a ragged ROI count
        for roi in range(1, 2 + f % 3):
            rows.append({"frame": f, "roi": roi, "area_px": 100 * f + roi, "ratio_mean": f + roi / 10.0,
                         "time": None})
    merged = parallel.gather_rows(rows, ["frame", "roi", "area_px", "ratio_mean", "time"], dist)
    comps = np.zeros(hi - lo, dtype=[("sum_i", "u8"), ("area", "u4"), ("crop", "i4")])
    comps["sum_i"], comps["area"], comps["crop"] = np.arange(lo, hi) * 7, np.arange(lo, hi), rank
    tabs = parallel.gather_tables(comps, dist)
    if rank == 0:
        q.put((merged, [t.tolist() for t in tabs]))
    else:
        assert merged is None and tabs is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    world, n_frames = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged, tabs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = [(f, roi) for f in range(n_frames) for roi in range(1, 2 + f % 3)]
    assert [(int(r["frame"]), int(r["roi"])) for r in merged] == want          # (frame, roi) order kept
    assert all(r["time"] is None for r in merged)
    assert [int(r["area_px"]) for r in merged] == [100 * f + roi for f, roi in want]
    flat = [row for t in tabs for row in t]
    assert [r[1] for r in flat] == list(range(n_frames)) and [r[0] for r in flat] == [7 * i for i in range(n_frames)]
