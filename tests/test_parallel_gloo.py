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


def _job_worker(rank, world, port, q):
    try:
        _job_worker_body(rank, world, port, q)
    except Exception as e:                                   # surface the failure instead of a queue timeout
        q.put(f"rank {rank}: {type(e).__name__}: {e}")
        raise


def _job_worker_body(rank, world, port, q):
    """Each rank runs the product's FrameBatchJob (kernels in the emulated build) on its block of
    frames; every step all-gathers the packed tables; rank 0 checks what it received."""
    import torch.distributed as dist
    from imageprocess_b200 import batch
    from imageprocess_b200.ops import Engine
    from oracle.gen_golden import small_scene
    from tests.emu.emu_backend import NumpyMem, emu_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = Engine(emu_lib(), NumpyMem())
    scenes = [small_scene(s, H=96, W=128, n_cells=2, blobs=4) for s in (51, 52, 53, 54)]
    fa_params = {"alpha": 2.0, "min_area_um": 0.05, "max_area_um": 5.0, "close_radius": 1, "subtract_bg": True}
    task = {"bg_scope": "full", "bg_mode": "percentile", "clip_neg": True, "bg_stride": 4, "percentile": 1.0,
            "per_channel_p": False, "ch_p_map": {}}

    def run_block(r, with_dist):
        lo, hi = parallel.shard_range(len(scenes), r, world)
        planes = np.stack([np.stack([d, a]) for d, a, _ in scenes[lo:hi]])
        job = batch.FrameBatchJob(eng, planes.shape, stages=("int", "fa"), int_task=task, fa_params=fa_params, fa_px=0.112)
        job.dist = dist if with_dist else None
        return job, job.run(eng.mem.from_host(planes), [sc[2] for sc in scenes[lo:hi]])

    job, res = run_block(rank, True)
    ok = True
    if rank == 0:
        assert res.gathered is not None and len(res.gathered) == world
        for r in range(world):
            jr, want = run_block(r, False)                    # what rank r must have produced
            O = jr._plans[next(iter(jr._plans))].O
            arena, comps = res.gathered[r]
            ok &= np.array_equal(O.view(arena, "comp_off")[: want.n_rois + 1], want.fa_comp_off)
            # the header names the section, so a receiver with a different layout can read it too
            ok &= np.array_equal(res.gathered_comp_off[r], want.fa_comp_off)
            n = int(want.fa_comp_off[-1])
            ok &= np.array_equal(comps[:n], want.fa_comps)
            so = O.view(arena, "stat_out")[: want.int_stat.size].reshape(want.int_stat.shape)
            ok &= bool((so["n"] == want.int_stat["n"]).all() and (so["q"] == want.int_stat["q"]).all())
        q.put(bool(ok))
    else:
        assert res.gathered is None
    dist.barrier()
    dist.destroy_process_group()


def test_job_all_gather_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_job_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    assert got is True, got
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
