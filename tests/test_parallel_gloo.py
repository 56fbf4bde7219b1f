"""N > 1 path on the CPU: two gloo ranks shard a frame list, run the host-side table logic on
their block and gather the row tables to rank 0.  The per-pixel kernels are
not involved here (they are exercised per rank in the emu / GPU tiers)."""
import os
import socket

import numpy as np
import pytest
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


def _job_worker(rank, world, port, q, via):
    try:
        _job_worker_body(rank, world, port, q, via)
    except Exception as e:                                   # surface the failure instead of a queue timeout
        q.put(f"rank {rank}: {type(e).__name__}: {e}")
        raise


def _job_worker_body(rank, world, port, q, via):
    """Each rank runs the product's FrameBatchJob (kernels in the emulated build) on its block of
    frames for several steps; the staged tables go out with ONE all-gather per `gather_every`
    steps (+ the partial last group at finish()); rank 0 checks what it received.  Rank 1's second
    step carries a bright plane: its sampled percentile window misses deterministically and the
    step is repeated with full histograms -- the rerun must stay rank-local (a
    rerun that issued its own collective desynchronised the ranks)."""
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
    n_steps = {0: 3, 1: 2}                                    # ranks own a different number of steps

    def planes_of(r, step):
        lo, hi = parallel.shard_range(len(scenes), r, world)
        planes = np.stack([np.stack([d, a]) for d, a, _ in scenes[lo:hi]])
        planes = np.roll(planes, 3 * step, axis=-1).copy()    # a different batch every step
        if r == 1 and step == 1:
            planes[0, 0] = 40000 + planes[0, 0] % 512          # above the 15-bit sample histogram: deterministic window miss
        return planes, [sc[2] for sc in scenes[lo:hi]]

    def make_job(r, with_dist):
        planes, _ = planes_of(r, 0)
        job = batch.FrameBatchJob(eng, planes.shape, stages=("int", "fa"), int_task=task, fa_params=fa_params, fa_px=0.112)
        job.pq_min_px = 0                                     # sampled windows even on these small frames
        job.gather_every = 2
        job.gather_via = via
        job.dist = dist if with_dist else None
        return job

    job = make_job(rank, True)
    p0, polys0 = planes_of(rank, 0)
    job.n_slots = 1                                           # prime() runs 3 setup steps per slot: keep the emulated test short
    job.prime(eng.mem.from_host(p0), polys0)                  # setup steps take no part in the gathers; the first one
    assert job._pc_hint is not None and job._gather_cap is not None      # sizes the gather capacity (agreed on the second)
    job.begin_distributed(n_steps[rank])
    got = []
    for step in range(n_steps[rank]):
        planes, polys = planes_of(rank, step)
        job.run(eng.mem.from_host(planes), polys)
        got += job.gathered(copy=True)          # kept until the end of the test: private copies
    got += job.finish(copy=True)
    misses = job.window_misses
    ok = True
    if rank == 0:
        assert (job._shm is not None) == (via == "shm")
        assert [g["group"] for g in got] == sorted(g["group"] for g in got), [g["group"] for g in got]
        if via == "nccl":
            assert len(got) == 2, len(got)                        # 3 steps in groups of 2: one full, one partial gather
        per_rank = [[e for g in got for e in g["per_rank"][r]] for r in range(world)]
        assert [len(x) for x in per_rank] == [n_steps[0], n_steps[1]], [len(x) for x in per_rank]
        for r in range(world):
            jr = make_job(r, False)
            for step in range(n_steps[r]):
                planes, polys = planes_of(r, step)
                want = jr.run(eng.mem.from_host(planes), polys)          # what rank r must have produced
                O = jr._plans[next(iter(jr._plans))].O
                arena, comps, comp_off = per_rank[r][step]
                ok &= np.array_equal(O.view(arena, "comp_off")[: want.n_rois + 1], want.fa_comp_off)
                # the header names the section, so a receiver with a different layout can read it too
                ok &= np.array_equal(comp_off, want.fa_comp_off)
                n = int(want.fa_comp_off[-1])
                ok &= np.array_equal(comps[:n], want.fa_comps)
                so = O.view(arena, "stat_out")[: want.int_stat.size].reshape(want.int_stat.shape)
                ok &= bool((so["n"] == want.int_stat["n"]).all() and (so["q"] == want.int_stat["q"]).all())
                bg = O.view(arena, "params")[jr._plans[next(iter(jr._plans))].P_INT:][: want.int_bg.size]
                ok &= np.array_equal(bg.reshape(want.int_bg.shape), want.int_bg)
        q.put((bool(ok), misses))
    else:
        assert got == [] and misses >= 1, misses               # the bright plane did miss, and was repeated locally
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("via", ["shm", "nccl"])
def test_job_all_gather_gloo(via):
    """via = "shm": shared-memory table ring (ranks of one host, the default); "nccl": the
    all-gather path (gloo stands in for NCCL here)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_job_worker, args=(r, world, port, q, via)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    assert got == (True, 0), got
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0


def _ring_worker(rank, world, port, q):
    """Producer (rank 1) publishes 50 entries into a 16-entry ring while the consumer (rank 0, also a
    producer of 5) polls: order, content, back-pressure (the producer can never be more than a ring
    ahead of what the consumer released) and the end-of-run handshake."""
    try:
        import torch.distributed as dist
        from tests.emu.emu_backend import NumpyMem
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ring = parallel.ShmTableRing(dist, NumpyMem(), cap=4096, n_entries=16, dst=0)
        assert ring.ok
        n_mine = 50 if rank == 1 else 5
        seen = {0: [], 1: []}
        prev = [0, 0]

        def consume():
            for r in range(world):
                ring.release(r, prev[r])
            for r in range(world):
                for pos, blob in ring.poll_rank(r):
                    assert ring.next[r] - ring.released[r] <= ring.n
                    assert int(blob[:8].view(np.int64)[0]) == 1000 * r + pos and int(blob[-1]) == pos % 251
                    seen[r].append(pos)
                prev[r] = ring.next[r]

        for pos in range(n_mine):
            arr, _ = ring.entry(pos)                   # blocks while the consumer has not released pos - 16
            arr[:8].view(np.int64)[0] = 1000 * rank + pos
            arr[-1] = pos % 251
            ring.publish(pos)
            if rank == 0:
                consume()
        ring.end_run(n_mine)
        if rank == 0:
            import time
            t0 = time.time()
            while not ring.drained(ring.run):
                consume()
                assert time.time() - t0 < 60
                time.sleep(1e-4)
            assert seen[0] == list(range(5)) and seen[1] == list(range(50)), (seen[0], len(seen[1]))
            q.put("ok")
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        q.put(f"rank {rank}: {type(e).__name__}: {e}")
        raise


def test_shm_table_ring_backpressure():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    assert q.get(timeout=120) == "ok"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("ipb200_tables_")] if os.path.isdir("/dev/shm") else True
