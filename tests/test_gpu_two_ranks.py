"""The N-rank product path on ONE GPU: two processes share cuda:0 (gloo rendezvous -- NCCL refuses
two ranks on one device), each runs the CUDA FrameBatchJob on its own frames for several steps, the
step tables reach rank 0 through the shared-memory ring (parallel.ShmTableRing: cudaHostRegister on
the mapping, the step's D2H lands in it) and rank 0 compares what it received with its own run of
the other rank's frames.  Covers, on real hardware, what tests/test_parallel_gloo.py covers on the
emulated build."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        import imageprocess_b200 as ipb
        from imageprocess_b200 import batch
        from oracle.gen_golden import small_scene
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        eng = ipb.engine("cuda:0")
        fa_params = {"alpha": 2.0, "min_area_um": 0.05, "max_area_um": 5.0, "close_radius": 1, "subtract_bg": True}
        task = {"bg_scope": "full", "bg_mode": "percentile", "clip_neg": True, "bg_stride": 4, "percentile": 1.0,
                "per_channel_p": False, "ch_p_map": {}}
        n_steps = 5

        def frames_of(r, step):
            sc = [small_scene(300 + 10 * r + k, H=160, W=224, n_cells=2, blobs=6) for k in range(2)]
            planes = np.stack([np.stack([d, a]) for d, a, _ in sc])
            return np.roll(planes, 5 * step, axis=-1).copy(), [s[2] for s in sc]

        def make_job(with_dist):
            planes, _ = frames_of(0, 0)
            job = batch.FrameBatchJob(eng, planes.shape, stages=("int", "fa"), int_task=task, fa_params=fa_params, fa_px=0.112)
            job.pq_min_px = 0
            job.dist = dist if with_dist else None
            return job

        job = make_job(True)
        p0, polys0 = frames_of(rank, 0)
        dev = eng.mem.from_host(p0)
        job.prime(dev, polys0)                               # graphs captured; the ring's capacity agreed
        assert job._shm is not None and job.gather_via == "shm"
        got, pend = [], []
        for step in range(n_steps):
            planes, polys = frames_of(rank, step)
            dev = eng.mem.from_host(planes)
            pend.append((job.submit(dev, polys), dev))
            if len(pend) > 2:
                job.collect(pend.pop(0)[0])
                got += job.gathered(copy=True)
        while pend:
            job.collect(pend.pop(0)[0])
        got += job.finish(copy=True)
        if rank == 0:
            per_rank = [[e for g in got for e in g["per_rank"][r]] for r in range(world)]
            assert [len(x) for x in per_rank] == [n_steps] * world, [len(x) for x in per_rank]
            ok = True
            for r in range(world):
                jr = make_job(False)
                for step in range(n_steps):
                    planes, polys = frames_of(r, step)
                    want = jr.run(eng.mem.from_host(planes), polys)
                    arena, comps, comp_off = per_rank[r][step]
                    O = jr._plans[next(iter(jr._plans))].O
                    ok &= np.array_equal(comp_off, want.fa_comp_off)
                    ok &= np.array_equal(comps[: int(want.fa_comp_off[-1])], want.fa_comps)
                    so = O.view(arena, "stat_out")[: want.int_stat.size].reshape(want.int_stat.shape)
                    ok &= bool((so["n"] == want.int_stat["n"]).all() and (so["q"] == want.int_stat["q"]).all())
            q.put("ok" if ok else "tables differ")
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                   # surface the failure instead of a queue timeout
        q.put(f"rank {rank}: {type(e).__name__}: {e}")
        raise


def test_two_ranks_one_gpu_shared_ring():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        assert q.get(timeout=300) == "ok"
    finally:
        for p in procs:
            p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
