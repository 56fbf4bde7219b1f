#!/usr/bin/env python
"""Concurrent host-to-device bandwidth of 1 / 2 / 4 / 8 GPUs of one box (the
end-to-end number scales 1.00 / 0.99 / 0.54 / 0.43 while the device-timed one scales 0.92: is it the
host?).  Every rank pins a 1 GiB buffer (plain pinned and write-combined) and copies it to its GPU
back to back for ~1.5 s, all ranks at the same time; rank 0 prints per-rank and aggregate GB/s.
Third variant: the rank first moves to the cores of its GPU's NUMA node (parallel.bind_near_gpu),
so the buffer is first-touched there.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 profiles/h2d_matrix.py          (under gpurun --gpus 8)
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = 1 << 30
    dev = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local}")
    out = {}
    numa = None
    for kind in ("pinned", "write_combined", "pinned_near_gpu"):
        if kind == "pinned_near_gpu":
            # the process moves to the cores of its GPU's NUMA node, then allocates and first-touches
            import sys
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
            from imageprocess_b200.parallel import bind_near_gpu
            numa = bind_near_gpu(local)
        if kind in ("pinned", "pinned_near_gpu"):
            host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            host.fill_(1)
            copy = lambda: dev.copy_(host, non_blocking=True)
        else:
            try:
                from cuda import cudart
                err, ptr = cudart.cudaHostAlloc(n, cudart.cudaHostAllocWriteCombined)
                if int(err) != 0:
                    raise RuntimeError(str(err))
                stream = torch.cuda.current_stream().cuda_stream
                copy = lambda: cudart.cudaMemcpyAsync(dev.data_ptr(), ptr, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, stream)
            except Exception as e:          # cuda-python missing or refused: report and go on
                out[kind] = {"error": str(e)[:120]}
                continue
        for _ in range(2):
            copy()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 0
        e0.record()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 1.5:
            copy()
            reps += 1
            if reps % 4 == 0:
                torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        gbs = reps * n / (e0.elapsed_time(e1) / 1e3) / 1e9
        t = torch.tensor([gbs], device=f"cuda:{local}")
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v.item()) for v in allv]
        else:
            vals = [gbs]
        out[kind] = {"per_rank_gbs": [round(v, 1) for v in vals], "aggregate_gbs": round(sum(vals), 1)}
    if world > 1:
        infos = [None] * world
        dist.all_gather_object(infos, numa)
    else:
        infos = [numa]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy": n, "h2d": out, "cpu_count": os.cpu_count(), "numa": infos}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
