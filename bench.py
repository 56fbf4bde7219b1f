#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: Mpix/s on synthetic 2048x2048 uint16 two-channel
time-lapse frames (config C4 of SURVEY.md 8(d)), per GPU and over N GPUs, with the HBM
roofline of the dominant kernel and the reference CPU path timed beside it.

  python bench.py --gpus 1 --steps K --warmup W              our arm (CUDA, this repo)
  python bench.py --impl reference --steps K --warmup W      the reference's CPU path (oracle
                                                             port of it)
For N > 1 the driver launches it under torchrun (one rank per GPU); frames are sharded by
rank with no data-path collective (weak scaling), only the timing is all-reduced (max).

A "step" is one pass of the complete per-frame hot path (ROI rasterisation, backgrounds,
FRET ratio image, per-ROI intensity + ratio statistics, FA segmentation) over one batch of
`--frames` frames per GPU.  Inputs are larger than L2 (frames*16 MiB), so no L2 flush is
needed between timed iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H = W = 2048
N_CELLS = 24
BLOBS = 60
BYTES_PER_PX = 8.0          # 4 B read (2 x uint16) + 4 B ratio image written

FRET_P = {"bg_scope": "full", "bg_mode": "percentile", "percentile": 1.0, "per_channel_p": False,
          "donor_p": 1.0, "fret_p": 1.0, "clip_neg": True, "eps_percentile": 1.0,
          "ratio_mode": "Donor/FRET"}                       # fret_ratio_builder.py:567-589
INT_TASK = {"bg_scope": "full", "bg_mode": "percentile", "clip_neg": True, "bg_stride": 4,
            "percentile": 1.0, "per_channel_p": False, "ch_p_map": {}}   # Fluor_INT defaults
FA_PARAMS = {"alpha": 2.0, "min_area_um": 1.5, "max_area_um": 30.0, "close_radius": 1,
             "subtract_bg": True}                           # FA_Analyzer.py:298-304
FA_PX = 0.112
# "spread": the adhesion blobs of a cell are kept apart all over the cell (~770 separate adhesions per
# frame); "centre" is round 1's layout (they overlap into 1-2 large adhesions per cell)
BLOB_LAYOUT = os.environ.get("IPB_BENCH_BLOBS", "spread")
FRAME_INFO = {}


def make_frames(n_frames, seed=1234, n_unique=2):
    """C4 frames: n_unique synthetic base frames (same cell geometry), the rest derived by a
    per-frame integer jitter so that every frame holds different pixel data."""
    from imageprocess_b200 import synth
    base = []
    polys = None
    for u in range(n_unique):
        d, a, polys_u = synth.fret_frame(seed=seed, H=H, W=W, n_cells=N_CELLS, r_min=80, r_max=160,
                                         blobs_per_cell=BLOBS, drift=1.0 + 0.2 * (u / max(1, n_unique - 1) - 0.5),
                                         blob_layout=BLOB_LAYOUT, info=FRAME_INFO)
        polys = polys_u
        base.append(np.stack([d, a]))
    rng = np.random.default_rng(seed + 17)
    out = np.empty((n_frames, 2, H, W), dtype=np.uint16)
    for t in range(n_frames):
        b = base[t % n_unique]
        if t < n_unique:
            out[t] = b
        else:
            j = rng.integers(0, 7, size=(2, H, W), dtype=np.uint16)
            out[t] = np.minimum(b.astype(np.uint32) + j, 65535).astype(np.uint16)
    return out, polys


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons = index, [], set()
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                    "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                self.samples.append((float(f[0]), float(f[1])))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.index is not None:
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.index is not None:
            self._t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1],
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU reference
def _ref_frame(args):
    """One frame through the reference's CPU path (oracle port): FRET + ROI intensity + FA."""
    planes, polys = args
    from oracle import port
    d, a = planes[0], planes[1]
    D, A = d.astype(np.float32), a.astype(np.float32)
    port.fret_process_pair(D, A, polys, FRET_P)
    port.int_process_key({1: D.copy(), 2: A.copy()}, polys, None, INT_TASK)
    port.fa_batch_rows(D, polys, FA_PARAMS, FA_PX, with_contours=False)
    return planes.shape[-1] * planes.shape[-2]


def cpu_reference(frames, polys, workers):
    import multiprocessing as mp
    t0 = time.perf_counter()
    if workers <= 1:
        for f in range(frames.shape[0]):
            _ref_frame((frames[f], polys))
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_ref_frame, [(frames[f], polys) for f in range(frames.shape[0])])
    dt = time.perf_counter() - t0
    return frames.shape[0] * H * W / dt / 1e6, dt


def _ref_item(args):
    """One (possibly scaled-down) frame through the reference's CPU path (oracle port)."""
    planes, polys, fa_px = args
    from oracle import port
    D, A = planes[0].astype(np.float32), planes[1].astype(np.float32)
    port.fret_process_pair(D, A, polys, FRET_P)
    port.int_process_key({1: D.copy(), 2: A.copy()}, polys, None, INT_TASK)
    port.fa_batch_rows(D, polys, FA_PARAMS, fa_px, with_contours=False)
    return planes.shape[-1] * planes.shape[-2]


def run_reference(args):
    """The reference's CPU path (oracle port of it) on WHOLE 2048 x 2048 frames of the same workload
    (same generator, seed, ROIs and blob layout as our arm), with every host thread the reference
    would use: its own pool size min(cpu_count, 8) (Fluor_INT.py:2211-2216).

    A step is a bounded sample: `workers` whole frames, one per worker, all in parallel (~30 s).
    When K such steps would not end within the budget (IPB_REF_BUDGET_S, default 780 s) a step is
    ONE whole frame whose ROIs are dealt to the workers (each repeats the cheap whole-frame parts).
    The frame is never scaled down: config equals our arm's."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    from imageprocess_b200 import synth
    cores = os.cpu_count() or 1
    workers = min(cores, 8)
    budget_s = float(os.environ.get("IPB_REF_BUDGET_S", "780"))
    frames, polys = make_frames(2, seed=1234)                 # the two base frames of our arm's rank 0
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None

    def run(items):
        t0 = time.perf_counter()
        if pool is None:
            for it in items:
                _ref_item(it)
        else:
            pool.map(_ref_item, items, chunksize=1)
        return time.perf_counter() - t0

    if args.warmup > 0:             # imports, page faults, pool start-up: a quarter-size scene is enough for that
        d, a, wp = synth.fret_frame(seed=7, H=H // 4, W=W // 4, n_cells=N_CELLS, r_min=20, r_max=40, blobs_per_cell=4,
                                    blob_area=(8, 50))
        run([(np.stack([d, a]), wp, FA_PX)] * workers)
    by_frame = [(frames[i % 2], polys, FA_PX) for i in range(workers)]
    by_roi = lambda k: [(frames[k % 2], polys[i::workers], FA_PX) for i in range(workers)]
    t_first = run(by_frame)                                   # the first timed step, in the reference's own mode
    whole = t_first * args.steps <= budget_s
    t_all, px = t_first, workers * H * W
    for k in range(1, args.steps):
        if whole:
            t_all += run(by_frame)
            px += workers * H * W
        else:
            t_all += run(by_roi(k))
            px += H * W
    if pool is not None:
        pool.close()
    value = px / t_all / 1e6
    mode = (f"{workers} whole 2048x2048 frames per step, one per worker" if whole else
            f"step 1: {workers} whole 2048x2048 frames, one per worker; later steps: one whole 2048x2048 frame, its "
            f"{len(polys)} ROIs dealt to the {workers} workers")
    line = {"impl": "reference", "metric": "Mpix/s (2048x2048 uint16 2ch FRET+FA+ROI-intensity time-lapse)",
            "value": value, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.frames),
            "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": workers, "kind": "port",
                             "sample": f"{mode}; oracle port (FRET+INT+FA, find_contours skipped), multiprocessing "
                                       f"pool of {workers} (the reference's own pool size), {t_all:.0f} s in all"},
            "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(frames_per_step):
    """The `config` object both arms print (same workload, same keys)."""
    placed = FRAME_INFO.get("blobs_placed", 0) // 2            # make_frames paints two base frames
    return {"workload": f"C4-synth 2048x2048x2ch uint16 time-lapse, {N_CELLS} cell ROIs, {BLOBS} FA blobs drawn per cell "
                        f"({BLOB_LAYOUT} layout: {placed} separate blobs per frame)", "frames_per_step_per_gpu": frames_per_step,
            "l2": "inputs larger than L2 (no flush needed)", "stages": ["fret", "int", "fa"]}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import imageprocess_b200 as ipb
    from imageprocess_b200 import batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    eng = ipb.engine(f"cuda:{local}")
    F = args.frames
    frames, polys = make_frames(F, seed=1234 + 1000 * rank)
    polys_pf = [polys] * F              # a time-lapse stage: every frame carries the stage's ROI list (one rasterisation per step)
    if args.per_frame_rois:             # every frame its own ROI list (vertices shifted by a frame-dependent sub-pixel offset):
        polys_pf = [[np.asarray(P, dtype=float) + np.array([0.25 * (f % 4), 0.125 * (f // 4 % 8)]) for P in polys]
                    for f in range(F)]  # F x 24 polygons rasterised per step instead of 24
    shape = (F, 2, H, W)
    pinned_np, pinned_t = eng.mem.pinned(shape, np.uint16)
    pinned_np[...] = frames
    planes = eng.mem.empty(shape, np.uint16)
    eng.mem.upload_async(planes, pinned_t)
    eng.mem.sync()
    job = batch.FrameBatchJob(eng, shape, stages=("fret", "int", "fa"), fret_p=FRET_P, int_task=INT_TASK,
                              fa_params=FA_PARAMS, fa_px=FA_PX)

    tab_dev = torch.device(f"cuda:{local}")

    job.dist = dist                     # N > 1: the packed row tables go to rank 0, ONE NCCL all-gather per 8 steps
    LAG = max(1, args.lag)              # tickets in flight: step k is collected after step k + LAG was submitted

    seen = {"adhesions": 0, "roi_fallbacks": 0, "steps": 0}

    def consume(res):
        """Host side of one step: per-adhesion table with the reference's dtypes; on rank 0 of an
        N-rank job also the other ranks' gathered tables (their adhesion counts are read here)."""
        t = batch.fa_table(res, job.fa_cfg)
        seen["adhesions"] += int(t["label"].shape[0])
        seen["roi_fallbacks"] += int(getattr(res, "roi_fallbacks", 0))
        seen["steps"] += 1
        return res.d2h_bytes

    def consume_gathered(groups):
        """Rank 0: the other ranks' tables as they arrive (their adhesion counts are read here)."""
        for g in groups:
            for ents in g["per_rank"]:
                for arena, comps, comp_off in ents:          # every rank has its own ROI set / arena layout
                    seen["gathered_adhesions"] = seen.get("gathered_adhesions", 0) + int(comp_off[-1])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(loop, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d2h = loop(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=f"cuda:{local}")
            every = torch.zeros(world, device=f"cuda:{local}")
            dist.all_gather_into_tensor(every, t)
            timed.per_rank = [round(float(v) / steps, 4) for v in every.tolist()]      # ms per step of every rank
            ms = float(every.max().item())                                              # the job's time: the slowest rank
        barrier()
        return ms, d2h

    host = {"submit_s": 0.0, "collect_s": 0.0, "gathered_s": 0.0, "n": 0}

    def collect_one(pend):
        t0 = time.perf_counter()
        tk = pend.pop(0)
        tk.event.synchronize()
        host.setdefault("waits", []).append(round(1e3 * (time.perf_counter() - t0), 3))
        d = consume(job.collect(tk))
        t1 = time.perf_counter()
        consume_gathered(job.gathered())
        host["collect_s"] += t1 - t0
        host["gathered_s"] += time.perf_counter() - t1
        return d

    def finish_gathers():
        t0 = time.perf_counter()
        consume_gathered(job.finish())                      # the partial last group; waits for every gather
        host["gathered_s"] += time.perf_counter() - t0

    def loop_resident(steps):
        """Steps over HBM-resident frames; up to LAG steps are in flight before a step's tables are
        unpacked, as consecutive batches of a time-lapse are in the product."""
        pend, d2h = [], 0
        for _ in range(steps):
            t0 = time.perf_counter()
            pend.append(job.submit(planes, polys_pf))
            host["submit_s"] += time.perf_counter() - t0
            host["n"] += 1
            if len(pend) > LAG:
                d2h = collect_one(pend)
        while pend:
            d2h = collect_one(pend)
        finish_gathers()
        return d2h

    # end-to-end: host frames in pinned memory, H2D of step k+1 on a copy stream while step k computes
    planes_b = [planes, eng.mem.empty(shape, np.uint16)]
    copy_stream = torch.cuda.Stream(device=tab_dev)

    def loop_e2e(steps):
        main = torch.cuda.current_stream(tab_dev)
        up = [None, None]          # upload-finished events per buffer
        done = [None, None]        # compute-finished events per buffer
        def upload(k):
            b = k & 1
            with torch.cuda.stream(copy_stream):
                if done[b] is not None:
                    copy_stream.wait_event(done[b])
                eng.mem.upload_async(planes_b[b], pinned_t)
                up[b] = torch.cuda.Event()
                up[b].record(copy_stream)
        upload(0)
        pend, d2h = [], 0
        for k in range(steps):
            if k + 1 < steps:
                upload(k + 1)
            b = k & 1
            main.wait_event(up[b])
            pend.append(job.submit(planes_b[b], polys_pf))
            done[b] = torch.cuda.Event()
            done[b].record(main)
            if len(pend) > LAG:
                d2h = collect_one(pend)
        while pend:
            d2h = collect_one(pend)
        finish_gathers()
        return d2h

    eng.mem.copy_bytes(planes_b[1], 0, planes, 0, planes.nbytes)
    for b in planes_b:                  # setup: buffers, plan tables and the CUDA graphs of both input buffers
        job.prime(b, polys_pf)
    loop_resident(args.warmup)
    torch.cuda.synchronize()
    launches0 = eng.launches
    with ClockSampler(local if rank == 0 else None) as clk:      # one sampler per job, on rank 0's GPU
        for k in host:
            host[k] = [] if k == "waits" else 0
        ms, d2h = timed(loop_resident, args.steps)
        ms_ranks = getattr(timed, "per_rank", None)
        host_submit_ms = 1e3 * host["submit_s"] / max(1, host["n"])
        host_ms = {k[:-2]: round(1e3 * host[k] / max(1, host["n"]), 4) for k in ("submit_s", "collect_s", "gathered_s")}
        if os.environ.get("IPB_BENCH_TRACE"):
            host_ms["waits"] = list(host.get("waits", []))
        launches = eng.launches - launches0
        # per-kernel CUDA-event times: the same steps once more with the branches of a step
        # serialised on one stream (overlapped kernels cannot be timed one by one)
        job.overlap = False
        loop_resident(1)
        eng.profile_start()
        ms_ser, _ = timed(loop_resident, args.steps)
        prof = eng.profile_stop()
        job.overlap = True
    loop_e2e(min(args.warmup, 2))
    ms_e2e, _ = timed(loop_e2e, args.steps)

    mpix_total = world * F * args.steps * H * W / 1e6
    value = mpix_total / (ms / 1e3)
    e2e_value = mpix_total / (ms_e2e / 1e3)
    line = {"metric": "Mpix/s (2048x2048 uint16 2ch FRET+FA+ROI-intensity time-lapse)",
            "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16/f32", "data": "synthetic",
            "config": dict(workload_config(F), **({"rois": "one ROI list per frame (rasterised every step)"} if args.per_frame_rois else {})),
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": int(frames.nbytes),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clk.summary()}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
        dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else None
        kern = {k: {"calls": v[0], "ms": round(v[1], 4), "share": round(v[1] / max(ms_ser, 1e-9), 4)}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        if dom is not None:
            # the entry point with the largest share of the step; bytes and time are both taken
            # per step (an entry point may be called more than once per step)
            name, (ncalls, tot_ms) = dom
            per_step_ms = tot_ms / max(1, args.steps)
            alg_bytes = job.algorithmic_bytes(name)
            achieved = alg_bytes / (per_step_ms / 1e3) / 1e9
            traffic = None
            try:                   # dram bytes per step from the committed ncu --set full capture
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(name)
            except Exception:
                pass
            line["roofline"] = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak,
                                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                                "launch_ms": per_step_ms, "calls_per_step": ncalls / max(1, args.steps),
                                # the same work in the domain's unit: ROI pixels measured per second (three value
                                # sources per pixel: donor, acceptor, ratio)
                                "roi_px_per_s": (job.n_roi_px or 0) / (per_step_ms / 1e3)}
        line["pipeline_roofline"] = {"bytes_per_px": BYTES_PER_PX,
                                     "achieved_gbs": value / world * 1e6 * BYTES_PER_PX / 1e9,
                                     "frac_of_peak": value / world * 1e6 * BYTES_PER_PX / 1e9 / peak}
        line["kernels"] = kern
        line["ms_per_step_serialized"] = ms_ser / args.steps
        line["host_submit_ms_per_step"] = host_submit_ms
        if ms_ranks is not None:
            line["gather_via"] = job.gather_via       # how the other ranks' tables reached rank 0 (shared-memory ring | NCCL all-gather)
            line["ms_per_step_ranks"] = ms_ranks      # device-timed, per rank; the job's ms_per_step is their max
        line["host_ms_per_step"] = host_ms          # rank 0: submit / collect (waits for the step) / gathered tables of all ranks
        line["window_misses"] = int(job.window_misses)       # steps repeated with full histograms (exact either way)
        line["adhesions_per_frame"] = seen["adhesions"] / max(1, seen["steps"] * F)
        line["roi_fallbacks_per_step"] = seen["roi_fallbacks"] / max(1, seen["steps"])   # ROIs repeated by the full-histogram kernels
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            t0 = time.perf_counter()
            v, dt = cpu_reference(frames[:1], polys, 1)
            line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                    "sample": f"1 frame of the same workload through the oracle port "
                                              f"(FRET+INT+FA, find_contours skipped), {dt:.1f} s"}
        print(json.dumps(line))
    # leave without waiting on NCCL / CUDA teardown (a hung shutdown would stall the whole launch):
    # the line is out, every rank has passed the last barrier
    sys.stdout.flush()
    if dist is not None:
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout=10)
        os._exit(0)


# ----------------------------------------------------------------------------- other BASELINE configs
N2_PARAMS = {"px_um": 0.223, "rim_um": 1.12, "annulus_on": False, "ann_in_um": 1.2, "ann_out_um": 2.5,
             "bg_scope": "full", "bg_mode": "percentile", "percentile": 1.0, "per_channel_p": False,
             "donor_p": 1.0, "fret_p": 1.0, "clip_neg": True, "eps_percentile": 1.0, "ratio_mode": "FRET/Donor",
             "use_spectral": True, "alpha": 0.12, "beta": 0.05, "g_factor": 1.1,
             "sat_filter_on": True, "sat_threshold": 65535.0, "clip_ratio_on": True, "clip_ratio_max": 20.0}
C5_FA = {"alpha": 1.0, "min_area_um": 1.5, "max_area_um": 30.0, "close_radius": 1, "subtract_bg": True}


def _load_c1_fixture(exp):
    """One shipped 1Intensity fixture (tests/golden/intensity/<exp>: the reference's own Testsamples frame,
    1536 x 2048, channels 2 and 3 as uint16 + its ROI JSON), read without the test helpers."""
    import lzma
    d = os.path.join(ROOT, "tests", "golden", "intensity", exp)
    rois = json.load(open(os.path.join(d, "rois.json")))
    h, w = rois["image_shape"]["height"], rois["image_shape"]["width"]
    planes = []
    for ch in (2, 3):
        raw = lzma.decompress(open(os.path.join(d, f"ch{ch}.u16.xz"), "rb").read())
        hi = np.frombuffer(raw[:h * w], dtype=np.uint8).astype(np.uint16)
        lo = np.frombuffer(raw[h * w:], dtype=np.uint8).astype(np.uint16)
        planes.append(((hi << 8) | lo).reshape(h, w))
    return np.stack(planes), [np.asarray(P, dtype=float) for P in rois["rois"] if len(P) >= 3]


def other_workload(args, eng):
    """Builds the step of `--config c1 | c2 | c5 | c3s` on `eng` (kept apart from the timing so that the CPU
    suite can drive the same code through the emulated engine, tests/test_emu_host.py)."""
    from imageprocess_b200 import batch, nesprin2, synth
    mem = eng.mem
    if args.config == "c1":
        # BASELINE config 1: 1Intensity.bat on the shipped BCC P0 / P1 frames, the CSV's own settings
        sets = [_load_c1_fixture(e) for e in ("e1_P0", "e2_P1")]
        frames = np.stack([s[0] for s in sets])
        polys_pf = [s[1] for s in sets]
        shape = frames.shape
        pinned_np, pinned_t = mem.pinned(shape, np.uint16)
        pinned_np[...] = frames
        planes = mem.empty(shape, np.uint16)
        mem.upload_async(planes, pinned_t)
        job = batch.FrameBatchJob(eng, shape, stages=("int",), int_task=INT_TASK, int_channels=[0, 1])
        job.ch_names = [2, 3]
        seen = {}

        def step():
            res = job.run(planes, polys_pf)
            seen["roi_rows"] = int(len(res.frame))
            return res.d2h_bytes

        def step_e2e():
            mem.upload_async(planes, pinned_t)
            return step()
        return {"step": step, "step_e2e": step_e2e, "prime": lambda: job.prime(planes, polys_pf), "px": shape[0] * shape[2] * shape[3], "h2d": int(frames.nbytes),
                "bpp": 4.0,                                        # 2 x uint16 read, tables only
                "workload": (f"C1 shipped 1Flu_Intensity frames (BCC P0 + P1, 1536x2048 uint16, channels 2+3, "
                             f"{len(polys_pf[0])} + {len(polys_pf[1])} ROIs), per-ROI intensity tables, 2 frames per step"),
                "metric": "Mpix/s (1536x2048 uint16 2ch per-ROI fluorescence intensity, shipped fixture)",
                "dtype": "u16/f32", "stages": ["int"], "seen": seen, "l2": "inputs SMALLER than L2 (25 MB)"}
    if args.config == "c2":
        # BASELINE config 2: 2FocalAdhesion.bat; the sample's images are not shipped, its ROI JSONs are
        rois = json.load(open(os.path.join(ROOT, "tests", "golden", "fa_rois.json")))
        names = sorted(rois)
        h, w = rois[names[0]]["image_shape"]["height"], rois[names[0]]["image_shape"]["width"]
        polys_pf = [[np.asarray(P, dtype=float) for P in rois[n]["rois"]] for n in names]
        frames = np.stack([synth.fa_cells_frame(100 + k, h, w, polys, blobs_per_cell=40)
                           for k, polys in enumerate(polys_pf)])[:, None]
        shape = frames.shape
        pinned_np, pinned_t = mem.pinned(shape, np.uint16)
        pinned_np[...] = frames
        planes = mem.empty(shape, np.uint16)
        mem.upload_async(planes, pinned_t)
        job = batch.FrameBatchJob(eng, shape, stages=("fa",), fa_params=FA_PARAMS, fa_px=FA_PX)
        seen = {}

        def step():
            res = job.run(planes, polys_pf)
            seen["adhesions"] = int(res.fa_comp_off[-1])
            seen["cell_crops"] = int(len(res.frame))
            return res.d2h_bytes

        def step_e2e():
            mem.upload_async(planes, pinned_t)
            return step()
        return {"step": step, "step_e2e": step_e2e, "prime": lambda: job.prime(planes, polys_pf), "px": shape[0] * h * w, "h2d": int(frames.nbytes),
                "bpp": 2.0,                                        # FA on one channel, tables only
                "workload": (f"C2-synth {h}x{w} uint16 FA images painted under the shipped FA-sample ROI JSONs "
                             f"({len(names)} images, {sum(len(p) for p in polys_pf)} cell outlines of 62-540 vertices), "
                             f"per-adhesion tables, {len(names)} frames per step"),
                "metric": "Mpix/s (2200x3200 uint16 FA segmentation, shipped ROI outlines)",
                "dtype": "u16", "stages": ["fa"], "seen": seen, "l2": "inputs SMALLER than L2 (56 MB)"}
    if args.config == "c5":
        size = 8192
        img, polys = synth.fa_mosaic(seed=99, H=size, W=size, n_blobs=100000)
        shape = (1, 1, size, size)
        pinned_np, pinned_t = mem.pinned(shape, np.uint16)
        pinned_np[...] = img[None, None]
        planes = mem.empty(shape, np.uint16)
        mem.upload_async(planes, pinned_t)
        job = batch.FrameBatchJob(eng, shape, stages=("fa",), fa_params=C5_FA, fa_px=FA_PX, want_labels=True)
        seen = {}

        def step():
            res = job.run(planes, [polys])
            seen["adhesions"] = int(res.fa_comp_off[-1])
            return res.d2h_bytes

        def step_e2e():
            mem.upload_async(planes, pinned_t)
            return step()
        return {"step": step, "step_e2e": step_e2e, "prime": lambda: job.prime(planes, [polys]), "px": size * size, "h2d": int(img.nbytes),
                "bpp": 6.0,                                        # 2 B read + 4 B label map written
                "workload": "C5-synth 8192x8192 uint16 FA mosaic, one ROI over the field, label map returned",
                "metric": "Mpix/s (8192x8192 uint16 stitched FA mosaic: CCL + regionprops)",
                "dtype": "u16", "stages": ["fa"], "seen": seen, "l2": "inputs larger than L2 (no flush needed)"}
    F = max(1, min(args.frames, 8))
    d, a, polys = synth.fret_frame(seed=1234, H=H, W=W, n_cells=N_CELLS, r_min=80, r_max=160)
    ao = (0.3 * a + np.random.default_rng(5).poisson(50, d.shape)).astype(np.uint16)
    frames = np.stack([np.stack([d, a, ao])] * F)
    rng = np.random.default_rng(3)
    for f in range(1, F):                                      # every frame holds different pixel data
        frames[f] = np.minimum(frames[f].astype(np.uint32) + rng.integers(0, 7, frames[f].shape), 65535).astype(np.uint16)
    shape = frames.shape
    pinned_np, pinned_t = mem.pinned(shape, np.uint16)
    pinned_np[...] = frames
    planes = mem.empty(shape, np.uint16)
    mem.upload_async(planes, pinned_t)
    seen = {}

    def step():
        out = nesprin2.nesprin2_batch(eng, planes, shape, [polys] * F, N2_PARAMS, donor_ch=0, acc_ch=1, aonly_ch=2)
        seen["rows"] = sum(len(r) for r in out["rows_per_frame"])
        return 8 * F + 200 * seen["rows"]

    def step_e2e():
        mem.upload_async(planes, pinned_t)
        return step()
    return {"step": step, "step_e2e": step_e2e, "px": F * H * W, "h2d": int(frames.nbytes),
            "bpp": 22.0,                                       # 3 x uint16 read + R, Ralt, Dcorr, Acorr float32 written
            "workload": (f"C3-synth 2048x2048 donor/FRET/acceptor-only uint16 triples, {N_CELLS} ROIs, Nesprin2 builder "
                         f"with use_spectral alpha=0.12 beta=0.05 g=1.1, inner rim, {F} frames per step"),
            "metric": "Mpix/s (2048x2048 uint16 3ch Nesprin2 FRET with spectral correction)",
            "dtype": "u16/f32", "stages": ["nesprin2"], "seen": seen,
            "l2": "inputs larger than L2 (no flush needed)" if frames.nbytes > 126e6 else "inputs SMALLER than L2"}


def run_other(args):
    """`--config c1`: the shipped 1Intensity frames through the per-ROI intensity stage; `--config c2`:
    2200 x 3200 FA images under the shipped FA-sample outlines; `--config c5`: one 8192 x 8192 uint16
    stitched FA mosaic (~1e5 adhesions, label map returned) per step through the one-kernel-per-phase FA
    path; `--config c3s`: the Nesprin2 builder with spectral bleed-through correction on 2048 x 2048
    donor / FRET / acceptor-only triples.  Same JSON line as the headline (C4) run; these are single-GPU
    figures kept under profiles/, not the driver's metric.  Workloads that fit in L2 are timed step by
    step with a 512 MB write between the steps (outside the timed events)."""
    import torch
    import imageprocess_b200 as ipb
    torch.cuda.set_device(0)
    eng = ipb.engine("cuda:0")
    mem = eng.mem
    wl = other_workload(args, eng)
    step, step_e2e, px, h2d, bpp, seen = wl["step"], wl["step_e2e"], wl["px"], wl["h2d"], wl["bpp"], wl["seen"]
    flush = None
    if "SMALLER" in wl["l2"]:
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
        wl["l2"] += ": L2 flushed (512 MB written) before every timed step, outside the step's events"
    if wl.get("prime"):                    # buffers allocated, every output slot's CUDA graph captured before the warm-up
        wl["prime"]()
    mem.sync()

    def timed(fn, steps):
        torch.cuda.synchronize()
        if flush is not None:
            evs, d2h = [], 0
            for _ in range(steps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                d2h = fn()
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs), d2h
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d2h = 0
        for _ in range(steps):
            d2h = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), d2h

    for _ in range(max(args.warmup, 3)):
        step()
    l0 = eng.launches
    with ClockSampler(0) as clk:
        ms, d2h = timed(step, args.steps)
        launches = eng.launches - l0
        eng.profile_start()
        ms_prof, _ = timed(step, args.steps)
        prof = eng.profile_stop()
    ms_e2e, _ = timed(step_e2e, args.steps)
    value = args.steps * px / 1e6 / (ms / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    line = {"metric": wl["metric"], "value": value, "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["dtype"],
            "data": "synthetic" if args.config != "c1" else "shipped fixture", "config": {"workload": wl["workload"], "l2": wl["l2"], "stages": wl["stages"]},
            "e2e": {"value": args.steps * px / 1e6 / (ms_e2e / 1e3), "unit": "Mpix/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clk.summary(),
            "pipeline_roofline": {"bytes_per_px": bpp, "achieved_gbs": value * 1e6 * bpp / 1e9,
                                  "frac_of_peak": value * 1e6 * bpp / 1e9 / peak, "peak": peak},
            "kernels": {k: {"calls": v[0], "ms": round(v[1], 4), "share": round(v[1] / max(ms_prof, 1e-9), 4)}
                        for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
            "seen": seen}
    print(json.dumps(line))


def run_folder(args):
    """`--config folder`: the `.bat`-level entry points on a folder of TIFFs: a
    synthetic C4 time-lapse is written to disk (uncompressed uint16 TIFF pairs + one ROI JSON per
    frame), then fret_ratio_builder.run_headless (3FRET.bat) and Fluor_INT.run_headless
    (1Intensity.bat) process it: threaded decode -> pinned ring -> async upload -> ONE persistent
    FrameBatchJob -> CSV.  Wall-clock Mpix/s, decode and file output included."""
    import shutil
    import tempfile
    import torch
    import imageprocess_b200 as ipb
    from imageprocess_b200.host import Fluor_INT, common, fret_ratio_builder
    import pandas  # noqa: F401  (process start-up, ~2 s: the table writers import it on first use; not part of a folder's time)
    torch.cuda.set_device(0)
    eng = ipb.engine("cuda:0")
    n = int(args.folder_frames)
    root = tempfile.mkdtemp(prefix="ipb_folder_", dir=os.environ.get("IPB_TMP", None))
    try:
        frames, polys = make_frames(min(n, 8), seed=1234, n_unique=2)
        roi_dir = os.path.join(root, "roi")
        os.makedirs(roi_dir)
        t0 = time.perf_counter()
        roi_json = json.dumps({"name": "S01", "image_shape": {"height": H, "width": W},
                               "rois": [np.asarray(P).tolist() for P in polys]})
        for t in range(n):
            f = frames[t % frames.shape[0]]
            common.write_tiff(os.path.join(root, f"S01_t{t:03d}_1.tif"), f[0])
            common.write_tiff(os.path.join(root, f"S01_t{t:03d}_2.tif"), f[1])
            with open(os.path.join(roi_dir, f"S01_t{t:02d}.json"), "w") as fh:      # the mirrors' own name: S01_t07, S01_t123
                fh.write(roi_json)
        t_write = time.perf_counter() - t0
        out = {}
        for name, fn in (("fret_ratio_builder", lambda tm: fret_ratio_builder.run_headless(
                              root, roi_dir, out_root=os.path.join(root, "RES_FRET"), eng=eng, log=lambda s: None,
                              p={"timelapse": True, "out_tif": False, "ratio_mode": "Donor/FRET"},
                              frames_per_batch=args.frames if args.frames < 64 else 32, timing=tm)),
                         ("Fluor_INT", lambda tm: Fluor_INT.run_headless(
                              root, roi_dir, out_root=os.path.join(root, "RES_INT"), eng=eng, log=lambda s: None,
                              cfg={"timelapse": True, "channels_to_quant": [1, 2]}, timing=tm))):
            tm = {}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rows = fn(tm)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out[name] = {"seconds": round(dt, 3), "frames_per_s": round(n / dt, 2), "mpix_per_s": round(n * H * W / dt / 1e6, 1),
                         "rows": len(rows), "decode_thread_seconds": round(tm.get("decode_s", 0.0), 3),
                         "host_waited_for_decode_s": round(tm.get("wait_decode_s", 0.0), 3), "batches": tm.get("batches", 0),
                         "host_seconds": {k[:-2]: round(v, 3) for k, v in tm.items()
                                          if k.endswith("_s") and k not in ("decode_s", "wait_decode_s")}}
        line = {"metric": "Mpix/s (folder of 2048x2048 uint16 2ch TIFF pairs through the headless entry points, wall clock)",
                "value": out["fret_ratio_builder"]["mpix_per_s"], "unit": "Mpix/s", "n_gpus": 1, "steps": 1, "warmup": 0,
                "higher_is_better": True, "data": "synthetic", "dtype": "u16/f32",
                "config": {"workload": f"{n} frames of the C4 scene as uncompressed TIFF pairs on disk + ROI JSON per frame",
                           "tiff_write_seconds": round(t_write, 1), "cpu_count": os.cpu_count()},
                "entry_points": out}
        print(json.dumps(line))
    finally:
        shutil.rmtree(root, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--per-frame-rois", action="store_true",
                    help="give every frame of the step its own ROI list (default: one list per step, a time-lapse stage)")
    ap.add_argument("--lag", type=int, default=2, help="steps in flight before a step's tables are unpacked")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c4", choices=["c4", "c1", "c2", "c5", "c3s", "folder"],
                    help="c4: the BASELINE metric (default); c1: shipped 1Intensity frames; c2: FA under the shipped outlines; c5: 8192^2 FA mosaic; c3s: Nesprin2 with spectral correction; "
                         "folder: the headless entry points on a folder of TIFFs")
    ap.add_argument("--folder-frames", type=int, default=256, help="--config folder: frames written to disk")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "folder":
        run_folder(args)
    elif args.config != "c4":
        run_other(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
