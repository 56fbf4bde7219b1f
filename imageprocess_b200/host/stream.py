"""Input side of the `.bat`-level entry points: threaded TIFF decode into
a ring of pinned host buffers, asynchronous H2D on a copy stream, ONE FrameBatchJob that lives for
the whole folder (its plans, device workspace and CUDA graphs are reused by every batch), results
collected two batches behind the submit.

It replaces the reference's process pool over (stage, time) keys (INT/Fluor_INT.py:2211-2229,
FRET/fret_ratio_builder.py:945-970): the pool's workers decoded AND computed; here the workers
only decode (PIL releases the GIL while it reads and decompresses), the device computes.

    stream = FrameStream(eng, (C, H, W), make_job, frames_per_batch=32)
    for idx, res in stream.run(items, load, polys_of):      # idx: positions of the batch's items
        ...                                                 # res: BatchResult; frames idx[k] <-> k

Frames are decoded per batch, never all up front (a long time-lapse does not fit host RAM:
ADVICE round 1); batches keep the caller's item order.
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np


class FrameStream:
    def __init__(self, eng, chw, make_job, frames_per_batch=32, decode_threads=8, depth=3, lag=2):
        t_c = time.perf_counter()
        self.eng, self.mem = eng, eng.mem
        self.C, self.H, self.W = (int(v) for v in chw)
        self.F = int(frames_per_batch)
        self.depth = max(2, int(depth))
        shape = (self.F, self.C, self.H, self.W)
        self.shape = shape
        self.job = make_job(shape)
        # ring buffers are allocated when their slot is first used (page-locking 0.5 GB takes ~0.2 s:
        # slots 1 and 2 are pinned while the pool already decodes into slot 0)
        self.pin = [None] * self.depth
        self.dev = [None] * self.depth
        if os.environ.get("IPB_STREAM_EAGER"):               # measurement switch: the whole ring up front
            self.pin = [self.mem.pinned(shape, np.uint16) for _ in range(self.depth)]
            self.dev = [self.mem.empty(shape, np.uint16) for _ in range(self.depth)]
        self.pool = ThreadPoolExecutor(max_workers=max(1, int(decode_threads)))
        self.timing = {"decode_s": 0.0, "wait_decode_s": 0.0, "batches": 0, "frames": 0,
                       "submit_s": 0.0, "collect_s": 0.0, "construct_s": 0.0}     # host seconds by phase
        self.lag = max(0, min(int(lag), 2))  # batches in flight before one is collected (0: device images of a
                                             # result -- single-buffered per job -- are read before the next submit)
        self.errors = {}                     # item position -> exception raised by its load(); its frame is zeros
        self.timing["construct_s"] = time.perf_counter() - t_c      # job + pinned / device rings

    def _decode_batch(self, slot, chunk, load):
        """Starts the decode of `chunk` into pinned slot `slot`; returns the futures.  A loader that
        takes a second argument gets the frame's slice of the pinned buffer and may fill it in place
        (returning None): plain TIFFs go from the file into page-locked memory without a copy."""
        import inspect
        try:
            self._load_into = len(inspect.signature(load).parameters) >= 2
        except (TypeError, ValueError):
            self._load_into = False
        if self.pin[slot] is None:
            t0 = time.perf_counter()
            self.pin[slot] = self.mem.pinned(self.shape, np.uint16)
            self.dev[slot] = self.mem.empty(self.shape, np.uint16)
            self.timing["construct_s"] += time.perf_counter() - t0
        dst = self.pin[slot][0]

        def one(k, pos, item):
            t0 = time.perf_counter()
            try:                                                 # one unreadable image must not lose the batch
                a = load(item, dst[k]) if self._load_into else load(item)      # (the reference logs the key and goes on)
                if a is not None:                                # None: the loader filled dst[k] itself
                    a = np.asarray(a)
                    if a.shape != (self.C, self.H, self.W):
                        raise ValueError(f"frame of shape {a.shape}, expected {(self.C, self.H, self.W)}")
                    dst[k] = a
            except Exception as e:
                self.errors[pos] = e
                dst[k] = 0
            return time.perf_counter() - t0
        return [self.pool.submit(one, k, pos, it) for k, (pos, it) in enumerate(chunk)]

    def run(self, items, load, polys_of):
        """Generator over batches: yields (list of item positions, BatchResult).  `load(item)` returns
        the frame's uint16 planes [C][H][W]; `polys_of(item)` its ROI polygon list (or None).

        Ring of `depth` (pinned, device) buffer pairs; batch b uses pair b % depth:
          decode(b + 2) runs on the pool while batch b is uploaded and computed;
          upload(b) goes to the copy stream once the step that last read the device buffer is done;
          the step waits for its upload; tickets are collected two batches behind."""
        F, mem, job, D = self.F, self.mem, self.job, self.depth
        chunks = [list(range(b0, min(b0 + F, len(items)))) for b0 in range(0, len(items), F)]
        decoding, uploaded, computed = {}, {}, [None] * D
        inflight = []                                            # (positions, ticket)

        def start_decode(b):
            if b < len(chunks):
                if b - D in uploaded:                            # the pinned buffer's previous upload has run
                    uploaded.pop(b - D).synchronize()
                decoding[b] = self._decode_batch(b % D, [(i, items[i]) for i in chunks[b]], load)

        start_decode(0)
        start_decode(1)
        for b, pos in enumerate(chunks):
            slot = b % D
            t0 = time.perf_counter()
            self.timing["decode_s"] += sum(f.result() for f in decoding.pop(b))
            self.timing["wait_decode_s"] += time.perf_counter() - t0
            n = len(pos)
            if n < F:                                            # tail: repeat the last frame, no ROIs on the padding
                self.pin[slot][0][n:] = self.pin[slot][0][n - 1]
            t1 = time.perf_counter()
            polys = [(polys_of(items[i]) if i not in self.errors else []) for i in pos] + [[] for _ in range(F - n)]
            up = mem.upload_on_copy_stream(self.dev[slot], self.pin[slot][1], after=computed[slot])
            uploaded[b] = up
            mem.wait_event(up)
            tk = job.submit(self.dev[slot], polys)
            computed[slot] = mem.event()
            computed[slot].record()
            inflight.append((pos, tk))
            self.timing["submit_s"] += time.perf_counter() - t1
            if b == 0:
                self.timing["first_submit_s"] = time.perf_counter() - t1      # eager step: workspace allocations, module load
            self.timing["batches"] += 1
            self.timing["frames"] += n
            start_decode(b + 2)
            while len(inflight) > self.lag:
                yield self._collect(inflight.pop(0))
        while inflight:
            yield self._collect(inflight.pop(0))

    def _collect(self, entry):
        t0 = time.perf_counter()
        pos, tk = entry
        res = self.job.collect(tk)
        self.timing["collect_s"] += time.perf_counter() - t0
        return pos, res

    def close(self):
        self.pool.shutdown(wait=True)
