"""Device-memory plumbing.  PyTorch is used only for what the task allows it for: device
allocations, streams and (in pipeline.py) torch.distributed.  Every buffer handed to the
C ABI is a raw byte tensor; dtype/shape live in the small DevBuf wrapper."""
import numpy as np


class DevBuf:
    """A typed view over a flat byte buffer owned by a memory backend."""
    __slots__ = ("raw", "dtype", "shape", "mem", "_keep")

    def __init__(self, raw, dtype, shape, mem):
        self.raw, self.dtype, self.shape, self.mem = raw, np.dtype(dtype), tuple(shape), mem

    @property
    def ptr(self):
        return self.mem.raw_ptr(self.raw)

    @property
    def nbytes(self):
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    def host(self):
        return self.mem.to_host(self)

    def zero_(self):
        self.mem.zero(self)
        return self


class _Branch:
    def __init__(self, mem, idx, detach=False):
        self.mem, self.idx, self.detach, self.event = mem, idx, detach, None

    def __enter__(self):
        mem, torch = self.mem, self.mem.torch
        side = mem.__dict__.setdefault("_side", {})
        if self.idx not in side:
            side[self.idx] = torch.cuda.Stream(device=mem.device)
        s = side[self.idx]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(mem.device))
        s.wait_event(ev)
        self.ctx = torch.cuda.stream(s)
        self.ctx.__enter__()
        self.s = s
        return self

    def __exit__(self, *a):
        ev = self.mem.torch.cuda.Event()
        ev.record(self.s)
        self.event = ev
        if not self.detach:
            self.mem.__dict__.setdefault("_pending", []).append(ev)
        return self.ctx.__exit__(*a)


class _Side:
    def __init__(self, mem, idx, events):
        self.mem, self.idx, self.events, self.event = mem, idx, list(events), None

    def __enter__(self):
        mem, torch = self.mem, self.mem.torch
        side = mem.__dict__.setdefault("_side", {})
        if self.idx not in side:
            side[self.idx] = torch.cuda.Stream(device=mem.device)
        self.s = side[self.idx]
        for ev in self.events:
            self.s.wait_event(ev)
        self.ctx = torch.cuda.stream(self.s)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        ev = self.mem.torch.cuda.Event()
        ev.record(self.s)
        self.event = ev
        return self.ctx.__exit__(*a)


class TorchMem:
    """CUDA memory through torch (caching allocator, current stream)."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("imageprocess_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")

    def n_sms(self):
        return int(self.torch.cuda.get_device_properties(self.device).multi_processor_count)

    def nvtx_mark(self, name=None):
        """Closes the open NVTX range of this backend and opens `name` (None: just close): one range
        per stage of a step, visible in nsys / ncu timelines, free otherwise."""
        nv = self.torch.cuda.nvtx
        if getattr(self, "_nvtx_open", False):
            nv.range_pop()
            self._nvtx_open = False
        if name is not None:
            nv.range_push(name)
            self._nvtx_open = True

    # -- allocation
    def empty(self, shape, dtype):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        raw = self.torch.empty(max(n, 16), dtype=self.torch.uint8, device=self.device)
        return DevBuf(raw, dtype, shape, self)

    def zeros(self, shape, dtype):
        return self.empty(shape, dtype).zero_()

    def zero(self, buf):
        buf.raw.zero_()

    def from_host(self, arr, pinned=None):
        arr = np.ascontiguousarray(arr)
        buf = self.empty(arr.shape, arr.dtype)
        if arr.nbytes:
            src = self.torch.from_numpy(arr.view(np.uint8).reshape(-1))
            buf.raw[: arr.nbytes].copy_(src, non_blocking=False)
        return buf

    def wrap_tensor(self, t, dtype=None, shape=None):
        """Wrap an existing contiguous CUDA tensor (no copy)."""
        assert t.is_cuda and t.is_contiguous()
        raw = t.view(self.torch.uint8).reshape(-1)
        return DevBuf(raw, dtype or np.dtype(str(t.dtype).replace("torch.", "")), shape or tuple(t.shape), self)

    def to_host(self, buf):
        n = buf.nbytes
        if n == 0:
            return np.zeros(buf.shape, dtype=buf.dtype)
        h = buf.raw[:n].cpu().numpy()
        return h.view(buf.dtype).reshape(buf.shape).copy()

    def raw_ptr(self, raw):
        return int(raw.data_ptr())

    @property
    def stream(self):
        return int(self.torch.cuda.current_stream(self.device).cuda_stream)

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def pinned(self, shape, dtype):
        """Pinned host staging buffer as a numpy array (+ the torch tensor that owns it)."""
        n = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        t = self.torch.empty(n, dtype=self.torch.uint8, pin_memory=True)
        return t.numpy().view(dtype).reshape(shape), t

    def register_host(self, buffer, array):
        """Page-locks existing host memory (a shared-memory mapping) and returns the torch tensor
        over it, so that copies into it are asynchronous like those into pinned() buffers."""
        t = self.torch.frombuffer(buffer, dtype=self.torch.uint8)
        rc = self.torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel(), 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed: {rc}")
        return t

    def upload_async(self, buf, pinned_tensor, nbytes=None):
        """H2D copy of a pinned tensor into an existing DevBuf on the current stream."""
        n = pinned_tensor.numel() if nbytes is None else int(nbytes)
        buf.raw[:n].copy_(pinned_tensor[:n], non_blocking=True)

    def upload_on_copy_stream(self, buf, pinned_tensor, after=None, nbytes=None):
        """H2D copy on the backend's copy stream, ordered after event `after` (the last reader of
        `buf`); returns the event that marks its end.  Lets the next batch's upload run beside the
        current batch's kernels."""
        torch = self.torch
        cs = self.__dict__.get("_copy_stream")
        if cs is None:
            cs = self.__dict__["_copy_stream"] = torch.cuda.Stream(device=self.device)
        if after is not None:
            cs.wait_event(after)
        n = pinned_tensor.numel() if nbytes is None else int(nbytes)
        with torch.cuda.stream(cs):
            buf.raw[:n].copy_(pinned_tensor[:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
        return ev

    def download_async(self, pinned_tensor, buf, nbytes, offset=0):
        """D2H copy into a pinned tensor on the current stream (caller syncs); `offset` applies to both sides."""
        o, n = int(offset), int(nbytes)
        pinned_tensor[o: o + n].copy_(buf.raw[o: o + n], non_blocking=True)

    def zero_bytes(self, buf, nbytes, offset=0):
        buf.raw[int(offset): int(offset) + int(nbytes)].zero_()

    def copy_bytes(self, dst, dst_off, src, src_off, nbytes):
        dst.raw[int(dst_off): int(dst_off) + int(nbytes)].copy_(src.raw[int(src_off): int(src_off) + int(nbytes)])

    def sync(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    # -- side streams: independent branches of one step run concurrently (fork / join by events)
    def branch(self, idx, detach=False):
        """Context manager: work enqueued inside goes to side stream `idx`, ordered after
        everything enqueued on the current stream so far.  join() makes the current stream
        wait for every branch opened since the last join(); a detached branch is not joined
        (its .event is the caller's to wait on)."""
        return _Branch(self, idx, detach)

    def wait_event(self, ev):
        self.torch.cuda.current_stream(self.device).wait_event(ev)

    def side(self, idx, events):
        """Context manager: work enqueued inside goes to side stream `idx`, ordered after `events`
        only (NOT after the rest of the current stream); never joined -- its .event is the caller's."""
        return _Side(self, idx, events)

    def graph(self):
        """(CUDA graph, capture context): work enqueued inside the context on the current stream
        (and on branches forked from it) is recorded, not run."""
        g = self.torch.cuda.CUDAGraph()
        return g, self.torch.cuda.graph(g, capture_error_mode="thread_local")

    def join(self):
        cur = self.torch.cuda.current_stream(self.device)
        for ev in getattr(self, "_pending", []):
            cur.wait_event(ev)
        self._pending = []

    def all_reduce_max(self, value, dist):
        """max over ranks of a Python int (one small collective + host read; used once per job)."""
        t = self.torch.tensor([int(value)], dtype=self.torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def all_gather_bytes(self, dst, src, nbytes, dist):
        """dst[r * nbytes : (r + 1) * nbytes] = rank r's src[:nbytes] (NCCL all-gather, enqueued
        asynchronously with respect to the host, ordered after the current stream's work)."""
        n = int(nbytes)
        dist.all_gather_into_tensor(dst.raw[: n * dist.get_world_size()], src.raw[:n])
