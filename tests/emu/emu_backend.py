"""tests/emu/emu_backend.py -- TEST INFRASTRUCTURE.  numpy-backed memory + the emulated
library, so the product's host logic (imageprocess_b200/ops.py etc.) can drive the product's
kernels on the CPU in `pytest -m "not gpu"`.  Never imported by the product."""
import numpy as np

from imageprocess_b200._lib import Lib
from imageprocess_b200.device import DevBuf
from tests.emu import build_emu


class NumpyMem:
    def empty(self, shape, dtype):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        n = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        raw = np.full(max(n, 16) + 64, 0xA5, dtype=np.uint8)     # poison: catches missing init
        off = (-raw.ctypes.data) % 64
        return DevBuf(raw[off:off + max(n, 16)], dtype, shape, self)

    def zeros(self, shape, dtype):
        return self.empty(shape, dtype).zero_()

    def zero(self, buf):
        buf.raw[:] = 0

    def from_host(self, arr, pinned=None):
        arr = np.ascontiguousarray(arr)
        buf = self.empty(arr.shape, arr.dtype)
        buf.raw[: arr.nbytes] = arr.view(np.uint8).reshape(-1)
        return buf

    def to_host(self, buf):
        return buf.raw[: buf.nbytes].view(buf.dtype).reshape(buf.shape).copy()

    def raw_ptr(self, raw):
        return int(raw.ctypes.data)

    stream = 0

    def n_sms(self):
        return 2                     # persistent grids of a few CTAs: job loops get exercised

    def nvtx_mark(self, name=None):
        self.last_nvtx = name

    def pinned(self, shape, dtype):
        n = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        t = np.zeros(n, dtype=np.uint8)
        return t.view(dtype).reshape(shape), t

    def upload_async(self, buf, pinned_tensor, nbytes=None):
        n = pinned_tensor.size if nbytes is None else int(nbytes)
        buf.raw[:n] = pinned_tensor[:n]

    def upload_on_copy_stream(self, buf, pinned_tensor, after=None, nbytes=None):
        self.upload_async(buf, pinned_tensor, nbytes)
        return NumpyMem._Event()

    def register_host(self, buffer, array):
        return array

    def download_async(self, pinned_tensor, buf, nbytes, offset=0):
        o, n = int(offset), int(nbytes)
        pinned_tensor[o: o + n] = buf.raw[o: o + n]

    def zero_bytes(self, buf, nbytes, offset=0):
        buf.raw[int(offset): int(offset) + int(nbytes)] = 0

    def copy_bytes(self, dst, dst_off, src, src_off, nbytes):
        dst.raw[int(dst_off): int(dst_off) + int(nbytes)] = src.raw[int(src_off): int(src_off) + int(nbytes)]

    class _Event:
        def record(self):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return 0.0

        def query(self):
            return True

    class _SideCtx:
        def __init__(self):
            self.event = NumpyMem._Event()

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def side(self, idx, events):
        return NumpyMem._SideCtx()

    def event(self):
        return NumpyMem._Event()

    def sync(self):
        pass

    def branch(self, idx, detach=False):
        import contextlib
        return contextlib.nullcontext()

    def wait_event(self, ev):
        pass

    def join(self):
        pass

    def all_reduce_max(self, value, dist):
        import torch
        t = torch.tensor([int(value)], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def all_gather_bytes(self, dst, src, nbytes, dist):
        import torch
        n = int(nbytes)
        out = torch.from_numpy(dst.raw[: n * dist.get_world_size()])
        dist.all_gather_into_tensor(out, torch.from_numpy(src.raw[:n].copy()))


_lib = None


def emu_lib():
    global _lib
    if _lib is None:
        _lib = Lib(build_emu.build())
        assert _lib.c.ipb_is_emulated() == 1
    return _lib
