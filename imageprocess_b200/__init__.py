"""imageprocess_b200 -- B200-native implementation of the per-pixel analysis hot path of
gavyek/ImageProcess (ROI intensity, focal-adhesion segmentation, FRET ratio imaging).

Hand-written sm_100a CUDA kernels behind a C ABI (include/ipb200.h, lib/libipb200.so),
driven from host Python that mirrors the reference's function boundaries.  PyTorch is used
only for device memory, streams and torch.distributed.  There is no CPU fallback.
"""
__version__ = "0.1.0"

_engine = None


def engine(device=None):
    """Process-wide Engine bound to libipb200.so and torch CUDA memory."""
    global _engine
    if _engine is None:
        from . import _lib, device as _device, ops
        lib = _lib.load()
        ops.check_struct_sizes(lib)
        _engine = ops.Engine(lib, _device.TorchMem(device))
    return _engine
