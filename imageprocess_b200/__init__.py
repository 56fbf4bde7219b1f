"""imageprocess_b200 -- B200-native implementation of the per-pixel analysis hot path of
gavyek/ImageProcess (ROI intensity, focal-adhesion segmentation, FRET ratio imaging).

Hand-written sm_100a CUDA kernels behind a C ABI (include/ipb200.h, lib/libipb200.so),
driven from host Python that mirrors the reference's function boundaries.  PyTorch is used
only for device memory, streams and torch.distributed.  There is no CPU fallback.
"""
__version__ = "0.1.0"

_engine = None          # the engine handed out for device=None (tests install an emulated one here)
_engines = {}           # device string -> Engine


def engine(device=None):
    """Engine bound to libipb200.so and torch CUDA memory: one per device per process
    (device=None: the first one created, else the current CUDA device's)."""
    global _engine
    if device is None and _engine is not None:
        return _engine
    from . import _lib, device as _device, ops
    mem = _device.TorchMem(device)
    key = str(mem.device)
    if key not in _engines:
        lib = _lib.load()
        ops.check_struct_sizes(lib)
        _engines[key] = ops.Engine(lib, mem)
    if _engine is None:
        _engine = _engines[key]
    return _engines[key]
