"""Device operators: thin, typed Python fronts of the C ABI (include/ipb200.h).

``Engine(lib, mem)`` is constructed by ``imageprocess_b200.engine()`` with the CUDA library
and torch-backed memory.  (The CPU test tier injects the emulated build of the same kernel
sources and a numpy memory backend; see tests/emu/.)
"""
import numpy as np

from . import geometry as geo


class RoiMasks:
    """Device-resident result of Engine.rasterize()."""

    def __init__(self, table, d, pool, area, union, union_wpr, frame_hw, n_frames, mem):
        self.table, self.d = table, d
        self.pool, self.area, self.union = pool, area, union
        self.union_wpr, self.frame_hw, self.n_frames, self.mem = union_wpr, frame_hw, n_frames, mem

    def mask_host(self, i):
        """ROI i's mask as a bool array over its storage rect (test/debug helper)."""
        t = self.table
        pool = self.pool.host()
        rows, wpr = int(t.rows[i]), int(t.wpr[i])
        words = pool[t.mask_off[i]: t.mask_off[i] + rows * wpr].reshape(rows, wpr)
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")
        sw = int(t.srect[i, 2] - t.srect[i, 0])
        return bits[:, :sw].astype(bool)

    def union_host(self):
        u = self.union.host()
        H, W = self.frame_hw
        bits = np.unpackbits(u.reshape(self.n_frames, H, self.union_wpr).view(np.uint8), axis=2,
                             bitorder="little")
        return bits[:, :, :W].astype(bool)


class Engine:
    def __init__(self, lib, mem):
        self.lib, self.mem = lib, mem

    # ------------------------------------------------------------------ a1 / a2
    def rasterize(self, rule, specs, frame_hw, n_frames=1, want_union=True):
        """Rasterise all ``specs`` (geometry.RoiSpec) of a batch in one launch."""
        mem = self.mem
        t = geo.RoiTable(specs)
        H, W = int(frame_hw[0]), int(frame_hw[1])
        d = {
            "verts": mem.from_host(t.verts if t.verts.size else np.zeros((1, 2))),
            "vert_off": mem.from_host(t.vert_off),
            "erect": mem.from_host(t.erect if t.n else np.zeros((1, 4), np.int32)),
            "srect": mem.from_host(t.srect if t.n else np.zeros((1, 4), np.int32)),
            "org": mem.from_host(t.org if t.n else np.zeros((1, 2), np.int32)),
            "frame": mem.from_host(t.frame if t.n else np.zeros(1, np.int32)),
            "mask_off": mem.from_host(t.mask_off),
            "wpr": mem.from_host(t.wpr if t.n else np.zeros(1, np.int32)),
        }
        pool = mem.empty(max(t.total_words, 1), np.uint32)
        area = mem.empty(max(t.n, 1), np.uint32)
        union_wpr = (W + 31) // 32
        union = mem.zeros((n_frames, H, union_wpr), np.uint32) if want_union else None
        self.lib.call("ipb_rasterize_rois", int(rule), t.n, d["verts"].ptr, d["vert_off"].ptr,
                      d["erect"].ptr, d["srect"].ptr, d["org"].ptr, d["frame"].ptr,
                      d["mask_off"].ptr, t.max_rows, t.max_wpr, pool.ptr, area.ptr,
                      union.ptr if union is not None else None, union_wpr, H, mem.stream)
        return RoiMasks(t, d, pool, area, union, union_wpr, (H, W), n_frames, mem)
