"""Device operators: thin, typed Python fronts of the C ABI (include/ipb200.h).

``Engine(lib, mem)`` is constructed by ``imageprocess_b200.engine()`` with the CUDA library
and torch-backed memory.  (The CPU test tier injects the emulated build of the same kernel
sources and a numpy memory backend; see tests/emu/.)
"""
import math

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


# kernels launched by each C-ABI entry point (memsets not counted)
KERNELS_PER_CALL = {"ipb_fa_segment": 6,   # per-crop shared-memory path (14 on the one-kernel-per-phase path)
                     "ipb_rasterize_rois": 1, "ipb_hist_u16": 1, "ipb_hist_quantiles": 1,
                    "ipb_scatter_qvalues": 1, "ipb_fret_eps": 1, "ipb_fa_params": 1,
                    "ipb_fret_pixels": 1, "ipb_region_stats": 1, "ipb_roi_stats_fused": 2, "ipb_region_dilate": 2, "ipb_hist_select": 3, "ipb_hist_planes": 1}


class Engine:
    def __init__(self, lib, mem):
        self.lib, self.mem = lib, mem
        self.launches = 0          # kernels launched through this engine
        self.prof = None           # name -> list of (start_event, end_event) when profiling

    def call(self, name, *args):
        """Checked C-ABI call; counts kernel launches and, when profiling is on, brackets the
        call with CUDA events on the launching stream."""
        self.launches += KERNELS_PER_CALL.get(name, 1)
        if self.prof is None:
            return self.lib.call(name, *args)
        e0, e1 = self.mem.event(), self.mem.event()
        e0.record()
        rc = self.lib.call(name, *args)
        e1.record()
        self.prof.setdefault(name, []).append((e0, e1))
        return rc

    def n_sms(self):
        """Streaming multiprocessors of the device (sizes the persistent grids)."""
        f = getattr(self.mem, "n_sms", None)
        return int(f()) if f else 148

    def profile_start(self):
        self.prof = {}

    def profile_stop(self):
        """Returns {entry point: (n_calls, total_ms)}; caller must have synchronised."""
        out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (self.prof or {}).items()}
        self.prof = None
        return out

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
        self.call("ipb_rasterize_rois", int(rule), t.n, d["verts"].ptr, d["vert_off"].ptr,
                      d["erect"].ptr, d["srect"].ptr, d["org"].ptr, d["frame"].ptr,
                      d["mask_off"].ptr, t.max_rows, t.max_wpr, pool.ptr, area.ptr,
                      union.ptr if union is not None else None, union_wpr, H, mem.stream)
        return RoiMasks(t, d, pool, area, union, union_wpr, (H, W), n_frames, mem)


# ---------------------------------------------------------------------- struct mirrors
PAT_FULL, PAT_STRIDE1D, PAT_STRIDE2D, PAT_MASKED, PAT_MASKED_STRIDE = 0, 1, 2, 3, 4
SRC_U16, SRC_F32, SRC_RATIO = 0, 1, 2
QK_NONE, QK_PCT, QK_MEDIAN = 0, 1, 2
FP_BD, FP_BA, FP_EPS, FP_BAO, FP_STRIDE = 0, 1, 2, 3, 4

HIST_JOB = np.dtype([("plane", "i4"), ("pattern", "i4"), ("k", "i4"), ("mask_frame", "i4"),
                     ("moments", "i4"), ("excl_plane1", "i4"), ("sat_min", "i4"), ("pad", "i4")])
Q_JOB = np.dtype([("hist", "i4"), ("q32", "f4"), ("pad", "i4", 2)])
Q_OUT = np.dtype([("prev", "i4"), ("next", "i4"), ("gamma", "f4"), ("value", "f4"), ("n", "u8")])
REGION = np.dtype([("mask_off", "i8"), ("x0", "i4"), ("y0", "i4"), ("w", "i4"), ("h", "i4"),
                   ("wpr", "i4"), ("frame", "i4"), ("use_and", "i4"), ("and_plane", "i4")])
STAT_JOB = np.dtype([("region", "i4"), ("src", "i4"), ("plane", "i4"), ("n_views", "i4"),
                     ("bidx", "i4", 2), ("clip_neg", "i4", 2), ("qkind", "i4", 3), ("q32", "f4", 3),
                     ("out", "i4", 2)])
STAT_OUT = np.dtype([("n", "u8"), ("area", "u8"), ("sum", "f8"), ("ssd", "f8"), ("vmin", "f4"),
                     ("vmax", "f4"), ("q", "f4", 3), ("pad0", "f4")])
FRET_CFG = np.dtype([("numer_is_acceptor", "i4"), ("clip_neg", "i4"), ("sat_on", "i4"),
                     ("sat_thr", "f4"), ("use_spectral", "i4"), ("alpha", "f4"), ("beta", "f4"),
                     ("g_factor", "f4"), ("clip_on", "i4"), ("clip_max", "f4"), ("donor_ch", "i4"),
                     ("acc_ch", "i4"), ("aonly_ch", "i4"), ("n_ch", "i4")])
CROP = np.dtype([("bit_off", "i8"), ("pix_off", "i8"), ("row_off", "i8"), ("mask_off", "i8"),
                 ("ox", "i4"), ("oy", "i4"),
                 ("w", "i4"), ("h", "i4"), ("wpr", "i4"), ("plane", "i4"), ("frame", "i4"), ("pad0", "i4")])
COMP = np.dtype([("sum_i", "u8"), ("sum_y", "u8"), ("sum_x", "u8"), ("area", "u4"), ("crop", "i4")])
CROP_JOB = np.dtype([("plane", "i4"), ("x0", "i4"), ("y0", "i4"), ("w", "i4"), ("h", "i4"), ("region", "i4"),
                     ("out_off", "i8")])
PLANE_PASS = np.dtype([("plane", "i4"), ("excl_plane1", "i4"), ("sat_min", "i4"), ("n_jobs", "i4"), ("job", "i4", 4)])
HIST_WIN = np.dtype([("wlo", "i4"), ("whi", "i4"), ("mode", "i4"), ("pad", "i4")])
ROI_JOB = np.dtype([("region", "i4"), ("plane", "i4", 2), ("n_views", "i4", 2), ("bidx", "i4", (2, 2)),
                    ("clip", "i4", (2, 2)), ("out", "i4", (2, 2)), ("qkind", "i4", (2, 3)), ("q32", "f4", (2, 3)),
                    ("ratio_on", "i4"), ("ratio_out", "i4"), ("fp_idx", "i4"), ("numer_slot", "i4"),
                    ("ratio_clip_neg", "i4"), ("rqkind", "i4", 3), ("rq32", "f4", 3)])
_SIZEOF = [HIST_JOB, Q_JOB, Q_OUT, REGION, STAT_JOB, STAT_OUT, FRET_CFG, CROP, COMP, CROP_JOB, PLANE_PASS, HIST_WIN,
           ROI_JOB]
RF_CTAS_PER_SM = 2


def plane_passes(hist_jobs):
    """Groups histogram jobs that read the same plane (and saturation partner) into passes of
    up to four jobs for ipb_hist_select's single read of each plane."""
    groups = {}
    for j, hj in enumerate(hist_jobs):
        groups.setdefault((int(hj["plane"]), int(hj["excl_plane1"]), int(hj["sat_min"])), []).append(j)
    out = []
    for (plane, excl, sat), js in groups.items():
        for i in range(0, len(js), 4):
            pp = np.zeros(1, dtype=PLANE_PASS)[0]
            pp["plane"], pp["excl_plane1"], pp["sat_min"], pp["n_jobs"] = plane, excl, sat, len(js[i: i + 4])
            pp["job"][: len(js[i: i + 4])] = js[i: i + 4]
            out.append(pp)
    return np.array(out, dtype=PLANE_PASS) if out else np.zeros(0, dtype=PLANE_PASS)


PQ_WIN = 2048


def pq_servable(hist_jobs, passes):
    """True when every plane pass can take ipb_hist_select's sampled-window path (mirror of
    ipb_pq_roles in csrc/ipb_pq.cuh): at most one FULL job, one flat-stride job with k in
    {2, 4, 8} and one [::k, ::k] job per pass, no masks, no saturation filter."""
    for pp in passes:
        if int(pp["sat_min"]) > 0:
            return False
        roles = set()
        for j in pp["job"][: int(pp["n_jobs"])]:
            hj = hist_jobs[int(j)]
            if int(hj["sat_min"]) > 0 or int(hj["excl_plane1"]) > 0:
                return False
            pat, k = int(hj["pattern"]), int(hj["k"])
            if pat == PAT_FULL and "F" not in roles:
                roles.add("F")
            elif pat == PAT_STRIDE1D and k in (2, 4, 8) and "S" not in roles:
                roles.add("S")
            elif pat == PAT_STRIDE2D and k >= 2 and "P" not in roles:
                roles.add("P")
            else:
                return False
    return True


def q32_of(p):
    """numpy.percentile's quantile for a float32 sample: true_divide(p, float32(100))."""
    return np.float32(np.true_divide(p, np.float32(100)))


def check_struct_sizes(lib):
    for i, dt in enumerate(_SIZEOF):
        n = lib.c.ipb_sizeof(i)
        if n != dt.itemsize:
            raise RuntimeError(f"struct {i}: library says {n} bytes, python mirror {dt.itemsize}")


class HistResult:
    def __init__(self, hist, stats, n_jobs):
        self.hist, self.stats, self.n_jobs = hist, stats, n_jobs


def check_hist_jobs(jobs):
    """include/ipb200.h: the stride of a strided pattern is >= 2 (a stride of 1 is the FULL / MASKED pattern:
    the kernels' multiply-high remainder has no 32-bit constant for k = 1)."""
    strided = np.isin(jobs["pattern"], (PAT_STRIDE1D, PAT_STRIDE2D, PAT_MASKED_STRIDE))
    if bool((strided & (jobs["k"] < 2)).any()):
        raise ValueError("histogram job with a strided pattern and k < 2: use the FULL / MASKED pattern for a stride of 1")
    return jobs


def _engine_hist(self, planes, H, W, jobs, union=None, union_wpr=0):
    """jobs: structured array HIST_JOB.  Returns HistResult (device)."""
    mem = self.mem
    jobs = check_hist_jobs(np.ascontiguousarray(jobs, dtype=HIST_JOB))
    n = jobs.shape[0]
    d_jobs = mem.from_host(jobs if n else np.zeros(1, HIST_JOB))
    hist = mem.empty((max(n, 1), 65536), np.uint32)
    stats = mem.empty((max(n, 1), 4), np.uint64)
    has_ms = bool((jobs["pattern"] == PAT_MASKED_STRIDE).any()) if n else False
    scratch = mem.empty((max(n, 1), H), np.uint64) if has_ms else None
    self.call("ipb_hist_u16", planes.ptr, H, W, d_jobs.ptr, n, int(has_ms),
                  union.ptr if union is not None else None, int(union_wpr),
                  scratch.ptr if scratch is not None else None, hist.ptr, stats.ptr, mem.stream)
    res = HistResult(hist, stats, n)
    res._keep = (d_jobs, scratch)
    return res


def _engine_quantiles(self, hres, qjobs):
    """qjobs: structured array Q_JOB.  Returns device buffer of Q_OUT."""
    mem = self.mem
    qjobs = np.ascontiguousarray(qjobs, dtype=Q_JOB)
    n = qjobs.shape[0]
    d_q = mem.from_host(qjobs if n else np.zeros(1, Q_JOB))
    qout = mem.empty(max(n, 1), Q_OUT)
    self.call("ipb_hist_quantiles", hres.hist.ptr, hres.stats.ptr, d_q.ptr, n, qout.ptr, mem.stream)
    qout._keep = d_q
    return qout


def _engine_scatter_qvalues(self, qout, dst_idx, dst):
    idx = self.mem.from_host(np.ascontiguousarray(dst_idx, dtype=np.int32))
    self.call("ipb_scatter_qvalues", qout.ptr, idx.ptr, int(len(dst_idx)), dst.ptr, self.mem.stream)


def _engine_fret_eps(self, qout_eps, n_frames, denom_slot, clip_neg, fparams, eps_abs=5.0):
    self.call("ipb_fret_eps", qout_eps.ptr, int(n_frames), int(denom_slot), int(bool(clip_neg)),
                  float(eps_abs), fparams.ptr, self.mem.stream)


def _engine_fa_params(self, hres, stat_idx, qout_bg, n_frames, npx, alpha, fa):
    idx = self.mem.from_host(np.ascontiguousarray(stat_idx, dtype=np.int32))
    self.call("ipb_fa_params", hres.stats.ptr, idx.ptr, qout_bg.ptr, int(n_frames), int(npx),
                  float(np.float32(alpha)), fa.ptr, self.mem.stream)


def _engine_fret_pixels(self, planes, F, H, W, cfg, fparams, union=None, union_wpr=0, R=None,
                        Ralt=None, Rroi=None, Dcorr=None, Acorr=None, union_idx=None):
    cfg = np.ascontiguousarray(cfg, dtype=FRET_CFG).reshape(1)
    p = lambda b: b.ptr if b is not None else None
    self.call("ipb_fret_pixels", planes.ptr, int(F), int(H), int(W), cfg.ctypes.data, fparams.ptr,
                  p(union), int(union_wpr), p(union_idx), p(R), p(Ralt), p(Rroi), p(Dcorr), p(Acorr),
                  None, None, 0, self.mem.stream)


def _engine_region_stats(self, regions, jobs, mask_pool, H, W, planes=None, images=None, bvals=None,
                         and_bits=None, and_wpr=0):
    """regions: REGION array, jobs: STAT_JOB array (output rows are assigned here when left
    zero: one row per view, in job order).  Returns device buffer of STAT_OUT."""
    mem = self.mem
    regions = np.ascontiguousarray(regions, dtype=REGION)
    jobs = np.ascontiguousarray(jobs, dtype=STAT_JOB).copy()
    n = jobs.shape[0]
    jobs["n_views"] = np.maximum(jobs["n_views"], 1)
    if n and not jobs["out"].any():
        first = np.concatenate([[0], np.cumsum(jobs["n_views"])[:-1]])
        jobs["out"][:, 0] = first
        jobs["out"][:, 1] = first + 1
    n_out = int(jobs["n_views"].sum()) if n else 0
    srcs = np.unique(jobs["src"]) if n else np.zeros(0, np.int32)
    uniform = int(srcs[0]) if srcs.size == 1 else -1
    d_r = mem.from_host(regions if regions.shape[0] else np.zeros(1, REGION))
    d_j = mem.from_host(jobs if n else np.zeros(1, STAT_JOB))
    out = mem.empty(max(n_out, 1), STAT_OUT)
    p = lambda b: b.ptr if b is not None else None
    self.call("ipb_region_stats", d_r.ptr, d_j.ptr, n, uniform, mask_pool.ptr, p(and_bits), int(and_wpr),
                  int(H), int(W), p(planes), p(images), p(bvals), out.ptr, None, mem.stream)
    out._keep = (d_r, d_j)
    return out


Engine.hist = _engine_hist
Engine.quantiles = _engine_quantiles
Engine.scatter_qvalues = _engine_scatter_qvalues
Engine.fret_eps = _engine_fret_eps
Engine.fa_params = _engine_fa_params
Engine.fret_pixels = _engine_fret_pixels
Engine.region_stats = _engine_region_stats


def regions_from_masks(rm):
    """REGION rows for every ROI of a RoiMasks (frame coordinates)."""
    t = rm.table
    reg = np.zeros(t.n, dtype=REGION)
    reg["mask_off"] = t.mask_off[:-1]
    reg["x0"] = t.org[:, 0] + t.srect[:, 0]
    reg["y0"] = t.org[:, 1] + t.srect[:, 1]
    reg["w"] = t.srect[:, 2] - t.srect[:, 0]
    reg["h"] = t.srect[:, 3] - t.srect[:, 1]
    reg["wpr"] = t.wpr
    reg["frame"] = t.frame
    return reg


class FaResult:
    """Device-resident result of Engine.fa_segment()."""

    def __init__(self, crops, bw, comp_off, comps, labels, cap, keep):
        self.crops, self.bw, self.comp_off, self.comps, self.labels, self.cap = crops, bw, comp_off, comps, labels, cap
        self._keep = keep

    def bw_host(self, i):
        c = self.crops[i]
        words = self.bw.host()[c["bit_off"]: c["bit_off"] + c["h"] * c["wpr"]].reshape(c["h"], c["wpr"])
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")
        return bits[:, :c["w"]].astype(bool)

    def labels_host(self, i):
        c = self.crops[i]
        return self.labels.host()[c["pix_off"]: c["pix_off"] + c["h"] * c["w"]].reshape(c["h"], c["w"])


def crops_from_masks(rm, planes_of_roi):
    """CROP rows for every (skimage-rule, full-crop) ROI of a RoiMasks."""
    t = rm.table
    cr = np.zeros(t.n, dtype=CROP)
    w = (t.srect[:, 2] - t.srect[:, 0]).astype(np.int64)
    h = (t.srect[:, 3] - t.srect[:, 1]).astype(np.int64)
    cr["bit_off"] = t.mask_off[:-1]
    cr["mask_off"] = t.mask_off[:-1]
    cr["pix_off"][1:] = np.cumsum(w * h)[:-1]
    cr["row_off"][1:] = np.cumsum(h)[:-1]
    cr["ox"] = t.org[:, 0] + t.srect[:, 0]
    cr["oy"] = t.org[:, 1] + t.srect[:, 1]
    cr["w"], cr["h"], cr["wpr"] = w, h, t.wpr
    cr["plane"] = planes_of_roi
    cr["frame"] = t.frame
    return cr, int((w * h).sum()), int(h.sum())


def _engine_fa_segment(self, rm, crops, total_px, total_rows, planes, H, W, fa_params, min_size,
                       close_radius, want_labels=False, comp_cap=None):
    mem = self.mem
    n = crops.shape[0]
    d_crops = mem.from_host(crops if n else np.zeros(1, CROP))
    words = max(rm.table.total_words, 1)
    bw_a, bw_b, bw_f, rootbits = (mem.empty(words, np.uint32) for _ in range(4))
    L = mem.empty(max(total_px, 1), np.int32)
    csize = mem.empty(max(total_px, 1), np.uint32)
    row_roots = mem.empty(max(total_rows, 1), np.int32)
    row_base = mem.empty(max(total_rows, 1), np.int32)
    crop_count = mem.empty(max(n, 1), np.int32)
    comp_off = mem.zeros(n + 1, np.int32)
    if comp_cap is None:   # rigorous bound on 8-connected components of an h x w image
        comp_cap = int(sum(((int(c["h"]) + 1) // 2) * ((int(c["w"]) + 1) // 2) for c in crops)) or 1
    comps = mem.empty(comp_cap, COMP)
    labels = mem.empty(max(total_px, 1), np.int32) if want_labels else None
    self.call("ipb_fa_segment", d_crops.ptr, n, int(crops["h"].max()) if n else 0, int(total_rows),
              planes.ptr, int(H), int(W), fa_params.ptr, rm.pool.ptr, float(min_size), int(close_radius),
              bw_a.ptr, bw_b.ptr, L.ptr, csize.ptr, rootbits.ptr, row_roots.ptr, row_base.ptr,
              crop_count.ptr, bw_f.ptr, comp_off.ptr, comps.ptr, int(comp_cap),
              labels.ptr if labels is not None else None, 0, None, 8, mem.stream)
    return FaResult(crops, bw_f, comp_off, comps, labels, comp_cap,
                    (d_crops, bw_a, bw_b, rootbits, L, csize, row_roots, row_base, crop_count))


Engine.fa_segment = _engine_fa_segment


# ---------------------------------------------------------------------- morphology / moments / previews
def ball_gmax(d2max):
    """gmax table of the Euclidean ball dx^2 + dy^2 <= d2max (integers)."""
    R = int(math.isqrt(int(d2max)))
    return np.array([math.isqrt(int(d2max) - dx * dx) for dx in range(R + 1)], dtype=np.uint8), R


def square_gmax(p):
    return np.full(int(p) + 1, int(p), dtype=np.uint8), int(p)


def rim_d2max(rim_px):
    """Largest integer d2 with sqrt(d2) <= rim_px as numpy evaluates `dist_in <= rim_px` on the
    float64 EDT (reference Nesprin2_FRET_Builder.py:412-413); rim_px may be fractional."""
    r = float(rim_px)
    d2 = int(math.floor(r * r)) + 2
    while d2 > 0 and not (math.sqrt(d2) <= r):
        d2 -= 1
    return d2


def _engine_region_dilate(self, regions, in_pool, gmax, R, invert=False, and_pool=None, andnot_pool=None,
                          out_pool=None, d_regions=None):
    """Dilates every region mask (REGION rows; all pools share mask_off).  Returns the out pool."""
    mem = self.mem
    regions = np.ascontiguousarray(regions, dtype=REGION)
    n = regions.shape[0]
    if out_pool is None:
        out_pool = mem.empty(in_pool.shape, np.uint32)
    if n == 0:
        return out_pool
    w = regions["w"].astype(np.int64)
    h = regions["h"].astype(np.int64)
    g_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(w * h, out=g_off[1:])
    d_r = d_regions if d_regions is not None else mem.from_host(regions)
    d_goff = mem.from_host(g_off)
    g = mem.empty(max(int(g_off[-1]), 1), np.uint8)
    gm = np.ascontiguousarray(gmax, dtype=np.uint8)
    p = lambda b: b.ptr if b is not None else None
    self.call("ipb_region_dilate", d_r.ptr, n, int(w.max()), int(h.max()), in_pool.ptr, int(bool(invert)),
              gm.ctypes.data, int(R), g.ptr, d_goff.ptr, p(and_pool), p(andnot_pool), out_pool.ptr, mem.stream)
    out_pool._keep = (d_r, d_goff, g)
    return out_pool


def _engine_region_moments(self, regions, mask_pool):
    """Exact integer sums per region: columns n, sx, sy, sxx, syy, sxy (uint64, frame coords)."""
    mem = self.mem
    regions = np.ascontiguousarray(regions, dtype=REGION)
    n = regions.shape[0]
    out = mem.empty((max(n, 1), 6), np.uint64)
    if n:
        d_r = mem.from_host(regions)
        self.call("ipb_region_moments", d_r.ptr, n, mask_pool.ptr, out.ptr, mem.stream)
        out._keep = d_r
    return out.host()[:n]


def _engine_preview_u16(self, images, px_per_image, n_images, lohi):
    """images: device float32 [n_images][px]; lohi: host float32 [n_images][3] = lo, hi, den."""
    mem = self.mem
    out = mem.empty((n_images, px_per_image), np.uint16)
    d_l = mem.from_host(np.ascontiguousarray(lohi, dtype=np.float32).reshape(-1))
    self.call("ipb_preview_u16", images.ptr, int(px_per_image), int(n_images), d_l.ptr, out.ptr, mem.stream)
    out._keep = d_l
    return out


Engine.region_dilate = _engine_region_dilate
Engine.region_moments = _engine_region_moments
Engine.preview_u16 = _engine_preview_u16
