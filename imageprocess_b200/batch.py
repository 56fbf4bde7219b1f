"""FrameBatchJob: the fused per-frame hot path for a batch of F frames resident in HBM
(BASELINE.json config C4).

One ``run()`` = geometry tables built with vectorised numpy -> ONE pinned H2D copy of all
tables -> ~25 kernel launches on the current stream with no host synchronisation in
between -> ONE D2H copy of the packed result tables (+ a second, exact-size copy of the
per-adhesion table).  Device workspace is allocated once and reused across steps.

Stages (any subset):
  "fret"  fret_ratio_builder.process_one_stage numeric body (reference
          src/FRET/fret_ratio_builder.py:454-474,493-507): backgrounds, epsilon, ratio image,
          per-ROI ratio / donor / acceptor statistics
  "int"   Fluor_INT._process_key_task numeric body (src/INT/Fluor_INT.py:839-870):
          background per channel, per-ROI 9 statistics per channel
  "fa"    FA_Analyzer batch body (src/INT/FA_Analyzer.py:984-1039): global stats, per-cell
          crop + skimage mask, analyze_fa_crop, per-adhesion table
Results are numpy structured tables; ``rows_*`` helpers turn them into the reference's row
dicts for the unchanged pandas writers.
"""
import hashlib
import math
import contextlib
import copy
import os

import numpy as np

from . import geometry as geo
from . import ops
from .ops import (COMP, CROP, FP_BA, FP_BD, FP_STRIDE, FRET_CFG, HIST_JOB, PAT_FULL, PAT_MASKED,
                  PAT_MASKED_STRIDE, PAT_STRIDE1D, PAT_STRIDE2D, Q_JOB, Q_OUT, QK_MEDIAN, QK_PCT,
                  REGION, ROI_JOB, SRC_F32, SRC_U16, STAT_JOB, STAT_OUT, q32_of)

_ALIGN = 256


def _al(n):
    return (int(n) + _ALIGN - 1) // _ALIGN * _ALIGN


class _Arena:
    """Named sections inside one byte buffer (host pinned or device)."""

    def __init__(self):
        self.sections, self.size = {}, 0

    def add(self, name, dtype, shape):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        self.sections[name] = (self.size, np.dtype(dtype), shape, nbytes)
        self.size = _al(self.size + max(nbytes, 1))

    def view(self, base_np, name):
        off, dt, shape, nbytes = self.sections[name]
        return base_np[off: off + nbytes].view(dt).reshape(shape)

    def ptr(self, base_ptr, name):
        return base_ptr + self.sections[name][0]


def _copy_records(a):
    """Private copy of a structured array by bytes (numpy copies structured dtypes field by field:
    1 ms for the 1.5 MB adhesion table of a step against 0.1 ms for the same bytes)."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).copy().view(a.dtype).reshape(a.shape)


class _Plan:
    """Host tables + sizes of one batch layout (see FrameBatchJob._plan_for)."""
    pass


def _digest(polys):
    h = hashlib.blake2b(digest_size=16)
    for P in polys:
        a = np.ascontiguousarray(np.asarray(P, dtype=np.float64))
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.digest()


class _Ticket:
    """A submitted, not yet collected step."""
    pass


class BatchResult:
    """Host tables of one run() (numpy) + handles of the device-resident images."""
    pass


class FrameBatchJob:
    def __init__(self, eng, shape, stages=("fret", "int", "fa"), fret_p=None, int_task=None,
                 fa_params=None, fa_px=0.112, donor_ch=0, acc_ch=1, fa_ch=0, int_channels=None,
                 want_roi_image=False, want_labels=False, fa_config=None, fa_save_ok_only=True, hist_select=None,
                 want_contours=False):
        self.eng, self.mem = eng, eng.mem
        self.F, self.C, self.H, self.W = (int(s) for s in shape)
        self.stages = tuple(s for s in ("fret", "int", "fa") if s in stages)
        self.fret_p, self.int_task, self.fa_user = fret_p, int_task, fa_params
        self.fa_px, self.fa_save_ok_only = fa_px, fa_save_ok_only
        self.fa_cfg = fa_config or (fa_um_to_px_config(fa_params, fa_px) if fa_params else None)
        self.donor_ch, self.acc_ch, self.fa_ch = donor_ch, acc_ch, fa_ch
        self.int_ch = list(int_channels) if int_channels is not None else list(range(self.C))
        self.want_roi_image, self.want_labels = want_roi_image, want_labels or want_contours
        self.want_contours = want_contours       # marching-squares cell records of every adhesion (needs the label maps)
        self._bufs = {}
        self._plans = {}
        self._plan_serial = 0
        # N ranks: every step stages its packed tables in a device ring; ONE all-gather per
        # `gather_every` collected steps (see _issue_gather), never one per step
        self.n_slots = 3                 # output slots: a ticket may be collected two submits later
        self.gather_every = int(os.environ.get("IPB_GATHER_EVERY", "2"))
        # buffers of the gather path (ring of staged steps, gathered blob, the destination's pinned copy):
        # a rank may run gather_rings x gather_every steps ahead of the slowest one.  Small groups keep the
        # tail short (the last group's tables reach the destination after the last step: all-gather + D2H
        # of N x gather_every steps); the gathers run on their own stream, so their count costs nothing
        self.gather_rings = int(os.environ.get("IPB_GATHER_RINGS", "4"))
        # how the step tables of N ranks reach the destination rank: "shm" = every rank downloads into a
        # shared-memory ring the destination reads in place (ranks of one host; parallel.ShmTableRing: no
        # device work, no collective, nothing couples the ranks); "nccl" = all-gather + one download on the
        # destination (ranks on several hosts, or when the shared segments cannot be had)
        self.gather_via = os.environ.get("IPB_GATHER_VIA", "shm")
        self._shm = None
        self._g_pos = 0                  # steps staged so far
        self._g_open = {}                # group -> [entries collected, their ring-copy events]
        self._g_issued = []              # issued gathers not yet handed out: dicts
        self._g_last = {}                # ring buffer index -> event of the last gather that read it
        self._g_total, self._g_count = None, 0
        self._priming = False
        self._slot_busy = {}         # output slot -> event of its last table download
        self._pc_hint = None         # adhesion rows fetched with the step's tables: 1.25x the largest step seen so far
        self._graphs = {}           # (plan, input buffer, output slot, full_hist, ...) -> (CUDA graph, ticket template) | (None, times seen)
        self.use_graphs = bool(int(os.environ.get("IPB_GRAPHS", "1")))
        self.window_misses = 0
        self.comp_refetches = 0      # steps repeated because they had more adhesions than rows staged
        self._miss_streak = 0        # consecutive steps whose sampled windows missed: two in a row make the
        self._sticky_full = False    # full histograms sticky (bright or constant planes miss every time)
        self._slot = 0
        self.dist = None            # torch.distributed module when the job is one rank of N (see parallel.py)
        self.gather_dst = 0
        self._gather_cap = None     # bytes every rank contributes to the per-step all-gather
        self.overlap = bool(int(os.environ.get("IPB_OVERLAP", "1")))   # branches of a step on side streams
        self.fa_path = 0            # ipb_fa_segment path: 0 auto, 1 one CTA per crop (shared memory), 2 one kernel per phase, 3 one CTA per crop (global memory)
        # percentiles by sampled windows (ipb_hist_select) instead of full histograms wherever the
        # plane passes allow it (ops.pq_servable): exact either way
        self.hist_select = bool(int(os.environ.get("IPB_HIST_SELECT", "1"))) if hist_select is None else bool(hist_select)
        # per-ROI statistics: ONE walk of each ROI for both channels and the ratio (ipb_roi_stats_fused,
        # sampled value windows); the regions it cannot serve are repeated by the full-histogram kernels
        self.fa_threshold = None     # "otsu": optional stage, Otsu's threshold of the frame instead of mean + alpha * std
        self.fused_roi = bool(int(os.environ.get("IPB_FUSED_ROI", "1")))
        # FA moments on the FRET pass instead of the percentile pass: measured 2.545 vs 2.468 ms per step
        # (the FRET pass loses 30 us, the percentile pass gains 25 us, the FA chain starts later): off
        self.fret_moments = bool(int(os.environ.get("IPB_FRET_MOMENTS", "0")))
        self.rf_ctas = int(os.environ.get("IPB_RF_CTAS", ops.RF_CTAS_PER_SM)) * eng.n_sms()      # persistent CTAs of the fused ROI kernel
        self.roi_fallbacks = 0       # regions repeated by the full-histogram kernels so far
        self.pq_min_px = 1 << 18     # smaller planes take the full histograms (the sample would be most of the plane)
        self._pin = None
        self.n_roi_px = 0
        self.union_wpr = (self.W + 31) // 32
        if "fret" in self.stages:
            assert fret_p is not None
        if "int" in self.stages:
            assert int_task is not None
        if "fa" in self.stages:
            assert self.fa_cfg is not None

    # ------------------------------------------------------------------ buffers
    def _dev(self, name, nbytes):
        b = self._bufs.get(name)
        if b is None or b.nbytes < nbytes:
            if b is not None:
                self._graphs.clear()            # captured graphs hold the old buffer's address
            b = self.mem.empty(int(nbytes * 1.25) + 256, np.uint8)
            self._bufs[name] = b
        return b

    def _pinned(self, name, nbytes):
        p = self._bufs.get(name)
        if p is None or p[0].nbytes < nbytes:
            if p is not None:
                self._graphs.clear()
            p = self.mem.pinned(int(nbytes * 1.25) + 256, np.uint8)
            self._bufs[name] = p
        return p

    # ------------------------------------------------------------------ plan (host tables)
    def _plan_for(self, polys_per_frame):
        """Host-side tables for one batch.  Frames that carry the same ROI set (a time-lapse
        stage read with one ROI JSON, reference Fluor_INT.py:333-346) share one rasterised
        mask per ROI: the unique sets are found by object identity first, by content digest
        otherwise.  Plans are cached by content, so a steady time-lapse pays for the numpy
        table building once; every step still uploads the tables and runs every kernel."""
        F = self.F
        if len(polys_per_frame) != F:
            raise ValueError(f"polys_per_frame has {len(polys_per_frame)} entries for {F} frames")
        set_of_frame = np.zeros(F, dtype=np.int32)
        usets, digests, by_id, by_dg = [], [], {}, {}
        for f, pl in enumerate(polys_per_frame):
            pl = pl if pl is not None else ()
            s = by_id.get(id(pl))
            if s is None:
                dg = _digest(pl)
                s = by_dg.get(dg)
                if s is None:
                    s = len(usets)
                    usets.append(pl)
                    digests.append(dg)
                    by_dg[dg] = s
                by_id[id(pl)] = s
            set_of_frame[f] = s
        key = (set_of_frame.tobytes(), tuple(digests))
        plan = self._plans.get(key)
        if plan is None:
            plan = self._build_plan(set_of_frame, usets)
            self._plan_serial += 1
            plan.serial = self._plan_serial      # graphs are keyed by it (an id() can come back after eviction)
            if len(self._plans) >= 4:
                old = self._plans.pop(next(iter(self._plans)))
                self._graphs = {k: v for k, v in self._graphs.items() if k[0] != old.serial}
            self._plans[key] = plan
        return plan

    def _build_plan(self, set_of_frame, usets):
        F, C, H, W = self.F, self.C, self.H, self.W
        st = self.stages
        pl = _Plan()
        need_mpl = pl.need_mpl = ("fret" in st) or ("int" in st)
        S = pl.S = max(len(usets), 1)
        verts, off, uset, uroi = geo.flatten_polys(usets)
        NU = pl.NU = off.shape[0] - 1
        ucnt = np.bincount(uset, minlength=S).astype(np.int64)
        ustart = np.concatenate([[0], np.cumsum(ucnt)])[:-1]
        cnt_f = ucnt[set_of_frame]
        NR = pl.NR = int(cnt_f.sum())
        frame = np.repeat(np.arange(F, dtype=np.int32), cnt_f)
        inst_off = np.concatenate([[0], np.cumsum(cnt_f)])
        uidx = (ustart[set_of_frame[frame]] + (np.arange(NR) - inst_off[frame])).astype(np.int64)
        pl.frame, pl.roi, pl.uidx = frame, uroi[uidx].astype(np.int32), uidx
        pl.set_of_frame = set_of_frame

        T = pl.T = _Arena()                       # uploaded tables
        if need_mpl:
            m_rect = geo.mpl_tables(verts, off, W, H)
            m_wpr, m_rows, m_moff = geo._mask_layout(m_rect)
            pl.m_words = max(int(m_moff[-1]), 1)
            pl.m_max_rows = int(m_rows.max()) if NU else 0
            pl.m_max_wpr = int(m_wpr.max()) if NU else 0
        if "fa" in st:
            fa = geo.fa_tables(verts, off, W, H)
            f_wpr, f_rows, f_moff = geo._mask_layout(fa["srect"])
            pl.f_words = max(int(f_moff[-1]), 1)
            pl.f_max_rows = int(f_rows.max()) if NU else 0
            pl.f_max_wpr = int(f_wpr.max()) if NU else 0
            fw = fa["srect"][:, 2].astype(np.int64)[uidx]
            fh = fa["srect"][:, 3].astype(np.int64)[uidx]
            iw = f_wpr.astype(np.int64)[uidx]
            pl.fa_rect = fa["crop_rect"][uidx]
        T.add("vert_off", np.int32, NU + 1)
        T.add("uset", np.int32, max(NU, 1))
        T.add("union_idx", np.int32, F)
        if need_mpl:
            T.add("m_verts", np.float64, (max(verts.shape[0], 1), 2))
            T.add("m_rect", np.int32, (max(NU, 1), 4))
            T.add("m_org", np.int32, (max(NU, 1), 2))
            T.add("m_moff", np.int64, NU + 1)
            T.add("regions", REGION, max(NR, 1))
        if "fa" in st:
            T.add("f_verts", np.float64, (max(verts.shape[0], 1), 2))
            T.add("f_erect", np.int32, (max(NU, 1), 4))
            T.add("f_srect", np.int32, (max(NU, 1), 4))
            T.add("f_org", np.int32, (max(NU, 1), 2))
            T.add("f_moff", np.int64, NU + 1)
            T.add("crops", CROP, max(NR, 1))
            T.add("crop_order", np.int32, max(NR, 1))

        # ---- histogram / quantile job tables
        has_rois = cnt_f > 0
        hist_jobs = []
        hidx = pl.hidx = {}
        if "fret" in st:
            p = self.fret_p
            masked = (p["bg_scope"] == "roi_union")
            for key, ch in (("fret_d", self.donor_ch), ("fret_a", self.acc_ch)):
                j = np.zeros(F, dtype=HIST_JOB)
                j["plane"] = np.arange(F) * C + ch
                j["mask_frame"] = set_of_frame
                j["pattern"] = np.where(masked & has_rois, PAT_MASKED, PAT_FULL)
                hidx[key] = sum(x.shape[0] for x in hist_jobs)
                hist_jobs.append(j)
        if "int" in st:
            t = self.int_task
            stride = int(t["bg_stride"]) if t.get("bg_stride") else 1
            masked = (t["bg_scope"] == "roi_union")
            for ci, ch in enumerate(self.int_ch):
                j = np.zeros(F, dtype=HIST_JOB)
                j["plane"] = np.arange(F) * C + ch
                j["mask_frame"] = set_of_frame
                j["k"] = stride
                if stride > 1:
                    j["pattern"] = np.where(masked & has_rois, PAT_MASKED_STRIDE, PAT_STRIDE1D)
                else:
                    j["pattern"] = np.where(masked & has_rois, PAT_MASKED, PAT_FULL)
                hidx[("int", ci)] = sum(x.shape[0] for x in hist_jobs)
                hist_jobs.append(j)
        # the FA channel's integer moments (mean / std of the frame) ride on the fused FRET pass when it
        # reads that channel anyway; else on the percentile pass over the plane
        pl.fa_mom_from_fret = ("fa" in st and "fret" in st and self.fa_ch in (self.donor_ch, self.acc_ch) and
                               self.fret_moments)
        if "fa" in st:
            j = np.zeros(F, dtype=HIST_JOB)
            j["plane"] = np.arange(F) * C + self.fa_ch
            j["pattern"], j["k"], j["moments"] = PAT_STRIDE2D, 10, 0 if pl.fa_mom_from_fret else 1
            hidx["fa"] = sum(x.shape[0] for x in hist_jobs)
            hist_jobs.append(j)
        hist_jobs = ops.check_hist_jobs(np.concatenate(hist_jobs)) if hist_jobs else np.zeros(0, dtype=HIST_JOB)
        NH = pl.NH = hist_jobs.shape[0]
        pl.has_ms = bool((hist_jobs["pattern"] == PAT_MASKED_STRIDE).any())
        pl.pq_ok = False             # set once the plane passes are known

        # params layout (float32): [fret F*4 | int F*Ci | fa F*4]
        Ci = pl.Ci = len(self.int_ch)
        P_FRET, P_INT, P_FA = 0, F * FP_STRIDE, F * FP_STRIDE + F * Ci
        pl.P_FRET, pl.P_INT, pl.P_FA = P_FRET, P_INT, P_FA
        NP = pl.NP = P_FA + F * 4
        qjobs, qdst = [], []
        qidx = pl.qidx = {}

        def add_q(key, h0, q32, dst):
            q = np.zeros(F, dtype=Q_JOB)
            q["hist"] = h0 + np.arange(F)
            q["q32"] = q32
            qidx[key] = sum(x.shape[0] for x in qjobs)
            qjobs.append(q)
            qdst.append(np.asarray(dst, dtype=np.int32))

        host_bg = pl.host_bg = {}
        if "fret" in st:
            p = self.fret_p
            per_ch = bool(p["per_channel_p"])
            d_p = float(p["donor_p"]) if per_ch else float(p["percentile"])
            a_p = float(p["fret_p"]) if per_ch else float(p["percentile"])
            pl.numer_is_acc = p["ratio_mode"] == "FRET/Donor"
            pct = p["bg_mode"] == "percentile"
            neg = np.full(F, -1)
            add_q("fret_d", hidx["fret_d"], q32_of(d_p), P_FRET + np.arange(F) * FP_STRIDE + FP_BD if pct else neg)
            add_q("fret_a", hidx["fret_a"], q32_of(a_p), P_FRET + np.arange(F) * FP_STRIDE + FP_BA if pct else neg)
            add_q("fret_eps", hidx["fret_d"] if pl.numer_is_acc else hidx["fret_a"], q32_of(p["eps_percentile"]), neg)
            if p["bg_mode"] == "hist-mode":
                host_bg["fret"] = (d_p, a_p)
        if "int" in st:
            t = self.int_task
            p_glob = float(t["percentile"])
            self._int_p = [float(t["ch_p_map"].get(self._ch_name(ci), p_glob)) if t.get("per_channel_p") else p_glob
                           for ci in range(Ci)]
            pct = t["bg_mode"] == "percentile"
            for ci in range(Ci):
                add_q(("int", ci), hidx[("int", ci)], q32_of(self._int_p[ci]),
                      P_INT + np.arange(F) * Ci + ci if pct else np.full(F, -1))
            if t["bg_mode"] == "hist-mode":
                host_bg["int"] = True
        if "fa" in st:
            add_q("fa", hidx["fa"], q32_of(1.0), np.full(F, -1))
        qjobs = np.concatenate(qjobs) if qjobs else np.zeros(0, dtype=Q_JOB)
        qdst = np.concatenate(qdst) if qdst else np.zeros(0, dtype=np.int32)
        NQ = pl.NQ = qjobs.shape[0]

        # ---- region-stat jobs.  Output rows per ROI: [fret: R, donor, acceptor][int: one per
        # channel].  uint16 views of the same channel (FRET builder + Fluor_INT backgrounds)
        # share one job; every job template is laid out contiguously: uint16 first, then float.
        rpr = pl.rpr = (3 if "fret" in st else 0) + (Ci if "int" in st else 0)
        pl.int_col0 = 3 if "fret" in st else 0
        q3 = ((QK_PCT, QK_MEDIAN, QK_PCT), (q32_of(5), 0.0, q32_of(95)))
        qmed = ((0, QK_MEDIAN, 0), (0.0, 0.0, 0.0))
        views = {}                                   # channel -> list of (bidx array, clip, slot, qspec)
        if "fret" in st:
            clipn = int(bool(self.fret_p["clip_neg"]))
            for ch, slot, out_slot in ((self.donor_ch, FP_BD, 1), (self.acc_ch, FP_BA, 2)):
                views.setdefault(ch, []).append((P_FRET + frame * FP_STRIDE + slot, clipn, out_slot, qmed))
        if "int" in st:
            clipn = int(bool(self.int_task["clip_neg"]))
            for ci, ch in enumerate(self.int_ch):
                views.setdefault(ch, []).append((P_INT + frame * Ci + ci, clipn, pl.int_col0 + ci, q3))
        rr = np.arange(NR)
        views_all = {ch: list(vl) for ch, vl in views.items()}
        u16_tmpl = []
        for ch, vl in views.items():
            while vl:
                take, vl = vl[:2], vl[2:]
                s = np.zeros(NR, dtype=STAT_JOB)
                s["region"], s["src"], s["plane"], s["n_views"] = rr, SRC_U16, frame * C + ch, len(take)
                qs = q3 if any(v[3] is q3 for v in take) else qmed
                s["qkind"], s["q32"] = qs
                for v, (bidx, clipn, out_slot, _) in enumerate(take):
                    s["bidx"][:, v] = bidx
                    s["clip_neg"][:, v] = clipn
                    s["out"][:, v] = rr * rpr + out_slot
                u16_tmpl.append(s)
        f32_tmpl = []
        if "fret" in st:
            s = np.zeros(NR, dtype=STAT_JOB)
            s["region"], s["src"], s["plane"], s["n_views"] = rr, SRC_F32, frame, 1
            s["bidx"] = -1
            s["qkind"], s["q32"] = q3
            s["out"][:, 0] = rr * rpr
            f32_tmpl.append(s)
        pl.n_u16 = NR * len(u16_tmpl)
        pl.n_f32 = NR * len(f32_tmpl)
        if NR and need_mpl:
            # longest jobs first (one CTA per job, scheduled in index order): the last wave of a
            # launch then holds the small regions.  Every view names its own output row, so the
            # job order is free.
            mr = m_rect[uidx]
            big_first = np.argsort(-((mr[:, 2] - mr[:, 0]).astype(np.int64) * (mr[:, 3] - mr[:, 1])), kind="stable")
            u16_tmpl = [np.concatenate(u16_tmpl)[np.argsort(np.tile(np.argsort(big_first, kind="stable"), len(u16_tmpl)), kind="stable")]] if u16_tmpl else []
            f32_tmpl = [t[big_first] for t in f32_tmpl]
        sj = np.concatenate(u16_tmpl + f32_tmpl) if (u16_tmpl or f32_tmpl) and NR else np.zeros(0, dtype=STAT_JOB)
        NS = pl.NS = sj.shape[0]
        pl.n_out = NR * rpr

        # ---- fused ROI jobs (ipb_roi_stats_fused): per region, channel pairs -- the FRET pair with
        #      its ratio first -- each channel with its views; same output rows as the jobs above,
        #      which stay as the rerun list for the regions the fused kernel flags
        fj = []
        if NR and need_mpl and W % 8 == 0:
            pairs = []
            if "fret" in st:
                pairs.append((self.donor_ch, self.acc_ch, True))
            rest = [ch for ch in views_all if not ("fret" in st and ch in (self.donor_ch, self.acc_ch))]
            for i in range(0, len(rest), 2):
                pairs.append((rest[i], rest[i + 1] if i + 1 < len(rest) else None, False))
            for c0, c1, ratio in pairs:
                j = np.zeros(NR, dtype=ROI_JOB)
                j["region"] = rr
                j["plane"][:, 0] = frame * C + c0
                j["plane"][:, 1] = frame * C + c1 if c1 is not None else -1
                j["bidx"] = -1
                for slot, ch in enumerate((c0, c1)):
                    vl = views_all.get(ch, []) if ch is not None else []
                    if len(vl) > 2:
                        raise ValueError("more than two views of one channel")
                    j["n_views"][:, slot] = len(vl)
                    qs = q3 if any(v[3] is q3 for v in vl) else qmed
                    j["qkind"][:, slot], j["q32"][:, slot] = qs
                    for v, (bidx, clipn, out_slot, _) in enumerate(vl):
                        j["bidx"][:, slot, v] = bidx
                        j["clip"][:, slot, v] = clipn
                        j["out"][:, slot, v] = rr * rpr + out_slot
                if ratio:
                    j["ratio_on"], j["ratio_out"] = 1, rr * rpr
                    j["fp_idx"] = P_FRET + frame * FP_STRIDE
                    j["numer_slot"] = 1 if pl.numer_is_acc else 0
                    j["ratio_clip_neg"] = int(bool(self.fret_p["clip_neg"]))
                    j["rqkind"], j["rq32"] = q3
                fj.append(j[big_first])
        fj = np.concatenate(fj) if fj else np.zeros(0, dtype=ROI_JOB)
        NF = pl.NF = fj.shape[0]
        if NF:
            mr = m_rect[uidx]
            rw, rh = (mr[:, 2] - mr[:, 0]).astype(np.int64), (mr[:, 3] - mr[:, 1]).astype(np.int64)
            pl.rf_stride = int(min(2 * ((rw + 16) * rh).max() + 16384, 1 << 19))   # words per CTA: see ipb_roi_stats_fused
            pl.rf_max_w, pl.rf_max_h = int(rw.max()), int(rh.max())             # what ipb_roi_stats_fused_sizes is asked with

        passes = ops.plane_passes(hist_jobs)
        pl.n_passes = passes.shape[0]
        pl.pq_ok = bool(pl.n_passes) and self.W % 8 == 0 and self.W >= 16 and self.H * self.W >= self.pq_min_px and \
            ops.pq_servable(hist_jobs, passes)
        T.add("hist_jobs", HIST_JOB, max(NH, 1))
        T.add("passes", ops.PLANE_PASS, max(pl.n_passes, 1))
        T.add("qjobs", Q_JOB, max(NQ, 1))
        T.add("qdst", np.int32, max(NQ, 1))
        T.add("stat_jobs", STAT_JOB, max(NS, 1))
        T.add("roi_jobs", ROI_JOB, max(NF, 1))
        T.add("fa_stat_idx", np.int32, F)

        # ---- fill the pinned table buffer (uploaded with one H2D copy every step)
        pl.pin_np, pl.pin_t = self.mem.pinned(T.size, np.uint8)
        V = lambda name: T.view(pl.pin_np, name)
        V("vert_off")[:] = off.astype(np.int32)
        if NU:
            V("uset")[:NU] = uset
        V("union_idx")[:] = set_of_frame
        if need_mpl:
            V("m_verts")[: verts.shape[0]] = verts
            V("m_rect")[:NU] = m_rect
            V("m_org")[:NU] = 0
            V("m_moff")[:] = m_moff
            if NR:
                r = V("regions")[:NR]
                mr = m_rect[uidx]
                r["mask_off"] = m_moff[:-1][uidx]
                r["x0"], r["y0"] = mr[:, 0], mr[:, 1]
                r["w"], r["h"] = mr[:, 2] - mr[:, 0], mr[:, 3] - mr[:, 1]
                r["wpr"], r["frame"], r["use_and"], r["and_plane"] = m_wpr[uidx], frame, 0, 0
        if "fa" in st:
            V("f_verts")[: verts.shape[0]] = fa["local_verts"]
            V("f_erect")[:NU] = fa["erect"]
            V("f_srect")[:NU] = fa["srect"]
            V("f_org")[:NU] = fa["org"]
            V("f_moff")[:] = f_moff
            bit_off = np.zeros(NR + 1, dtype=np.int64)
            pix_off = np.zeros(NR + 1, dtype=np.int64)
            row_off = np.zeros(NR + 1, dtype=np.int64)
            np.cumsum(iw * fh, out=bit_off[1:])
            np.cumsum(fw * fh, out=pix_off[1:])
            np.cumsum(fh, out=row_off[1:])
            if NR:
                c = V("crops")[:NR]
                c["bit_off"], c["pix_off"], c["row_off"] = bit_off[:-1], pix_off[:-1], row_off[:-1]
                c["mask_off"] = f_moff[:-1][uidx]
                org = fa["org"][uidx]
                c["ox"], c["oy"] = org[:, 0], org[:, 1]
                c["w"], c["h"], c["wpr"] = fw, fh, iw
                c["plane"], c["frame"], c["pad0"] = frame * C + self.fa_ch, frame, 0
            pl.fa_crops = V("crops")[:NR].copy()
            V("crop_order")[:NR] = np.argsort(-(fw * fh), kind="stable")       # biggest crops start first
            pl.fa_words = max(int(bit_off[-1]), 1)
            pl.total_px, pl.total_rows = int(pix_off[-1]), int(row_off[-1])
            pl.fa_max_h = int(fh.max()) if NR else 0
            pl.comp_cap = int((((fh + 1) // 2) * ((fw + 1) // 2)).sum()) or 1
        if NH:
            V("hist_jobs")[:NH] = hist_jobs
            V("passes")[: pl.n_passes] = passes
        if NQ:
            V("qjobs")[:NQ] = qjobs
            V("qdst")[:NQ] = qdst
        if NS:
            V("stat_jobs")[:NS] = sj
        if NF:
            V("roi_jobs")[:NF] = fj
        if "fa" in st:
            V("fa_stat_idx")[:] = hidx["fa"] + np.arange(F)

        # ---- device output arena (one D2H at the end)
        O = pl.O = _Arena()
        O.add("params", np.float32, max(NP, 1))
        O.add("m_area", np.uint32, max(NU, 1))
        O.add("f_area", np.uint32, max(NU, 1))
        O.add("stat_out", STAT_OUT, max(pl.n_out, 1))
        O.add("comp_off", np.int32, NR + 1)
        O.add("miss", np.uint32, 1)
        O.add("rf_flags", np.uint8, max(NR, 1))
        return pl

    # ------------------------------------------------------------------ the step
    def submit(self, planes, polys_per_frame, full_hist=False, _pos=None):
        """Enqueues one step on the current stream and returns a ticket for collect(); nothing
        here waits for the device, so consecutive steps of a time-lapse overlap the host's table
        unpacking with the device's next batches (three output slots: collect a ticket before
        the third submit after it).  full_hist = True forces exact full-range histograms instead
        of sample-selected windows (automatic after a window miss).

        A step whose plan, input buffer and output slot were already seen twice is captured into
        a CUDA graph (table upload, ~25 launches on four streams, result downloads) and replayed
        from then on: one launch instead of ~60 driver calls from Python.  In an N-rank job the
        step's staged tables are copied into the gather ring; the all-gather itself is issued by
        collect() once per `gather_every` steps."""
        mem = self.mem
        pl = self._plan_for(polys_per_frame)
        full_hist = bool(full_hist) or self._sticky_full
        slot = self._slot
        self._slot = (self._slot + 1) % self.n_slots
        # everything the enqueued work depends on besides the (fixed) job parameters
        key = (pl.serial, int(planes.ptr), slot, bool(full_hist), bool(self.hist_select), bool(self.fused_roi), bool(self.fret_moments),
               bool(self.overlap), int(self.fa_path), int(self.pq_min_px), self._staged(), self._pc_rows(pl))
        graphable = self.use_graphs and hasattr(mem, "graph") and self.eng.prof is None and not pl.host_bg and \
            self.fa_threshold is None
        ent = self._graphs.get(key) if graphable else None
        if self._slot_busy.get(slot) is not None:
            mem.wait_event(self._slot_busy[slot])            # the slot's previous download has read its stage
        if ent is not None and ent[0] is not None:
            graph, tmpl = ent
            graph.replay()
            tk = copy.copy(tmpl)
            tk.res = copy.copy(tmpl.res)
            self.eng.launches += tmpl.launches
        elif graphable and ent is not None and ent[1] >= 2:
            l0 = self.eng.launches
            try:
                graph, ctx = mem.graph()
                with ctx:
                    tmpl = self._enqueue(planes, polys_per_frame, full_hist, pl, slot)
                tmpl.launches = self.eng.launches - l0
                self._graphs[key] = (graph, tmpl)
                graph.replay()
                tk = copy.copy(tmpl)
                tk.res = copy.copy(tmpl.res)
            except Exception:                                   # capture refused: stay eager for this job
                self.use_graphs = False
                mem.sync()
                tk = self._enqueue(planes, polys_per_frame, full_hist, pl, slot)
        else:
            if graphable:
                self._graphs[key] = (None, (ent[1] if ent else 0) + 1)
            tk = self._enqueue(planes, polys_per_frame, full_hist, pl, slot)
        self._stage_for_gather(tk, _pos)
        done = mem.event()
        done.record()
        if self._shm is not None and tk.g_pos is not None:   # N ranks, one host: straight into the shared ring
            pin_np, pin_t = self._shm.entry(tk.g_pos)
        else:
            pin_np, pin_t = self._pinned(f"pin_stage{slot}", tk.stage_bytes)
        O, a_rows = tk.pl.O, 32 + _al(tk.pl.O.size)
        n_dl = min(tk.stage_bytes, pin_np.nbytes)            # (a ring entry holds `cap` bytes: rows beyond it stay behind)
        tk.pout_np = pin_np[32: 32 + O.size]
        tk.pc_np = pin_np[a_rows: n_dl]
        tk.pc_rows_host = (n_dl - a_rows) // COMP.itemsize
        with mem.side(5, [done]) as dl:                      # the download stream: behind this step only
            mem.download_async(pin_t, tk.d_stage, n_dl)
        tk.event = self._slot_busy[slot] = dl.event
        return tk

    def _pc_rows(self, pl):
        """Adhesion rows downloaded (and staged for the gather) with a step's tables; collect()
        fetches the rest when a step has more.  Before the job has seen a step: 96 per ROI."""
        if not hasattr(pl, "comp_cap"):                      # no FA stage in this job
            return 0
        return min(pl.comp_cap, self._pc_hint or max(4096, 96 * pl.NR))

    def _staged(self):
        """N > 1 ranks: does the step stage its tables for the gather?  Not on the very first setup
        step of prime(): the capacity all ranks agree on is sized from that step's adhesion count."""
        if self.dist is None or self.dist.get_world_size() == 1:
            return False
        return not (self._priming and self._pc_hint is None and "fa" in self.stages)

    def prime(self, planes, polys_per_frame):
        """Setup for a long run over `planes`' buffer: steps it until both output slots have their
        CUDA graphs (buffers allocated, kernels loaded, graphs captured), so that the first real
        step already is one graph launch."""
        self._priming = True             # setup steps stage their tables (same graphs) but take no part in the gathers
        try:
            for _ in range(3 * self.n_slots):
                self.run(planes, polys_per_frame)
        finally:
            self._priming = False

    def _enqueue(self, planes, polys_per_frame, full_hist, pl, slot):
        """The stream work of one step (see submit)."""
        eng, mem, F, C, H, W = self.eng, self.mem, self.F, self.C, self.H, self.W
        st = self.stages
        T, O, NR, NU, NH, NQ, NS = pl.T, pl.O, pl.NR, pl.NU, pl.NH, pl.NQ, pl.NS
        P_FRET, P_INT, P_FA, Ci, NP = pl.P_FRET, pl.P_INT, pl.P_FA, pl.Ci, pl.NP
        res = BatchResult()
        res.frame, res.roi, res.n_rois = pl.frame, pl.roi, NR
        lib_call = eng.call
        # independent branches of the step (FA chain, uint16 region statistics, FRET pass + ratio
        # statistics) go to side streams and meet again before the result download
        branch = mem.branch if self.overlap else (lambda i: contextlib.nullcontext())

        d_tab = self._dev("d_tables", T.size)
        mem.upload_async(d_tab, pl.pin_t, T.size)
        tp = lambda name: T.ptr(d_tab.ptr, name)
        d_out = self._dev("d_out", O.size)
        op = lambda name: O.ptr(d_out.ptr, name)

        mem.nvtx_mark("ipb:tables+raster")
        # ---- rasterise every unique ROI once
        if pl.need_mpl:
            m_pool = self._dev("m_pool", 4 * pl.m_words)
            d_union = self._dev("union", 4 * pl.S * H * self.union_wpr)
            # the rasterisers run beside the percentile chain (masked histogram jobs need the union first)
            with (branch(3) if pl.pq_ok else contextlib.nullcontext()):
                mem.zero_bytes(d_union, 4 * pl.S * H * self.union_wpr)
                lib_call("ipb_rasterize_rois", geo.RULE_MPL, NU, tp("m_verts"), tp("vert_off"), tp("m_rect"),
                         tp("m_rect"), tp("m_org"), tp("uset"), tp("m_moff"), pl.m_max_rows, pl.m_max_wpr,
                         m_pool.ptr, op("m_area"), d_union.ptr, self.union_wpr, H, mem.stream)
            union_ptr = d_union.ptr
        else:
            union_ptr = None
        if "fa" in st:
            f_pool = self._dev("f_pool", 4 * pl.f_words)
            with branch(2):
                lib_call("ipb_rasterize_rois", geo.RULE_SK, NU, tp("f_verts"), tp("vert_off"), tp("f_erect"),
                         tp("f_srect"), tp("f_org"), tp("uset"), tp("f_moff"), pl.f_max_rows, pl.f_max_wpr,
                         f_pool.ptr, op("f_area"), None, self.union_wpr, H, mem.stream)

        mem.nvtx_mark("ipb:percentiles")
        # ---- histograms -> percentiles -> per-frame scalars
        d_hist = self._dev("hist", 4 * 65536 * max(NH, 1))
        d_hstat = self._dev("hstat", 8 * 4 * max(NH, 1))
        d_scr = self._dev("rank_scratch", 8 * max(NH, 1) * H) if pl.has_ms else None
        d_qout = self._dev("qout", Q_OUT.itemsize * max(NQ, 1))
        mem.zero_bytes(d_out, O.sections["params"][3], O.sections["params"][0])
        mem.zero_bytes(d_out, O.sections["miss"][3], O.sections["miss"][0])
        use_select = NH and self.hist_select and not full_hist and not pl.host_bg and pl.pq_ok
        if use_select:
            # percentiles by sampled windows (exact; a window miss repeats the step with full histograms)
            d_hw = self._dev("hist_win", 4 * ops.PQ_WIN * NH)
            d_win = self._dev("hist_winrange", ops.HIST_WIN.itemsize * NH)
            d_cnt = self._dev("hist_cnt", 8 * NH)
            lib_call("ipb_hist_select", planes.ptr, H, W, tp("hist_jobs"), NH, tp("passes"), pl.n_passes, tp("qjobs"), NQ,
                     d_hw.ptr, d_win.ptr, d_cnt.ptr, d_hstat.ptr, d_qout.ptr, op("miss"), mem.stream)
            lib_call("ipb_scatter_qvalues", d_qout.ptr, tp("qdst"), NQ, op("params"), mem.stream)
        elif NH:
            lib_call("ipb_hist_planes", planes.ptr, H, W, tp("hist_jobs"), NH, tp("passes"), pl.n_passes, int(pl.has_ms),
                     union_ptr, self.union_wpr, d_scr.ptr if d_scr is not None else None, d_hist.ptr, d_hstat.ptr, mem.stream)
            lib_call("ipb_hist_quantiles", d_hist.ptr, d_hstat.ptr, tp("qjobs"), NQ, d_qout.ptr, mem.stream)
            lib_call("ipb_scatter_qvalues", d_qout.ptr, tp("qdst"), NQ, op("params"), mem.stream)
        if pl.host_bg:
            self._host_hist_mode(pl.host_bg, d_hist, pl.hidx, d_out, O, NH, P_FRET, P_INT, Ci)
        if "fret" in st:
            den_slot = FP_BD if pl.numer_is_acc else FP_BA
            lib_call("ipb_fret_eps", d_qout.ptr + Q_OUT.itemsize * pl.qidx["fret_eps"], F, den_slot,
                     int(bool(self.fret_p["clip_neg"])), 5.0, op("params") + 4 * P_FRET, mem.stream)
        fa_params_call = lambda: lib_call("ipb_fa_params", d_hstat.ptr, tp("fa_stat_idx"),
                                          d_qout.ptr + Q_OUT.itemsize * pl.qidx["fa"], F, H * W,
                                          float(np.float32(self.fa_cfg["alpha"])), op("params") + 4 * P_FA, mem.stream)
        late_fa = "fa" in st and pl.fa_mom_from_fret         # mean / std arrive with the FRET pass
        if "fa" in st and not late_fa:
            fa_params_call()
            if self.fa_threshold == "otsu":
                self._host_otsu(planes, d_out, O, P_FA)

        mem.join()                                   # masks and per-frame scalars are ready from here on
        use_fused = self.fused_roi and pl.NF > 0

        mem.nvtx_mark("ipb:fa_chain")
        # ---- focal adhesions
        if "fa" in st and NR and pl.total_px > 0:
            words = pl.fa_words
            bwA, bwB, bwF, rootb = (self._dev(n, 4 * words) for n in ("bwA", "bwB", "bwF", "rootbits"))
            d_L = self._dev("L", 4 * pl.total_px)
            d_cs = self._dev("csize", 4 * pl.total_px)
            d_rr = self._dev("row_roots", 4 * pl.total_rows)
            d_rb = self._dev("row_base", 4 * pl.total_rows)
            d_cc = self._dev("crop_count", 4 * NR)
            d_comps = self._dev("comps", COMP.itemsize * pl.comp_cap)
            d_lab = self._dev("labels", 4 * pl.total_px) if self.want_labels else None
            res.fa_bw = ops_view(bwF, np.uint32, (words,), mem)
            res.fa_labels = ops_view(d_lab, np.int32, (pl.total_px,), mem) if d_lab is not None else None
            res.fa_rec = res.fa_rec_count = None
            if self.want_contours:
                d_rec = self._dev("fa_rec", 8 * pl.total_px)
                d_recn = self._dev("fa_rec_count", 4 * NR)
                res.fa_rec = ops_view(d_rec, np.uint32, (pl.total_px, 2), mem)
                res.fa_rec_count = ops_view(d_recn, np.uint32, (NR,), mem)
            cfgf = self.fa_cfg

            def run_fa():
                with branch(2):
                    lib_call("ipb_fa_segment", tp("crops"), NR, pl.fa_max_h, pl.total_rows, planes.ptr, H, W,
                             op("params") + 4 * P_FA, f_pool.ptr,
                             float(cfgf["min_px"]) if cfgf["min_px"] > 0 else 0.0,
                             int(cfgf["close_radius"]) if cfgf["close_radius"] > 0 else 0,
                             bwA.ptr, bwB.ptr, d_L.ptr, d_cs.ptr, rootb.ptr, d_rr.ptr, d_rb.ptr, d_cc.ptr,
                             bwF.ptr, op("comp_off"), d_comps.ptr, pl.comp_cap,
                             d_lab.ptr if d_lab is not None else None, int(self.fa_path), tp("crop_order"), 8, mem.stream)
                    if self.want_contours:
                        lib_call("ipb_fa_contour_cells", tp("crops"), NR, int((pl.fa_crops["w"].astype(np.int64) * pl.fa_crops["h"]).max()),
                                 d_lab.ptr, res.fa_rec.ptr, res.fa_rec_count.ptr, mem.stream)
            if not late_fa:
                run_fa()
            fa_ran = True
        else:
            fa_ran = False
            if "fa" in st:
                mem.zero_bytes(d_out, O.sections["comp_off"][3], O.sections["comp_off"][0])
        mem.nvtx_mark("ipb:roi_stats")
        # ---- per-ROI statistics (side stream): one walk of each ROI for both channels and the ratio;
        #      else the uint16 jobs of the full-histogram kernel
        with branch(1):
            if use_fused:
                d_sc = self._dev("rf_scratch", 4 * pl.rf_stride * self.rf_ctas)
                d_ctr = self._dev("rf_counter", 256)
                d_wide = self._dev("rf_wide", pl.NF)
                lib_call("ipb_roi_stats_fused", tp("regions"), NR, tp("roi_jobs"), pl.NF, m_pool.ptr, H, W, planes.ptr,
                         op("params"), op("stat_out"), d_sc.ptr, pl.rf_stride, self.rf_ctas, d_ctr.ptr, op("rf_flags"),
                         d_wide.ptr, mem.stream)
            elif pl.n_u16:
                lib_call("ipb_region_stats", tp("regions"), tp("stat_jobs"), pl.n_u16, SRC_U16, m_pool.ptr, None, 0,
                         H, W, planes.ptr, None, op("params"), op("stat_out"), None, mem.stream)

        mem.nvtx_mark("ipb:fret_pixels")
        # ---- fused FRET pass
        d_R = None
        if "fret" in st:
            d_R = self._dev("R", 4 * F * H * W)
            d_Rroi = self._dev("Rroi", 4 * F * H * W) if self.want_roi_image else None
            cfg = fret_cfg(self.fret_p, C, self.donor_ch, self.acc_ch)
            lib_call("ipb_fret_pixels", planes.ptr, F, H, W, cfg.ctypes.data, op("params") + 4 * P_FRET,
                     union_ptr, self.union_wpr, tp("union_idx"), d_R.ptr, None,
                     d_Rroi.ptr if d_Rroi is not None else None, None, None,
                     d_hstat.ptr if late_fa else None, tp("fa_stat_idx") if late_fa else None,
                     int(self.fa_ch == self.acc_ch and self.fa_ch != self.donor_ch), mem.stream)
            if late_fa:                                      # the frame's mean / std are complete now
                fa_params_call()
                if fa_ran:
                    run_fa()
            res.R = ops_view(d_R, np.float32, (F, H, W), mem)
            res.R_roi = ops_view(d_Rroi, np.float32, (F, H, W), mem) if d_Rroi is not None else None

        mem.nvtx_mark("ipb:roi_stats_rerun")
        # ---- per-ROI statistics of the ratio image (full-histogram kernel), or -- after the fused ROI
        #      kernel -- the rerun of the regions it flagged (usually none: a few CTAs scan the flags)
        if use_fused:
            mem.join()
            if pl.n_u16:
                lib_call("ipb_region_stats", tp("regions"), tp("stat_jobs"), pl.n_u16, SRC_U16, m_pool.ptr, None, 0,
                         H, W, planes.ptr, None, op("params"), op("stat_out"), op("rf_flags"), mem.stream)
            if pl.n_f32:
                lib_call("ipb_region_stats", tp("regions"), tp("stat_jobs") + STAT_JOB.itemsize * pl.n_u16, pl.n_f32,
                         SRC_F32, m_pool.ptr, None, 0, H, W, planes.ptr, d_R.ptr, op("params"), op("stat_out"),
                         op("rf_flags"), mem.stream)
        elif pl.n_f32:
            lib_call("ipb_region_stats", tp("regions"), tp("stat_jobs") + STAT_JOB.itemsize * pl.n_u16, pl.n_f32,
                     SRC_F32, m_pool.ptr, None, 0, H, W, planes.ptr, d_R.ptr, op("params"), op("stat_out"), None, mem.stream)
        mem.join()
        if "fa" in st:
            res.fa_rect, res.fa_crops = pl.fa_rect, pl.fa_crops

        mem.nvtx_mark("ipb:results_stage")
        # ---- results: the step packs its tables into this slot's stage (device-to-device, inside
        #      the step's graph): 32-byte header | table arena | first adhesion rows.  submit() then
        #      downloads the stage with ONE D2H on the download stream, off the critical path: the next
        #      step's kernels do not queue behind this step's PCIe copy (nor, on the destination rank of
        #      an N-rank job, behind the gathered tables of the other ranks on the same copy engine).
        tk = _Ticket()
        tk.pl, tk.res, tk.fa_ran = pl, res, fa_ran
        tk.planes, tk.polys, tk.full_hist = planes, polys_per_frame, full_hist
        tk.pc_rows = self._pc_rows(pl) if fa_ran else 0              # usual batches fit; collect() repeats the step when not
        a_rows = 32 + _al(O.size)
        tk.stage_bytes = a_rows + COMP.itemsize * tk.pc_rows
        staged = self._staged()
        if staged and self._gather_cap is None:
            # N > 1 ranks: every rank sends the same number of bytes per step, a capacity agreed once
            # per job (max over ranks of the stage size + 64 KiB; + 25% while no step was seen yet).
            # Rows that do not fit are not sent; the receiver sees that from comp_off and reports it.
            world = self.dist.get_world_size()
            need = tk.stage_bytes
            self._gather_cap = _al(mem.all_reduce_max(need + (need // 4 if self._pc_hint is None else 0) + (1 << 16), self.dist))
            # every buffer of the gather path now, not inside the run: pinning the destination
            # rank's host buffers alone takes ~25 ms each (measured: it showed up as 1 ms per step
            # of a 24-step run when the second ring's buffers were first used inside it)
            K = self.gather_every
            if self.gather_via == "shm":
                from . import parallel
                ring = parallel.ShmTableRing(self.dist, mem, self._gather_cap, n_entries=16, dst=self.gather_dst)
                if ring.ok:
                    self._shm = ring
                else:
                    self.gather_via = "nccl"                  # no room for the segments (or no page-locking): collective path
            for b in range(self.gather_rings if self._shm is None else 0):
                self._dev(f"gather_ring{b}", self._gather_cap * K)
                self._dev(f"gather_all{b}", self._gather_cap * K * world)
                if self.dist.get_rank() == self.gather_dst:
                    self._pinned(f"pin_gather{b}", self._gather_cap * K * world)
        cap = self._gather_cap if staged else 0
        if staged and a_rows > cap:
            raise RuntimeError("table arena larger than the agreed gather capacity; pass gather_cap_bytes")
        rows_sent = min(tk.pc_rows, (cap - a_rows) // COMP.itemsize) if staged else tk.pc_rows
        d_stage = self._dev(f"stage{slot}", max(tk.stage_bytes, cap))
        # ranks may carry different ROI sets, hence different arena layouts: the header names the
        # section every receiver needs to read the adhesion table (comp_off: offset, entries).  One
        # pinned header per (slot, plan): a replayed graph reads it at replay time
        hdr_np, hdr_t = self._pinned(f"pin_hdr{slot}_{pl.serial}", 32)
        hdr_np[:32].view(np.int64)[:] = (O.size, rows_sent, O.sections["comp_off"][0], NR + 1)
        mem.upload_async(d_stage, hdr_t, 32)
        mem.copy_bytes(d_stage, 32, d_out, 0, O.size)
        if tk.pc_rows:
            mem.copy_bytes(d_stage, a_rows, d_comps, 0, COMP.itemsize * tk.pc_rows)
        self._pinned(f"pin_stage{slot}", tk.stage_bytes)          # (allocated here, outside the timed steps)
        tk.d_stage, tk.slot = d_stage, slot
        tk.g_stage = (d_stage, cap, self.dist.get_world_size()) if staged else None   # the collective: _issue_gather()
        mem.nvtx_mark(None)
        return tk

    # ------------------------------------------------------------------ N ranks: table gathers
    def begin_distributed(self, n_local_steps):
        """Agrees on the number of gathers of the job (ranks may own a different number of steps:
        shards differ by one frame block); call it on every rank before the first step.  Without
        it every rank must run the same number of steps."""
        if self.dist is None or self.dist.get_world_size() == 1:
            return
        groups = -(-int(n_local_steps) // self.gather_every)
        self._g_total = self.mem.all_reduce_max(groups, self.dist)

    def _stage_for_gather(self, tk, pos):
        """Copies the step's staged tables (header + table arena + first adhesion rows, written by
        the step itself) into entry `pos` of the gather ring and records when that is done."""
        tk.g_pos = None
        if getattr(tk, "g_stage", None) is None or self._priming:
            return
        mem, K = self.mem, self.gather_every
        d_stage, cap, world = tk.g_stage
        if pos is None:
            pos = self._g_pos
            self._g_pos += 1
        if self._shm is not None:                            # the download itself goes into the ring entry (submit)
            tk.g_pos, tk.g_ev = pos, None
            return
        g, i = divmod(pos, K)
        if (g - self.gather_rings) in self._g_open:
            raise RuntimeError("collect() earlier tickets first: the gather ring still holds their group")
        b = g % self.gather_rings
        ring = self._dev(f"gather_ring{b}", cap * K)
        if self._g_last.get(b) is not None and g not in self._g_open:
            mem.wait_event(self._g_last[b])              # the gather of group g - gather_rings has read this ring
            self._g_last[b] = None
        self._g_open.setdefault(g, [0, []])
        mem.copy_bytes(ring, i * cap, d_stage, 0, cap)
        ev = mem.event()
        ev.record()
        tk.g_pos, tk.g_ev = pos, ev

    def _entry_done(self, tk):
        """collect() has accepted the step's tables: its ring entry is final.  The K-th accepted
        entry of a group issues the group's all-gather."""
        if tk.g_pos is None:
            return
        if self._shm is not None:
            self._shm.publish(tk.g_pos)
            return
        g = tk.g_pos // self.gather_every
        ent = self._g_open[g]
        ent[0] += 1
        ent[1].append(tk.g_ev)
        if ent[0] == self.gather_every:
            self._issue_gather(g, self.gather_every)

    def _issue_gather(self, g, n_valid):
        """ONE NCCL all-gather (NVLink 5 / NVSwitch) of a group of `gather_every` staged steps, on
        its own stream behind the ring copies only, so no kernel of a later step queues behind
        the collective (which waits for the slowest rank).  The destination rank brings the
        gathered blob to the host; gathered() hands it out.  Always issued from the host in
        program order (a collective inside a replayed CUDA graph left the process group unable
        to shut down)."""
        mem, K = self.mem, self.gather_every
        cap, world = self._gather_cap, self.dist.get_world_size()
        b = g % self.gather_rings
        ring = self._dev(f"gather_ring{b}", cap * K)
        events = self._g_open.pop(g, [0, []])[1]
        for i in range(n_valid, K):                           # entries of a partial (or empty) last group
            mem.zero_bytes(ring, 32, i * cap)
        if n_valid < K:
            ev = mem.event()
            ev.record()
            events = events + [ev]
        d_all = self._dev(f"gather_all{b}", cap * K * world)
        rec = {"group": g, "n": n_valid, "np": None, "cap": cap, "world": world}
        with mem.side(4, events) as br:
            if not os.environ.get("IPB_DEBUG_NO_GATHER"):
                mem.all_gather_bytes(d_all, ring, cap * K, self.dist)
            if self.dist.get_rank() == self.gather_dst and not os.environ.get("IPB_DEBUG_NO_GATHER_D2H"):
                g_np, g_t = self._pinned(f"pin_gather{b}", cap * K * world)
                for r in range(world):                        # one copy per rank: a step's own small download on the
                    mem.download_async(g_t, d_all, cap * K, offset=r * cap * K)          # same copy engine slips in between
                rec["np"] = g_np
        rec["event"] = br.event
        self._g_last[b] = br.event
        self._g_issued.append(rec)
        self._g_count += 1

    def gathered(self, block=False, copy=False):
        """Finished gathers since the last call, oldest first: on the destination rank a list of
        {"group", "per_rank": [rank][entry] -> (table arena, adhesion rows, comp_off)}; [] elsewhere
        (the collectives are still waited for when block is set).

        The arrays are VIEWS of the pinned buffer the gather was downloaded into (two buffers,
        `gather_rings` buffers used in turn): they stay valid until the gather of group + gather_rings is
        issued, i.e. for at least (gather_rings - 1) x gather_every further steps.  copy = True hands out private copies instead (measured:
        copying 8 steps x N ranks of tables in one go stalls the destination rank's submit loop for
        ~6 ms per rank and group, which is what made N ranks slower than one -- profiles/README.md)."""
        out = []
        keep = _copy_records if copy else (lambda a: a)
        if self._shm is not None:
            # shared-memory ring: whatever became final since the last call, read in place.  The views are
            # valid until the NEXT call of gathered() / finish(), which releases them to their producers
            if self.dist.get_rank() != self.gather_dst:
                return out
            import time
            ring, world = self._shm, self.dist.get_world_size()
            run = getattr(self, "_shm_wait_run", None)
            prev = getattr(self, "_shm_prev", [0] * world)
            for r in range(world):
                ring.release(r, prev[r])                         # handed out by earlier calls
            per_rank, private = [[] for _ in range(world)], [0] * world
            t0 = time.perf_counter()
            while True:
                for r in range(world):
                    for _pos, blob in ring.poll_rank(r):
                        arena_b, rows, co_off, co_n = (int(v) for v in blob[:32].view(np.int64))
                        a0 = 32 + _al(arena_b)
                        arena = keep(blob[32: 32 + arena_b])
                        comp_off = arena[co_off: co_off + 4 * co_n].view(np.int32)
                        rows = min(rows, int(comp_off[-1]) if co_n else 0)
                        per_rank[r].append((arena, keep(blob[a0: a0 + COMP.itemsize * rows]).view(COMP), comp_off, co_off))
                if not block or run is None or ring.drained(run):
                    break
                # waiting for ranks that still have steps to run: do not sit on half of a producer's ring
                for r in range(world):
                    if ring.next[r] - ring.released[r] >= ring.n // 2:
                        for k in range(private[r], len(per_rank[r])):
                            a, c, o, co_off = per_rank[r][k]
                            a = a.copy()
                            per_rank[r][k] = (a, c.copy(), a[co_off: co_off + o.nbytes].view(np.int32), co_off)
                        private[r] = len(per_rank[r])
                        ring.release(r, ring.next[r])
                if time.perf_counter() - t0 > 120.0:
                    raise RuntimeError("finish(): a rank did not end its run")
                time.sleep(50e-6)
            self._shm_prev = list(ring.next)
            if any(per_rank):
                self._shm_seq = getattr(self, "_shm_seq", -1) + 1
                out.append({"group": self._shm_seq, "per_rank": [[e[:3] for e in ents] for ents in per_rank]})
            return out
        while self._g_issued:
            rec = self._g_issued[0]
            if block:
                rec["event"].synchronize()
            elif not rec["event"].query():
                break
            self._g_issued.pop(0)
            if rec["np"] is None:
                continue
            K, cap, world = self.gather_every, rec["cap"], rec["world"]
            blobs = rec["np"][: world * K * cap].reshape(world, K, cap)
            hdr = blobs[:, :, :32].copy().view(np.int64).reshape(world, K, 4)
            per_rank = []
            for r in range(world):
                ents = []
                for i in range(K):
                    arena_b, rows, co_off, co_n = (int(v) for v in hdr[r, i])
                    if arena_b <= 0:
                        continue
                    a0 = 32 + _al(arena_b)
                    arena = keep(blobs[r, i, 32: 32 + arena_b])
                    comp_off = arena[co_off: co_off + 4 * co_n].view(np.int32)
                    rows = min(rows, int(comp_off[-1]) if co_n else 0)      # rows staged vs rows the step produced
                    ents.append((arena, keep(blobs[r, i, a0: a0 + COMP.itemsize * rows]).view(COMP), comp_off))
                per_rank.append(ents)
            out.append({"group": rec["group"], "per_rank": per_rank})
        return out

    def finish(self, copy=False):
        """End of the job (or of a timed run): gathers the partial last group, issues the empty
        gathers a rank with fewer steps still owes (begin_distributed), waits for every gather
        and returns what gathered() has not handed out yet.  Every rank must call it."""
        if self.dist is None or self.dist.get_world_size() == 1 or self._gather_cap is None:
            return []
        if self._shm is not None:
            # every step of this rank is published (collect() did that): tell the destination the run is
            # over; the destination waits until it has every rank's steps.  No collective, no barrier
            self._shm.end_run(self._g_pos)
            self._shm_wait_run = self._shm.run
            try:
                return self.gathered(block=True, copy=copy)
            finally:
                self._shm_wait_run = None
        for g in sorted(self._g_open):
            self._issue_gather(g, self._g_open[g][0])
        while self._g_total is not None and self._g_count < self._g_total:
            g = self._g_pos // self.gather_every + 1 + self._g_count          # any unused group id
            self._g_open[g] = [0, []]
            if self._g_last.get(g % self.gather_rings) is not None:
                self.mem.wait_event(self._g_last[g % self.gather_rings])
            self._issue_gather(g, 0)
        out = self.gathered(block=True, copy=copy)
        # a later run starts on fresh groups
        self._g_pos = (self._g_pos + self.gather_every - 1) // self.gather_every * self.gather_every
        self._g_total, self._g_count = None, 0
        return out

    def collect(self, tk):
        """Waits for a submitted step and unpacks its host tables."""
        mem, F, st = self.mem, self.F, self.stages
        pl, res, O = tk.pl, tk.res, tk.pl.O
        NR, NU, NP, Ci = pl.NR, pl.NU, pl.NP, pl.Ci
        P_FRET, P_INT, P_FA = pl.P_FRET, pl.P_INT, pl.P_FA
        tk.event.synchronize()
        OV = lambda name: O.view(tk.pout_np, name)
        if int(OV("miss")[0]) != 0:
            # a sampled window missed a wanted rank: exact rerun with full histograms.  The rerun is
            # rank-local (it re-stages the SAME entry of the gather ring and issues no collective),
            # so ranks that miss and ranks that do not keep the same sequence of collectives
            self.window_misses += 1
            self._miss_streak += 1
            if self._miss_streak >= 2:                     # data that misses deterministically: stop sampling
                self._sticky_full = True
            return self.collect(self.submit(tk.planes, tk.polys, full_hist=True, _pos=tk.g_pos))
        if not tk.full_hist:
            self._miss_streak = 0
        if "fa" in st and tk.fa_ran:
            total = int(OV("comp_off")[NR])
            if total > pl.comp_cap:
                raise RuntimeError("fa_segment: component table overflow")
            want = -(-(total + total // 4 + 1024) // 4096) * 4096
            if self._pc_hint is None or want > self._pc_hint:         # grows only: a new size means new graphs
                self._pc_hint = want
            if total > min(tk.pc_rows, tk.pc_rows_host):
                if total <= tk.pc_rows:
                    raise RuntimeError("adhesion rows beyond the agreed table capacity of the N-rank job; pass a larger gather capacity")
                # more adhesions than the rows staged with the tables: the step is repeated with the
                # larger fetch (rank-local, same ring entry; the device table of a step is overwritten
                # by the steps behind it, so the missing rows cannot be fetched after the fact)
                self.comp_refetches += 1
                return self.collect(self.submit(tk.planes, tk.polys, full_hist=tk.full_hist, _pos=tk.g_pos))
        self._entry_done(tk)
        params = OV("params")[:NP].copy()
        res.d2h_bytes = tk.stage_bytes                       # what the step's one D2H copied
        # regions the fused ROI kernel left to the full-histogram kernels, by reason (ipb_roifused.cuh)
        why = np.bincount(OV("rf_flags")[:NR], minlength=6) if (self.fused_roi and pl.NF) else np.zeros(6, np.int64)
        res.roi_fallbacks = int(why[1:].sum())
        res.roi_fallback_why = {n: int(why[k]) for k, n in enumerate(("geometry", "windows", "empty", "rank", "list"), 1) if why[k]}
        self.roi_fallbacks += res.roi_fallbacks
        if "fret" in st:
            res.fret_params = params[P_FRET: P_FRET + F * FP_STRIDE].reshape(F, FP_STRIDE)
        if "int" in st:
            res.int_bg = params[P_INT: P_INT + F * Ci].reshape(F, Ci)
            res.int_p = list(self._int_p)
        if "fa" in st:
            res.fa_stats = params[P_FA: P_FA + F * 4].reshape(F, 4)
        if pl.need_mpl:
            res.area = OV("m_area")[:NU].copy()[pl.uidx] if NR else np.zeros(0, np.uint32)
            self.n_roi_px = int(res.area.sum())
        so = _copy_records(OV("stat_out")[:pl.n_out]).reshape(NR, pl.rpr) if pl.n_out else np.zeros((NR, 0), dtype=STAT_OUT)
        if "fret" in st:
            res.fret_stat = so[:, 0:3]
        if "int" in st:
            res.int_stat = so[:, pl.int_col0: pl.int_col0 + Ci]
        if "fa" in st:
            comp_off = OV("comp_off")[: NR + 1].copy()
            res.fa_comp_off = comp_off
            total = int(comp_off[-1]) if tk.fa_ran else 0
            res.fa_comps = tk.pc_np[: COMP.itemsize * total].copy().view(COMP) if total else np.zeros(0, dtype=COMP)
        return res

    def run(self, planes, polys_per_frame, full_hist=False, _pos=None):
        """One synchronous step: submit + collect."""
        return self.collect(self.submit(planes, polys_per_frame, full_hist, _pos))

    def _n_hist_planes(self):
        """Distinct uint16 planes per frame that the histogram stage reads."""
        chs = set()
        if "fret" in self.stages:
            chs |= {self.donor_ch, self.acc_ch}
        if "int" in self.stages:
            chs |= set(self.int_ch)
        if "fa" in self.stages:
            chs.add(self.fa_ch)
        return len(chs)

    def _ch_name(self, ci):
        names = getattr(self, "ch_names", None)
        return names[ci] if names else self.int_ch[ci] + 1

    def _host_hist_mode(self, host_bg, d_hist, hidx, d_out, O, NH, P_FRET, P_INT, Ci):
        """'hist-mode' background (non-default GUI option): the 2048-bin np.histogram level is
        derived on the host from the exact device histogram (<= 65536 distinct values)."""
        mem, F = self.mem, self.F
        mem.sync()
        hh = ops_view(d_hist, np.uint32, (NH, 65536), mem).host()
        off, _, _, nbytes = O.sections["params"]
        params = ops_view(d_out, np.uint8, (d_out.nbytes,), mem).host()[off: off + nbytes].view(np.float32).copy()
        if "fret" in host_bg:
            d_p, a_p = host_bg["fret"]
            for f in range(F):
                for key, slot, pp in (("fret_d", FP_BD, d_p), ("fret_a", FP_BA, a_p)):
                    lvl = hist_mode_level(hh[hidx[key] + f], pp)
                    params[P_FRET + f * FP_STRIDE + slot] = 0.0 if lvl is None else lvl
        if "int" in host_bg:
            for f in range(F):
                for ci in range(Ci):
                    lvl = hist_mode_level(hh[hidx[("int", ci)] + f], self._int_p[ci])
                    params[P_INT + f * Ci + ci] = 0.0 if lvl is None else lvl
        tmp = mem.from_host(params)
        mem.copy_bytes(d_out, off, tmp, 0, nbytes)

    def _host_otsu(self, planes, d_out, O, P_FA):
        """Optional stage: the FA threshold of every frame becomes Otsu's threshold of its FA channel
        (exact device histogram, 65536-term scan on the host; the step waits for it)."""
        from . import filters
        mem, F, C = self.mem, self.F, self.C
        thr = filters.threshold_otsu(self.eng, planes, self.H, self.W, [f * C + self.fa_ch for f in range(F)])
        off, _, _, nbytes = O.sections["params"]
        params = ops_view(d_out, np.uint8, (d_out.nbytes,), mem).host()[off: off + nbytes].view(np.float32).copy()
        params[P_FA + 3: P_FA + 4 * F: 4] = np.asarray(thr, dtype=np.float32)
        tmp = mem.from_host(params)
        mem.copy_bytes(d_out, off, tmp, 0, nbytes)

    def algorithmic_bytes(self, entry):
        """Compulsory bytes one launch of `entry` moves (DESIGN.md 'Kernels')."""
        F, C, H, W = self.F, self.C, self.H, self.W
        px = F * H * W
        roi_px = self.n_roi_px or 0
        n_hist = (2 if "fret" in self.stages else 0) + (len(self.int_ch) if "int" in self.stages else 0) + \
                 (1 if "fa" in self.stages else 0)
        entry = {"ipb_hist_select": "ipb_hist_planes"}.get(entry, entry)
        return {
            "ipb_hist_planes": 2 * px * self._n_hist_planes(),   # every sampled plane read once per launch
            "ipb_fret_pixels": 8 * px,                   # 2 x uint16 in, float32 ratio out
            # ratio job: float32 under the mask; one uint16 job per measured channel
            "ipb_region_stats": (4 * roi_px if "fret" in self.stages else 0) + 2 * roi_px * len(
                set(([self.donor_ch, self.acc_ch] if "fret" in self.stages else []) +
                    (list(self.int_ch) if "int" in self.stages else []))),
            # one read of both uint16 channels under the mask (the ratio is recomputed, not read)
            "ipb_roi_stats_fused": 2 * roi_px * len(
                set(([self.donor_ch, self.acc_ch] if "fret" in self.stages else []) +
                    (list(self.int_ch) if "int" in self.stages else []))),
            "ipb_rasterize_rois": roi_px // 8 + 1,       # bit masks written
            "ipb_fa_segment": 2 * roi_px,                # crop pixels of the FA channel read once
        }.get(entry, 0)


def ops_view(buf, dtype, shape, mem):
    from .device import DevBuf
    return DevBuf(buf.raw, dtype, shape, mem)


def fret_cfg(p, n_ch=2, donor_ch=0, acc_ch=1):
    cfg = np.zeros(1, dtype=FRET_CFG)
    cfg["numer_is_acceptor"] = int(p["ratio_mode"] == "FRET/Donor")
    cfg["clip_neg"] = int(bool(p["clip_neg"]))
    cfg["g_factor"] = 1.0
    cfg["donor_ch"], cfg["acc_ch"], cfg["aonly_ch"], cfg["n_ch"] = donor_ch, acc_ch, -1, n_ch
    return cfg


def fa_um_to_px_config(params, px_size):
    """FA_Analyzer.py:527-535."""
    return {"alpha": params["alpha"], "min_px": params["min_area_um"] / (px_size ** 2),
            "max_px": params["max_area_um"] / (px_size ** 2),
            "close_radius": params["close_radius"], "subtract_bg": params.get("subtract_bg", True)}


def hist_mode_level(counts, p):
    """'hist-mode' background (reference Fluor_INT.py:474-483) from an exact integer
    histogram: np.histogram of the distinct values weighted by their counts places every
    value in the same bin as np.histogram of the full sample (same float32 edges)."""
    vals = np.flatnonzero(counts)
    if vals.size == 0:
        return 0.0
    w = counts[vals].astype(np.float64)
    hist, bins = np.histogram(vals.astype(np.float32), bins=2048, weights=w)
    if hist.sum() <= 0:
        return None
    cdf = np.cumsum(hist).astype(float)
    cdf /= cdf[-1]
    idx = int(np.searchsorted(cdf, float(p) / 100.0, side="left"))
    if idx >= len(bins) - 1:
        return float(bins[-1])
    return float(0.5 * (bins[idx] + bins[idx + 1]))


# ====================================================================== tables -> reference rows
def _f32(x):
    return float(np.float32(x))


def _mean_std(o):
    n = int(o["n"])
    if n == 0:
        return 0, math.nan, math.nan
    mean = float(o["sum"]) / n
    return n, _f32(mean), _f32(math.sqrt(max(float(o["ssd"]) / n, 0.0)))


def rows_intensity(res, F, ch_names):
    """Per-frame lists of the dicts quantify_per_roi_multi returns (Fluor_INT.py:509-538)."""
    out = [[] for _ in range(F)]
    for r in range(res.n_rois):
        row = {"roi": int(res.roi[r]), "area_px": int(res.area[r])}
        for ci, ch in enumerate(ch_names):
            o = res.int_stat[r, ci]
            n, mean, std = _mean_std(o)
            if n == 0:
                st = dict(mean=math.nan, median=math.nan, std=math.nan, p5=math.nan, p95=math.nan,
                          vmin=math.nan, vmax=math.nan, vsum=math.nan, npx=0)
            else:
                st = dict(mean=mean, median=float(o["q"][1]), std=std, p5=float(o["q"][0]),
                          p95=float(o["q"][2]), vmin=float(o["vmin"]), vmax=float(o["vmax"]),
                          vsum=_f32(o["sum"]), npx=n)
            for k, v in st.items():
                row[f"ch{ch}_{k}"] = v
        out[int(res.frame[r])].append(row)
    return out


def rows_fret(res, F):
    """Per-frame lists of the dicts quantify_per_roi returns (fret_ratio_builder.py:342-362)."""
    out = [[] for _ in range(F)]
    for r in range(res.n_rois):
        o = res.fret_stat[r, 0]
        n, mean, std = _mean_std(o)
        row = {"roi": int(res.roi[r]), "area_px": int(res.area[r])}
        if n == 0:
            row.update({f"ratio_{k}": math.nan for k in ("mean", "median", "std", "p5", "p95")})
        else:
            row.update({"ratio_mean": mean, "ratio_median": float(o["q"][1]), "ratio_std": std,
                        "ratio_p5": float(o["q"][0]), "ratio_p95": float(o["q"][2])})
        for name, oo in (("donor", res.fret_stat[r, 1]), ("yfret", res.fret_stat[r, 2])):
            nn, mm, _ = _mean_std(oo)
            row[f"{name}_mean"] = mm if nn else math.nan
            row[f"{name}_median"] = float(oo["q"][1]) if nn else math.nan
        out[int(res.frame[r])].append(row)
    return out


FA_CATS = ("OK", "Large", "Small")


def fa_table(res, cfg):
    """Vectorised per-adhesion table with the reference's dtypes (FA_Analyzer.py:166-193):
    area float64, mean float32, integrated densities float64, centroid float64.  Per-crop values are
    spread with np.repeat (fancy indexing and boolean-mask stores cost 2-3x as much at ~50 k rows)."""
    comps, off = res.fa_comps, res.fa_comp_off
    n = comps.shape[0]
    counts = np.diff(off)
    crop = np.repeat(np.arange(res.n_rois), counts) if n else np.zeros(0, dtype=np.int64)
    label = (np.arange(n) - np.repeat(off[:-1], counts) + 1) if n else np.zeros(0, dtype=np.int64)
    area = comps["area"].astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean_raw = (comps["sum_i"].astype(np.float64) / area).astype(np.float32)
        cy = comps["sum_y"].astype(np.float64) / area
        cx = comps["sum_x"].astype(np.float64) / area
    frame = np.repeat(res.frame, counts) if n else np.zeros(0, dtype=np.int32)
    bg = np.repeat(res.fa_stats[res.frame, 2].astype(np.float32), counts) if n else np.zeros(0, dtype=np.float32)
    if cfg.get("subtract_bg", True):
        mean_corr = np.maximum(np.float32(0), mean_raw - bg)
    else:
        mean_corr = mean_raw
    cat = np.where(area < cfg["min_px"], np.int8(2), (area > cfg["max_px"]).view(np.int8))      # 0 OK, 1 Large, 2 Small
    return {"crop": crop, "frame": frame, "cell_id": np.repeat(res.roi, counts) if n else np.zeros(0, np.int32),
            "label": label, "cat": cat, "area": area, "mean_raw": mean_raw, "mean_corr": mean_corr,
            "int_den_raw": mean_raw.astype(np.float64) * area, "int_den_corr": mean_corr.astype(np.float64) * area,
            "cy": cy, "cx": cx, "bg": bg,
            "thr": np.repeat(res.fa_stats[res.frame, 3].astype(np.float32), counts) if n else bg}


def fa_items(res, cfg, contours=None):
    """Per-crop results dicts in analyze_fa_crop's format (FA_Analyzer.py:164-193).  contours: per
    crop {label: [polylines]} (contours.contours_of_crop); then 'contour' is the label's first one
    and a label without any is skipped, as the reference does (FA_Analyzer.py:168-170)."""
    t = fa_table(res, cfg)
    out = [{"OK": [], "Large": [], "Small": []} for _ in range(res.n_rois)]
    for i in range(t["label"].shape[0]):
        contour = None
        if contours is not None:
            cl = contours[int(t["crop"][i])].get(int(t["label"][i]), [])
            if not cl:
                continue
            contour = cl[0]
        mean_raw, mean_corr, area = t["mean_raw"][i], t["mean_corr"][i], t["area"][i]
        if cfg.get("subtract_bg", True) and not (mean_raw - t["bg"][i] > 0):
            mean_corr = 0                        # python max(0, x) keeps the int 0
        out[int(t["crop"][i])][FA_CATS[t["cat"][i]]].append({
            "label": int(t["label"][i]), "area": area, "contour": contour,
            "centroid": (float(t["cy"][i]), float(t["cx"][i])), "mean_int_raw": mean_raw,
            "mean_int_corr": mean_corr, "int_den_raw": mean_raw * area, "int_den_corr": mean_corr * area,
            "bg_level": t["bg"][i]})
    return out


def fa_contours(res):
    """Per crop {label: [polylines]} from the device's cell records of a step run with
    want_contours (read before the next submit: the records live in the job's device buffers)."""
    from . import contours as ct
    rec, cnt = res.fa_rec.host(), res.fa_rec_count.host()
    out = []
    for i in range(res.n_rois):
        c = res.fa_crops[i]
        n = min(int(cnt[i]), int(c["w"]) * int(c["h"]))
        out.append(ct.contours_of_crop(rec[int(c["pix_off"]): int(c["pix_off"]) + n], n, int(c["w"])))
    return out


def rows_fa(res, cfg, user_params, px_size, F, save_ok_only=True):
    """Per-frame CSV row dicts in the reference's order (FA_Analyzer.py:1019-1039): per cell,
    categories OK / Large / Small, labels ascending."""
    t = fa_table(res, cfg)
    out = [[] for _ in range(F)]
    n = t["label"].shape[0]
    if n == 0:
        return out
    order = np.lexsort((t["label"], t["cat"], t["crop"]))
    for i in order:
        cat = FA_CATS[t["cat"][i]]
        if save_ok_only and cat != "OK":
            continue
        area = t["area"][i]
        out[int(t["frame"][i])].append({
            "Cell_ID": int(t["cell_id"][i]), "Category": cat, "Area_px": area,
            "Area_um2": area * (px_size ** 2), "Mean_Intensity_Raw": t["mean_raw"][i],
            "Mean_Intensity_Corr": t["mean_corr"][i], "Int_Density_Raw": t["int_den_raw"][i],
            "Int_Density_Corr": t["int_den_corr"][i], "Background_Level": t["bg"][i],
            "Used_Alpha": user_params["alpha"], "Global_Threshold": t["thr"][i],
            "Min_Area_Setting": user_params["min_area_um"], "Max_Area_Setting": user_params["max_area_um"],
            "Close_Radius_Setting": user_params["close_radius"],
            "Subtract_BG_Setting": user_params.get("subtract_bg", True)})
    return out
