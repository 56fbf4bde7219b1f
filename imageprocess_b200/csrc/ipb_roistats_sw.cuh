// ipb_roistats_sw.cuh -- per-region statistics with exact order statistics by SAMPLED WINDOWS.
// Included by ipb_roistats.cuh; same jobs, same outputs, same exactness as ipb_k_region_stats
// (quantify_stats, quantify_per_roi).
//
// ipb_k_region_stats histograms every pixel of a region with one shared-memory atomic (~2 per
// clock per SM on B200) from a single 1024-thread CTA per SM whose ~10 dependent phases cannot
// overlap.  Here a region is measured by a 256-thread CTA (4 resident per SM, persistent over the
// job list, largest regions first) with almost no atomics:
//   walk     row popcounts -> row offsets (block scan); a warp per mask row gathers the region's
//            pixels with 4 loads in flight and writes their ordered-uint32 keys, compacted in
//            raster order, to the CTA's slice of a global scratch buffer (L2 resident: later
//            passes read it back coalesced).  n, min, max and the float64 sums ride along.
//   sample   every s-th key (<= 2048) is bitonic-sorted in shared memory.  For each wanted
//            quantile the sorted sample gives a key window [klo, khi] that holds the wanted ranks
//            with overwhelming probability (sample position -/+ 5 sigma), and its ~200 sample
//            keys inside the window are the SPLITTERS of a fine histogram.
//   count    one pass over the keys: keys below a window are counted in registers; keys inside
//            are binary-searched among the splitters (the only atomics, ~20 % of the keys).
//            float32: the squared deviations from the exact mean ride along.
//   collect  the bin that holds a wanted rank spans one sample interval (~n / 2048 keys); a
//            second pass counts the keys equal to its lower splitter and lists the keys strictly
//            inside; the rank is resolved by brute force on that short list.
// Anything that does not fit (AND planes, > 2048 rows, more pixels than the scratch slice, a rank
// outside its window, a list overflow) raises *miss: the caller repeats the step with
// ipb_k_region_stats, so results are exact in every case.
#pragma once

#define IPB_SW_THREADS 256
#define IPB_SW_SAMP 2048
#define IPB_SW_BINS 512
#define IPB_SW_LIST 1024
#define IPB_SW_MAXROWS 2048
#define IPB_SW_TARGETS (2 * IPB_RS_MAXQ)

struct IpbSwSh {
    unsigned samp[IPB_SW_SAMP];
    unsigned rowbase[IPB_SW_MAXROWS];
    unsigned ihist[IPB_RS_MAXQ][IPB_SW_BINS];
    unsigned tlist[IPB_SW_TARGETS][IPB_SW_LIST];
    unsigned klo[IPB_RS_MAXQ], khi[IPB_RS_MAXQ], below[IPB_RS_MAXQ];
    int wa[IPB_RS_MAXQ], wT[IPB_RS_MAXQ], wact[IPB_RS_MAXQ];
    unsigned tlo[IPB_SW_TARGETS], thi[IPB_SW_TARGETS], trr[IPB_SW_TARGETS], teq[IPB_SW_TARGETS], tn[IPB_SW_TARGETS];
    unsigned tval[IPB_SW_TARGETS];
    int tact[IPB_SW_TARGETS], talias[IPB_SW_TARGETS];
    unsigned wred[IPB_SW_THREADS / 32][4];
    double dred[IPB_SW_THREADS / 32][4];
    unsigned ns_v;
    int missed;
};

// number of splitters <= key among S[0 .. cnt)  (S ascending)
__device__ __forceinline__ unsigned ipb_sw_upper_bound(const unsigned* S, unsigned cnt, unsigned key) {
    unsigned lo = 0, hi = cnt;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (S[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int SRC>
__global__ void __launch_bounds__(IPB_SW_THREADS, 4)
ipb_k_region_stats_sw(const IpbRegion* __restrict__ regions, const IpbStatJob* __restrict__ jobs, int n_jobs,
                      const unsigned* __restrict__ mask_pool, int H, int W,
                      const unsigned short* __restrict__ planes, const float* __restrict__ images,
                      const float* __restrict__ bvals, IpbStatOut* __restrict__ out,
                      unsigned* __restrict__ scratch, unsigned long long stride, unsigned* __restrict__ miss)
{
    __shared__ IpbSwSh sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = IPB_SW_THREADS / 32;
    const unsigned lt = (1u << lane) - 1u;
    const float fnan = __uint_as_float(0x7fc00000u);
    unsigned* keys = scratch + (size_t)blockIdx.x * stride;

    for (int jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
        __syncthreads();                                          // shared state of the previous job is dead
        const IpbStatJob job = jobs[jb];
        if (job.src != SRC) continue;                             // block-uniform
        const IpbRegion rg = regions[job.region];
        const int nv = (SRC == IPB_SRC_U16) ? (job.n_views < 1 ? 1 : (job.n_views > IPB_RS_MAXV ? IPB_RS_MAXV : job.n_views)) : 1;
        float vB[IPB_RS_MAXV]; int vclip[IPB_RS_MAXV];
#pragma unroll
        for (int v = 0; v < IPB_RS_MAXV; ++v) {
            vB[v] = (SRC == IPB_SRC_U16 && v < nv && job.bidx[v] >= 0) ? bvals[job.bidx[v]] : 0.0f;
            vclip[v] = (v < nv) ? job.clip_neg[v] : 0;
        }
        if (tid == 0) sh.missed = 0;
        if (rg.use_and || rg.h > IPB_SW_MAXROWS || rg.h <= 0 || rg.wpr <= 0) {     // block-uniform
            if (tid == 0) atomicAdd(miss, 1u);
            continue;
        }
        const unsigned* mask = mask_pool + rg.mask_off;
        const int h = rg.h, wpr = rg.wpr;

        // ---- row popcounts -> exclusive row offsets, area
        for (int r = tid; r < h; r += IPB_SW_THREADS) {
            unsigned c = 0;
            for (int j = 0; j < wpr; ++j) c += (unsigned)__popc(mask[(size_t)r * wpr + j]);
            sh.rowbase[r] = c;
        }
        __syncthreads();
        unsigned area;
        {
            const int per = (h + IPB_SW_THREADS - 1) / IPB_SW_THREADS;
            const int r0 = tid * per;
            unsigned cnt = 0;
            for (int k = 0; k < per; ++k) if (r0 + k < h) cnt += sh.rowbase[r0 + k];
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) sh.wred[warp][0] = incl;
            __syncthreads();
            unsigned before = 0, total = 0;
            for (int i = 0; i < nwarps; ++i) { const unsigned t = sh.wred[i][0]; total += t; if (i < warp) before += t; }
            unsigned run = before + incl - cnt;
            for (int k = 0; k < per; ++k) if (r0 + k < h) { const unsigned c = sh.rowbase[r0 + k]; sh.rowbase[r0 + k] = run; run += c; }
            area = total;
            __syncthreads();
        }
        if ((unsigned long long)area > stride) { if (tid == 0) atomicAdd(miss, 1u); continue; }

        // ---- walk: keys in raster order -> scratch; n, key range, float64 sums
        unsigned n_t = 0, lo_t = 0xffffffffu, hi_t = 0u;
        double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0;
        if (area > 0) {
            const unsigned short* u16 = (SRC == IPB_SRC_U16) ? planes + (size_t)job.plane * H * W : nullptr;
            const float* f32 = (SRC == IPB_SRC_F32) ? images + (size_t)job.plane * H * W : nullptr;
            for (int r = warp; r < h; r += nwarps) {
                unsigned base = sh.rowbase[r];
                const unsigned* mrow = mask + (size_t)r * wpr;
                const size_t p0 = (size_t)(rg.y0 + r) * W + rg.x0 + lane;
                for (int j0 = 0; j0 < wpr; j0 += 4) {
                    unsigned m[4], raw[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) m[u] = (j0 + u < wpr) ? mrow[j0 + u] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        raw[u] = 0u;
                        if ((m[u] >> lane) & 1u) {
                            const size_t a = p0 + 32 * (size_t)(j0 + u);
                            if (SRC == IPB_SRC_U16) raw[u] = (unsigned)u16[a]; else raw[u] = __float_as_uint(f32[a]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if ((m[u] >> lane) & 1u) {
                            const unsigned pos = base + (unsigned)__popc(m[u] & lt);
                            unsigned key;
                            if (SRC == IPB_SRC_U16) {
                                key = raw[u];
                                const double t0 = (double)ipb_rs_transform(vB[0], vclip[0], key);
                                s0 += t0; q0 += t0 * t0;
                                if (nv > 1) { const double t1 = (double)ipb_rs_transform(vB[1], vclip[1], key); s1 += t1; q1 += t1 * t1; }
                            } else {
                                const float v = __uint_as_float(raw[u]);
                                if (isfinite(v)) { key = ipb_f32_key(v); s0 += (double)v; }
                                else key = IPB_RS_SENTINEL;
                            }
                            keys[pos] = key;
                            if (key != IPB_RS_SENTINEL) { ++n_t; lo_t = key < lo_t ? key : lo_t; hi_t = key > hi_t ? key : hi_t; }
                        }
                        base += (unsigned)__popc(m[u]);
                    }
                }
            }
        }
        // ---- block reductions
        n_t = ipb_warp_sum(n_t); lo_t = ipb_warp_min(lo_t); hi_t = ipb_warp_max(hi_t);
        s0 = ipb_warp_sum(s0); q0 = ipb_warp_sum(q0); s1 = ipb_warp_sum(s1); q1 = ipb_warp_sum(q1);
        if (lane == 0) {
            sh.wred[warp][0] = n_t; sh.wred[warp][1] = lo_t; sh.wred[warp][2] = hi_t;
            sh.dred[warp][0] = s0; sh.dred[warp][1] = q0; sh.dred[warp][2] = s1; sh.dred[warp][3] = q1;
        }
        __syncthreads();                                          // also publishes the keys to the whole CTA
        unsigned long long n = 0;
        unsigned kmin = 0xffffffffu, kmax = 0u;
        double S0 = 0.0, Q0 = 0.0, S1 = 0.0, Q1 = 0.0;
        for (int i = 0; i < nwarps; ++i) {
            n += sh.wred[i][0];
            kmin = sh.wred[i][1] < kmin ? sh.wred[i][1] : kmin;
            kmax = sh.wred[i][2] > kmax ? sh.wred[i][2] : kmax;
            S0 += sh.dred[i][0]; Q0 += sh.dred[i][1]; S1 += sh.dred[i][2]; Q1 += sh.dred[i][3];
        }
        if (n == 0) {
            if (tid < nv) {
                IpbStatOut o;
                o.n = 0; o.area = area; o.sum = 0.0; o.ssd = 0.0; o.pad0 = 0.f;
                o.vmin = fnan; o.vmax = fnan;
                for (int i = 0; i < IPB_RS_MAXQ; ++i) o.q[i] = fnan;
                out[job.out[tid]] = o;
            }
            continue;
        }
        double v_sum[IPB_RS_MAXV], v_ssd[IPB_RS_MAXV];
        v_sum[0] = S0; v_sum[1] = S1;
        v_ssd[0] = 0.0; v_ssd[1] = 0.0;
        if (SRC == IPB_SRC_U16) {
            // counts times values with <= 17 significant bits: sumsq - sum^2 / n in float64, as in
            // ipb_k_region_stats' histogram path; a constant region is exactly 0
            v_ssd[0] = (kmin == kmax) ? 0.0 : Q0 - S0 * (S0 / (double)n);
            v_ssd[1] = (kmin == kmax) ? 0.0 : Q1 - S1 * (S1 / (double)n);
        }
        const double f_mean = S0 / (double)n;

        // ---- wanted ranks
        IpbQIdx qi[IPB_RS_MAXQ];
#pragma unroll
        for (int i = 0; i < IPB_RS_MAXQ; ++i) {
            qi[i].prev = qi[i].next = 0; qi[i].gamma = 0.f;
            if (job.qkind[i] == IPB_QKIND_PCT) qi[i] = ipb_np_qidx_f32((long long)n, job.q32[i]);
            else if (job.qkind[i] == IPB_QKIND_MEDIAN) {
                if (n & 1ull) qi[i].prev = qi[i].next = (long long)(n >> 1);
                else { qi[i].prev = (long long)(n >> 1) - 1; qi[i].next = (long long)(n >> 1); }
            }
        }

        // ---- sample: every s-th key, bitonic sort
        const unsigned sstep = (area + IPB_SW_SAMP - 1u) / IPB_SW_SAMP;
        const unsigned ns = (area + sstep - 1u) / sstep;
        unsigned P = 64;
        while (P < ns) P <<= 1;
        for (unsigned i = tid; i < P; i += IPB_SW_THREADS) sh.samp[i] = i < ns ? keys[(size_t)i * sstep] : IPB_RS_SENTINEL;
        __syncthreads();
        for (unsigned k = 2; k <= P; k <<= 1) {
            for (unsigned j = k >> 1; j > 0; j >>= 1) {
                for (unsigned i = tid; i < (P >> 1); i += IPB_SW_THREADS) {
                    const unsigned a = ((i & ~(j - 1u)) << 1) | (i & (j - 1u)), b = a | j;
                    const unsigned x = sh.samp[a], y = sh.samp[b];
                    const bool up = (a & k) == 0u;
                    if ((x > y) == up) { sh.samp[a] = y; sh.samp[b] = x; }
                }
                __syncthreads();
            }
        }
        if (tid == 0) sh.ns_v = ipb_sw_upper_bound(sh.samp, P, IPB_RS_SENTINEL - 1u);     // valid sample keys
        if (tid < IPB_SW_TARGETS) { sh.tact[tid] = 0; sh.talias[tid] = -1; sh.teq[tid] = 0; sh.tn[tid] = 0; sh.tval[tid] = 0; }
        if (tid < IPB_RS_MAXQ) { sh.wact[tid] = 0; sh.below[tid] = 0; }
        for (int i = tid; i < IPB_RS_MAXQ * IPB_SW_BINS; i += IPB_SW_THREADS) (&sh.ihist[0][0])[i] = 0u;
        __syncthreads();
        const unsigned ns_v = sh.ns_v;
        const bool direct = sstep == 1u;                          // the sorted sample IS the sorted region
        double ssd_t = 0.0;

        if (direct) {
            if (tid < IPB_SW_TARGETS && job.qkind[tid >> 1] != IPB_QKIND_NONE) {
                const long long r = (tid & 1) ? qi[tid >> 1].next : qi[tid >> 1].prev;
                sh.tval[tid] = sh.samp[r];
            }
            if (SRC != IPB_SRC_U16)
                for (unsigned i = tid; i < ns_v; i += IPB_SW_THREADS) { const double d = (double)ipb_key_f32(sh.samp[i]) - f_mean; ssd_t += d * d; }
        } else {
            // ---- windows: one per wanted quantile
            if (tid < IPB_RS_MAXQ && job.qkind[tid] != IPB_QKIND_NONE) {
                const int i = tid;
                const double scale = (double)ns_v / (double)n;
                const double f_lo = ((double)qi[i].prev + 0.5) * scale - 0.5, f_hi = ((double)qi[i].next + 0.5) * scale - 0.5;
                const double qq = ((double)qi[i].prev + 0.5) / (double)n;
                const double d = 5.0 * sqrt(fmax(qq * (1.0 - qq), 0.0) * (double)ns_v) + 3.0;
                long long a = (long long)floor(f_lo - d) - 1, b = (long long)ceil(f_hi + d) + 1;
                unsigned klo, khi;
                if (a < 0) { a = 0; klo = 0u; } else klo = sh.samp[a];
                if (b > (long long)ns_v - 1) { b = (long long)ns_v - 1; khi = IPB_RS_SENTINEL - 1u; } else khi = sh.samp[b];
                if (b < a) b = a;
                if (b - a + 2 > IPB_SW_BINS) sh.missed = 1;
                else { sh.wa[i] = (int)a; sh.wT[i] = (int)(b - a); sh.klo[i] = klo; sh.khi[i] = khi; sh.wact[i] = 1; }
            }
            __syncthreads();
            // ---- count pass
            {
                unsigned klo[IPB_RS_MAXQ], khi[IPB_RS_MAXQ], bl[IPB_RS_MAXQ];
                int wa[IPB_RS_MAXQ], wT[IPB_RS_MAXQ];
                bool act[IPB_RS_MAXQ];
#pragma unroll
                for (int i = 0; i < IPB_RS_MAXQ; ++i) { act[i] = sh.wact[i] != 0; klo[i] = sh.klo[i]; khi[i] = sh.khi[i]; wa[i] = sh.wa[i]; wT[i] = sh.wT[i]; bl[i] = 0; }
                for (unsigned idx = tid; idx < area; idx += IPB_SW_THREADS) {
                    const unsigned key = keys[idx];
                    if (key == IPB_RS_SENTINEL) continue;
                    if (SRC != IPB_SRC_U16) { const double d = (double)ipb_key_f32(key) - f_mean; ssd_t += d * d; }
#pragma unroll
                    for (int i = 0; i < IPB_RS_MAXQ; ++i) {
                        if (!act[i]) continue;
                        if (key < klo[i]) ++bl[i];
                        else if (key <= khi[i]) atomicAdd(&sh.ihist[i][ipb_sw_upper_bound(sh.samp + wa[i], (unsigned)wT[i] + 1u, key)], 1u);
                    }
                }
#pragma unroll
                for (int i = 0; i < IPB_RS_MAXQ; ++i) {
                    const unsigned t = ipb_warp_sum(bl[i]);
                    if (lane == 0 && t) atomicAdd(&sh.below[i], t);
                }
            }
            __syncthreads();
            // ---- the bin of every wanted rank: warp t <-> target t (quantile t / 2, prev / next)
            if (warp < IPB_SW_TARGETS && sh.wact[warp >> 1]) {
                const int t = warp, i = t >> 1;
                const unsigned long long r = (unsigned long long)((t & 1) ? qi[i].next : qi[i].prev);
                const unsigned long long below = sh.below[i];
                const unsigned nb = (unsigned)sh.wT[i] + 2u;
                bool found = false;
                if (r >= below) {
                    const unsigned long long rr = r - below;
                    const unsigned per = IPB_SW_BINS / 32;
                    unsigned c[IPB_SW_BINS / 32], mine = 0;
#pragma unroll
                    for (unsigned b = 0; b < per; ++b) { const unsigned bi = (unsigned)lane * per + b; c[b] = bi < nb ? sh.ihist[i][bi] : 0u; mine += c[b]; }
                    unsigned long long incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const unsigned long long x = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += x; }
                    unsigned long long acc = incl - mine;
                    if (rr >= acc && rr < incl) {
#pragma unroll
                        for (unsigned b = 0; b < per; ++b) {
                            if (rr >= acc && rr < acc + c[b]) {
                                const unsigned u = (unsigned)lane * per + b;
                                sh.tlo[t] = u == 0u ? sh.klo[i] : sh.samp[sh.wa[i] + (int)u - 1];
                                sh.thi[t] = u <= (unsigned)sh.wT[i] ? sh.samp[sh.wa[i] + (int)u] : sh.khi[i] + 1u;     // exclusive
                                sh.trr[t] = (unsigned)(rr - acc);
                                sh.tact[t] = 1;
                            }
                            acc += c[b];
                        }
                    }
                    found = __any_sync(IPB_FULL, rr >= incl - mine && rr < incl);
                }
                if (!found && lane == 0) sh.missed = 1;
            }
            __syncthreads();
            // next shares prev's bin (the usual case): one list serves both
            if (tid < IPB_RS_MAXQ) {
                const int a = 2 * tid, b = a + 1;
                if (sh.tact[a] && sh.tact[b] && sh.tlo[a] == sh.tlo[b] && sh.thi[a] == sh.thi[b]) { sh.talias[b] = a; sh.tact[b] = 2; }
            }
            __syncthreads();
            // ---- collect pass: keys equal to the bin's lower splitter are counted, keys strictly inside listed
            {
                unsigned tlo[IPB_SW_TARGETS], thi[IPB_SW_TARGETS], eq[IPB_SW_TARGETS];
                bool act[IPB_SW_TARGETS];
#pragma unroll
                for (int t = 0; t < IPB_SW_TARGETS; ++t) { act[t] = sh.tact[t] == 1; tlo[t] = sh.tlo[t]; thi[t] = sh.thi[t]; eq[t] = 0; }
                for (unsigned idx = tid; idx < area; idx += IPB_SW_THREADS) {
                    const unsigned key = keys[idx];
                    if (key == IPB_RS_SENTINEL) continue;
#pragma unroll
                    for (int t = 0; t < IPB_SW_TARGETS; ++t) {
                        if (!act[t] || key < tlo[t] || key >= thi[t]) continue;
                        if (key == tlo[t]) ++eq[t];
                        else { const unsigned p = atomicAdd(&sh.tn[t], 1u); if (p < IPB_SW_LIST) sh.tlist[t][p] = key; }
                    }
                }
#pragma unroll
                for (int t = 0; t < IPB_SW_TARGETS; ++t) {
                    const unsigned e = ipb_warp_sum(eq[t]);
                    if (lane == 0 && e) atomicAdd(&sh.teq[t], e);
                }
            }
            __syncthreads();
            // ---- resolve, target after target with the whole CTA: the (rr - eq)-th smallest listed key by
            //      counting (lists are one sample interval long: tens of keys, rarely hundreds)
            for (int t = 0; t < IPB_SW_TARGETS; ++t) {
                if (!sh.tact[t]) continue;                        // block-uniform
                const int src = sh.tact[t] == 2 ? sh.talias[t] : t;
                const unsigned rr = sh.trr[t], eqc = sh.teq[src], L = sh.tn[src];
                if (rr < eqc) { if (tid == 0) sh.tval[t] = sh.tlo[src]; }
                else if (L > IPB_SW_LIST || rr - eqc >= L) { if (tid == 0) sh.missed = 1; }
                else {
                    const unsigned m = rr - eqc;
                    for (unsigned j = tid; j < L; j += IPB_SW_THREADS) {
                        const unsigned e = sh.tlist[src][j];
                        unsigned less = 0, same = 0;
                        for (unsigned k = 0; k < L; ++k) { const unsigned x = sh.tlist[src][k]; less += x < e; same += x == e; }
                        if (m >= less && m < less + same) sh.tval[t] = e;
                    }
                }
            }
        }
        if (SRC != IPB_SRC_U16) {
            ssd_t = ipb_warp_sum(ssd_t);
            if (lane == 0) sh.dred[warp][0] = ssd_t;
        }
        __syncthreads();
        if (SRC != IPB_SRC_U16) { double t = 0.0; for (int i = 0; i < nwarps; ++i) t += sh.dred[i][0]; v_ssd[0] = t; }
        if (sh.missed) { if (tid == 0) atomicAdd(miss, 1u); continue; }        // block-uniform; the step is repeated

        if (tid < nv) {
            const int v = tid;
            const float B = vB[v]; const int clip = vclip[v];
            IpbStatOut o;
            o.n = n; o.area = area; o.sum = v_sum[v]; o.ssd = v_ssd[v] > 0.0 ? v_ssd[v] : 0.0; o.pad0 = 0.f;
            float rv[IPB_SW_TARGETS];
            for (int r = 0; r < IPB_SW_TARGETS; ++r) {
                const unsigned key = sh.tval[r];
                rv[r] = (SRC == IPB_SRC_U16) ? ipb_rs_transform(B, clip, key) : ipb_key_f32(key);
            }
            o.vmin = (SRC == IPB_SRC_U16) ? ipb_rs_transform(B, clip, kmin) : ipb_key_f32(kmin);
            o.vmax = (SRC == IPB_SRC_U16) ? ipb_rs_transform(B, clip, kmax) : ipb_key_f32(kmax);
            for (int i = 0; i < IPB_RS_MAXQ; ++i) {
                o.q[i] = fnan;
                if (job.qkind[i] == IPB_QKIND_PCT) o.q[i] = ipb_np_lerp_f32(rv[2 * i], rv[2 * i + 1], qi[i].gamma);
                else if (job.qkind[i] == IPB_QKIND_MEDIAN)
                    o.q[i] = (n & 1ull) ? rv[2 * i] : ipb_np_mid2_f32(rv[2 * i], rv[2 * i + 1]);
            }
            out[job.out[v]] = o;
        }
    }
}
