// ipb_roistats.cuh -- per-region statistics with EXACT order statistics
// (quantify_stats, quantify_per_roi, pick_epsilon on a
// masked denominator, annulus medians).
//
// A "region" is a bit-packed mask over a rect of a frame (an ROI mask from the rasteriser,
// or a whole-frame bitmask such as the ROI union) optionally ANDed with a second
// whole-frame bitmask (Nesprin2: roi & rim).  A job measures one value source over one
// region:
//   IPB_SRC_U16  v = raw uint16 pixel, value = T(v) = float32(v) - B, optionally clipped at 0
//                (the reference's bg_correct).  T is monotone, so order statistics are
//                selected on the integer key v and transformed afterwards: exact.  Because
//                the key set does not depend on B, one job can carry several VIEWS
//                (B, clip) of the same pixels -- e.g. the FRET builder's and Fluor_INT's
//                background levels of one channel -- and all views share the gather and the
//                selection; each view gets its own output row.
//   IPB_SRC_F32  value = a float32 image pixel (ratio image); non-finite values are
//                dropped like the reference's np.isfinite filter; key = ordered-uint32.
//
// One CTA (1024 threads, one per SM) per job:
//   gather   warps walk the region's mask words, 4 words per iteration so that 4 coalesced
//            pixel loads per lane are in flight (lane b <-> bit b); valid values are
//            compacted as keys into shared memory (ballot/popc positions) while n and the
//            key range are accumulated.  Regions that do not fit in shared memory keep the
//            same code path but re-walk global memory in every pass.
//   select   pass 1 is ONE wide histogram over the top bits of (key - kmin) using all shared
//            memory the key store left free (up to 32768 bins): for uint16 data this resolves
//            every rank at once AND gives the exact per-view sums (sum of count*T(bin)), so
//            the uint16 path does no per-pixel floating point at all.  The histogram is
//            scanned warp-cooperatively (lanes read consecutive bins: no bank conflicts).
//            If low bits remain (float keys), the handful of keys that share a wanted prefix
//            are compacted into a small list which is ranked by counting (<= 512 entries) or
//            a bitonic sort; pathological tie-heavy regions fall back to 8-bit digit passes
//            with one histogram per distinct prefix.
// numpy's float32 percentile / median arithmetic is replayed from ipb_exact.cuh.
#pragma once
#include "ipb_rt.cuh"
#include "ipb_exact.cuh"

#define IPB_SRC_U16 0
#define IPB_SRC_F32 1
#define IPB_RS_THREADS 1024
#define IPB_RS_MAXQ 3
#define IPB_RS_MAXR (2 * IPB_RS_MAXQ)
#define IPB_RS_MAXV 2
#define IPB_RS_DIGIT 8
#define IPB_RS_BINS (1 << IPB_RS_DIGIT)
#define IPB_RS_UNROLL 4
#define IPB_RS_SMEM_BYTES (212 * 1024)      // dynamic shared memory for the key store

#define IPB_QKIND_NONE 0
#define IPB_QKIND_PCT 1      // np.percentile(vals, p): q32 = f32(p)/f32(100)
#define IPB_QKIND_MEDIAN 2   // np.median(vals)

struct IpbRegion {            // one per region
    long long mask_off;       // word offset of the region's bit rows in `mask_pool`
    int x0, y0, w, h;         // rect in FRAME coordinates (bit b of word j of row r <-> x0+32j+b, y0+r)
    int wpr;                  // words per mask row
    int frame;                // frame index (informational)
    int use_and;              // != 0: AND with and_bits[and_plane]
    int and_plane;            // index of the H x and_wpr bit plane in and_bits
};

struct IpbStatJob {           // 64 bytes
    int region;
    int src;                  // IPB_SRC_*
    int plane;                // U16: plane index (frame*C+ch); F32: image index
    int n_views;              // U16: 1..IPB_RS_MAXV ; F32: 1
    int bidx[IPB_RS_MAXV];    // U16: index into bvals (background B); < 0 -> B = 0
    int clip_neg[IPB_RS_MAXV];// U16: clip T(v) at 0
    int qkind[IPB_RS_MAXQ];
    float q32[IPB_RS_MAXQ];
    int out[IPB_RS_MAXV];     // view v writes out[out[v]]
};

struct IpbStatOut {
    unsigned long long n;     // finite values measured (npx)
    unsigned long long area;  // pixels in the region (before the finite filter)
    double sum;               // sum of values (float64 accumulation)
    double ssd;               // sum of squared deviations from the float64 mean
    float vmin, vmax;
    float q[IPB_RS_MAXQ];     // requested order-statistic results (NaN if n == 0)
    float pad0;
};

struct IpbRsCtx {
    const unsigned* mask; const unsigned* androw0; int and_wpr;
    int x0, y0, w, h, wpr;
    const unsigned short* u16; const float* f32; int W;
};

__device__ __forceinline__ float ipb_rs_transform(float B, int clip, unsigned v) {
    float t = __fsub_rn((float)v, B);
    if (clip && t < 0.0f) t = 0.0f;
    return t;
}

// One warp-iteration of the region walk: up to IPB_RS_UNROLL consecutive mask words of row
// r starting at word j0; lane b owns bit b.  Fills for this lane key[u] and ok[u] (pixel
// belongs to the region and is measured) and counts region pixels.
template <int SRC>
__device__ __forceinline__ void ipb_rs_load_group(const IpbRsCtx& c, int r, int j0, int lane,
                                                  unsigned (&key)[IPB_RS_UNROLL], bool (&ok)[IPB_RS_UNROLL],
                                                  unsigned& region_px) {
    const unsigned* mrow = c.mask + (size_t)r * c.wpr;
    const int y = c.y0 + r;
    unsigned m[IPB_RS_UNROLL];
#pragma unroll
    for (int u = 0; u < IPB_RS_UNROLL; ++u) m[u] = (j0 + u < c.wpr) ? mrow[j0 + u] : 0u;
    if (c.androw0) {
        const unsigned* ar = c.androw0 + (size_t)y * c.and_wpr;
#pragma unroll
        for (int u = 0; u < IPB_RS_UNROLL; ++u) {
            if (m[u]) {
                const int xb = c.x0 + 32 * (j0 + u), k = xb >> 5, s = xb & 31;
                unsigned lo = ar[k] >> s;
                if (s && k + 1 < c.and_wpr) lo |= ar[k + 1] << (32 - s);
                m[u] &= lo;
            }
        }
    }
    const size_t p0 = (size_t)y * c.W + (c.x0 + 32 * j0 + lane);
    unsigned raw[IPB_RS_UNROLL];
#pragma unroll
    for (int u = 0; u < IPB_RS_UNROLL; ++u) {
        ok[u] = (m[u] >> lane) & 1u;
        raw[u] = 0u;
        if (ok[u]) raw[u] = (SRC == IPB_SRC_U16) ? (unsigned)c.u16[p0 + 32 * u] : __float_as_uint(c.f32[p0 + 32 * u]);
    }
#pragma unroll
    for (int u = 0; u < IPB_RS_UNROLL; ++u) {
        region_px += ok[u] ? 1u : 0u;
        if (SRC == IPB_SRC_U16) key[u] = raw[u];
        else {
            const float v = __uint_as_float(raw[u]);
            if (ok[u] && !isfinite(v)) ok[u] = false;
            key[u] = ipb_f32_key(v);
        }
    }
}

// f(key) for every measured value: from the shared-memory key store when it holds the whole
// region, else by re-walking global memory.
template <int SRC, typename F>
__device__ __forceinline__ void ipb_rs_foreach_key(const IpbRsCtx& c, bool in_smem, unsigned n,
                                                   const unsigned* k32, const unsigned short* k16, F f) {
    if (in_smem) {
        if (SRC == IPB_SRC_U16) for (unsigned i = threadIdx.x; i < n; i += blockDim.x) f((unsigned)k16[i]);
        else for (unsigned i = threadIdx.x; i < n; i += blockDim.x) f(k32[i]);
    } else {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int r = warp; r < c.h; r += nwarps)
            for (int j0 = 0; j0 < c.wpr; j0 += IPB_RS_UNROLL) {
                unsigned key[IPB_RS_UNROLL], dummy = 0;
                bool ok[IPB_RS_UNROLL];
                ipb_rs_load_group<SRC>(c, r, j0, lane, key, ok, dummy);
#pragma unroll
                for (int u = 0; u < IPB_RS_UNROLL; ++u) if (ok[u]) f(key[u]);
            }
    }
}

__device__ __forceinline__ double ipb_block_sum_d(double v, double* red) {
    v = ipb_warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ unsigned long long ipb_block_sum_u64(unsigned long long v, unsigned long long* red) {
    v = ipb_warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    unsigned long long t = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

#define IPB_RS_LISTCAP 4096
#define IPB_RS_COUNTSORT 512           // candidate lists up to this size are ranked by counting
#define IPB_RS_RESERVE (16 * 1024)     // bytes of the dynamic store always left to the histogram

template <int SRC>
__global__ void __launch_bounds__(IPB_RS_THREADS, 1)
ipb_k_region_stats(const IpbRegion* __restrict__ regions, const IpbStatJob* __restrict__ jobs,
                   const unsigned* __restrict__ mask_pool, const unsigned* __restrict__ and_bits,
                   int and_wpr, int H, int W,
                   const unsigned short* __restrict__ planes, const float* __restrict__ images,
                   const float* __restrict__ bvals, IpbStatOut* __restrict__ out, int smem_bytes)
{
    IPB_DYN_SMEM(unsigned, keystore);
    __shared__ unsigned ghist[IPB_RS_MAXR][IPB_RS_BINS];   // generic (fallback) per-group histograms
    __shared__ double red_d[32];
    __shared__ unsigned long long red_u[32];
    __shared__ unsigned red_k[2][32];
    __shared__ unsigned long long r_rank[IPB_RS_MAXR];   // remaining rank inside its group
    __shared__ unsigned r_prefix[IPB_RS_MAXR];           // key' bits resolved so far (>> shift)
    __shared__ int r_group[IPB_RS_MAXR];
    __shared__ unsigned g_prefix[IPB_RS_MAXR];
    __shared__ int g_n;
    __shared__ unsigned n_stored, list_n;

    const IpbStatJob job = jobs[blockIdx.x];
    if (job.src != SRC) return;                          // mixed job lists: the other instantiation takes it
    const IpbRegion rg = regions[job.region];
    IpbRsCtx c;
    c.mask = mask_pool + rg.mask_off;
    c.androw0 = (rg.use_and && and_bits) ? and_bits + (size_t)rg.and_plane * H * and_wpr : nullptr;
    c.and_wpr = and_wpr;
    c.x0 = rg.x0; c.y0 = rg.y0; c.w = rg.w; c.h = rg.h; c.wpr = rg.wpr;
    c.W = W;
    c.u16 = planes ? planes + (size_t)job.plane * H * W : nullptr;
    c.f32 = images ? images + (size_t)job.plane * H * W : nullptr;
    const int nv = (SRC == IPB_SRC_U16) ? (job.n_views < 1 ? 1 : (job.n_views > IPB_RS_MAXV ? IPB_RS_MAXV : job.n_views)) : 1;
    float vB[IPB_RS_MAXV]; int vclip[IPB_RS_MAXV];
#pragma unroll
    for (int v = 0; v < IPB_RS_MAXV; ++v) {
        vB[v] = (SRC == IPB_SRC_U16 && v < nv && job.bidx[v] >= 0) ? bvals[job.bidx[v]] : 0.0f;
        vclip[v] = (v < nv) ? job.clip_neg[v] : 0;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    unsigned* k32 = keystore;
    unsigned short* k16 = reinterpret_cast<unsigned short*>(keystore);
    const int keysize = SRC == IPB_SRC_U16 ? 2 : 4;
    const unsigned cap = (unsigned)((smem_bytes - IPB_RS_RESERVE) / keysize);
    if (tid == 0) { n_stored = 0u; list_n = 0u; }
    __syncthreads();

    // ---- gather: n, area, key range (F32: also sum); keys compacted into shared memory
    unsigned n_t = 0, area_t = 0;
    double s_t = 0.0, s2_t = 0.0;
    unsigned kmin_t = 0xffffffffu, kmax_t = 0u;
    for (int r = warp; r < c.h; r += nwarps) {
        for (int j0 = 0; j0 < c.wpr; j0 += IPB_RS_UNROLL) {
            unsigned key[IPB_RS_UNROLL];
            bool ok[IPB_RS_UNROLL];
            ipb_rs_load_group<SRC>(c, r, j0, lane, key, ok, area_t);
            unsigned vm[IPB_RS_UNROLL], tot = 0;
#pragma unroll
            for (int u = 0; u < IPB_RS_UNROLL; ++u) { vm[u] = __ballot_sync(IPB_FULL, ok[u]); tot += __popc(vm[u]); }
            if (tot == 0) continue;                                  // warp-uniform
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&n_stored, tot);
            base = __shfl_sync(IPB_FULL, base, 0);
#pragma unroll
            for (int u = 0; u < IPB_RS_UNROLL; ++u) {
                if (ok[u]) {
                    const unsigned pos = base + __popc(vm[u] & ((1u << lane) - 1u));
                    if (pos < cap) { if (SRC == IPB_SRC_U16) k16[pos] = (unsigned short)key[u]; else k32[pos] = key[u]; }
                    ++n_t;
                    if (SRC == IPB_SRC_F32) { const double v = (double)ipb_key_f32(key[u]); s_t += v; s2_t += v * v; }
                    kmin_t = key[u] < kmin_t ? key[u] : kmin_t;
                    kmax_t = key[u] > kmax_t ? key[u] : kmax_t;
                }
                base += __popc(vm[u]);
            }
        }
    }
    const unsigned long long n = ipb_block_sum_u64((unsigned long long)n_t, red_u);
    const unsigned long long area = ipb_block_sum_u64((unsigned long long)area_t, red_u);
    kmin_t = ipb_warp_min(kmin_t); kmax_t = ipb_warp_max(kmax_t);
    __syncthreads();
    if (lane == 0) { red_k[0][warp] = kmin_t; red_k[1][warp] = kmax_t; }
    __syncthreads();
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = 0; i < nwarps; ++i) { kmin = red_k[0][i] < kmin ? red_k[0][i] : kmin; kmax = red_k[1][i] > kmax ? red_k[1][i] : kmax; }
    const bool in_smem = n <= (unsigned long long)cap;

    const float fnan = __uint_as_float(0x7fc00000u);
    if (n == 0) {
        if (tid < nv) {
            IpbStatOut o;
            o.n = 0; o.area = area; o.sum = 0.0; o.ssd = 0.0; o.pad0 = 0.f;
            o.vmin = fnan; o.vmax = fnan;
            for (int i = 0; i < IPB_RS_MAXQ; ++i) o.q[i] = fnan;
            out[job.out[tid]] = o;
        }
        return;
    }

    // ---- ranks wanted
    IpbQIdx qi[IPB_RS_MAXQ];
    int nr = 0;
    for (int i = 0; i < IPB_RS_MAXQ; ++i) {
        qi[i].prev = qi[i].next = 0; qi[i].gamma = 0.f;
        if (job.qkind[i] == IPB_QKIND_PCT) qi[i] = ipb_np_qidx_f32((long long)n, job.q32[i]);
        else if (job.qkind[i] == IPB_QKIND_MEDIAN) {
            if (n & 1ull) qi[i].prev = qi[i].next = (long long)(n >> 1);
            else { qi[i].prev = (long long)(n >> 1) - 1; qi[i].next = (long long)(n >> 1); }
        }
        if (job.qkind[i] != IPB_QKIND_NONE) nr = 2 * (i + 1);
    }
    if (tid == 0) {
        for (int i = 0; i < IPB_RS_MAXQ; ++i) {
            r_rank[2 * i] = (unsigned long long)qi[i].prev; r_rank[2 * i + 1] = (unsigned long long)qi[i].next;
            r_prefix[2 * i] = r_prefix[2 * i + 1] = 0u;
            r_group[2 * i] = r_group[2 * i + 1] = 0;
        }
        g_prefix[0] = 0u; g_n = 1;
    }
    __syncthreads();

    // ---- pass 1: one wide histogram (as many bins as the free part of the store allows)
    const unsigned nn = (unsigned)(in_smem ? n : 0);
    const unsigned range = kmax - kmin;
    const int bits = range ? (32 - __clz((int)range)) : 0;
    const unsigned key_words = in_smem ? (unsigned)(((size_t)n * keysize + 3) / 4) : 0u;
    unsigned* whist = keystore + key_words;
    const unsigned hb_cap = (unsigned)(smem_bytes / 4) - key_words;        // >= RESERVE/4 = 4096
    int d1 = 31 - __clz((int)hb_cap);
    if (d1 > 15) d1 = 15;
    if (d1 > bits) d1 = bits;
    const int rb = bits - d1;                                               // bits left after pass 1
    const unsigned nb = 1u << d1;
    for (unsigned i = tid; i < nb; i += blockDim.x) whist[i] = 0u;
    __syncthreads();
    ipb_rs_foreach_key<SRC>(c, in_smem, nn, k32, k16, [&](unsigned key) { atomicAdd(&whist[(key - kmin) >> rb], 1u); });
    __syncthreads();
    {
        // warp-cooperative scan: warp w owns a contiguous band of 32-bin rows; lanes read
        // consecutive bins (conflict-free).  Only the bands that contain a wanted rank are
        // scanned row by row.
        unsigned long long kk_l[IPB_RS_MAXR];
        for (int r = 0; r < IPB_RS_MAXR; ++r) kk_l[r] = r_rank[r];
        const unsigned rows = (nb + 31u) >> 5;
        const unsigned rpw = (rows + (unsigned)nwarps - 1u) / (unsigned)nwarps;
        const unsigned row0 = (unsigned)warp * rpw;
        unsigned row1 = row0 + rpw;
        if (row1 > rows) row1 = rows;
        unsigned long long band = 0;
        for (unsigned row = row0; row < row1; ++row) {
            const unsigned b = (row << 5) + (unsigned)lane;
            if (b < nb) band += whist[b];
        }
        band = ipb_warp_sum(band);
        __syncthreads();
        if (lane == 0) red_u[warp] = band;
        __syncthreads();
        unsigned long long base = 0;
        for (int i = 0; i < warp; ++i) base += red_u[i];
        bool mine = false;
        for (int r = 0; r < nr; ++r) mine = mine || (kk_l[r] >= base && kk_l[r] < base + band);
        if (mine) {                                                        // warp-uniform
            unsigned long long run = base;
            for (unsigned row = row0; row < row1; ++row) {
                const unsigned b = (row << 5) + (unsigned)lane;
                const unsigned v = b < nb ? whist[b] : 0u;
                unsigned incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
                const unsigned rowtot = __shfl_sync(IPB_FULL, incl, 31);
                for (int r = 0; r < nr; ++r) {
                    const unsigned long long kk = kk_l[r];
                    if (kk >= run && kk < run + rowtot) {
                        const unsigned off = (unsigned)(kk - run);
                        if (off >= incl - v && off < incl) { r_prefix[r] = b; r_rank[r] = (unsigned long long)(off - (incl - v)); }
                    }
                }
                run += rowtot;
            }
        }
        __syncthreads();
    }

    // ---- sums.  U16: per view from the exact histogram (or the keys when bins are coarse);
    //      F32: sum from the gather, squared deviations from the stored keys.
    double v_sum[IPB_RS_MAXV], v_ssd[IPB_RS_MAXV];
    if (SRC == IPB_SRC_U16) {
        for (int v = 0; v < nv; ++v) {
            const float B = vB[v]; const int clip = vclip[v];
            double s = 0.0;
            if (rb == 0) { for (unsigned b = tid; b < nb; b += blockDim.x) { const unsigned cnt = whist[b]; if (cnt) s += (double)cnt * (double)ipb_rs_transform(B, clip, kmin + b); } }
            else ipb_rs_foreach_key<SRC>(c, in_smem, nn, k32, k16, [&](unsigned key) { s += (double)ipb_rs_transform(B, clip, key); });
            const double sum = ipb_block_sum_d(s, red_d);
            const double mean = sum / (double)n;
            double q = 0.0;
            if (rb == 0) { for (unsigned b = tid; b < nb; b += blockDim.x) { const unsigned cnt = whist[b]; if (cnt) { const double d = (double)ipb_rs_transform(B, clip, kmin + b) - mean; q += (double)cnt * d * d; } } }
            else ipb_rs_foreach_key<SRC>(c, in_smem, nn, k32, k16, [&](unsigned key) { const double d = (double)ipb_rs_transform(B, clip, key) - mean; q += d * d; });
            v_sum[v] = sum;
            v_ssd[v] = ipb_block_sum_d(q, red_d);
        }
    } else {
        const double sum = ipb_block_sum_d(s_t, red_d);
        double ssd;
        if (in_smem) {
            const double mean = sum / (double)n;
            double q = 0.0;
            for (unsigned i = tid; i < nn; i += blockDim.x) { const double d = (double)ipb_key_f32(k32[i]) - mean; q += d * d; }
            ssd = ipb_block_sum_d(q, red_d);
        } else {
            const double sumsq = ipb_block_sum_d(s2_t, red_d);
            ssd = sumsq - sum * (sum / (double)n);
            if (!(ssd > 0.0)) ssd = 0.0;
        }
        v_sum[0] = sum; v_ssd[0] = ssd;
    }

    if (rb > 0 && nr > 0) {
        // ---- regroup, then compact the few keys that share a wanted prefix
        if (tid == 0) {
            int m = 0;
            for (int r = 0; r < nr; ++r) {
                int gg = -1;
                for (int t = 0; t < m; ++t) if (g_prefix[t] == r_prefix[r]) { gg = t; break; }
                if (gg < 0) { g_prefix[m] = r_prefix[r]; gg = m++; }
                r_group[r] = gg;
            }
            g_n = m > 0 ? m : 1;
        }
        __syncthreads();
        const int ng = g_n;
        unsigned gp[IPB_RS_MAXR];
        for (int g = 0; g < IPB_RS_MAXR; ++g) gp[g] = g < ng ? g_prefix[g] : 0xffffffffu;
        unsigned* list = whist;                                  // the wide histogram is no longer needed
        const unsigned lowmask = (rb >= 32) ? 0xffffffffu : ((1u << rb) - 1u);
        const bool packable = rb <= 28;
        __syncthreads();
        if (packable) {
            ipb_rs_foreach_key<SRC>(c, in_smem, nn, k32, k16, [&](unsigned key) {
                const unsigned kp = key - kmin, hi = kp >> rb;
#pragma unroll
                for (int g = 0; g < IPB_RS_MAXR; ++g) {
                    if (hi == gp[g]) {
                        const unsigned idx = atomicAdd(&list_n, 1u);
                        if (idx < IPB_RS_LISTCAP) list[idx] = ((unsigned)g << 28) | (kp & lowmask);
                    }
                }
            });
        }
        __syncthreads();
        const unsigned m = list_n;
        if (packable && m <= IPB_RS_LISTCAP) {
            // order the candidate list (group in the top 4 bits, low key bits below); rank r is
            // then a direct index into its group's segment
            unsigned* sorted = list;
            if (m <= IPB_RS_COUNTSORT) {
                // small list: every thread places one element by counting (stable on index)
                sorted = list + IPB_RS_LISTCAP / 2;
                if ((unsigned)tid < m) {
                    const unsigned a = list[tid];
                    unsigned pos = 0;
                    for (unsigned j = 0; j < m; ++j) { const unsigned b = list[j]; pos += (b < a || (b == a && j < (unsigned)tid)) ? 1u : 0u; }
                    sorted[pos] = a;
                }
                __syncthreads();
            } else {
                unsigned P = 1;
                while (P < m) P <<= 1;
                for (unsigned i = m + tid; i < P; i += blockDim.x) list[i] = 0xffffffffu;
                __syncthreads();
                for (unsigned k = 2; k <= P; k <<= 1) {
                    for (unsigned j = k >> 1; j > 0; j >>= 1) {
                        for (unsigned i = tid; i < P; i += blockDim.x) {
                            const unsigned ixj = i ^ j;
                            if (ixj > i) {
                                const unsigned a = list[i], b = list[ixj];
                                const bool up = (i & k) == 0;
                                if ((a > b) == up) { list[i] = b; list[ixj] = a; }
                            }
                        }
                        __syncthreads();
                    }
                }
            }
            if (tid < nr) {
                const int r = tid;
                const unsigned g = (unsigned)r_group[r];
                // first index of group g: binary search for (g << 28)
                unsigned lo = 0, hi = m;
                while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (sorted[mid] < (g << 28)) lo = mid + 1; else hi = mid; }
                const unsigned ve = sorted[lo + (unsigned)r_rank[r]];
                r_prefix[r] = (gp[g] << rb) | (ve & lowmask);                    // full key'
            }
            __syncthreads();
        } else {
            // ---- generic fallback: 8-bit digit passes with one histogram per distinct prefix
            int shift_prev = rb;
            do {
                const int shift = shift_prev > IPB_RS_DIGIT ? shift_prev - IPB_RS_DIGIT : 0;
                const unsigned dmask = (1u << (shift_prev - shift)) - 1u;
                const int ngc = g_n;
                for (int i = tid; i < ngc * IPB_RS_BINS; i += blockDim.x) (&ghist[0][0])[i] = 0u;
                __syncthreads();
                ipb_rs_foreach_key<SRC>(c, in_smem, nn, k32, k16, [&](unsigned key) {
                    const unsigned kp = key - kmin;
                    const unsigned hi = shift_prev >= 32 ? 0u : (kp >> shift_prev);
                    const unsigned dg = (kp >> shift) & dmask;
                    for (int g = 0; g < ngc; ++g)
                        if (hi == g_prefix[g]) atomicAdd(&ghist[g][dg], 1u);
                });
                __syncthreads();
                if (warp < ngc) {
                    const int g = warp;
                    unsigned cnt[IPB_RS_BINS / 32], mine = 0;
#pragma unroll
                    for (int b = 0; b < IPB_RS_BINS / 32; ++b) { cnt[b] = ghist[g][lane * (IPB_RS_BINS / 32) + b]; mine += cnt[b]; }
                    unsigned long long incl = mine;
#pragma unroll
                    for (int o2 = 1; o2 < 32; o2 <<= 1) { unsigned long long t = __shfl_up_sync(IPB_FULL, incl, o2); if (lane >= o2) incl += t; }
                    const unsigned long long lo = incl - mine, hi = incl;
                    for (int r = 0; r < nr; ++r) {
                        if (r_group[r] != g) continue;
                        const unsigned long long kk = r_rank[r];
                        __syncwarp();
                        if (kk >= lo && kk < hi) {
                            unsigned long long acc = lo;
#pragma unroll
                            for (int b = 0; b < IPB_RS_BINS / 32; ++b) {
                                if (kk >= acc && kk < acc + cnt[b]) {
                                    r_prefix[r] = (g_prefix[g] << (shift_prev - shift)) | (unsigned)(lane * (IPB_RS_BINS / 32) + b);
                                    r_rank[r] = kk - acc;
                                }
                                acc += cnt[b];
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    int mm = 0;
                    for (int r = 0; r < nr; ++r) {
                        int gg = -1;
                        for (int t = 0; t < mm; ++t) if (g_prefix[t] == r_prefix[r]) { gg = t; break; }
                        if (gg < 0) { g_prefix[mm] = r_prefix[r]; gg = mm++; }
                        r_group[r] = gg;
                    }
                    g_n = mm > 0 ? mm : 1;
                }
                __syncthreads();
                shift_prev = shift;
            } while (shift_prev > 0);
        }
    }

    if (tid < nv) {
        const int v = tid;
        const float B = vB[v]; const int clip = vclip[v];
        IpbStatOut o;
        o.n = n; o.area = area; o.sum = v_sum[v]; o.ssd = v_ssd[v] > 0.0 ? v_ssd[v] : 0.0; o.pad0 = 0.f;
        float rv[IPB_RS_MAXR];
        for (int r = 0; r < IPB_RS_MAXR; ++r) {
            const unsigned key = kmin + r_prefix[r];
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
