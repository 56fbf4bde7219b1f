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
//                selected on the integer key v and transformed afterwards: exact.
//   IPB_SRC_F32  value = a float32 image pixel (ratio image); non-finite values are
//                dropped like the reference's np.isfinite filter; key = ordered-uint32.
//
// One CTA (1024 threads, one per SM) per job:
//   gather   warps walk the region's mask words, 4 words per iteration so that 4 coalesced
//            pixel loads per lane are in flight (lane b <-> bit b); valid values are
//            compacted as keys into shared memory (ballot/popc positions), while n, sum and
//            the key range are accumulated.  Regions that do not fit in shared memory keep
//            the same code path but re-walk global memory in every pass.
//   select   radix select with 8-bit digits over (key - kmin): only as many passes as the
//            region's own key range needs, all requested ranks (p5 / median / p95 -> up to 6)
//            resolved together with one 256-bin histogram per distinct prefix.  The first
//            pass also accumulates the squared deviations from the float64 mean.
// numpy's float32 percentile / median arithmetic is replayed from ipb_exact.cuh.
#pragma once
#include "ipb_rt.cuh"
#include "ipb_exact.cuh"

#define IPB_SRC_U16 0
#define IPB_SRC_F32 1
#define IPB_RS_THREADS 1024
#define IPB_RS_MAXQ 3
#define IPB_RS_MAXR (2 * IPB_RS_MAXQ)
#define IPB_RS_DIGIT 8
#define IPB_RS_BINS (1 << IPB_RS_DIGIT)
#define IPB_RS_UNROLL 4
#define IPB_RS_SMEM_BYTES (200 * 1024)      // dynamic shared memory for the key store

#define IPB_QKIND_NONE 0
#define IPB_QKIND_PCT 1      // np.percentile(vals, p): q32 = f32(p)/f32(100)
#define IPB_QKIND_MEDIAN 2   // np.median(vals)

struct IpbRegion {            // one per region
    long long mask_off;       // word offset of the region's bit rows in `mask_pool`
    int x0, y0, w, h;         // rect in FRAME coordinates (bit b of word j of row r <-> x0+32j+b, y0+r)
    int wpr;                  // words per mask row
    int frame;                // frame index (selects image plane and AND-mask)
    int use_and;              // != 0: AND with and_bits[frame]
    int pad0;
};

struct IpbStatJob {
    int region;
    int src;                  // IPB_SRC_*
    int plane;                // U16: plane index (frame*C+ch); F32: image index
    int bidx;                 // U16: index into bvals (background B); < 0 -> B = 0
    int clip_neg;             // U16: clip T(v) at 0
    int qkind[IPB_RS_MAXQ];
    float q32[IPB_RS_MAXQ];
    int pad0;
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
    float B; int clip; int src;
};

__device__ __forceinline__ float ipb_rs_transform(const IpbRsCtx& c, unsigned v) {
    float t = __fsub_rn((float)v, c.B);
    if (c.clip && t < 0.0f) t = 0.0f;
    return t;
}
__device__ __forceinline__ float ipb_rs_value(const IpbRsCtx& c, unsigned key) {
    return c.src == IPB_SRC_U16 ? ipb_rs_transform(c, key) : ipb_key_f32(key);
}

// One warp-iteration of the region walk: IPB_RS_UNROLL consecutive mask words starting at
// word w0; lane b owns bit b.  Fills m[u] (warp-uniform word after the AND mask), and for
// this lane key[u] and ok[u] (pixel belongs to the region and is measured).
__device__ __forceinline__ void ipb_rs_load_group(const IpbRsCtx& c, long long w0, long long nwords, int lane,
                                                  unsigned (&m)[IPB_RS_UNROLL], unsigned (&key)[IPB_RS_UNROLL],
                                                  bool (&ok)[IPB_RS_UNROLL], unsigned& region_px) {
    int yy[IPB_RS_UNROLL], xx[IPB_RS_UNROLL];
#pragma unroll
    for (int u = 0; u < IPB_RS_UNROLL; ++u) {
        const long long wi = w0 + u;
        m[u] = 0u;
        yy[u] = 0; xx[u] = 0;
        if (wi < nwords) {
            const int r = (int)(wi / c.wpr), j = (int)(wi - (long long)r * c.wpr);
            unsigned mw = c.mask[wi];
            const int y = c.y0 + r, xb = c.x0 + 32 * j;
            if (mw && c.androw0) {
                const unsigned* ar = c.androw0 + (size_t)y * c.and_wpr;
                const int k = xb >> 5, s = xb & 31;
                unsigned lo = ar[k] >> s;
                if (s && k + 1 < c.and_wpr) lo |= ar[k + 1] << (32 - s);
                mw &= lo;
            }
            m[u] = mw;
            yy[u] = y; xx[u] = xb + lane;
        }
    }
    unsigned raw[IPB_RS_UNROLL];
#pragma unroll
    for (int u = 0; u < IPB_RS_UNROLL; ++u) {
        ok[u] = (m[u] >> lane) & 1u;
        raw[u] = 0u;
        if (ok[u]) {
            const size_t p = (size_t)yy[u] * c.W + xx[u];
            raw[u] = (c.src == IPB_SRC_U16) ? (unsigned)c.u16[p] : __float_as_uint(c.f32[p]);
        }
    }
#pragma unroll
    for (int u = 0; u < IPB_RS_UNROLL; ++u) {
        region_px += ok[u] ? 1u : 0u;
        if (c.src == IPB_SRC_U16) key[u] = raw[u];
        else {
            const float v = __uint_as_float(raw[u]);
            if (ok[u] && !isfinite(v)) ok[u] = false;
            key[u] = ipb_f32_key(v);
        }
    }
}

// f(key) for every measured value: from the shared-memory key store when it holds the whole
// region, else by re-walking global memory.
template <typename F>
__device__ __forceinline__ void ipb_rs_foreach_key(const IpbRsCtx& c, bool in_smem, unsigned long long n,
                                                   const unsigned* k32, const unsigned short* k16, F f) {
    if (in_smem) {
        if (c.src == IPB_SRC_U16) for (unsigned i = threadIdx.x; i < n; i += blockDim.x) f((unsigned)k16[i]);
        else for (unsigned i = threadIdx.x; i < n; i += blockDim.x) f(k32[i]);
    } else {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        const long long nwords = (long long)c.h * c.wpr;
        for (long long w0 = (long long)warp * IPB_RS_UNROLL; w0 < nwords; w0 += (long long)nwarps * IPB_RS_UNROLL) {
            unsigned m[IPB_RS_UNROLL], key[IPB_RS_UNROLL], dummy = 0;
            bool ok[IPB_RS_UNROLL];
            ipb_rs_load_group(c, w0, nwords, lane, m, key, ok, dummy);
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

__global__ void __launch_bounds__(IPB_RS_THREADS, 1)
ipb_k_region_stats(const IpbRegion* __restrict__ regions, const IpbStatJob* __restrict__ jobs,
                   const unsigned* __restrict__ mask_pool, const unsigned* __restrict__ and_bits,
                   int and_wpr, int H, int W,
                   const unsigned short* __restrict__ planes, const float* __restrict__ images,
                   const float* __restrict__ bvals, IpbStatOut* __restrict__ out, int smem_bytes)
{
    IPB_DYN_SMEM(unsigned, keystore);
    __shared__ unsigned hist[IPB_RS_MAXR][IPB_RS_BINS];
    __shared__ double red_d[32];
    __shared__ unsigned long long red_u[32];
    __shared__ unsigned red_k[2][32];
    __shared__ unsigned long long r_rank[IPB_RS_MAXR];   // remaining rank inside its group
    __shared__ unsigned r_prefix[IPB_RS_MAXR];           // key' bits resolved so far (>> shift)
    __shared__ int r_group[IPB_RS_MAXR];
    __shared__ unsigned g_prefix[IPB_RS_MAXR];
    __shared__ int g_n;
    __shared__ unsigned n_stored;

    const IpbStatJob job = jobs[blockIdx.x];
    const IpbRegion rg = regions[job.region];
    IpbRsCtx c;
    c.mask = mask_pool + rg.mask_off;
    c.androw0 = (rg.use_and && and_bits) ? and_bits + (size_t)rg.frame * H * and_wpr : nullptr;
    c.and_wpr = and_wpr;
    c.x0 = rg.x0; c.y0 = rg.y0; c.w = rg.w; c.h = rg.h; c.wpr = rg.wpr;
    c.W = W; c.src = job.src;
    c.u16 = planes ? planes + (size_t)job.plane * H * W : nullptr;
    c.f32 = images ? images + (size_t)job.plane * H * W : nullptr;
    c.B = (job.src == IPB_SRC_U16 && job.bidx >= 0) ? bvals[job.bidx] : 0.0f;
    c.clip = job.clip_neg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    unsigned* k32 = keystore;
    unsigned short* k16 = reinterpret_cast<unsigned short*>(keystore);
    const unsigned cap = (unsigned)(smem_bytes / (job.src == IPB_SRC_U16 ? 2 : 4));
    if (tid == 0) n_stored = 0u;
    __syncthreads();

    // ---- gather: n, area, sum, key range; keys compacted into shared memory while they fit
    unsigned long long n_t = 0;
    unsigned area_t = 0;
    double s_t = 0.0;
    unsigned kmin_t = 0xffffffffu, kmax_t = 0u;
    {
        const long long nwords = (long long)c.h * c.wpr;
        for (long long w0 = (long long)warp * IPB_RS_UNROLL; w0 < nwords; w0 += (long long)nwarps * IPB_RS_UNROLL) {
            unsigned m[IPB_RS_UNROLL], key[IPB_RS_UNROLL];
            bool ok[IPB_RS_UNROLL];
            ipb_rs_load_group(c, w0, nwords, lane, m, key, ok, area_t);
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
                    if (pos < cap) { if (job.src == IPB_SRC_U16) k16[pos] = (unsigned short)key[u]; else k32[pos] = key[u]; }
                    ++n_t;
                    s_t += (double)ipb_rs_value(c, key[u]);
                    kmin_t = key[u] < kmin_t ? key[u] : kmin_t;
                    kmax_t = key[u] > kmax_t ? key[u] : kmax_t;
                }
                base += __popc(vm[u]);
            }
        }
    }
    const unsigned long long n = ipb_block_sum_u64(n_t, red_u);
    const unsigned long long area = ipb_block_sum_u64((unsigned long long)area_t, red_u);
    const double sum = ipb_block_sum_d(s_t, red_d);
    kmin_t = ipb_warp_min(kmin_t); kmax_t = ipb_warp_max(kmax_t);
    __syncthreads();
    if (lane == 0) { red_k[0][warp] = kmin_t; red_k[1][warp] = kmax_t; }
    __syncthreads();
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = 0; i < nwarps; ++i) { kmin = red_k[0][i] < kmin ? red_k[0][i] : kmin; kmax = red_k[1][i] > kmax ? red_k[1][i] : kmax; }
    const bool in_smem = n <= (unsigned long long)cap;

    IpbStatOut o;
    o.n = n; o.area = area; o.sum = sum; o.ssd = 0.0; o.pad0 = 0.f;
    const float fnan = __uint_as_float(0x7fc00000u);
    o.vmin = fnan; o.vmax = fnan;
    for (int i = 0; i < IPB_RS_MAXQ; ++i) o.q[i] = fnan;
    if (n == 0) {
        if (tid == 0) out[blockIdx.x] = o;
        return;
    }
    const double mean = sum / (double)n;

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

    // ---- radix select over key' = key - kmin
    const unsigned range = kmax - kmin;
    const int bits = range ? (32 - __clz((int)range)) : 0;
    int shift_prev = bits;                   // bits below this are unresolved
    double ssd_t = 0.0;
    bool first = true;
    do {
        const int shift = shift_prev > IPB_RS_DIGIT ? shift_prev - IPB_RS_DIGIT : 0;
        const unsigned dmask = (1u << (shift_prev - shift)) - 1u;
        const int ng = g_n;
        for (int i = tid; i < ng * IPB_RS_BINS; i += blockDim.x) (&hist[0][0])[i] = 0u;
        __syncthreads();
        ipb_rs_foreach_key(c, in_smem, n, k32, k16, [&](unsigned key) {
            const unsigned kp = key - kmin;
            if (first) { const double d = (double)ipb_rs_value(c, key) - mean; ssd_t += d * d; }
            const unsigned hi = shift_prev >= 32 ? 0u : (kp >> shift_prev);
            const unsigned dg = (kp >> shift) & dmask;
            for (int g = 0; g < ng; ++g)
                if (first || hi == g_prefix[g]) atomicAdd(&hist[g][dg], 1u);
        });
        __syncthreads();
        // locate every rank inside its group's histogram: warp g scans group g (256 bins)
        if (warp < ng) {
            const int g = warp;
            unsigned cnt[IPB_RS_BINS / 32], mine = 0;
#pragma unroll
            for (int b = 0; b < IPB_RS_BINS / 32; ++b) { cnt[b] = hist[g][lane * (IPB_RS_BINS / 32) + b]; mine += cnt[b]; }
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
        if (tid == 0) {                                            // regroup by distinct prefix
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
        shift_prev = shift;
        first = false;
    } while (shift_prev > 0);

    const double ssd = ipb_block_sum_d(ssd_t, red_d);
    if (tid == 0) {
        o.ssd = ssd;
        float rv[IPB_RS_MAXR];
        for (int r = 0; r < IPB_RS_MAXR; ++r) rv[r] = ipb_rs_value(c, kmin + r_prefix[r]);
        o.vmin = ipb_rs_value(c, kmin);
        o.vmax = ipb_rs_value(c, kmax);
        for (int i = 0; i < IPB_RS_MAXQ; ++i) {
            if (job.qkind[i] == IPB_QKIND_PCT) o.q[i] = ipb_np_lerp_f32(rv[2 * i], rv[2 * i + 1], qi[i].gamma);
            else if (job.qkind[i] == IPB_QKIND_MEDIAN)
                o.q[i] = (n & 1ull) ? rv[2 * i] : ipb_np_mid2_f32(rv[2 * i], rv[2 * i + 1]);
        }
        out[blockIdx.x] = o;
    }
}
