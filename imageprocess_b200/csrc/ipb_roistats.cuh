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
// One CTA per job: pass A gets n, sum, min/max key; then a radix select with 10-bit digits
// over (key - kmin), i.e. only as many passes as the region's own key range needs (1-2 for
// uint16 data, <= 4 for floats), resolving all requested ranks (p5 / median / p95 -> up to 6
// ranks) in the same passes with one shared-memory histogram per distinct prefix.
// numpy's float32 percentile / median arithmetic is replayed from ipb_exact.cuh.
#pragma once
#include "ipb_rt.cuh"
#include "ipb_exact.cuh"

#define IPB_SRC_U16 0
#define IPB_SRC_F32 1
#define IPB_RS_THREADS 256
#define IPB_RS_MAXQ 3
#define IPB_RS_MAXR (2 * IPB_RS_MAXQ)
#define IPB_RS_DIGIT 10
#define IPB_RS_BINS (1 << IPB_RS_DIGIT)

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

// Calls f(key, value) for every measured pixel of the region; returns via `area` the number
// of region pixels seen by this thread.
template <typename F>
__device__ __forceinline__ void ipb_rs_foreach(const IpbRsCtx& c, unsigned long long& area, F f) {
    const long long nwords = (long long)c.h * c.wpr;
    for (long long wi = threadIdx.x; wi < nwords; wi += blockDim.x) {
        const int r = (int)(wi / c.wpr), j = (int)(wi % c.wpr);
        unsigned m = c.mask[wi];
        if (!m) continue;
        const int y = c.y0 + r, xb = c.x0 + 32 * j;
        if (c.androw0) {
            const unsigned* ar = c.androw0 + (size_t)y * c.and_wpr;
            const int k = xb >> 5, s = xb & 31;
            unsigned lo = ar[k] >> s;
            if (s && k + 1 < c.and_wpr) lo |= ar[k + 1] << (32 - s);
            m &= lo;
        }
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1;
            const int x = xb + b;
            ++area;
            if (c.src == IPB_SRC_U16) {
                const unsigned v = c.u16[(size_t)y * c.W + x];
                f(v, ipb_rs_transform(c, v));
            } else {
                const float v = c.f32[(size_t)y * c.W + x];
                if (isfinite(v)) f(ipb_f32_key(v), v);
            }
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

// exclusive block scan of one u64 per thread (blockDim.x <= 1024); `sm` holds >= 32 entries
__device__ __forceinline__ unsigned long long ipb_block_excl_scan_u64(unsigned long long v, unsigned long long* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    unsigned long long base = 0;
    for (int i = 0; i < warp; ++i) base += sm[i];
    return base + incl - v;
}

__global__ void __launch_bounds__(IPB_RS_THREADS)
ipb_k_region_stats(const IpbRegion* __restrict__ regions, const IpbStatJob* __restrict__ jobs,
                   const unsigned* __restrict__ mask_pool, const unsigned* __restrict__ and_bits,
                   int and_wpr, int H, int W,
                   const unsigned short* __restrict__ planes, const float* __restrict__ images,
                   const float* __restrict__ bvals, IpbStatOut* __restrict__ out)
{
    __shared__ unsigned hist[IPB_RS_MAXR][IPB_RS_BINS];
    __shared__ double red_d[32];
    __shared__ unsigned long long red_u[32];
    __shared__ unsigned red_k[2][32];
    __shared__ unsigned long long r_rank[IPB_RS_MAXR];   // remaining rank inside its group
    __shared__ unsigned r_prefix[IPB_RS_MAXR];           // key' bits resolved so far (>> shift)
    __shared__ int r_group[IPB_RS_MAXR];
    __shared__ unsigned g_prefix[IPB_RS_MAXR];
    __shared__ int g_n;
    __shared__ unsigned long long part[IPB_RS_THREADS];

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

    // ---- pass A: n, area, sum, key range
    unsigned long long n_t = 0, area_t = 0;
    double s_t = 0.0;
    unsigned kmin_t = 0xffffffffu, kmax_t = 0u;
    ipb_rs_foreach(c, area_t, [&](unsigned key, float val) {
        ++n_t; s_t += (double)val;
        kmin_t = key < kmin_t ? key : kmin_t;
        kmax_t = key > kmax_t ? key : kmax_t;
    });
    const unsigned long long n = ipb_block_sum_u64(n_t, red_u);
    const unsigned long long area = ipb_block_sum_u64(area_t, red_u);
    const double sum = ipb_block_sum_d(s_t, red_d);
    kmin_t = ipb_warp_min(kmin_t); kmax_t = ipb_warp_max(kmax_t);
    __syncthreads();
    if (lane == 0) { red_k[0][warp] = kmin_t; red_k[1][warp] = kmax_t; }
    __syncthreads();
    unsigned kmin = 0xffffffffu, kmax = 0u;
    for (int i = 0; i < nwarps; ++i) { kmin = red_k[0][i] < kmin ? red_k[0][i] : kmin; kmax = red_k[1][i] > kmax ? red_k[1][i] : kmax; }

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
        const unsigned dmask = (shift_prev - shift) >= 32 ? 0xffffffffu : ((1u << (shift_prev - shift)) - 1u);
        const int ng = g_n;
        for (int i = tid; i < ng * IPB_RS_BINS; i += blockDim.x) (&hist[0][0])[i] = 0u;
        __syncthreads();
        unsigned long long dummy = 0;
        ipb_rs_foreach(c, dummy, [&](unsigned key, float val) {
            const unsigned kp = key - kmin;
            if (first) { const double d = (double)val - mean; ssd_t += d * d; }
            const unsigned hi = shift_prev >= 32 ? 0u : (kp >> shift_prev);
            const unsigned dg = (kp >> shift) & dmask;
            for (int g = 0; g < ng; ++g)
                if (first || hi == g_prefix[g]) atomicAdd(&hist[g][dg], 1u);
        });
        __syncthreads();
        // locate every rank inside its group's histogram (block-wide scan per group)
        unsigned long long kk_l[IPB_RS_MAXR];
        int grp_l[IPB_RS_MAXR];
        for (int r = 0; r < IPB_RS_MAXR; ++r) { kk_l[r] = r_rank[r]; grp_l[r] = r_group[r]; }
        __syncthreads();
        for (int g = 0; g < ng; ++g) {
            const int per = IPB_RS_BINS / IPB_RS_THREADS;        // 4 bins per thread
            unsigned long long mine = 0;
            for (int b = 0; b < per; ++b) mine += hist[g][tid * per + b];
            const unsigned long long lo = ipb_block_excl_scan_u64(mine, part), hi = lo + mine;
            for (int r = 0; r < nr; ++r) {
                if (grp_l[r] != g) continue;
                const unsigned long long kk = kk_l[r];
                if (kk >= lo && kk < hi) {
                    unsigned long long acc = lo;
                    for (int b = 0; b < per; ++b) {
                        const unsigned cnt = hist[g][tid * per + b];
                        if (kk < acc + cnt) {
                            r_prefix[r] = (g_prefix[g] << (shift_prev - shift)) | (unsigned)(tid * per + b);
                            r_rank[r] = kk - acc;
                            break;
                        }
                        acc += cnt;
                    }
                }
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
        for (int r = 0; r < IPB_RS_MAXR; ++r) {
            const unsigned key = kmin + r_prefix[r];
            rv[r] = (job.src == IPB_SRC_U16) ? ipb_rs_transform(c, key) : ipb_key_f32(key);
        }
        o.vmin = (job.src == IPB_SRC_U16) ? ipb_rs_transform(c, kmin) : ipb_key_f32(kmin);
        o.vmax = (job.src == IPB_SRC_U16) ? ipb_rs_transform(c, kmax) : ipb_key_f32(kmax);
        for (int i = 0; i < IPB_RS_MAXQ; ++i) {
            if (job.qkind[i] == IPB_QKIND_PCT) o.q[i] = ipb_np_lerp_f32(rv[2 * i], rv[2 * i + 1], qi[i].gamma);
            else if (job.qkind[i] == IPB_QKIND_MEDIAN)
                o.q[i] = (n & 1ull) ? rv[2 * i] : ipb_np_mid2_f32(rv[2 * i], rv[2 * i + 1]);
        }
        out[blockIdx.x] = o;
    }
}
