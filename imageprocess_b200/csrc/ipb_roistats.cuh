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
//                background levels of one channel -- and all views share the walk and the
//                selection; each view gets its own output row.
//   IPB_SRC_F32  value = a float32 image pixel (ratio image); non-finite values are
//                dropped like the reference's np.isfinite filter; key = ordered-uint32.
//   IPB_SRC_RATIO value = (eff(N - bg_n) + eps) / (eff(D - bg_d) + eps), eff = max(., 0) when
//                clip_neg, NaN above clip_max: the per-ROI ratio Nesprin2 re-derives after its
//                annulus background (Nesprin2_FRET_Builder.py:1528-1535) from the corrected
//                channel images N = images[plane], D = images[clip_neg[0]]; the six float
//                parameters {bg_n, bg_d, eps, clip_neg, clip_on, clip_max} sit at
//                bvals[bidx[0]].  Otherwise handled like IPB_SRC_F32.
//
// One CTA (1024 threads, one per SM) per job.
//   walk     the region's mask words are taken 1024 at a time: every thread fetches one
//            word (+ the AND plane), a block scan of the popcounts gives every word its
//            output position, and the words, positions and pixel offsets are parked in
//            shared memory.  Each warp then owns 32 consecutive words and issues the
//            coalesced pixel loads of 8 words back to back (lane b <-> bit b), so 8
//            independent loads per lane are in flight instead of a mask->pixel chain.
//   uint16, n <= 65535 (every ROI of the C1-C4 workloads): the walk counts straight into a
//            FULL-RANGE 65536-bin histogram of packed 16-bit counters (128 KB of shared
//            memory, no key store, no dependence on the value range -- saturated pixels
//            cost nothing).  One warp-cooperative pass over the bins then yields n, min,
//            max, every rank and the exact per-view sums (sum of count * T(bin)); a second
//            one the squared deviations.  No per-pixel floating point at all.
//   float32 (and uint16 regions with more pixels): keys are compacted into shared memory;
//            pass 1 is ONE wide histogram over the top bits of (key - kmin) using all shared
//            memory the key store left free; the few keys that share a wanted bin are
//            compacted into a candidate list, and the remaining low bits are resolved by
//            8-bit digit passes over that short list (over all keys if the list overflows).
//            Regions that do not fit in shared memory keep the same code path but re-walk
//            global memory in every pass.
// numpy's float32 percentile / median arithmetic is replayed from ipb_exact.cuh.
#pragma once
#include "ipb_rt.cuh"
#include "ipb_exact.cuh"
#include "ipb_scan.cuh"

#define IPB_SRC_U16 0
#define IPB_SRC_F32 1
#define IPB_SRC_RATIO 2     // float32 ratio re-derived per pixel from two float32 images (Nesprin2 annulus)
#define IPB_RS_THREADS 1024
#define IPB_RS_MAXQ 3
#define IPB_RS_MAXR (2 * IPB_RS_MAXQ)
#define IPB_RS_MAXV 2
#define IPB_RS_DIGIT 8
#define IPB_RS_BINS (1 << IPB_RS_DIGIT)
#define IPB_RS_BATCH 8                      // pixel loads in flight per lane
#define IPB_RS_SMEM_BYTES (212 * 1024)      // dynamic shared memory: key store + histogram
#define IPB_RS_LISTCAP 4096
// bytes of the dynamic store always left to the histogram / candidate list + digit histograms
#define IPB_RS_RESERVE ((IPB_RS_LISTCAP + IPB_RS_MAXR * IPB_RS_BINS) * 4)
#define IPB_RS_SENTINEL 0xffffffffu         // key slot of a dropped (non-finite) float pixel
#define IPB_RS_PACKED_MAX 65535u            // packed 16-bit counters are exact up to this many pixels

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
    const float* f32b; float rp[6];      // IPB_SRC_RATIO: second image and its parameters
};

// np.maximum(x, 0.0) semantics (NaN propagates)
__device__ __forceinline__ float ipb_np_max0(float x) { return (x != x) ? x : (x > 0.0f ? x : 0.0f); }

__device__ __forceinline__ float ipb_rs_ratio(const IpbRsCtx& c, float nv, float dv) {
    float n = __fsub_rn(nv, c.rp[0]), d = __fsub_rn(dv, c.rp[1]);
    if (c.rp[3] != 0.0f) { n = ipb_np_max0(n); d = ipb_np_max0(d); }
    float r = __fdiv_rn(__fadd_rn(n, c.rp[2]), __fadd_rn(d, c.rp[2]));
    if (c.rp[4] != 0.0f && r > c.rp[5]) r = __uint_as_float(0x7fc00000u);
    return r;
}

struct IpbRsWalkSh {          // shared staging of one 1024-word chunk of the walk
    unsigned mw[IPB_RS_THREADS];    // mask word
    unsigned mb[IPB_RS_THREADS];    // output position of the word's first pixel
    unsigned mp[IPB_RS_THREADS];    // plane offset of the word's bit 0
    unsigned wt[32];                // per-warp popcount totals
};

__device__ __forceinline__ float ipb_rs_transform(float B, int clip, unsigned v) {
    float t = __fsub_rn((float)v, B);
    if (clip && t < 0.0f) t = 0.0f;
    return t;
}

// mask word `wi` (row r, word j) of the region, ANDed with the AND plane when there is one
__device__ __forceinline__ unsigned ipb_rs_mask_word(const IpbRsCtx& c, unsigned wi, unsigned r, unsigned j) {
    unsigned m = c.mask[wi];
    if (c.androw0 && m) {
        const unsigned* ar = c.androw0 + (size_t)(c.y0 + (int)r) * c.and_wpr;
        const int xb = c.x0 + 32 * (int)j, k = xb >> 5, s = xb & 31;
        unsigned lo = ar[k] >> s;
        if (s && k + 1 < c.and_wpr) lo |= ar[k + 1] << (32 - s);
        m &= lo;
    }
    return m;
}

// Walks the region: f(raw, pos) for every region pixel, raw = the uint16 value or the
// float32 bit pattern, pos = the pixel's raster-order index inside the region (only when
// NEED_POS).  Returns the number of region pixels.  All threads of the CTA must call it.
template <int SRC, bool NEED_POS, typename F>
__device__ __forceinline__ unsigned ipb_rs_walk(const IpbRsCtx& c, IpbRsWalkSh& sh, F f) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nwords = (unsigned)c.h * (unsigned)c.wpr;
    const unsigned lt = (1u << lane) - 1u;
    unsigned run = 0;
    for (unsigned c0 = 0; c0 < nwords; c0 += IPB_RS_THREADS) {
        const unsigned wi = c0 + (unsigned)tid;
        unsigned m = 0, p = 0;
        if (wi < nwords) {
            const unsigned r = wi / (unsigned)c.wpr, j = wi - r * (unsigned)c.wpr;
            m = ipb_rs_mask_word(c, wi, r, j);
            p = (unsigned)(c.y0 + (int)r) * (unsigned)c.W + (unsigned)(c.x0 + 32 * (int)j);
        }
        unsigned base = 0;
        if (NEED_POS) {
            const unsigned cnt = (unsigned)__popc(m);
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) sh.wt[warp] = incl;
            __syncthreads();
            unsigned before = 0, total = 0;
            for (int i = 0; i < (IPB_RS_THREADS >> 5); ++i) { const unsigned t = sh.wt[i]; total += t; if (i < warp) before += t; }
            base = run + before + incl - cnt;
            run += total;
        }
        sh.mw[tid] = m; sh.mb[tid] = base; sh.mp[tid] = p;
        __syncthreads();
#pragma unroll 1
        for (int k = 0; k < 32; k += IPB_RS_BATCH) {
            unsigned word[IPB_RS_BATCH], raw[IPB_RS_BATCH], any = 0;
#pragma unroll
            for (int u = 0; u < IPB_RS_BATCH; ++u) { word[u] = sh.mw[warp * 32 + k + u]; any |= word[u]; }
            if (!any) continue;                                              // warp-uniform
#pragma unroll
            for (int u = 0; u < IPB_RS_BATCH; ++u) {
                raw[u] = 0u;
                if ((word[u] >> lane) & 1u) {
                    const size_t a = (size_t)sh.mp[warp * 32 + k + u] + (size_t)lane;
                    if (SRC == IPB_SRC_U16) raw[u] = (unsigned)c.u16[a];
                    else if (SRC == IPB_SRC_F32) raw[u] = __float_as_uint(c.f32[a]);
                    else raw[u] = __float_as_uint(ipb_rs_ratio(c, c.f32[a], c.f32b[a]));
                }
            }
#pragma unroll
            for (int u = 0; u < IPB_RS_BATCH; ++u)
                if ((word[u] >> lane) & 1u)
                    f(raw[u], NEED_POS ? sh.mb[warp * 32 + k + u] + (unsigned)__popc(word[u] & lt) : 0u);
        }
        __syncthreads();
    }
    if (!NEED_POS) return 0u;
    return run;
}

// uint16 regions without an AND plane in frames whose rows are 16-byte aligned: the walk by
// aligned 8-pixel units.  Every thread takes units (row, unit) of the region's rect in flat order,
// cuts the unit's 8 mask bits out of the (unaligned) mask row with one funnel shift, and reads the
// 8 pixels with ONE 128-bit load; four units are in flight per thread.  ~6 instructions per pixel
// instead of the ~50 of the lane-per-pixel walk (which also serves every other case).
// f(value, in_region): called for all 8 pixels of a unit that holds a region pixel.
template <typename F>
__device__ __forceinline__ void ipb_rs_walk_u16x8(const IpbRsCtx& c, F f) {
    const int k0 = c.x0 >> 3, s = c.x0 & 7;
    const unsigned nunits = (unsigned)(((c.x0 + c.w + 7) >> 3) - k0);
    const unsigned total = (unsigned)c.h * nunits;
    // row of a unit by one multiply-high; nunits == 1 (a rect inside ONE aligned 8-pixel column) has no 32-bit
    // magic (2^32) and takes the division, as do rects too large for the exactness bound
    const unsigned magic = (unsigned)((0x100000000ull + nunits - 1) / nunits);     // 0 for nunits == 1
    const bool fastdiv = magic != 0u && (unsigned long long)total * nunits < 0xffffffffull;
    for (unsigned i0 = threadIdx.x; i0 < total; i0 += 4u * blockDim.x) {
        uint4 q[4];
        unsigned bits[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const unsigned idx = i0 + (unsigned)g * blockDim.x;
            bits[g] = 0u;
            q[g] = make_uint4(0, 0, 0, 0);
            if (idx < total) {
                const unsigned r = fastdiv ? __umulhi(idx, magic) : idx / nunits;
                const unsigned u = idx - r * nunits;
                const int b0 = 8 * (int)u - s;                           // region x of the unit's first pixel (>= -7)
                const int j = b0 >> 5, sh = b0 & 31;
                const unsigned* mrow = c.mask + (size_t)r * c.wpr;
                const unsigned lo = (j >= 0 && j < c.wpr) ? mrow[j] : 0u;
                const unsigned hi = (j + 1 < c.wpr) ? mrow[j + 1] : 0u;
                unsigned b = __funnelshift_r(lo, hi, (unsigned)sh) & 0xffu;
                if (b0 + 8 > c.w) b &= (1u << (c.w - b0)) - 1u;            // pixels beyond the rect
                bits[g] = b;
                if (b) q[g] = __ldg(reinterpret_cast<const uint4*>(c.u16 + (size_t)(c.y0 + (int)r) * c.W) + k0 + u);
            }
        }
        // no per-pixel branch: f(value, 0 / 1) for all 8 pixels of a unit that holds any region pixel
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (!bits[g]) continue;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const unsigned w = t < 4 ? (t < 2 ? q[g].x : q[g].y) : (t < 6 ? q[g].z : q[g].w);
                f((t & 1) ? (w >> 16) : (w & 0xffffu), (bits[g] >> t) & 1u);
            }
        }
    }
}

// block sums: warp partials -> shared memory -> warp 0 folds them with shuffles -> one broadcast
// value (red: >= 33 entries).  Every thread summing all 32 partials itself cost ~10 % of the
// uint16 kernel's instructions.
__device__ __forceinline__ double ipb_block_sum_d(double v, double* red) {
    v = ipb_warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nw ? red[lane] : 0.0;
        t = ipb_warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ unsigned long long ipb_block_sum_u64(unsigned long long v, unsigned long long* red) {
    v = ipb_warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long t = lane < nw ? red[lane] : 0ull;
        t = ipb_warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// f(key) for every measured key: from the shared-memory key store when it holds the whole
// region (slots of dropped pixels hold the sentinel), else by re-walking global memory.
template <int SRC, typename F>
__device__ __forceinline__ void ipb_rs_foreach_key(const IpbRsCtx& c, IpbRsWalkSh& sh, bool in_smem, unsigned n_slots,
                                                   const unsigned* k32, const unsigned short* k16, F f) {
    if (in_smem) {
        if (SRC == IPB_SRC_U16) for (unsigned i = threadIdx.x; i < n_slots; i += blockDim.x) f((unsigned)k16[i]);
        else for (unsigned i = threadIdx.x; i < n_slots; i += blockDim.x) { const unsigned k = k32[i]; if (k != IPB_RS_SENTINEL) f(k); }
    } else {
        ipb_rs_walk<SRC, false>(c, sh, [&](unsigned raw, unsigned) {
            if (SRC == IPB_SRC_U16) f(raw);
            else { const float v = __uint_as_float(raw); if (isfinite(v)) f(ipb_f32_key(v)); }
        });
    }
}

// Shared rank bookkeeping of the selection passes
struct IpbRsSel {
    unsigned long long rank[IPB_RS_MAXR];   // remaining rank inside its group
    unsigned prefix[IPB_RS_MAXR];           // key' bits resolved so far
    int group[IPB_RS_MAXR];
    unsigned gprefix[IPB_RS_MAXR];
    int gn;
};

// thread 0: distinct prefixes -> groups
__device__ __forceinline__ void ipb_rs_regroup(IpbRsSel& s, int nr) {
    int m = 0;
    for (int r = 0; r < nr; ++r) {
        int gg = -1;
        for (int t = 0; t < m; ++t) if (s.gprefix[t] == s.prefix[r]) { gg = t; break; }
        if (gg < 0) { s.gprefix[m] = s.prefix[r]; gg = m++; }
        s.group[r] = gg;
    }
    s.gn = m > 0 ? m : 1;
}

// one job by the whole CTA (every thread of the block calls it)
template <int SRC>
__device__ __forceinline__ void ipb_rs_job(unsigned jidx, const IpbRegion* __restrict__ regions, const IpbStatJob* __restrict__ jobs,
                   const unsigned* __restrict__ mask_pool, const unsigned* __restrict__ and_bits,
                   int and_wpr, int H, int W,
                   const unsigned short* __restrict__ planes, const float* __restrict__ images,
                   const float* __restrict__ bvals, IpbStatOut* __restrict__ out, int smem_bytes,
                   const unsigned char* __restrict__ only)
{
    IPB_DYN_SMEM(unsigned, keystore);
    __shared__ IpbRsWalkSh wsh;
    __shared__ double red_d[33];
    __shared__ unsigned long long red_u[33];
    __shared__ unsigned red_k[2][33];
    __shared__ IpbRsSel sel;
    __shared__ unsigned list_n;
    __shared__ unsigned char gtab[64];

    const IpbStatJob job = jobs[jidx];
    if (job.src != SRC) return;                          // mixed job lists: the other instantiation takes it
    if (only && !only[job.region]) return;               // rerun of the regions the fused ROI kernel could not serve
    const IpbRegion rg = regions[job.region];
    IpbRsCtx c;
    c.mask = mask_pool + rg.mask_off;
    c.androw0 = (rg.use_and && and_bits) ? and_bits + (size_t)rg.and_plane * H * and_wpr : nullptr;
    c.and_wpr = and_wpr;
    c.x0 = rg.x0; c.y0 = rg.y0; c.w = rg.w; c.h = rg.h; c.wpr = rg.wpr;
    c.W = W;
    c.u16 = planes ? planes + (size_t)job.plane * H * W : nullptr;
    c.f32 = images ? images + (size_t)job.plane * H * W : nullptr;
    c.f32b = nullptr;
#pragma unroll
    for (int i = 0; i < 6; ++i) c.rp[i] = 0.0f;
    if (SRC == IPB_SRC_RATIO) {
        c.f32b = images + (size_t)job.clip_neg[0] * H * W;
        if (bvals && job.bidx[0] >= 0) { for (int i = 0; i < 6; ++i) c.rp[i] = bvals[job.bidx[0] + i]; }
    }
    const int nv = (SRC == IPB_SRC_U16) ? (job.n_views < 1 ? 1 : (job.n_views > IPB_RS_MAXV ? IPB_RS_MAXV : job.n_views)) : 1;
    float vB[IPB_RS_MAXV]; int vclip[IPB_RS_MAXV];
#pragma unroll
    for (int v = 0; v < IPB_RS_MAXV; ++v) {
        vB[v] = (SRC == IPB_SRC_U16 && v < nv && job.bidx[v] >= 0) ? bvals[job.bidx[v]] : 0.0f;
        vclip[v] = (v < nv) ? job.clip_neg[v] : 0;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    const float fnan = __uint_as_float(0x7fc00000u);
    const unsigned nwords = (unsigned)c.h * (unsigned)c.wpr;

    // ---- region size (mask popcount): `area`, and the choice of the uint16 strategy
    unsigned long long area;
    {
        unsigned cnt = 0;
        for (unsigned wi = tid; wi < nwords; wi += blockDim.x) {
            const unsigned r = wi / (unsigned)c.wpr, j = wi - r * (unsigned)c.wpr;
            cnt += (unsigned)__popc(ipb_rs_mask_word(c, wi, r, j));
        }
        area = ipb_block_sum_u64((unsigned long long)cnt, red_u);
    }

    unsigned long long n = 0;
    unsigned kmin = 0xffffffffu, kmax = 0u;     // key range of the sample
    unsigned kbase = 0u;                        // sel.prefix is relative to this key
    IpbQIdx qi[IPB_RS_MAXQ];
    int nr = 0;
    double v_sum[IPB_RS_MAXV], v_ssd[IPB_RS_MAXV];
#pragma unroll
    for (int v = 0; v < IPB_RS_MAXV; ++v) { v_sum[v] = 0.0; v_ssd[v] = 0.0; }
    for (int i = 0; i < IPB_RS_MAXQ; ++i) { qi[i].prev = qi[i].next = 0; qi[i].gamma = 0.f; }

    // ranks wanted for a sample of nn_ values -> sel (thread 0 writes, everyone syncs)
    auto setup_ranks = [&](unsigned long long nn_) {
        nr = 0;
        for (int i = 0; i < IPB_RS_MAXQ; ++i) {
            if (job.qkind[i] == IPB_QKIND_PCT) qi[i] = ipb_np_qidx_f32((long long)nn_, job.q32[i]);
            else if (job.qkind[i] == IPB_QKIND_MEDIAN) {
                if (nn_ & 1ull) qi[i].prev = qi[i].next = (long long)(nn_ >> 1);
                else { qi[i].prev = (long long)(nn_ >> 1) - 1; qi[i].next = (long long)(nn_ >> 1); }
            }
            if (job.qkind[i] != IPB_QKIND_NONE) nr = 2 * (i + 1);
        }
        if (tid == 0) {
            for (int i = 0; i < IPB_RS_MAXQ; ++i) {
                sel.rank[2 * i] = (unsigned long long)qi[i].prev; sel.rank[2 * i + 1] = (unsigned long long)qi[i].next;
                sel.prefix[2 * i] = sel.prefix[2 * i + 1] = 0u;
                sel.group[2 * i] = sel.group[2 * i + 1] = 0;
            }
            sel.gprefix[0] = 0u; sel.gn = 1;
        }
        __syncthreads();
    };
    // block-wide min / max of per-thread key bounds
    auto block_minmax = [&](unsigned lo_t, unsigned hi_t) {
        lo_t = ipb_warp_min(lo_t); hi_t = ipb_warp_max(hi_t);
        __syncthreads();
        if (lane == 0) { red_k[0][warp] = lo_t; red_k[1][warp] = hi_t; }
        __syncthreads();
        if (warp == 0) {
            unsigned a = lane < nwarps ? red_k[0][lane] : 0xffffffffu, b = lane < nwarps ? red_k[1][lane] : 0u;
            a = ipb_warp_min(a); b = ipb_warp_max(b);
            if (lane == 0) { red_k[0][32] = a; red_k[1][32] = b; }
        }
        __syncthreads();
        kmin = red_k[0][32] < kmin ? red_k[0][32] : kmin;
        kmax = red_k[1][32] > kmax ? red_k[1][32] : kmax;
    };

    if (SRC == IPB_SRC_U16 && area > 0 && area <= (unsigned long long)IPB_RS_PACKED_MAX) {
        // ================= uint16 fast path: full-range histogram of packed 16-bit counters
        unsigned* h16 = keystore;                                  // 32768 words = 65536 counters
        for (unsigned i = tid; i < 32768u; i += blockDim.x) h16[i] = 0u;
        __syncthreads();
        if (!c.androw0 && (W & 7) == 0 && (((size_t)c.u16) & 15) == 0)                    // block-uniform
            ipb_rs_walk_u16x8(c, [&](unsigned raw, unsigned on) { atomicAdd(&h16[raw >> 1], on << ((raw & 1u) << 4)); });
        else
            ipb_rs_walk<SRC, false>(c, wsh, [&](unsigned raw, unsigned) { atomicAdd(&h16[raw >> 1], 1u << ((raw & 1u) << 4)); });
        n = area;
        setup_ranks(n);
        // one pass over the bins: key range, per-view sum and sum of squares (the squared
        // deviations follow as sumsq - sum^2 / n in float64: the terms are counts times values
        // with at most 17 significant bits, far inside float64's exact-integer range)
        unsigned lo_t = 0xffffffffu, hi_t = 0u;
        double s[IPB_RS_MAXV], q[IPB_RS_MAXV];
#pragma unroll
        for (int v = 0; v < IPB_RS_MAXV; ++v) { s[v] = 0.0; q[v] = 0.0; }
        // every thread owns the 32 consecutive words [32 tid, 32 tid + 32) (rotated start: the lanes
        // of a warp hit different banks) and keeps their pixel count for the rank location below
        unsigned mine = 0;
        for (unsigned ii = 0; ii < 32u; ++ii) {
            const unsigned i = 32u * (unsigned)tid + ((ii + (unsigned)tid) & 31u);
            const unsigned w = h16[i];
            if (!w) continue;
            const unsigned lo = w & 0xffffu, hi = w >> 16, k0 = 2u * i;
            mine += lo + hi;
            const unsigned first = lo ? k0 : k0 + 1u, last = hi ? k0 + 1u : k0;
            lo_t = first < lo_t ? first : lo_t;
            hi_t = last > hi_t ? last : hi_t;
#pragma unroll
            for (int v = 0; v < IPB_RS_MAXV; ++v)
                if (v < nv) {
                    const double t0 = (double)ipb_rs_transform(vB[v], vclip[v], k0);
                    const double t1 = (double)ipb_rs_transform(vB[v], vclip[v], k0 + 1u);
                    const double c0 = (double)lo * t0, c1 = (double)hi * t1;
                    s[v] += c0 + c1;
                    q[v] += c0 * t0 + c1 * t1;
                }
        }
        block_minmax(lo_t, hi_t);
        for (int v = 0; v < nv; ++v) {
            v_sum[v] = ipb_block_sum_d(s[v], red_d);
            const double sumsq = ipb_block_sum_d(q[v], red_d);
            v_ssd[v] = (kmin == kmax) ? 0.0 : sumsq - v_sum[v] * (v_sum[v] / (double)n);   // constant region: exactly 0
        }
        // ranks: the word first, then the low / high counter inside the word
        unsigned long long want[IPB_RS_MAXR];
        for (int r = 0; r < IPB_RS_MAXR; ++r) want[r] = sel.rank[r];
        ipb_locate_ranks_chunks(32768u, 32768u / IPB_RS_THREADS, (unsigned long long)mine, want, nr, red_u,
                      [&](unsigned i) { const unsigned w = h16[i]; return (w & 0xffffu) + (w >> 16); },
                      [&](int r, unsigned i, unsigned inside) { sel.prefix[r] = 2u * i + (inside >= (h16[i] & 0xffffu) ? 1u : 0u); });
        kbase = 0u;                                                // prefixes are absolute keys
    } else {
        // ================= keyed path (float32; uint16 regions above 65535 px)
        unsigned* k32 = keystore;
        unsigned short* k16 = reinterpret_cast<unsigned short*>(keystore);
        const int keysize = SRC == IPB_SRC_U16 ? 2 : 4;
        const unsigned cap = (unsigned)((smem_bytes - IPB_RS_RESERVE) / keysize);
        const bool in_smem = area <= (unsigned long long)cap;
        unsigned n_t = 0, lo_t = 0xffffffffu, hi_t = 0u;
        double s_t = 0.0, s2_t = 0.0;
        auto take = [&](unsigned raw, unsigned pos) {
                unsigned key;
                if (SRC == IPB_SRC_U16) key = raw;
                else {
                    const float v = __uint_as_float(raw);
                    if (isfinite(v)) { key = ipb_f32_key(v); const double d = (double)v; s_t += d; if (!in_smem) s2_t += d * d; }
                    else key = IPB_RS_SENTINEL;
                }
                if (in_smem) { if (SRC == IPB_SRC_U16) k16[pos] = (unsigned short)key; else k32[pos] = key; }
                if (SRC == IPB_SRC_U16 || key != IPB_RS_SENTINEL) {
                    ++n_t;
                    lo_t = key < lo_t ? key : lo_t;
                    hi_t = key > hi_t ? key : hi_t;
                }
            };
        if (area > 0) {
            // (a walk by aligned 4-pixel float units, ipb_rs_walk_f32x4, measured slower here -- 345 vs
            // 310 us per step: four pixels do not amortise a unit's mask cut-out and slot scan)
            ipb_rs_walk<SRC, true>(c, wsh, take);
        }
        n = ipb_block_sum_u64((unsigned long long)n_t, red_u);
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
        block_minmax(lo_t, hi_t);
        setup_ranks(n);
        kbase = kmin;

        // ---- pass 1: one wide histogram (as many bins as the free part of the store allows)
        const unsigned n_slots = in_smem ? (unsigned)area : 0u;
        const unsigned range = kmax - kmin;
        const int bits = range ? (32 - __clz((int)range)) : 0;
        const unsigned key_words = in_smem ? (unsigned)(((size_t)area * keysize + 3) / 4) : 0u;
        unsigned* whist = keystore + key_words;
        const unsigned hb_cap = (unsigned)(smem_bytes / 4) - key_words;        // >= RESERVE/4 words
        int d1 = 31 - __clz((int)hb_cap);
        if (d1 > 15) d1 = 15;
        if (d1 > bits) d1 = bits;
        const int rb = bits - d1;                                               // bits left after pass 1
        const unsigned nb = 1u << d1;
        for (unsigned i = tid; i < nb; i += blockDim.x) whist[i] = 0u;
        // float32: the sum is known from the walk, so the squared deviations from the exact mean
        // ride along with the histogram pass over the stored keys
        double f_sum = 0.0, f_q = 0.0;
        if (SRC != IPB_SRC_U16) f_sum = ipb_block_sum_d(s_t, red_d);
        const double f_mean = f_sum / (double)n;
        __syncthreads();
        if (SRC != IPB_SRC_U16 && in_smem) {
            for (unsigned i = tid; i < n_slots; i += blockDim.x) {
                const unsigned k = k32[i];
                if (k != IPB_RS_SENTINEL) {
                    atomicAdd(&whist[(k - kmin) >> rb], 1u);
                    const double d = (double)ipb_key_f32(k) - f_mean;
                    f_q += d * d;
                }
            }
        } else {
            ipb_rs_foreach_key<SRC>(c, wsh, in_smem, n_slots, k32, k16, [&](unsigned key) { atomicAdd(&whist[(key - kmin) >> rb], 1u); });
        }
        __syncthreads();
        {
            unsigned long long want[IPB_RS_MAXR];
            for (int r = 0; r < IPB_RS_MAXR; ++r) want[r] = sel.rank[r];
            ipb_locate_ranks_smem(nb, (nb + IPB_RS_THREADS - 1u) / IPB_RS_THREADS, want, nr, red_u,
                          [&](unsigned i) { return whist[i]; },
                          [&](int r, unsigned i, unsigned inside) { sel.prefix[r] = i; sel.rank[r] = (unsigned long long)inside; });
        }

        // ---- sums.  uint16: per view from the exact histogram (or the keys when bins are
        //      coarse); float32: sum from the walk, squared deviations from the stored keys.
        if (SRC == IPB_SRC_U16) {
            for (int v = 0; v < nv; ++v) {
                const float B = vB[v]; const int clip = vclip[v];
                double s = 0.0;
                if (rb == 0) { for (unsigned b = tid; b < nb; b += blockDim.x) { const unsigned cnt = whist[b]; if (cnt) s += (double)cnt * (double)ipb_rs_transform(B, clip, kmin + b); } }
                else ipb_rs_foreach_key<SRC>(c, wsh, in_smem, n_slots, k32, k16, [&](unsigned key) { s += (double)ipb_rs_transform(B, clip, key); });
                const double sum = ipb_block_sum_d(s, red_d);
                const double mean = sum / (double)n;
                double q = 0.0;
                if (rb == 0) { for (unsigned b = tid; b < nb; b += blockDim.x) { const unsigned cnt = whist[b]; if (cnt) { const double d = (double)ipb_rs_transform(B, clip, kmin + b) - mean; q += (double)cnt * d * d; } } }
                else ipb_rs_foreach_key<SRC>(c, wsh, in_smem, n_slots, k32, k16, [&](unsigned key) { const double d = (double)ipb_rs_transform(B, clip, key) - mean; q += d * d; });
                v_sum[v] = sum;
                v_ssd[v] = ipb_block_sum_d(q, red_d);
            }
        } else {
            const double sum = f_sum;
            double ssd;
            if (in_smem) {
                ssd = ipb_block_sum_d(f_q, red_d);
            } else {
                const double sumsq = ipb_block_sum_d(s2_t, red_d);
                ssd = sumsq - sum * (sum / (double)n);
            }
            v_sum[0] = sum; v_ssd[0] = ssd;
        }

        if (rb > 0 && nr > 0) {
            // ---- candidates: the few keys that share a wanted pass-1 bin
            if (tid == 0) { ipb_rs_regroup(sel, nr); list_n = 0u; }
            __syncthreads();
            const int ng0 = sel.gn;
            unsigned gp0[IPB_RS_MAXR];
            for (int g = 0; g < IPB_RS_MAXR; ++g) gp0[g] = g < ng0 ? sel.gprefix[g] : 0xffffffffu;
            unsigned* list = whist;                              // the wide histogram is no longer needed
            unsigned* ghist = whist + IPB_RS_LISTCAP;            // [group][256] digit histograms
            const unsigned lowmask = (rb >= 32) ? 0xffffffffu : ((1u << rb) - 1u);
            const bool packable = rb <= 28;
            __syncthreads();
            if (packable) {
                // most keys share no pass-1 bin with a wanted rank: gtab[bin mod 64] = the groups whose
                // bin has that residue (usually none): one shared-memory byte instead of six compares
                if (tid < 64) {
                    unsigned mk = 0;
                    for (int g = 0; g < ng0; ++g) if ((gp0[g] & 63u) == (unsigned)tid) mk |= 1u << g;
                    gtab[tid] = (unsigned char)mk;
                }
                __syncthreads();
                ipb_rs_foreach_key<SRC>(c, wsh, in_smem, n_slots, k32, k16, [&](unsigned key) {
                    const unsigned kp = key - kmin, hi = kp >> rb;
                    unsigned mk = gtab[hi & 63u];
                    while (mk) {
                        const int g = __ffs((int)mk) - 1;
                        mk &= mk - 1u;
                        if (hi == sel.gprefix[g]) {
                            const unsigned idx = atomicAdd(&list_n, 1u);
                            if (idx < IPB_RS_LISTCAP) list[idx] = ((unsigned)g << 28) | (kp & lowmask);
                        }
                    }
                });
            }
            __syncthreads();
            const unsigned m = list_n;
            const bool use_list = packable && m <= IPB_RS_LISTCAP;
            // ---- 8-bit digit passes over the candidates (or over every key), one histogram
            //      per distinct prefix, until all bits are resolved
            int shift_prev = rb;
            do {
                const int shift = shift_prev > IPB_RS_DIGIT ? shift_prev - IPB_RS_DIGIT : 0;
                const unsigned dmask = (1u << (shift_prev - shift)) - 1u;
                const int ngc = sel.gn;
                for (int i = tid; i < ngc * IPB_RS_BINS; i += blockDim.x) ghist[i] = 0u;
                __syncthreads();
                auto count = [&](unsigned kp) {
                    const unsigned hi = shift_prev >= 32 ? 0u : (kp >> shift_prev);
                    const unsigned dg = (kp >> shift) & dmask;
                    for (int g = 0; g < ngc; ++g)
                        if (hi == sel.gprefix[g]) atomicAdd(&ghist[g * IPB_RS_BINS + dg], 1u);
                };
                if (use_list) {
                    for (unsigned i = tid; i < m; i += blockDim.x) {
                        const unsigned e = list[i];
                        count((gp0[e >> 28] << rb) | (e & lowmask));
                    }
                } else {
                    ipb_rs_foreach_key<SRC>(c, wsh, in_smem, n_slots, k32, k16, [&](unsigned key) { count(key - kmin); });
                }
                __syncthreads();
                if (warp < ngc) {
                    const int g = warp;
                    unsigned cnt[IPB_RS_BINS / 32], mine = 0;
#pragma unroll
                    for (int b = 0; b < IPB_RS_BINS / 32; ++b) { cnt[b] = ghist[g * IPB_RS_BINS + lane * (IPB_RS_BINS / 32) + b]; mine += cnt[b]; }
                    unsigned long long incl = mine;
#pragma unroll
                    for (int o2 = 1; o2 < 32; o2 <<= 1) { unsigned long long t = __shfl_up_sync(IPB_FULL, incl, o2); if (lane >= o2) incl += t; }
                    const unsigned long long lo = incl - mine, hi = incl;
                    for (int r = 0; r < nr; ++r) {
                        if (sel.group[r] != g) continue;
                        const unsigned long long kk = sel.rank[r];
                        __syncwarp();
                        if (kk >= lo && kk < hi) {
                            unsigned long long acc = lo;
#pragma unroll
                            for (int b = 0; b < IPB_RS_BINS / 32; ++b) {
                                if (kk >= acc && kk < acc + cnt[b]) {
                                    sel.prefix[r] = (sel.gprefix[g] << (shift_prev - shift)) | (unsigned)(lane * (IPB_RS_BINS / 32) + b);
                                    sel.rank[r] = kk - acc;
                                }
                                acc += cnt[b];
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                if (tid == 0) ipb_rs_regroup(sel, nr);
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
            const unsigned key = kbase + sel.prefix[r];
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

// grid = n_jobs (one job per CTA), or -- with `only` -- a small grid whose CTAs scan the job list
// 32 jobs at a time and measure only the regions flagged by the fused ROI kernel (ipb_roifused.cuh)
template <int SRC>
__global__ void __launch_bounds__(IPB_RS_THREADS, 1)
ipb_k_region_stats(const IpbRegion* __restrict__ regions, const IpbStatJob* __restrict__ jobs, int n_jobs,
                   const unsigned* __restrict__ mask_pool, const unsigned* __restrict__ and_bits,
                   int and_wpr, int H, int W,
                   const unsigned short* __restrict__ planes, const float* __restrict__ images,
                   const float* __restrict__ bvals, IpbStatOut* __restrict__ out, int smem_bytes,
                   const unsigned char* __restrict__ only /* null, or per region: != 0 -> measure it */)
{
    if (!only) {
        for (unsigned j = blockIdx.x; j < (unsigned)n_jobs; j += gridDim.x) {
            __syncthreads();                               // the previous job's shared state is dead
            ipb_rs_job<SRC>(j, regions, jobs, mask_pool, and_bits, and_wpr, H, W, planes, images, bvals, out, smem_bytes, only);
        }
        return;
    }
    __shared__ unsigned s_pick;
    for (unsigned c0 = 32u * blockIdx.x; c0 < (unsigned)n_jobs; c0 += 32u * gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 32) {
            const unsigned j = c0 + threadIdx.x;
            const bool f = j < (unsigned)n_jobs && jobs[j].src == SRC && only[jobs[j].region] != 0;
            const unsigned b = __ballot_sync(IPB_FULL, f);
            if (threadIdx.x == 0) s_pick = b;
        }
        __syncthreads();
        unsigned pick = s_pick;
        while (pick) {                                     // block-uniform
            const unsigned k = (unsigned)__ffs((int)pick) - 1u;
            pick &= pick - 1u;
            __syncthreads();
            ipb_rs_job<SRC>(c0 + k, regions, jobs, mask_pool, and_bits, and_wpr, H, W, planes, images, bvals, out, smem_bytes, only);
        }
    }
}
