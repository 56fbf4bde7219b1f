// ipb_hist.cuh -- exact 65 536-bin integer histograms of uint16 planes and the order
// statistics / float32 percentiles derived from them.
//
// Every background / epsilon / preview percentile of the reference is np.percentile over a
// float32 copy of uint16 pixels (all, a strided subsample, or the pixels under a mask).  The
// order statistics of such a sample are integers, so one exact integer histogram per
// (plane, sampling pattern) reproduces them bit-for-bit; numpy's float32 index/lerp
// arithmetic is then replayed by ipb_exact.cuh.
//
// ipb_k_hist_u16: grid (chunks, jobs); one CTA streams a band of rows with 128-bit loads and
// counts into a shared-memory window of the low IPB_HIST_WIN bins (fluorescence data lives
// there); brighter values go straight to global atomics.  The window is flushed with one
// global atomic per non-empty bin.  Optionally the exact integer moments (sum, sum of
// squares) of ALL pixels of the plane are accumulated in the same pass (FA global stats).
#pragma once
#include "ipb_rt.cuh"
#include "ipb_exact.cuh"

#define IPB_HIST_BINS 65536
#define IPB_HIST_WIN 24576
#define IPB_HIST_THREADS 1024

#define IPB_PAT_FULL 0
#define IPB_PAT_STRIDE1D 1   // flat index % k == 0      (vals[::k] of the raveled plane)
#define IPB_PAT_STRIDE2D 2   // y % k == 0 && x % k == 0 (img[::k, ::k])
#define IPB_PAT_MASKED 3     // pixels under the frame's union bitmask
#define IPB_PAT_MASKED_STRIDE 4  // every k-th masked pixel in raster order (img[mask][::k])

struct IpbHistJob {
    int plane;        // index of the uint16 plane (frame * C + channel)
    int pattern;      // IPB_PAT_*
    int k;            // stride for the strided patterns
    int mask_frame;   // frame index into the union bitmask (masked patterns)
    int moments;      // != 0: also accumulate sum / sumsq of ALL pixels of the plane
    int excl_plane1;  // 1 + index of a second plane for the saturation filter, 0 = none
    int sat_min;      // > 0: drop pixels whose value, or the second plane's, is >= sat_min
                      //      (Nesprin2 saturation -> NaN -> np.isfinite filter, 1415-1421,435)
    int pad2;
};

// ---- selection by sampling (ipb_hist_select): per-job value window derived from a sample
#define IPB_HSEL_WIN 4096        // window bins per job in the tail pass
#define IPB_HSEL_MAXJ 4          // jobs fused into one pass over a plane
#define IPB_HSEL_WINDOWED 0
#define IPB_HSEL_FULL 1          // exact full-range histogram instead (sparse patterns, wide windows)
#define IPB_HSEL_NONE 2          // no order statistic wanted from this job
struct IpbHistWin { int wlo, whi, mode, pad; };     // window [wlo, whi)
struct IpbPlanePass { int plane, excl_plane1, sat_min, n_jobs; int job[IPB_HSEL_MAXJ]; };

// deterministic 1/16 sample of the 8-pixel groups (hash of the group's row and column index:
// no row / column periodicity of the image can line up with it)
__device__ __forceinline__ bool ipb_hist_sampled(unsigned y, unsigned xv) {
    unsigned h = y * 0x9E3779B1u + xv * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return (h & 15u) == 0u;
}

// out_stats[job] = { n_selected, sum_all, sumsq_all, reserved }
__global__ void __launch_bounds__(IPB_HIST_THREADS)
ipb_k_hist_u16(const unsigned short* __restrict__ planes, int H, int W,
               const IpbHistJob* __restrict__ jobs, int rows_per_chunk,
               const unsigned* __restrict__ union_bits, int union_wpr,
               unsigned* __restrict__ hist, unsigned long long* __restrict__ out_stats,
               int sample /* != 0: only the hash-sampled 1/16 of the 8-pixel groups */,
               const IpbHistWin* __restrict__ only_full /* non-null: jobs whose mode != FULL are skipped */)
{
    IPB_DYN_SMEM(unsigned, sh);
    if (only_full && only_full[blockIdx.y].mode != IPB_HSEL_FULL) return;
    IpbHistJob job = jobs[blockIdx.y];
    if (sample || only_full) job.moments = 0;        // the selection path takes moments in its tail pass
    const int y_beg = (int)blockIdx.x * rows_per_chunk;
    int y_end = y_beg + rows_per_chunk;
    if (y_end > H) y_end = H;
    unsigned* gh = hist + (size_t)blockIdx.y * IPB_HIST_BINS;
    for (int b = threadIdx.x; b < IPB_HIST_WIN; b += blockDim.x) sh[b] = 0;
    __syncthreads();

    const unsigned short* img = planes + (size_t)job.plane * H * W;
    const unsigned sat_min = job.sat_min > 0 ? (unsigned)job.sat_min : 0xffffffffu;
    const unsigned short* img2 = (job.sat_min > 0 && job.excl_plane1 > 0) ? planes + (size_t)(job.excl_plane1 - 1) * H * W : nullptr;
    const unsigned* ubits = (job.pattern == IPB_PAT_MASKED) ? union_bits + (size_t)job.mask_frame * H * union_wpr : nullptr;
    unsigned long long s1 = 0, s2 = 0, nsel = 0;
    const int k = job.k > 0 ? job.k : 1;
    const bool vec_ok = ((W & 7) == 0) && ((((size_t)img) & 15) == 0) && ((((size_t)img2) & 15) == 0);

#define IPB_HIST_COUNT(v)                                      \
    do {                                                       \
        unsigned vv_ = (v);                                    \
        if (vv_ < IPB_HIST_WIN) atomicAdd(&sh[vv_], 1u);       \
        else atomicAdd(&gh[vv_], 1u);                          \
        ++nsel;                                                \
    } while (0)

    if (y_beg < y_end) {
        if (vec_ok) {
            const int vpr = W >> 3;                                   // 8-pixel vectors per row
            const long long nvec = (long long)(y_end - y_beg) * vpr;
            for (long long i = threadIdx.x; i < nvec; i += blockDim.x) {
                const int y = y_beg + (int)(i / vpr);
                const int x0 = ((int)(i % vpr)) << 3;
                if (job.pattern == IPB_PAT_STRIDE2D && !job.moments && (y % k) != 0) continue;
                if (sample && !ipb_hist_sampled((unsigned)y, (unsigned)(x0 >> 3))) continue;
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(img + (size_t)y * W + x0));
                const unsigned w[4] = {q.x, q.y, q.z, q.w};
                uint4 q2 = make_uint4(0, 0, 0, 0);
                if (img2) q2 = __ldg(reinterpret_cast<const uint4*>(img2 + (size_t)y * W + x0));
                const unsigned w2[4] = {q2.x, q2.y, q2.z, q2.w};
                unsigned sel = 0;                                     // bit i: pixel x0+i selected
                if (job.pattern == IPB_PAT_FULL) sel = 0xffu;
                else if (job.pattern == IPB_PAT_STRIDE1D) {
                    long long flat = (long long)y * W + x0;
                    int first = (int)((k - (flat % k)) % k);
                    for (int t = first; t < 8; t += k) sel |= 1u << t;
                } else if (job.pattern == IPB_PAT_STRIDE2D) {
                    if ((y % k) == 0) {
                        int first = (k - (x0 % k)) % k;
                        for (int t = first; t < 8; t += k) sel |= 1u << t;
                    }
                } else if (job.pattern == IPB_PAT_MASKED) {
                    sel = (ubits[(size_t)y * union_wpr + (x0 >> 5)] >> (x0 & 31)) & 0xffu;
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                    const unsigned o = (t & 1) ? (w2[t >> 1] >> 16) : (w2[t >> 1] & 0xffffu);
                    if (job.moments) { s1 += v; s2 += (unsigned long long)v * v; }
                    if (((sel >> t) & 1u) && v < sat_min && o < sat_min) IPB_HIST_COUNT(v);
                }
            }
        } else {
            const long long npx = (long long)(y_end - y_beg) * W;
            for (long long i = threadIdx.x; i < npx; i += blockDim.x) {
                const int y = y_beg + (int)(i / W), x = (int)(i % W);
                if (sample && !ipb_hist_sampled((unsigned)y, (unsigned)(x >> 3))) continue;
                const unsigned v = img[(size_t)y * W + x];
                if (job.moments) { s1 += v; s2 += (unsigned long long)v * v; }
                bool sel;
                if (job.pattern == IPB_PAT_FULL) sel = true;
                else if (job.pattern == IPB_PAT_STRIDE1D) sel = (((long long)y * W + x) % k) == 0;
                else if (job.pattern == IPB_PAT_STRIDE2D) sel = (y % k) == 0 && (x % k) == 0;
                else if (job.pattern == IPB_PAT_MASKED) sel = (ubits[(size_t)y * union_wpr + (x >> 5)] >> (x & 31)) & 1u;
                else sel = false;
                if (sel && img2 && (unsigned)img2[(size_t)y * W + x] >= sat_min) sel = false;
                if (sel && v < sat_min) IPB_HIST_COUNT(v);
            }
        }
    }
#undef IPB_HIST_COUNT
    __syncthreads();
    for (int b = threadIdx.x; b < IPB_HIST_WIN; b += blockDim.x) {
        const unsigned c = sh[b];
        if (c) atomicAdd(&gh[b], c);
    }
    // block reduction of the three 64-bit counters
    s1 = ipb_warp_sum(s1); s2 = ipb_warp_sum(s2); nsel = ipb_warp_sum(nsel);
    __shared__ unsigned long long red[3][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; red[2][warp] = nsel; }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        unsigned long long a = lane < nw ? red[0][lane] : 0ull;
        unsigned long long b = lane < nw ? red[1][lane] : 0ull;
        unsigned long long c = lane < nw ? red[2][lane] : 0ull;
        a = ipb_warp_sum(a); b = ipb_warp_sum(b); c = ipb_warp_sum(c);
        if (lane == 0) {
            unsigned long long* st = out_stats + (size_t)blockIdx.y * 4;
            if (c) atomicAdd(&st[0], c);
            if (a) atomicAdd(&st[1], a);
            if (b) atomicAdd(&st[2], b);
        }
    }
}

// img[mask][::k]: every k-th masked pixel in raster order.  One CTA per job walks the rows
// in order with a running rank (exclusive scan of per-row popcounts first).  Rare path
// (Fluor_INT bg_scope = "roi_union" with bg_stride > 1, reference Fluor_INT.py:465-470).
__global__ void __launch_bounds__(256)
ipb_k_hist_masked_stride(const unsigned short* __restrict__ planes, int H, int W,
                         const IpbHistJob* __restrict__ jobs, const unsigned* __restrict__ union_bits,
                         int union_wpr, unsigned long long* __restrict__ row_rank /* [jobs][H] scratch */,
                         unsigned* __restrict__ hist, unsigned long long* __restrict__ out_stats)
{
    const IpbHistJob job = jobs[blockIdx.x];
    if (job.pattern != IPB_PAT_MASKED_STRIDE) return;
    const unsigned short* img = planes + (size_t)job.plane * H * W;
    const unsigned* ubits = union_bits + (size_t)job.mask_frame * H * union_wpr;
    unsigned* gh = hist + (size_t)blockIdx.x * IPB_HIST_BINS;
    unsigned long long* rr = row_rank + (size_t)blockIdx.x * H;
    const int k = job.k > 0 ? job.k : 1;
    // per-row popcounts
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        unsigned c = 0;
        for (int j = 0; j < union_wpr; ++j) c += __popc(ubits[(size_t)y * union_wpr + j]);
        rr[y] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                      // H <= 16384: a serial exclusive scan is fine here
        unsigned long long acc = 0;
        for (int y = 0; y < H; ++y) { unsigned long long c = rr[y]; rr[y] = acc; acc += c; }
    }
    __syncthreads();
    unsigned long long nsel = 0;
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        unsigned long long rank = rr[y];
        for (int j = 0; j < union_wpr; ++j) {
            unsigned m = ubits[(size_t)y * union_wpr + j];
            while (m) {
                const int b = __ffs((int)m) - 1;
                m &= m - 1;
                if ((rank % (unsigned long long)k) == 0) {
                    const int x = 32 * j + b;
                    if (x < W) { atomicAdd(&gh[img[(size_t)y * W + x]], 1u); ++nsel; }
                }
                ++rank;
            }
        }
    }
    if (nsel) atomicAdd(&out_stats[(size_t)blockIdx.x * 4], nsel);
}

// ---------------------------------------------------------------- order statistics
struct IpbQJob {
    int hist;       // histogram index
    float q32;      // quantile = float32(p) / float32(100), computed by the host like numpy
    int pad0, pad1;
};
struct IpbQOut {
    int prev, next;        // the two order statistics (raw uint16 values); -1 if n == 0
    float gamma;           // numpy's interpolation weight
    float value;           // numpy percentile of the float32 copy of the raw sample
    unsigned long long n;  // sample size
};

// one CTA (256 threads) per quantile job; thread t owns bins [256 t, 256 t + 256)
__global__ void __launch_bounds__(256)
ipb_k_hist_quantiles(const unsigned* __restrict__ hist, const unsigned long long* __restrict__ stats,
                     const IpbQJob* __restrict__ qjobs, IpbQOut* __restrict__ out)
{
    const IpbQJob qj = qjobs[blockIdx.x];
    const unsigned* h = hist + (size_t)qj.hist * IPB_HIST_BINS;
    const unsigned long long n = stats[(size_t)qj.hist * 4];
    __shared__ unsigned long long part[256];
    __shared__ int res[2];
    const int t = threadIdx.x;
    unsigned long long mine = 0;
    for (int b = 0; b < 256; ++b) mine += h[t * 256 + b];
    part[t] = mine;
    if (t < 2) res[t] = -1;
    __syncthreads();
    if (t == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < 256; ++i) { unsigned long long c = part[i]; part[i] = acc; acc += c; }
    }
    __syncthreads();
    IpbQIdx qi;
    qi.prev = 0; qi.next = 0; qi.gamma = 0.f;
    if (n > 0) {
        qi = ipb_np_qidx_f32((long long)n, qj.q32);
        const unsigned long long lo = part[t], hi = lo + mine;
        const long long want[2] = {qi.prev, qi.next};
        for (int w = 0; w < 2; ++w) {
            const unsigned long long kk = (unsigned long long)want[w];
            if (kk >= lo && kk < hi) {
                unsigned long long acc = lo;
                for (int b = 0; b < 256; ++b) {
                    acc += h[t * 256 + b];
                    if (kk < acc) { res[w] = t * 256 + b; break; }
                }
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        IpbQOut o;
        o.prev = res[0]; o.next = res[1]; o.gamma = qi.gamma; o.n = n;
        o.value = (n > 0 && res[0] >= 0 && res[1] >= 0)
                      ? ipb_np_lerp_f32((float)res[0], (float)res[1], qi.gamma) : 0.0f;
        out[blockIdx.x] = o;
    }
}

// ================================================================ selection by sampling
// np.percentile needs two order statistics, not the whole distribution.  ipb_hist_select:
//   1. sample   exact histogram of a hashed 1/16 sample of each job's pixels (ipb_k_hist_u16,
//               sample = 1): 16x fewer shared-memory atomics, 1/8 of the DRAM sectors.
//   2. windows  per job, the value window that contains every wanted rank with overwhelming
//               probability: sample ranks r = q (n_s - 1) -/+ (6 sqrt(q (1-q) n_s) + 8).
//   3. tail     ONE read of each plane for all jobs that sample it: pixels below / above the
//               window are only counted, pixels inside go to a 4096-bin shared histogram.
//   4. fallback sparse patterns, tiny samples and wide windows take the exact full-range
//               histogram (ipb_k_hist_u16 restricted to those jobs).
//   5. select   exact ranks inside the window.  A wanted rank outside its window is reported
//               in *miss (the caller then repeats the step with full histograms), so the
//               result is exact in every case.

// one CTA (256 threads) per histogram job
__global__ void __launch_bounds__(256)
ipb_k_hist_windows(const unsigned* __restrict__ hs /* sample histograms */,
                   const unsigned long long* __restrict__ stats_s, const IpbHistJob* __restrict__ jobs,
                   const IpbQJob* __restrict__ qjobs, int n_q, IpbHistWin* __restrict__ win)
{
    const int job = blockIdx.x, t = threadIdx.x;
    const unsigned* h = hs + (size_t)job * IPB_HIST_BINS;
    const unsigned long long ns = stats_s[(size_t)job * 4];
    __shared__ unsigned long long part[256];
    __shared__ long long want[2];
    __shared__ int res[2];
    __shared__ int any_q;
    if (t == 0) {
        // rank range over every quantile wanted from this job
        long long lo = 0x7fffffffffffffffll, hi = -1;
        int any = 0;
        for (int i = 0; i < n_q; ++i) {
            if (qjobs[i].hist != job) continue;
            any = 1;
            if (ns == 0) continue;
            const double q = (double)qjobs[i].q32;
            const double r = q * (double)(ns - 1);
            const double d = 6.0 * sqrt(fmax(q * (1.0 - q), 0.0) * (double)ns) + 8.0;
            long long a = (long long)floor(r - d), b = (long long)ceil(r + d) + 1;
            if (a < 0) a = 0;
            if (b > (long long)ns - 1) b = (long long)ns - 1;
            lo = a < lo ? a : lo;
            hi = b > hi ? b : hi;
        }
        want[0] = lo; want[1] = hi; any_q = any;
        res[0] = res[1] = -1;
    }
    unsigned long long mine = 0;
    for (int b = 0; b < 256; ++b) mine += h[t * 256 + b];
    part[t] = mine;
    __syncthreads();
    if (t == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < 256; ++i) { unsigned long long c = part[i]; part[i] = acc; acc += c; }
    }
    __syncthreads();
    const int pattern = jobs[job].pattern;
    const bool sparse = pattern == IPB_PAT_STRIDE2D || pattern == IPB_PAT_MASKED_STRIDE;
    if (any_q && ns >= 256 && !sparse) {
        const unsigned long long lo = part[t], hi = lo + mine;
        for (int w = 0; w < 2; ++w) {
            const unsigned long long kk = (unsigned long long)want[w];
            if (kk >= lo && kk < hi) {
                unsigned long long acc = lo;
                for (int b = 0; b < 256; ++b) {
                    acc += h[t * 256 + b];
                    if (kk < acc) { res[w] = t * 256 + b; break; }
                }
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        IpbHistWin o;
        o.pad = 0;
        if (!any_q) { o.wlo = 0; o.whi = 0; o.mode = IPB_HSEL_NONE; }
        else if (ns < 256 || sparse || res[0] < 0 || res[1] < 0) { o.wlo = 0; o.whi = 0; o.mode = IPB_HSEL_FULL; }
        else {
            // a window that starts / ends at the sample's extreme rank is opened to the end of
            // the value range: the true extremes may lie beyond the sample's
            o.wlo = want[0] == 0 ? 0 : res[0];
            o.whi = want[1] >= (long long)ns - 1 ? IPB_HIST_BINS : res[1] + 1;
            o.mode = (o.whi - o.wlo <= IPB_HSEL_WIN) ? IPB_HSEL_WINDOWED : IPB_HSEL_FULL;
        }
        win[job] = o;
    }
}

// tail pass: grid (chunks, plane passes), 512 threads; shared = n_jobs windows of 4096 bins.
// cnt[job] = { below, inside, above, 0 }; moments (sum, sumsq of ALL pixels) -> stats[job][1..2],
// n_selected -> stats[job][0].
#define IPB_HSEL_THREADS 512
__global__ void __launch_bounds__(IPB_HSEL_THREADS)
ipb_k_hist_tail(const unsigned short* __restrict__ planes, int H, int W,
                const IpbPlanePass* __restrict__ passes, const IpbHistJob* __restrict__ jobs,
                const IpbHistWin* __restrict__ win, int rows_per_chunk,
                const unsigned* __restrict__ union_bits, int union_wpr,
                unsigned* __restrict__ hw /* [jobs][IPB_HSEL_WIN] */,
                unsigned long long* __restrict__ cnt, unsigned long long* __restrict__ stats)
{
    IPB_DYN_SMEM(unsigned, sh);
    const IpbPlanePass pp = passes[blockIdx.y];
    const int y_beg = (int)blockIdx.x * rows_per_chunk;
    int y_end = y_beg + rows_per_chunk;
    if (y_end > H) y_end = H;
    IpbHistJob jb[IPB_HSEL_MAXJ];
    IpbHistWin wn[IPB_HSEL_MAXJ];
    const unsigned* ub[IPB_HSEL_MAXJ];
    bool need = false;
#pragma unroll
    for (int u = 0; u < IPB_HSEL_MAXJ; ++u) {
        if (u < pp.n_jobs) {
            jb[u] = jobs[pp.job[u]];
            wn[u] = win[pp.job[u]];
            ub[u] = (jb[u].pattern == IPB_PAT_MASKED) ? union_bits + (size_t)jb[u].mask_frame * H * union_wpr : nullptr;
            need = need || wn[u].mode == IPB_HSEL_WINDOWED || jb[u].moments;
        } else { jb[u].pattern = -1; jb[u].moments = 0; jb[u].k = 1; wn[u].mode = IPB_HSEL_NONE; wn[u].wlo = wn[u].whi = 0; ub[u] = nullptr; }
    }
    if (!need || y_beg >= y_end) return;
    for (int b = threadIdx.x; b < pp.n_jobs * IPB_HSEL_WIN; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    const unsigned short* img = planes + (size_t)pp.plane * H * W;
    const unsigned sat_min = pp.sat_min > 0 ? (unsigned)pp.sat_min : 0xffffffffu;
    const unsigned short* img2 = (pp.sat_min > 0 && pp.excl_plane1 > 0) ? planes + (size_t)(pp.excl_plane1 - 1) * H * W : nullptr;
    unsigned below[IPB_HSEL_MAXJ], inside[IPB_HSEL_MAXJ], above[IPB_HSEL_MAXJ];
    unsigned long long s1 = 0, s2 = 0;
    bool moments = false;
#pragma unroll
    for (int u = 0; u < IPB_HSEL_MAXJ; ++u) { below[u] = inside[u] = above[u] = 0; moments = moments || jb[u].moments; }
    const bool vec_ok = ((W & 7) == 0) && ((((size_t)img) & 15) == 0) && ((((size_t)img2) & 15) == 0);
    const int step = vec_ok ? 8 : 1;
    const int upr = vec_ok ? (W >> 3) : W;                            // units (vectors or pixels) per row
    const long long nunits = (long long)(y_end - y_beg) * upr;
    for (long long i = threadIdx.x; i < nunits; i += blockDim.x) {
        const int y = y_beg + (int)(i / upr);
        const int x0 = ((int)(i % upr)) * step;
        unsigned w[4] = {0, 0, 0, 0}, w2[4] = {0, 0, 0, 0};
        if (vec_ok) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(img + (size_t)y * W + x0));
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
            if (img2) {
                const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(img2 + (size_t)y * W + x0));
                w2[0] = q2.x; w2[1] = q2.y; w2[2] = q2.z; w2[3] = q2.w;
            }
        } else {
            w[0] = img[(size_t)y * W + x0];
            if (img2) w2[0] = img2[(size_t)y * W + x0];
        }
        unsigned sel[IPB_HSEL_MAXJ];
#pragma unroll
        for (int u = 0; u < IPB_HSEL_MAXJ; ++u) {
            sel[u] = 0;
            if (wn[u].mode != IPB_HSEL_WINDOWED) continue;
            const int k = jb[u].k > 0 ? jb[u].k : 1;
            if (jb[u].pattern == IPB_PAT_FULL) sel[u] = 0xffu;
            else if (jb[u].pattern == IPB_PAT_STRIDE1D) {
                const long long flat = (long long)y * W + x0;
                const int first = (int)((k - (flat % k)) % k);
                for (int t = first; t < step; t += k) sel[u] |= 1u << t;
            } else if (jb[u].pattern == IPB_PAT_MASKED) {
                sel[u] = (ub[u][(size_t)y * union_wpr + (x0 >> 5)] >> (x0 & 31)) & 0xffu;
            }
            if (!vec_ok) sel[u] &= 1u;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (t >= step) break;
            const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
            const unsigned o = (t & 1) ? (w2[t >> 1] >> 16) : (w2[t >> 1] & 0xffffu);
            if (moments) { s1 += v; s2 += (unsigned long long)v * v; }
            const bool keep = v < sat_min && o < sat_min;
#pragma unroll
            for (int u = 0; u < IPB_HSEL_MAXJ; ++u) {
                if (((sel[u] >> t) & 1u) && keep) {
                    if (v < (unsigned)wn[u].wlo) ++below[u];
                    else if (v >= (unsigned)wn[u].whi) ++above[u];
                    else { ++inside[u]; atomicAdd(&sh[u * IPB_HSEL_WIN + (int)v - wn[u].wlo], 1u); }
                }
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < pp.n_jobs * IPB_HSEL_WIN; b += blockDim.x) {
        const unsigned c = sh[b];
        if (c) atomicAdd(&hw[(size_t)pp.job[b / IPB_HSEL_WIN] * IPB_HSEL_WIN + (b % IPB_HSEL_WIN)], c);
    }
    // block reduction: per job three counters (+ the two moment sums)
    __shared__ unsigned long long red[16];
    __shared__ unsigned long long acc[3 * IPB_HSEL_MAXJ + 2];
    if (threadIdx.x < 3 * IPB_HSEL_MAXJ + 2) acc[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < IPB_HSEL_MAXJ; ++u) {
        unsigned long long a = ipb_warp_sum((unsigned long long)below[u]);
        unsigned long long b = ipb_warp_sum((unsigned long long)inside[u]);
        unsigned long long c = ipb_warp_sum((unsigned long long)above[u]);
        if (lane == 0) { if (a) atomicAdd(&acc[3 * u], a); if (b) atomicAdd(&acc[3 * u + 1], b); if (c) atomicAdd(&acc[3 * u + 2], c); }
    }
    s1 = ipb_warp_sum(s1); s2 = ipb_warp_sum(s2);
    if (lane == 0) { if (s1) atomicAdd(&acc[3 * IPB_HSEL_MAXJ], s1); if (s2) atomicAdd(&acc[3 * IPB_HSEL_MAXJ + 1], s2); }
    (void)red;
    __syncthreads();
    if (threadIdx.x < pp.n_jobs) {
        const int u = threadIdx.x, j = pp.job[u];
        const unsigned long long a = acc[3 * u], b = acc[3 * u + 1], c = acc[3 * u + 2];
        if (a) atomicAdd(&cnt[(size_t)j * 4], a);
        if (b) atomicAdd(&cnt[(size_t)j * 4 + 1], b);
        if (c) atomicAdd(&cnt[(size_t)j * 4 + 2], c);
        if (a + b + c) atomicAdd(&stats[(size_t)j * 4], a + b + c);
        if (jobs[j].moments) {
            if (acc[3 * IPB_HSEL_MAXJ]) atomicAdd(&stats[(size_t)j * 4 + 1], acc[3 * IPB_HSEL_MAXJ]);
            if (acc[3 * IPB_HSEL_MAXJ + 1]) atomicAdd(&stats[(size_t)j * 4 + 2], acc[3 * IPB_HSEL_MAXJ + 1]);
        }
    }
}

// one CTA (256 threads) per quantile job
__global__ void __launch_bounds__(256)
ipb_k_hist_select_q(const IpbQJob* __restrict__ qjobs, const IpbHistWin* __restrict__ win,
                    const unsigned long long* __restrict__ cnt, const unsigned* __restrict__ hw,
                    const unsigned* __restrict__ hf, const unsigned long long* __restrict__ stats,
                    IpbQOut* __restrict__ out, unsigned* __restrict__ miss)
{
    const IpbQJob qj = qjobs[blockIdx.x];
    const IpbHistWin wn = win[qj.hist];
    const unsigned long long n = stats[(size_t)qj.hist * 4];
    const bool windowed = wn.mode == IPB_HSEL_WINDOWED;
    const unsigned* h = windowed ? hw + (size_t)qj.hist * IPB_HSEL_WIN : hf + (size_t)qj.hist * IPB_HIST_BINS;
    const int per = windowed ? IPB_HSEL_WIN / 256 : IPB_HIST_BINS / 256;
    const unsigned long long base = windowed ? cnt[(size_t)qj.hist * 4] : 0ull;
    const int v0 = windowed ? wn.wlo : 0;
    __shared__ unsigned long long part[256];
    __shared__ int res[2];
    const int t = threadIdx.x;
    unsigned long long mine = 0;
    for (int b = 0; b < per; ++b) mine += h[t * per + b];
    part[t] = mine;
    if (t < 2) res[t] = -1;
    __syncthreads();
    if (t == 0) {
        unsigned long long acc = base;
        for (int i = 0; i < 256; ++i) { unsigned long long c = part[i]; part[i] = acc; acc += c; }
    }
    __syncthreads();
    IpbQIdx qi;
    qi.prev = 0; qi.next = 0; qi.gamma = 0.f;
    if (n > 0) {
        qi = ipb_np_qidx_f32((long long)n, qj.q32);
        const unsigned long long lo = part[t], hi = lo + mine;
        const long long want[2] = {qi.prev, qi.next};
        for (int w = 0; w < 2; ++w) {
            const unsigned long long kk = (unsigned long long)want[w];
            if (kk >= lo && kk < hi) {
                unsigned long long acc = lo;
                for (int b = 0; b < per; ++b) {
                    acc += h[t * per + b];
                    if (kk < acc) { res[w] = v0 + t * per + b; break; }
                }
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        IpbQOut o;
        o.prev = res[0]; o.next = res[1]; o.gamma = qi.gamma; o.n = n;
        const bool ok = n > 0 && res[0] >= 0 && res[1] >= 0;
        o.value = ok ? ipb_np_lerp_f32((float)res[0], (float)res[1], qi.gamma) : 0.0f;
        if (n > 0 && !ok) atomicAdd(miss, 1u);          // the sample-derived window missed a rank
        out[blockIdx.x] = o;
    }
}
