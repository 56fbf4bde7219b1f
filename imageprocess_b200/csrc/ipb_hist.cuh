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
#include "ipb_scan.cuh"

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

// ---- plane passes: the jobs that read one plane, fused into one read (ipb_hist_planes, ipb_hist_select)
#define IPB_HSEL_MAXJ 4          // jobs fused into one pass over a plane
struct IpbHistWin { int wlo, whi, mode, pad; };     // window [wlo, whi); mode IPB_PQ_*
struct IpbPlanePass { int plane, excl_plane1, sat_min, n_jobs; int job[IPB_HSEL_MAXJ]; };

// out_stats[job] = { n_selected, sum_all, sumsq_all, reserved }
__global__ void __launch_bounds__(IPB_HIST_THREADS)
ipb_k_hist_u16(const unsigned short* __restrict__ planes, int H, int W,
               const IpbHistJob* __restrict__ jobs, int rows_per_chunk,
               const unsigned* __restrict__ union_bits, int union_wpr,
               unsigned* __restrict__ hist, unsigned long long* __restrict__ out_stats)
{
    IPB_DYN_SMEM(unsigned, sh);
    const IpbHistJob job = jobs[blockIdx.y];
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
    unsigned nsel32 = 0;                       // per-thread count of the vector path (<= 2^27 pixels per thread)
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
            // (row, vector) walk without a division per iteration
            const int dy = (int)blockDim.x / vpr, dx = (int)blockDim.x % vpr;
            int y = y_beg + (int)threadIdx.x / vpr, xv = (int)threadIdx.x % vpr;
            // one 8-pixel group: selection mask, then counting
            auto process = [&](int y, int x0, const uint4& q, const uint4& q2) {
                const unsigned w[4] = {q.x, q.y, q.z, q.w};
                const unsigned w2[4] = {q2.x, q2.y, q2.z, q2.w};
                unsigned sel = 0;                                     // bit i: pixel x0+i selected
                if (job.pattern == IPB_PAT_FULL) sel = 0xffu;
                else if (job.pattern == IPB_PAT_STRIDE1D) {
                    const unsigned long long flat = (unsigned long long)y * W + x0;
                    const unsigned fm = flat < 0xffffffffull ? (unsigned)flat % (unsigned)k : (unsigned)(flat % (unsigned long long)k);
                    int first = (int)(((unsigned)k - fm) % (unsigned)k);
                    for (int t = first; t < 8; t += k) sel |= 1u << t;
                } else if (job.pattern == IPB_PAT_STRIDE2D) {
                    if ((y % k) == 0) {
                        int first = (k - (x0 % k)) % k;
                        for (int t = first; t < 8; t += k) sel |= 1u << t;
                    }
                } else if (job.pattern == IPB_PAT_MASKED) {
                    sel = (ubits[(size_t)y * union_wpr + (x0 >> 5)] >> (x0 & 31)) & 0xffu;
                }
                if (job.moments) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned a = w[j] & 0xffffu, b = w[j] >> 16;
                        s1 += a + b;
                        s2 += (unsigned long long)a * a + (unsigned long long)b * b;
                    }
                }
                if (img2 || job.sat_min > 0) {                        // saturation filter (rare)
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                        const unsigned o = (t & 1) ? (w2[t >> 1] >> 16) : (w2[t >> 1] & 0xffffu);
                        if (!(v < sat_min && o < sat_min)) sel &= ~(1u << t);
                    }
                }
                nsel32 += (unsigned)__popc(sel);
                if (sel == 0xffu) {                                   // the common case: no tests per pixel
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                        if (v < IPB_HIST_WIN) atomicAdd(&sh[v], 1u); else atomicAdd(&gh[v], 1u);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                        if ((sel >> t) & 1u) { if (v < IPB_HIST_WIN) atomicAdd(&sh[v], 1u); else atomicAdd(&gh[v], 1u); }
                    }
                }
            };
            // 4 groups per trip: their (up to 8) 128-bit loads are issued before any counting,
            // so every thread keeps several loads in flight
            while (y < y_end) {
                int ys[4], xs[4];
                bool ok[4];
                uint4 q[4], q2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ys[u] = y; xs[u] = xv << 3;
                    ok[u] = y < y_end;
                    if (ok[u] && job.pattern == IPB_PAT_STRIDE2D && !job.moments && (y % k) != 0) ok[u] = false;
                    q[u] = make_uint4(0, 0, 0, 0); q2[u] = make_uint4(0, 0, 0, 0);
                    if (ok[u]) {
                        q[u] = __ldg(reinterpret_cast<const uint4*>(img + (size_t)y * W + xs[u]));
                        if (img2) q2[u] = __ldg(reinterpret_cast<const uint4*>(img2 + (size_t)y * W + xs[u]));
                    }
                    xv += dx; y += dy;
                    if (xv >= vpr) { xv -= vpr; ++y; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) if (ok[u]) process(ys[u], xs[u], q[u], q2[u]);
            }
        } else {
            const int dy = (int)blockDim.x / W, dx = (int)blockDim.x % W;
            int y = y_beg + (int)threadIdx.x / W, x = (int)threadIdx.x % W;
            for (; y < y_end; x += dx, y += dy, y += (x >= W) ? 1 : 0, x -= (x >= W) ? W : 0) {
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
    nsel += nsel32;
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

// one CTA (256 threads) per quantile job; warps scan contiguous bands of the histogram with
// coalesced reads (ipb_locate_ranks)
__global__ void __launch_bounds__(1024)
ipb_k_hist_quantiles(const unsigned* __restrict__ hist, const unsigned long long* __restrict__ stats,
                     const IpbQJob* __restrict__ qjobs, IpbQOut* __restrict__ out)
{
    const IpbQJob qj = qjobs[blockIdx.x];
    const unsigned* h = hist + (size_t)qj.hist * IPB_HIST_BINS;
    const unsigned long long n = stats[(size_t)qj.hist * 4];
    __shared__ unsigned long long red_u[32];
    __shared__ int res[2];
    if (threadIdx.x < 2) res[threadIdx.x] = -1;
    IpbQIdx qi;
    qi.prev = 0; qi.next = 0; qi.gamma = 0.f;
    if (n > 0) {                                              // block-uniform
        qi = ipb_np_qidx_f32((long long)n, qj.q32);
        const unsigned long long want[2] = {(unsigned long long)qi.prev, (unsigned long long)qi.next};
        ipb_locate_ranks(IPB_HIST_BINS / 32, want, 2, red_u, [&](unsigned i) { return h[i]; },
                         [&](int r, unsigned i, unsigned) { res[r] = (int)i; });
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        IpbQOut o;
        o.prev = res[0]; o.next = res[1]; o.gamma = qi.gamma; o.n = n;
        o.value = (n > 0 && res[0] >= 0 && res[1] >= 0)
                      ? ipb_np_lerp_f32((float)res[0], (float)res[1], qi.gamma) : 0.0f;
        out[blockIdx.x] = o;
    }
}

#include "ipb_pq.cuh"

// ================================================================ fused full histograms per plane
// ipb_k_hist_planes: the exact 65 536-bin histograms of EVERY job that samples a plane from ONE
// read of that plane (ipb_hist_planes; passes built like ipb_hist_select's).  A time-lapse frame
// typically has a full / masked job (FRET background + epsilon), a flat-stride job (Fluor_INT
// background on vals[::k]) and the FA job (moments of all pixels + a sparse [::k, ::k] sample) on
// the same channel: three reads of the plane become one, and the per-8-pixel bookkeeping is paid
// once.  Dense jobs share the shared-memory window (IPB_HIST_WIN / n_dense bins each, brighter
// values go to L2 atomics); sparse [::k, ::k] jobs count straight into global memory.  A
// flat-stride job next to a FULL job selects a subset of its pixels: those pixels are counted
// once, in the stride job's window, and that window is added to both histograms at the end --
// the kernel is bound by shared-memory atomics (~2 per clock per SM), so every atomic saved counts.
// n % k for n < 2^32 / k with magic = ceil(2^32 / k): one multiply-high instead of a division
__device__ __forceinline__ unsigned ipb_fastmod(unsigned n, unsigned k, unsigned magic) {
    return n - __umulhi(n, magic) * k;
}

__global__ void __launch_bounds__(IPB_HIST_THREADS)
ipb_k_hist_planes(const unsigned short* __restrict__ planes, int H, int W,
                  const IpbPlanePass* __restrict__ passes, const IpbHistJob* __restrict__ jobs, int rows_per_chunk,
                  const unsigned* __restrict__ union_bits, int union_wpr,
                  unsigned* __restrict__ hist, unsigned long long* __restrict__ out_stats)
{
    IPB_DYN_SMEM(unsigned, sh);
    const IpbPlanePass pp = passes[blockIdx.y];
    const int y_beg = (int)blockIdx.x * rows_per_chunk;
    int y_end = y_beg + rows_per_chunk;
    if (y_end > H) y_end = H;
    if (y_beg >= y_end) return;
    // The jobs of a pass are sorted into fixed roles held in scalars (no indexed register arrays
    // in the pixel loop):  F = a FULL job, S = a flat-stride job (subset of F when both exist: its
    // pixels are counted once, in S's window, which is added to both histograms at the flush),
    // M = a masked job, P = a sparse [::k, ::k] job.  A second job of a role takes the generic slot G.
    int jF = -1, jS = -1, jM = -1, jP = -1, jG = -1, patG = -1;
    unsigned kS = 1, kP = 1, kG = 1;
    bool moments = false;
    for (int u = 0; u < pp.n_jobs; ++u) {
        const IpbHistJob j = jobs[pp.job[u]];
        moments = moments || j.moments != 0;
        const unsigned k = j.k > 0 ? (unsigned)j.k : 1u;
        if (j.pattern == IPB_PAT_FULL && jF < 0) jF = u;
        else if (j.pattern == IPB_PAT_STRIDE1D && jS < 0) { jS = u; kS = k; }
        else if (j.pattern == IPB_PAT_MASKED && jM < 0) jM = u;
        else if (j.pattern == IPB_PAT_STRIDE2D && jP < 0) { jP = u; kP = k; }
        else if (j.pattern != IPB_PAT_MASKED_STRIDE && jG < 0) { jG = u; patG = j.pattern; kG = k; }
        // IPB_PAT_MASKED_STRIDE has its own kernel; a third job of one role cannot occur (<= 4 jobs, 5 roles)
    }
    const int n_dense = (jF >= 0) + (jS >= 0) + (jM >= 0) + (jG >= 0 && patG != IPB_PAT_STRIDE2D);
    const unsigned win = n_dense ? (unsigned)(IPB_HIST_WIN / n_dense) : 0u;
    int d = 0;
    const int bF = jF >= 0 ? (d++) * (int)win : 0, bS = jS >= 0 ? (d++) * (int)win : 0, bM = jM >= 0 ? (d++) * (int)win : 0;
    const int bG = (jG >= 0 && patG != IPB_PAT_STRIDE2D) ? (d++) * (int)win : -1;
    auto ghist = [&](int slot) { return hist + (size_t)pp.job[slot < 0 ? 0 : slot] * IPB_HIST_BINS; };
    unsigned* gF = ghist(jF); unsigned* gS = ghist(jS); unsigned* gM = ghist(jM); unsigned* gP = ghist(jP); unsigned* gG = ghist(jG);
    const unsigned* ubM = jM >= 0 ? union_bits + (size_t)jobs[pp.job[jM]].mask_frame * H * union_wpr : nullptr;
    const unsigned* ubG = (jG >= 0 && patG == IPB_PAT_MASKED) ? union_bits + (size_t)jobs[pp.job[jG]].mask_frame * H * union_wpr : nullptr;
    unsigned p16S = 0, p16P = 0, p16G = 0;
    for (unsigned t = 0; t < 16; t += kS) p16S |= 1u << t;
    for (unsigned t = 0; t < 16; t += kP) p16P |= 1u << t;
    for (unsigned t = 0; t < 16; t += kG) p16G |= 1u << t;
    const unsigned mgP = (unsigned)((0x100000000ull + kP - 1) / kP), mgG = (unsigned)((0x100000000ull + kG - 1) / kG);
    const bool pairFS = jF >= 0 && jS >= 0;
    for (int b = threadIdx.x; b < n_dense * (int)win; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    const unsigned short* img = planes + (size_t)pp.plane * H * W;
    const unsigned sat_min = pp.sat_min > 0 ? (unsigned)pp.sat_min : 0xffffffffu;
    const unsigned short* img2 = (pp.sat_min > 0 && pp.excl_plane1 > 0) ? planes + (size_t)(pp.excl_plane1 - 1) * H * W : nullptr;
    unsigned nF = 0, nS = 0, nM = 0, nP = 0, nG = 0;
    unsigned long long s1 = 0, s2 = 0;
    const bool vec_ok = ((W & 7) == 0) && ((((size_t)img) & 15) == 0) && ((((size_t)img2) & 15) == 0);
    const int step = vec_ok ? 8 : 1;
    const int upr = vec_ok ? (W >> 3) : W;
    const int dy = (int)blockDim.x / upr, dx = (int)blockDim.x % upr;
    int y = y_beg + (int)threadIdx.x / upr, xu = (int)threadIdx.x % upr;

    // flat-stride selection of one unit: bits of the pixels with flat index % k == 0
    auto stride1d = [&](int y, int x0, unsigned k, unsigned p16) -> unsigned {
        const unsigned long long flat = (unsigned long long)y * W + x0;
        const unsigned fm = (k & (k - 1u)) == 0u ? (unsigned)flat & (k - 1u)
                            : (flat < 0xffffffffull ? (unsigned)flat % k : (unsigned)(flat % (unsigned long long)k));
        const unsigned first = fm ? k - fm : 0u;
        return first < 8u ? (p16 << first) & 0xffu : 0u;
    };
    auto stride2d = [&](int y, int x0, unsigned k, unsigned p16, unsigned magic) -> unsigned {
        if (ipb_fastmod((unsigned)y, k, magic) != 0u) return 0u;
        const unsigned xm = ipb_fastmod((unsigned)x0, k, magic);
        const unsigned first = xm ? k - xm : 0u;
        return first < 8u ? (p16 << first) & 0xffu : 0u;
    };
    // count the selected pixels of one unit into a dense job's window / global histogram
    auto count_dense = [&](unsigned sel, int base, unsigned* g, const unsigned (&w)[4]) {
        if (sel == 0xffu) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                if (v < win) atomicAdd(&sh[base + v], 1u); else atomicAdd(&g[v], 1u);
            }
        } else {
            while (sel) {
                const int t = __ffs((int)sel) - 1;
                sel &= sel - 1;
                const unsigned word = t < 4 ? (t < 2 ? w[0] : w[1]) : (t < 6 ? w[2] : w[3]);
                const unsigned v = (t & 1) ? (word >> 16) : (word & 0xffffu);
                if (v < win) atomicAdd(&sh[base + v], 1u); else atomicAdd(&g[v], 1u);
            }
        }
    };
    // one unit (8-pixel group, or one pixel when the row length is not a multiple of 8)
    auto process = [&](int y, int x0, const unsigned (&w)[4], const unsigned (&w2)[4]) {
        if (moments) {
            if (step == 8) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned a = w[j] & 0xffffu, b = w[j] >> 16;
                    s1 += a + b;
                    s2 += (unsigned long long)a * a + (unsigned long long)b * b;
                }
            } else { const unsigned a = w[0] & 0xffffu; s1 += a; s2 += (unsigned long long)a * a; }
        }
        unsigned keep = step == 8 ? 0xffu : 1u;
        if (pp.sat_min > 0) {
            keep = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (t >= step) break;
                const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                const unsigned o = (t & 1) ? (w2[t >> 1] >> 16) : (w2[t >> 1] & 0xffffu);
                keep |= (v < sat_min && o < sat_min) ? (1u << t) : 0u;
            }
        }
        const unsigned selS = jS >= 0 ? stride1d(y, x0, kS, p16S) & keep : 0u;
        if (pairFS) {                                  // every kept pixel once; S's pixels in S's window
            nF += (unsigned)__popc(keep); nS += (unsigned)__popc(selS);
            if (keep == 0xffu) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const unsigned v = (t & 1) ? (w[t >> 1] >> 16) : (w[t >> 1] & 0xffffu);
                    const bool sub = (selS >> t) & 1u;
                    if (v < win) atomicAdd(&sh[(sub ? bS : bF) + (int)v], 1u);
                    else { atomicAdd(&gF[v], 1u); if (sub) atomicAdd(&gS[v], 1u); }
                }
            } else {
                unsigned todo = keep;
                while (todo) {
                    const int t = __ffs((int)todo) - 1;
                    todo &= todo - 1;
                    const unsigned word = t < 4 ? (t < 2 ? w[0] : w[1]) : (t < 6 ? w[2] : w[3]);
                    const unsigned v = (t & 1) ? (word >> 16) : (word & 0xffffu);
                    const bool sub = (selS >> t) & 1u;
                    if (v < win) atomicAdd(&sh[(sub ? bS : bF) + (int)v], 1u);
                    else { atomicAdd(&gF[v], 1u); if (sub) atomicAdd(&gS[v], 1u); }
                }
            }
        } else {
            if (jF >= 0) { nF += (unsigned)__popc(keep); count_dense(keep, bF, gF, w); }
            if (jS >= 0 && selS) { nS += (unsigned)__popc(selS); count_dense(selS, bS, gS, w); }
        }
        if (jM >= 0) {
            const unsigned sel = ((ubM[(size_t)y * union_wpr + (x0 >> 5)] >> (x0 & 31)) & 0xffu) & keep;
            if (sel) { nM += (unsigned)__popc(sel); count_dense(sel, bM, gM, w); }
        }
        if (jP >= 0) {
            unsigned sel = stride2d(y, x0, kP, p16P, mgP) & keep;
            nP += (unsigned)__popc(sel);
            while (sel) {                              // sparse: straight to the global histogram
                const int t = __ffs((int)sel) - 1;
                sel &= sel - 1;
                const unsigned word = t < 4 ? (t < 2 ? w[0] : w[1]) : (t < 6 ? w[2] : w[3]);
                atomicAdd(&gP[(t & 1) ? (word >> 16) : (word & 0xffffu)], 1u);
            }
        }
        if (jG >= 0) {
            unsigned sel = 0;
            if (patG == IPB_PAT_FULL) sel = 0xffu;
            else if (patG == IPB_PAT_STRIDE1D) sel = stride1d(y, x0, kG, p16G);
            else if (patG == IPB_PAT_STRIDE2D) sel = stride2d(y, x0, kG, p16G, mgG);
            else if (patG == IPB_PAT_MASKED) sel = (ubG[(size_t)y * union_wpr + (x0 >> 5)] >> (x0 & 31)) & 0xffu;
            sel &= keep;
            if (sel) {
                nG += (unsigned)__popc(sel);
                if (bG >= 0) count_dense(sel, bG, gG, w);
                else while (sel) {
                    const int t = __ffs((int)sel) - 1;
                    sel &= sel - 1;
                    const unsigned word = t < 4 ? (t < 2 ? w[0] : w[1]) : (t < 6 ? w[2] : w[3]);
                    atomicAdd(&gG[(t & 1) ? (word >> 16) : (word & 0xffffu)], 1u);
                }
            }
        }
    };
    if (vec_ok) {
        while (y < y_end) {
            int ys[4], xs[4];
            bool ok[4];
            uint4 q[4], q2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ys[u] = y; xs[u] = xu << 3;
                ok[u] = y < y_end;
                q[u] = make_uint4(0, 0, 0, 0); q2[u] = make_uint4(0, 0, 0, 0);
                if (ok[u]) {
                    q[u] = __ldg(reinterpret_cast<const uint4*>(img + (size_t)y * W + xs[u]));
                    if (img2) q2[u] = __ldg(reinterpret_cast<const uint4*>(img2 + (size_t)y * W + xs[u]));
                }
                xu += dx; y += dy;
                if (xu >= upr) { xu -= upr; ++y; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!ok[u]) continue;
                const unsigned w[4] = {q[u].x, q[u].y, q[u].z, q[u].w}, w2[4] = {q2[u].x, q2[u].y, q2[u].z, q2[u].w};
                process(ys[u], xs[u], w, w2);
            }
        }
    } else {
        for (; y < y_end; xu += dx, y += dy, y += (xu >= upr) ? 1 : 0, xu -= (xu >= upr) ? upr : 0) {
            const unsigned w[4] = {img[(size_t)y * W + xu], 0u, 0u, 0u};
            const unsigned w2[4] = {img2 ? (unsigned)img2[(size_t)y * W + xu] : 0u, 0u, 0u, 0u};
            process(y, xu, w, w2);
        }
    }
    __syncthreads();
    // flush the windows (S's window also belongs to F when they are paired)
    for (int b = threadIdx.x; b < n_dense * (int)win; b += blockDim.x) {
        const unsigned c = sh[b];
        if (!c) continue;
        const int dd = b / (int)win, v = b - dd * (int)win;
        const int base = dd * (int)win;
        if (jF >= 0 && base == bF) atomicAdd(&gF[v], c);
        else if (jS >= 0 && base == bS) { atomicAdd(&gS[v], c); if (pairFS) atomicAdd(&gF[v], c); }
        else if (jM >= 0 && base == bM) atomicAdd(&gM[v], c);
        else atomicAdd(&gG[v], c);
    }
    // per job: selected-pixel count; moments of the plane (all pixels) to every job that asked
    __shared__ unsigned long long acc[7];
    if (threadIdx.x < 7) acc[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    {
        const unsigned long long a0 = ipb_warp_sum((unsigned long long)nF), a1 = ipb_warp_sum((unsigned long long)nS),
                                 a2 = ipb_warp_sum((unsigned long long)nM), a3 = ipb_warp_sum((unsigned long long)nP),
                                 a4 = ipb_warp_sum((unsigned long long)nG);
        s1 = ipb_warp_sum(s1); s2 = ipb_warp_sum(s2);
        if (lane == 0) {
            if (a0) atomicAdd(&acc[0], a0);
            if (a1) atomicAdd(&acc[1], a1);
            if (a2) atomicAdd(&acc[2], a2);
            if (a3) atomicAdd(&acc[3], a3);
            if (a4) atomicAdd(&acc[4], a4);
            if (s1) atomicAdd(&acc[5], s1);
            if (s2) atomicAdd(&acc[6], s2);
        }
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        const int slot = threadIdx.x == 0 ? jF : threadIdx.x == 1 ? jS : threadIdx.x == 2 ? jM : threadIdx.x == 3 ? jP : jG;
        if (slot >= 0 && acc[threadIdx.x]) atomicAdd(&out_stats[(size_t)pp.job[slot] * 4], acc[threadIdx.x]);
    }
    if (threadIdx.x < pp.n_jobs && jobs[pp.job[threadIdx.x]].moments) {
        const int j = pp.job[threadIdx.x];
        if (acc[5]) atomicAdd(&out_stats[(size_t)j * 4 + 1], acc[5]);
        if (acc[6]) atomicAdd(&out_stats[(size_t)j * 4 + 2], acc[6]);
    }
}
