// ipb_pq.cuh -- plane percentiles by sampled windows (ipb_hist_select).  Included by ipb_hist.cuh.
//
// np.percentile needs two order statistics, not the distribution, and a shared-memory atomic
// costs ~16x a shared-memory load (B200: ~2 per clock per SM), so the exact full histogram of
// a plane is bound by its one atomic per pixel, not by HBM.  Instead:
//   1. sample   ONE CTA per plane pass histograms a stratified sample of 8-pixel units (one
//               hashed unit out of every stratum of consecutive units, <= 65 536 pixels) and
//               derives, for all quantiles wanted from the pass's dense jobs, ONE value window
//               [wlo, whi) that holds every wanted rank with overwhelming probability (sample
//               ranks -/+ 12 sigma: 6 sigma widened 2x for the 8-pixel clusters), and a second
//               window for the sparse [::k, ::k] job (same sample, widened by that job's own
//               sampling noise).
//   2. count    one read of each plane: a unit whose packed minimum is >= whi needs nothing more
//               (the usual case for a low background percentile).  Units with a pixel below whi
//               are parked in a per-warp shared-memory queue and handled 32 at a time with every
//               lane busy: pixels below wlo are counted in registers, pixels inside the window go
//               to a 2048-bin shared-memory histogram.  The plane's integer moments (FA global
//               statistics) and the sparse job's pixels (own window) ride along.
//   3. select   per quantile: exact rank inside the window.  A wanted rank outside its window,
//               or a pass the fast path cannot serve, raises *miss; the caller then repeats the
//               step with full histograms (ipb_hist_planes), so results are exact in every case.
// Roles of a pass: F = a FULL job, S = a flat-stride job vals[::k] with k in {2, 4, 8} and
// W % 8 == 0 (its pixels sit at fixed positions of every 8-pixel unit, and are a subset of F's),
// P = a sparse [::k, ::k] job.  Anything else (masks, saturation filter, two jobs of a role) is
// not served here.
#pragma once

#define IPB_PQ_WIN 2048            // window bins per job
#define IPB_PQ_SBINS 32768         // sample histogram bins (values clipped to the last one)
#define IPB_PQ_SUNITS 8192         // sampled 8-pixel units per plane
#define IPB_PQ_THREADS 256         // count kernel
#define IPB_PQ_OK 0
#define IPB_PQ_FALLBACK 1          // the pass (or this job's window) needs the full-histogram path
#define IPB_PQ_IDLE 3              // no quantile wanted

struct IpbPqRoles {
    int jF, jS, jP;                // job indices (global) or -1
    unsigned kS, kP, pS;           // strides; pS = bits of a unit's pixels that belong to S
    int moments;                   // some job of the pass wants the plane's sum / sum of squares
    int ok;                        // the fast path can serve this pass
};

__device__ __forceinline__ IpbPqRoles ipb_pq_roles(const IpbPlanePass& pp, const IpbHistJob* __restrict__ jobs, int W) {
    IpbPqRoles r;
    r.jF = r.jS = r.jP = -1; r.kS = 1; r.kP = 1; r.pS = 0; r.moments = 0;
    r.ok = (pp.sat_min <= 0) && ((W & 7) == 0);
    for (int u = 0; u < pp.n_jobs; ++u) {
        const IpbHistJob j = jobs[pp.job[u]];
        r.moments |= j.moments != 0;
        if (j.sat_min > 0 || j.excl_plane1 > 0) r.ok = 0;
        if (j.pattern == IPB_PAT_FULL && r.jF < 0) r.jF = pp.job[u];
        else if (j.pattern == IPB_PAT_STRIDE1D && r.jS < 0 && (j.k == 2 || j.k == 4 || j.k == 8)) { r.jS = pp.job[u]; r.kS = (unsigned)j.k; }
        else if (j.pattern == IPB_PAT_STRIDE2D && r.jP < 0 && j.k >= 1) { r.jP = pp.job[u]; r.kP = (unsigned)j.k; }
        else r.ok = 0;
    }
    if (r.jS >= 0) for (unsigned t = 0; t < 8; t += r.kS) r.pS |= 1u << t;
    return r;
}

__device__ __forceinline__ unsigned ipb_pq_px(const uint4& q, int t) {
    const unsigned w = t < 4 ? (t < 2 ? q.x : q.y) : (t < 6 ? q.z : q.w);
    return (t & 1) ? (w >> 16) : (w & 0xffffu);
}

// smallest of the 8 pixels of a unit
__device__ __forceinline__ unsigned ipb_pq_min8(const uint4& q) {
#ifdef IPB_EMULATE
    unsigned m = 0xffffu;
    for (int t = 0; t < 8; ++t) { const unsigned v = ipb_pq_px(q, t); m = v < m ? v : m; }
    return m;
#else
    const unsigned m = __vminu2(__vminu2(q.x, q.y), __vminu2(q.z, q.w));
    return min(m & 0xffffu, m >> 16);
#endif
}

// ---------------------------------------------------------------- 1. sample -> windows
// grid (n_passes), 1024 threads, IPB_PQ_SBINS words of dynamic shared memory.  Also initialises
// the per-job outputs of the pass: stats = {n, 0, 0, 0}, cnt = 0, window histogram = 0.
__global__ void __launch_bounds__(1024)
ipb_k_pq_sample(const unsigned short* __restrict__ planes, int H, int W,
                const IpbPlanePass* __restrict__ passes, const IpbHistJob* __restrict__ jobs,
                const IpbQJob* __restrict__ qjobs, int n_q,
                unsigned* __restrict__ hist_win, IpbHistWin* __restrict__ win,
                unsigned long long* __restrict__ cnt, unsigned long long* __restrict__ stats)
{
    IPB_DYN_SMEM(unsigned, sh);
    __shared__ unsigned long long scan32[32];
    __shared__ int want_i[6];               // sample rank range over the quantiles of F: [0], [1]; of S: [2], [3]; of P: [4], [5]
    __shared__ int any_q[3];
    __shared__ int res[6];
    const int tid = threadIdx.x;
    const IpbPlanePass pp = passes[blockIdx.x];
    const IpbPqRoles r = ipb_pq_roles(pp, jobs, W);
    const unsigned long long npx = (unsigned long long)H * (unsigned long long)W;
    const unsigned long long nP = (unsigned long long)((H + (int)r.kP - 1) / (int)r.kP) * (unsigned long long)((W + (int)r.kP - 1) / (int)r.kP);

    // ---- per-job outputs
    for (int u = 0; u < pp.n_jobs; ++u) {
        const int j = pp.job[u];
        unsigned* hw = hist_win + (size_t)j * IPB_PQ_WIN;
        for (int b = tid; b < IPB_PQ_WIN; b += blockDim.x) hw[b] = 0u;
        if (tid == 0) {
            unsigned long long n = 0;
            if (j == r.jF) n = npx;
            else if (j == r.jS) n = (npx + r.kS - 1) / r.kS;
            else if (j == r.jP) n = nP;
            stats[(size_t)j * 4] = n; stats[(size_t)j * 4 + 1] = 0; stats[(size_t)j * 4 + 2] = 0; stats[(size_t)j * 4 + 3] = 0;
            cnt[j] = 0;
        }
    }
    if (!r.ok) {                                                 // block-uniform
        if (tid < pp.n_jobs) {
            IpbHistWin o; o.wlo = 0; o.whi = 0; o.pad = 0; o.mode = IPB_PQ_FALLBACK;
            win[pp.job[tid]] = o;
        }
        return;
    }

    // ---- sample histogram: low half = pixels outside S, high half = S's pixels
    for (int b = tid; b < IPB_PQ_SBINS; b += blockDim.x) sh[b] = 0u;
    if (tid < 3) { want_i[2 * tid] = 0x7fffffff; want_i[2 * tid + 1] = -1; any_q[tid] = 0; res[2 * tid] = res[2 * tid + 1] = -1; }
    __syncthreads();
    const unsigned long long U = npx >> 3;                       // units of the plane (W % 8 == 0)
    const unsigned nsu = U < (unsigned long long)IPB_PQ_SUNITS ? (unsigned)U : (unsigned)IPB_PQ_SUNITS;
    const unsigned stratum = nsu ? (unsigned)(U / nsu) : 1u;
    const uint4* img = reinterpret_cast<const uint4*>(planes + (size_t)pp.plane * H * W);
    for (unsigned i0 = 0; i0 < nsu; i0 += 4u * blockDim.x) {
        uint4 q[4];
        bool ok[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const unsigned i = i0 + (unsigned)g * blockDim.x + (unsigned)tid;
            ok[g] = i < nsu;
            q[g] = make_uint4(0, 0, 0, 0);
            if (ok[g]) {
                unsigned h = i * 0x9E3779B1u ^ ((unsigned)pp.plane + 1u) * 0x85EBCA77u;
                h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
                const unsigned long long unit = (unsigned long long)i * stratum + __umulhi(h, stratum);
                q[g] = __ldg(img + unit);
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (!ok[g]) continue;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                unsigned v = ipb_pq_px(q[g], t);
                v = v < (unsigned)(IPB_PQ_SBINS - 1) ? v : (unsigned)(IPB_PQ_SBINS - 1);
                atomicAdd(&sh[v], ((r.pS >> t) & 1u) ? 0x10000u : 1u);
            }
        }
    }
    __syncthreads();
    const long long nsS = (long long)nsu * __popc(r.pS);
    const long long nsA = (long long)nsu * 8;                    // all sampled pixels: F's sample, and P's stand-in

    // ---- sample rank range over every quantile wanted from F / S / P
    for (int i = tid; i < n_q; i += blockDim.x) {
        const int hj = qjobs[i].hist;
        int which = -1;
        if (hj == r.jF) which = 0; else if (hj == r.jS) which = 1; else if (hj == r.jP) which = 2;
        if (which < 0) continue;
        const long long ns = which == 1 ? nsS : nsA;
        atomicOr(&any_q[which], 1);
        if (ns < 2) continue;
        const double q = (double)qjobs[i].q32;
        const double rr = q * (double)(ns - 1);
        const double pq = fmax(q * (1.0 - q), 0.0);
        double d = 12.0 * sqrt(pq * (double)ns) + 16.0;
        // P is read off the all-pixel sample: add 6 sigma of P's own quantile noise
        if (which == 2) d += 6.0 * sqrt(pq / (double)(nP > 0 ? nP : 1)) * (double)ns;
        double a = floor(rr - d), b = ceil(rr + d) + 1.0;
        if (a < 0.0) a = 0.0;
        if (b > (double)(ns - 1)) b = (double)(ns - 1);
        atomicMin(&want_i[2 * which], (int)a);
        atomicMax(&want_i[2 * which + 1], (int)b);
    }
    __syncthreads();
    auto all_px = [&](unsigned i) { const unsigned w = sh[i]; return (w & 0xffffu) + (w >> 16); };
    if (any_q[0] && nsA >= 2) {                                  // block-uniform
        const unsigned long long wv[2] = {(unsigned long long)want_i[0], (unsigned long long)want_i[1]};
        ipb_locate_ranks_smem(IPB_PQ_SBINS, IPB_PQ_SBINS / 1024, wv, 2, scan32, all_px,
                              [&](int k, unsigned i, unsigned) { res[k] = (int)i; });
    }
    if (any_q[1] && nsS >= 2) {
        const unsigned long long wv[2] = {(unsigned long long)want_i[2], (unsigned long long)want_i[3]};
        ipb_locate_ranks_smem(IPB_PQ_SBINS, IPB_PQ_SBINS / 1024, wv, 2, scan32,
                              [&](unsigned i) { return sh[i] >> 16; },
                              [&](int k, unsigned i, unsigned) { res[2 + k] = (int)i; });
    }
    if (any_q[2] && nsA >= 2) {
        const unsigned long long wv[2] = {(unsigned long long)want_i[4], (unsigned long long)want_i[5]};
        ipb_locate_ranks_smem(IPB_PQ_SBINS, IPB_PQ_SBINS / 1024, wv, 2, scan32, all_px,
                              [&](int k, unsigned i, unsigned) { res[4 + k] = (int)i; });
    }
    __syncthreads();
    if (tid == 0) {
        // a window runs from the lowest wanted sample rank's value (0 when that rank is the
        // sample's first: the true minimum may lie below the sample's) to the highest one's.  A
        // highest rank at the sample's end, a clipped value or a sample too small sends the jobs to
        // the full-histogram path; a window wider than IPB_PQ_WIN bins keeps its upper part (a rank
        // below it is then reported as a miss).
        auto window = [&](int w0, int w1) {
            int lo = 0x7fffffff, hi = -1;
            bool bad = false, any = false;
            for (int w = w0; w <= w1; ++w) {
                const long long ns = w == 1 ? nsS : nsA;
                if (!any_q[w] || (w == 0 && r.jF < 0) || (w == 1 && r.jS < 0) || (w == 2 && r.jP < 0)) continue;
                any = true;
                if (ns < 64 || res[2 * w] < 0 || res[2 * w + 1] < 0 || want_i[2 * w + 1] >= (int)(ns - 1)) { bad = true; continue; }
                const int l = want_i[2 * w] == 0 ? 0 : res[2 * w];
                lo = l < lo ? l : lo;
                hi = res[2 * w + 1] > hi ? res[2 * w + 1] : hi;
            }
            IpbHistWin o; o.pad = 0; o.wlo = 0; o.whi = 0;
            if (!any) o.mode = IPB_PQ_IDLE;
            else if (bad || hi >= IPB_PQ_SBINS - 1) o.mode = IPB_PQ_FALLBACK;
            else {
                o.whi = hi + 1;
                o.wlo = lo;
                if (o.whi - o.wlo > IPB_PQ_WIN) o.wlo = o.whi - IPB_PQ_WIN;
                o.mode = IPB_PQ_OK;
            }
            return o;
        };
        const IpbHistWin oD = window(0, 1), oP = window(2, 2);
        for (int u = 0; u < pp.n_jobs; ++u) {
            const int j = pp.job[u];
            IpbHistWin oj = j == r.jP ? oP : oD;
            if ((j == r.jF && !any_q[0]) || (j == r.jS && !any_q[1])) oj.mode = oj.mode == IPB_PQ_FALLBACK ? IPB_PQ_FALLBACK : IPB_PQ_IDLE;
            win[j] = oj;
        }
    }
}

// ---------------------------------------------------------------- 2. count
// one 8-pixel unit against the dense window: pixels below wlo -> register counters, pixels
// inside -> shared-memory bins (S's pixels in the second copy)
__device__ __forceinline__ void ipb_pq_count_unit(const uint4& q, unsigned wlo, unsigned whi, unsigned pS, bool haveF,
                                                  unsigned* sh, unsigned& cF, unsigned& cS) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {                     // (a fully predicated variant measured 6 % slower)
        const unsigned v = ipb_pq_px(q, t);
        if (v >= whi) continue;
        const bool inS = (pS >> t) & 1u;
        if (!inS && !haveF) continue;
        if (v < wlo) { if (inS) ++cS; else ++cF; }
        else atomicAdd(&sh[(inS ? IPB_PQ_WIN : 0) + (int)(v - wlo)], 1u);
    }
}

// grid (chunks, n_passes), IPB_PQ_THREADS threads; a CTA streams a contiguous band of units with
// four 128-bit loads in flight per thread.  cnt[job] = pixels below the window; hist_win[job] =
// the window's bins; stats[job][1..2] = moments of the plane for jobs that asked.
// (Measured variants: one instantiation per pass kind -- plain passes alone stream at 3.9 TB/s, but
// the two half-empty launches took 156 + 68 us against 188 us for this mixed one; a main loop of
// whole trips without bounds tests behind a per-unit lambda: 231 us.)
#define IPB_PQ_QCAP 64             // queue entries per warp (at most 31 left over + 32 new)
__global__ void __launch_bounds__(IPB_PQ_THREADS, 4)
ipb_k_pq_count(const unsigned short* __restrict__ planes, int H, int W,
               const IpbPlanePass* __restrict__ passes, const IpbHistJob* __restrict__ jobs,
               const IpbHistWin* __restrict__ win, unsigned units_per_chunk,
               unsigned* __restrict__ hist_win,
               unsigned long long* __restrict__ cnt, unsigned long long* __restrict__ stats)
{
    __shared__ unsigned sh[3 * IPB_PQ_WIN];       // [0, WIN): F's pixels outside S; [WIN, 2 WIN): S's pixels; [2 WIN, 3 WIN): P's
    __shared__ uint4 queue[IPB_PQ_THREADS / 32][IPB_PQ_QCAP];
    __shared__ unsigned long long acc[5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const IpbPlanePass pp = passes[blockIdx.y];
    const IpbPqRoles r = ipb_pq_roles(pp, jobs, W);
    if (!r.ok) return;
    const unsigned long long U = ((unsigned long long)H * (unsigned long long)W) >> 3;
    const unsigned long long u_beg = (unsigned long long)blockIdx.x * units_per_chunk;
    if (u_beg >= U) return;
    unsigned long long u_end = u_beg + units_per_chunk;
    if (u_end > U) u_end = U;
    const int jw = r.jF >= 0 ? r.jF : r.jS;
    IpbHistWin w; w.wlo = 0; w.whi = 0; w.mode = IPB_PQ_IDLE; w.pad = 0;
    IpbHistWin wP = w;
    if (jw >= 0) w = win[jw];
    if (r.jP >= 0) wP = win[r.jP];
    const bool windowed = w.mode == IPB_PQ_OK;
    const bool moments = r.moments != 0;
    const bool sparse = r.jP >= 0 && wP.mode == IPB_PQ_OK;
    if (!windowed && !moments && !sparse) return;
    const unsigned wlo = windowed ? (unsigned)w.wlo : 0u, whi = windowed ? (unsigned)w.whi : 0u;
    const unsigned wloP = sparse ? (unsigned)wP.wlo : 0u, whiP = sparse ? (unsigned)wP.whi : 0u;
    const bool haveF = r.jF >= 0;
    for (int b = tid; b < 3 * IPB_PQ_WIN; b += blockDim.x) sh[b] = 0u;
    if (tid < 5) acc[tid] = 0;
    __syncthreads();

    const uint4* img = reinterpret_cast<const uint4*>(planes + (size_t)pp.plane * H * W);
    const unsigned upr = (unsigned)W >> 3;
    // row of a unit: one multiply-high while unit * upr stays below 2^32 (every frame up to 4096^2)
    const bool fastdiv = U * (unsigned long long)upr < 0xffffffffull;
    const unsigned mg_upr = (unsigned)((0x100000000ull + upr - 1) / upr);
    const unsigned mg_kP = (unsigned)((0x100000000ull + r.kP - 1) / r.kP);
    unsigned cF = 0, cS = 0, cP = 0;                          // below the window: outside S / in S / of P
    unsigned long long s1 = 0, s2 = 0;
    unsigned qn = 0;                                          // entries waiting in this warp's queue (warp-uniform)
    uint4* wq = queue[warp];
    const unsigned lt = (1u << lane) - 1u;

    for (unsigned long long base = u_beg; base < u_end; base += 4ull * IPB_PQ_THREADS) {      // block-uniform trips
        uint4 q[4];
        bool ok[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const unsigned long long u = base + (unsigned long long)g * IPB_PQ_THREADS + (unsigned)tid;
            ok[g] = u < u_end;
            q[g] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            if (ok[g]) q[g] = __ldg(img + u);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (moments && ok[g]) {
                const unsigned ws[4] = {q[g].x, q[g].y, q[g].z, q[g].w};
                unsigned s = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned a = ws[j] & 0xffffu, b = ws[j] >> 16;
                    s += a + b;
                    s2 += (unsigned long long)a * a;
                    s2 += (unsigned long long)b * b;
                }
                s1 += s;
            }
            if (windowed) {
                const bool low = ok[g] && ipb_pq_min8(q[g]) < whi;
                const unsigned m = __ballot_sync(IPB_FULL, low);
                if (m) {                                               // warp-uniform
                    if (low) wq[qn + (unsigned)__popc(m & lt)] = q[g];
                    qn += (unsigned)__popc(m);
                    __syncwarp();
                    if (qn >= 32u) {
                        const uint4 e = wq[lane];
                        const uint4 rest = wq[32 + lane];
                        __syncwarp();
                        ipb_pq_count_unit(e, wlo, whi, r.pS, haveF, sh, cF, cS);
                        qn -= 32u;
                        if ((unsigned)lane < qn) wq[lane] = rest;
                        __syncwarp();
                    }
                }
            }
            if (sparse && ok[g]) {
                const unsigned long long u = base + (unsigned long long)g * IPB_PQ_THREADS + (unsigned)tid;
                const unsigned y = fastdiv ? __umulhi((unsigned)u, mg_upr) : (unsigned)(u / upr);
                if (y - __umulhi(y, mg_kP) * r.kP == 0u) {                   // y % kP == 0 (y < 2^32 / kP)
                    const unsigned x0 = ((unsigned)(u - (unsigned long long)y * upr)) << 3;
                    const unsigned xm = x0 - __umulhi(x0, mg_kP) * r.kP;
                    unsigned sel = 0;
                    for (unsigned t = xm ? r.kP - xm : 0u; t < 8u; t += r.kP) sel |= 1u << t;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (!((sel >> t) & 1u)) continue;
                        const unsigned v = ipb_pq_px(q[g], t);
                        if (v < wloP) ++cP;
                        else if (v < whiP) atomicAdd(&sh[2 * IPB_PQ_WIN + (int)(v - wloP)], 1u);
                    }
                }
            }
        }
    }
    if (windowed && (unsigned)lane < qn) ipb_pq_count_unit(wq[lane], wlo, whi, r.pS, haveF, sh, cF, cS);
    __syncthreads();
    // ---- flush: S's window also belongs to F (S is a subset of F)
    if (windowed) {
        unsigned* gF = haveF ? hist_win + (size_t)r.jF * IPB_PQ_WIN : nullptr;
        unsigned* gS = r.jS >= 0 ? hist_win + (size_t)r.jS * IPB_PQ_WIN : nullptr;
        for (int b = tid; b < IPB_PQ_WIN; b += blockDim.x) {
            const unsigned a = sh[b], s = sh[IPB_PQ_WIN + b];
            if (gF && (a + s)) atomicAdd(&gF[b], a + s);
            if (gS && s) atomicAdd(&gS[b], s);
        }
    }
    if (sparse) {
        unsigned* gP = hist_win + (size_t)r.jP * IPB_PQ_WIN;
        for (int b = tid; b < IPB_PQ_WIN; b += blockDim.x) { const unsigned p = sh[2 * IPB_PQ_WIN + b]; if (p) atomicAdd(&gP[b], p); }
    }
    const unsigned long long a0 = ipb_warp_sum((unsigned long long)cF), a1 = ipb_warp_sum((unsigned long long)cS),
                             a2 = ipb_warp_sum((unsigned long long)cP);
    s1 = ipb_warp_sum(s1); s2 = ipb_warp_sum(s2);
    if (lane == 0) {
        if (a0) atomicAdd(&acc[0], a0);
        if (a1) atomicAdd(&acc[1], a1);
        if (s1) atomicAdd(&acc[2], s1);
        if (s2) atomicAdd(&acc[3], s2);
        if (a2) atomicAdd(&acc[4], a2);
    }
    __syncthreads();
    if (tid == 0) {
        if (haveF && (acc[0] + acc[1])) atomicAdd(&cnt[r.jF], acc[0] + acc[1]);
        if (r.jS >= 0 && acc[1]) atomicAdd(&cnt[r.jS], acc[1]);
        if (r.jP >= 0 && acc[4]) atomicAdd(&cnt[r.jP], acc[4]);
    }
    if (tid < pp.n_jobs && jobs[pp.job[tid]].moments) {
        const int j = pp.job[tid];
        if (acc[2]) atomicAdd(&stats[(size_t)j * 4 + 1], acc[2]);
        if (acc[3]) atomicAdd(&stats[(size_t)j * 4 + 2], acc[3]);
    }
}

// ---------------------------------------------------------------- 3. select
// one CTA (256 threads) per quantile job
__global__ void __launch_bounds__(256)
ipb_k_pq_select(const IpbQJob* __restrict__ qjobs, const IpbHistWin* __restrict__ win,
                const unsigned long long* __restrict__ cnt, const unsigned* __restrict__ hist_win,
                const unsigned long long* __restrict__ stats,
                IpbQOut* __restrict__ out, unsigned* __restrict__ miss)
{
    const IpbQJob qj = qjobs[blockIdx.x];
    const IpbHistWin w = win[qj.hist];
    const unsigned long long n = stats[(size_t)qj.hist * 4];
    __shared__ unsigned long long scan32[32];
    __shared__ int res[2];
    if (threadIdx.x < 2) res[threadIdx.x] = -1;
    __syncthreads();
    IpbQIdx qi;
    qi.prev = 0; qi.next = 0; qi.gamma = 0.f;
    bool missed = w.mode != IPB_PQ_OK;
    if (n > 0 && !missed) {                                       // block-uniform
        qi = ipb_np_qidx_f32((long long)n, qj.q32);
        const unsigned long long below = cnt[qj.hist];
        if ((unsigned long long)qi.prev < below) missed = true;
        else {
            const unsigned* h = hist_win + (size_t)qj.hist * IPB_PQ_WIN;
            const unsigned long long want[2] = {(unsigned long long)qi.prev - below, (unsigned long long)qi.next - below};
            ipb_locate_ranks_smem(IPB_PQ_WIN, IPB_PQ_WIN / 256, want, 2, scan32, [&](unsigned i) { return h[i]; },
                                  [&](int k, unsigned i, unsigned) { res[k] = (int)i; });
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        IpbQOut o;
        const bool ok = n > 0 && !missed && res[0] >= 0 && res[1] >= 0;
        o.prev = ok ? w.wlo + res[0] : -1; o.next = ok ? w.wlo + res[1] : -1; o.gamma = qi.gamma; o.n = n;
        o.value = ok ? ipb_np_lerp_f32((float)o.prev, (float)o.next, qi.gamma) : 0.0f;
        if (n > 0 && !ok) atomicAdd(miss, 1u);
        out[blockIdx.x] = o;
    }
}
