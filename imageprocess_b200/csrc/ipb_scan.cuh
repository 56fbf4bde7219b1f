// ipb_scan.cuh -- block-wide rank location in a histogram, shared by the percentile kernels.
#pragma once
#include "ipb_rt.cuh"

// Warp-cooperative scan of `nrows` rows of 32 counters: finds, for every wanted rank, the
// counter that holds it.  value(i) = counter i (0 beyond the end).  hit(r, i, rank_inside).
template <typename V, typename HIT>
__device__ __forceinline__ void ipb_locate_ranks(unsigned nrows, const unsigned long long* want, int nr,
                                              unsigned long long* red_u, V value, HIT hit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const unsigned rpw = (nrows + (unsigned)nwarps - 1u) / (unsigned)nwarps;
    const unsigned row0 = (unsigned)warp * rpw;
    unsigned row1 = row0 + rpw;
    if (row1 > nrows) row1 = nrows;
    unsigned long long band = 0;
#pragma unroll 8
    for (unsigned row = row0; row < row1; ++row) band += value((row << 5) + (unsigned)lane);   // independent loads
    band = ipb_warp_sum(band);
    __syncthreads();
    if (lane == 0) red_u[warp] = band;
    __syncthreads();
    unsigned long long base = 0;
    for (int i = 0; i < warp; ++i) base += red_u[i];
    bool mine = false;
    for (int r = 0; r < nr; ++r) mine = mine || (want[r] >= base && want[r] < base + band);
    if (mine) {                                                            // warp-uniform
        unsigned long long run = base;
        for (unsigned row = row0; row < row1; ++row) {
            const unsigned i = (row << 5) + (unsigned)lane;
            const unsigned v = value(i);
            unsigned incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
            const unsigned rowtot = __shfl_sync(IPB_FULL, incl, 31);
            for (int r = 0; r < nr; ++r) {
                const unsigned long long kk = want[r];
                if (kk >= run && kk < run + rowtot) {
                    const unsigned off = (unsigned)(kk - run);
                    if (off >= incl - v && off < incl) hit(r, i, off - (incl - v));
                }
            }
            run += rowtot;
        }
    }
    __syncthreads();
}


// Same task for counters that live in SHARED memory (or registers behind `value`): every thread
// owns `per` consecutive counters and sums them with a rotated start (thread t begins at
// counter (t mod per) of its chunk, so the lanes of a warp hit different banks), a block scan
// of the per-thread sums follows, and only the threads whose chunk holds a wanted rank walk
// their chunk in order.  All threads work in parallel; no warp walks a long serial chain.
// `per` = counters per thread (nbins <= per * blockDim.x); scan32 = >= 32 words of shared scratch.
// second half of ipb_locate_ranks_smem for callers that already hold `mine`, the sum of their own
// chunk [tid * per, tid * per + per) (e.g. from a pass over the bins they had to make anyway)
template <typename V, typename HIT>
__device__ __forceinline__ void ipb_locate_ranks_chunks(unsigned nbins, unsigned per, unsigned long long mine,
                                                        const unsigned long long* want, int nr,
                                                        unsigned long long* scan32, V value, HIT hit);

template <typename V, typename HIT>
__device__ __forceinline__ void ipb_locate_ranks_smem(unsigned nbins, unsigned per, const unsigned long long* want,
                                                      int nr, unsigned long long* scan32, V value, HIT hit) {
    const unsigned c0 = (unsigned)threadIdx.x * per;
    unsigned long long mine = 0;
    const unsigned rot = per ? (unsigned)threadIdx.x % per : 0u;
    for (unsigned i = 0; i < per; ++i) {
        unsigned k = i + rot;
        if (k >= per) k -= per;
        const unsigned b = c0 + k;
        if (b < nbins) mine += value(b);
    }
    ipb_locate_ranks_chunks(nbins, per, mine, want, nr, scan32, value, hit);
}

template <typename V, typename HIT>
__device__ __forceinline__ void ipb_locate_ranks_chunks(unsigned nbins, unsigned per, unsigned long long mine,
                                                        const unsigned long long* want, int nr,
                                                        unsigned long long* scan32, V value, HIT hit) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const unsigned c0 = (unsigned)tid * per;
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) scan32[warp] = incl;
    __syncthreads();
    unsigned long long base = 0;
    for (int i = 0; i < warp && i < nwarps; ++i) base += scan32[i];
    const unsigned long long lo = base + incl - mine, hi = base + incl;
    for (int r = 0; r < nr; ++r) {
        const unsigned long long kk = want[r];
        if (kk >= lo && kk < hi) {
            unsigned long long acc = lo;
            for (unsigned i = 0; i < per; ++i) {
                const unsigned b = c0 + i;
                const unsigned v = b < nbins ? value(b) : 0u;
                if (kk < acc + v) { hit(r, b, (unsigned)(kk - acc)); break; }
                acc += v;
            }
        }
    }
    __syncthreads();
}
