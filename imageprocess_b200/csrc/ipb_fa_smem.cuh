// ipb_fa_smem.cuh -- the focal-adhesion chain of one crop entirely in shared memory.
// Included by ipb_fa.cuh (same phases, same exact semantics as the kernels there).
//
// ipb_k_fa_fused walks its union-find and its bit images through L2: ~15 dependent phases of a
// few global round trips each, ~80 us per cell crop although a crop is only ~2-4 K words of bits.
// Here ONE CTA keeps everything of its crop on chip:
//   * the two bit images it ping-pongs between (threshold -> size filter -> dilation -> erosion),
//   * a compact run index: base[word] = number of run starts before the word (block scan), so a
//     run's id is base[word] + its rank among the starts of the word; ids follow raster order,
//   * union-find parents and component sizes over run ids (shared-memory atomicMin / atomicAdd),
//   * per-component sums (area, intensity, coordinates) accumulated per run.
// Component ranks (skimage's raster-order label numbers) are a block scan over the root flags of
// the run ids.  Results leave the CTA as: the final bit rows (bw_final), a staged component table
// (csize slice: 8 words per adhesion) and a run table (L slice: 2 words per run: y<<16 | x0,
// len<<16 | rank) from which the optional label map is painted.  A crop that does not fit
// (> IPB_FAS_MAXWORDS words, > IPB_FAS_MAXRUNS runs, > IPB_FAS_MAXCOMP adhesions, staging larger
// than its slices) is flagged (crop_count = -1) and taken by ipb_k_fa_fused afterwards.
#pragma once

#define IPB_FAS_THREADS 512
#define IPB_FAS_MAXWORDS 4096
#define IPB_FAS_MAXRUNS 8192
#define IPB_FAS_MAXCOMP 512
#define IPB_FAS_MARK (-7)          // row_base[c.row_off]: this crop was finished by the shared-memory kernel
#define IPB_FAS_STAGES 4           // row buffers per warp of the threshold phase's TMA pipeline
#define IPB_FAS_ROWBUF 1024        // bytes per row buffer: crop rows of up to 504 + 7 pixels (16 warps x 4 x 1 KB = the union-find area)
#define IPB_FAS_SMEM_BYTES (IPB_FAS_MAXWORDS * 4 * 2 + IPB_FAS_MAXWORDS * 2 + IPB_FAS_MAXRUNS * 4 * 2)

__device__ __forceinline__ unsigned ipb_fas_find(unsigned* parent, unsigned i) {
    volatile unsigned* P = parent;
    while (true) {
        const unsigned p = P[i];
        if (p == i) return i;
        const unsigned gp = P[p];
        if (gp == p) return p;
        atomicMin(&parent[i], gp);          // path halving; parents only ever decrease
        i = gp;
    }
}
__device__ __forceinline__ void ipb_fas_union(unsigned* parent, unsigned a, unsigned b) {
    while (true) {
        a = ipb_fas_find(parent, a);
        b = ipb_fas_find(parent, b);
        if (a == b) return;
        if (a > b) { const unsigned t = a; a = b; b = t; }
        const unsigned old = atomicMin(&parent[b], a);
        if (old == b) return;
        b = old;
    }
}

// base[word] = run starts before the word (raster order); returns the number of runs.
// All threads must call it; ends with a barrier.
__device__ __forceinline__ unsigned ipb_fas_index_runs(const IpbCrop& c, const unsigned* bits, unsigned short* base, unsigned* wsum) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int nwords = c.h * c.wpr;
    const int per = (nwords + (int)blockDim.x - 1) / (int)blockDim.x;
    const int w0 = tid * per;
    unsigned cnt = 0;
    for (int k = 0; k < per; ++k) {
        const int wi = w0 + k;
        if (wi < nwords) { const int y = wi / c.wpr, j = wi - y * c.wpr; cnt += (unsigned)__popc(ipb_bits_starts(bits + (size_t)y * c.wpr, j)); }
    }
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
    __syncthreads();                        // wsum may still be read from an earlier call
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int i = 0; i < nwarps; ++i) { const unsigned t = wsum[i]; total += t; if (i < warp) before += t; }
    unsigned run = before + incl - cnt;
    for (int k = 0; k < per; ++k) {
        const int wi = w0 + k;
        if (wi < nwords) {
            const int y = wi / c.wpr, j = wi - y * c.wpr;
            base[wi] = (unsigned short)(run < 0xffffu ? run : 0xffffu);
            run += (unsigned)__popc(ipb_bits_starts(bits + (size_t)y * c.wpr, j));
        }
    }
    __syncthreads();
    return total;
}

// id of the run that starts at pixel x of row y
__device__ __forceinline__ unsigned ipb_fas_run_id(const IpbCrop& c, const unsigned* bits, const unsigned short* base, int y, int x) {
    const int j = x >> 5;
    const unsigned s = ipb_bits_starts(bits + (size_t)y * c.wpr, j);
    return (unsigned)base[y * c.wpr + j] + (unsigned)__popc(s & ((1u << (x & 31)) - 1u));
}

#define IPB_FAS_FOREACH_RUN(c, bits, base, ...)                                                \
    for (int i_ = threadIdx.x; i_ < (c).h * (c).wpr; i_ += blockDim.x) {                       \
        const int y = i_ / (c).wpr, j = i_ - y * (c).wpr;                                      \
        const unsigned* row = (bits) + (size_t)y * (c).wpr;                                    \
        unsigned s_ = ipb_bits_starts(row, j);                                                 \
        unsigned id = (base)[i_];                                                              \
        while (s_) {                                                                           \
            const int b_ = __ffs((int)s_) - 1;                                                 \
            s_ &= s_ - 1;                                                                      \
            const int a = 32 * j + b_;                                                         \
            const int e = ipb_bits_next_clear(row, a, (c).w);       /* run = [a, e) */         \
            __VA_ARGS__                                                                        \
            ++id;                                                                              \
        }                                                                                      \
    }

// ---- threshold & ROI mask with 128-bit loads.  A warp owns a crop row; lane L reads the two aligned
// 8-pixel units 4 j0 + 2 L, + 1 of the frame row (one pass covers 15 crop words), turns them into 16
// comparison bits, and lanes 0..14 assemble the crop words from three neighbouring lanes' bits (the
// crop's left edge is not unit-aligned: shift by ox & 7).  (float)px > thr is evaluated
// as the equivalent integer test px > floor(thr).
__device__ __forceinline__ unsigned ipb_fas_unit_bits(const uint4& q, int ithr) {
    unsigned b = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const unsigned w = t < 4 ? (t < 2 ? q.x : q.y) : (t < 6 ? q.z : q.w);
        const int v = (int)((t & 1) ? (w >> 16) : (w & 0xffffu));
        b |= (v > ithr ? 1u : 0u) << t;
    }
    return b;
}
__device__ __forceinline__ void ipb_fas_threshold(const IpbCrop& c, const unsigned short* __restrict__ planes, int H, int W,
                                                  const float* __restrict__ fa_params, const unsigned* __restrict__ roi_mask, unsigned* A,
                                                  unsigned char* ring /* nwarps x STAGES x ROWBUF bytes, 16-byte aligned; null: no TMA */,
                                                  IpbMbar* bars /* nwarps x STAGES */)
{
    const unsigned short* img = planes + (size_t)c.plane * H * W;
    if ((W & 7) != 0 || (((size_t)img) & 15) != 0) {               // block-uniform: unaligned frames take the scalar phase
        IpbCrop cs = c;
        cs.bit_off = 0;
        ipb_k_fa_threshold_phase(cs, 0, c.h, planes, H, W, fa_params, roi_mask, A);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float thr = fa_params[(size_t)c.frame * 4 + 3];
    int ithr;
    if (!(thr == thr)) ithr = 65536;                               // NaN: no pixel passes
    else if (thr < 0.0f) ithr = -1;
    else if (thr >= 65535.0f) ithr = 65535;
    else ithr = (int)floorf(thr);
    const int k0 = c.ox >> 3, s = c.ox & 7;
    const int nunits = ((c.ox + c.w + 7) >> 3) - k0;
    const int src = 2 * (lane < 15 ? lane : 0);
    // Rows arrive by TMA: every warp owns IPB_FAS_STAGES row buffers in the (still unused) union-find
    // area and one mbarrier per buffer; lane 0 starts the bulk copy of row y + STAGES * nwarps as
    // soon as the warp has read row y, so the loads cost no issue slots and their latency hides
    // behind the comparisons of the rows in between.  Rows wider than a buffer take the LDG loop.
    const unsigned row_bytes = 16u * (unsigned)nunits;
    const bool tma = ring != nullptr && row_bytes <= (unsigned)IPB_FAS_ROWBUF;          // block-uniform
    unsigned char* mybuf = tma ? ring + (size_t)warp * IPB_FAS_STAGES * IPB_FAS_ROWBUF : nullptr;
    IpbMbar* mybar = bars + warp * IPB_FAS_STAGES;
    if (tma) {
        if (lane == 0) {
#pragma unroll
            for (int st = 0; st < IPB_FAS_STAGES; ++st) ipb_mbar_init(mybar + st, 1u);
            ipb_mbar_fence_init();
#pragma unroll
            for (int st = 0; st < IPB_FAS_STAGES; ++st) {
                const int y = warp + st * nwarps;
                if (y < c.h) ipb_bulk_load(mybuf + st * IPB_FAS_ROWBUF,
                                           reinterpret_cast<const uint4*>(img + (size_t)(c.oy + y) * W) + k0, row_bytes, mybar + st);
            }
        }
        __syncwarp();
    }
    int it = 0;
    for (int y = warp; y < c.h; y += nwarps, ++it) {
        const uint4* row = reinterpret_cast<const uint4*>(img + (size_t)(c.oy + y) * W) + k0;
        const int st = it % IPB_FAS_STAGES;
        const uint4* srow = reinterpret_cast<const uint4*>(mybuf + st * IPB_FAS_ROWBUF);
        if (tma) ipb_mbar_wait(mybar + st, (unsigned)(it / IPB_FAS_STAGES) & 1u);
        const unsigned* mrow = roi_mask + c.mask_off + (size_t)y * c.wpr;
        // a lane takes two adjacent units (16 pixels): one pass of the warp covers 15 crop words
        for (int j0 = 0; j0 < c.wpr; j0 += 15) {
            const int unit = 4 * j0 + 2 * lane;
            uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
            const bool in0 = unit < nunits, in1 = unit + 1 < nunits;
            if (tma) {
                if (in0) q0 = srow[unit];
                if (in1) q1 = srow[unit + 1];
            } else {
                if (in0) q0 = __ldg(row + unit);
                if (in1) q1 = __ldg(row + unit + 1);
            }
            const unsigned half = (in0 ? ipb_fas_unit_bits(q0, ithr) : 0u) | ((in1 ? ipb_fas_unit_bits(q1, ithr) : 0u) << 8);
            unsigned long long v = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) v |= (unsigned long long)__shfl_sync(IPB_FULL, half, src + k) << (16 * k);
            unsigned word = (unsigned)(v >> s);
            const int j = j0 + lane;
            if (lane < 15 && j < c.wpr) {
                const int rem = c.w - 32 * j;
                if (rem < 32) word &= (1u << rem) - 1u;
                A[(size_t)y * c.wpr + j] = word & mrow[j];
            }
        }
        if (tma) {
            __syncwarp();                                          // every lane has read the buffer
            const int yn = y + IPB_FAS_STAGES * nwarps;
            if (lane == 0 && yn < c.h)
                ipb_bulk_load(mybuf + st * IPB_FAS_ROWBUF, reinterpret_cast<const uint4*>(img + (size_t)(c.oy + yn) * W) + k0,
                              row_bytes, mybar + st);
        }
    }
}

template <int CONN>
__device__ __forceinline__ void ipb_fas_merge(const IpbCrop& c, const unsigned* bits, const unsigned short* base, unsigned* parent) {
    IPB_FAS_FOREACH_RUN(c, bits, base, {
        if (y > 0) {
            const unsigned* prow = row - c.wpr;
            int lo = a, hi = e - 1;
            if (CONN == 8) { lo = a > 0 ? a - 1 : 0; hi = e < c.w ? e : c.w - 1; }
            int pos = lo;
            while (pos <= hi) {
                const int t = ipb_bits_next_set(prow, pos, hi);
                if (t < 0) break;
                const int ps = ipb_bits_run_start(prow, t);
                ipb_fas_union(parent, id, ipb_fas_run_id(c, bits, base, y - 1, ps));
                pos = ipb_bits_next_clear(prow, t, c.w) + 1;
            }
        }
    })
}

__global__ void __launch_bounds__(IPB_FAS_THREADS, 2)
ipb_k_fa_fused_smem(const IpbCrop* __restrict__ crops, const unsigned short* __restrict__ planes, int H, int W,
                    const float* __restrict__ fa_params, const unsigned* __restrict__ roi_mask,
                    double min_size, IpbDisk disk, int* __restrict__ L /* run tables */, unsigned* __restrict__ csize /* staged comps */,
                    int* __restrict__ row_roots, int* __restrict__ row_base, int* __restrict__ crop_count,
                    unsigned* __restrict__ bw_final, int* __restrict__ labels /* nullable: zeroed here, painted by the gather */,
                    const int* __restrict__ order)
{
    IPB_DYN_SMEM(unsigned, smem);
    unsigned* A = smem;
    unsigned* B = A + IPB_FAS_MAXWORDS;
    unsigned* parent = B + IPB_FAS_MAXWORDS;
    unsigned* size = parent + IPB_FAS_MAXRUNS;
    unsigned short* base = reinterpret_cast<unsigned short*>(size + IPB_FAS_MAXRUNS);
    __shared__ unsigned wsum[IPB_FAS_THREADS / 32];
    __shared__ IpbMbar bars[(IPB_FAS_THREADS / 32) * IPB_FAS_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int ci = order ? order[blockIdx.x] : (int)blockIdx.x;
    const IpbCrop c = crops[ci];
    const int nwords = c.h * c.wpr;
    const long long npx = (long long)c.w * c.h;
    if (nwords > IPB_FAS_MAXWORDS || c.w > 8192 || c.h >= 65536) {                // block-uniform
        if (tid == 0) crop_count[ci] = -1;
        return;
    }
    IpbCrop cs = c;
    cs.bit_off = 0;                                         // bit rows live in shared memory from here on

    // ---- 1. threshold & ROI mask -> A
    ipb_fas_threshold(c, planes, H, W, fa_params, roi_mask, A, reinterpret_cast<unsigned char*>(parent), bars);
    __syncthreads();
    unsigned* cur = A;
    unsigned* other = B;

    // ---- 2. remove_small_objects: 4-connected components over run ids, sizes, filter -> other
    if (min_size > 0) {
        const unsigned nruns = ipb_fas_index_runs(cs, cur, base, wsum);
        if (nruns > IPB_FAS_MAXRUNS) { if (tid == 0) crop_count[ci] = -1; return; }
        for (unsigned i = tid; i < nruns; i += blockDim.x) { parent[i] = i; size[i] = 0u; }
        __syncthreads();
        ipb_fas_merge<4>(cs, cur, base, parent);
        __syncthreads();
        IPB_FAS_FOREACH_RUN(cs, cur, base, {
            const unsigned r = ipb_fas_find(parent, id);
            parent[id] = r;
            atomicAdd(&size[r], (unsigned)(e - a));
        })
        __syncthreads();
        for (int i = tid; i < nwords; i += blockDim.x) {
            const int y = i / c.wpr, j = i - y * c.wpr;
            const unsigned* row = cur + (size_t)y * c.wpr;
            const unsigned wv = row[j];
            unsigned keep = 0u, todo = wv;
            while (todo) {
                const int b = __ffs((int)todo) - 1;
                const int x = 32 * j + b;
                const int a = ipb_bits_run_start(row, x);
                const unsigned r = parent[ipb_fas_run_id(cs, cur, base, y, a)];      // flattened: the run's root
                const bool ok = !((double)size[r] < min_size);
                int e = ipb_bits_next_clear(row, x, c.w);
                if (e > 32 * j + 32) e = 32 * j + 32;
                const int nb = e - x;
                const unsigned seg = (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u)) << b;
                if (ok) keep |= seg;
                todo &= ~seg;
            }
            other[i] = keep;
        }
        __syncthreads();
        unsigned* t = cur; cur = other; other = t;
    }

    // ---- 3. binary closing with disk(r): dilation (outside = 0), erosion (outside = 1)
    if (disk.r > 0) {
        ipb_k_bits_morph_phase<0>(cs, 0, c.h, cur, disk, other);
        __syncthreads();
        ipb_k_bits_morph_phase<1>(cs, 0, c.h, other, disk, cur);
        __syncthreads();
    }
    const unsigned* fin = cur;
    for (int i = tid; i < nwords; i += blockDim.x) bw_final[(size_t)c.bit_off + i] = fin[i];

    // ---- 4. label: 8-connected components, raster-order ranks
    const unsigned nruns = ipb_fas_index_runs(cs, fin, base, wsum);
    if (nruns > IPB_FAS_MAXRUNS || 2ll * nruns > npx) { if (tid == 0) crop_count[ci] = -1; return; }
    for (unsigned i = tid; i < nruns; i += blockDim.x) parent[i] = i;
    __syncthreads();
    ipb_fas_merge<8>(cs, fin, base, parent);
    __syncthreads();
    for (unsigned i = tid; i < nruns; i += blockDim.x) {
        const unsigned r = ipb_fas_find(parent, i);
        parent[i] = r;
        size[i] = (r == i) ? 1u : 0u;                       // root flag, turned into the root's rank below
    }
    __syncthreads();
    unsigned nroots;
    {
        const unsigned per = (nruns + blockDim.x - 1) / blockDim.x;
        const unsigned i0 = (unsigned)tid * per;
        unsigned cnt = 0;
        for (unsigned k = 0; k < per; ++k) if (i0 + k < nruns) cnt += size[i0 + k];
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int i = 0; i < nwarps; ++i) { const unsigned t = wsum[i]; total += t; if (i < warp) before += t; }
        unsigned run = before + incl - cnt;
        for (unsigned k = 0; k < per; ++k) if (i0 + k < nruns) { const unsigned f = size[i0 + k]; size[i0 + k] = run; run += f; }
        nroots = total;
        __syncthreads();
    }
    if (nroots > IPB_FAS_MAXCOMP || 8ll * nroots > npx) { if (tid == 0) crop_count[ci] = -1; return; }

    // ---- 5. per-adhesion sums (shared accumulators in the bit buffer that is free now) + run table.
    // All accumulators are 32-bit (native shared-memory atomics; 64-bit ones are CAS loops): a run's
    // intensity sum fits 32 bits (<= 4096 px x 65535) and goes in as two 16-bit halves; the
    // coordinate sums of a <= 131072-px crop at most 8192 wide stay below 2^30.
    unsigned* acc_a = other;                                   // [MAXCOMP] each: 5 x 2 KB <= 16 KB
    unsigned* acc_il = acc_a + IPB_FAS_MAXCOMP;
    unsigned* acc_ih = acc_il + IPB_FAS_MAXCOMP;
    unsigned* acc_y = acc_ih + IPB_FAS_MAXCOMP;
    unsigned* acc_x = acc_y + IPB_FAS_MAXCOMP;
    for (unsigned i = tid; i < 5u * IPB_FAS_MAXCOMP; i += blockDim.x) acc_a[i] = 0u;
    __syncthreads();
    const unsigned short* img = planes + (size_t)c.plane * H * W;
    int* Lc = L + c.pix_off;
    // word by word, not run by run: a thread sums at most the 32 pixels of its word's segments, so the
    // long runs of large adhesions spread over many threads; the run table is only written when the
    // label map is wanted
    for (int wi = tid; wi < nwords; wi += blockDim.x) {
        const int y = wi / c.wpr, j = wi - y * c.wpr;
        const unsigned* row = fin + (size_t)y * c.wpr;
        const unsigned wv = row[j];
        if (!wv) continue;
        const unsigned starts = ipb_bits_starts(row, j);
        const unsigned id0 = base[wi];                         // id of the first run that STARTS in this word
        const unsigned short* irow = img + (size_t)(c.oy + y) * W + c.ox + 32 * j;
        unsigned todo = wv, k = 0;
        while (todo) {
            const int b = __ffs((int)todo) - 1;
            const unsigned inv = ~wv >> b;                     // bit 0 is clear (the segment starts at b)
            const int nb = inv ? __ffs((int)inv) - 1 : 32 - b;
            const unsigned seg = (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u)) << b;
            todo &= ~seg;
            const bool is_start = (starts >> b) & 1u;          // else: the run continues from the previous word (b == 0)
            const unsigned rid = is_start ? id0 + k : id0 - 1u;
            k += is_start ? 1u : 0u;
            const unsigned rank = size[parent[rid]];
            unsigned si = 0;
            for (int t = 0; t < nb; t += 8) {                  // eight pixel loads in flight (one at a time left the
                unsigned v[8];                                 // thread waiting on L2 once per pixel: 8.5 % of the
#pragma unroll                                                 // kernel's stall samples)
                for (int u = 0; u < 8; ++u) v[u] = (t + u < nb) ? (unsigned)irow[b + t + u] : 0u;
#pragma unroll
                for (int u = 0; u < 8; ++u) si += v[u];
            }
            const unsigned len = (unsigned)nb, x0 = (unsigned)(32 * j + b);
            atomicAdd(&acc_a[rank], len);
            atomicAdd(&acc_il[rank], si & 0xffffu);
            atomicAdd(&acc_ih[rank], si >> 16);
            atomicAdd(&acc_y[rank], len * (unsigned)y);
            atomicAdd(&acc_x[rank], (2u * x0 + len - 1u) * len / 2u);
            if (labels && is_start) {
                const unsigned e = (unsigned)ipb_bits_next_clear(row, (int)x0, c.w);
                Lc[2 * rid] = (int)(((unsigned)y << 16) | x0);
                Lc[2 * rid + 1] = (int)(((e - x0) << 16) | rank);
            }
        }
    }
    __syncthreads();
    unsigned* st = csize + c.pix_off;
    for (unsigned i = tid; i < nroots; i += blockDim.x) {
        const unsigned long long si = ((unsigned long long)acc_ih[i] << 16) + acc_il[i];
        st[8 * i + 0] = (unsigned)si; st[8 * i + 1] = (unsigned)(si >> 32);
        st[8 * i + 2] = acc_y[i]; st[8 * i + 3] = 0u;
        st[8 * i + 4] = acc_x[i]; st[8 * i + 5] = 0u;
        st[8 * i + 6] = acc_a[i]; st[8 * i + 7] = (unsigned)ci;
    }
    if (labels) for (long long i = tid; i < npx; i += blockDim.x) labels[c.pix_off + i] = 0;
    if (tid == 0) {
        crop_count[ci] = (int)nroots;
        row_roots[c.row_off] = (int)nruns;
        row_base[c.row_off] = IPB_FAS_MARK;
    }
}

// staged component tables -> the compact table (after the crop scan); optional label map from
// the run tables.  grid (8, n_crops), 256 threads.  Crops of the global-memory path are skipped.
__global__ void __launch_bounds__(256)
ipb_k_fa_gather_smem(const IpbCrop* __restrict__ crops, const int* __restrict__ L, const unsigned* __restrict__ csize,
                     const int* __restrict__ row_roots, const int* __restrict__ row_base,
                     const int* __restrict__ crop_count, const int* __restrict__ comp_off, int cap,
                     IpbComp* __restrict__ comps, int* __restrict__ labels /* nullable, zeroed */)
{
    const int ci = blockIdx.y;
    const IpbCrop c = crops[ci];
    if (row_base[c.row_off] != IPB_FAS_MARK) return;
    const int n = crop_count[ci], off = comp_off[ci];
    const unsigned* st = csize + c.pix_off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (off + i >= cap) break;
        IpbComp o;
        o.sum_i = (unsigned long long)st[8 * i] | ((unsigned long long)st[8 * i + 1] << 32);
        o.sum_y = (unsigned long long)st[8 * i + 2] | ((unsigned long long)st[8 * i + 3] << 32);
        o.sum_x = (unsigned long long)st[8 * i + 4] | ((unsigned long long)st[8 * i + 5] << 32);
        o.area = st[8 * i + 6];
        o.crop = ci;
        comps[off + i] = o;
    }
    if (labels) {
        const int nruns = row_roots[c.row_off];
        const int* Lc = L + c.pix_off;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nruns; i += gridDim.x * blockDim.x) {
            const unsigned p = (unsigned)Lc[2 * i], q = (unsigned)Lc[2 * i + 1];
            const int y = (int)(p >> 16), a = (int)(p & 0xffffu), len = (int)(q >> 16), rank = (int)(q & 0xffffu);
            int* lrow = labels + c.pix_off + (size_t)y * c.w + a;
            for (int k = 0; k < len; ++k) lrow[k] = rank + 1;
        }
    }
}
