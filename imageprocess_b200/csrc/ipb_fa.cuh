// ipb_fa.cuh -- focal-adhesion segmentation chain on a ragged batch of crops
// (reference src/INT/FA_Analyzer.py:123-195 analyze_fa_crop).
//
//   bw = (crop > thr) & roi_mask                      FA_Analyzer.py:146-147
//   remove_small_objects(bw, min_px)  (4-conn CCL)    FA_Analyzer.py:151
//   binary_closing(bw, disk(r))                        FA_Analyzer.py:155-156
//   label (8-conn, raster-order numbering)             FA_Analyzer.py:158
//   regionprops: area, mean intensity, centroid        FA_Analyzer.py:159-186
//
// Every crop keeps the reference's per-crop semantics (crop-border behaviour of the closing,
// labels restarting per crop).  All binary images are bit-packed rows (32 px per word, same
// layout as the rasteriser's mask pool), so thresholding, the size filter and the
// morphology are word-parallel.  Connected components use a run-based union-find: a run is
// a maximal horizontal sequence of set bits, its label slot is the pixel index of its first
// pixel; runs of adjacent rows that touch (4- or 8-connectivity) are united with atomicMin
// on a global label array (block-local work, merged globally through L2 atomics), so a
// component's root is its first pixel in raster order -- exactly the order in which
// skimage.measure.label / scipy.ndimage.label number components.  Sequential label ids are
// then a popcount-prefix over root bits, and per-component sums are accumulated per run
// (not per pixel) with atomics into a compact table.
// The same kernels serve 24 small cell crops per frame and one 8192x8192 mosaic crop.
#pragma once
#include "ipb_rt.cuh"

#define IPB_FA_THREADS 256
#define IPB_FA_ROWS 32         // rows of one crop handled by one CTA (~1-2 words per thread)

struct IpbCrop {
    long long bit_off;   // word offset of the crop's bit rows (all bit pools share it)
    long long pix_off;   // element offset of the crop's per-pixel arrays (labels)
    long long row_off;   // element offset of the crop's per-row arrays
    long long mask_off;  // word offset of the crop's ROI mask rows in the (shared) roi_mask pool
    int ox, oy;          // frame position of crop pixel (0,0)
    int w, h;
    int wpr;             // words per bit row
    int plane;           // uint16 plane index (frame * C + channel)
    int frame;           // frame index (threshold / params lookup)
    int pad0;
};

// ---------------------------------------------------------------- bit-row helpers
// invariant: bits >= w of the last word of every row are zero
__device__ __forceinline__ int ipb_bits_get(const unsigned* row, int x) { return (row[x >> 5] >> (x & 31)) & 1u; }

// smallest x in [from, to] with the bit set, else -1
__device__ __forceinline__ int ipb_bits_next_set(const unsigned* row, int from, int to) {
    if (from > to) return -1;
    int j = from >> 5;
    const int jl = to >> 5;
    unsigned m = row[j] & (0xffffffffu << (from & 31));
    while (true) {
        if (j == jl) m &= (0xffffffffu >> (31 - (to & 31)));
        if (m) return 32 * j + __ffs((int)m) - 1;
        if (j == jl) return -1;
        ++j;
        m = row[j];
    }
}
// smallest x in [from, w) with the bit clear, else w
__device__ __forceinline__ int ipb_bits_next_clear(const unsigned* row, int from, int w) {
    const int wpr = (w + 31) >> 5;
    int j = from >> 5;
    if (j >= wpr) return w;
    unsigned m = ~row[j] & (0xffffffffu << (from & 31));
    while (true) {
        if (m) { const int x = 32 * j + __ffs((int)m) - 1; return x < w ? x : w; }
        ++j;
        if (j >= wpr) return w;
        m = ~row[j];
    }
}
// first pixel of the run containing x (bit x must be set)
__device__ __forceinline__ int ipb_bits_run_start(const unsigned* row, int x) {
    int j = x >> 5;
    const int b = x & 31;
    unsigned m = ~row[j] & (b == 31 ? 0xffffffffu : ((1u << (b + 1)) - 1u));
    while (true) {
        if (m) return 32 * j + (32 - __clz((int)m));
        if (j == 0) return 0;
        --j;
        m = ~row[j];
    }
}
// bit mask of run starts inside word j of a row
__device__ __forceinline__ unsigned ipb_bits_starts(const unsigned* row, int j) {
    const unsigned wv = row[j];
    const unsigned carry = j > 0 ? (row[j - 1] >> 31) : 0u;
    return wv & ~((wv << 1) | carry);
}

// ---------------------------------------------------------------- union-find on run starts
__device__ __forceinline__ int ipb_uf_load(const int* L, int i) {
#ifdef IPB_EMULATE
    return L[i];
#else
    return __ldcg(L + i);          // L2 (coherent) read: parents only ever decrease
#endif
}
// find with path halving: every visited node is re-pointed at its grandparent (atomicMin, so
// parents only ever decrease and concurrent unions stay correct); later finds walk short paths
__device__ __forceinline__ int ipb_uf_find(int* L, int i) {
    while (true) {
        const int p = ipb_uf_load(L, i);
        if (p == i) return i;
        const int gp = ipb_uf_load(L, p);
        if (gp == p) return p;
        atomicMin(&L[i], gp);
        i = gp;
    }
}
__device__ __forceinline__ void ipb_uf_union(int* L, int a, int b) {
    while (true) {
        a = ipb_uf_find(L, a);
        b = ipb_uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a);
        if (old == b) return;
        b = old;
    }
}

// every word (y, j) of the row band [fa_y0, fa_y0 + fa_nrow) of a crop, spread over the CTA
#define IPB_FA_FOREACH_WORD(crop, ...)                                                         \
    {                                                                                          \
        for (int i_ = threadIdx.x; i_ < fa_nrow * (crop).wpr; i_ += blockDim.x) {              \
            const int y = fa_y0 + i_ / (crop).wpr, j = i_ % (crop).wpr;                        \
            __VA_ARGS__                                                                        \
        }                                                                                      \
    }
// the band of one CTA of the multi-kernel path: IPB_FA_ROWS rows starting at blockIdx.x * IPB_FA_ROWS
#define IPB_FA_BAND(crop)                                                                      \
    const int fa_y0 = (int)blockIdx.x * IPB_FA_ROWS;                                           \
    if (fa_y0 >= (crop).h) return;                                                             \
    const int fa_nrow = (crop).h - fa_y0 > IPB_FA_ROWS ? IPB_FA_ROWS : (crop).h - fa_y0;

// ---------------------------------------------------------------- 1. threshold & mask
// a warp owns a row at a time; lane b reads pixel 32 j + b of four consecutive words (four
// independent coalesced 64-byte loads in flight), ballot -> word.
__device__ __forceinline__ void ipb_k_fa_threshold_phase(const IpbCrop& c, int fa_y0, int fa_nrow,
                                                         const unsigned short* planes, int H, int W, const float* fa_params,
                                                         const unsigned* roi_mask, unsigned* bw)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float thr = fa_params[(size_t)c.frame * 4 + 3];
    const unsigned short* img = planes + (size_t)c.plane * H * W;
    for (int r = warp; r < fa_nrow; r += nwarps) {
        const int y = fa_y0 + r;
        const unsigned short* irow = img + (size_t)(c.oy + y) * W + c.ox;
        const size_t w0 = (size_t)c.bit_off + (size_t)y * c.wpr, m0 = (size_t)c.mask_off + (size_t)y * c.wpr;
        for (int j0 = 0; j0 < c.wpr; j0 += 4) {
            unsigned short px[4];
            bool in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int x = 32 * (j0 + u) + lane;
                in[u] = (j0 + u < c.wpr) && x < c.w;
                px[u] = 0;
                if (in[u]) px[u] = irow[x];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned word = __ballot_sync(IPB_FULL, in[u] && (float)px[u] > thr);
                if (lane == 0 && j0 + u < c.wpr) bw[w0 + j0 + u] = word & roi_mask[m0 + j0 + u];
            }
        }
    }
}
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_fa_threshold(const IpbCrop* __restrict__ crops, const unsigned short* __restrict__ planes,
                   int H, int W, const float* __restrict__ fa_params /* [F][4], [3] = thr */,
                   const unsigned* __restrict__ roi_mask, unsigned* __restrict__ bw)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    ipb_k_fa_threshold_phase(c, fa_y0, fa_nrow, planes, H, W, fa_params, roi_mask, bw);
}

// ---------------------------------------------------------------- 2. CCL
__device__ __forceinline__ void ipb_k_ccl_init_phase(const IpbCrop& c, int fa_y0, int fa_nrow, const unsigned* bits,
               int* L, unsigned* csize)
{
    IPB_FA_FOREACH_WORD(c, {
        const unsigned* row = bits + c.bit_off + (size_t)y * c.wpr;
        unsigned s = ipb_bits_starts(row, j);
        while (s) {
            const int b = __ffs((int)s) - 1;
            s &= s - 1;
            const int idx = y * c.w + 32 * j + b;
            L[c.pix_off + idx] = idx;
            if (csize) csize[c.pix_off + idx] = 0u;
        }
    })
}
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_init(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
               int* __restrict__ L, unsigned* __restrict__ csize)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    ipb_k_ccl_init_phase(c, fa_y0, fa_nrow, bits, L, csize);
}

// unite every run starting in this word with the runs of the previous row it touches
template <int CONN>
__device__ __forceinline__ void ipb_k_ccl_merge_phase(const IpbCrop& c, int fa_y0, int fa_nrow, const unsigned* bits, int* L)
{
    int* Lc = L + c.pix_off;
    IPB_FA_FOREACH_WORD(c, {
        if (y > 0) {
            const unsigned* row = bits + c.bit_off + (size_t)y * c.wpr;
            const unsigned* prow = row - c.wpr;
            unsigned s = ipb_bits_starts(row, j);
            while (s) {
                const int b = __ffs((int)s) - 1;
                s &= s - 1;
                const int a = 32 * j + b;
                const int e = ipb_bits_next_clear(row, a, c.w) - 1;          // run = [a, e]
                int lo = a, hi = e;
                if (CONN == 8) { lo = a > 0 ? a - 1 : 0; hi = e + 1 < c.w ? e + 1 : c.w - 1; }
                int pos = lo;
                while (pos <= hi) {
                    const int t = ipb_bits_next_set(prow, pos, hi);
                    if (t < 0) break;
                    const int ps = ipb_bits_run_start(prow, t);
                    ipb_uf_union(Lc, y * c.w + a, (y - 1) * c.w + ps);
                    pos = ipb_bits_next_clear(prow, t, c.w) + 1;
                }
            }
        }
    })
}
template <int CONN>
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_merge(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits, int* __restrict__ L)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    ipb_k_ccl_merge_phase<CONN>(c, fa_y0, fa_nrow, bits, L);
}

// L[start] = root ; component size accumulated per run (4-conn pass: remove_small_objects)
__device__ __forceinline__ void ipb_k_ccl_flatten_size_phase(const IpbCrop& c, int fa_y0, int fa_nrow, const unsigned* bits,
                       int* L, unsigned* csize)
{
    int* Lc = L + c.pix_off;
    IPB_FA_FOREACH_WORD(c, {
        const unsigned* row = bits + c.bit_off + (size_t)y * c.wpr;
        unsigned s = ipb_bits_starts(row, j);
        while (s) {
            const int b = __ffs((int)s) - 1;
            s &= s - 1;
            const int a = 32 * j + b;
            const int e = ipb_bits_next_clear(row, a, c.w);
            const int r = ipb_uf_find(Lc, y * c.w + a);
            Lc[y * c.w + a] = r;
            atomicAdd(&csize[c.pix_off + r], (unsigned)(e - a));
        }
    })
}
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_flatten_size(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
                       int* __restrict__ L, unsigned* __restrict__ csize)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    ipb_k_ccl_flatten_size_phase(c, fa_y0, fa_nrow, bits, L, csize);
}

// keep a pixel iff its component has size >= min_size (skimage: sizes < min_size removed)
__device__ __forceinline__ void ipb_k_fa_size_filter_phase(const IpbCrop& c, int fa_y0, int fa_nrow, const unsigned* bits,
                     const int* L, const unsigned* csize, double min_size,
                     unsigned* out)
{
    IPB_FA_FOREACH_WORD(c, {
        const unsigned* row = bits + c.bit_off + (size_t)y * c.wpr;
        const unsigned wv = row[j];
        unsigned keep = 0u, todo = wv;
        while (todo) {
            const int b = __ffs((int)todo) - 1;
            const int x = 32 * j + b;
            const int a = ipb_bits_run_start(row, x);
            const int r = L[c.pix_off + y * c.w + a];                 // flattened: root of the run
            const bool ok = !((double)csize[c.pix_off + r] < min_size);
            // all bits of this run inside the word share the verdict
            int e = ipb_bits_next_clear(row, x, c.w);
            if (e > 32 * j + 32) e = 32 * j + 32;
            const int nb = e - x;
            const unsigned seg = (nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u)) << b;
            if (ok) keep |= seg;
            todo &= ~seg;
        }
        out[(size_t)c.bit_off + (size_t)y * c.wpr + j] = keep;
    })
}
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_fa_size_filter(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
                     const int* __restrict__ L, const unsigned* __restrict__ csize, double min_size,
                     unsigned* __restrict__ out)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    ipb_k_fa_size_filter_phase(c, fa_y0, fa_nrow, bits, L, csize, min_size, out);
}

// ---------------------------------------------------------------- 3. morphology with disk(r)
// OP = 0 dilation (outside = 0), OP = 1 erosion (outside = 1): scipy.ndimage semantics used by
// skimage.morphology.binary_closing.  halfw[dy + r] = floor(sqrt(r^2 - dy^2)).
struct IpbDisk { int r; int halfw[11]; };

__device__ __forceinline__ unsigned ipb_row_word(const unsigned* bits, const IpbCrop& c, int y, int j, unsigned outside) {
    if (y < 0 || y >= c.h || j < 0 || j >= c.wpr) return outside;
    unsigned v = bits[c.bit_off + (size_t)y * c.wpr + j];
    if (outside && j == c.wpr - 1 && (c.w & 31)) v |= ~((1u << (c.w & 31)) - 1u);   // pixels beyond w count as 1
    return v;
}

template <int OP>
__device__ __forceinline__ void ipb_k_bits_morph_phase(const IpbCrop& c, int fa_y0, int fa_nrow, const unsigned* in, IpbDisk disk,
                 unsigned* out)
{
    const unsigned outside = OP ? 0xffffffffu : 0u;
    IPB_FA_FOREACH_WORD(c, {
        unsigned acc = OP ? 0xffffffffu : 0u;
        for (int dy = -disk.r; dy <= disk.r; ++dy) {
            const int k = disk.halfw[dy + disk.r];
            const unsigned m = ipb_row_word(in, c, y + dy, j, outside);
            // the neighbouring words only matter where the disk row is wider than one pixel
            const unsigned l = k > 0 ? ipb_row_word(in, c, y + dy, j - 1, outside) : 0u;
            const unsigned r = k > 0 ? ipb_row_word(in, c, y + dy, j + 1, outside) : 0u;
            unsigned v = m;
            for (int sft = 1; sft <= k; ++sft) {
                const unsigned from_left = (m << sft) | (l >> (32 - sft));     // pixel x - sft
                const unsigned from_right = (m >> sft) | (r << (32 - sft));    // pixel x + sft
                if (OP) v &= from_left & from_right; else v |= from_left | from_right;
            }
            if (OP) acc &= v; else acc |= v;
        }
        if (j == c.wpr - 1 && (c.w & 31)) acc &= (1u << (c.w & 31)) - 1u;
        out[(size_t)c.bit_off + (size_t)y * c.wpr + j] = acc;
    })
}
template <int OP>
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_bits_morph(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ in, IpbDisk disk,
                 unsigned* __restrict__ out)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    ipb_k_bits_morph_phase<OP>(c, fa_y0, fa_nrow, in, disk, out);
}

struct IpbComp {               // one row of the per-adhesion table (exact integer sums)
    unsigned long long sum_i;  // sum of raw intensities
    unsigned long long sum_y;  // sum of row coordinates (crop-local)
    unsigned long long sum_x;  // sum of column coordinates (crop-local)
    unsigned area;
    int crop;
};

#include "ipb_fa_smem.cuh"

// ---------------------------------------------------------------- 4. labels & regionprops
// L[start] = root ; root bits ; roots per row
__device__ __forceinline__ void ipb_k_ccl_flatten_roots_phase(const IpbCrop& c, int fa_y0, int fa_nrow, const unsigned* bits,
                        int* L, unsigned* rootbits, int* row_roots)
{
    int* Lc = L + c.pix_off;
    IPB_FA_FOREACH_WORD(c, {
        const unsigned* row = bits + c.bit_off + (size_t)y * c.wpr;
        unsigned s = ipb_bits_starts(row, j);
        unsigned rb = 0u;
        while (s) {
            const int b = __ffs((int)s) - 1;
            s &= s - 1;
            const int idx = y * c.w + 32 * j + b;
            const int r = ipb_uf_find(Lc, idx);
            Lc[idx] = r;
            if (r == idx) rb |= 1u << b;
        }
        rootbits[(size_t)c.bit_off + (size_t)y * c.wpr + j] = rb;
        if (rb) atomicAdd(&row_roots[c.row_off + y], __popc(rb));
    })
}
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_flatten_roots(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
                        int* __restrict__ L, unsigned* __restrict__ rootbits, int* __restrict__ row_roots)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    ipb_k_ccl_flatten_roots_phase(c, fa_y0, fa_nrow, bits, L, rootbits, row_roots);
}

// per crop: exclusive scan of roots per row -> row_base ; crop_count[crop] = total
__global__ void __launch_bounds__(256)
ipb_k_fa_row_scan(const IpbCrop* __restrict__ crops, const int* __restrict__ row_roots,
                  int* __restrict__ row_base, int* __restrict__ crop_count)
{
    const IpbCrop c = crops[blockIdx.x];
    __shared__ int wsum[8];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int y0 = 0; y0 < c.h; y0 += 256) {
        const int y = y0 + threadIdx.x;
        const int v = y < c.h ? row_roots[c.row_off + y] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int base = carry;
        for (int i = 0; i < warp; ++i) base += wsum[i];
        if (y < c.h) row_base[c.row_off + y] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) crop_count[blockIdx.x] = carry;
}

// ---------------------------------------------------------------- fused per-crop chain
// One CTA per crop runs every phase up to the per-row root counts back to back (phases are
// separated by __syncthreads; the union-find works through global memory, which the barrier
// makes visible inside the CTA).  Used when the batch has many small crops (cell ROIs); big
// single crops (the 8192^2 mosaic) take the multi-kernel path, which spreads one crop over
// the whole chip.
#define IPB_FA_FUSED_THREADS 512
__global__ void __launch_bounds__(IPB_FA_FUSED_THREADS)
ipb_k_fa_fused(const IpbCrop* __restrict__ crops, const unsigned short* __restrict__ planes, int H, int W,
               const float* __restrict__ fa_params, const unsigned* __restrict__ roi_mask,
               double min_size, IpbDisk disk, unsigned* bw_a, unsigned* bw_b, int* L, unsigned* csize,
               unsigned* rootbits, int* row_roots, int* row_base, int* crop_count, unsigned* bw_final,
               const int* __restrict__ order /* nullable: CTA b runs crop order[b] (biggest first) */,
               int only_flagged /* != 0: only crops the shared-memory kernel flagged (crop_count == -1) */)
{
    const int ci = order ? order[blockIdx.x] : (int)blockIdx.x;
    if (only_flagged && crop_count[ci] != -1) return;
    const IpbCrop c = crops[ci];
    const int fa_y0 = 0, fa_nrow = c.h;
    __shared__ int wsum[IPB_FA_FUSED_THREADS / 32];
    __shared__ int carry;
    ipb_k_fa_threshold_phase(c, fa_y0, fa_nrow, planes, H, W, fa_params, roi_mask, bw_a);
    __syncthreads();
    unsigned* cur = bw_a;
    unsigned* other = bw_b;
    if (min_size > 0) {
        ipb_k_ccl_init_phase(c, fa_y0, fa_nrow, cur, L, csize);
        __syncthreads();
        ipb_k_ccl_merge_phase<4>(c, fa_y0, fa_nrow, cur, L);
        __syncthreads();
        ipb_k_ccl_flatten_size_phase(c, fa_y0, fa_nrow, cur, L, csize);
        __syncthreads();
        ipb_k_fa_size_filter_phase(c, fa_y0, fa_nrow, cur, L, csize, min_size, other);
        __syncthreads();
        unsigned* t = cur; cur = other; other = t;
    }
    if (disk.r > 0) {
        ipb_k_bits_morph_phase<0>(c, fa_y0, fa_nrow, cur, disk, other);
        __syncthreads();
        ipb_k_bits_morph_phase<1>(c, fa_y0, fa_nrow, other, disk, bw_final);
    } else {
        ipb_k_bits_morph_phase<0>(c, fa_y0, fa_nrow, cur, disk, bw_final);
    }
    for (int y = threadIdx.x; y < c.h; y += blockDim.x) row_roots[c.row_off + y] = 0;
    __syncthreads();
    ipb_k_ccl_init_phase(c, fa_y0, fa_nrow, bw_final, L, (unsigned*)nullptr);
    __syncthreads();
    ipb_k_ccl_merge_phase<8>(c, fa_y0, fa_nrow, bw_final, L);
    __syncthreads();
    ipb_k_ccl_flatten_roots_phase(c, fa_y0, fa_nrow, bw_final, L, rootbits, row_roots);
    __syncthreads();
    // exclusive scan of roots per row -> row_base ; crop_count[crop] = total
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int y0 = 0; y0 < c.h; y0 += blockDim.x) {
        const int y = y0 + threadIdx.x;
        const int v = y < c.h ? row_roots[c.row_off + y] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int base = carry;
        for (int i = 0; i < warp; ++i) base += wsum[i];
        if (y < c.h) row_base[c.row_off + y] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = base + incl;
        __syncthreads();
    }
    (void)nw;
    if (threadIdx.x == 0) crop_count[ci] = carry;
}

// comp_off = exclusive scan of crop_count (single CTA), comp_off[n] = total
__global__ void __launch_bounds__(256)
ipb_k_fa_crop_scan(const int* __restrict__ crop_count, int n, int* __restrict__ comp_off)
{
    __shared__ int wsum[8];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const int v = i < n ? crop_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(IPB_FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int base = carry;
        for (int k = 0; k < warp; ++k) base += wsum[k];
        if (i < n) comp_off[i] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) comp_off[n] = carry;
}

__global__ void ipb_k_fa_zero_comps(const int* __restrict__ comp_off, int n_crops, int cap, IpbComp* __restrict__ comps)
{
    int total = comp_off[n_crops];
    if (total > cap) total = cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        IpbComp z; z.sum_i = 0; z.sum_y = 0; z.sum_x = 0; z.area = 0; z.crop = -1;
        comps[i] = z;
    }
}

// sequential id (0-based within the crop) of the component rooted at crop-local pixel idx
__device__ __forceinline__ int ipb_fa_root_rank(const IpbCrop& c, const unsigned* rootbits,
                                                const int* row_base, int ridx) {
    const int ry = ridx / c.w, rx = ridx % c.w;
    const unsigned* rrow = rootbits + c.bit_off + (size_t)ry * c.wpr;
    int rank = row_base[c.row_off + ry];
    const int jw = rx >> 5;
    for (int k = 0; k < jw; ++k) rank += __popc(rrow[k]);
    rank += __popc(rrow[jw] & ((1u << (rx & 31)) - 1u));
    return rank;
}

// per run: area, sum of intensities, sum of coordinates -> compact component table;
// optionally the int32 label map (1-based ids, 0 background) of every pixel of the run
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_fa_props(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
               const int* __restrict__ L, const unsigned* __restrict__ rootbits,
               const int* __restrict__ row_base, const int* __restrict__ comp_off, int cap,
               const unsigned short* __restrict__ planes, int H, int W,
               IpbComp* __restrict__ comps, int* __restrict__ labels /* nullable */)
{
    const IpbCrop c = crops[blockIdx.y];
    IPB_FA_BAND(c);
    if (row_base[c.row_off] == IPB_FAS_MARK) return;   // finished by ipb_k_fa_fused_smem / ipb_k_fa_gather_smem
    const unsigned short* img = planes + (size_t)c.plane * H * W;
    IPB_FA_FOREACH_WORD(c, {
        const unsigned* row = bits + c.bit_off + (size_t)y * c.wpr;
        if (labels) {                                  // background of this word
            const unsigned wv = row[j];
            for (int b = 0; b < 32 && 32 * j + b < c.w; ++b)
                if (!((wv >> b) & 1u)) labels[c.pix_off + (size_t)y * c.w + 32 * j + b] = 0;
        }
        unsigned s = ipb_bits_starts(row, j);
        while (s) {
            const int b = __ffs((int)s) - 1;
            s &= s - 1;
            const int a = 32 * j + b;
            const int e = ipb_bits_next_clear(row, a, c.w);                     // run = [a, e)
            const int r = L[c.pix_off + y * c.w + a];
            const int rank = ipb_fa_root_rank(c, rootbits, row_base, r);
            const int cid = comp_off[blockIdx.y] + rank;
            unsigned long long si = 0;
            const unsigned short* irow = img + (size_t)(c.oy + y) * W + c.ox;
            for (int x = a; x < e; ++x) si += irow[x];
            if (labels) for (int x = a; x < e; ++x) labels[c.pix_off + (size_t)y * c.w + x] = rank + 1;
            if (cid < cap) {
                const unsigned len = (unsigned)(e - a);
                atomicAdd(&comps[cid].area, len);
                atomicAdd(&comps[cid].sum_i, si);
                atomicAdd(&comps[cid].sum_y, (unsigned long long)len * (unsigned long long)y);
                atomicAdd(&comps[cid].sum_x, (unsigned long long)(a + e - 1) * len / 2ull);
                comps[cid].crop = (int)blockIdx.y;
            }
        }
    })
}
