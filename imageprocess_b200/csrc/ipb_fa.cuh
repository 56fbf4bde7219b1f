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
__device__ __forceinline__ int ipb_uf_find(const int* L, int i) {
    int p = ipb_uf_load(L, i);
    while (p != i) { i = p; p = ipb_uf_load(L, i); }
    return i;
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

#define IPB_FA_FOREACH_WORD(crop, ...)                                                         \
    {                                                                                          \
        const int y_beg_ = (int)blockIdx.x * IPB_FA_ROWS;                                      \
        int nrow_ = (crop).h - y_beg_;                                                         \
        if (nrow_ > IPB_FA_ROWS) nrow_ = IPB_FA_ROWS;                                          \
        for (int i_ = threadIdx.x; i_ < nrow_ * (crop).wpr; i_ += blockDim.x) {                \
            const int y = y_beg_ + i_ / (crop).wpr, j = i_ % (crop).wpr;                       \
            __VA_ARGS__                                                                        \
        }                                                                                      \
    }

// ---------------------------------------------------------------- 1. threshold & mask
// warp per word: lane b reads pixel 32 j + b (coalesced 64 B), ballot -> word.
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_fa_threshold(const IpbCrop* __restrict__ crops, const unsigned short* __restrict__ planes,
                   int H, int W, const float* __restrict__ fa_params /* [F][4], [3] = thr */,
                   const unsigned* __restrict__ roi_mask, unsigned* __restrict__ bw)
{
    const IpbCrop c = crops[blockIdx.y];
    const int y_beg = (int)blockIdx.x * IPB_FA_ROWS;
    if (y_beg >= c.h) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    int nrow = c.h - y_beg;
    if (nrow > IPB_FA_ROWS) nrow = IPB_FA_ROWS;
    const float thr = fa_params[(size_t)c.frame * 4 + 3];
    const unsigned short* img = planes + (size_t)c.plane * H * W;
    const int nitems = nrow * c.wpr;
    // 4 words per warp-iteration: four independent coalesced pixel loads in flight per lane
    for (int i0 = warp * 4; i0 < nitems; i0 += nwarps * 4) {
        unsigned short px[4];
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            in[u] = false;
            px[u] = 0;
            if (i < nitems) {
                const int y = y_beg + i / c.wpr, j = i % c.wpr, x = 32 * j + lane;
                in[u] = x < c.w;
                if (in[u]) px[u] = img[(size_t)(c.oy + y) * W + (c.ox + x)];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u;
            const unsigned word = __ballot_sync(IPB_FULL, in[u] && (float)px[u] > thr);
            if (lane == 0 && i < nitems) {
                const int y = y_beg + i / c.wpr, j = i % c.wpr;
                const size_t wi = (size_t)c.bit_off + (size_t)y * c.wpr + j;
                bw[wi] = word & roi_mask[(size_t)c.mask_off + (size_t)y * c.wpr + j];
            }
        }
    }
}

// ---------------------------------------------------------------- 2. CCL
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_init(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
               int* __restrict__ L, unsigned* __restrict__ csize)
{
    const IpbCrop c = crops[blockIdx.y];
    if ((int)blockIdx.x * IPB_FA_ROWS >= c.h) return;
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

// unite every run starting in this word with the runs of the previous row it touches
template <int CONN>
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_merge(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits, int* __restrict__ L)
{
    const IpbCrop c = crops[blockIdx.y];
    if ((int)blockIdx.x * IPB_FA_ROWS >= c.h) return;
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

// L[start] = root ; component size accumulated per run (4-conn pass: remove_small_objects)
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_flatten_size(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
                       int* __restrict__ L, unsigned* __restrict__ csize)
{
    const IpbCrop c = crops[blockIdx.y];
    if ((int)blockIdx.x * IPB_FA_ROWS >= c.h) return;
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

// keep a pixel iff its component has size >= min_size (skimage: sizes < min_size removed)
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_fa_size_filter(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
                     const int* __restrict__ L, const unsigned* __restrict__ csize, double min_size,
                     unsigned* __restrict__ out)
{
    const IpbCrop c = crops[blockIdx.y];
    if ((int)blockIdx.x * IPB_FA_ROWS >= c.h) return;
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
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_bits_morph(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ in, IpbDisk disk,
                 unsigned* __restrict__ out)
{
    const IpbCrop c = crops[blockIdx.y];
    if ((int)blockIdx.x * IPB_FA_ROWS >= c.h) return;
    const unsigned outside = OP ? 0xffffffffu : 0u;
    IPB_FA_FOREACH_WORD(c, {
        unsigned acc = OP ? 0xffffffffu : 0u;
        for (int dy = -disk.r; dy <= disk.r; ++dy) {
            const int k = disk.halfw[dy + disk.r];
            const unsigned m = ipb_row_word(in, c, y + dy, j, outside);
            const unsigned l = ipb_row_word(in, c, y + dy, j - 1, outside);
            const unsigned r = ipb_row_word(in, c, y + dy, j + 1, outside);
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

// ---------------------------------------------------------------- 4. labels & regionprops
// L[start] = root ; root bits ; roots per row
__global__ void __launch_bounds__(IPB_FA_THREADS)
ipb_k_ccl_flatten_roots(const IpbCrop* __restrict__ crops, const unsigned* __restrict__ bits,
                        int* __restrict__ L, unsigned* __restrict__ rootbits, int* __restrict__ row_roots)
{
    const IpbCrop c = crops[blockIdx.y];
    if ((int)blockIdx.x * IPB_FA_ROWS >= c.h) return;
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

struct IpbComp {               // one row of the per-adhesion table (exact integer sums)
    unsigned long long sum_i;  // sum of raw intensities
    unsigned long long sum_y;  // sum of row coordinates (crop-local)
    unsigned long long sum_x;  // sum of column coordinates (crop-local)
    unsigned area;
    int crop;
};

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
    if ((int)blockIdx.x * IPB_FA_ROWS >= c.h) return;
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
