// ipb_gauss.cuh -- separable Gaussian filter on shared-memory halo tiles loaded by TMA
// (the band-pass / unsharp display filters of the interactive ROI drawer,
// roi_manual_drawer.py:870-876, ndi.gaussian_filter(im, sigma); also the optional, default-OFF
// Gaussian pre-filter stage BASELINE.json's north_star names and the reference does not have).
//
// scipy.ndimage.gaussian_filter = one correlate1d per axis (last axis first is NOT scipy's order:
// it filters axis 0, then axis 1), mode 'reflect' (d c b a | a b c d | d c b a), weights
// exp(-x^2 / (2 sigma^2)) / sum for |x| <= int(4 sigma + 0.5), every line converted to float64,
// the symmetric-kernel loop of ni_filters.c NI_Correlate1D:
//     tmp = line[i] * w[0];  for j = radius .. 1:  tmp += (line[i - j] + line[i + j]) * w[j]
// and the result of each pass stored in the array's dtype (float32 here).  The kernels below do
// exactly that arithmetic (float64, separate multiply and add, same order), so the output equals
// scipy's bit for bit; the weights come from the host, computed with numpy as scipy does.
//
// Data movement: a CTA owns a tile of the output.  One elected thread starts one TMA bulk copy per
// tile row (cp.async.bulk, global -> shared, 16-byte aligned segments including the halo) on one
// mbarrier; the block waits once and computes from shared memory.  Reflection at the image border:
//   axis 0 (columns of the tile walk rows): the halo ROWS are fetched from the reflected row index,
//                                           so the tile is complete after the copies;
//   axis 1: the copies are clamped to the row and the reflected columns are read through an index map.
#pragma once
#include "ipb_rt.cuh"

#define IPB_GS_THREADS 256
#define IPB_GS_MAXR 160                 // largest kernel radius served (sigma 40): the axis-0 tile with its halo stays below 192 KB
#define IPB_GS_TILE_W 128               // output columns per tile (both passes), multiple of 4
#define IPB_GS_TILE_H0 32               // output rows per tile of the axis-0 pass
#define IPB_GS_TILE_H1 8                // output rows per tile of the axis-1 pass

__device__ __forceinline__ int ipb_reflect(int i, int n) {      // scipy 'reflect': -1 -> 0, -2 -> 1, n -> n-1, ...
    if (n == 1) return 0;
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

// AXIS = 0: out[y][x] = sum_j w[|j|] in[reflect(y + j)][x];  AXIS = 1: along x.
// in / out: float32 [n_images][H][W], W % 4 == 0 and 16-byte aligned images for the TMA path.
// Dynamic shared memory: rows x pitch floats (pitch = tile width incl. halo, multiple of 4).
template <int AXIS>
__global__ void __launch_bounds__(IPB_GS_THREADS)
ipb_k_gauss_pass(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                 const double* __restrict__ weights /* [radius + 1]: centre, then offsets 1..radius */, int radius, int use_tma)
{
    IPB_DYN_SMEM(float, tile);
    __shared__ IpbMbar bar;
    __shared__ double w[IPB_GS_MAXR + 1];
    const int tid = threadIdx.x;
    const size_t img_off = (size_t)blockIdx.z * H * W;
    const float* src = in + img_off;
    float* dst = out + img_off;
    for (int i = tid; i <= radius; i += blockDim.x) w[i] = weights[i];
    const int x0 = blockIdx.x * IPB_GS_TILE_W;
    const int tw = min(IPB_GS_TILE_W, W - x0);                        // output columns of this tile
    if (AXIS == 0) {
        const int y0 = blockIdx.y * IPB_GS_TILE_H0;
        const int th = min(IPB_GS_TILE_H0, H - y0);
        const int rows = th + 2 * radius, pitch = IPB_GS_TILE_W;
        // tile row r <-> image row reflect(y0 - radius + r), columns x0 .. x0 + tw
        if (use_tma) {
            if (tid == 0) { ipb_mbar_init(&bar, 1u); ipb_mbar_fence_init(); }
            __syncthreads();
            if (tid == 0) {
                ipb_mbar_expect(&bar, 4u * (unsigned)tw * (unsigned)rows);
                for (int r = 0; r < rows; ++r)
                    ipb_bulk_copy(tile + (size_t)r * pitch, src + (size_t)ipb_reflect(y0 - radius + r, H) * W + x0,
                                  4u * (unsigned)tw, &bar);
            }
            ipb_mbar_wait(&bar, 0u);
        } else {
            for (int i = tid; i < rows * tw; i += blockDim.x) {
                const int r = i / tw, c = i - r * tw;
                tile[(size_t)r * pitch + c] = src[(size_t)ipb_reflect(y0 - radius + r, H) * W + x0 + c];
            }
            __syncthreads();
        }
        for (int i = tid; i < th * tw; i += blockDim.x) {
            const int r = i / tw, c = i - r * tw;
            const float* col = tile + (size_t)(r + radius) * pitch + c;
            double acc = __dmul_rn((double)col[0], w[0]);
            for (int j = radius; j >= 1; --j)
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)col[-(ptrdiff_t)j * pitch], (double)col[(ptrdiff_t)j * pitch]), w[j]));
            dst[(size_t)(y0 + r) * W + x0 + c] = (float)acc;
        }
    } else {
        const int y0 = blockIdx.y * IPB_GS_TILE_H1;
        const int th = min(IPB_GS_TILE_H1, H - y0);
        const int hr = (radius + 3) & ~3;                              // halo rounded up to 16 bytes
        const int pitch = IPB_GS_TILE_W + 2 * hr;
        // tile column c <-> image column x0 - hr + c; only the columns inside the image are loaded
        const int lo = max(x0 - hr, 0), hi = min(x0 + tw + hr, W);     // [lo, hi): multiples of 4 (W % 4 == 0)
        if (use_tma) {
            if (tid == 0) { ipb_mbar_init(&bar, 1u); ipb_mbar_fence_init(); }
            __syncthreads();
            if (tid == 0) {
                ipb_mbar_expect(&bar, 4u * (unsigned)(hi - lo) * (unsigned)th);
                for (int r = 0; r < th; ++r)
                    ipb_bulk_copy(tile + (size_t)r * pitch + (lo - (x0 - hr)), src + (size_t)(y0 + r) * W + lo,
                                  4u * (unsigned)(hi - lo), &bar);
            }
            ipb_mbar_wait(&bar, 0u);
        } else {
            for (int i = tid; i < th * (hi - lo); i += blockDim.x) {
                const int r = i / (hi - lo), c = i - r * (hi - lo);
                tile[(size_t)r * pitch + (lo - (x0 - hr)) + c] = src[(size_t)(y0 + r) * W + lo + c];
            }
            __syncthreads();
        }
        // a reflected column lies inside [lo, hi) whenever the radius does not exceed the image
        // width (the caller checks); else it is read from global memory
        for (int i = tid; i < th * tw; i += blockDim.x) {
            const int r = i / tw, c = i - r * tw;
            const int x = x0 + c;
            const float* row = tile + (size_t)r * pitch - (x0 - hr);   // row[xx] = image column xx of this tile row
            const float* grow = src + (size_t)(y0 + r) * W;
            auto at = [&](int xx) -> double {
                const int q = (xx >= 0 && xx < W) ? xx : ipb_reflect(xx, W);
                return (double)((q >= lo && q < hi) ? row[q] : grow[q]);
            };
            double acc = __dmul_rn(at(x), w[0]);
            for (int j = radius; j >= 1; --j)
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(at(x - j), at(x + j)), w[j]));
            dst[(size_t)(y0 + r) * W + x] = (float)acc;
        }
    }
}

// out = a - b (band-pass: gaussian(small) - gaussian(large)) or out = a + amount * (a - b) (unsharp:
// im + amount * (im - gaussian(im, r))), float32 ops rounded separately as numpy evaluates them
__global__ void ipb_k_gauss_combine(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                    long long n, int unsharp, float amount)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = __fsub_rn(a[i], b[i]);
        out[i] = unsharp ? __fadd_rn(a[i], __fmul_rn(amount, d)) : d;
    }
}

// uint16 <-> float32 planes for the optional Gaussian pre-filter of an integer channel:
// to_f32: out = float32(in);  to_u16: out = uint16(clip(rint(in), 0, 65535)) (NaN -> 0)
__global__ void ipb_k_u16_to_f32(const unsigned short* __restrict__ in, float* __restrict__ out, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}
__global__ void ipb_k_f32_to_u16(const float* __restrict__ in, unsigned short* __restrict__ out, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = rintf(in[i]);                                    // round half to even, as np.rint
        v = v > 0.0f ? v : 0.0f;                                   // NaN and negatives -> 0
        out[i] = (unsigned short)(v < 65535.0f ? v : 65535.0f);
    }
}
