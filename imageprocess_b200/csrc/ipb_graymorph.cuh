// ipb_graymorph.cuh -- flat grey-scale erosion / dilation on uint16 planes with a (2r+1) x (2r+1)
// square footprint, separable (min / max along axis 0, then along axis 1), on shared-memory halo
// tiles filled by TMA bulk copies.  scipy.ndimage.grey_erosion / grey_dilation(input, size=2r+1)
// with the default mode 'reflect'; white_tophat = input - dilation(erosion(input)).
// The reference has no top-hat stage (BASELINE.json's north_star names one): optional,
// OFF by default, oracle = scipy.ndimage.white_tophat ("parity unpinned by reference").
// Same tile scheme as ipb_gauss.cuh; min / max are exact, so the result equals scipy's bit for bit.
#pragma once
#include "ipb_gauss.cuh"

#define IPB_GM_TILE_W 256               // output columns per tile, multiple of 8 (16 bytes of uint16)
#define IPB_GM_TILE_H0 32
#define IPB_GM_TILE_H1 8
#define IPB_GM_MAXR 64

template <int AXIS, bool MAX>
__global__ void __launch_bounds__(IPB_GS_THREADS)
ipb_k_graymorph_pass(const unsigned short* __restrict__ in, unsigned short* __restrict__ out, int H, int W, int radius, int use_tma)
{
    IPB_DYN_SMEM(unsigned short, tile);
    __shared__ IpbMbar bar;
    const int tid = threadIdx.x;
    const size_t img_off = (size_t)blockIdx.z * H * W;
    const unsigned short* src = in + img_off;
    unsigned short* dst = out + img_off;
    const int x0 = blockIdx.x * IPB_GM_TILE_W;
    const int tw = min(IPB_GM_TILE_W, W - x0);
    auto pick = [](unsigned a, unsigned b) { return MAX ? (a > b ? a : b) : (a < b ? a : b); };
    if (AXIS == 0) {
        const int y0 = blockIdx.y * IPB_GM_TILE_H0;
        const int th = min(IPB_GM_TILE_H0, H - y0);
        const int rows = th + 2 * radius, pitch = IPB_GM_TILE_W;
        if (use_tma) {
            if (tid == 0) { ipb_mbar_init(&bar, 1u); ipb_mbar_fence_init(); }
            __syncthreads();
            if (tid == 0) {
                ipb_mbar_expect(&bar, 2u * (unsigned)tw * (unsigned)rows);
                for (int r = 0; r < rows; ++r)
                    ipb_bulk_copy(tile + (size_t)r * pitch, src + (size_t)ipb_reflect(y0 - radius + r, H) * W + x0,
                                  2u * (unsigned)tw, &bar);
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
            const unsigned short* col = tile + (size_t)r * pitch + c;          // window rows r .. r + 2 radius
            unsigned v = col[0];
            for (int j = 1; j <= 2 * radius; ++j) v = pick(v, (unsigned)col[(size_t)j * pitch]);
            dst[(size_t)(y0 + r) * W + x0 + c] = (unsigned short)v;
        }
    } else {
        const int y0 = blockIdx.y * IPB_GM_TILE_H1;
        const int th = min(IPB_GM_TILE_H1, H - y0);
        const int hr = (radius + 7) & ~7;
        const int pitch = IPB_GM_TILE_W + 2 * hr;
        const int lo = max(x0 - hr, 0), hi = min(x0 + tw + hr, W);
        if (use_tma) {
            if (tid == 0) { ipb_mbar_init(&bar, 1u); ipb_mbar_fence_init(); }
            __syncthreads();
            if (tid == 0) {
                ipb_mbar_expect(&bar, 2u * (unsigned)(hi - lo) * (unsigned)th);
                for (int r = 0; r < th; ++r)
                    ipb_bulk_copy(tile + (size_t)r * pitch + (lo - (x0 - hr)), src + (size_t)(y0 + r) * W + lo,
                                  2u * (unsigned)(hi - lo), &bar);
            }
            ipb_mbar_wait(&bar, 0u);
        } else {
            for (int i = tid; i < th * (hi - lo); i += blockDim.x) {
                const int r = i / (hi - lo), c = i - r * (hi - lo);
                tile[(size_t)r * pitch + (lo - (x0 - hr)) + c] = src[(size_t)(y0 + r) * W + lo + c];
            }
            __syncthreads();
        }
        for (int i = tid; i < th * tw; i += blockDim.x) {
            const int r = i / tw, c = i - r * tw;
            const int x = x0 + c;
            const unsigned short* row = tile + (size_t)r * pitch - (x0 - hr);
            const unsigned short* grow = src + (size_t)(y0 + r) * W;
            unsigned v = MAX ? 0u : 0xffffu;
            for (int j = -radius; j <= radius; ++j) {
                const int xx = x + j;
                const int q = (xx >= 0 && xx < W) ? xx : ipb_reflect(xx, W);
                v = pick(v, (unsigned)((q >= lo && q < hi) ? row[q] : grow[q]));
            }
            dst[(size_t)(y0 + r) * W + x] = (unsigned short)v;
        }
    }
}

// out = a - b on uint16 (white top-hat: the opening never exceeds the input)
__global__ void ipb_k_sub_u16(const unsigned short* __restrict__ a, const unsigned short* __restrict__ b,
                              unsigned short* __restrict__ out, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (unsigned short)(a[i] - b[i]);
}
