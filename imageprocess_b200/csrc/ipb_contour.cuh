// ipb_contour.cuh -- marching-squares cells of every labelled adhesion of a crop in ONE pass
// (SURVEY.md 8(f) item 2).  The reference calls skimage.measure.find_contours(labeled_img == k, 0.5)
// once PER ADHESION over the whole crop (INT/FA_Analyzer.py:166-170): O(adhesions x crop pixels).
// For a binary mask at level 0.5 every contour point is the midpoint of a pixel pair, so a contour
// is fully described by the 2 x 2 cells its mask cuts and their 4-bit case.  One thread per cell
// looks at the (up to four) distinct labels of the cell and emits, per label, a record
//     { cell index r0 * w + c0,  label << 4 | case }        case: bit 0 ul, 1 ur, 2 ll, 3 lr
// for the cases 1..14.  The host orders a crop's records by (label, cell) -- the raster order in
// which skimage's _get_contour_segments visits the cells of that label's mask -- and links the
// segments exactly as skimage's _assemble_contours does (imageprocess_b200/contours.py).
#pragma once
#include "ipb_fa.cuh"

// labels: int32 [total_px] crop-local label maps (0 background) at IpbCrop.pix_off
// rec:    uint2 [total_px] record slices at pix_off; rec_count: uint32 [n_crops] (zeroed by the caller)
__global__ void __launch_bounds__(256)
ipb_k_fa_contour_cells(const IpbCrop* __restrict__ crops, const int* __restrict__ labels,
                       uint2* __restrict__ rec, unsigned* __restrict__ rec_count)
{
    const IpbCrop c = crops[blockIdx.y];
    if (c.w < 2 || c.h < 2) return;
    const int* L = labels + c.pix_off;
    uint2* R = rec + c.pix_off;
    const long long cap = (long long)c.w * c.h;
    const long long ncell = (long long)(c.w - 1) * (c.h - 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) {
        const int r0 = (int)(i / (c.w - 1)), c0 = (int)(i - (long long)r0 * (c.w - 1));
        const int* p = L + (size_t)r0 * c.w + c0;
        const int v[4] = {p[0], p[1], p[c.w], p[c.w + 1]};             // ul, ur, ll, lr
        if (v[0] == v[1] && v[0] == v[2] && v[0] == v[3]) continue;      // uniform cell: no label is cut here
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int k = v[a];
            if (k <= 0) continue;
            bool first = true;
#pragma unroll
            for (int b = 0; b < 4; ++b) if (b < a && v[b] == k) first = false;
            if (!first) continue;                                        // every distinct label once
            const unsigned cs = (v[0] == k ? 1u : 0u) | (v[1] == k ? 2u : 0u) | (v[2] == k ? 4u : 0u) | (v[3] == k ? 8u : 0u);
            if (cs == 15u) continue;
            const unsigned slot = atomicAdd(&rec_count[blockIdx.y], 1u);
            if ((long long)slot < cap) R[slot] = make_uint2((unsigned)(r0 * c.w + c0), ((unsigned)k << 4) | cs);
        }
    }
}
