// ipb_morph.cuh -- region morphology and shape sums on bit-packed region masks
// (SURVEY.md 8(a) a12, a13, a17):
//
//   * dilation of a region mask by a symmetric, row-convex structuring element given as a
//     table  gmax[|dx|] = largest vertical offset covered at horizontal offset dx :
//       - Euclidean ball of squared radius d2max: gmax[dx] = floor(sqrt(d2max - dx^2))
//         -> the inner rim  0 < EDT(mask) <= rim_px  of make_inside_rim_mask
//            (reference src/FRET/Nesprin2_FRET_Builder.py:409-414) is
//            mask & dilate(~mask, ball(rim_px^2)): a pixel is within Euclidean distance r of
//            a background pixel iff the ball around that background pixel covers it.  All
//            arithmetic is on integers (dx^2 + dy^2 <= d2max), so the result equals the
//            float64 EDT comparison bit for bit.
//       - full square (2p+1)^2: gmax[dx] = p  -> scipy.ndimage.binary_dilation(base, ones)
//         of annulus_mask_from_poly (Nesprin2_FRET_Builder.py:416-427); outside the frame
//         counts as 0 (scipy border_value = 0).
//     Two kernels: a column pass computes, per pixel, the vertical distance g to the nearest
//     source pixel of its column (capped at 255); a row pass sets a pixel iff some
//     horizontal offset dx has g(x+dx, y) <= gmax[|dx|].  O(H) + O(R) per pixel instead of
//     O(R^2), independent of how many pixels are set.
//   * exact integer first / second moments of a region mask (area, sum x, sum y, sum x^2,
//     sum y^2, sum xy in frame coordinates) for morphology_from_polygon
//     (reference src/MOR_by_ROI.py:193-241: area_px, centroid, np.cov of pixel coordinates).
//   * elementwise normalisation to uint16 previews (Fluor_INT.py:930-943,
//     fret_ratio_builder.py:479-483) and the ROI cropper's clip / mask / gamma chain
//     (roi_channel_cropper.py:923-953).
#pragma once
#include "ipb_rt.cuh"
#include "ipb_roistats.cuh"

#define IPB_MORPH_MAXR 254

struct IpbSpan { int R; unsigned char gmax[IPB_MORPH_MAXR + 2]; };

// ---------------------------------------------------------------- column pass
// grid (ceil(max_w / 128), n_regions), block 128: thread = one column of the region rect.
// g[g_off[r] + y * w + x] = min(255, vertical distance from (x, y) to the nearest source
// pixel of column x inside the rect); source = bit (invert: !bit).
__global__ void __launch_bounds__(128)
ipb_k_morph_coldist(const IpbRegion* __restrict__ regions, const unsigned* __restrict__ in_pool,
                    int invert, const long long* __restrict__ g_off, unsigned char* __restrict__ g)
{
    const IpbRegion rg = regions[blockIdx.y];
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= rg.w) return;
    const unsigned* m = in_pool + rg.mask_off;
    unsigned char* gr = g + g_off[blockIdx.y];
    const int j = x >> 5, b = x & 31;
    int d = 255;
    for (int y = 0; y < rg.h; ++y) {
        unsigned bit = (m[(size_t)y * rg.wpr + j] >> b) & 1u;
        if (invert) bit ^= 1u;
        d = bit ? 0 : (d < 255 ? d + 1 : 255);
        gr[(size_t)y * rg.w + x] = (unsigned char)d;
    }
    d = 255;
    for (int y = rg.h - 1; y >= 0; --y) {
        const int up = gr[(size_t)y * rg.w + x];
        d = up == 0 ? 0 : (d < 255 ? d + 1 : 255);
        if (d < up) gr[(size_t)y * rg.w + x] = (unsigned char)d;
    }
}

// ---------------------------------------------------------------- row pass
// grid (ceil(max_h / 8), n_regions), block 256: warp per output word, lane b <-> pixel 32j+b.
// out = dilated & and_pool & ~andnot_pool (either pool may be null), same layout as in_pool.
__global__ void __launch_bounds__(256)
ipb_k_morph_rowtest(const IpbRegion* __restrict__ regions, const long long* __restrict__ g_off,
                    const unsigned char* __restrict__ g, IpbSpan span,
                    const unsigned* __restrict__ and_pool, const unsigned* __restrict__ andnot_pool,
                    unsigned* __restrict__ out_pool)
{
    const IpbRegion rg = regions[blockIdx.y];
    const int y0 = (int)blockIdx.x * 8;
    if (y0 >= rg.h) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    int nrow = rg.h - y0;
    if (nrow > 8) nrow = 8;
    const unsigned char* gr = g + g_off[blockIdx.y];
    for (int i = warp; i < nrow * rg.wpr; i += nwarps) {
        const int y = y0 + i / rg.wpr, j = i % rg.wpr;
        const int x = 32 * j + lane;
        bool on = false;
        if (x < rg.w) {
            const unsigned char* row = gr + (size_t)y * rg.w;
            int lo = x - span.R, hi = x + span.R;
            if (lo < 0) lo = 0;
            if (hi > rg.w - 1) hi = rg.w - 1;
            for (int xx = lo; xx <= hi && !on; ++xx) {
                const int dx = xx > x ? xx - x : x - xx;
                on = row[xx] <= span.gmax[dx];
            }
        }
        unsigned word = __ballot_sync(IPB_FULL, on);
        if (lane == 0) {
            const size_t wi = (size_t)rg.mask_off + (size_t)y * rg.wpr + j;
            if (and_pool) word &= and_pool[wi];
            if (andnot_pool) word &= ~andnot_pool[wi];
            out_pool[wi] = word;
        }
    }
}

// ---------------------------------------------------------------- moments
// out[r][6] = { n, sum x, sum y, sum x^2, sum y^2, sum x*y } over the region's set pixels,
// x / y in FRAME coordinates, exact uint64.  One CTA (256 threads) per region.
__global__ void __launch_bounds__(256)
ipb_k_region_moments(const IpbRegion* __restrict__ regions, const unsigned* __restrict__ mask_pool,
                     unsigned long long* __restrict__ out)
{
    const IpbRegion rg = regions[blockIdx.x];
    const unsigned* m = mask_pool + rg.mask_off;
    unsigned long long s[6] = {0, 0, 0, 0, 0, 0};
    const unsigned nwords = (unsigned)rg.h * (unsigned)rg.wpr;
    for (unsigned wi = threadIdx.x; wi < nwords; wi += blockDim.x) {
        unsigned w = m[wi];
        if (!w) continue;
        const unsigned r = wi / (unsigned)rg.wpr, j = wi - r * (unsigned)rg.wpr;
        const unsigned long long y = (unsigned long long)(rg.y0 + (int)r);
        const unsigned long long xb = (unsigned long long)(rg.x0 + 32 * (int)j);
        unsigned long long c = 0, sx = 0, sxx = 0;
        while (w) {
            const int b = __ffs((int)w) - 1;
            w &= w - 1;
            const unsigned long long x = xb + (unsigned long long)b;
            ++c; sx += x; sxx += x * x;
        }
        s[0] += c; s[1] += sx; s[2] += c * y; s[3] += sxx; s[4] += c * y * y; s[5] += sx * y;
    }
    __shared__ unsigned long long red[6][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; ++k) { s[k] = ipb_warp_sum(s[k]); if (lane == 0) red[k][warp] = s[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
        out[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
    }
}

// ---------------------------------------------------------------- normalisation
// Preview (Fluor_INT.py:932-943): clip(img, lo, hi) -> (c - lo) / den -> * 65535 -> uint16
// (truncation; NaN -> 0).  lohi[img][3] = {lo, hi, den} float32 per image, den = float32 of
// the float64 expression (hi - lo + 1e-12) evaluated by the caller.
__global__ void __launch_bounds__(256)
ipb_k_preview_u16(const float* __restrict__ img, long long px_per_image, int n_images,
                  const float* __restrict__ lohi, unsigned short* __restrict__ out)
{
    const long long total = px_per_image * n_images;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / px_per_image);
        const float lo = lohi[3 * k], hi = lohi[3 * k + 1], den = lohi[3 * k + 2];
        float v = img[i];
        unsigned short o = 0;
        if (v == v) {
            v = v < lo ? lo : (v > hi ? hi : v);
            const float nrm = __fmul_rn(__fdiv_rn(__fsub_rn(v, lo), den), 65535.0f);
            o = (unsigned short)(int)nrm;
        }
        out[i] = o;
    }
}

// ROI cropper (roi_channel_cropper.py:923-953): norm = clip((crop - lo) / (hi - lo), 0, 1);
// norm *= mask (when mask_outside); norm_gamma = norm ** inv_gamma (float32 powf);
// out16 = (clip(norm_gamma, 0, 1) * 65535) -> uint16.  One crop per blockIdx.y.
// crop[k] = {plane, x0, y0, w, h, out_off (elements), region (mask, or < 0)}; params[k] = {lo, hi - lo}
struct IpbCropJob { int plane, x0, y0, w, h, region; long long out_off; };
__global__ void __launch_bounds__(256)
ipb_k_crop_normalize(const IpbCropJob* __restrict__ jobs, const unsigned short* __restrict__ planes,
                     int H, int W, const float* __restrict__ params, float inv_gamma,
                     const IpbRegion* __restrict__ regions, const unsigned* __restrict__ mask_pool,
                     float* __restrict__ out_norm, unsigned short* __restrict__ out16)
{
    const IpbCropJob cj = jobs[blockIdx.y];
    const float lo = params[2 * blockIdx.y], span = params[2 * blockIdx.y + 1];
    const unsigned short* img = planes + (size_t)cj.plane * H * W;
    const long long n = (long long)cj.w * cj.h;
    IpbRegion rg;
    const unsigned* m = nullptr;
    if (cj.region >= 0) { rg = regions[cj.region]; m = mask_pool + rg.mask_off; }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / cj.w), x = (int)(i % cj.w);
        const float v = (float)img[(size_t)(cj.y0 + y) * W + (cj.x0 + x)];
        float nrm = __fdiv_rn(__fsub_rn(v, lo), span);
        nrm = nrm < 0.0f ? 0.0f : (nrm > 1.0f ? 1.0f : nrm);
        if (m) {
            const int lx = cj.x0 + x - rg.x0, ly = cj.y0 + y - rg.y0;
            bool in = lx >= 0 && ly >= 0 && lx < rg.w && ly < rg.h &&
                      ((m[(size_t)ly * rg.wpr + (lx >> 5)] >> (lx & 31)) & 1u);
            nrm = in ? nrm : 0.0f;
        }
        const float gm = inv_gamma == 1.0f ? nrm : powf(nrm, inv_gamma);
        if (out_norm) out_norm[cj.out_off + i] = gm;
        if (out16) {
            const float c = gm < 0.0f ? 0.0f : (gm > 1.0f ? 1.0f : gm);
            out16[cj.out_off + i] = (unsigned short)(int)__fmul_rn(c, 65535.0f);
        }
    }
}

// eps = max(eps_abs, percentile of the corrected denominator under the ROI union) from a
// float32 region-statistics row (Nesprin2 pick_epsilon, Nesprin2_FRET_Builder.py:470-476,1484-1486)
__global__ void ipb_k_eps_from_stat(const IpbStatOut* __restrict__ so, const int* __restrict__ row_of_frame,
                                    int F, float eps_abs, float* __restrict__ fparams)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float eps = eps_abs;
    const int row = row_of_frame[f];
    if (row >= 0) {
        const IpbStatOut o = so[row];
        if (o.n > 0 && o.q[0] > eps) eps = o.q[0];
    }
    fparams[(size_t)f * 4 + 2] = eps;
}
