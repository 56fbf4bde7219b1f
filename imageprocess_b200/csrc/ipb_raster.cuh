// ipb_raster.cuh -- batched, bit-exact ROI polygon rasterisation.
//
// Replaces, for every ROI of every frame of a batch in ONE launch:
//   rule IPB_RULE_MPL  matplotlib.path.Path(poly).contains_points(grid) as called by
//                      rasterize_polygon (reference src/INT/Fluor_INT.py:398-403 and copies)
//   rule IPB_RULE_SK   skimage.draw.polygon(r, c, shape) (reference src/INT/FA_Analyzer.py:1014)
//
// Both library rules test every grid point against every edge in float64.  Here one warp
// owns one (ROI, row): each lane takes edges, and for an edge that straddles the row it
// evaluates the library's own float64 expression (explicit _rn intrinsics, no FMA) only
//   * once at a representative pixel left of the edge's x-range and once right of it --
//     outside [floor(xmin)-1, ceil(xmax)+1] the sign of the expression provably cannot
//     change (no cancellation; margin >= 1 px), so one evaluation stands for the whole
//     half-line and becomes a range toggle in a per-row "delta" bit array, and
//   * per pixel inside that (short) middle zone, as a direct bit toggle.
// A warp-wide prefix-XOR over the delta words then yields the crossing parity of every
// pixel of the row.  Results are therefore identical to evaluating the library expression
// at every pixel, at O(edges + row words) per row instead of O(edges * row pixels).
//
// Output: one bit-packed mask per ROI over its storage rect (row-major, 32 px per word,
// bit b of word j <-> local x = sx0 + 32 j + b), the ROI pixel count, and optionally the
// per-frame union bitmask (bit b of word j <-> frame x = 32 j + b).
#pragma once
#include "ipb_rt.cuh"

#define IPB_RULE_MPL 0
#define IPB_RULE_SK 1
#define IPB_RASTER_WARPS 8
#define IPB_RASTER_LONG 48   // middle zones longer than this are evaluated by the whole warp

__device__ __forceinline__ int ipb_clamp_d2i(double v) {
    if (!(v > -1073741824.0)) return -1073741824;
    if (!(v < 1073741824.0)) return 1073741824;
    return (int)v;
}

// toggle [a, b) (local x, a < b) in the delta array of a row stored from sx0 to sx1
__device__ __forceinline__ void ipb_toggle_range(unsigned* d, int a, int b, int sx0, int sx1) {
    int pa = a - sx0;
    atomicXor(&d[pa >> 5], 1u << (pa & 31));
    if (b < sx1) {
        int pb = b - sx0;
        atomicXor(&d[pb >> 5], 1u << (pb & 31));
    }
}
__device__ __forceinline__ void ipb_toggle_bit(unsigned* d, int x, int sx0) {
    int p = x - sx0;
    atomicXor(&d[p >> 5], 1u << (p & 31));
}

// ---- matplotlib rule: edge v0 -> v1, row ty.  Returns predicate at integer tx.
struct IpbMplEdge {
    double lhs, dy, v1x;
    bool f1;
    __device__ __forceinline__ bool hit(int tx) const {
        double rhs = __dmul_rn(__dsub_rn(v1x, (double)tx), dy);
        return (lhs >= rhs) == f1;
    }
};

// ---- skimage rule: edge (v_i, v_{i-1}) seen from row y.  sign of the quotient at x.
struct IpbSkEdge {
    double y0, y1, den, vix, vpx;
    __device__ __forceinline__ double quot(int x) const {
        double x0 = __dsub_rn(vix, (double)x);
        double x1 = __dsub_rn(vpx, (double)x);
        return __ddiv_rn(__dsub_rn(__dmul_rn(x0, y1), __dmul_rn(x1, y0)), den);
    }
};

template <int RULE>
__global__ void __launch_bounds__(IPB_RASTER_WARPS * 32)
ipb_k_raster(int n_rois, const double2* __restrict__ verts, const int* __restrict__ vert_off,
             const int4* __restrict__ erect, const int4* __restrict__ srect,
             const int2* __restrict__ org, const int* __restrict__ roi_frame,
             const long long* __restrict__ mask_off, int max_wpr,
             unsigned* __restrict__ mask_pool, unsigned* __restrict__ area,
             unsigned* __restrict__ union_bits, int union_wpr, int frame_h)
{
    IPB_DYN_SMEM(unsigned, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.y;
    if (r >= n_rois) return;
    const int4 S = srect[r];
    const int yl = S.y + (int)blockIdx.x * IPB_RASTER_WARPS + warp;
    if (yl >= S.w) return;                                   // whole warp leaves together
    const int4 E = erect[r];
    const int sw = S.z - S.x;
    const int wpr = (sw + 31) >> 5;
    unsigned* dR = smem + (size_t)warp * 5 * max_wpr;
    unsigned* dL = dR + max_wpr;
    unsigned* xR = dL + max_wpr;
    unsigned* xL = xR + max_wpr;
    unsigned* vt = xL + max_wpr;
    for (int j = lane; j < wpr; j += 32) { dR[j] = 0; dL[j] = 0; xR[j] = 0; xL[j] = 0; vt[j] = 0; }
    __syncwarp();

    const int voff = vert_off[r], nv = vert_off[r + 1] - voff;
    const bool row_active = (yl >= E.y && yl < E.w && E.x < E.z && nv >= (RULE == IPB_RULE_MPL ? 3 : 1));
    if (row_active) {
        const double ty = (double)yl;
        for (int e0 = 0; e0 < nv; e0 += 32) {
            const int e = e0 + lane;
            bool is_long = false;
            int mlo = 0, mhi = -1;
            IpbMplEdge me; IpbSkEdge se; bool strR = false, strL = false;
            me.lhs = 0; me.dy = 0; me.v1x = 0; me.f1 = false;
            se.y0 = 0; se.y1 = 0; se.den = 1; se.vix = 0; se.vpx = 0;
            if (e < nv) {
                double2 a, b;     // MPL: a = v0 = P[e], b = v1 = P[e+1];  SK: a = v_i, b = v_{i-1}
                if (RULE == IPB_RULE_MPL) {
                    a = verts[voff + e];
                    b = verts[voff + (e + 1 == nv ? 0 : e + 1)];
                    bool f0 = a.y >= ty;
                    me.f1 = b.y >= ty;
                    strR = (f0 != me.f1);
                    me.lhs = __dmul_rn(__dsub_rn(b.y, ty), __dsub_rn(a.x, b.x));
                    me.dy = __dsub_rn(a.y, b.y);
                    me.v1x = b.x;
                } else {
                    a = verts[voff + e];
                    b = verts[voff + (e == 0 ? nv - 1 : e - 1)];
                    se.y0 = __dsub_rn(a.y, ty);
                    se.y1 = __dsub_rn(b.y, ty);
                    se.den = __dsub_rn(se.y1, se.y0);
                    se.vix = a.x; se.vpx = b.x;
                    strR = ((se.y0 > 0) != (se.y1 > 0));
                    strL = ((se.y0 < 0) != (se.y1 < 0));
                    // vertex with eps tolerance: "cdef float eps = 1e-12"
                    const double eps = (double)1e-12f;
                    if (-eps < se.y0 && se.y0 < eps) {
                        int xc = ipb_clamp_d2i(rint(a.x));
                        for (int x = xc - 1; x <= xc + 1; ++x) {
                            double x0 = __dsub_rn(a.x, (double)x);
                            if (-eps < x0 && x0 < eps && x >= E.x && x < E.z) atomicOr(&vt[(x - S.x) >> 5], 1u << ((x - S.x) & 31));
                        }
                    }
                }
                if (strR || strL) {
                    double xa = a.x < b.x ? a.x : b.x, xb = a.x < b.x ? b.x : a.x;
                    mlo = ipb_clamp_d2i(floor(xa)) - 1;
                    mhi = ipb_clamp_d2i(ceil(xb)) + 1;
                    // left half-line: tx <= mlo - 1
                    if (E.x < mlo) {
                        int b_end = mlo < E.z ? mlo : E.z;
                        if (RULE == IPB_RULE_MPL) {
                            if (me.hit(mlo - 1)) ipb_toggle_range(dR, E.x, b_end, S.x, S.z);
                        } else {
                            double q = se.quot(mlo - 1);
                            if (strR && q > 0) ipb_toggle_range(dR, E.x, b_end, S.x, S.z);
                            if (strL && q < 0) ipb_toggle_range(dL, E.x, b_end, S.x, S.z);
                        }
                    }
                    // right half-line: tx >= mhi + 1
                    if (mhi + 1 < E.z) {
                        int a_beg = (mhi + 1) > E.x ? (mhi + 1) : E.x;
                        if (RULE == IPB_RULE_MPL) {
                            if (me.hit(mhi + 1)) ipb_toggle_range(dR, a_beg, E.z, S.x, S.z);
                        } else {
                            double q = se.quot(mhi + 1);
                            if (strR && q > 0) ipb_toggle_range(dR, a_beg, E.z, S.x, S.z);
                            if (strL && q < 0) ipb_toggle_range(dL, a_beg, E.z, S.x, S.z);
                        }
                    }
                    if (mlo < E.x) mlo = E.x;
                    if (mhi > E.z - 1) mhi = E.z - 1;
                    if (mhi - mlo + 1 > IPB_RASTER_LONG) {
                        is_long = true;
                    } else {
                        for (int x = mlo; x <= mhi; ++x) {
                            if (RULE == IPB_RULE_MPL) {
                                if (me.hit(x)) ipb_toggle_bit(xR, x, S.x);
                            } else {
                                double q = se.quot(x);
                                if (strR && q > 0) ipb_toggle_bit(xR, x, S.x);
                                if (strL && q < 0) ipb_toggle_bit(xL, x, S.x);
                            }
                        }
                    }
                }
            }
            // long middle zones: the whole warp evaluates one edge's pixels together
            unsigned lm = __ballot_sync(IPB_FULL, is_long);
            while (lm) {
                const int src = __ffs((int)lm) - 1;
                lm &= lm - 1;
                const int b_lo = __shfl_sync(IPB_FULL, mlo, src);
                const int b_hi = __shfl_sync(IPB_FULL, mhi, src);
                if (RULE == IPB_RULE_MPL) {
                    IpbMplEdge g;
                    g.lhs = __shfl_sync(IPB_FULL, me.lhs, src);
                    g.dy = __shfl_sync(IPB_FULL, me.dy, src);
                    g.v1x = __shfl_sync(IPB_FULL, me.v1x, src);
                    g.f1 = __shfl_sync(IPB_FULL, (int)me.f1, src) != 0;
                    for (int x = b_lo + lane; x <= b_hi; x += 32)
                        if (g.hit(x)) ipb_toggle_bit(xR, x, S.x);
                } else {
                    IpbSkEdge g;
                    g.y0 = __shfl_sync(IPB_FULL, se.y0, src);
                    g.y1 = __shfl_sync(IPB_FULL, se.y1, src);
                    g.den = __shfl_sync(IPB_FULL, se.den, src);
                    g.vix = __shfl_sync(IPB_FULL, se.vix, src);
                    g.vpx = __shfl_sync(IPB_FULL, se.vpx, src);
                    const bool gR = __shfl_sync(IPB_FULL, (int)strR, src) != 0;
                    const bool gL = __shfl_sync(IPB_FULL, (int)strL, src) != 0;
                    for (int x = b_lo + lane; x <= b_hi; x += 32) {
                        double q = g.quot(x);
                        if (gR && q > 0) ipb_toggle_bit(xR, x, S.x);
                        if (gL && q < 0) ipb_toggle_bit(xL, x, S.x);
                    }
                }
            }
        }
    }
    __syncwarp();

    // prefix-XOR of the delta words -> crossing parity per pixel; combine; store
    unsigned carryR = 0, carryL = 0, cnt = 0;
    unsigned* out = mask_pool + mask_off[r] + (long long)(yl - S.y) * wpr;
    const int ox = org[r].x + S.x, oy = org[r].y + yl;
    unsigned* urow = nullptr;
    if (union_bits != nullptr && oy >= 0 && oy < frame_h)
        urow = union_bits + ((long long)roi_frame[r] * frame_h + oy) * union_wpr;
    for (int w0 = 0; w0 < wpr; w0 += 32) {
        const int j = w0 + lane;
        unsigned pr = (j < wpr) ? dR[j] : 0u, pl = 0u;
        pr ^= pr << 1; pr ^= pr << 2; pr ^= pr << 4; pr ^= pr << 8; pr ^= pr << 16;
        unsigned totR = pr >> 31, inclR = totR;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(IPB_FULL, inclR, o); if (lane >= o) inclR ^= t; }
        unsigned m = pr ^ (((inclR ^ totR ^ carryR) & 1u) ? 0xffffffffu : 0u);
        carryR ^= __shfl_sync(IPB_FULL, inclR, 31);
        if (j < wpr) m ^= xR[j];
        if (RULE == IPB_RULE_SK) {
            pl = (j < wpr) ? dL[j] : 0u;
            pl ^= pl << 1; pl ^= pl << 2; pl ^= pl << 4; pl ^= pl << 8; pl ^= pl << 16;
            unsigned totL = pl >> 31, inclL = totL;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(IPB_FULL, inclL, o); if (lane >= o) inclL ^= t; }
            unsigned ml = pl ^ (((inclL ^ totL ^ carryL) & 1u) ? 0xffffffffu : 0u);
            carryL ^= __shfl_sync(IPB_FULL, inclL, 31);
            if (j < wpr) { ml ^= xL[j]; m |= ml | vt[j]; }
        }
        if (j < wpr) {
            if (j == wpr - 1 && (sw & 31)) m &= (1u << (sw & 31)) - 1u;
            if (!row_active) m = 0u;
            out[j] = m;
            cnt += (unsigned)__popc(m);
            if (urow != nullptr && m) {
                const int X = ox + 32 * j;           // frame x of bit 0 (>= 0 by contract)
                const int k = X >> 5, s = X & 31;
                if (k >= 0 && k < union_wpr) atomicOr(&urow[k], m << s);
                if (s && (m >> (32 - s)) && k + 1 >= 0 && k + 1 < union_wpr) atomicOr(&urow[k + 1], m >> (32 - s));
            }
        }
    }
    cnt = ipb_warp_sum(cnt);
    if (lane == 0 && cnt) atomicAdd(&area[r], cnt);
}
