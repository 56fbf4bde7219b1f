// ipb_fret.cuh -- fused elementwise FRET pass and the tiny
// per-frame parameter kernels that turn histogram order statistics into the scalars the
// reference computes on the host (background B, epsilon, FA threshold).
//
// One read of the donor and acceptor uint16 planes (128-bit loads), one write of the
// float32 ratio image (128-bit stores); every arithmetic step is a separately rounded
// float32 operation in the reference's order (numpy evaluates (yf - alpha*d) - beta*ao as
// three ufunc calls; no FMA contraction here either):
//   saturation -> NaN            Nesprin2_FRET_Builder.py:1415-1421
//   J = img - B ; J[J<0] = 0     fret_ratio_builder.py:332-336 (NaN stays NaN)
//   spectral correction          Nesprin2_FRET_Builder.py:460-468
//   R = (numer+eps)/(denom+eps)  fret_ratio_builder.py:474 ; Nesprin2 1499-1500 (+ inverse)
//   R > clip_max -> NaN          Nesprin2_FRET_Builder.py:1502-1504
//   R_roi = R, NaN outside union fret_ratio_builder.py:494-495
#pragma once
#include "ipb_rt.cuh"
#include "ipb_exact.cuh"
#include "ipb_hist.cuh"

struct IpbFretCfg {
    int numer_is_acceptor;   // 1: ratio_mode "FRET/Donor" (numer = acceptor), 0: "Donor/FRET"
    int clip_neg;
    int sat_on;   float sat_thr;
    int use_spectral; float alpha, beta, g_factor;
    int clip_on;  float clip_max;
    int donor_ch, acc_ch, aonly_ch;   // channel indices inside a frame; aonly_ch < 0: none
    int n_ch;
};

// per-frame scalars, float32 as the reference uses them: [frame][4] = {Bd, Ba, eps, Bao}
#define IPB_FP_BD 0
#define IPB_FP_BA 1
#define IPB_FP_EPS 2
#define IPB_FP_BAO 3
#define IPB_FP_STRIDE 4

__device__ __forceinline__ float ipb_bgsub(float v, float B, int clip_neg) {
    float j = __fsub_rn(v, B);
    if (clip_neg && j < 0.0f) j = 0.0f;
    return j;
}

struct IpbFretPx { float R, Ralt, dcorr, acorr; };

// PLAIN = no saturation filter, no spectral correction, no ratio clip (the general FRET builder,
// fret_ratio_builder.py:466-474): those per-pixel tests are compiled out
template <bool WANT_ALT, bool PLAIN>
__device__ __forceinline__ IpbFretPx ipb_fret_px(const IpbFretCfg& cfg, const float* fp,
                                                 unsigned dv, unsigned av, unsigned aov) {
    const float fnan = __uint_as_float(0x7fc00000u);
    float d = (float)dv, a = (float)av;
    if (!PLAIN && cfg.sat_on && (d >= cfg.sat_thr || a >= cfg.sat_thr)) { d = fnan; a = fnan; }
    const float dbc = ipb_bgsub(d, fp[IPB_FP_BD], cfg.clip_neg);
    const float abc = ipb_bgsub(a, fp[IPB_FP_BA], cfg.clip_neg);
    float acorr = abc;
    if (!PLAIN && cfg.use_spectral) {
        float t = __fsub_rn(abc, __fmul_rn(cfg.alpha, dbc));
        if (cfg.aonly_ch >= 0) {
            const float aobc = ipb_bgsub((float)aov, fp[IPB_FP_BAO], cfg.clip_neg);
            t = __fsub_rn(t, __fmul_rn(cfg.beta, aobc));
        }
        acorr = __fmul_rn(t, cfg.g_factor);
    }
    const float eps = fp[IPB_FP_EPS];
    const float numer = cfg.numer_is_acceptor ? acorr : dbc;
    const float denom = cfg.numer_is_acceptor ? dbc : acorr;
    IpbFretPx o;
    o.R = __fdiv_rn(__fadd_rn(numer, eps), __fadd_rn(denom, eps));
    o.Ralt = WANT_ALT ? __fdiv_rn(__fadd_rn(denom, eps), __fadd_rn(numer, eps)) : 0.0f;
    if (!PLAIN && cfg.clip_on) {
        if (o.R > cfg.clip_max) o.R = fnan;
        if (WANT_ALT && o.Ralt > cfg.clip_max) o.Ralt = fnan;
    }
    o.dcorr = dbc; o.acorr = acorr;
    return o;
}

// planes: uint16 [F][n_ch][H][W].  Outputs (any may be null): R, Ralt, Rroi, Dcorr, Acorr,
// each float32 [F][H][W].  Requires W % 8 == 0 for the vector path (scalar path otherwise).
template <bool WANT_ALT, bool PLAIN>
__global__ void __launch_bounds__(256)
ipb_k_fret_pixels(const unsigned short* __restrict__ planes, int F, int H, int W, IpbFretCfg cfg,
                  const float* __restrict__ fparams, const unsigned* __restrict__ union_bits, int union_wpr,
                  const int* __restrict__ union_idx /* frame -> union plane; null: identity */,
                  float* __restrict__ R, float* __restrict__ Ralt, float* __restrict__ Rroi,
                  float* __restrict__ Dcorr, float* __restrict__ Acorr,
                  unsigned long long* __restrict__ mom_stats /* nullable: [.][4] = {n, sum, sumsq, 0} rows */,
                  const int* __restrict__ mom_idx /* [F] row of frame f */, int mom_acceptor)
{
    const long long plane_px = (long long)H * W;
    const float fnan = __uint_as_float(0x7fc00000u);
    // integer moments of one channel ride along (FA global statistics, FA_Analyzer.py:984-987): the pass
    // is bound by HBM, the two integer multiply-adds per pixel find free issue slots
    unsigned m1 = 0;
    unsigned long long m2 = 0;
    if ((W & 7) == 0) {
        // grid = (chunks, frames): indices inside a frame are walked without 64-bit divisions;
        // two 8-pixel groups per trip, all of their 128-bit loads issued before the arithmetic
        const long long vpf = plane_px >> 3;
        const int f = blockIdx.y;
        const unsigned short* base = planes + (size_t)f * cfg.n_ch * plane_px;
        const unsigned short* dptr = base + (size_t)cfg.donor_ch * plane_px;
        const unsigned short* aptr = base + (size_t)cfg.acc_ch * plane_px;
        const bool have_ao = !PLAIN && cfg.use_spectral && cfg.aonly_ch >= 0;
        const unsigned short* optr = have_ao ? base + (size_t)cfg.aonly_ch * plane_px : nullptr;
        const float* fp = fparams + (size_t)f * IPB_FP_STRIDE;
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long iv0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv0 < vpf; iv0 += 2 * stride) {
            uint4 dq[2], aq[2], oq[2];
            bool ok[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const long long iv = iv0 + g * stride;
                ok[g] = iv < vpf;
                dq[g] = aq[g] = oq[g] = make_uint4(0, 0, 0, 0);
                if (ok[g]) {
                    dq[g] = __ldg(reinterpret_cast<const uint4*>(dptr + (iv << 3)));
                    aq[g] = __ldg(reinterpret_cast<const uint4*>(aptr + (iv << 3)));
                    if (have_ao) oq[g] = __ldg(reinterpret_cast<const uint4*>(optr + (iv << 3)));
                }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (!ok[g]) continue;
                const long long p0 = (iv0 + g * stride) << 3;
                const unsigned dw[4] = {dq[g].x, dq[g].y, dq[g].z, dq[g].w}, aw[4] = {aq[g].x, aq[g].y, aq[g].z, aq[g].w},
                               ow[4] = {oq[g].x, oq[g].y, oq[g].z, oq[g].w};
                if (mom_stats) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned w = mom_acceptor ? aw[j] : dw[j];
                        const unsigned a = w & 0xffffu, b = w >> 16;
                        m1 += a + b;
                        m2 += (unsigned long long)a * a;
                        m2 += (unsigned long long)b * b;
                    }
                }
                unsigned ub = 0xffu;
                if (Rroi) {
                    const int y = (int)((unsigned long long)p0 / (unsigned)W), x0 = (int)(p0 - (long long)y * W);
                    const int uf = union_idx ? union_idx[f] : f;
                    ub = union_bits ? ((union_bits[((size_t)uf * H + y) * union_wpr + (x0 >> 5)] >> (x0 & 31)) & 0xffu) : 0u;
                }
                float r[8], ra[8], rr[8], dc[8], ac[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const unsigned dv = (t & 1) ? (dw[t >> 1] >> 16) : (dw[t >> 1] & 0xffffu);
                    const unsigned av = (t & 1) ? (aw[t >> 1] >> 16) : (aw[t >> 1] & 0xffffu);
                    const unsigned ov = (t & 1) ? (ow[t >> 1] >> 16) : (ow[t >> 1] & 0xffffu);
                    const IpbFretPx o = ipb_fret_px<WANT_ALT, PLAIN>(cfg, fp, dv, av, ov);
                    r[t] = o.R; ra[t] = o.Ralt; dc[t] = o.dcorr; ac[t] = o.acorr;
                    rr[t] = ((ub >> t) & 1u) ? o.R : fnan;
                }
                const size_t o0 = (size_t)f * plane_px + p0;
                if (R)     { reinterpret_cast<float4*>(R + o0)[0] = make_float4(r[0], r[1], r[2], r[3]);         reinterpret_cast<float4*>(R + o0)[1] = make_float4(r[4], r[5], r[6], r[7]); }
                if (WANT_ALT && Ralt) { reinterpret_cast<float4*>(Ralt + o0)[0] = make_float4(ra[0], ra[1], ra[2], ra[3]);  reinterpret_cast<float4*>(Ralt + o0)[1] = make_float4(ra[4], ra[5], ra[6], ra[7]); }
                if (Rroi)  { reinterpret_cast<float4*>(Rroi + o0)[0] = make_float4(rr[0], rr[1], rr[2], rr[3]);  reinterpret_cast<float4*>(Rroi + o0)[1] = make_float4(rr[4], rr[5], rr[6], rr[7]); }
                if (Dcorr) { reinterpret_cast<float4*>(Dcorr + o0)[0] = make_float4(dc[0], dc[1], dc[2], dc[3]); reinterpret_cast<float4*>(Dcorr + o0)[1] = make_float4(dc[4], dc[5], dc[6], dc[7]); }
                if (Acorr) { reinterpret_cast<float4*>(Acorr + o0)[0] = make_float4(ac[0], ac[1], ac[2], ac[3]); reinterpret_cast<float4*>(Acorr + o0)[1] = make_float4(ac[4], ac[5], ac[6], ac[7]); }
            }
        }
    } else {
        const int f = blockIdx.y;
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < plane_px;
             p += (long long)gridDim.x * blockDim.x) {
            const long long i = (long long)f * plane_px + p;
            const unsigned short* base = planes + (size_t)f * cfg.n_ch * plane_px;
            const unsigned dv = base[(size_t)cfg.donor_ch * plane_px + p];
            const unsigned av = base[(size_t)cfg.acc_ch * plane_px + p];
            const unsigned ov = (cfg.use_spectral && cfg.aonly_ch >= 0) ? base[(size_t)cfg.aonly_ch * plane_px + p] : 0u;
            const IpbFretPx o = ipb_fret_px<WANT_ALT, PLAIN>(cfg, fparams + (size_t)f * IPB_FP_STRIDE, dv, av, ov);
            if (mom_stats) { const unsigned v = mom_acceptor ? av : dv; m1 += v; m2 += (unsigned long long)v * v; }
            if (R) R[i] = o.R;
            if (WANT_ALT && Ralt) Ralt[i] = o.Ralt;
            if (Dcorr) Dcorr[i] = o.dcorr;
            if (Acorr) Acorr[i] = o.acorr;
            if (Rroi) {
                const int y = (int)(p / W), x = (int)(p % W);
                const int uf = union_idx ? union_idx[f] : f;
                const bool in = union_bits && ((union_bits[((size_t)uf * H + y) * union_wpr + (x >> 5)] >> (x & 31)) & 1u);
                Rroi[i] = in ? o.R : fnan;
            }
        }
    }
    if (mom_stats) {                                               // block-uniform
        // (a thread sums at most 2^16 pixels of a frame: m1 < 2^32)
        __shared__ unsigned long long msum[2];
        if (threadIdx.x < 2) msum[threadIdx.x] = 0ull;
        __syncthreads();
        const unsigned long long w1 = ipb_warp_sum((unsigned long long)m1), w2 = ipb_warp_sum(m2);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&msum[0], w1); atomicAdd(&msum[1], w2); }
        __syncthreads();
        if (threadIdx.x < 2 && msum[threadIdx.x])
            atomicAdd(&mom_stats[(size_t)mom_idx[blockIdx.y] * 4 + 1 + threadIdx.x], msum[threadIdx.x]);
    }
}

// ---------------------------------------------------------------- per-frame scalars
// dst[dst_idx[i]] = percentile value of quantile job i (0 when the sample was empty:
// bg_value returns 0.0 for vals.size == 0, fret_ratio_builder.py:316-317)
__global__ void ipb_k_scatter_qvalues(const IpbQOut* __restrict__ qout, const int* __restrict__ dst_idx,
                                      int n, float* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && dst_idx[i] >= 0) dst[dst_idx[i]] = qout[i].n > 0 ? qout[i].value : 0.0f;
}

// eps = max(5.0, percentile(denominator values, p_floor))   (fret_ratio_builder.py:338-340)
// where the denominator is the bg-corrected donor or acceptor: a monotone transform of the
// raw sample, so its order statistics are the transformed raw order statistics.
// qout_eps[f]: quantile job on the denominator channel's histogram at the eps quantile.
__global__ void ipb_k_fret_eps(const IpbQOut* __restrict__ qout_eps, int F, int denom_slot /* IPB_FP_BD or _BA */,
                               int clip_neg, float eps_abs, float* __restrict__ fparams)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const IpbQOut q = qout_eps[f];
    float eps = eps_abs;
    if (q.n > 0 && q.prev >= 0 && q.next >= 0) {
        const float B = fparams[(size_t)f * IPB_FP_STRIDE + denom_slot];
        const float a = ipb_bgsub((float)q.prev, B, clip_neg), b = ipb_bgsub((float)q.next, B, clip_neg);
        const float p = ipb_np_lerp_f32(a, b, q.gamma);
        if (p > eps) eps = p;                   // python max(5.0, p): p wins only if p > 5.0
    }
    fparams[(size_t)f * IPB_FP_STRIDE + IPB_FP_EPS] = eps;
}

// FA global stats (FA_Analyzer.py:984-987) from exact integer moments + the [::10, ::10]
// sample percentile; threshold m + alpha*s in float32 (FA_Analyzer.py:143-144).
// fa[f] = {mean, std, bg, thr}
__global__ void ipb_k_fa_params(const unsigned long long* __restrict__ stats /* hist job stats */,
                                const int* __restrict__ stat_idx, const IpbQOut* __restrict__ qout_bg,
                                int F, long long npx, float alpha, float* __restrict__ fa)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const unsigned long long* st = stats + (size_t)stat_idx[f] * 4;
    const double n = (double)npx;
    const double mean = (double)st[1] / n;
    double var = (double)st[2] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float m = (float)mean, s = (float)sqrt(var);
    fa[f * 4 + 0] = m;
    fa[f * 4 + 1] = s;
    fa[f * 4 + 2] = qout_bg[f].n > 0 ? qout_bg[f].value : 0.0f;
    fa[f * 4 + 3] = __fadd_rn(m, __fmul_rn(alpha, s));
}
