// ipb_exact.cuh -- bit-exact scalar arithmetic shared by kernels: ordered float keys and the
// float32 arithmetic numpy performs for percentile / median of a float32 array
// (numpy/lib/_function_base_impl.py: percentile -> _quantile -> _get_indexes / _get_gamma /
// _lerp, method 'linear'; numpy 2.x evaluates all of it in the array's dtype).
// Every operation is an explicit round-to-nearest intrinsic so nvcc cannot contract a*b+c
// into an FMA (numpy never does).
#pragma once
#include "ipb_rt.cuh"

// monotone float32 <-> uint32 key (total order, -0 < +0)
IPB_HD uint32_t ipb_f32_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
IPB_HD float ipb_key_f32(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

struct IpbQIdx {
    long long prev, next;   // indices into the sorted sample
    float gamma;            // interpolation weight
};

// numpy _quantile(method='linear') index arithmetic for a float32 sample of size n (n >= 1)
// and quantile q32 = float32(p) / float32(100) (computed by the caller exactly that way).
IPB_HD IpbQIdx ipb_np_qidx_f32(long long n, float q32) {
    IpbQIdx r;
    float nm1 = __ll2float_rn(n - 1);
    float vi = __fmul_rn(nm1, q32);            // (n - 1) * quantiles
    float prevf = floorf(vi);                   // _get_indexes
    float nextf = __fadd_rn(prevf, 1.0f);
    if (vi >= nm1) {                            // indexes_above_bounds -> -1 (last element)
        r.prev = n - 1; r.next = n - 1;
    } else if (vi < 0.0f) {
        r.prev = 0; r.next = 0;
    } else {
        r.prev = (long long)prevf; r.next = (long long)nextf;
    }
    r.gamma = __fsub_rn(vi, prevf);             // _get_gamma (uses floor(vi), not the clipped index)
    return r;
}

// numpy _lerp in float32
IPB_HD float ipb_np_lerp_f32(float a, float b, float t) {
    float diff = __fsub_rn(b, a);
    float r = __fadd_rn(a, __fmul_rn(diff, t));
    if (t >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, t)));
    return r;
}

// numpy median of a float32 sample: mean of the one or two middle order statistics
// (np.mean of 2 float32 = float32 sum, then / 2).
IPB_HD float ipb_np_mid2_f32(float a, float b) { return __fdiv_rn(__fadd_rn(a, b), 2.0f); }
