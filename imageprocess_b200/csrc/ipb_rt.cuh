// ipb_rt.cuh -- launch/runtime abstraction shared by every kernel file.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a (this is the only build the
// package loads).  With -DIPB_EMULATE the same kernel sources compile with g++ against
// tests/emu/cuda_emu.h so that the CPU test tier can run the kernels' logic in the build
// container (no GPU there); that library lives under tests/ and is never loaded by the
// product.
#pragma once
#include <stdint.h>

#ifdef IPB_EMULATE
#include "cuda_emu.h"
#define IPB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    IPB_EMU_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
#define IPB_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::g_blk->dyn_smem)
#define IPB_HD static inline
#else
#include <cuda_runtime.h>
#define IPB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define IPB_DYN_SMEM(T, name)                                        \
    extern __shared__ __align__(16) unsigned char name##_raw_[];     \
    T* name = reinterpret_cast<T*>(name##_raw_)
#define IPB_HD __device__ __forceinline__
#endif

#define IPB_FULL 0xffffffffu

// hint: bring the line of `p` into L2 (no register result, no fault on a bad address)
#ifdef IPB_EMULATE
static inline void ipb_prefetch_l2(const void*) {}
#else
__device__ __forceinline__ void ipb_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
#endif

// ---- status codes (mirrored in include/ipb200.h)
#define IPB_OK 0
#define IPB_ERR_ARG (-1)
#define IPB_ERR_CUDA (-2)
#define IPB_ERR_WORKSPACE (-3)
#define IPB_ERR_UNSUPPORTED (-4)

void ipb_set_error(const char* fmt, ...);
int ipb_check_launch(const char* what);

#define IPB_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ipb_set_error(__VA_ARGS__);        \
            return IPB_ERR_ARG;                \
        }                                      \
    } while (0)

static inline unsigned ipb_div_up(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---- warp helpers
__device__ __forceinline__ unsigned ipb_lane() { return threadIdx.x & 31u; }

template <typename T>
__device__ __forceinline__ T ipb_warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(IPB_FULL, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T ipb_warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(IPB_FULL, v, o); v = u < v ? u : v; }
    return v;
}
template <typename T>
__device__ __forceinline__ T ipb_warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(IPB_FULL, v, o); v = u > v ? u : v; }
    return v;
}

// ---- TMA bulk copies (cp.async.bulk, 1-D) tracked by an mbarrier: global -> shared rows without a
// register round trip or an issue slot per 16 bytes.  Product build: inline PTX for sm_100a (SASS:
// UBLKCP + SYNCS).  Emulated build: the copy happens at once and completes the phase when the announced
// bytes are in; the wait yields until the phase of the given parity is complete (the hardware's rule).
struct IpbMbar { unsigned long long v; };          // emulated: low word = completed phases, high word = bytes pending

__device__ __forceinline__ void ipb_mbar_init(IpbMbar* bar, unsigned count) {
#ifdef IPB_EMULATE
    bar->v = 0; (void)count;
#else
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory");
#endif
}
// makes freshly initialised barriers visible to the async proxy (call once after the inits)
__device__ __forceinline__ void ipb_mbar_fence_init() {
#ifndef IPB_EMULATE
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// one thread: announces `bytes` of incoming data and starts the copy src -> dst (both 16-byte
// aligned, bytes a multiple of 16); the barrier completes its phase when the bytes have landed
__device__ __forceinline__ void ipb_bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, IpbMbar* bar) {
#ifdef IPB_EMULATE
    memcpy(dst_smem, src_gmem, bytes);
    bar->v = (bar->v & 0xffffffffull) + 1ull;
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(d), "l"(src_gmem), "r"(bytes), "r"(b) : "memory");
#endif
}
// several copies on one barrier phase: ONE thread announces the total first, then starts the copies
__device__ __forceinline__ void ipb_mbar_expect(IpbMbar* bar, unsigned total_bytes) {
#ifdef IPB_EMULATE
    bar->v = (bar->v & 0xffffffffull) | ((unsigned long long)total_bytes << 32);
    if (total_bytes == 0) bar->v = (bar->v & 0xffffffffull) + 1ull;
#else
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(total_bytes) : "memory");
#endif
}
__device__ __forceinline__ void ipb_bulk_copy(void* dst_smem, const void* src_gmem, unsigned bytes, IpbMbar* bar) {
#ifdef IPB_EMULATE
    memcpy(dst_smem, src_gmem, bytes);
    const unsigned long long left = (bar->v >> 32) - bytes;
    bar->v = left ? ((bar->v & 0xffffffffull) | (left << 32)) : ((bar->v & 0xffffffffull) + 1ull);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(d), "l"(src_gmem), "r"(bytes), "r"(b) : "memory");
#endif
}
// every consumer thread: blocks until the barrier's phase with parity `parity` has completed
__device__ __forceinline__ void ipb_mbar_wait(IpbMbar* bar, unsigned parity) {
#ifdef IPB_EMULATE
    while (((unsigned)(bar->v & 1ull)) == parity) emu::yield();
#else
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{ .reg .pred p;\n\t"
                 "IPB_WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra IPB_DONE_%=;\n\t"
                 "bra IPB_WAIT_%=;\n\t"
                 "IPB_DONE_%=: }" :: "r"(b), "r"(parity) : "memory");
#endif
}
