// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A minimal CPU emulator of the CUDA execution model, used only to run the product's
// kernels (imageprocess_b200/csrc/*.cuh, unmodified) inside the build container, which
// has no GPU.  It lets `pytest -m "not gpu"` exercise the kernels' *logic* before GPU
// minutes are spent.  The product package never loads the emulated library: it loads
// libipb200.so (nvcc, sm_100a) only and fails loudly when that is missing.
//
// Model: one block at a time; every CUDA thread of the block is a fiber (hand-rolled
// x86-64 context switch) scheduled round-robin on one OS thread.  __syncthreads() and
// the warp collectives yield until all participants arrived, so barrier/shuffle
// semantics are exact for converged code; atomics are plain read-modify-writes.
// Data races are not detected as such (use compute-sanitizer on the GPU for that), but the order in which
// the threads of a block (and the blocks of a grid) run is selectable -- IPB_EMU_ORDER = forward | reverse |
// random[:seed], each optionally with ":preempt" (a thread yields after every atomic) -- and correctly
// synchronised kernels give identical results under all of them (tests/test_emu_orders.py).
#pragma once
#if !defined(__x86_64__)
#error "cuda_emu.h: x86-64 only"
#endif
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <sys/mman.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

// ---------------------------------------------------------------- vector types
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct ushort4 { unsigned short x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace emu {

struct Fiber {
    void* sp = nullptr;
    bool done = false;
    uint3_ tid{0, 0, 0};
    unsigned linear = 0;
};

struct WarpState {
    unsigned long long buf[32];
    unsigned arrived = 0, departed = 0;
    unsigned gen = 0;
};

struct BlockCtx {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    unsigned nthreads = 0, alive = 0;
    unsigned bar_arrived = 0, bar_gen = 0;
    void* sched_sp = nullptr;
    std::function<void()> body;
    unsigned char* dyn_smem = nullptr;
};

extern BlockCtx* g_blk;
extern Fiber* g_cur;
extern uint3_ g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern int g_preempt;

extern "C" void emu_switch(void** from_sp, void* to_sp);
void yield();
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body);

}  // namespace emu

#define threadIdx (emu::g_cur->tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

// ---------------------------------------------------------------- barriers / collectives
static inline void __syncthreads() {
    emu::BlockCtx* b = emu::g_blk;
    unsigned my = b->bar_gen;
    b->bar_arrived++;
    while (b->bar_gen == my) {
        if (b->bar_arrived >= b->alive) { b->bar_arrived = 0; b->bar_gen++; break; }
        emu::yield();
    }
}

namespace emu {
// All lanes in `mask` contribute v; returns pointer to a private snapshot of 32 values.
static inline void warp_exchange(unsigned mask, unsigned long long v, unsigned long long out[32]) {
    BlockCtx* b = g_blk;
    unsigned lane = g_cur->linear & 31u;
    WarpState& w = b->warps[g_cur->linear >> 5];
    w.buf[lane] = v;
    w.arrived |= (1u << lane);
    while ((w.arrived & mask) != mask) yield();
    for (int i = 0; i < 32; ++i) out[i] = w.buf[i];
    w.departed |= (1u << lane);
    if ((w.departed & mask) == mask) {
        w.arrived &= ~mask;
        w.departed &= ~mask;
        w.gen++;
    } else {
        unsigned g = w.gen;
        while (w.gen == g) yield();
    }
}
static inline unsigned lane_id() { return g_cur->linear & 31u; }
}  // namespace emu

static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    unsigned long long t[32];
    emu::warp_exchange(mask, 0, t);
}
static inline unsigned __activemask() { return 0xffffffffu; }

template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shfl");
    unsigned long long raw = 0, t[32];
    std::memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, t);
    unsigned lane = emu::lane_id();
    unsigned s = (lane & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
    T r;
    std::memcpy(&r, &t[s], sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    unsigned long long raw = 0, t[32];
    std::memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, t);
    unsigned s = emu::lane_id() ^ (unsigned)lanemask;
    T r;
    std::memcpy(&r, &t[s & 31], sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned long long raw = 0, t[32];
    std::memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, t);
    unsigned lane = emu::lane_id();
    unsigned base = lane & ~(unsigned)(width - 1);
    int s = (int)lane - (int)delta;
    if (s < (int)base) s = (int)lane;
    T r;
    std::memcpy(&r, &t[s], sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned long long raw = 0, t[32];
    std::memcpy(&raw, &v, sizeof(T));
    emu::warp_exchange(mask, raw, t);
    unsigned lane = emu::lane_id();
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned s = lane + delta;
    if (s >= base + (unsigned)width) s = lane;
    T r;
    std::memcpy(&r, &t[s], sizeof(T));
    return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned long long t[32];
    emu::warp_exchange(mask, pred ? 1ull : 0ull, t);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r |= (t[i] ? 1u : 0u) << i;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    unsigned long long t[32];
    emu::warp_exchange(mask, v, t);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r += (unsigned)t[i];
    return r;
}

// ---------------------------------------------------------------- atomics (one OS thread)
// With IPB_EMU_ORDER=...:preempt a thread gives up its turn right after every atomic, so the sections of
// different threads interleave at their atomics instead of running one whole section after the other.
#define EMU_ATOMIC_RMW(name, T, expr) \
    static inline T name(T* a, T v) { T old = *a; *a = (expr); if (emu::g_preempt) emu::yield(); return old; }
EMU_ATOMIC_RMW(atomicAdd, int, old + v)
EMU_ATOMIC_RMW(atomicAdd, unsigned, old + v)
EMU_ATOMIC_RMW(atomicAdd, unsigned long long, old + v)
EMU_ATOMIC_RMW(atomicAdd, float, old + v)
EMU_ATOMIC_RMW(atomicAdd, double, old + v)
EMU_ATOMIC_RMW(atomicSub, int, old - v)
EMU_ATOMIC_RMW(atomicSub, unsigned, old - v)
EMU_ATOMIC_RMW(atomicMin, int, std::min(old, v))
EMU_ATOMIC_RMW(atomicMin, unsigned, std::min(old, v))
EMU_ATOMIC_RMW(atomicMin, unsigned long long, std::min(old, v))
EMU_ATOMIC_RMW(atomicMin, long long, std::min(old, v))
EMU_ATOMIC_RMW(atomicMax, int, std::max(old, v))
EMU_ATOMIC_RMW(atomicMax, unsigned, std::max(old, v))
EMU_ATOMIC_RMW(atomicMax, unsigned long long, std::max(old, v))
EMU_ATOMIC_RMW(atomicMax, long long, std::max(old, v))
EMU_ATOMIC_RMW(atomicOr, int, old | v)
EMU_ATOMIC_RMW(atomicOr, unsigned, old | v)
EMU_ATOMIC_RMW(atomicOr, unsigned long long, old | v)
EMU_ATOMIC_RMW(atomicAnd, int, old & v)
EMU_ATOMIC_RMW(atomicAnd, unsigned, old & v)
EMU_ATOMIC_RMW(atomicXor, int, old ^ v)
EMU_ATOMIC_RMW(atomicXor, unsigned, old ^ v)
EMU_ATOMIC_RMW(atomicExch, int, v)
EMU_ATOMIC_RMW(atomicExch, unsigned, v)
EMU_ATOMIC_RMW(atomicExch, unsigned long long, v)
EMU_ATOMIC_RMW(atomicExch, float, v)
#undef EMU_ATOMIC_RMW
template <typename T>
static inline T atomicCAS(T* a, T cmp, T v) { T old = *a; if (old == cmp) *a = v; if (emu::g_preempt) emu::yield(); return old; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---------------------------------------------------------------- intrinsics
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    s &= 31; unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)((v << s) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    s &= 31; unsigned long long v = ((unsigned long long)hi << 32) | lo; return (unsigned)(v >> s);
}
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; std::memcpy(&d, &u, 8); return d; }
// compiled with -ffp-contract=off -msse2: every op below rounds once, like the _rn intrinsics
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float __ll2float_rn(long long v) { return (float)v; }
static inline float __ull2float_rn(unsigned long long v) { return (float)v; }
static inline float __int2float_rn(int v) { return (float)v; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }
static inline float __double2float_rn(double v) { return (float)v; }
static inline double __ll2double_rn(long long v) { return (double)v; }
static inline int __float2int_rz(float v) { return (int)v; }
static inline int __double2int_rz(double v) { return (int)v; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
using std::isfinite;
using std::isnan;
using std::max;
using std::min;
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

// ---------------------------------------------------------------- runtime shims
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

#define IPB_EMU_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
