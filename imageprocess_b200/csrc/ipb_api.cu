// ipb_api.cu -- the C ABI of libipb200.so (declared in include/ipb200.h).
//
// Conventions: every entry point is asynchronous on the caller's
// cudaStream_t, takes caller-owned DEVICE pointers plus sizes, never allocates or frees,
// keeps no global mutable state (the last-error string is thread-local) and returns
// 0 or a negative IPB_ERR_* code.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "ipb_rt.cuh"
#include "ipb_exact.cuh"
#include "ipb_raster.cuh"

static thread_local char g_ipb_err[512] = "";

void ipb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ipb_err, sizeof(g_ipb_err), fmt, ap);
    va_end(ap);
}

int ipb_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ipb_set_error("%s: %s", what, cudaGetErrorString(e));
        return IPB_ERR_CUDA;
    }
    return IPB_OK;
}

#define IPB_CUDA_TRY(expr, what)                                          \
    do {                                                                  \
        cudaError_t e__ = (expr);                                         \
        if (e__ != cudaSuccess) {                                         \
            ipb_set_error("%s: %s", what, cudaGetErrorString(e__));       \
            return IPB_ERR_CUDA;                                          \
        }                                                                 \
    } while (0)

extern "C" {

const char* ipb_last_error(void) { return g_ipb_err; }

int ipb_version(void) { return 100; }

int ipb_is_emulated(void) {
#ifdef IPB_EMULATE
    return 1;
#else
    return 0;
#endif
}

int ipb_rasterize_rois(int rule, int n_rois, const double* verts_xy, const int32_t* vert_off,
                       const int32_t* erect, const int32_t* srect, const int32_t* org,
                       const int32_t* roi_frame, const int64_t* mask_off, int max_rows, int max_wpr,
                       uint32_t* mask_pool, uint32_t* area, uint32_t* union_bits, int union_wpr,
                       int frame_h, void* stream)
{
    IPB_REQUIRE(rule == IPB_RULE_MPL || rule == IPB_RULE_SK, "ipb_rasterize_rois: bad rule %d", rule);
    IPB_REQUIRE(n_rois >= 0 && n_rois <= 65535, "ipb_rasterize_rois: n_rois %d out of range", n_rois);
    IPB_REQUIRE(max_wpr >= 0 && max_wpr <= 1024, "ipb_rasterize_rois: max_wpr %d out of range", max_wpr);
    if (n_rois == 0) return IPB_OK;
    IPB_REQUIRE(verts_xy && vert_off && erect && srect && org && roi_frame && mask_off && mask_pool && area,
                "ipb_rasterize_rois: null pointer");
    IPB_CUDA_TRY(cudaMemsetAsync(area, 0, sizeof(uint32_t) * (size_t)n_rois, (cudaStream_t)stream), "memset area");
    if (max_rows <= 0 || max_wpr <= 0) return IPB_OK;
    dim3 grid(ipb_div_up(max_rows, IPB_RASTER_WARPS), (unsigned)n_rois);
    dim3 block(IPB_RASTER_WARPS * 32);
    size_t smem = (size_t)IPB_RASTER_WARPS * 5 * (size_t)max_wpr * sizeof(unsigned);
    if (rule == IPB_RULE_MPL) {
        auto k = ipb_k_raster<IPB_RULE_MPL>;
        if (smem > 48 * 1024) IPB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "raster smem");
        IPB_LAUNCH(k, grid, block, smem, stream, n_rois, (const double2*)verts_xy, vert_off, (const int4*)erect,
                   (const int4*)srect, (const int2*)org, roi_frame, (const long long*)mask_off, max_wpr,
                   mask_pool, area, union_bits, union_wpr, frame_h);
    } else {
        auto k = ipb_k_raster<IPB_RULE_SK>;
        if (smem > 48 * 1024) IPB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "raster smem");
        IPB_LAUNCH(k, grid, block, smem, stream, n_rois, (const double2*)verts_xy, vert_off, (const int4*)erect,
                   (const int4*)srect, (const int2*)org, roi_frame, (const long long*)mask_off, max_wpr,
                   mask_pool, area, union_bits, union_wpr, frame_h);
    }
    return ipb_check_launch("ipb_k_raster");
}

}  // extern "C"
