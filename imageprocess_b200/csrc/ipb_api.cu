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
#include "ipb_hist.cuh"
#include "ipb_roistats.cuh"
#include "ipb_roifused.cuh"
#include "ipb_fret.cuh"
#include "ipb_fa.cuh"
#include "ipb_morph.cuh"
#include "ipb_contour.cuh"
#include "ipb_gauss.cuh"
#include "ipb_graymorph.cuh"

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

template <int SRC>
static int ipb_launch_region_stats(const void* regions, const void* jobs, int n_jobs, const uint32_t* mask_pool,
                                   const uint32_t* and_bits, int and_wpr, int H, int W, const uint16_t* planes,
                                   const float* images, const float* bvals, void* out, const uint8_t* only, void* stream)
{
    const int smem = IPB_RS_SMEM_BYTES;
    IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_region_stats<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "region_stats smem");
    // a rerun of flagged regions: every CTA scans 32 jobs of the list at a time
    const int grid = only ? (((n_jobs + 31) / 32) < 148 ? ((n_jobs + 31) / 32) : 148) : n_jobs;
    IPB_LAUNCH(ipb_k_region_stats<SRC>, dim3(grid), dim3(IPB_RS_THREADS), (size_t)smem, stream,
               (const IpbRegion*)regions, (const IpbStatJob*)jobs, n_jobs, mask_pool, and_bits, and_wpr, H, W,
               planes, images, bvals, (IpbStatOut*)out, smem, (const unsigned char*)only);
    return ipb_check_launch("ipb_k_region_stats");
}

template <bool MAX>
static int ipb_launch_graymorph(const uint16_t* in, uint16_t* tmp, uint16_t* out, int n, int H, int W, int radius, void* stream)
{
    const int use_tma = (W % 8 == 0) && (((size_t)in | (size_t)tmp) % 16 == 0) && ((size_t)H * W % 8 == 0);
    const dim3 block(IPB_GS_THREADS);
    const size_t smem0 = (size_t)(IPB_GM_TILE_H0 + 2 * radius) * IPB_GM_TILE_W * sizeof(uint16_t);
    const size_t smem1 = (size_t)IPB_GM_TILE_H1 * (IPB_GM_TILE_W + 2 * ((radius + 7) & ~7)) * sizeof(uint16_t);
    auto k0 = ipb_k_graymorph_pass<0, MAX>;
    auto k1 = ipb_k_graymorph_pass<1, MAX>;
    IPB_CUDA_TRY(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0), "graymorph smem");
    IPB_LAUNCH(k0, dim3(ipb_div_up(W, IPB_GM_TILE_W), ipb_div_up(H, IPB_GM_TILE_H0), (unsigned)n), block, smem0, stream,
               in, tmp, H, W, radius, use_tma);
    int rc = ipb_check_launch("ipb_k_graymorph_pass<0>");
    if (rc) return rc;
    IPB_LAUNCH(k1, dim3(ipb_div_up(W, IPB_GM_TILE_W), ipb_div_up(H, IPB_GM_TILE_H1), (unsigned)n), block, smem1, stream,
               tmp, out, H, W, radius, use_tma);
    return ipb_check_launch("ipb_k_graymorph_pass<1>");
}

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

// ---------------------------------------------------------------- histograms / quantiles
static int ipb_launch_hist_full(const uint16_t* planes, int H, int W, const void* jobs, int n_jobs,
                                const uint32_t* union_bits, int union_wpr, uint32_t* hist, uint64_t* stats,
                                void* stream);

int ipb_hist_u16(const uint16_t* planes, int H, int W, const void* jobs, int n_jobs,
                 int has_masked_stride, const uint32_t* union_bits, int union_wpr,
                 uint64_t* row_rank_scratch, uint32_t* hist, uint64_t* stats, void* stream)
{
    IPB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535, "ipb_hist_u16: n_jobs %d out of range", n_jobs);
    if (n_jobs == 0) return IPB_OK;
    IPB_REQUIRE(planes && jobs && hist && stats && H > 0 && W > 0, "ipb_hist_u16: bad argument");
    IPB_REQUIRE(!has_masked_stride || (union_bits && row_rank_scratch), "ipb_hist_u16: masked stride needs union + scratch");
    cudaStream_t st = (cudaStream_t)stream;
    IPB_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)IPB_HIST_BINS * n_jobs, st), "memset hist");
    IPB_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(uint64_t) * 4 * (size_t)n_jobs, st), "memset stats");
    int rc = ipb_launch_hist_full(planes, H, W, jobs, n_jobs, union_bits, union_wpr, hist, stats, stream);
    if (rc) return rc;
    if (has_masked_stride) {
        IPB_LAUNCH(ipb_k_hist_masked_stride, dim3(n_jobs), dim3(256), 0, stream, planes, H, W,
                   (const IpbHistJob*)jobs, union_bits, union_wpr, (unsigned long long*)row_rank_scratch,
                   hist, (unsigned long long*)stats);
        rc = ipb_check_launch("ipb_k_hist_masked_stride");
    }
    return rc;
}

// one launch of the full-range histogram kernel over all jobs
static int ipb_launch_hist_full(const uint16_t* planes, int H, int W, const void* jobs, int n_jobs,
                                const uint32_t* union_bits, int union_wpr, uint32_t* hist, uint64_t* stats,
                                void* stream)
{
    int chunks = (296 * 6 + n_jobs - 1) / n_jobs;       // ~6 waves of the 296 resident CTAs
    int max_chunks = H / 16 > 0 ? H / 16 : 1;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    const int rows_per_chunk = (H + chunks - 1) / chunks;
    chunks = (H + rows_per_chunk - 1) / rows_per_chunk;
    const size_t smem = sizeof(unsigned) * IPB_HIST_WIN;
    IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_hist_u16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "hist smem");
    IPB_LAUNCH(ipb_k_hist_u16, dim3(chunks, n_jobs), dim3(IPB_HIST_THREADS), smem, stream,
               planes, H, W, (const IpbHistJob*)jobs, rows_per_chunk, union_bits, union_wpr,
               hist, (unsigned long long*)stats);
    return ipb_check_launch("ipb_k_hist_u16");
}

int ipb_hist_select(const uint16_t* planes, int H, int W, const void* jobs, int n_jobs,
                    const void* passes, int n_passes, const void* qjobs, int n_q,
                    uint32_t* hist_win, void* win, uint64_t* cnt, uint64_t* stats,
                    void* qout, uint32_t* miss, void* stream)
{
    IPB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535 && n_passes >= 0 && n_passes <= 65535, "ipb_hist_select: job count out of range");
    if (n_jobs == 0 || n_passes == 0) return IPB_OK;
    IPB_REQUIRE(planes && jobs && passes && hist_win && win && cnt && stats && miss && H > 0 && W > 0,
                "ipb_hist_select: bad argument");
    IPB_REQUIRE(n_q == 0 || (qjobs && qout), "ipb_hist_select: quantile jobs without buffers");
    IPB_REQUIRE((W & 7) == 0 && W >= 16 && (((size_t)planes) & 15) == 0,
                "ipb_hist_select: needs W %% 8 == 0, W >= 16 and 16-byte aligned planes");    // W == 8: one unit per row has no 32-bit multiply-high constant
    int rc;
    const size_t smem = sizeof(unsigned) * IPB_PQ_SBINS;
    IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_pq_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pq sample smem");
    IPB_LAUNCH(ipb_k_pq_sample, dim3(n_passes), dim3(1024), smem, stream, planes, H, W, (const IpbPlanePass*)passes,
               (const IpbHistJob*)jobs, (const IpbQJob*)qjobs, n_q, hist_win, (IpbHistWin*)win,
               (unsigned long long*)cnt, (unsigned long long*)stats);
    if ((rc = ipb_check_launch("ipb_k_pq_sample"))) return rc;
    // 4 CTAs of 256 threads per SM = 592 resident CTAs; ~4 waves, at least 4 trips of 4 units per thread
    const unsigned long long U = ((unsigned long long)H * (unsigned long long)W) >> 3;
    unsigned long long chunks = (592ull * 4 + n_passes - 1) / n_passes;
    const unsigned long long min_units = 16ull * IPB_PQ_THREADS;
    if (chunks * min_units > U) chunks = U / min_units;
    if (chunks < 1) chunks = 1;
    const unsigned long long upc = (U + chunks - 1) / chunks;
    IPB_REQUIRE(upc < 0xffffffffull, "ipb_hist_select: plane too large");
    chunks = (U + upc - 1) / upc;
    IPB_LAUNCH(ipb_k_pq_count, dim3((unsigned)chunks, n_passes), dim3(IPB_PQ_THREADS), 0, stream, planes, H, W,
               (const IpbPlanePass*)passes, (const IpbHistJob*)jobs, (const IpbHistWin*)win, (unsigned)upc,
               hist_win, (unsigned long long*)cnt, (unsigned long long*)stats);
    if ((rc = ipb_check_launch("ipb_k_pq_count"))) return rc;
    if (n_q > 0) {
        IPB_LAUNCH(ipb_k_pq_select, dim3(n_q), dim3(256), 0, stream, (const IpbQJob*)qjobs, (const IpbHistWin*)win,
                   (const unsigned long long*)cnt, hist_win, (const unsigned long long*)stats,
                   (IpbQOut*)qout, miss);
        if ((rc = ipb_check_launch("ipb_k_pq_select"))) return rc;
    }
    return IPB_OK;
}

int ipb_hist_planes(const uint16_t* planes, int H, int W, const void* jobs, int n_jobs,
                    const void* passes, int n_passes, int has_masked_stride, const uint32_t* union_bits,
                    int union_wpr, uint64_t* row_rank_scratch, uint32_t* hist, uint64_t* stats, void* stream)
{
    IPB_REQUIRE(n_jobs >= 0 && n_jobs <= 65535 && n_passes >= 0 && n_passes <= 65535, "ipb_hist_planes: job count out of range");
    if (n_jobs == 0) return IPB_OK;
    IPB_REQUIRE(planes && jobs && passes && hist && stats && H > 0 && W > 0, "ipb_hist_planes: bad argument");
    IPB_REQUIRE(!has_masked_stride || (union_bits && row_rank_scratch), "ipb_hist_planes: masked stride needs union + scratch");
    cudaStream_t st = (cudaStream_t)stream;
    IPB_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)IPB_HIST_BINS * n_jobs, st), "memset hist");
    IPB_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(uint64_t) * 4 * (size_t)n_jobs, st), "memset stats");
    int rc = IPB_OK;
    if (n_passes > 0) {
        // 2 CTAs of 1024 threads per SM = 296 resident CTAs; ~6 waves keep the last, partly filled
        // wave short against the whole launch
        int chunks = (296 * 6 + n_passes - 1) / n_passes;
        int max_chunks = H / 16 > 0 ? H / 16 : 1;
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        const int rows_per_chunk = (H + chunks - 1) / chunks;
        chunks = (H + rows_per_chunk - 1) / rows_per_chunk;
        const size_t smem = sizeof(unsigned) * IPB_HIST_WIN;
        IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_hist_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "hist planes smem");
        IPB_LAUNCH(ipb_k_hist_planes, dim3(chunks, n_passes), dim3(IPB_HIST_THREADS), smem, stream, planes, H, W,
                   (const IpbPlanePass*)passes, (const IpbHistJob*)jobs, rows_per_chunk, union_bits, union_wpr,
                   hist, (unsigned long long*)stats);
        if ((rc = ipb_check_launch("ipb_k_hist_planes"))) return rc;
    }
    if (has_masked_stride) {
        IPB_LAUNCH(ipb_k_hist_masked_stride, dim3(n_jobs), dim3(256), 0, stream, planes, H, W,
                   (const IpbHistJob*)jobs, union_bits, union_wpr, (unsigned long long*)row_rank_scratch,
                   hist, (unsigned long long*)stats);
        rc = ipb_check_launch("ipb_k_hist_masked_stride");
    }
    return rc;
}

int ipb_hist_quantiles(const uint32_t* hist, const uint64_t* stats, const void* qjobs, int n_q,
                       void* qout, void* stream)
{
    if (n_q <= 0) return IPB_OK;
    IPB_REQUIRE(hist && stats && qjobs && qout, "ipb_hist_quantiles: null pointer");
    IPB_LAUNCH(ipb_k_hist_quantiles, dim3(n_q), dim3(1024), 0, stream, hist,
               (const unsigned long long*)stats, (const IpbQJob*)qjobs, (IpbQOut*)qout);
    return ipb_check_launch("ipb_k_hist_quantiles");
}

int ipb_scatter_qvalues(const void* qout, const int32_t* dst_idx, int n, float* dst, void* stream)
{
    if (n <= 0) return IPB_OK;
    IPB_REQUIRE(qout && dst_idx && dst, "ipb_scatter_qvalues: null pointer");
    IPB_LAUNCH(ipb_k_scatter_qvalues, dim3(ipb_div_up(n, 128)), dim3(128), 0, stream,
               (const IpbQOut*)qout, dst_idx, n, dst);
    return ipb_check_launch("ipb_k_scatter_qvalues");
}

int ipb_fret_eps(const void* qout_eps, int n_frames, int denom_slot, int clip_neg, float eps_abs,
                 float* fparams, void* stream)
{
    if (n_frames <= 0) return IPB_OK;
    IPB_REQUIRE(qout_eps && fparams && (denom_slot == IPB_FP_BD || denom_slot == IPB_FP_BA), "ipb_fret_eps: bad argument");
    IPB_LAUNCH(ipb_k_fret_eps, dim3(ipb_div_up(n_frames, 128)), dim3(128), 0, stream,
               (const IpbQOut*)qout_eps, n_frames, denom_slot, clip_neg, eps_abs, fparams);
    return ipb_check_launch("ipb_k_fret_eps");
}

int ipb_fa_params(const uint64_t* stats, const int32_t* stat_idx, const void* qout_bg, int n_frames,
                  int64_t npx, float alpha, float* fa, void* stream)
{
    if (n_frames <= 0) return IPB_OK;
    IPB_REQUIRE(stats && stat_idx && qout_bg && fa && npx > 0, "ipb_fa_params: bad argument");
    IPB_LAUNCH(ipb_k_fa_params, dim3(ipb_div_up(n_frames, 128)), dim3(128), 0, stream,
               (const unsigned long long*)stats, stat_idx, (const IpbQOut*)qout_bg, n_frames,
               (long long)npx, alpha, fa);
    return ipb_check_launch("ipb_k_fa_params");
}

// ---------------------------------------------------------------- fused FRET pass
int ipb_fret_pixels(const uint16_t* planes, int n_frames, int H, int W, const void* cfg_host,
                    const float* fparams, const uint32_t* union_bits, int union_wpr,
                    const int32_t* union_idx, float* R, float* Ralt, float* Rroi, float* Dcorr,
                    float* Acorr, uint64_t* mom_stats, const int32_t* mom_idx, int mom_acceptor, void* stream)
{
    if (n_frames <= 0) return IPB_OK;
    IPB_REQUIRE(planes && cfg_host && fparams && H > 0 && W > 0, "ipb_fret_pixels: bad argument");
    IPB_REQUIRE(!mom_stats || mom_idx, "ipb_fret_pixels: moments need mom_idx");
    IpbFretCfg cfg;
    memcpy(&cfg, cfg_host, sizeof(cfg));
    IPB_REQUIRE(cfg.n_ch > 0 && cfg.donor_ch >= 0 && cfg.donor_ch < cfg.n_ch && cfg.acc_ch >= 0 &&
                cfg.acc_ch < cfg.n_ch && cfg.aonly_ch < cfg.n_ch, "ipb_fret_pixels: bad channel indices");
    IPB_REQUIRE(n_frames <= 65535, "ipb_fret_pixels: n_frames %d out of range", n_frames);
    const long long work = (long long)H * W / 8;                      // per frame
    long long blocks = (work + 255) / 256;
    long long cap = (148LL * 8 * 4 + n_frames - 1) / n_frames;        // ~32 CTAs per SM over the whole grid
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const bool plain = !cfg.sat_on && !cfg.use_spectral && !cfg.clip_on;
#define IPB_FRET_LAUNCH(ALT, PLAIN)                                                                                    \
    IPB_LAUNCH((ipb_k_fret_pixels<ALT, PLAIN>), dim3((unsigned)blocks, (unsigned)n_frames), dim3(256), 0, stream, planes, \
               n_frames, H, W, cfg, fparams, union_bits, union_wpr, union_idx, R, Ralt, Rroi, Dcorr, Acorr,                \
               (unsigned long long*)mom_stats, mom_idx, mom_acceptor)
    if (Ralt) { if (plain) IPB_FRET_LAUNCH(true, true); else IPB_FRET_LAUNCH(true, false); }
    else      { if (plain) IPB_FRET_LAUNCH(false, true); else IPB_FRET_LAUNCH(false, false); }
#undef IPB_FRET_LAUNCH
    return ipb_check_launch("ipb_k_fret_pixels");
}

// ---------------------------------------------------------------- region statistics
int ipb_region_stats(const void* regions, const void* jobs, int n_jobs, int uniform_src,
                     const uint32_t* mask_pool, const uint32_t* and_bits, int and_wpr, int H, int W,
                     const uint16_t* planes, const float* images, const float* bvals, void* out,
                     const uint8_t* only, void* stream)
{
    if (n_jobs <= 0) return IPB_OK;
    IPB_REQUIRE(regions && jobs && mask_pool && out && H > 0 && W > 0, "ipb_region_stats: bad argument");
    IPB_REQUIRE(uniform_src >= -1 && uniform_src <= IPB_SRC_RATIO, "ipb_region_stats: bad uniform_src %d", uniform_src);
    int rc = IPB_OK;
    // uniform_src = -1: the job list may mix sources; each instantiation skips the other's jobs
    // (a source whose buffer is NULL is not launched at all)
    if (uniform_src == IPB_SRC_U16 || (uniform_src < 0 && planes)) {
        IPB_REQUIRE(planes, "ipb_region_stats: uint16 jobs need planes");
        rc = ipb_launch_region_stats<IPB_SRC_U16>(regions, jobs, n_jobs, mask_pool, and_bits, and_wpr, H, W,
                                                  planes, images, bvals, out, only, stream);
        if (rc) return rc;
    }
    if (uniform_src == IPB_SRC_F32 || (uniform_src < 0 && images)) {
        IPB_REQUIRE(images, "ipb_region_stats: float32 jobs need images");
        rc = ipb_launch_region_stats<IPB_SRC_F32>(regions, jobs, n_jobs, mask_pool, and_bits, and_wpr, H, W,
                                                  planes, images, bvals, out, only, stream);
        if (rc) return rc;
    }
    if (uniform_src == IPB_SRC_RATIO || (uniform_src < 0 && images && bvals)) {
        IPB_REQUIRE(images && bvals, "ipb_region_stats: ratio jobs need images and parameters");
        rc = ipb_launch_region_stats<IPB_SRC_RATIO>(regions, jobs, n_jobs, mask_pool, and_bits, and_wpr, H, W,
                                                    planes, images, bvals, out, only, stream);
    }
    return rc;
}

int ipb_roi_stats_fused(const void* regions, int n_regions, const void* jobs, int n_jobs, const uint32_t* mask_pool,
                        int H, int W, const uint16_t* planes, const float* bvals, void* out, uint32_t* scratch,
                        int64_t stride_words, int n_ctas, uint32_t* counter, uint8_t* flags, uint8_t* wide_flags,
                        void* stream)
{
    IPB_REQUIRE(n_regions >= 0 && n_jobs >= 0, "ipb_roi_stats_fused: negative count");
    IPB_REQUIRE(flags && counter && wide_flags, "ipb_roi_stats_fused: null flags / counter");
    IPB_CUDA_TRY(cudaMemsetAsync(counter, 0, 2 * sizeof(uint32_t), (cudaStream_t)stream), "memset counter");
    if (n_regions > 0) IPB_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)n_regions, (cudaStream_t)stream), "memset flags");
    if (n_jobs == 0) return IPB_OK;
    IPB_CUDA_TRY(cudaMemsetAsync(wide_flags, 0, (size_t)n_jobs, (cudaStream_t)stream), "memset wide flags");
    IPB_REQUIRE(regions && jobs && mask_pool && planes && bvals && out && scratch && stride_words > 0 &&
                stride_words * (int64_t)n_ctas < (1ll << 32) && n_ctas > 0 && H > 0 && W > 0, "ipb_roi_stats_fused: bad argument");
    const int smem = IPB_RF_SMEM_BYTES;
    IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_roi_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "roi_fused smem");
    IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_roi_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "roi_fused smem");
    const int grid = n_ctas < n_jobs ? n_ctas : n_jobs;
    IPB_LAUNCH(ipb_k_roi_fused<false>, dim3(grid), dim3(IPB_RF_THREADS), (size_t)smem, stream, (const IpbRegion*)regions,
               (const IpbRoiJob*)jobs, n_jobs, mask_pool, H, W, planes, bvals, (IpbStatOut*)out, scratch,
               (unsigned long long)stride_words, counter, (unsigned char*)flags, (unsigned char*)wide_flags);
    int rc = ipb_check_launch("ipb_k_roi_fused");
    if (rc) return rc;
    // the jobs handed on by the first launch (usually none: every CTA scans its share of the flags and leaves)
    IPB_LAUNCH(ipb_k_roi_fused<true>, dim3(grid), dim3(IPB_RF_THREADS), (size_t)smem, stream, (const IpbRegion*)regions,
               (const IpbRoiJob*)jobs, n_jobs, mask_pool, H, W, planes, bvals, (IpbStatOut*)out, scratch,
               (unsigned long long)stride_words, counter + 1, (unsigned char*)flags, (unsigned char*)wide_flags);
    return ipb_check_launch("ipb_k_roi_fused<wide>");
}

// ---------------------------------------------------------------- focal-adhesion chain
int ipb_fa_segment(const void* crops, int n_crops, int max_rows, int64_t total_rows,
                   const uint16_t* planes, int H, int W, const float* fa_params,
                   const uint32_t* roi_mask, double min_size, int close_radius,
                   uint32_t* bw_a, uint32_t* bw_b, int32_t* L, uint32_t* csize, uint32_t* rootbits,
                   int32_t* row_roots, int32_t* row_base, int32_t* crop_count,
                   uint32_t* bw_final, int32_t* comp_off, void* comps, int comp_cap,
                   int32_t* labels, int path, const int32_t* crop_order, int label_conn, void* stream)
{
    if (n_crops <= 0) return IPB_OK;
    IPB_REQUIRE(path >= 0 && path <= 3, "ipb_fa_segment: path %d not in 0..3", path);
    IPB_REQUIRE(label_conn == 8 || label_conn == 4, "ipb_fa_segment: label_conn %d not 4 or 8", label_conn);
    IPB_REQUIRE(n_crops <= 65535, "ipb_fa_segment: n_crops %d out of range", n_crops);
    IPB_REQUIRE(crops && planes && fa_params && roi_mask && bw_a && bw_b && L && csize && rootbits &&
                row_roots && row_base && crop_count && bw_final && comp_off && comps,
                "ipb_fa_segment: null pointer");
    IPB_REQUIRE(close_radius >= 0 && close_radius <= 5, "ipb_fa_segment: close_radius %d not in 0..5", close_radius);
    IPB_REQUIRE(max_rows > 0 && total_rows > 0 && comp_cap > 0, "ipb_fa_segment: bad sizes");
    const IpbCrop* cr = (const IpbCrop*)crops;
    const dim3 grid(ipb_div_up(max_rows, IPB_FA_ROWS), (unsigned)n_crops), block(IPB_FA_THREADS);
    int rc;
    IpbDisk disk;
    memset(&disk, 0, sizeof(disk));
    disk.r = close_radius;
    for (int dy = -close_radius; dy <= close_radius; ++dy) {
        int k = 0;
        while ((k + 1) * (k + 1) + dy * dy <= close_radius * close_radius) ++k;
        disk.halfw[dy + close_radius] = k;
    }
    // many small crops (cell ROIs): one CTA per crop runs the whole chain; few / huge crops
    // (the mosaic): one kernel per phase, each crop spread over the chip
    // many small crops (cell ROIs): one CTA per crop runs the whole chain in shared memory (crops
    // that do not fit are flagged and taken by the global-memory variant right after); few / huge
    // crops (the mosaic): one kernel per phase, each crop spread over the chip.  path 3 = the
    // global-memory per-crop kernel for every crop.
    // (4-connected final labels -- scipy.ndimage.label's default, used by the ROI drawer's assist -- are
    // served by the one-kernel-per-phase path only)
    const bool fused = label_conn == 8 && (path == 1 || path == 3 || (path == 0 && n_crops >= 64 && max_rows <= 1024));
    if (fused) {
        if (path != 3) {
            IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_fa_fused_smem, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)IPB_FAS_SMEM_BYTES), "fa smem");
            IPB_LAUNCH(ipb_k_fa_fused_smem, dim3(n_crops), dim3(IPB_FAS_THREADS), IPB_FAS_SMEM_BYTES, stream, cr, planes, H, W,
                       fa_params, roi_mask, min_size, disk, L, csize, row_roots, row_base, crop_count, bw_final, labels, crop_order);
            if ((rc = ipb_check_launch("ipb_k_fa_fused_smem"))) return rc;
        }
        IPB_LAUNCH(ipb_k_fa_fused, dim3(n_crops), dim3(IPB_FA_FUSED_THREADS), 0, stream, cr, planes, H, W, fa_params,
                   roi_mask, min_size, disk, bw_a, bw_b, L, csize, rootbits, row_roots, row_base, crop_count, bw_final, crop_order,
                   path != 3 ? 1 : 0);
        if ((rc = ipb_check_launch("ipb_k_fa_fused"))) return rc;
    } else {
        IPB_LAUNCH(ipb_k_fa_threshold, grid, block, 0, stream, cr, planes, H, W, fa_params, roi_mask, bw_a);
        if ((rc = ipb_check_launch("ipb_k_fa_threshold"))) return rc;
        uint32_t* cur = bw_a;
        uint32_t* other = bw_b;
        if (min_size > 0) {
            IPB_LAUNCH(ipb_k_ccl_init, grid, block, 0, stream, cr, (const unsigned*)cur, L, csize);
            IPB_LAUNCH(ipb_k_ccl_merge<4>, grid, block, 0, stream, cr, (const unsigned*)cur, L);
            IPB_LAUNCH(ipb_k_ccl_flatten_size, grid, block, 0, stream, cr, (const unsigned*)cur, L, csize);
            IPB_LAUNCH(ipb_k_fa_size_filter, grid, block, 0, stream, cr, (const unsigned*)cur, (const int*)L,
                       (const unsigned*)csize, min_size, other);
            if ((rc = ipb_check_launch("ipb_fa small-object removal"))) return rc;
            uint32_t* t = cur; cur = other; other = t;
        }
        if (close_radius > 0) {
            IPB_LAUNCH(ipb_k_bits_morph<0>, grid, block, 0, stream, cr, (const unsigned*)cur, disk, other);
            IPB_LAUNCH(ipb_k_bits_morph<1>, grid, block, 0, stream, cr, (const unsigned*)other, disk, bw_final);
        } else {
            IPB_LAUNCH(ipb_k_bits_morph<0>, grid, block, 0, stream, cr, (const unsigned*)cur, disk, bw_final);
        }
        if ((rc = ipb_check_launch("ipb_k_bits_morph"))) return rc;
        IPB_CUDA_TRY(cudaMemsetAsync(row_roots, 0, sizeof(int32_t) * (size_t)total_rows, (cudaStream_t)stream), "memset row_roots");
        IPB_LAUNCH(ipb_k_ccl_init, grid, block, 0, stream, cr, (const unsigned*)bw_final, L, (unsigned*)nullptr);
        if (label_conn == 4) IPB_LAUNCH(ipb_k_ccl_merge<4>, grid, block, 0, stream, cr, (const unsigned*)bw_final, L);
        else IPB_LAUNCH(ipb_k_ccl_merge<8>, grid, block, 0, stream, cr, (const unsigned*)bw_final, L);
        IPB_LAUNCH(ipb_k_ccl_flatten_roots, grid, block, 0, stream, cr, (const unsigned*)bw_final, L, rootbits, row_roots);
        IPB_LAUNCH(ipb_k_fa_row_scan, dim3(n_crops), dim3(256), 0, stream, cr, (const int*)row_roots, row_base, crop_count);
        if ((rc = ipb_check_launch("ipb_fa labelling"))) return rc;
    }
    IPB_LAUNCH(ipb_k_fa_crop_scan, dim3(1), dim3(256), 0, stream, (const int*)crop_count, n_crops, comp_off);
    IPB_LAUNCH(ipb_k_fa_zero_comps, dim3(296), dim3(256), 0, stream, (const int*)comp_off, n_crops, comp_cap, (IpbComp*)comps);
    if (fused && path != 3) {
        IPB_LAUNCH(ipb_k_fa_gather_smem, dim3(8, (unsigned)n_crops), dim3(256), 0, stream, cr, (const int*)L, (const unsigned*)csize,
                   (const int*)row_roots, (const int*)row_base, (const int*)crop_count, (const int*)comp_off, comp_cap,
                   (IpbComp*)comps, labels);
    }
    IPB_LAUNCH(ipb_k_fa_props, grid, block, 0, stream, cr, (const unsigned*)bw_final, (const int*)L,
               (const unsigned*)rootbits, (const int*)row_base, (const int*)comp_off, comp_cap, planes, H, W,
               (IpbComp*)comps, labels);
    return ipb_check_launch("ipb_fa labelling");
}

// ---------------------------------------------------------------- morphology / moments / previews
int ipb_region_dilate(const void* regions, int n_regions, int max_w, int max_h, const uint32_t* in_pool,
                      int invert, const uint8_t* gmax_host, int R, uint8_t* g_scratch, const int64_t* g_off,
                      const uint32_t* and_pool, const uint32_t* andnot_pool, uint32_t* out_pool, void* stream)
{
    if (n_regions <= 0 || max_w <= 0 || max_h <= 0) return IPB_OK;
    IPB_REQUIRE(n_regions <= 65535, "ipb_region_dilate: n_regions %d out of range", n_regions);
    IPB_REQUIRE(regions && in_pool && gmax_host && g_scratch && g_off && out_pool, "ipb_region_dilate: null pointer");
    IPB_REQUIRE(R >= 0 && R <= IPB_MORPH_MAXR, "ipb_region_dilate: radius %d not in 0..%d", R, IPB_MORPH_MAXR);
    IpbSpan span;
    memset(&span, 0, sizeof(span));
    span.R = R;
    for (int i = 0; i <= R; ++i) {
        IPB_REQUIRE(gmax_host[i] <= IPB_MORPH_MAXR, "ipb_region_dilate: gmax[%d] too large", i);
        span.gmax[i] = gmax_host[i];
    }
    IPB_LAUNCH(ipb_k_morph_coldist, dim3(ipb_div_up(max_w, 128), (unsigned)n_regions), dim3(128), 0, stream,
               (const IpbRegion*)regions, in_pool, invert, (const long long*)g_off, g_scratch);
    int rc = ipb_check_launch("ipb_k_morph_coldist");
    if (rc) return rc;
    IPB_LAUNCH(ipb_k_morph_rowtest, dim3(ipb_div_up(max_h, 8), (unsigned)n_regions), dim3(256), 0, stream,
               (const IpbRegion*)regions, (const long long*)g_off, (const unsigned char*)g_scratch, span,
               and_pool, andnot_pool, out_pool);
    return ipb_check_launch("ipb_k_morph_rowtest");
}

int ipb_region_moments(const void* regions, int n_regions, const uint32_t* mask_pool, uint64_t* out, void* stream)
{
    if (n_regions <= 0) return IPB_OK;
    IPB_REQUIRE(regions && mask_pool && out, "ipb_region_moments: null pointer");
    IPB_LAUNCH(ipb_k_region_moments, dim3(n_regions), dim3(256), 0, stream, (const IpbRegion*)regions, mask_pool,
               (unsigned long long*)out);
    return ipb_check_launch("ipb_k_region_moments");
}

int ipb_preview_u16(const float* images, int64_t px_per_image, int n_images, const float* lohi,
                    uint16_t* out, void* stream)
{
    if (n_images <= 0 || px_per_image <= 0) return IPB_OK;
    IPB_REQUIRE(images && lohi && out, "ipb_preview_u16: null pointer");
    long long blocks = (px_per_image * n_images + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    IPB_LAUNCH(ipb_k_preview_u16, dim3((unsigned)blocks), dim3(256), 0, stream, images, (long long)px_per_image,
               n_images, lohi, out);
    return ipb_check_launch("ipb_k_preview_u16");
}

int ipb_crop_normalize(const void* jobs, int n_jobs, int64_t max_px, const uint16_t* planes, int H, int W,
                       const float* params, float inv_gamma, const void* regions, const uint32_t* mask_pool,
                       float* out_norm, uint16_t* out16, void* stream)
{
    if (n_jobs <= 0 || max_px <= 0) return IPB_OK;
    IPB_REQUIRE(n_jobs <= 65535, "ipb_crop_normalize: n_jobs %d out of range", n_jobs);
    IPB_REQUIRE(jobs && planes && params && (out_norm || out16), "ipb_crop_normalize: null pointer");
    long long bx = (max_px + 255) / 256;
    if (bx > 64) bx = 64;
    IPB_LAUNCH(ipb_k_crop_normalize, dim3((unsigned)bx, (unsigned)n_jobs), dim3(256), 0, stream,
               (const IpbCropJob*)jobs, planes, H, W, params, inv_gamma, (const IpbRegion*)regions, mask_pool,
               out_norm, out16);
    return ipb_check_launch("ipb_k_crop_normalize");
}

int ipb_eps_from_stat(const void* stat_out, const int32_t* row_of_frame, int n_frames, float eps_abs,
                      float* fparams, void* stream)
{
    if (n_frames <= 0) return IPB_OK;
    IPB_REQUIRE(stat_out && row_of_frame && fparams, "ipb_eps_from_stat: null pointer");
    IPB_LAUNCH(ipb_k_eps_from_stat, dim3(ipb_div_up(n_frames, 128)), dim3(128), 0, stream,
               (const IpbStatOut*)stat_out, row_of_frame, n_frames, eps_abs, fparams);
    return ipb_check_launch("ipb_k_eps_from_stat");
}

int ipb_fa_contour_cells(const void* crops, int n_crops, int64_t max_px, const int32_t* labels, void* rec,
                         uint32_t* rec_count, void* stream)
{
    if (n_crops <= 0) return IPB_OK;
    IPB_REQUIRE(crops && labels && rec && rec_count && max_px >= 0, "ipb_fa_contour_cells: bad argument");
    IPB_REQUIRE(n_crops <= 65535, "ipb_fa_contour_cells: n_crops %d out of range", n_crops);
    IPB_CUDA_TRY(cudaMemsetAsync(rec_count, 0, sizeof(uint32_t) * (size_t)n_crops, (cudaStream_t)stream), "memset rec_count");
    if (max_px == 0) return IPB_OK;
    unsigned gx = ipb_div_up(max_px, 256 * 4);
    if (gx > 2048u) gx = 2048u;
    IPB_LAUNCH(ipb_k_fa_contour_cells, dim3(gx, (unsigned)n_crops), dim3(256), 0, stream, (const IpbCrop*)crops, labels,
               (uint2*)rec, rec_count);
    return ipb_check_launch("ipb_k_fa_contour_cells");
}

int ipb_gaussian_f32(const float* in, float* tmp, float* out, int n_images, int H, int W,
                     const double* weights, int radius, void* stream)
{
    if (n_images <= 0) return IPB_OK;
    IPB_REQUIRE(in && tmp && out && weights && H > 0 && W > 0, "ipb_gaussian_f32: bad argument");
    IPB_REQUIRE(radius >= 0 && radius <= IPB_GS_MAXR, "ipb_gaussian_f32: radius %d not in 0..%d", radius, IPB_GS_MAXR);
    IPB_REQUIRE(n_images <= 65535, "ipb_gaussian_f32: n_images %d out of range", n_images);
    // TMA needs 16-byte aligned row segments; other shapes take the plain loads of the same kernels
    const int use_tma = (W % 4 == 0) && (((size_t)in | (size_t)tmp) % 16 == 0) && ((size_t)H * W % 4 == 0);
    const dim3 block(IPB_GS_THREADS);
    const size_t smem0 = (size_t)(IPB_GS_TILE_H0 + 2 * radius) * IPB_GS_TILE_W * sizeof(float);
    const size_t smem1 = (size_t)IPB_GS_TILE_H1 * (IPB_GS_TILE_W + 2 * ((radius + 3) & ~3)) * sizeof(float);
    IPB_CUDA_TRY(cudaFuncSetAttribute(ipb_k_gauss_pass<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0), "gauss smem");
    IPB_LAUNCH(ipb_k_gauss_pass<0>, dim3(ipb_div_up(W, IPB_GS_TILE_W), ipb_div_up(H, IPB_GS_TILE_H0), (unsigned)n_images), block,
               smem0, stream, in, tmp, H, W, weights, radius, use_tma);
    int rc = ipb_check_launch("ipb_k_gauss_pass<0>");
    if (rc) return rc;
    IPB_LAUNCH(ipb_k_gauss_pass<1>, dim3(ipb_div_up(W, IPB_GS_TILE_W), ipb_div_up(H, IPB_GS_TILE_H1), (unsigned)n_images), block,
               smem1, stream, tmp, out, H, W, weights, radius, use_tma);
    return ipb_check_launch("ipb_k_gauss_pass<1>");
}

int ipb_gauss_combine(const float* a, const float* b, float* out, int64_t n, int unsharp, float amount, void* stream)
{
    if (n <= 0) return IPB_OK;
    IPB_REQUIRE(a && b && out, "ipb_gauss_combine: null pointer");
    IPB_LAUNCH(ipb_k_gauss_combine, dim3(ipb_div_up(n, 256 * 8) > 4736u ? 4736u : ipb_div_up(n, 256 * 8)), dim3(256), 0, stream,
               a, b, out, (long long)n, unsharp, amount);
    return ipb_check_launch("ipb_k_gauss_combine");
}

int ipb_graymorph_u16(const uint16_t* in, uint16_t* tmp, uint16_t* out, int n_images, int H, int W, int radius, int dilate,
                      void* stream)
{
    if (n_images <= 0) return IPB_OK;
    IPB_REQUIRE(in && tmp && out && H > 0 && W > 0, "ipb_graymorph_u16: bad argument");
    IPB_REQUIRE(radius >= 0 && radius <= IPB_GM_MAXR, "ipb_graymorph_u16: radius %d not in 0..%d", radius, IPB_GM_MAXR);
    IPB_REQUIRE(n_images <= 65535, "ipb_graymorph_u16: n_images %d out of range", n_images);
    return dilate ? ipb_launch_graymorph<true>(in, tmp, out, n_images, H, W, radius, stream)
                  : ipb_launch_graymorph<false>(in, tmp, out, n_images, H, W, radius, stream);
}

int ipb_sub_u16(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n, void* stream)
{
    if (n <= 0) return IPB_OK;
    IPB_REQUIRE(a && b && out, "ipb_sub_u16: null pointer");
    IPB_LAUNCH(ipb_k_sub_u16, dim3(ipb_div_up(n, 256 * 8) > 4736u ? 4736u : ipb_div_up(n, 256 * 8)), dim3(256), 0, stream, a, b, out,
               (long long)n);
    return ipb_check_launch("ipb_k_sub_u16");
}

int ipb_convert_planes(const void* in, void* out, int64_t n, int to_u16, void* stream)
{
    if (n <= 0) return IPB_OK;
    IPB_REQUIRE(in && out, "ipb_convert_planes: null pointer");
    const unsigned g = ipb_div_up(n, 256 * 8) > 4736u ? 4736u : ipb_div_up(n, 256 * 8);
    if (to_u16) IPB_LAUNCH(ipb_k_f32_to_u16, dim3(g), dim3(256), 0, stream, (const float*)in, (unsigned short*)out, (long long)n);
    else IPB_LAUNCH(ipb_k_u16_to_f32, dim3(g), dim3(256), 0, stream, (const unsigned short*)in, (float*)out, (long long)n);
    return ipb_check_launch("ipb_convert_planes");
}

int ipb_selftest_fdiv(const float* a, const float* b, int64_t n, uint32_t* mismatches, void* stream)
{
    IPB_REQUIRE(a && b && mismatches && n >= 0, "ipb_selftest_fdiv: bad argument");
    if (n == 0) return IPB_OK;
    IPB_LAUNCH(ipb_k_selftest_fdiv, dim3(592), dim3(256), 0, stream, a, b, (long long)n, mismatches);
    return ipb_check_launch("ipb_k_selftest_fdiv");
}

// ---------------------------------------------------------------- workspace sizes (host-side, no device work)
int ipb_hist_sizes(int n_jobs, int frame_h, int has_masked_stride, int64_t* bytes)
{
    IPB_REQUIRE(bytes && n_jobs >= 0 && frame_h >= 0, "ipb_hist_sizes: bad argument");
    const int64_t n = n_jobs > 0 ? n_jobs : 1;
    bytes[0] = n * 65536 * (int64_t)sizeof(uint32_t);                 // hist
    bytes[1] = n * 4 * (int64_t)sizeof(uint64_t);                     // stats
    bytes[2] = has_masked_stride ? n * frame_h * (int64_t)sizeof(uint64_t) : 0;   // row_rank_scratch
    return IPB_OK;
}

int ipb_hist_select_sizes(int n_jobs, int n_q, int64_t* bytes)
{
    IPB_REQUIRE(bytes && n_jobs >= 0 && n_q >= 0, "ipb_hist_select_sizes: bad argument");
    const int64_t n = n_jobs > 0 ? n_jobs : 1, q = n_q > 0 ? n_q : 1;
    bytes[0] = n * IPB_PQ_WIN * (int64_t)sizeof(uint32_t);            // hist_win
    bytes[1] = n * (int64_t)sizeof(IpbHistWin);                       // win
    bytes[2] = n * (int64_t)sizeof(uint64_t);                         // cnt
    bytes[3] = n * 4 * (int64_t)sizeof(uint64_t);                     // stats
    bytes[4] = q * (int64_t)sizeof(IpbQOut);                          // qout
    return IPB_OK;
}

int ipb_roi_stats_fused_sizes(int n_regions, int n_jobs, int max_rect_w, int max_rect_h, int n_sms, int64_t* out)
{
    IPB_REQUIRE(out && n_regions >= 0 && n_jobs >= 0 && max_rect_w >= 0 && max_rect_h >= 0 && n_sms > 0,
                "ipb_roi_stats_fused_sizes: bad argument");
    // list slots of a CTA: every in-window key of the largest region even when half of its pixels are
    // in a window, for three sources (half of the slice for the ratio, a quarter per uint16 slot)
    int64_t stride = 2 * ((int64_t)max_rect_w + 16) * max_rect_h + 16384;
    if (stride > (1ll << 19)) stride = 1ll << 19;
    const int n_ctas = 2 * n_sms;                                     // __launch_bounds__(256, 2)
    out[0] = stride;                                                  // stride_words
    out[1] = n_ctas;                                                  // n_ctas
    out[2] = stride * n_ctas * (int64_t)sizeof(uint32_t);             // scratch bytes
    out[3] = 2 * (int64_t)sizeof(uint32_t);                           // counter bytes
    out[4] = n_regions > 0 ? n_regions : 1;                           // flags bytes
    out[5] = n_jobs > 0 ? n_jobs : 1;                                 // wide_flags bytes
    return IPB_OK;
}

int ipb_fa_segment_sizes(int n_crops, const int32_t* crop_wh /* [host] w, h per crop */, int want_labels, int64_t* out)
{
    IPB_REQUIRE(out && n_crops >= 0 && (n_crops == 0 || crop_wh), "ipb_fa_segment_sizes: bad argument");
    int64_t words = 0, px = 0, rows = 0, cap = 0;
    for (int i = 0; i < n_crops; ++i) {
        const int64_t w = crop_wh[2 * i], h = crop_wh[2 * i + 1];
        IPB_REQUIRE(w >= 0 && h >= 0, "ipb_fa_segment_sizes: negative crop size");
        words += ((w + 31) / 32) * h; px += w * h; rows += h;
        cap += ((h + 1) / 2) * ((w + 1) / 2);                         // bound on 8-connected components of an h x w image
    }
    if (words < 1) words = 1;
    if (px < 1) px = 1;
    if (rows < 1) rows = 1;
    if (cap < 1) cap = 1;
    out[0] = words * 4;                                               // bw_a, bw_b, rootbits, bw_final: each
    out[1] = px * 4;                                                  // L, csize: each
    out[2] = rows * 4;                                                // row_roots, row_base: each
    out[3] = (int64_t)(n_crops > 0 ? n_crops : 1) * 4;                // crop_count
    out[4] = ((int64_t)n_crops + 1) * 4;                              // comp_off
    out[5] = cap;                                                     // comp_cap (rows)
    out[6] = cap * (int64_t)sizeof(IpbComp);                          // comps
    out[7] = want_labels ? px * 4 : 0;                                // labels
    out[8] = rows;                                                    // total_rows
    return IPB_OK;
}

int ipb_region_dilate_sizes(int n_regions, const int32_t* region_wh /* [host] w, h per region */, int64_t* out)
{
    IPB_REQUIRE(out && n_regions >= 0 && (n_regions == 0 || region_wh), "ipb_region_dilate_sizes: bad argument");
    int64_t g = 0;
    for (int i = 0; i < n_regions; ++i) g += (int64_t)region_wh[2 * i] * region_wh[2 * i + 1];
    out[0] = g > 0 ? g : 1;                                           // g_scratch bytes (g_off[i] = bytes before region i)
    return IPB_OK;
}

int ipb_sizeof(int what)
{
    switch (what) {
        case 0: return (int)sizeof(IpbHistJob);
        case 1: return (int)sizeof(IpbQJob);
        case 2: return (int)sizeof(IpbQOut);
        case 3: return (int)sizeof(IpbRegion);
        case 4: return (int)sizeof(IpbStatJob);
        case 5: return (int)sizeof(IpbStatOut);
        case 6: return (int)sizeof(IpbFretCfg);
        case 7: return (int)sizeof(IpbCrop);
        case 8: return (int)sizeof(IpbComp);
        case 9: return (int)sizeof(IpbCropJob);
        case 10: return (int)sizeof(IpbPlanePass);
        case 11: return (int)sizeof(IpbHistWin);
        case 12: return (int)sizeof(IpbRoiJob);
        default: return -1;
    }
}

}  // extern "C"
