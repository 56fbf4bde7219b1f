/*
 * ipb200.h -- C ABI of libipb200.so: the B200-native (sm_100a) per-pixel analysis hot path of
 * gavyek/ImageProcess.
 *
 * The reference has no FFI of its own: its hot path is a set of module-level Python functions
 * (numpy / scipy / scikit-image / matplotlib.path calls).  This header is the boundary a
 * maintainer binds instead of those library calls; INTEGRATION.md shows the ctypes stub for
 * each reference function.  Reference citations below are relative to /root/reference/src/.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ or torch types.
 *   - Every pointer marked [dev] is a caller-owned DEVICE pointer; [host] is host memory.
 *     The library never allocates, frees or keeps pointers.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - Returns IPB_OK (0) or a negative IPB_ERR_* code; ipb_last_error() gives the
 *     thread-local message.  No global mutable state: safe from many host threads and from
 *     one process per GPU.
 *   - Struct layouts are fixed below and can be checked with ipb_sizeof().
 */
#ifndef IPB200_H
#define IPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPB_OK 0
#define IPB_ERR_ARG (-1)
#define IPB_ERR_CUDA (-2)
#define IPB_ERR_WORKSPACE (-3)
#define IPB_ERR_UNSUPPORTED (-4)

const char* ipb_last_error(void);
int ipb_version(void);
int ipb_is_emulated(void);      /* 1 only in the CPU test build of the same sources */
int ipb_sizeof(int which);      /* 0 HistJob 1 QJob 2 QOut 3 Region 4 StatJob 5 StatOut 6 FretCfg 7 Crop 8 Comp 9 CropJob 10 PlanePass 11 HistWin 12 RoiJob */

/* ------------------------------------------------------------------ ROI rasterisation
 * Replaces rasterize_polygon (INT/Fluor_INT.py:398-403; copies FRET/fret_ratio_builder.py:292,
 * FRET/Nesprin2_FRET_Builder.py:388, MOR_by_ROI.py:160, roi_channel_cropper.py:286), i.e.
 * matplotlib.path.Path.contains_points over the whole grid  (rule IPB_RULE_MPL), and
 * skimage.draw.polygon at INT/FA_Analyzer.py:805,883,1014,1225     (rule IPB_RULE_SK),
 * for all ROIs of a batch of frames in one launch.  Bit-exact with both libraries.
 *
 * Per ROI r (all tables [dev]):
 *   verts_xy[2*vert_off[r] .. 2*vert_off[r+1])  float64 (x,y) in the ROI's LOCAL grid
 *   erect[4r..]  (x0,y0,x1,y1) half-open rect where the rule is evaluated
 *   srect[4r..]  (x0,y0,x1,y1) half-open rect the stored mask covers (contains erect)
 *   org[2r..]    frame position of local (0,0);  roi_frame[r] frame index
 *   mask_off[r]  offset (32-bit words) of the ROI's mask in mask_pool; rows of
 *                ceil((sx1-sx0)/32) words; bit b of word j <-> local x = sx0 + 32 j + b
 * Outputs: mask_pool, area[r] (pixel count), union_bits (optional; must be zeroed by the
 * caller): per frame H rows of union_wpr words, bit b of word j <-> frame x = 32 j + b.    */
#define IPB_RULE_MPL 0
#define IPB_RULE_SK 1
int ipb_rasterize_rois(int rule, int n_rois, const double* verts_xy, const int32_t* vert_off,
                       const int32_t* erect, const int32_t* srect, const int32_t* org,
                       const int32_t* roi_frame, const int64_t* mask_off, int max_rows, int max_wpr,
                       uint32_t* mask_pool, uint32_t* area, uint32_t* union_bits, int union_wpr,
                       int frame_h, void* stream);

/* ------------------------------------------------------------------ histograms / percentiles
 * Replaces every np.percentile / np.histogram over pixels of a uint16-derived float32 image:
 * bg_value (INT/Fluor_INT.py:464-485, FRET/fret_ratio_builder.py:314-330,
 * FRET/Nesprin2_FRET_Builder.py:432-451), FA global stats (INT/FA_Analyzer.py:984-987),
 * pick_epsilon (fret_ratio_builder.py:338-340).  One exact 65536-bin histogram per job.     */
#define IPB_PAT_FULL 0           /* all pixels                        vals = img.ravel()      */
#define IPB_PAT_STRIDE1D 1       /* flat index % k == 0               vals[::k]               */
#define IPB_PAT_STRIDE2D 2       /* y % k == 0 and x % k == 0         img[::k, ::k]           */
#define IPB_PAT_MASKED 3         /* pixels under union_bits           img[scope_mask]         */
#define IPB_PAT_MASKED_STRIDE 4  /* every k-th masked pixel           img[scope_mask][::k]    */
typedef struct {
    int32_t plane;        /* plane index into planes[][H][W] (frame * C + channel) */
    int32_t pattern;      /* IPB_PAT_* */
    int32_t k;            /* stride of the strided patterns, >= 2 (a stride of 1 is IPB_PAT_FULL / IPB_PAT_MASKED) */
    int32_t mask_frame;   /* frame index into union_bits (masked patterns) */
    int32_t moments;      /* != 0: also accumulate sum / sum of squares of ALL pixels */
    int32_t excl_plane1;  /* 1 + index of a second plane for the saturation filter, 0 = none */
    int32_t sat_min;      /* > 0: drop pixels whose value (or the second plane's) is >= sat_min:
                             Nesprin2 saturation filter (Nesprin2_FRET_Builder.py:1415-1421) */
    int32_t pad;
} ipb_hist_job;
/* hist: [n_jobs][65536] uint32;  stats: [n_jobs][4] uint64 = {n_selected, sum_all, sumsq_all, 0}
 * row_rank_scratch: [n_jobs][H] uint64, only for IPB_PAT_MASKED_STRIDE jobs.               */
int ipb_hist_u16(const uint16_t* planes, int H, int W, const void* jobs /* ipb_hist_job[] dev */,
                 int n_jobs, int has_masked_stride, const uint32_t* union_bits, int union_wpr,
                 uint64_t* row_rank_scratch, uint32_t* hist, uint64_t* stats, void* stream);

/* Same histograms and stats as ipb_hist_u16, from ONE read of each plane for all the jobs that
 * sample it (passes: ipb_plane_pass[], up to 4 jobs sharing a plane and saturation partner).   */
int ipb_hist_planes(const uint16_t* planes, int H, int W, const void* jobs, int n_jobs,
                    const void* passes, int n_passes, int has_masked_stride, const uint32_t* union_bits,
                    int union_wpr, uint64_t* row_rank_scratch, uint32_t* hist, uint64_t* stats, void* stream);

/* Percentiles by sampled windows: the same percentiles as ipb_hist_planes + ipb_hist_quantiles,
 * exact, without the full histograms (a shared-memory atomic per pixel bounds those, not HBM).
 * Per plane pass, a stratified sample of 8-pixel units fixes ONE value window [wlo, whi) that holds
 * every wanted rank with overwhelming probability; one read of the plane counts the pixels below
 * the window in registers and histograms the pixels inside it (2048 bins); the plane's integer
 * moments and a sparse [::k, ::k] job (own window from the same sample) ride along.  Served roles per pass: one FULL job, one flat-stride job vals[::k] (k in {2, 4, 8}),
 * one [::k, ::k] job; W % 8 == 0 and W >= 16.  *miss (zeroed by the caller) counts quantiles whose rank fell
 * outside the window or whose pass cannot be served: the caller must then repeat with
 * ipb_hist_planes + ipb_hist_quantiles.  Replaces the np.percentile calls of bg_value
 * (Fluor_INT.py:464-485, fret_ratio_builder.py:314-330), pick_epsilon (fret_ratio_builder.py:338)
 * and the FA global statistics (FA_Analyzer.py:984-987).
 * Buffers: hist_win uint32 [n_jobs][2048], win ipb_hist_win[n_jobs], cnt uint64 [n_jobs],
 * stats uint64 [n_jobs][4]
 * ({n, sum, sumsq, 0} as in ipb_hist_u16).  No buffer needs clearing by the caller.            */
typedef struct { int32_t plane, excl_plane1, sat_min, n_jobs; int32_t job[4]; } ipb_plane_pass;
typedef struct { int32_t wlo, whi, mode, pad; } ipb_hist_win;
int ipb_hist_select(const uint16_t* planes, int H, int W, const void* jobs, int n_jobs,
                    const void* passes /* ipb_plane_pass[] dev */, int n_passes,
                    const void* qjobs /* ipb_q_job[] dev */, int n_q,
                    uint32_t* hist_win, void* win, uint64_t* cnt, uint64_t* stats,
                    void* qout /* ipb_q_out[n_q] */, uint32_t* miss, void* stream);

typedef struct {
    int32_t hist;         /* histogram (job) index */
    float q32;            /* float32(p) / float32(100), as numpy computes it */
    int32_t pad[2];
} ipb_q_job;
typedef struct {
    int32_t prev, next;   /* the two order statistics (raw values); -1 if the sample is empty */
    float gamma;          /* numpy's interpolation weight */
    float value;          /* np.percentile of the float32 copy of the sample */
    uint64_t n;           /* sample size */
} ipb_q_out;
int ipb_hist_quantiles(const uint32_t* hist, const uint64_t* stats, const void* qjobs /* dev */,
                       int n_q, void* qout /* ipb_q_out[] dev */, void* stream);
/* dst[dst_idx[i]] = qout[i].value (0 for an empty sample)                                  */
int ipb_scatter_qvalues(const void* qout, const int32_t* dst_idx, int n, float* dst, void* stream);
/* fparams[f][2] = max(eps_abs, percentile of the bg-corrected denominator)                 */
int ipb_fret_eps(const void* qout_eps, int n_frames, int denom_slot, int clip_neg, float eps_abs,
                 float* fparams, void* stream);
/* fa[f] = {mean, std, bg, mean + alpha*std} in float32 (INT/FA_Analyzer.py:143-144,984-987) */
int ipb_fa_params(const uint64_t* stats, const int32_t* stat_idx, const void* qout_bg, int n_frames,
                  int64_t npx, float alpha, float* fa, void* stream);

/* ------------------------------------------------------------------ fused FRET pass
 * Replaces, per pixel and in one pass: saturation filter (Nesprin2_FRET_Builder.py:1415-1421),
 * bg_correct (fret_ratio_builder.py:332-336), spectral_correct (Nesprin2 460-468), the
 * epsilon-regularised ratio and its inverse (fret_ratio_builder.py:474; Nesprin2 1499-1500),
 * ratio clipping (Nesprin2 1502-1504) and the ROI-masked copy (fret_ratio_builder.py:494-495).
 * planes: uint16 [F][n_ch][H][W]; fparams: float32 [F][4] = {Bd, Ba, eps, Bao};
 * outputs float32 [F][H][W], any of them may be NULL.
 * mom_stats (optional): the integer moments of the donor (mom_acceptor = 0) or acceptor channel ride
 * along for the FA global statistics (INT/FA_Analyzer.py:984-987): sum and sum of squares of frame f
 * are ADDED to mom_stats[mom_idx[f]][1] and [2] (rows {n, sum, sumsq, 0} as ipb_hist_u16 writes
 * them; the caller clears those two words or lets ipb_hist_select do it).                     */
typedef struct {
    int32_t numer_is_acceptor;    /* 1: "FRET/Donor", 0: "Donor/FRET" */
    int32_t clip_neg;
    int32_t sat_on;  float sat_thr;
    int32_t use_spectral;  float alpha, beta, g_factor;
    int32_t clip_on; float clip_max;
    int32_t donor_ch, acc_ch, aonly_ch /* < 0: none */, n_ch;
} ipb_fret_cfg;
int ipb_fret_pixels(const uint16_t* planes, int n_frames, int H, int W,
                    const void* cfg_host /* ipb_fret_cfg, [host] */, const float* fparams,
                    const uint32_t* union_bits, int union_wpr,
                    const int32_t* union_idx /* [F] frame -> union plane, NULL: identity */,
                    float* R, float* Ralt, float* Rroi, float* Dcorr, float* Acorr,
                    uint64_t* mom_stats, const int32_t* mom_idx, int mom_acceptor, void* stream);

/* ------------------------------------------------------------------ per-region statistics
 * Replaces quantify_stats / quantify_per_roi_multi (INT/Fluor_INT.py:494-538),
 * quantify_per_roi (fret_ratio_builder.py:342-362) and the Nesprin2 row statistics
 * (Nesprin2_FRET_Builder.py:1537-1581): n, sum, sum of squared deviations, min, max and up
 * to three exact order-statistic results (np.percentile / np.median float32 arithmetic).
 * A uint16 job may carry up to two VIEWS (background B, clip) of the same pixels: the order
 * statistics of value = float32(raw) - B are selected once on the raw integer keys and each
 * view gets its own output row.                                                            */
#define IPB_SRC_U16 0            /* value = float32(raw) - B, optionally clipped at 0 */
#define IPB_SRC_F32 1            /* value = float32 image pixel, non-finite dropped */
#define IPB_SRC_RATIO 2          /* value = (eff(N-bg_n)+eps)/(eff(D-bg_d)+eps), N = images[plane],
                                    D = images[clip_neg[0]], six float32 parameters {bg_n, bg_d, eps,
                                    clip_neg, clip_on, clip_max} at bvals[bidx[0]]: the per-ROI ratio
                                    after Nesprin2's annulus background (Nesprin2_FRET_Builder.py:1528-1535) */
#define IPB_QKIND_NONE 0
#define IPB_QKIND_PCT 1
#define IPB_QKIND_MEDIAN 2
typedef struct {
    int64_t mask_off;             /* word offset of the region's bit rows in mask_pool */
    int32_t x0, y0, w, h;         /* rect in frame coordinates */
    int32_t wpr, frame, use_and;
    int32_t and_plane;            /* use_and != 0: AND with bit plane and_bits[and_plane][H][and_wpr] */
} ipb_region;
typedef struct {
    int32_t region, src, plane;
    int32_t n_views;              /* uint16: 1..2, float32: 1 */
    int32_t bidx[2];              /* per view: index into bvals, < 0: B = 0 */
    int32_t clip_neg[2];
    int32_t qkind[3];
    float q32[3];
    int32_t out[2];               /* view v writes output row out[v] */
} ipb_stat_job;
typedef struct {
    uint64_t n, area;
    double sum, ssd;
    float vmin, vmax;
    float q[3];
    float pad0;
} ipb_stat_out;
/* uniform_src: IPB_SRC_* when every job has that source (one launch), -1 for
 * a mixed list (one launch per source, each skipping the other's jobs).                    */
int ipb_region_stats(const void* regions, const void* jobs, int n_jobs, int uniform_src,
                     const uint32_t* mask_pool, const uint32_t* and_bits, int and_wpr, int H, int W,
                     const uint16_t* planes, const float* images, const float* bvals, void* out,
                     const uint8_t* only /* NULL, or [n_regions]: measure only regions with only[region] != 0 */,
                     void* stream);

/* Per-ROI statistics of the FRET + intensity stages in ONE walk of each ROI: up to two uint16
 * channels of a region (each with up to two (B, clip) views) and the epsilon-regularised ratio of the
 * two, recomputed per pixel with the arithmetic of ipb_fret_pixels' plain configuration
 * (fret_ratio_builder.py:466-474), so the ratio image is not read back and the mask is decoded once.
 * Order statistics by sampled value windows: a sample of the ROI fixes, per source and wanted
 * quantile, a window of values; pixels outside the windows only feed integer moments and "below
 * the window" counters, pixels inside go to a fine histogram.  Output rows are ipb_stat_out rows,
 * bit-identical in n / area / min / max / q to ipb_region_stats on the equivalent jobs.
 * Exact in every case: a region the scheme cannot serve (AND plane, tiny or huge region, windows
 * wider than the fine histogram, a rank outside its window) gets flags[region] = 1 and no output;
 * the caller then runs ipb_region_stats on the equivalent job list with only = flags.
 * Replaces quantify_stats / quantify_per_roi_multi (INT/Fluor_INT.py:494-538) and quantify_per_roi
 * (FRET/fret_ratio_builder.py:342-362).
 *   scratch   uint32 [n_ctas][stride_words]: per-CTA lists of in-window keys; ipb_roi_stats_fused_sizes()
 *             gives a sufficient stride, the CTA count and every buffer size
 *   counter   uint32 [2] job counters, flags uint8 [n_regions] (0, or why the region was left to
 *             ipb_region_stats), wide_flags uint8 [n_jobs] (jobs with a very broad uint16 distribution,
 *             handed from the first launch to the second): all cleared by the call
 *   n_ctas    persistent CTAs (two per SM)                                                         */
typedef struct {
    int32_t region;
    int32_t plane[2];             /* uint16 plane of channel slot 0 / 1; < 0: slot unused */
    int32_t n_views[2];
    int32_t bidx[2][2];           /* view's background at bvals[bidx]; < 0: B = 0 */
    int32_t clip[2][2];
    int32_t out[2][2];            /* output row of the view */
    int32_t qkind[2][3];
    float q32[2][3];
    int32_t ratio_on, ratio_out;
    int32_t fp_idx;               /* bvals[fp_idx + {0,1,2}] = {B of slot 0, B of slot 1, eps} */
    int32_t numer_slot;           /* slot of the numerator */
    int32_t ratio_clip_neg;
    int32_t rqkind[3];
    float rq32[3];
} ipb_roi_job;
int ipb_roi_stats_fused(const void* regions, int n_regions, const void* jobs /* ipb_roi_job[] dev */, int n_jobs,
                        const uint32_t* mask_pool, int H, int W, const uint16_t* planes, const float* bvals,
                        void* out /* ipb_stat_out[] */, uint32_t* scratch, int64_t stride_words, int n_ctas,
                        uint32_t* counter, uint8_t* flags, uint8_t* wide_flags, void* stream);

/* ------------------------------------------------------------------ focal-adhesion chain
 * Replaces analyze_fa_crop (INT/FA_Analyzer.py:123-195) for a ragged batch of crops in one
 * call: bw = (crop > thr) & mask (146-147); remove_small_objects, 4-connectivity, float
 * min_size (151); binary_closing with disk(close_radius) (155-156); label, 8-connectivity,
 * raster-order numbering (158); regionprops area / intensity sum / centroid sums (159-186).
 * Crop-border semantics and per-crop label numbering are the reference's.
 *   crops        ipb_crop[n_crops] [dev]; bit rows of a crop live at bit_off in every bit pool
 *   fa_params    float32 [F][4] = {mean, std, bg, thr} from ipb_fa_params
 *   roi_mask     the IPB_RULE_SK mask pool (rows of wpr words at ipb_crop.mask_off)
 *   min_size     <= 0: no small-object removal;  close_radius 0..5 (0: no closing)
 *   bw_a, bw_b, rootbits, bw_final  uint32 pools of the mask pool's size (scratch / result)
 *   L, csize     int32 / uint32 [total_px] scratch;  row_roots, row_base int32 [total_rows]
 *   crop_count   int32 [n_crops];  comp_off int32 [n_crops + 1] (exclusive scan, result)
 *   comps        ipb_comp[comp_cap] result table: components of crop c are rows
 *                comp_off[c] .. comp_off[c+1], label k <-> row comp_off[c] + k - 1
 *   labels       optional int32 [total_px] label maps (crop-local, 0 background)
 *   path         0: many small crops (cell ROIs) run one CTA per crop through the whole chain,
 *                few / huge crops (stitched mosaic) one kernel per phase over the whole chip   */
typedef struct {
    int64_t bit_off, pix_off, row_off;
    int64_t mask_off;     /* word offset of the crop's ROI mask rows in roi_mask (crops of
                             different frames may share one rasterised mask) */
    int32_t ox, oy, w, h, wpr, plane, frame, pad0;
} ipb_crop;
typedef struct {
    uint64_t sum_i, sum_y, sum_x;
    uint32_t area;
    int32_t crop;
} ipb_comp;
int ipb_fa_segment(const void* crops, int n_crops, int max_rows, int64_t total_rows,
                   const uint16_t* planes, int H, int W, const float* fa_params,
                   const uint32_t* roi_mask, double min_size, int close_radius,
                   uint32_t* bw_a, uint32_t* bw_b, int32_t* L, uint32_t* csize, uint32_t* rootbits,
                   int32_t* row_roots, int32_t* row_base, int32_t* crop_count,
                   uint32_t* bw_final, int32_t* comp_off, void* comps, int comp_cap,
                   int32_t* labels, int path /* 0 auto, 1 one CTA per crop, 2 one kernel per phase */,
                   const int32_t* crop_order /* [dev] optional: crops by decreasing size */,
                   int label_conn /* 8: skimage.measure.label (the FA chain); 4: scipy.ndimage.label's default */,
                   void* stream);

/* Outlines of every labelled adhesion of every crop in ONE pass over the label maps: replaces the
 * reference's per-adhesion skimage.measure.find_contours(labeled_img == k, 0.5) loop
 * (INT/FA_Analyzer.py:166-170, O(adhesions x crop pixels)).  For each 2 x 2 cell of a crop's label map
 * and each label k cut by it, one record {cell = r0 * w + c0, k << 4 | case} with the marching-squares
 * case (bit 0 ul, 1 ur, 2 ll, 3 lr of the mask label == k; cases 1..14).  The host links a label's
 * cells in raster order into skimage's contour polylines (imageprocess_b200/contours.py).
 *   labels     int32 [total_px]: the label maps ipb_fa_segment wrote (crop c at ipb_crop.pix_off)
 *   rec        uint32 [total_px][2]: record slice of crop c at pix_off (at most w*h records are kept)
 *   rec_count  uint32 [n_crops]: records emitted per crop (cleared by the call)                     */
int ipb_fa_contour_cells(const void* crops, int n_crops, int64_t max_px, const int32_t* labels, void* rec,
                         uint32_t* rec_count, void* stream);

/* ------------------------------------------------------------------ morphology, moments, previews
 * ipb_region_dilate: dilation of region masks (ipb_region layout, all pools share mask_off)
 * by a symmetric row-convex structuring element  gmax[|dx|] = largest |dy| covered at
 * horizontal offset dx  (dx = 0..R, R <= 254; [host] table):
 *   Euclidean ball dx^2+dy^2 <= d2max, source = ~mask (invert = 1), and_pool = mask:
 *     make_inside_rim_mask, 0 < EDT <= rim_px (FRET/Nesprin2_FRET_Builder.py:409-414), exact;
 *   full square gmax[dx] = p: scipy binary_dilation with ones((2p+1,2p+1)), outside = 0
 *     (annulus_mask_from_poly, Nesprin2_FRET_Builder.py:416-427; ring = outer & ~inner via
 *     andnot_pool).
 * g_scratch: uint8, w*h bytes per region at g_off[region] ([dev] int64 offsets).
 * out = dilate(src) & and_pool & ~andnot_pool (NULL pools are skipped).                     */
int ipb_region_dilate(const void* regions, int n_regions, int max_w, int max_h, const uint32_t* in_pool,
                      int invert, const uint8_t* gmax_host, int R, uint8_t* g_scratch, const int64_t* g_off,
                      const uint32_t* and_pool, const uint32_t* andnot_pool, uint32_t* out_pool, void* stream);
/* out[r][6] = {n, sum x, sum y, sum x^2, sum y^2, sum xy} of the set pixels (frame coordinates):
 * morphology_from_polygon / second_moments (MOR_by_ROI.py:193-241).                         */
int ipb_region_moments(const void* regions, int n_regions, const uint32_t* mask_pool, uint64_t* out, void* stream);
/* 16-bit preview (INT/Fluor_INT.py:930-943, FRET/fret_ratio_builder.py:479-483):
 * out = uint16(((clip(img, lo, hi) - lo) / den) * 65535), lohi[k] = {lo, hi, den} per image. */
int ipb_preview_u16(const float* images, int64_t px_per_image, int n_images, const float* lohi,
                    uint16_t* out, void* stream);
/* ROI cropper chain (roi_channel_cropper.py:923-953): clip((crop-lo)/(hi-lo),0,1) * mask,
 * ** inv_gamma, float32 and/or uint16 crops.  params[k] = {lo, hi-lo}.                      */
typedef struct { int32_t plane, x0, y0, w, h, region; int64_t out_off; } ipb_crop_job;
int ipb_crop_normalize(const void* jobs, int n_jobs, int64_t max_px, const uint16_t* planes, int H, int W,
                       const float* params, float inv_gamma, const void* regions, const uint32_t* mask_pool,
                       float* out_norm, uint16_t* out16, void* stream);
/* fparams[f][2] = max(eps_abs, q[0] of float32 stat row row_of_frame[f]) (Nesprin2 pick_epsilon
 * on a spectrally corrected denominator, Nesprin2_FRET_Builder.py:470-476,1484-1486).       */
int ipb_eps_from_stat(const void* stat_out, const int32_t* row_of_frame, int n_frames, float eps_abs,
                      float* fparams, void* stream);

/* ------------------------------------------------------------------ Gaussian filter
 * scipy.ndimage.gaussian_filter(img, sigma) on float32 images, bit for bit: axis 0 then axis 1, mode
 * 'reflect', float64 accumulation in scipy's order, float32 stored after each pass.  Replaces the
 * display filters of the interactive ROI drawer (roi_manual_drawer.py:870-876: band-pass =
 * gaussian(sigma_small) - gaussian(sigma_large); unsharp = im + amount * (im - gaussian(radius))) and
 * serves the optional, default-OFF Gaussian pre-filter stage.  Shared-memory halo tiles are filled by
 * TMA bulk copies (cp.async.bulk + mbarrier) when W % 4 == 0 and the images are 16-byte aligned.
 *   weights  float64 [radius + 1] [dev]: centre weight, then offsets 1..radius; the caller computes them
 *            as scipy does: x = arange(-r, r + 1); w = exp(-0.5 / sigma^2 * x^2); w /= w.sum();
 *            radius = int(4 * sigma + 0.5) (truncate = 4.0), at most 160
 *   tmp      float32 scratch of the input's size; in, tmp, out: [n_images][H][W]
 * ipb_gauss_combine: out = a - b (unsharp = 0) or a + amount * (a - b), float32 ops rounded one by one. */
int ipb_gaussian_f32(const float* in, float* tmp, float* out, int n_images, int H, int W,
                     const double* weights, int radius, void* stream);
int ipb_gauss_combine(const float* a, const float* b, float* out, int64_t n, int unsharp, float amount, void* stream);

/* ------------------------------------------------------------------ grey-scale morphology (optional stage)
 * scipy.ndimage.grey_erosion / grey_dilation(input, size = 2 * radius + 1) on uint16 planes, mode
 * 'reflect', separable min / max on TMA-filled shared-memory halo tiles (W % 8 == 0, else plain loads).
 * white_tophat(input, size) = input - grey_dilation(grey_erosion(input)) = ipb_graymorph_u16 twice +
 * ipb_sub_u16.  The reference has no such stage (north_star names a top-hat filter): optional, OFF by
 * default; the oracle is scipy.ndimage itself.  tmp: scratch of the input's size; radius <= 64.      */
int ipb_graymorph_u16(const uint16_t* in, uint16_t* tmp, uint16_t* out, int n_images, int H, int W, int radius, int dilate,
                      void* stream);
int ipb_sub_u16(const uint16_t* a, const uint16_t* b, uint16_t* out, int64_t n, void* stream);

/* uint16 <-> float32 planes around the optional Gaussian pre-filter of an integer channel:
 * to_u16 = 0: out float32 = float32(in uint16);  to_u16 = 1: out uint16 = clip(rint(in float32), 0, 65535).  */
int ipb_convert_planes(const void* in, void* out, int64_t n, int to_u16, void* stream);

/* Self-test of the branch-free division ipb_roi_stats_fused uses for ratios whose operands are known to
 * lie in [5, 65536 + eps]: *mismatches (zeroed by the caller) += number of pairs (a[i], b[i]) whose
 * quotient differs from the IEEE-rounded one (the reference's numpy division).  Must stay 0.        */
int ipb_selftest_fdiv(const float* a, const float* b, int64_t n, uint32_t* mismatches, void* stream);

/* ------------------------------------------------------------------ workspace sizes
 * Host-side helpers (no device work, no stream): the byte size of every caller-owned buffer whose
 * size depends on the batch, so a binding in any language can allocate from this header alone.
 * The reference has no counterpart (numpy allocates inside every call it makes on this path, e.g. the
 * H*W x 2 float64 point list of rasterize_polygon, INT/Fluor_INT.py:398-402).  Buffers not listed
 * here have the sizes stated at their entry point (one struct per job / region / frame).
 *   ipb_hist_sizes            bytes[3] = {hist, stats, row_rank_scratch}      (ipb_hist_u16 / ipb_hist_planes)
 *   ipb_hist_select_sizes     bytes[5] = {hist_win, win, cnt, stats, qout}    (ipb_hist_select)
 *   ipb_roi_stats_fused_sizes out[6]   = {stride_words, n_ctas, scratch bytes, counter bytes, flags bytes,
 *                                         wide_flags bytes} for regions whose rects are at most max_rect_w x max_rect_h
 *   ipb_fa_segment_sizes      out[9]   = {bw_a = bw_b = rootbits = bw_final bytes (each), L = csize bytes (each),
 *                                         row_roots = row_base bytes (each), crop_count bytes, comp_off bytes,
 *                                         comp_cap (rows), comps bytes, labels bytes (0 unless wanted), total_rows}
 *                             for crops of crop_wh[2i] x crop_wh[2i+1] pixels
 *   ipb_region_dilate_sizes   out[1]   = {g_scratch bytes}                                              */
int ipb_hist_sizes(int n_jobs, int frame_h, int has_masked_stride, int64_t* bytes);
int ipb_hist_select_sizes(int n_jobs, int n_q, int64_t* bytes);
int ipb_roi_stats_fused_sizes(int n_regions, int n_jobs, int max_rect_w, int max_rect_h, int n_sms, int64_t* out);
int ipb_fa_segment_sizes(int n_crops, const int32_t* crop_wh, int want_labels, int64_t* out);
int ipb_region_dilate_sizes(int n_regions, const int32_t* region_wh, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* IPB200_H */
