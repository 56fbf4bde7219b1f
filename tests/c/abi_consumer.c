/* tests/c/abi_consumer.c -- TEST INFRASTRUCTURE.  A plain C99 caller of include/ipb200.h, the way a
 * non-Python binding (cgo / JNI / a C host program) would see the library: the header must compile as C,
 * every declared entry point must link, and the host-side queries (struct sizes, workspace sizes) must
 * work without a GPU.  It launches nothing.  Output: one "name value..." line per query, compared by
 * tests/test_abi.py with what the Python binding gets from the same library. */
#include <stdio.h>
#include "ipb200.h"

typedef void (*fn_t)(void);

int main(void) {
    /* the address of every entry point: an undeclared or unexported one fails to compile / link */
    fn_t all[] = {
        (fn_t)ipb_last_error, (fn_t)ipb_version, (fn_t)ipb_is_emulated, (fn_t)ipb_sizeof,
        (fn_t)ipb_rasterize_rois, (fn_t)ipb_hist_u16, (fn_t)ipb_hist_planes, (fn_t)ipb_hist_select,
        (fn_t)ipb_hist_quantiles, (fn_t)ipb_scatter_qvalues, (fn_t)ipb_fret_eps, (fn_t)ipb_fa_params,
        (fn_t)ipb_fret_pixels, (fn_t)ipb_region_stats, (fn_t)ipb_roi_stats_fused, (fn_t)ipb_fa_segment,
        (fn_t)ipb_fa_contour_cells, (fn_t)ipb_region_dilate, (fn_t)ipb_region_moments, (fn_t)ipb_preview_u16,
        (fn_t)ipb_crop_normalize, (fn_t)ipb_eps_from_stat, (fn_t)ipb_gaussian_f32, (fn_t)ipb_gauss_combine,
        (fn_t)ipb_graymorph_u16, (fn_t)ipb_sub_u16, (fn_t)ipb_convert_planes, (fn_t)ipb_selftest_fdiv,
        (fn_t)ipb_hist_sizes, (fn_t)ipb_hist_select_sizes, (fn_t)ipb_roi_stats_fused_sizes,
        (fn_t)ipb_fa_segment_sizes, (fn_t)ipb_region_dilate_sizes,
    };
    int n = (int)(sizeof all / sizeof all[0]), i;
    for (i = 0; i < n; i++)
        if (!all[i]) return 2;
    printf("entry_points %d\n", n);
    printf("version %d\n", ipb_version());
    printf("emulated %d\n", ipb_is_emulated());
    printf("sizeof");
    for (i = 0; i <= 12; i++) printf(" %d", ipb_sizeof(i));
    printf("\n");
    {
        int64_t b[16] = {0};
        int32_t wh[6] = {141, 135, 1008, 469, 33, 7};      /* three crops / regions, odd widths */
        int rc, k;
        rc = ipb_hist_sizes(5, 2048, 1, b);
        printf("hist_sizes %d %lld %lld %lld\n", rc, (long long)b[0], (long long)b[1], (long long)b[2]);
        rc = ipb_hist_select_sizes(7, 11, b);
        printf("hist_select_sizes %d", rc);
        for (k = 0; k < 5; k++) printf(" %lld", (long long)b[k]);
        printf("\n");
        rc = ipb_roi_stats_fused_sizes(24, 24, 330, 325, 148, b);
        printf("roi_stats_fused_sizes %d", rc);
        for (k = 0; k < 6; k++) printf(" %lld", (long long)b[k]);
        printf("\n");
        rc = ipb_fa_segment_sizes(3, wh, 1, b);
        printf("fa_segment_sizes %d", rc);
        for (k = 0; k < 9; k++) printf(" %lld", (long long)b[k]);
        printf("\n");
        rc = ipb_region_dilate_sizes(3, wh, b);
        printf("region_dilate_sizes %d %lld\n", rc, (long long)b[0]);
        /* argument errors come back as a code + a thread-local message, never as a crash */
        rc = ipb_hist_select_sizes(-1, 0, b);
        printf("bad_arg %d %s\n", rc, rc < 0 && ipb_last_error()[0] ? "message" : "no-message");
    }
    return 0;
}
