/*
 * oracle/c/polygon_rules.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the two third-party polygon rasterisation rules the
 * reference's hot path calls.  Neither library's source is under
 * /root/reference (both are unpinned entries of requirements.txt:3,8), so the
 * published algorithms are restated here and pinned against the reference's
 * own shipped artefacts (see tests/test_oracle_golden.py):
 *
 *   ipbo_mpl_points_in_path  matplotlib.path.Path(poly).contains_points(pts)
 *                            as called by rasterize_polygon
 *                            (src/INT/Fluor_INT.py:398-403 and the four copies
 *                            listed in SURVEY.md 8(a) a1).  Algorithm:
 *                            matplotlib src/_path.h point_in_path_impl
 *                            ("crossings-multiply" test, radius 0).
 *                            Pinned by area_px of the 29 ROIs in
 *                            Testsamples/1Flu_Intensity.../fluor_intensity_perROI.csv
 *
 *   ipbo_sk_polygon_mask     skimage.draw.polygon(r, c, shape) as called at
 *                            src/INT/FA_Analyzer.py:805,883,1014,1225 and
 *                            src/roi_manual_drawer.py:1336.  Algorithm:
 *                            skimage/draw/_draw.pyx _polygon +
 *                            skimage/_shared/geometry.pyx point_in_polygon
 *                            (O'Rourke crossing test, vertex/edge inclusive).
 *                            Pinned by roi/mask/S01_mask.tif of both
 *                            intensity experiments.
 *
 * Both functions deliberately keep the cost profile of the originals (every
 * point is tested against every edge) so that timing them is a fair
 * "reference CPU path" baseline.  Build with -ffp-contract=off: the originals
 * are compiled without FMA contraction on x86-64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* matplotlib _path.h point_in_path_impl for a single sub-path with nv >= 3
 * vertices, no codes (MOVETO, LINETO..., STOP), identity transform, radius 0.
 * verts = [x0,y0,x1,y1,...]; pts = [tx0,ty0,tx1,ty1,...] (n points);
 * out[i] = 1 if point i is inside.  Loop order (edges outer, points inner,
 * per-point yflag0 / subpath_flag arrays) follows the original.              */
int ipbo_mpl_points_in_path(const double *verts, int nv, const double *pts, size_t n,
                            uint8_t *out)
{
    memset(out, 0, n);
    if (nv < 3) return 0;                 /* points_in_path: total_vertices() < 3 */
    uint8_t *yflag0 = (uint8_t *)malloc(n ? n : 1);
    uint8_t *sub = (uint8_t *)calloc(n ? n : 1, 1);
    if (!yflag0 || !sub) { free(yflag0); free(sub); return -1; }

    double sx, sy, vtx0, vty0, vtx1, vty1, x, y;
    sx = vtx0 = vtx1 = verts[0];
    sy = vty0 = vty1 = verts[1];
    for (size_t i = 0; i < n; ++i) {
        double ty = pts[2 * i + 1];
        if (isfinite(ty)) { yflag0[i] = (vty0 >= ty); sub[i] = 0; }
    }
    /* inner do-while: one iteration per vertex read (1..nv-1, then STOP) */
    for (int k = 1; k <= nv; ++k) {
        if (k < nv) { x = verts[2 * k]; y = verts[2 * k + 1]; }
        else        { x = sx; y = sy; }             /* STOP: close to start */
        for (size_t i = 0; i < n; ++i) {
            double tx = pts[2 * i], ty = pts[2 * i + 1];
            if (!(isfinite(tx) && isfinite(ty))) continue;
            uint8_t yflag1 = (vty1 >= ty);
            if (yflag0[i] != yflag1) {
                if (((vty1 - ty) * (vtx0 - vtx1) >= (vtx1 - tx) * (vty0 - vty1)) == yflag1)
                    sub[i] ^= 1;
            }
            yflag0[i] = yflag1;
        }
        vtx0 = vtx1; vty0 = vty1;
        vtx1 = x;    vty1 = y;
    }
    /* closing edge after the loop */
    for (size_t i = 0; i < n; ++i) {
        double tx = pts[2 * i], ty = pts[2 * i + 1];
        if (!(isfinite(tx) && isfinite(ty))) continue;
        uint8_t yflag1 = (vty1 >= ty);
        if (yflag0[i] != yflag1) {
            if (((vty1 - ty) * (vtx0 - vtx1) >= (vtx1 - tx) * (vty0 - vty1)) == yflag1)
                sub[i] ^= 1;
        }
        out[i] |= sub[i];
    }
    free(yflag0); free(sub);
    return 0;
}

/* skimage/_shared/geometry.pyx point_in_polygon: 0 outside, 1 inside,
 * 2 vertex, 3 edge.  xp/yp are the polygon's x (column) and y (row). */
static unsigned char sk_point_in_polygon(const double *xp, const double *yp, int nv,
                                         double x, double y)
{
    unsigned int l_cross = 0, r_cross = 0;
    const float eps = 1e-12f;            /* "cdef float eps = 1e-12" */
    double x1 = xp[nv - 1] - x;
    double y1 = yp[nv - 1] - y;
    for (int i = 0; i < nv; ++i) {
        double x0 = xp[i] - x;
        double y0 = yp[i] - y;
        if ((-eps < x0 && x0 < eps) && (-eps < y0 && y0 < eps)) return 2;
        if ((y0 > 0) != (y1 > 0)) {
            if (((x0 * y1 - x1 * y0) / (y1 - y0)) > 0) r_cross += 1;
        }
        if ((y0 < 0) != (y1 < 0)) {
            if (((x0 * y1 - x1 * y0) / (y1 - y0)) < 0) l_cross += 1;
        }
        x1 = x0; y1 = y0;
    }
    if ((r_cross & 1) != (l_cross & 1)) return 3;
    if (r_cross & 1) return 1;
    return 0;
}

/* skimage/draw/_draw.pyx _polygon(r, c, shape): sets out[row*W+col] = 1 for
 * every returned (rr, cc).  Returns the number of pixels set, <0 on error. */
long ipbo_sk_polygon_mask(const double *r, const double *c, int nv, int H, int W, uint8_t *out)
{
    memset(out, 0, (size_t)H * (size_t)W);
    if (nv <= 0 || H <= 0 || W <= 0) return 0;
    double rmin = r[0], rmax = r[0], cmin = c[0], cmax = c[0];
    for (int i = 1; i < nv; ++i) {
        if (r[i] < rmin) rmin = r[i];
        if (r[i] > rmax) rmax = r[i];
        if (c[i] < cmin) cmin = c[i];
        if (c[i] > cmax) cmax = c[i];
    }
    long minr = (long)(rmin > 0 ? rmin : 0);          /* int(max(0, r.min())) */
    long maxr = (long)ceil(rmax);
    long minc = (long)(cmin > 0 ? cmin : 0);
    long maxc = (long)ceil(cmax);
    if (maxr > H - 1) maxr = H - 1;
    if (maxc > W - 1) maxc = W - 1;
    long cnt = 0;
    for (long ri = minr; ri <= maxr; ++ri)
        for (long ci = minc; ci <= maxc; ++ci)
            if (sk_point_in_polygon(c, r, nv, (double)ci, (double)ri)) {
                out[(size_t)ri * W + ci] = 1;
                ++cnt;
            }
    return cnt;
}
