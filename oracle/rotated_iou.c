/* CPU oracle: IoU of two convex quadrilaterals (BEV footprints of 3-D boxes).  TEST INFRASTRUCTURE ONLY.
 *
 * Restates what /root/reference/opencood/utils/common_utils.py:196-218 (`compute_iou`) asks of shapely:
 *     iou = box.intersection(b).area / box.union(b).area
 * for polygons built from the first four corners of each box (`convert_format`, common_utils.py:221-236),
 * as used by `nms_rotated` (/root/reference/opencood/utils/box_utils.py:693-738).
 *
 * PARITY UNPINNED: shapely (GEOS) is a third-party dependency that is absent from /root/reference and from this
 * image (requirements.txt lists `shapely`, no version pin).  For two convex polygons GEOS' overlay returns the convex
 * intersection polygon; this file computes the same region with Sutherland-Hodgman clipping in float64 and its area
 * with the shoelace formula; union area = area(a) + area(b) - area(intersection).  Pinned only by analytic
 * known-answer cases (tests/test_postprocess_cpu.py).
 */
#include <math.h>
#include <stdint.h>

static double poly_area_signed(const double* p, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        s += p[2 * i] * p[2 * j + 1] - p[2 * j] * p[2 * i + 1];
    }
    return 0.5 * s;
}

/* Intersection area of convex polygons a (na vertices) and b (nb), any orientation. */
double oracle_convex_intersection_area(const double* a_in, int na, const double* b_in, int nb) {
    double a[16], b[16], cur[32], nxt[32];
    /* make both counter-clockwise */
    const double sa = poly_area_signed(a_in, na), sb = poly_area_signed(b_in, nb);
    for (int i = 0; i < na; ++i) {
        const int k = sa >= 0 ? i : na - 1 - i;
        a[2 * i] = a_in[2 * k]; a[2 * i + 1] = a_in[2 * k + 1];
    }
    for (int i = 0; i < nb; ++i) {
        const int k = sb >= 0 ? i : nb - 1 - i;
        b[2 * i] = b_in[2 * k]; b[2 * i + 1] = b_in[2 * k + 1];
    }
    int n = na;
    for (int i = 0; i < 2 * na; ++i) cur[i] = a[i];
    for (int e = 0; e < nb && n > 0; ++e) {
        const double x1 = b[2 * e], y1 = b[2 * e + 1];
        const int e2 = (e + 1 == nb) ? 0 : e + 1;
        const double x2 = b[2 * e2], y2 = b[2 * e2 + 1];
        const double ex = x2 - x1, ey = y2 - y1;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const double px = cur[2 * i], py = cur[2 * i + 1], qx = cur[2 * j], qy = cur[2 * j + 1];
            const double dp = ex * (py - y1) - ey * (px - x1);      /* >= 0: inside (left of the edge) */
            const double dq = ex * (qy - y1) - ey * (qx - x1);
            if (dp >= 0) { nxt[2 * m] = px; nxt[2 * m + 1] = py; ++m; }
            if ((dp >= 0) != (dq >= 0)) {
                const double t = dp / (dp - dq);
                nxt[2 * m] = px + t * (qx - px); nxt[2 * m + 1] = py + t * (qy - py); ++m;
            }
        }
        n = m;
        for (int i = 0; i < 2 * n; ++i) cur[i] = nxt[i];
    }
    if (n < 3) return 0.0;
    return fabs(poly_area_signed(cur, n));
}

/* iou[k] = IoU(quad `box`, quad boxes[k]) as float32 (common_utils.py:218 casts to np.float32). */
void oracle_quad_iou_one_to_many(const double* box, const double* boxes, int n, float* iou) {
    const double a0 = fabs(poly_area_signed(box, 4));
    for (int k = 0; k < n; ++k) {
        const double* b = boxes + 8 * k;
        const double a1 = fabs(poly_area_signed(b, 4));
        const double inter = oracle_convex_intersection_area(box, 4, b, 4);
        iou[k] = (float)(inter / (a0 + a1 - inter));
    }
}
