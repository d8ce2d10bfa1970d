/*
 * oracle/cc_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never shipped, never on the product path).
 *
 * Plain-C restatement of the CPU algorithms of the reference's CC stage, used by tests/, smoke() and
 * bench.py's cpu_baseline leg as the checker for the CUDA path.
 *
 * Parity status: the reference ships no tests / golden vectors, so this oracle is
 * pinned by (a) oracle/_ref/accessmath_lib_ref.so = the reference's own accessmath_lib.c compiled
 * unmodified (oracle/Makefile), (b) outputs of the imported Python reference captured in tests/golden/
 * by oracle/gen_golden.py, (c) scipy.ndimage.label itself (the reference's third-party labeler).
 *
 * Citations: R/ = /root/reference/ACCESS2021_release/
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- orc_label4 -----------------------------------------------------------------------------
 * Restates scipy.ndimage.label(content) as called at R/AccessMath/preprocessing/content/labeler.py:126
 * (third-party: SciPy, unpinned by the reference; 1.18.1 in this image): default structuring element
 * = 4-connectivity, foreground = (content != 0), int32 output, labels 1..n numbered in raster order of
 * each component's first pixel. Returns n.
 */
static int32_t uf_find(int32_t *p, int32_t a) {
    while (p[a] != a) { p[a] = p[p[a]]; a = p[a]; }
    return a;
}
static void uf_union(int32_t *p, int32_t a, int32_t b) {
    a = uf_find(p, a); b = uf_find(p, b);
    if (a == b) return;
    if (a < b) p[b] = a; else p[a] = b;          /* root = minimum linear index */
}

int orc_label4(const uint8_t *content, int width, int height, int32_t *labels) {
    int64_t n_px = (int64_t)width * height;
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_px > 0 ? n_px : 1));
    for (int64_t i = 0; i < n_px; i++) parent[i] = (int32_t)i;
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            int32_t idx = y * width + x;
            if (!content[idx]) continue;
            if (x > 0 && content[idx - 1]) uf_union(parent, idx, idx - 1);
            if (y > 0 && content[idx - width]) uf_union(parent, idx, idx - width);
        }
    }
    int32_t next = 0;
    for (int64_t i = 0; i < n_px; i++) {
        if (!content[i]) { labels[i] = 0; continue; }
        int32_t r = uf_find(parent, (int32_t)i);
        if (r == i) labels[i] = ++next;              /* first raster pixel of its component */
        else labels[i] = labels[r];                  /* r < i, already numbered */
    }
    free(parent);
    return next;
}

/* ---- orc_age_boundaries ---------------------------------------------------------------------
 * Restates CC_AgeBoundaries, R/accessmath_lib.c:357-413: per label bbox (inclusive), pixel count and
 * minimum age with the "-1 = unset" rule, one raster pass.
 */
int orc_age_boundaries(const int32_t *labels, const float *ages, int width, int height, int count_labels,
                       int32_t *mins_y, int32_t *maxs_y, int32_t *mins_x, int32_t *maxs_x,
                       int32_t *counts, float *out_age) {
    for (int i = 0; i < count_labels; i++) {        /* accessmath_lib.c:364-374 */
        mins_y[i] = height; maxs_y[i] = 0; mins_x[i] = width; maxs_x[i] = 0; counts[i] = 0; out_age[i] = -1.0f;
    }
    int64_t idx = 0;
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++, idx++) {     /* accessmath_lib.c:378-409 */
            int32_t l = labels[idx];
            if (l <= 0) continue;
            int c = l - 1;
            if (mins_y[c] > y) mins_y[c] = y;
            if (maxs_y[c] < y) maxs_y[c] = y;
            if (mins_x[c] > x) mins_x[c] = x;
            if (maxs_x[c] < x) maxs_x[c] = x;
            counts[c]++;
            if (out_age[c] < 0.0f || ages[idx] < out_age[c]) out_age[c] = ages[idx];
        }
    }
    return 0;
}

/* ---- orc_overlap_count ----------------------------------------------------------------------
 * Restates the integer part of ConnectedComponent.getOverlapFMeasure,
 * R/AM_CommonTools/data/connected_component.py:202-228: number of pixels set in both crops inside the
 * intersection of the two (inclusive) bounding boxes; 0 when the boxes do not intersect (:250).
 * Crops are uint8 (h x w) 0/255 images, row-major, as built at labeler.py:183.
 */
int orc_overlap_count(const uint8_t *img_a, int a_min_x, int a_max_x, int a_min_y, int a_max_y,
                      const uint8_t *img_b, int b_min_x, int b_max_x, int b_min_y, int b_max_y) {
    if (!(a_max_y >= b_min_y && b_max_y >= a_min_y && a_max_x >= b_min_x && b_max_x >= a_min_x)) return 0;
    int x0 = a_min_x > b_min_x ? a_min_x : b_min_x, x1 = a_max_x < b_max_x ? a_max_x : b_max_x;
    int y0 = a_min_y > b_min_y ? a_min_y : b_min_y, y1 = a_max_y < b_max_y ? a_max_y : b_max_y;
    int aw = a_max_x - a_min_x + 1, bw = b_max_x - b_min_x + 1, match = 0;
    for (int y = y0; y <= y1; y++) {
        const uint8_t *ra = img_a + (size_t)(y - a_min_y) * aw - a_min_x;
        const uint8_t *rb = img_b + (size_t)(y - b_min_y) * bw - b_min_x;
        for (int x = x0; x <= x1; x++) match += (ra[x] & rb[x]) != 0;
    }
    return match;
}
