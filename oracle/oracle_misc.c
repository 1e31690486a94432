/*
 * oracle_misc.c -- CPU restatement of the small byte-moving / fp32 per-pixel ops
 * on the hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

/* cv::OrientationTransform (Discord-patched OpenCV 4.11; ref call site
 * opencv.cpp:217-221).  Behaviour pinned by the golden table the reference
 * produces for a 3x2 image:
 *  1 identity, 2 mirror-x, 3 rot180, 4 mirror-y, 5 transpose, 6 rot90 CW,
 *  7 transverse, 8 rot90 CCW.  5..8 swap the dimensions. */
int oracle_orient(const uint8_t* src, int w, int h, int cn, int o, uint8_t* dst, int* ow, int* oh) {
    if (o < 1 || o > 8) o = 1;
    int swap = o >= 5;
    int W = swap ? h : w, H = swap ? w : h;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int sx, sy;
            switch (o) {
                case 1: sx = x; sy = y; break;
                case 2: sx = w - 1 - x; sy = y; break;
                case 3: sx = w - 1 - x; sy = h - 1 - y; break;
                case 4: sx = x; sy = h - 1 - y; break;
                case 5: sx = y; sy = x; break;
                case 6: sx = y; sy = h - 1 - x; break;
                case 7: sx = w - 1 - y; sy = h - 1 - x; break;
                default: sx = w - 1 - y; sy = x; break; /* 8 */
            }
            memcpy(dst + ((size_t)y * W + x) * cn, src + ((size_t)sy * w + sx) * cn, cn);
        }
    *ow = W;
    *oh = H;
    return 0;
}

static inline uint8_t sat_rne(float f) {
    if (isnan(f)) return 0; /* cvtss2si(NaN) = INT_MIN -> saturates to 0 */
    long v = lrintf(f);
    return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}

/* ref opencv.cpp:556-667: every Mat expression is a separate fp32 pass, so each
 * operation rounds on its own (volatile stops the C compiler from contracting). */
int oracle_blend_over(const uint8_t* src, size_t sstep, int sc, uint8_t* dst, size_t dstep, int dc,
                      int w, int h) {
    if ((sc != 3 && sc != 4) || (dc != 3 && dc != 4)) return -1;
    const float k = (float)(1.0 / 255.0);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* s = src + (size_t)y * sstep + (size_t)x * sc;
            uint8_t* d = dst + (size_t)y * dstep + (size_t)x * dc;
            volatile float sa = (float)(sc == 4 ? s[3] : 255) * k;
            volatile float da = (float)(dc == 4 ? d[3] : 255) * k;
            volatile float oma = 1.0f - sa;
            volatile float t = da * oma;
            volatile float oa = sa + t;
            for (int c = 0; c < 3; c++) {
                volatile float scf = (float)s[c] * k, dcf = (float)d[c] * k;
                volatile float a = scf * sa;
                volatile float b = dcf * da;
                volatile float b2 = b * oma;
                volatile float num = a + b2;
                volatile float q = num / oa;
                d[c] = sat_rne(q * 255.0f);
            }
            if (dc == 4) d[3] = sat_rne(oa * 255.0f);
        }
    return 0;
}

/* ref opencv.go:331-363 (float64, int() truncation) */
void oracle_fit_rect(int sw, int sh, int dw, int dh, int* left, int* top, int* wc, int* hc) {
    double ai = (double)sw / (double)sh, ao = (double)dw / (double)dh;
    int w, h;
    if (ai > ao) { w = (int)(ao * (double)sh + 0.5); h = sh; }
    else { h = (int)((double)sw / ao + 0.5); w = sw; }
    if (w < 1) w = 1;
    if (h < 1) h = 1;
    int l = (int)((double)(sw - w) * 0.5), t = (int)((double)(sh - h) * 0.5);
    *left = l < 0 ? 0 : l;
    *top = t < 0 ? 0 : t;
    *wc = w;
    *hc = h;
}

/* ref ops.go:243-255 */
void oracle_expected_size(int ow, int oh, int rw, int rh, int* w, int* h) {
    int m = ow < oh ? ow : oh;
    if (rw == rh && rw > m) { *w = m; *h = m; }
    else if (rw > ow && rh > oh && rw != rh) { *w = ow; *h = oh; }
    else { *w = rw; *h = rh; }
}
