/*
 * oracle_resize.c -- CPU restatement of cv::resize for packed u8 as the reference
 * reaches it.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Reference call sites: opencv_mat_resize (ref opencv.cpp:196-208) on a view from
 * opencv_mat_crop (ref opencv.cpp:210-215), driven by Framebuffer.Fit / ResizeTo
 * (ref opencv.go:326-374, 294-309) with CV_INTER_AREA, and cv::resize(INTER_LINEAR)
 * inside opencv_copy_to_region* (ref opencv.cpp:585, 710).
 *
 * The algorithm is OpenCV 4.11.0's modules/imgproc/src/resize.cpp (not in the
 * reference tree; linked as deps/linux/amd64/lib/libopencv_imgproc.a), restated
 * from its published source and SURVEY.md Appendix E.1 / E.5:
 *   - both scales integer         -> resizeAreaFast_Invoker<uchar,int>
 *   - both scales >= 1            -> ResizeArea_Invoker<uchar,float> (fp32 FMA, RNE)
 *   - otherwise / INTER_LINEAR    -> fixed-point bilinear (11-bit coefficients)
 *   - INTER_CUBIC (exported by ref opencv.cpp:20, never passed by the Go side): sources under 4 px
 *     on an axis -> OpenCV's fixed-point bicubic (HResizeCubic<uchar,int,short> + VResizeCubic, the
 *     first width/16*16 samples of a row through the AVX2 float form); anything larger is taken by
 *     the vendored IPP (ippiResizeCubic_8u, B=0 C=0.75), a binary whose result equals the exact
 *     (double) evaluation of the a=-0.75 kernel to within 1 LSB on ~2e-5 of the samples
 *     (tests/test_oracle_live_reference.py states the bound).
 * Pinned against oracle/_ref in tests/test_oracle_live_reference.py and against
 * tests/golden/golden.npz.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static inline uint8_t sat_u8_from_int(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* saturate_cast<uchar>(float): cvRound (round-half-even) then clamp. */
static inline uint8_t sat_u8_from_float(float f) {
    int v = (int)lrintf(f); /* default rounding mode = RNE, like vcvtps2dq */
    return sat_u8_from_int(v);
}

/* OpenCV computeResizeAreaTab, one axis, expressed per destination index. */
int oracle_area_taps(int ssize, int dsize, int* first, int* count, float* w, int max_taps) {
    double scale = (double)ssize / (double)dsize;
    int used = 0;
    for (int d = 0; d < dsize; d++) {
        double f1 = d * scale;
        double f2 = f1 + scale;
        double cell = scale < (double)ssize - f1 ? scale : (double)ssize - f1;
        int s1 = (int)ceil(f1), s2 = (int)floor(f2);
        if (s2 > ssize - 1) s2 = ssize - 1;
        if (s1 > s2) s1 = s2;
        int k = 0, start = -1;
        float* wd = w + (size_t)d * max_taps;
        if (s1 - f1 > 1e-3) {
            if (k >= max_taps) return -1;
            start = s1 - 1;
            wd[k++] = (float)((s1 - f1) / cell);
        }
        for (int s = s1; s < s2; s++) {
            if (k >= max_taps) return -1;
            if (start < 0) start = s;
            wd[k++] = (float)(1.0 / cell);
        }
        if (f2 - s2 > 1e-3) {
            if (k >= max_taps) return -1;
            if (start < 0) start = s2;
            double t = f2 - s2;
            if (t > 1.0) t = 1.0;
            if (t > cell) t = cell;
            wd[k++] = (float)(t / cell);
        }
        first[d] = start < 0 ? 0 : start;
        count[d] = k;
        if (k > used) used = k;
    }
    return used;
}

/* ResizeArea_Invoker<uchar,float>: horizontal FMA chain per source row into buf,
 * vertical chain `sum = beta*buf` for the first tap then fma(beta, buf, sum). */
static int resize_area_general(const uint8_t* src, size_t sstep, int cn, int sw, int sh,
                               uint8_t* dst, size_t dstep, int dw, int dh) {
    int mx = (int)ceil((double)sw / dw) + 2, my = (int)ceil((double)sh / dh) + 2;
    int *xf = malloc(sizeof(int) * dw), *xc = malloc(sizeof(int) * dw);
    int *yf = malloc(sizeof(int) * dh), *yc = malloc(sizeof(int) * dh);
    float* xw = malloc(sizeof(float) * (size_t)dw * mx);
    float* yw = malloc(sizeof(float) * (size_t)dh * my);
    float* buf = malloc(sizeof(float) * (size_t)dw * cn);
    float* sum = malloc(sizeof(float) * (size_t)dw * cn);
    int rc = 0;
    if (oracle_area_taps(sw, dw, xf, xc, xw, mx) < 0 || oracle_area_taps(sh, dh, yf, yc, yw, my) < 0)
        rc = -2;
    for (int dy = 0; dy < dh && rc == 0; dy++) {
        for (int j = 0; j < yc[dy]; j++) {
            const uint8_t* S = src + (size_t)(yf[dy] + j) * sstep;
            float beta = yw[(size_t)dy * my + j];
            for (int dx = 0; dx < dw; dx++) {
                for (int c = 0; c < cn; c++) {
                    float b = 0.f;
                    for (int k = 0; k < xc[dx]; k++)
                        b = fmaf((float)S[(size_t)(xf[dx] + k) * cn + c], xw[(size_t)dx * mx + k], b);
                    buf[dx * cn + c] = b;
                }
            }
            for (int i = 0; i < dw * cn; i++) sum[i] = j == 0 ? beta * buf[i] : fmaf(beta, buf[i], sum[i]);
        }
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int i = 0; i < dw * cn; i++) D[i] = sat_u8_from_float(sum[i]);
    }
    free(xf); free(xc); free(yf); free(yc); free(xw); free(yw); free(buf); free(sum);
    return rc;
}

/* resizeAreaFast_Invoker<uchar,int,...>: 2x2 -> (a+b+c+d+2)>>2, otherwise
 * saturate_cast<uchar>(int_sum * (1.f/area)) with one fp32 multiply. */
static void resize_area_fast(const uint8_t* src, size_t sstep, int cn, uint8_t* dst, size_t dstep,
                             int dw, int dh, int kx, int ky) {
    float scale = 1.f / (float)(kx * ky);
    for (int dy = 0; dy < dh; dy++) {
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; dx++)
            for (int c = 0; c < cn; c++) {
                int s = 0;
                for (int y = 0; y < ky; y++) {
                    const uint8_t* S = src + (size_t)(dy * ky + y) * sstep + (size_t)dx * kx * cn + c;
                    for (int x = 0; x < kx; x++) s += S[x * cn];
                }
                if (kx == 2 && ky == 2)
                    D[dx * cn + c] = (uint8_t)((s + 2) >> 2);
                else
                    D[dx * cn + c] = sat_u8_from_float((float)s * scale);
            }
    }
}

/* Fixed-point bilinear (HResizeLinear<uchar,int,short> + VResizeLinear with
 * FixedPtCast<int,uchar,22>), with plain or area-mode coefficients (E.5). */
static void resize_linear(const uint8_t* src, size_t sstep, int cn, int sw, int sh, uint8_t* dst,
                          size_t dstep, int dw, int dh, int area_mode) {
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
    int* xo = malloc(sizeof(int) * dw);
    short* xa = malloc(sizeof(short) * 2 * dw);
    int* yo = malloc(sizeof(int) * dh);
    short* yb = malloc(sizeof(short) * 2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        int sx;
        float fx;
        if (!area_mode) {
            fx = (float)((dx + 0.5) * scale_x - 0.5);
            sx = (int)floorf(fx);
            fx -= sx;
        } else {
            sx = (int)floor(dx * scale_x);
            fx = (float)((dx + 1) - (sx + 1) * inv_x);
            fx = fx <= 0 ? 0.f : fx - floorf(fx);
        }
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xo[dx] = sx;
        xa[2 * dx] = (short)lrintf((1.f - fx) * 2048.f);
        xa[2 * dx + 1] = (short)lrintf(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        int sy;
        float fy;
        if (!area_mode) {
            fy = (float)((dy + 0.5) * scale_y - 0.5);
            sy = (int)floorf(fy);
            fy -= sy;
        } else {
            sy = (int)floor(dy * scale_y);
            fy = (float)((dy + 1) - (sy + 1) * inv_y);
            fy = fy <= 0 ? 0.f : fy - floorf(fy);
        }
        yo[dy] = sy; /* fy untouched; the two rows are clamped instead */
        yb[2 * dy] = (short)lrintf((1.f - fy) * 2048.f);
        yb[2 * dy + 1] = (short)lrintf(fy * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        int r0 = yo[dy], r1 = yo[dy] + 1;
        r0 = r0 < 0 ? 0 : r0 > sh - 1 ? sh - 1 : r0;
        r1 = r1 < 0 ? 0 : r1 > sh - 1 ? sh - 1 : r1;
        const uint8_t* S0 = src + (size_t)r0 * sstep;
        const uint8_t* S1 = src + (size_t)r1 * sstep;
        int b0 = yb[2 * dy], b1 = yb[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xo[dx], a0 = xa[2 * dx], a1 = xa[2 * dx + 1];
            int sx1 = sx + 1 < sw ? sx + 1 : sx; /* a1 == 0 there */
            for (int c = 0; c < cn; c++) {
                int t0 = S0[sx * cn + c] * a0 + S0[sx1 * cn + c] * a1;
                int t1 = S1[sx * cn + c] * a0 + S1[sx1 * cn + c] * a1;
                int v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
                D[dx * cn + c] = sat_u8_from_int(v);
            }
        }
    }
    free(xo); free(xa); free(yo); free(yb);
}


/* ---- INTER_CUBIC --------------------------------------------------------------------------- */

static void cubic_coeffs_f(float x, float* c) {
    /* interpolateCubic, A = -0.75, fp32 -- with the multiply-adds fused the way the vendored build
     * (gcc, x86-64-v3, -ffp-contract=fast) compiled it; found by comparing against oracle/_ref */
    const float A = -0.75f;
    const float t = x + 1.f, u = 1.f - x;
    c[0] = fmaf(fmaf(fmaf(A, t, -5 * A), t, 8 * A), t, -4 * A);
    c[1] = fmaf(fmaf(A + 2, x, -(A + 3)) * x, x, 1.f);
    c[2] = fmaf(fmaf(A + 2, u, -(A + 3)) * u, u, 1.f);
    c[3] = 1.f - c[0] - c[1] - c[2];
}

static void cubic_coeffs_d(double x, double* c) {
    const double A = -0.75;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.0 - c[0] - c[1] - c[2];
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* OpenCV's own bicubic (resizeGeneric_ with HResizeCubic<uchar,int,short>, VResizeCubic<...,
 * FixedPtCast<int,uchar,22>, VResizeCubicVec_32s8u>): the path cv::resize takes when IPP declines
 * (a source under 4 px on an axis). */
static void resize_cubic_fixed(const uint8_t* src, size_t sstep, int cn, int sw, int sh, uint8_t* dst,
                               size_t dstep, int dw, int dh) {
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    int* xo = malloc(sizeof(int) * dw);
    short* xa = malloc(sizeof(short) * 4 * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5), c[4];
        int sx = (int)floorf(fx);
        fx -= sx;
        cubic_coeffs_f(fx, c);
        xo[dx] = sx;
        for (int k = 0; k < 4; k++) xa[4 * dx + k] = (short)clampi((int)lrintf(c[k] * 2048.f), -32768, 32767);
    }
    int* H = malloc(sizeof(int) * (size_t)sh * dw * cn); /* horizontal pass of every source row */
    for (int y = 0; y < sh; y++) {
        const uint8_t* S = src + (size_t)y * sstep;
        for (int dx = 0; dx < dw; dx++)
            for (int c = 0; c < cn; c++) {
                int v = 0;
                for (int j = 0; j < 4; j++) v += S[clampi(xo[dx] - 1 + j, 0, sw - 1) * cn + c] * xa[4 * dx + j];
                H[((size_t)y * dw + dx) * cn + c] = v;
            }
    }
    const int W = dw * cn, nvec = W / 16 * 16;
    const float scale = 1.f / (2048.f * 2048.f);
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5), c[4];
        int sy = (int)floorf(fy);
        fy -= sy;
        cubic_coeffs_f(fy, c);
        short b[4];
        const int* R[4];
        for (int k = 0; k < 4; k++) {
            b[k] = (short)clampi((int)lrintf(c[k] * 2048.f), -32768, 32767);
            R[k] = H + (size_t)clampi(sy - 1 + k, 0, sh - 1) * W;
        }
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int x = 0; x < W; x++) {
            if (x < nvec) { /* v_muladd chain in fp32, v_round, saturating pack */
                float a = (float)R[3][x] * (b[3] * scale);
                a = fmaf((float)R[2][x], b[2] * scale, a);
                a = fmaf((float)R[1][x], b[1] * scale, a);
                a = fmaf((float)R[0][x], b[0] * scale, a);
                D[x] = sat_u8_from_float(a);
            } else {
                int v = R[0][x] * b[0] + R[1][x] * b[1] + R[2][x] * b[2] + R[3][x] * b[3];
                D[x] = sat_u8_from_int((v + (1 << 21)) >> 22);
            }
        }
    }
    free(xo); free(xa); free(H);
}

/* The a = -0.75 bicubic evaluated in double, rows first then columns, replicated borders, RNE:
 * what the vendored IPP returns up to the bound stated in the header. */
static void resize_cubic_exact(const uint8_t* src, size_t sstep, int cn, int sw, int sh, uint8_t* dst,
                               size_t dstep, int dw, int dh) {
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    int* xo = malloc(sizeof(int) * dw);
    double* xa = malloc(sizeof(double) * 4 * dw);
    for (int dx = 0; dx < dw; dx++) {
        double fx = (dx + 0.5) * scale_x - 0.5;
        int sx = (int)floor(fx);
        xo[dx] = sx;
        cubic_coeffs_d(fx - sx, xa + 4 * dx);
    }
    for (int dy = 0; dy < dh; dy++) {
        double fy = (dy + 0.5) * scale_y - 0.5, b[4];
        int sy = (int)floor(fy);
        cubic_coeffs_d(fy - sy, b);
        const uint8_t* R[4];
        for (int k = 0; k < 4; k++) R[k] = src + (size_t)clampi(sy - 1 + k, 0, sh - 1) * sstep;
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; dx++)
            for (int c = 0; c < cn; c++) {
                double sum = 0.0;
                for (int k = 0; k < 4; k++) {
                    double h = 0.0;
                    for (int j = 0; j < 4; j++)
                        h = fma((double)R[k][clampi(xo[dx] - 1 + j, 0, sw - 1) * cn + c], xa[4 * dx + j], h);
                    sum = fma(h, b[k], sum);
                }
                double r = rint(sum);
                D[dx * cn + c] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
            }
    }
    free(xo); free(xa);
}

int oracle_resize(const uint8_t* src, size_t sstep, int cn, int cx, int cy, int sw, int sh,
                  uint8_t* dst, size_t dstep, int dw, int dh, int interpolation) {
    if (interpolation != ORACLE_INTER_AREA && interpolation != ORACLE_INTER_LINEAR && interpolation != ORACLE_INTER_CUBIC) return -1;
    if (sw < 1 || sh < 1 || dw < 1 || dh < 1) return -1;
    const uint8_t* s0 = src + (size_t)cy * sstep + (size_t)cx * cn;
    if (sw == dw && sh == dh) { /* cv::resize: same size is a plain copy */
        for (int y = 0; y < dh; y++) memcpy(dst + (size_t)y * dstep, s0 + (size_t)y * sstep, (size_t)dw * cn);
        return 0;
    }
    if (interpolation == ORACLE_INTER_CUBIC) {
        if (sw < 4 || sh < 4)
            resize_cubic_fixed(s0, sstep, cn, sw, sh, dst, dstep, dw, dh);
        else
            resize_cubic_exact(s0, sstep, cn, sw, sh, dst, dstep, dw, dh);
        return 0;
    }
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
    int ix = (int)lrint(scale_x), iy = (int)lrint(scale_y); /* saturate_cast<int>(double) */
    int is_area_fast = fabs(scale_x - ix) < DBL_EPSILON && fabs(scale_y - iy) < DBL_EPSILON;
    if (interpolation == ORACLE_INTER_LINEAR && is_area_fast && ix == 2 && iy == 2)
        interpolation = ORACLE_INTER_AREA;
    if (interpolation == ORACLE_INTER_AREA && scale_x >= 1 && scale_y >= 1) {
        if (is_area_fast) {
            resize_area_fast(s0, sstep, cn, dst, dstep, dw, dh, ix, iy);
            return 0;
        }
        return resize_area_general(s0, sstep, cn, sw, sh, dst, dstep, dw, dh);
    }
    resize_linear(s0, sstep, cn, sw, sh, dst, dstep, dw, dh, interpolation == ORACLE_INTER_AREA);
    return 0;
}
