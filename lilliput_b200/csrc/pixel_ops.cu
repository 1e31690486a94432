// pixel_ops.cu -- orientation, region copy, alpha blend, clear/fill on packed u8 frames.
//
// Replaces: cv::OrientationTransform (ref opencv.cpp:217-221), opencv_copy_to_region
// (ref opencv.cpp:680-752), opencv_copy_to_region_with_alpha (ref opencv.cpp:556-667),
// opencv_mat_clear_to_transparent / reset / set_color (ref opencv.cpp:466-543).
// All are byte moves except the blend, which is fp32 with every OpenCV Mat expression
// rounded on its own (explicit __f*_rn intrinsics: no FMA contraction), RNE to u8.
#include "common.cuh"
#include "kernels.cuh"

namespace lp {

// EXIF orientation: destination pixel (x, y) <- source pixel, golden table SURVEY.md 8a R4.
__global__ void orient_kernel(const uint8_t* __restrict__ src, int w, int h, int C, int o,
                              uint8_t* __restrict__ dst, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y >= H) return;
    int sx, sy;
    switch (o) {
        case 2: sx = w - 1 - x; sy = y; break;
        case 3: sx = w - 1 - x; sy = h - 1 - y; break;
        case 4: sx = x; sy = h - 1 - y; break;
        case 5: sx = y; sy = x; break;
        case 6: sx = y; sy = h - 1 - x; break;
        case 7: sx = w - 1 - y; sy = h - 1 - x; break;
        case 8: sx = w - 1 - y; sy = x; break;
        default: sx = x; sy = y; break;
    }
    const uint8_t* s = src + ((size_t)sy * w + sx) * C;
    uint8_t* d = dst + ((size_t)y * W + x) * C;
    for (int c = 0; c < C; c++) d[c] = s[c];
}

int orient_launch(const uint8_t* src, int w, int h, int C, int o, uint8_t* dst, cudaStream_t st) {
    const bool swap = o >= 5 && o <= 8;
    const int W = swap ? h : w, H = swap ? w : h;
    dim3 grid(ceil_div(W, 128), H);
    orient_kernel<<<grid, 128, 0, st>>>(src, w, h, C, o, dst, W, H);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    return LP_OK;
}

// copyTo with cvtColor channel adaptation: 3->4 (A=255), 4->3 (drop A), 1->3/4 (replicate, A=255).
__global__ void copy_region_kernel(const uint8_t* __restrict__ src, size_t sstep, int sc,
                                   uint8_t* __restrict__ dst, size_t dstep, int dc, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* s = src + (size_t)y * sstep + (size_t)x * sc;
    uint8_t* d = dst + (size_t)y * dstep + (size_t)x * dc;
    if (sc == dc) {
        for (int c = 0; c < dc; c++) d[c] = s[c];
    } else if (sc == 1) {
        d[0] = d[1] = d[2] = s[0];
        if (dc == 4) d[3] = 255;
    } else {
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
        if (dc == 4) d[3] = 255;
    }
}

int copy_region_launch(const uint8_t* src, size_t sstep, int sc, uint8_t* dst, size_t dstep, int dc,
                       int w, int h, cudaStream_t st) {
    dim3 grid(ceil_div(w, 128), h);
    copy_region_kernel<<<grid, 128, 0, st>>>(src, sstep, sc, dst, dstep, dc, w, h);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    return LP_OK;
}

__device__ __forceinline__ uint8_t sat_rne(float f) {
    if (f != f) return 0;  // 0/0 -> NaN -> cvtss2si gives INT_MIN -> saturates to 0
    const int v = __float2int_rn(f);
    return (uint8_t)min(max(v, 0), 255);
}

// "over" compositing exactly as the reference spells it with cv::Mat expressions.
__global__ void blend_region_kernel(const uint8_t* __restrict__ src, size_t sstep, int sc,
                                    uint8_t* __restrict__ dst, size_t dstep, int dc, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* s = src + (size_t)y * sstep + (size_t)x * sc;
    uint8_t* d = dst + (size_t)y * dstep + (size_t)x * dc;
    const float k = (float)(1.0 / 255.0);
    const int g = sc == 1;  // grayscale source is expanded to BGR first
    const float sa = __fmul_rn((float)(sc == 4 ? s[3] : 255), k);
    const float da = __fmul_rn((float)(dc == 4 ? d[3] : 255), k);
    const float oma = __fsub_rn(1.0f, sa);
    const float oa = __fadd_rn(sa, __fmul_rn(da, oma));
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float scf = __fmul_rn((float)s[g ? 0 : c], k), dcf = __fmul_rn((float)d[c], k);
        const float num = __fadd_rn(__fmul_rn(scf, sa), __fmul_rn(__fmul_rn(dcf, da), oma));
        d[c] = sat_rne(__fmul_rn(__fdiv_rn(num, oa), 255.0f));
    }
    if (dc == 4) d[3] = sat_rne(__fmul_rn(oa, 255.0f));
}

int blend_region_launch(const uint8_t* src, size_t sstep, int sc, uint8_t* dst, size_t dstep, int dc,
                        int w, int h, cudaStream_t st) {
    dim3 grid(ceil_div(w, 128), h);
    blend_region_kernel<<<grid, 128, 0, st>>>(src, sstep, sc, dst, dstep, dc, w, h);
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    return LP_OK;
}

__global__ void fill_kernel(uint8_t* dst, size_t step, int C, int w, int h, uchar4 color) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    uint8_t* d = dst + (size_t)y * step + (size_t)x * C;
    const uint8_t v[4] = {color.x, color.y, color.z, color.w};
    for (int c = 0; c < C; c++) d[c] = v[c];
}

int fill_launch(uint8_t* dst, size_t step, int C, int w, int h, int b, int g, int r, int a,
                cudaStream_t st) {
    dim3 grid(ceil_div(w, 128), h);
    fill_kernel<<<grid, 128, 0, st>>>(dst, step, C, w, h, make_uchar4(b, g, r, a));
    g_launches++;
    LP_CUDA_OK(cudaGetLastError());
    return LP_OK;
}

}  // namespace lp
