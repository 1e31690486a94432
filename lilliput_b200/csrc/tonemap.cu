// tonemap.cu -- HDR (PQ / HLG) pixels of an 8-bit frame -> SDR BT.709, in place, on sm_100a.
//
// Replaces: tonemap_rgb_8u_inplace -> tonemap_rgb_to_sdr (ref color_info.cpp:112-236), which
// Framebuffer.TonemapToSDR (ref opencv.go:791-810) runs right after the decode of a source whose PNG cICP chunk
// (or AVIF colour box) signals a PQ or HLG transfer (ref ops.go:154-165, 511-517):
//   u8 / 255 -> EOTF (ST.2084 PQ or HLG) -> cv::TonemapReinhard(gamma 1.0, intensity 0.6, light_adapt 0.2,
//   color_adapt 0.3) -> 3x3 primaries matrix to BT.709 -> x 255, round to nearest, saturate.
// cv::TonemapReinhard is not a per-pixel map: it normalises the frame to [min, max] twice and uses the log-mean,
// log-min, log-max and the channel / grey means of the normalised frame -- global reductions inside ONE image
// (SURVEY 8(f)3).  Here: four passes over the frame, each a grid-wide reduction (warp shuffles -> one atomic per
// block) or the final map; the EOTF is recomputed in every pass instead of keeping a 12-byte-per-pixel fp32
// copy, so a pass reads 3-4 bytes per pixel and only the last one writes.  The scalar glue between the passes
// runs on the host in the reference's precisions.  The arithmetic is fp32 like OpenCV's; the reference's own
// SIMD evaluation order differs in the last ulp, which shows as +-1 LSB on ~0.01 % of the samples
// (tests/test_gpu_tonemap.py states the tolerance; oracle.tonemap_to_sdr is the restatement pinned on oracle/_ref).
#include <cfloat>
#include <cmath>

#include "common.cuh"
#include "kernels.cuh"

namespace lp {

struct TmScalars {
    int transfer;                 // 16 = PQ, 18 = HLG, anything else = none
    float a1, b1;                 // first normalisation: im = lin * a1 + b1
    float map_key, intensity;     // Reinhard
    float glob[3];                // color_adapt * channel mean + (1 - color_adapt) * grey mean
    float a2, b2;                 // second normalisation
    int use_matrix;
    float m[9];
};
struct TmReduce {
    float mn, mx;                 // pass 1 / pass 3: min, max over all channels
    float lmn, lmx;               // pass 2: min / max of log(max(grey, 1e-4))
    double slog, sgray, sch[3];   // pass 2: sums
};

__device__ __forceinline__ float tm_pq(float x) {
    const float m1 = 0.1593017578125f, m2 = 78.84375f, c1 = 0.8359375f, c2 = 18.8515625f, c3 = 18.6875f;
    const float xp = powf(x, 1.0f / m2);
    const float num = fmaxf(xp - c1, 0.0f), den = c2 - c3 * xp;
    return powf(num / den, 1.0f / m1);
}
__device__ __forceinline__ float tm_hlg(float x) {
    const float a = 0.17883277f, b = 0.28466892f, c = 0.55991073f;
    return x <= 0.5f ? x * x / 3.0f : (expf((x - c) / a) + b) / 12.0f;
}
__device__ __forceinline__ float tm_eotf(int transfer, uint8_t v) {
    const float x = (float)v * (1.0f / 255.0f);
    return transfer == 16 ? tm_pq(x) : transfer == 18 ? tm_hlg(x) : x;
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) > v) {
        const int assumed = old;
        old = atomicCAS(a, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) < v) {
        const int assumed = old;
        old = atomicCAS(a, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ float warp_min(float v) {
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// the pre-normalisation Reinhard output of one pixel (ref: cv::TonemapReinhardImpl::process)
__device__ __forceinline__ void tm_reinhard(const TmScalars& s, const float im[3], float out[3]) {
    const float gray = im[0] * 0.299f + im[1] * 0.587f + im[2] * 0.114f;
    const float ca = 0.3f, la = 0.2f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float adapt = ca * im[i] + (1.0f - ca) * gray;
        adapt = la * adapt + (1.0f - la) * s.glob[i];
        adapt = powf(s.intensity * adapt, s.map_key);
        out[i] = im[i] * (1.0f / (adapt + im[i]));
    }
}

template <int PASS>
__global__ void __launch_bounds__(256) tonemap_kernel(uint8_t* px, size_t step, int channels, int w, int h, TmScalars s, TmReduce* red) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const bool in = x < w && y < h;
    float lin[3] = {0, 0, 0};
    uint8_t* p = px + (size_t)y * step + (size_t)x * channels;
    if (in) {
#pragma unroll
        for (int i = 0; i < 3; i++) lin[i] = tm_eotf(s.transfer, p[i]);
    }
    if constexpr (PASS == 1) {
        float mn = in ? fminf(fminf(lin[0], lin[1]), lin[2]) : FLT_MAX, mx = in ? fmaxf(fmaxf(lin[0], lin[1]), lin[2]) : -FLT_MAX;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) {
            atomic_min_float(&red->mn, mn);
            atomic_max_float(&red->mx, mx);
        }
        return;
    } else {
    float im[3];
#pragma unroll
    for (int i = 0; i < 3; i++) im[i] = lin[i] * s.a1 + s.b1;
    if constexpr (PASS == 2) {
        const float gray = im[0] * 0.299f + im[1] * 0.587f + im[2] * 0.114f;
        const float lg = logf(fmaxf(gray, 1e-4f));
        const float lmn = warp_min(in ? lg : FLT_MAX), lmx = warp_max(in ? lg : -FLT_MAX);
        const double slog = warp_sum(in ? (double)lg : 0.0), sgray = warp_sum(in ? (double)gray : 0.0);
        const double s0 = warp_sum(in ? (double)im[0] : 0.0), s1 = warp_sum(in ? (double)im[1] : 0.0), s2 = warp_sum(in ? (double)im[2] : 0.0);
        if ((threadIdx.x & 31) == 0) {
            atomic_min_float(&red->lmn, lmn);
            atomic_max_float(&red->lmx, lmx);
            atomicAdd(&red->slog, slog);
            atomicAdd(&red->sgray, sgray);
            atomicAdd(&red->sch[0], s0);
            atomicAdd(&red->sch[1], s1);
            atomicAdd(&red->sch[2], s2);
        }
        return;
    } else {
    float out[3] = {0, 0, 0};
    if (in) tm_reinhard(s, im, out);
    if constexpr (PASS == 3) {
        float mn = in ? fminf(fminf(out[0], out[1]), out[2]) : FLT_MAX, mx = in ? fmaxf(fmaxf(out[0], out[1]), out[2]) : -FLT_MAX;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) {
            atomic_min_float(&red->mn, mn);
            atomic_max_float(&red->mx, mx);
        }
        return;
    } else {
    if (!in) return;
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = out[i] * s.a2 + s.b2;
    float c[3] = {t[0], t[1], t[2]};
    if (s.use_matrix) {
#pragma unroll
        for (int j = 0; j < 3; j++) c[j] = s.m[j * 3] * t[0] + s.m[j * 3 + 1] * t[1] + s.m[j * 3 + 2] * t[2];
    }
    if (s.transfer == 8) {  // linear light: display gamma (ref color_info.cpp:226-228); PQ / HLG already carry theirs
#pragma unroll
        for (int i = 0; i < 3; i++) c[i] = powf(c[i], 1.0f / 2.2f);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int v = __float2int_rn(c[i] * 255.0f);
        p[i] = (uint8_t)min(max(v, 0), 255);
    }
    // alpha, when present, is untouched (ref color_info.cpp:262-267)
    }
    }
    }
}

static void normalise_coeffs(float mn, float mx, float* a, float* b) {  // cv::TonemapImpl::process: (src - min) / (max - min)
    if ((double)mx - (double)mn > DBL_EPSILON) {
        const double alpha = 1.0 / ((double)mx - (double)mn);
        *a = (float)alpha;
        *b = (float)(-(double)mn * alpha);
    } else {
        *a = 1.0f;
        *b = 0.0f;
    }
}

int tonemap_to_sdr_launch(uint8_t* d_px, size_t step, int channels, int w, int h, int transfer, int primaries, cudaStream_t st) {
    if (!d_px || w <= 0 || h <= 0 || (channels != 3 && channels != 4)) return LP_OK;  // the reference returns silently
    TmScalars s;
    memset(&s, 0, sizeof(s));
    s.transfer = transfer;
    TmReduce* d_red = nullptr;
    LP_CUDA_OK(cudaMallocAsync(&d_red, sizeof(TmReduce), st));
    TmReduce r;
    dim3 grid(ceil_div(w, 256), h);
    auto reset = [&]() {
        memset(&r, 0, sizeof(r));
        r.mn = r.lmn = FLT_MAX;
        r.mx = r.lmx = -FLT_MAX;
        return cudaMemcpyAsync(d_red, &r, sizeof(r), cudaMemcpyHostToDevice, st);
    };
    auto fetch = [&]() {
        if (cudaMemcpyAsync(&r, d_red, sizeof(r), cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
        return cudaStreamSynchronize(st) == cudaSuccess;
    };
    int rc = LP_OK;
    do {
        if (reset() != cudaSuccess) { rc = LP_ERR_CUDA; break; }
        tonemap_kernel<1><<<grid, 256, 0, st>>>(d_px, step, channels, w, h, s, d_red);
        if (!fetch()) { rc = LP_ERR_CUDA; break; }
        normalise_coeffs(r.mn, r.mx, &s.a1, &s.b1);
        if (reset() != cudaSuccess) { rc = LP_ERR_CUDA; break; }
        tonemap_kernel<2><<<grid, 256, 0, st>>>(d_px, step, channels, w, h, s, d_red);
        if (!fetch()) { rc = LP_ERR_CUDA; break; }
        {
            const double total = (double)w * h;
            const float log_mean = (float)(r.slog / total);
            const double log_min = r.lmn, log_max = r.lmx;
            const float key = (float)((log_max - (double)log_mean) / (log_max - log_min));
            s.map_key = 0.3f + 0.7f * powf(key, 1.4f);
            s.intensity = expf(-0.6f);
            const float gray_mean = (float)(r.sgray / total);
            for (int i = 0; i < 3; i++) s.glob[i] = 0.3f * (float)(r.sch[i] / total) + (1.0f - 0.3f) * gray_mean;
        }
        if (reset() != cudaSuccess) { rc = LP_ERR_CUDA; break; }
        tonemap_kernel<3><<<grid, 256, 0, st>>>(d_px, step, channels, w, h, s, d_red);
        if (!fetch()) { rc = LP_ERR_CUDA; break; }
        normalise_coeffs(r.mn, r.mx, &s.a2, &s.b2);
        // ref color_info.cpp:160-197: primaries -> BT.709 (channel order as stored; unknown primaries pass through)
        static const float bt2020[9] = {1.6605f, -0.5876f, -0.0728f, -0.1246f, 1.1329f, -0.0083f, -0.0182f, -0.1006f, 1.1187f};
        static const float p3[9] = {1.2249f, -0.2247f, -0.0002f, -0.0420f, 1.0419f, 0.0001f, -0.0197f, 0.0754f, 0.9443f};
        static const float bt601[9] = {1.0440f, -0.0440f, 0.0000f, -0.0000f, 1.0000f, 0.0000f, 0.0000f, 0.0000f, 1.0000f};
        static const float xyz[9] = {1.0569715f, -0.2039770f, 0.0556301f, 0.0415551f, 1.8759675f, -0.9692436f, -0.4986108f, -1.5373832f, 3.2409699f};
        const float* m = primaries == 9 ? bt2020 : (primaries == 12 || primaries == 11) ? p3 : primaries == 6 ? bt601 : primaries == 10 ? xyz : nullptr;
        s.use_matrix = m != nullptr;
        if (m) memcpy(s.m, m, sizeof(s.m));
        tonemap_kernel<4><<<grid, 256, 0, st>>>(d_px, step, channels, w, h, s, d_red);
        g_launches += 4;
        if (cudaGetLastError() != cudaSuccess) rc = LP_ERR_CUDA;
    } while (0);
    cudaFreeAsync(d_red, st);
    return rc;
}

}  // namespace lp
