// common.cuh -- shared host/device helpers for liblilliput_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "lilliput_b200.h"

#define LP_CUDA_OK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            fprintf(stderr, "[lilliput_b200] CUDA error %s at %s:%d: %s\n",                \
                    cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e));     \
            return LP_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define LP_CUDA_OK_NULL(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            fprintf(stderr, "[lilliput_b200] CUDA error %s at %s:%d: %s\n",                \
                    cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e));     \
            return nullptr;                                                                \
        }                                                                                  \
    } while (0)

namespace lp {

constexpr int kNumSMs = 148;  // B200

// Process-wide lazy context: verifies a CUDA device exists (no CPU fallback).
// Returns LP_OK or LP_ERR_CUDA (logged once).
int ensure_device();
// Per-host-thread stream used by the synchronous per-image ABI.
cudaStream_t thread_stream();
// Count of kernel launches issued by this library on the calling thread
// (bench.py reports it as gpu_launches).
extern thread_local long g_launches;

// Makes `device` current for the scope and restores the caller's device afterwards (the batch entry points take a
// device ordinal; the per-image ABI on the same host thread must keep seeing the device it was using).
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = prev == device || cudaSetDevice(device) == cudaSuccess;
        if (prev == device) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
static inline T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}
template <class T>
static inline T round_up(T a, T b) {
    return ceil_div(a, b) * b;
}

}  // namespace lp
