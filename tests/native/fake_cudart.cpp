// A pretend CUDA runtime for sanitizer runs of the HOST side of the library on a machine without a GPU:
// "device" memory is host memory (so AddressSanitizer checks every cudaMemcpy* against the allocation it lands
// in), streams and events are inert, kernel launches succeed without running anything.  Outputs are therefore
// garbage; what is exercised is every allocation size, copy size and piece of bookkeeping the host code does.
// Probe infrastructure only (tests/native/host_fake_gpu_fuzz.sh).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {
void** __cudaRegisterFatBinary(void*) { static void* h; return &h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char*, char*, const char*, int, uint3*, uint3*, dim3*, dim3*, int*) {}
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}
unsigned __cudaPushCallConfiguration(dim3, dim3, size_t, cudaStream_t) { return 0; }
cudaError_t __cudaPopCallConfiguration(dim3*, dim3*, size_t*, void*) { return cudaSuccess; }

cudaError_t cudaLaunchKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "fake"; }
const char* cudaGetErrorName(cudaError_t) { return "fake"; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

static cudaError_t alloc(void** p, size_t n) {
    if (n > (size_t)6 << 30) return cudaErrorMemoryAllocation;
    *p = calloc(n ? n : 1, 1);  // zeroed: status words the host reads back say "ok"
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaMalloc(void** p, size_t n) { return alloc(p, n); }
cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return alloc(p, n); }
cudaError_t cudaMallocHost(void** p, size_t n) { return alloc(p, n); }
cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) { *free_b = (size_t)1 << 30; *total_b = (size_t)2 << 30; return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return alloc(p, n); }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }

// LP_FAKE_CHAOS=1: small device -> host reads (status words, lengths, counters) come back as random bytes: the host
// must not turn a nonsensical device answer into an out-of-bounds access of its own.
static void copy(void* d, const void* s, size_t n, cudaMemcpyKind k) {
    static const bool chaos = getenv("LP_FAKE_CHAOS") != nullptr;
    if (!n) return;
    if (chaos && k == cudaMemcpyDeviceToHost && n <= 64 && (rand() & 3) == 0) {
        for (size_t i = 0; i < n; i++) ((uint8_t*)d)[i] = (uint8_t)rand();
        return;
    }
    memmove(d, s, n);
    if (chaos && k == cudaMemcpyDeviceToHost && n <= (64u << 10) && (rand() & 3) == 0)  // arrays of lengths / states
        for (size_t i = 0; i + 4 <= n; i += 4)
            if (rand() % 10 == 0) {
                const uint32_t v = (uint32_t)rand() * 2654435761u;
                memcpy((uint8_t*)d + i, &v, 4);
            }
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k) { copy(d, s, n, k); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) { copy(d, s, n, k); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t y = 0; y < h; y++) memmove((char*)d + y * dp, (const char*)s + y * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
// device symbols (the entropy kernel's phase counters): no kernel ever runs here, so they read as zero
cudaError_t cudaMemcpyFromSymbol(void* d, const void*, size_t n, size_t, cudaMemcpyKind) { memset(d, 0, n); return cudaSuccess; }
cudaError_t cudaMemcpyToSymbol(const void*, const void*, size_t, size_t, cudaMemcpyKind) { return cudaSuccess; }
cudaError_t cudaMemset2DAsync(void* d, size_t p, int v, size_t w, size_t h, cudaStream_t) {
    for (size_t y = 0; y < h; y++) memset((char*)d + y * p, v, w);
    return cudaSuccess;
}

cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    switch (a) {  // (cub sizes its launch loops with these: a 0 would make them spin)
        case cudaDevAttrMultiProcessorCount: *v = 148; break;
        case cudaDevAttrMaxGridDimX: *v = 2147483647; break;
        case cudaDevAttrMaxGridDimY: case cudaDevAttrMaxGridDimZ: *v = 65535; break;
        case cudaDevAttrMaxThreadsPerBlock: *v = 1024; break;
        case cudaDevAttrMaxSharedMemoryPerBlock: *v = 48 * 1024; break;
        case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 227 * 1024; break;
        case cudaDevAttrWarpSize: *v = 32; break;
        case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
        case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
        default: *v = 1; break;
    }
    return cudaSuccess;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) { *n = 3; return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, const void*) {
    memset(a, 0, sizeof *a);
    a->maxThreadsPerBlock = 1024;
    a->ptxVersion = 100;
    a->binaryVersion = 100;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)calloc(1, 8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.f; return cudaSuccess; }
}
