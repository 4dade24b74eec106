/*
 * cuda_runtime.h -- CPU stand-in for the handful of CUDA runtime calls the reference's
 * host code makes (Deff2D.cuh:904-1021, 1163-1314).  TEST INFRASTRUCTURE ONLY: it lets
 * oracle/build_ref.sh compile the reference's UNMODIFIED host logic with g++ and run
 * its kernel body (Deff2D.cuh:69-92) on host threads, so the oracle restatement and the
 * golden vectors are pinned to the reference's own code.  Written from scratch; it
 * contains no reference text.
 */
#ifndef ORACLE_CUDA_SHIM_RUNTIME_H
#define ORACLE_CUDA_SHIM_RUNTIME_H
#include <chrono>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__

struct shim_uint3 { unsigned int x, y, z; };
static thread_local shim_uint3 blockIdx = {0, 0, 0};
static thread_local shim_uint3 blockDim = {1, 1, 1};
static thread_local shim_uint3 threadIdx = {0, 0, 0};

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind {
    cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
    cudaMemcpyDeviceToDevice = 3
};

static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceReset() { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "shim error" : "no error"; }
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = std::malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }

/* Device-to-device "swap by copy" (Deff2D.cuh:1281) is by far the hottest memcpy of the
 * reference loop; spread it over the host threads like the kernel itself. */
static inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind kind)
{
    if (kind == cudaMemcpyDeviceToDevice && n >= (1u << 20)) {
        const size_t chunk = 1u << 18;
        const long nchunks = (long)((n + chunk - 1) / chunk);
#pragma omp parallel for schedule(static)
        for (long c = 0; c < nchunks; c++) {
            size_t off = (size_t)c * chunk;
            size_t len = (off + chunk <= n) ? chunk : n - off;
            std::memcpy((char *)dst + off, (const char *)src + off, len);
        }
    } else {
        std::memcpy(dst, src, n);
    }
    return cudaSuccess;
}

struct shim_event { std::chrono::steady_clock::time_point t; };
typedef shim_event *cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new shim_event(); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, int) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}

/* kernel<<<grid, block>>>(args...) is rewritten by build_ref.sh's sed to
 * SHIM_LAUNCH(kernel, grid, block, args...): every (block, thread) pair of the launch
 * runs the kernel body once, blocks spread over OpenMP threads. */
#define SHIM_LAUNCH(kernel, nblocks, nthreads, ...)                                   \
    do {                                                                              \
        const int shim_nb = (int)(nblocks);                                           \
        const unsigned shim_nt = (unsigned)(nthreads);                                \
        _Pragma("omp parallel for schedule(static)")                                  \
        for (int shim_b = 0; shim_b < shim_nb; shim_b++) {                            \
            blockIdx.x = (unsigned)shim_b;                                            \
            blockDim.x = shim_nt;                                                     \
            for (unsigned shim_t = 0; shim_t < shim_nt; shim_t++) {                   \
                threadIdx.x = shim_t;                                                 \
                kernel(__VA_ARGS__);                                                  \
            }                                                                         \
        }                                                                             \
    } while (0)

#endif
