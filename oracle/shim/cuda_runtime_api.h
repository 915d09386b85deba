// Host stand-in for the CUDA runtime, used ONLY to compile the unmodified reference
// sources (/root/reference/PathTracerAP/*.cpp) for the CPU as the parity oracle
// (oracle/_ref/libptap_ref.so).  TEST INFRASTRUCTURE - never linked into the product.
//
// Every CUDA call the reference makes (GPUMemoryPool.h:19-38, Renderer.cpp:578-640,
// utility.h:25-40) is mapped to its libc equivalent; kernels launched with
// LAUNCH(k, grid, block, args...) run as an OpenMP loop over blocks with the
// usual threadIdx/blockIdx/blockDim built-ins provided as thread-locals.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>

#define __host__
#define __device__
#define __global__
#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
#ifndef __inline__
#define __inline__ inline
#endif

typedef int cudaError_t;
typedef int cudaError;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

template <typename T> static inline cudaError_t cudaMallocManaged(T** p, size_t n)
{
    *p = (T*)calloc(n ? n : 1, 1);
    return *p ? cudaSuccess : 2;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = 0; *t = 0; return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "host shim"; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct ptap_uint3 { unsigned x, y, z; };
extern thread_local ptap_uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim;

using std::min;
using std::max;

// Optional knobs set by the harness (oracle/ref_harness.cpp).
extern int ptap_ref_quiet;       // 1: the host launcher does nothing special; kept for symmetry

#define LAUNCH(kernel, grid, block, ...)                                          \
    do {                                                                          \
        const dim3 ptap_g = (grid), ptap_b = (block);                             \
        _Pragma("omp parallel for schedule(dynamic, 64)")                         \
        for (long long ptap_blk = 0; ptap_blk < (long long)ptap_g.x; ++ptap_blk) {\
            blockDim = ptap_b;                                                    \
            blockIdx.x = (unsigned)ptap_blk; blockIdx.y = blockIdx.z = 0;         \
            for (unsigned ptap_t = 0; ptap_t < ptap_b.x; ++ptap_t) {              \
                threadIdx.x = ptap_t; threadIdx.y = threadIdx.z = 0;              \
                kernel(__VA_ARGS__);                                              \
            }                                                                     \
        }                                                                         \
    } while (0)
