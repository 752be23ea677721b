// TEST INFRASTRUCTURE.  Host stand-in for <cuda_runtime.h>, so that the reference's VX3 "device" sources
// (/root/reference/Voxelyze3/src/VX3/*.cu, plain scalar C++ behind __device__) compile with g++ and run on the CPU as
// oracle/_ref/vx3_ref.  Nothing here is physics: qualifiers vanish, the CUDA allocation calls become malloc/memcpy, atomics
// become plain read-modify-writes, and a kernel launch `k<<<grid, block>>>(args)` (rewritten by the Makefile's sed into
// `vx3shim::launcher(grid, block, k)(args)`) runs every thread of the grid one after the other, in index order.
// A serial schedule is ONE valid execution of the reference's (racy) kernels; the header of oracle/vx3_ref_harness.cpp says
// which order is used and why.
#pragma once
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cmath>

#define __device__
#define __host__
#define __global__
#define __inline__ inline
#define __forceinline__ inline

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_shim {
    unsigned x = 0, y = 0, z = 0;
};
inline uint3_shim threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaLimit { cudaLimitMallocHeapSize = 2, cudaLimitStackSize = 0 };
inline const char *cudaGetErrorString(cudaError_t) { return "shim"; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) {
    *p = (T *)calloc(n ? n : 1, 1);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, void * = nullptr) {
    memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceReset() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaMemGetInfo(size_t *a, size_t *b) {
    *a = *b = (size_t)1 << 30;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxPotentialBlockSize(int *minGrid, int *block, F, size_t = 0, int = 0) {
    *minGrid = 1;
    *block = 512;
    return cudaSuccess;
}

template <class T> inline T atomicAdd(T *p, T v) {
    T old = *p;
    *p = old + v;
    return old;
}
inline int atomicCAS(int *p, int cmp, int v) {
    int old = *p;
    if (old == cmp)
        *p = v;
    return old;
}
inline int atomicExch(int *p, int v) {
    int old = *p;
    *p = v;
    return old;
}
inline double normcdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

// recording call the Makefile's sed adds after `VX3_Collision collision(voxel1, voxel2);` (defined in oracle/vx3_ref_harness.cpp)
void vx3shim_pair_hook(const void *voxel1, const void *voxel2);

namespace vx3shim {
// thread order of a launch: the x index (global) outermost and ascending, y inside it — for gpu_update_attach
// (first = x, second = y < first) this is "for first: for second < first", the order oracle/vx3_oracle.cpp sums contact forces in
template <class F> struct Launcher {
    dim3 g, b;
    F f;
    template <class... A> void operator()(A... a) {
        const auto st = threadIdx, sb = blockIdx;
        const dim3 sbd = blockDim, sgd = gridDim;
        blockDim = b;
        gridDim = g;
        for (unsigned z = 0; z < g.z * b.z; z++)
            for (unsigned x = 0; x < g.x * b.x; x++)
                for (unsigned y = 0; y < g.y * b.y; y++) {
                    blockIdx.x = x / b.x, threadIdx.x = x % b.x;
                    blockIdx.y = y / b.y, threadIdx.y = y % b.y;
                    blockIdx.z = z / b.z, threadIdx.z = z % b.z;
                    f(a...);
                }
        threadIdx = st, blockIdx = sb, blockDim = sbd, gridDim = sgd;
    }
};
template <class F> inline Launcher<F> launcher(dim3 g, dim3 b, F f) { return Launcher<F>{g, b, f}; }
} // namespace vx3shim
