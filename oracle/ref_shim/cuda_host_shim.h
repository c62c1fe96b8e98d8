/*
 * cuda_host_shim.h -- ORACLE BUILD INFRASTRUCTURE (not product code).
 *
 * Lets g++ compile the reference's per-thread CUDA kernel TEXT (src/kernel/*.cu of /root/reference)
 * as ordinary host C++, so that the reference's own arithmetic can be executed on the CPU and used
 * to pin oracle/dem_oracle.c.  Only kernels without intra-block cooperation are ever *called*
 * through this shim (one "thread" at a time); kernels that use __shared__/__syncthreads compile
 * but are never launched.
 */
#ifndef DEM_CUDA_HOST_SHIM_H
#define DEM_CUDA_HOST_SHIM_H

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h> /* vector types + make_* only; g++ sees __host__/__device__ as empty */

#ifndef __global__
#define __global__
#endif
#ifndef __shared__
#define __shared__ static
#endif
#ifndef __constant__
#define __constant__
#endif

struct ShimDim3 {
    unsigned int x, y, z;
};
static thread_local ShimDim3 shim_blockIdx = {0, 0, 0}, shim_blockDim = {1, 1, 1}, shim_threadIdx = {0, 0, 0};
#define blockIdx shim_blockIdx
#define blockDim shim_blockDim
#define threadIdx shim_threadIdx

static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline float atomicAdd(float* addr, float v) {
    float old = *addr;
    *addr = old + v;
    return old;
}
static inline unsigned int atomicAdd(unsigned int* addr, unsigned int v) {
    unsigned int old = *addr;
    *addr = old + v;
    return old;
}
/* round-up device intrinsics used by snap_to_face (src/kernel/DEMCollisionKernels.cu:76-78);
 * evaluated round-to-nearest on the host (<= 1 ulp apart). */
static inline double __drcp_ru(double x) { return 1.0 / x; }
static inline double __dmul_ru(double a, double b) { return a * b; }

using std::isfinite;
using std::sqrt;
using std::log;
using std::abs;

#endif
