/*
 * cuda_host_shim.h -- ORACLE BUILD INFRASTRUCTURE (not product code).
 *
 * Lets g++ compile the reference's per-thread CUDA kernel TEXT (src/kernel/*.cu of /root/reference)
 * as ordinary host C++, so that the reference's own arithmetic can be executed on the CPU and used
 * to pin oracle/dem_oracle.c.  Kernels without intra-block cooperation are called one "thread" at
 * a time.  Block-cooperative kernels (__shared__ + __syncthreads: the per-bin contact sweeps) run one
 * block at a time with every CUDA thread of the block on its own fiber (ref_harness.cpp: launch_coop):
 * __shared__ becomes a static (blocks run one after the other), __syncthreads hands control to the
 * block's scheduler, which resumes the fibers only once all of them have arrived.
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
/* (cuda_runtime.h defines __shared__ as nothing for a host compiler, which would make the arrays per-fiber locals) */
#undef __shared__
#define __shared__ static
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

static thread_local void (*shim_sync_hook)() = nullptr; /* set by launch_coop for the duration of a cooperative launch */
static inline void __syncthreads() {
    if (shim_sync_hook) shim_sync_hook();
}
static inline void __threadfence() {}
/* atomicAdd(float*): a real atomic (CAS loop) so that the harness may run the per-thread kernels on several host
 * threads; with one thread it degenerates to a plain add in program order (used by the bit-exactness tests). */
static inline float atomicAdd(float* addr, float v) {
    unsigned int* ia = reinterpret_cast<unsigned int*>(addr);
    unsigned int old = __atomic_load_n(ia, __ATOMIC_RELAXED), assumed;
    float fold;
    do {
        assumed = old;
        __builtin_memcpy(&fold, &assumed, 4);
        const float fnew = fold + v;
        unsigned int inew;
        __builtin_memcpy(&inew, &fnew, 4);
        if (__atomic_compare_exchange_n(ia, &old, inew, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
    } while (true);
    return fold;
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
