// kernels_mgpu.cu -- domain decomposition of the stepping path over the GPUs of one box (the reference
// has nothing of the kind: its "multi-GPU" is the kT/dT thread pair of src/DEM/APIPublic.cpp:35-48).
//
// Every rank keeps GLOBALLY indexed owner arrays (1M owners x 112 B is nothing next to 180 GB of HBM), so an owner
// never changes its index; what changes is who integrates it.  Rank r owns the owners whose centre lies in its
// x-slab and additionally holds exact copies ("ghosts") of the neighbours' owners within the halo width of a cut.
//   * per step   : own owners inside the halo are packed (state 64 B + spin 16 B) and sent to the neighbour, which
//                  scatters them into the same global slots -- ncclSend/ncclRecv pairs grouped on the compute stream;
//   * per rebuild: ownership is re-decided from positions by the same rule on both sides of a cut (the data is an
//                  exact copy, so both sides agree without talking), the halo membership lists are exchanged, and the
//                  ordinary rebuild runs over the active (own + ghost) spheres only.
// Contacts across a cut are evaluated on BOTH ranks from identical inputs (same roles, same arithmetic), each rank
// applying the wrench to its own owners only: no reverse force exchange, and the contact history stays identical on
// both sides.  Ghost--ghost contacts inside the halo are evaluated too (history only) so that an owner that later
// crosses the cut finds the history of all its contacts already present on the rank that takes it over.
#include "dem_kernels.h"

namespace demb {

__device__ __forceinline__ uint32_t warp_append(bool pred, uint32_t* cursor) {
    const uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(m & ((1u << lane) - 1u)) : 0xffffffffu;
}

// Re-decide ownership from positions and list the own owners that sit in the halo of either cut.
//   flag[g]: 0 unknown here, 1 own, 2 ghost.   counts: [0] own, [1] send-left, [2] send-right
__global__ void __launch_bounds__(256) k_mg_classify(const __grid_constant__ DevParams P, MgParams M) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    bool own = false, toL = false, toR = false;
    if (g < P.nOwners) {
        const uint8_t f = M.flag[g];
        if (f != 0) {
            if (g >= M.nClumpOwners) {
                own = true;  // boundary / analytical owners are replicated and "owned" everywhere
            } else {
                double X, Y, Z;
                pos_decode(P.state[g].pos, P, X, Y, Z);
                const float x = (float)X;  // LBF-relative
                own = (x >= M.cut_lo) && (x < M.cut_hi);
                if (own) {
                    const float halo = M.grid->halo;
                    toL = M.has_left && (x < M.cut_lo + halo);
                    toR = M.has_right && (x >= M.cut_hi - halo);
                }
            }
        }
        M.flag[g] = own ? 1 : 0;  // ghosts are re-flagged when the neighbour's list arrives
    }
    const uint32_t sl = warp_append(toL, &M.counts[1]);
    const uint32_t sr = warp_append(toR, &M.counts[2]);
    if (toL && sl < M.send_cap) M.send_gid[0][sl] = g;
    if (toR && sr < M.send_cap) M.send_gid[1][sr] = g;
    if ((toL && sl >= M.send_cap) || (toR && sr >= M.send_cap)) atomicOr(&P.flags[0], 8u);
}

// gather {state, spin} of the listed owners into a contiguous send buffer (80 B per owner)
__global__ void __launch_bounds__(256) k_mg_pack(const __grid_constant__ DevParams P, const uint32_t* __restrict__ gid,
                                                 uint32_t n, int4* __restrict__ buf) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 5u) return;
    const uint32_t i = t / 5u, part = t - i * 5u;
    const uint32_t g = gid[i];
    const int4* src = (part < 4) ? reinterpret_cast<const int4*>(P.state + g) + part
                                 : reinterpret_cast<const int4*>(P.spin + g);
    buf[(size_t)i * 5u + part] = *src;
}

// scatter received {state, spin} into the global slots and (at a rebuild) flag them as ghosts
__global__ void __launch_bounds__(256) k_mg_unpack(const __grid_constant__ DevParams P, const uint32_t* __restrict__ gid,
                                                   uint32_t n, const int4* __restrict__ buf, uint8_t* flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 5u) return;
    const uint32_t i = t / 5u, part = t - i * 5u;
    const uint32_t g = gid[i];
    int4* dst = (part < 4) ? reinterpret_cast<int4*>(P.state + g) + part : reinterpret_cast<int4*>(P.spin + g);
    *dst = buf[(size_t)i * 5u + part];
    if (flag && part == 0) flag[g] = 2;
}

// compact list of the active owners (own and ghost) for the integrator
__global__ void __launch_bounds__(256) k_mg_active_list(const __grid_constant__ DevParams P, MgParams M) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = (g < P.nOwners) && (M.flag[g] != 0);
    const uint32_t slot = warp_append(act, &M.counts[3]);
    if (act) M.active_list[slot] = g;
    const bool own = act && M.flag[g] == 1;
    warp_append(own, &M.counts[0]);
}

// compact list of the spheres of active owners: the rebuild walks this list instead of all spheres, so its cost follows
// the slab, not the whole bed.  Warps append in arrival order, lanes in sphere order: the spheres of a clump stay
// adjacent (up to a warp boundary), which is all the owner-major contact order needs.
__global__ void __launch_bounds__(256) k_mg_active_spheres(const __grid_constant__ DevParams P, MgParams M) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = (i < P.nSpheres) && (M.flag[P.sph[i].x] != 0);
    const uint32_t slot = warp_append(act, &M.counts[4]);
    if (act) M.act_sph[slot] = i;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-step exchange without a library call: every rank stores the {state, spin} records of its halo owners straight into
// the neighbour's receive buffer over NVLink (peer memory mapped with cudaIpc*), then publishes the epoch number in the
// neighbour's flag word; the neighbour's pull kernel waits for that word and scatters the records into its global
// slots.  Receive buffers are double buffered by epoch parity: a rank can run at most one exchange ahead of its
// neighbour (it cannot finish pull(e) before the neighbour has pushed epoch e), so the half written in epoch e+1 is
// the one the neighbour finished reading in epoch e-1.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) k_mg_push(const __grid_constant__ DevParams P, const __grid_constant__ MgP2P X) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n0 = X.n_send[0] * 5u, n1 = X.n_send[1] * 5u;
    if (t < n0 + n1) {
        const int d = t < n0 ? 0 : 1;
        const uint32_t u = d == 0 ? t : t - n0;
        const uint32_t i = u / 5u, part = u - i * 5u;
        const uint32_t g = X.send_gid[d][i];
        const int4* src = (part < 4) ? reinterpret_cast<const int4*>(P.state + g) + part
                                     : reinterpret_cast<const int4*>(P.spin + g);
        X.peer_recv[d][(size_t)i * 5u + part] = *src;
    }
    // publish: all stores of this grid, then the flag (last block to arrive does it)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(X.block_counter, 1u);
        if (done == gridDim.x - 1) {
            *X.block_counter = 0;
            __threadfence_system();
            if (X.has[0]) st_release_sys(X.peer_flag[0], X.epoch);
            if (X.has[1]) st_release_sys(X.peer_flag[1], X.epoch);
        }
    }
}

__global__ void __launch_bounds__(256) k_mg_pull(const __grid_constant__ DevParams P, const __grid_constant__ MgP2P X) {
    if (X.publish && blockIdx.x == 0 && threadIdx.x == 0) {
        // fused push: the integrator (the previous kernel in this stream) stored my halo records into the neighbours'
        // buffers; tell them the epoch is complete BEFORE waiting for theirs (no rank waits for another's wait)
        __threadfence_system();
        if (X.has[0]) st_release_sys(X.peer_flag[0], X.epoch);
        if (X.has[1]) st_release_sys(X.peer_flag[1], X.epoch);
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int d = 0; d < 2; d++) {
            if (!X.has[d]) continue;
            while (ld_acquire_sys(X.my_flag[d]) < X.epoch) {
                if (clock64() - t0 > 20000000000ll) {  // ~10 s: the neighbour is gone; report instead of hanging
                    atomicOr(&P.flags[0], 64u);
                    break;
                }
            }
        }
    }
    __syncthreads();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n0 = X.n_recv[0] * 5u, n1 = X.n_recv[1] * 5u;
    if (t >= n0 + n1) return;
    const int d = t < n0 ? 0 : 1;
    const uint32_t u = d == 0 ? t : t - n0;
    const uint32_t i = u / 5u, part = u - i * 5u;
    const uint32_t g = X.recv_gid[d][i];
    int4* dst = (part < 4) ? reinterpret_cast<int4*>(P.state + g) + part : reinterpret_cast<int4*>(P.spin + g);
    *dst = __ldcg(&X.my_recv[d][(size_t)i * 5u + part]);  // written by the peer: never through this SM's L1
}

int launch_mg_push(const DevParams& P, const MgP2P& X, cudaStream_t s) {
    const uint32_t n = (X.n_send[0] + X.n_send[1]) * 5u;
    k_mg_push<<<(n + 255) / 256 + (n == 0 ? 1 : 0), 256, 0, s>>>(P, X);
    return 1;
}
int launch_mg_pull(const DevParams& P, const MgP2P& X, cudaStream_t s) {
    const uint32_t n = (X.n_recv[0] + X.n_recv[1]) * 5u;
    k_mg_pull<<<(n + 255) / 256 + (n == 0 ? 1 : 0), 256, 0, s>>>(P, X);
    return 1;
}

int launch_mg_classify(const DevParams& P, const MgParams& M, cudaStream_t s) {
    cudaMemsetAsync(M.counts, 0, sizeof(uint32_t) * 8, s);
    if (P.nOwners) k_mg_classify<<<(P.nOwners + 255) / 256, 256, 0, s>>>(P, M);
    return 1;
}
int launch_mg_pack(const DevParams& P, const uint32_t* gid, uint32_t n, void* buf, cudaStream_t s) {
    if (n == 0) return 0;
    k_mg_pack<<<(n * 5u + 255) / 256, 256, 0, s>>>(P, gid, n, reinterpret_cast<int4*>(buf));
    return 1;
}
int launch_mg_unpack(const DevParams& P, const uint32_t* gid, uint32_t n, const void* buf, uint8_t* flag, cudaStream_t s) {
    if (n == 0) return 0;
    k_mg_unpack<<<(n * 5u + 255) / 256, 256, 0, s>>>(P, gid, n, reinterpret_cast<const int4*>(buf), flag);
    return 1;
}
// owner id -> slot in the neighbour's receive buffer (the integrator's fused push looks its owners up here)
__global__ void __launch_bounds__(256) k_mg_send_map(const uint32_t* __restrict__ gid, uint32_t n, int32_t* __restrict__ slot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot[gid[i]] = (int32_t)i;
}
int launch_mg_send_map(const uint32_t* gid, uint32_t n, int32_t* slot, uint32_t nOwners, cudaStream_t s) {
    cudaMemsetAsync(slot, 0xff, sizeof(int32_t) * (size_t)nOwners, s);
    if (n) k_mg_send_map<<<(n + 255) / 256, 256, 0, s>>>(gid, n, slot);
    return 1;
}

int launch_mg_active_spheres(const DevParams& P, const MgParams& M, cudaStream_t s) {
    if (P.nSpheres) k_mg_active_spheres<<<(P.nSpheres + 255) / 256, 256, 0, s>>>(P, M);
    return 1;
}
int launch_mg_active_list(const DevParams& P, const MgParams& M, cudaStream_t s) {
    if (P.nOwners) k_mg_active_list<<<(P.nOwners + 255) / 256, 256, 0, s>>>(P, M);
    return 1;
}

}  // namespace demb
