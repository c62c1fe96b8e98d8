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
int launch_mg_active_list(const DevParams& P, const MgParams& M, cudaStream_t s) {
    if (P.nOwners) k_mg_active_list<<<(P.nOwners + 255) / 256, 256, 0, s>>>(P, M);
    return 1;
}

}  // namespace demb
