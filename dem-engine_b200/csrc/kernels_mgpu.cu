// kernels_mgpu.cu -- domain decomposition of the stepping path over the GPUs of one box (the reference
// has nothing of the kind: its "multi-GPU" is the kT/dT thread pair of src/DEM/APIPublic.cpp:35-48).
//
// Every rank keeps GLOBALLY indexed owner arrays (1M owners x 112 B is nothing next to 180 GB of HBM), so an owner
// never changes its index; what changes is who integrates it.  Rank r owns the owners whose centre lies in its
// x-slab and additionally holds exact copies ("ghosts") of the neighbours' owners within the halo width of a cut.
// EVERYTHING here is device-driven over peer-mapped memory (NVLink stores + system-scope flag words): no library call,
// no host synchronisation, and no host-side argument changes from one step or rebuild to the next -- exchange numbers
// and all counts live in device memory -- so whole steps and whole rebuilds replay as CUDA graphs on every rank.
//   * per step   : k_mg_exchange, right after the integrator, copies the {state, spin} records (80 B) of the own owners
//                  inside a halo into the neighbours' receive buffers, publishes the exchange number to the neighbours,
//                  waits for theirs and scatters what they stored here into the same global slots;
//   * per rebuild: max |v| is all-gathered through per-rank mailboxes (same cell grid everywhere), ownership is
//                  re-decided from positions by the same rule on both sides of a cut (the data is an exact copy, so
//                  both sides agree without talking) walking only the owners this rank already holds, the halo
//                  membership lists + records + counts are pushed to the neighbours, and the ordinary rebuild runs over
//                  the active (own + ghost) spheres only; its verdict (overflow -> poison) is all-gathered too.
// Contacts across a cut are evaluated on BOTH ranks from identical inputs (same roles, same arithmetic), each rank
// applying the wrench to its own owners only: no reverse force exchange, and the contact history stays identical on
// both sides.  Ghost--ghost contacts inside the halo are evaluated too (history only) so that an owner that later
// crosses the cut finds the history of all its contacts already present on the rank that takes it over.
// Wall and mesh owners are replicated ("owned" everywhere): exact while they are fixed or follow a prescribed motion.
#include <algorithm>

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
__device__ __forceinline__ uint32_t warp_claim_n(uint32_t count, uint32_t* cursor) {
    const int lane = threadIdx.x & 31;
    uint32_t inc = count;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + inc - count;
}

// Append with ONE global atomic per CTA and list (16 000 warps adding to the same word serialise in L2: ~40 us at half a
// million owners): warps claim their stretch of the CTA's run in shared memory, thread 0 claims the CTA's run in the
// global cursor.  Must be called by all threads of the CTA; count = entries this thread appends (0..), returns its
// first slot.  sm: 2 words of shared memory.
__device__ __forceinline__ uint32_t block_claim(uint32_t count, uint32_t* cursor, uint32_t* sm) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) sm[0] = 0u;
    __syncthreads();
    uint32_t inc = count;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t wbase = 0;
    if (lane == 31 && total) wbase = atomicAdd(&sm[0], total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    __syncthreads();
    if (threadIdx.x == 0) sm[1] = sm[0] ? atomicAdd(cursor, sm[0]) : 0u;
    __syncthreads();
    const uint32_t r = sm[1] + wbase + inc - count;
    __syncthreads();  // (sm is reused by the next call)
    return r;
}

__device__ __forceinline__ unsigned long long* mg_flag(char* block, int dir) {
    return reinterpret_cast<unsigned long long*>(block + MG_HDR_STEP_FLAG) + dir;
}
__device__ __forceinline__ uint32_t* mg_recv_count(char* block, int par, int dir) {
    return reinterpret_cast<uint32_t*>(block + MG_HDR_RECV_COUNT) + par * 2 + dir;
}

// ---- step 2: re-decide ownership from positions, walking the owners of the previous cycle's active list only; own
// owners go to the new active list, those inside the halo of a cut additionally to that cut's send list ----
//   flag[g]: 0 unknown here, 1 own, 2 ghost.   counts[par]: [0] own, [1] send-left, [2] send-right, [3] active
__global__ void __launch_bounds__(256) k_mg_classify(const __grid_constant__ DevParams P, const __grid_constant__ MgDev M,
                                                     const GridInfo* __restrict__ grid, int par) {
    if (P.flags[DEM_FLAG_POISON]) return;
    __shared__ uint32_t sm[2];
    const uint32_t n = M.counts[par ^ 1][3];
    const uint32_t nround = (n + blockDim.x - 1u) / blockDim.x * blockDim.x;  // whole CTAs take part in the claims
    const float halo = grid->halo;
    uint32_t* cnt = M.counts[par];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += gridDim.x * blockDim.x) {
        bool own = false, toL = false, toR = false;
        uint32_t g = 0;
        if (t < n) {
            g = M.active_list[par ^ 1][t];
            if (g >= M.nClumpOwners) {
                own = true;  // boundary / analytical / mesh owners are replicated and "owned" everywhere
            } else {
                double X, Y, Z;
                pos_decode(P.state[g].pos, P, X, Y, Z);
                const float x = (float)X;  // LBF-relative
                own = (x >= M.cut_lo) && (x < M.cut_hi);
                if (own) {
                    toL = M.has[0] && (x < M.cut_lo + halo);
                    toR = M.has[1] && (x >= M.cut_hi - halo);
                }
            }
            // (ghosts are re-flagged when the neighbour's list arrives)
            M.flag[g] = own ? (toL ? 3 : (toR ? 4 : 1)) : 0;
        }
        const uint32_t sa = block_claim(own ? 1u : 0u, &cnt[3], sm);
        if (own) M.active_list[par][sa] = g;
        block_claim((own && g < M.nClumpOwners) ? 1u : 0u, &cnt[0], sm);
        const uint32_t sl = block_claim(toL ? 1u : 0u, &cnt[1], sm);
        const uint32_t sr = block_claim(toR ? 1u : 0u, &cnt[2], sm);
        if (toL && sl < M.cap) M.send_gid[par][0][sl] = g;
        if (toR && sr < M.cap) M.send_gid[par][1][sr] = g;
        if ((toL && sl >= M.cap) || (toR && sr >= M.cap)) atomicOr(&P.flags[DEM_FLAG_HALO], 8u);
    }
}

// ---- step 3: push the membership lists, their records and their lengths into the neighbours' blocks, then publish the
// exchange number (last block to finish) ----
__global__ void __launch_bounds__(256) k_mg_push_full(const __grid_constant__ DevParams P, const __grid_constant__ MgDev M, int par) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const unsigned long long e = *M.epoch + 1ull;
    const int half = (int)(e & 1ull);
    for (int d = 0; d < 2; d++) {
        if (!M.has[d]) continue;
        char* peer = M.peer_block[M.rank + (d == 0 ? -1 : 1)];
        const uint32_t n = min(M.counts[par][1 + d], M.cap);
        uint32_t* pg = reinterpret_cast<uint32_t*>(peer + mg_off_gid(M.cap, par, 1 - d));
        int4* pr = reinterpret_cast<int4*>(peer + mg_off_rec(M.cap, 1 - d, half));
        const uint32_t* gid = M.send_gid[par][d];
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n * 5u; t += gridDim.x * blockDim.x) {
            const uint32_t i = t / 5u, part = t - i * 5u;
            const uint32_t g = gid[i];
            const int4* src = (part < 4) ? reinterpret_cast<const int4*>(P.state + g) + part
                                         : reinterpret_cast<const int4*>(P.spin + g);
            pr[(size_t)i * 5u + part] = *src;
            if (part == 0) pg[i] = g;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) *mg_recv_count(peer, par, 1 - d) = n;
    }
    // publish: all stores of this grid, then the flag (last block to arrive does it)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&M.block_ctr[1], 1u);
        if (done == gridDim.x - 1) {
            M.block_ctr[1] = 0;
            __threadfence_system();
            for (int d = 0; d < 2; d++)
                if (M.has[d]) st_release_sys(mg_flag(M.peer_block[M.rank + (d == 0 ? -1 : 1)], 1 - d), e);
        }
    }
}

// wait until both neighbours have published exchange e (called by one thread per block)
__device__ __forceinline__ void mg_wait_neighbours(const DevParams& P, const MgDev& M, unsigned long long e) {
    const long long t0 = clock64();
    for (int d = 0; d < 2; d++) {
        if (!M.has[d]) continue;
        while (ld_acquire_sys(mg_flag(M.my_block, d)) < e) {
            if (clock64() - t0 > MG_SPIN_TIMEOUT_CYCLES) {
                atomicOr(&P.flags[DEM_FLAG_HALO], 64u);
                break;
            }
        }
    }
}
// the last block to finish advances the exchange counter (every block read it before taking part)
__device__ __forceinline__ void mg_finish_exchange(const MgDev& M, unsigned long long e) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(&M.block_ctr[0], 1u);
        if (done == gridDim.x - 1) {
            M.block_ctr[0] = 0;
            *M.epoch = e;
        }
    }
}

// ---- step 4: take in the neighbours' lists: scatter the records, flag the owners as ghosts, append them to the new
// active list ----
__global__ void __launch_bounds__(256) k_mg_pull_full(const __grid_constant__ DevParams P, const __grid_constant__ MgDev M, int par) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const unsigned long long e = *M.epoch + 1ull;
    const int half = (int)(e & 1ull);
    if (threadIdx.x == 0) mg_wait_neighbours(P, M, e);
    __syncthreads();
    for (int d = 0; d < 2; d++) {
        if (!M.has[d]) continue;
        const uint32_t n = min(__ldcg(mg_recv_count(M.my_block, par, d)), M.cap);
        const uint32_t* gid = reinterpret_cast<const uint32_t*>(M.my_block + mg_off_gid(M.cap, par, d));
        const int4* rec = reinterpret_cast<const int4*>(M.my_block + mg_off_rec(M.cap, d, half));
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n * 5u; t += gridDim.x * blockDim.x) {
            const uint32_t i = t / 5u, part = t - i * 5u;
            const uint32_t g = __ldcg(&gid[i]);  // written by the peer: never through this SM's L1
            int4* dst = (part < 4) ? reinterpret_cast<int4*>(P.state + g) + part : reinterpret_cast<int4*>(P.spin + g);
            *dst = __ldcg(&rec[(size_t)i * 5u + part]);
        }
        const uint32_t nround = (n + 31u) & ~31u;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
            const bool in = i < n;
            const uint32_t g = in ? __ldcg(&gid[i]) : 0u;
            if (in) M.flag[g] = 2;
            const uint32_t slot = warp_append(in, &M.counts[par][3]);
            if (in) M.active_list[par][slot] = g;
        }
    }
    mg_finish_exchange(M, e);
}

// ---- step 5: compact list of the spheres of the active owners: the rebuild walks this list instead of all spheres, so
// its cost follows the slab, not the whole bed.  The spheres of a clump stay adjacent, which is all the owner-major
// contact order needs. ----
__global__ void __launch_bounds__(256) k_mg_active_spheres(const __grid_constant__ DevParams P, const __grid_constant__ MgDev M, int par) {
    if (P.flags[DEM_FLAG_POISON]) return;
    if (M.owner_sph) {
        __shared__ uint32_t sm[2];
        const uint32_t n = M.counts[par][3];
        const uint32_t nround = (n + blockDim.x - 1u) / blockDim.x * blockDim.x;
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += gridDim.x * blockDim.x) {
            uint2 r = make_uint2(0u, 0u);
            if (t < n) r = M.owner_sph[M.active_list[par][t]];
            const uint32_t base = block_claim(r.y, &M.counts[par][4], sm);
            for (uint32_t k = 0; k < r.y; k++) M.act_sph[par][base + k] = r.x + k;
        }
    } else {
        // (spheres of an owner are not contiguous in this input: scan them all)
        const uint32_t nround = (P.nSpheres + 31u) & ~31u;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
            const bool act = (i < P.nSpheres) && (M.flag[P.sph[i].x] != 0);
            const uint32_t slot = warp_append(act, &M.counts[par][4]);
            if (act) M.act_sph[par][slot] = i;
        }
    }
}

// ---- step 1b: the list buffers of parity par were last written two rebuilds ago, for the active spheres of THAT cycle
// (still listed in act_sph[par]): empty their segments, so that spheres this rank no longer holds leave nothing stale
// behind for later history look-ups.  P carries the NEW lists. ----
__global__ void __launch_bounds__(256) k_mg_clear_segs(const __grid_constant__ DevParams P, const __grid_constant__ MgDev M, int par) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t n = M.counts[par][4];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t i = M.act_sph[par][t];
        P.ss.seg_count[i] = 0u;
        P.sn.seg_count[i] = 0u;
        P.sa.seg_count[i] = 0u;
        if (P.nTri) P.st.seg_count[i] = 0u;
    }
}

int launch_mg_redistribute(const DevParams& P, const MgDev& M, const GridInfo* grid, int par, int num_sms, cudaStream_t s) {
    const int g = num_sms * 2;
    k_mg_clear_segs<<<g, 256, 0, s>>>(P, M, par);
    launch_zero_u32(M.counts[par], 8, P.flags, num_sms, s);
    // (one owner per thread where the count allows: the per-owner work is a dependent chain of loads)
    const int gfull = (int)std::min<uint32_t>((P.nOwners + 255u) / 256u, (uint32_t)num_sms * 16u);
    k_mg_classify<<<std::max(gfull, 1), 256, 0, s>>>(P, M, grid, par);
    k_mg_push_full<<<std::min(g, 64), 256, 0, s>>>(P, M, par);
    k_mg_pull_full<<<std::min(g, 64), 256, 0, s>>>(P, M, par);
    k_mg_active_spheres<<<std::max(gfull, 1), 256, 0, s>>>(P, M, par);
    return 6;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-step exchange, ONE kernel right after the integrator.  Phase 1: every block copies its share of the {state, spin}
// records (80 B) of my halo owners -- the send lists of this cycle -- into the neighbours' receive buffers (coalesced
// 16-byte stores over NVLink); the last block to finish tells the neighbours that exchange e is complete.  Phase 2: every
// block waits for the neighbours' word and scatters the records they stored here into the global slots.  No rank waits
// for another rank's wait (phase 1 depends on nobody), so there is no cycle.  Receive buffers are double buffered by
// exchange parity: a rank can run at most one exchange ahead of its neighbour (it cannot finish exchange e before the
// neighbour has pushed e), so the half written in exchange e+1 is the one the neighbour finished reading in e-1.  The
// membership lists are double buffered by cycle parity for the same reason.  (Storing the records from inside the
// integrator instead -- one or two lanes of nearly every warp own a halo owner -- made the integrator as slow on half the
// owners (47 us) as it is on all of them on one GPU.)
__global__ void __launch_bounds__(256) k_mg_exchange(const __grid_constant__ DevParams P, const __grid_constant__ MgDev M, int par) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const unsigned long long e = *M.epoch + 1ull;
    const int half = (int)(e & 1ull);
    for (int d = 0; d < 2; d++) {
        if (!M.has[d]) continue;
        char* peer = M.peer_block[M.rank + (d == 0 ? -1 : 1)];
        const uint32_t n = min(M.counts[par][1 + d], M.cap);
        int4* pr = reinterpret_cast<int4*>(peer + mg_off_rec(M.cap, 1 - d, half));
        const uint32_t* gid = M.send_gid[par][d];
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n * 5u; t += gridDim.x * blockDim.x) {
            const uint32_t i = t / 5u, part = t - i * 5u;
            const uint32_t g = gid[i];
            const int4* src = (part < 4) ? reinterpret_cast<const int4*>(P.state + g) + part
                                         : reinterpret_cast<const int4*>(P.spin + g);
            pr[(size_t)i * 5u + part] = *src;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(&M.block_ctr[1], 1u);
        if (done == gridDim.x - 1) {
            M.block_ctr[1] = 0;
            __threadfence_system();
            for (int d = 0; d < 2; d++)
                if (M.has[d]) st_release_sys(mg_flag(M.peer_block[M.rank + (d == 0 ? -1 : 1)], 1 - d), e);
        }
        mg_wait_neighbours(P, M, e);
    }
    __syncthreads();
    for (int d = 0; d < 2; d++) {
        if (!M.has[d]) continue;
        const uint32_t n = min(__ldcg(mg_recv_count(M.my_block, par, d)), M.cap);
        const uint32_t* gid = reinterpret_cast<const uint32_t*>(M.my_block + mg_off_gid(M.cap, par, d));
        const int4* rec = reinterpret_cast<const int4*>(M.my_block + mg_off_rec(M.cap, d, half));
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n * 5u; t += gridDim.x * blockDim.x) {
            const uint32_t i = t / 5u, part = t - i * 5u;
            const uint32_t g = __ldcg(&gid[i]);
            int4* dst = (part < 4) ? reinterpret_cast<int4*>(P.state + g) + part : reinterpret_cast<int4*>(P.spin + g);
            *dst = __ldcg(&rec[(size_t)i * 5u + part]);
        }
    }
    mg_finish_exchange(M, e);
}

int launch_mg_pull(const DevParams& P, const MgDev& M, int par, int num_sms, cudaStream_t s) {
    k_mg_exchange<<<std::min(num_sms, 64), 256, 0, s>>>(P, M, par);
    return 1;
}

// device-side barrier over the ranks (one all-gather with nothing in it): what bench.py enqueues right before its first
// timing event, so that the timed region starts with all GPUs level
__global__ void __launch_bounds__(32) k_mg_barrier(const __grid_constant__ DevParams P, const __grid_constant__ MgDev M) {
    uint32_t o0, o1;
    mg_allgather(M, 0u, 0u, o0, o1, P.flags);
}
int launch_mg_barrier(const DevParams& P, const MgDev& M, cudaStream_t s) {
    k_mg_barrier<<<1, 32, 0, s>>>(P, M);
    return 1;
}

// In-process decomposition (all contexts in one process, peer access enabled): copy into MY arrays the records of the
// clump owners a peer owns, straight out of the peer's arrays -- afterwards this context holds the merged state of the
// whole system for trackers / writers / inspectors.
__global__ void __launch_bounds__(256) k_mg_gather_owned(const __grid_constant__ DevParams P,
                                                         const OwnerState* __restrict__ peer_state,
                                                         const float4* __restrict__ peer_spin,
                                                         const uint8_t* __restrict__ peer_flag, uint32_t nClumpOwners) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nClumpOwners * 5u; t += gridDim.x * blockDim.x) {
        const uint32_t o = t / 5u, part = t - o * 5u;
        if (peer_flag[o] != 1 && peer_flag[o] < 3) continue;  // (1, 3, 4: the peer owns it)
        int4* dst = (part < 4) ? reinterpret_cast<int4*>(P.state + o) + part : reinterpret_cast<int4*>(P.spin + o);
        const int4* src = (part < 4) ? reinterpret_cast<const int4*>(peer_state + o) + part
                                     : reinterpret_cast<const int4*>(peer_spin + o);
        *dst = *src;
    }
}
int launch_mg_gather_owned(const DevParams& P, const OwnerState* peer_state, const float4* peer_spin,
                           const uint8_t* peer_flag, uint32_t nClumpOwners, cudaStream_t s) {
    if (nClumpOwners == 0) return 0;
    k_mg_gather_owned<<<148 * 4, 256, 0, s>>>(P, peer_state, peer_spin, peer_flag, nClumpOwners);
    return 1;
}

}  // namespace demb
