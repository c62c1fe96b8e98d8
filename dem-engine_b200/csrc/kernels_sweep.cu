// kernels_sweep.cu -- the sphere--sphere contact-pair sweep of the rebuild, TMA-staged (sm_100a).
//
// Reference behaviour being reproduced: getNumberOfSphereContactsEachBin / populateSphSphContactPairsEachBin
// (src/kernel/DEMContactKernels_SphereSphere.cu:91-214,268-400): a per-bin all-pairs test staged in shared memory,
// run twice (count, then fill), with the acceptance rule of :57-89,172-214.  Here:
//   * cells are x-fastest, so the x-neighbours of a cell row are ONE contiguous run of the cell-sorted sphere stream;
//     every sphere looks only "forward" (the upper half of the 27-cell stencil = 5 runs: the own row behind the
//     sphere itself, (y+1,z), and (y-1..y+1, z+1)), so each pair is found once, by the sphere that comes first in
//     (cell, sphere id) order -- that sphere is geometry A of the contact;
//   * a CTA takes 128 consecutive positions of the sorted stream.  For each of the 5 run types the ranges its threads
//     need form one contiguous stretch of the stream (the map cell -> neighbour cell is monotonic), which ONE elected
//     thread stages into shared memory with 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx::bytes, SASS
//     UBLKCP + SYNCS): 16 B {x,y,z,r+margin} and 8 B {owner, r} per sphere.  All lanes then test pairs out of shared
//     memory -- no per-thread gathers from L2, which is what bound the previous one-thread-per-sphere kernel (LSU
//     wavefronts 70 %, 16 of 32 lanes active);
//   * two passes, COUNT (per-sphere counts -> exclusive scan in sphere-id order) and FILL, so the list is
//     deterministic, owner-major (the force kernel streams the A side and reduces it inside the warp) and there is NO
//     cap on the number of candidates per sphere (the reference allows up to 32768 spheres per bin,
//     DEMContactKernels_SphereSphere.cu:121-126).
#include "dem_kernels.h"

namespace demb {

constexpr int SW_THREADS = 128;  // sorted positions per work item
constexpr int SW_CH = 160;       // staged entries per run and phase (even: the 8-byte stream stays 16-byte aligned)

struct __align__(16) SweepSmem {
    float4 sph[5][SW_CH];
    uint2 aux[5][SW_CH];
    unsigned long long mbar;
    uint32_t rb[5], re[5];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// 1-D bulk copy global -> shared, completion signalled on the mbarrier (TMA engine; SASS UBLKCP.S.G)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// FILL = false: per-sphere counts into ss.seg_count / sn.seg_count.  FILL = true: the compiled records at
// seg_start[sphere] + k (seg_start = exclusive scan of the counts over sphere ids).
template <bool FILL>
__global__ void __launch_bounds__(SW_THREADS) k_sweep_tma(const __grid_constant__ DevParams P,
                                                          const __grid_constant__ CdParams C,
                                                          const uint32_t* __restrict__ keys) {
    if (P.flags[DEM_FLAG_POISON]) return;
    __shared__ SweepSmem sm;
    const int tid = threadIdx.x, lane = tid & 31;
    const GridInfo g = *C.grid;
    const uint32_t nSorted = C.cellStart[g.ncells];
    if (tid == 0) mbar_init(&sm.mbar, 1);
    __syncthreads();
    uint32_t parity = 0;
    const bool fam_on = (C.any_mask != 0) || (C.max_extra > 0.f);
    for (uint32_t base = blockIdx.x * SW_THREADS; base < nSorted; base += gridDim.x * SW_THREADS) {
        const uint32_t j = base + tid;
        const bool valid = j < nSorted;
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 myAux = make_uint2(0xffffffffu, 0u);
        uint32_t qb[5], qe[5];
#pragma unroll
        for (int r = 0; r < 5; r++) { qb[r] = 0xffffffffu; qe[r] = 0u; }
        if (valid) {
            me = C.sortedSph[j];
            myAux = C.sortedAux[j];
            const uint32_t key = keys[j];
            const int cx = (int)(key % g.nbx);
            const int cy = (int)((key / g.nbx) % g.nby);
            const int cz = (int)(key / (g.nbx * g.nby));
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, (int)g.nbx - 1);
#pragma unroll
            for (int r = 0; r < 5; r++) {
                // r=0: own row behind me; r=1: (dy=+1,dz=0); r=2..4: (dy=-1,0,+1; dz=+1)
                const int dy = (r == 0) ? 0 : (r == 1 ? 1 : r - 3);
                const int dz = (r < 2) ? 0 : 1;
                const int y = cy + dy, z = cz + dz;
                if (y < 0 || y >= (int)g.nby || z >= (int)g.nbz) continue;
                const uint32_t row = g.nbx * ((uint32_t)y + g.nby * (uint32_t)z);
                const uint32_t b = (r == 0) ? j + 1 : __ldg(&C.cellStart[row + x0]);
                const uint32_t e = __ldg(&C.cellStart[row + x1 + 1]);
                if (b < e) { qb[r] = b; qe[r] = e; }
            }
        }
        // ---- the stretch of the sorted stream this CTA needs for each run type ----
        if (tid < 5) { sm.rb[tid] = 0xffffffffu; sm.re[tid] = 0u; }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t lo = __reduce_min_sync(0xffffffffu, qb[r]);
            const uint32_t hi = __reduce_max_sync(0xffffffffu, qe[r]);
            if (lane == 0 && lo < hi) { atomicMin(&sm.rb[r], lo); atomicMax(&sm.re[r], hi); }
        }
        __syncthreads();
        uint32_t pos[5], end[5];
#pragma unroll
        for (int r = 0; r < 5; r++) {
            pos[r] = sm.rb[r] & ~1u;  // even start: the 8-byte stream is copied in 16-byte units
            end[r] = sm.re[r];
            if (sm.rb[r] >= sm.re[r]) { pos[r] = 0u; end[r] = 0u; }
        }
        __syncthreads();  // (rb / re are reset at the top of the next work item)
        uint32_t nT = 0, nN = 0;
        uint32_t slotT = 0, slotN = 0;
        uint32_t sid = 0, myMeta_z = 0, myFam = 0;
        float extraA = 0.f;
        if (valid && (FILL || fam_on)) {
            const uint4 meta = C.sortedMeta[j];
            sid = meta.y; myMeta_z = meta.z; myFam = meta.w;
            if (fam_on) extraA = P.familyExtraMargin[myFam];
            if (FILL) { slotT = P.ss.seg_start[sid]; slotN = P.sn.seg_start[sid]; }
        } else if (valid) {
            sid = C.sortedMeta[j].y;
        }
        // ---- phases: stage up to SW_CH entries of every run, test, advance ----
        for (;;) {
            uint32_t cnt[5];
            bool any = false;
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t left = (end[r] > pos[r]) ? end[r] - pos[r] : 0u;
                cnt[r] = min((uint32_t)SW_CH, (left + 1u) & ~1u);
                any |= cnt[r] != 0u;
            }
            if (!any) break;
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                uint32_t bytes = 0;
#pragma unroll
                for (int r = 0; r < 5; r++) bytes += cnt[r] * 24u;
                mbar_expect_tx(&sm.mbar, bytes);
#pragma unroll
                for (int r = 0; r < 5; r++)
                    if (cnt[r]) {
                        bulk_g2s(&sm.sph[r][0], C.sortedSph + pos[r], cnt[r] * 16u, &sm.mbar);
                        bulk_g2s(&sm.aux[r][0], C.sortedAux + pos[r], cnt[r] * 8u, &sm.mbar);
                    }
            }
            mbar_wait(&sm.mbar, parity);
            parity ^= 1u;
            if (valid) {
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const uint32_t lo = max(qb[r], pos[r]);
                    const uint32_t hi = min(qe[r], pos[r] + cnt[r]);
                    for (uint32_t q = lo; q < hi; q++) {
                        const float4 ot = sm.sph[r][q - pos[r]];
                        const float dx = me.x - ot.x, dy = me.y - ot.y, dz = me.z - ot.z;
                        const float d2 = dx * dx + dy * dy + dz * dz;
                        const float R = me.w + ot.w;
                        // superset of the double-precision test d2 <= R^2 && R - d > min(extraA, extraB)
                        // (DEMContactKernels_SphereSphere.cu:57-89)
                        if (d2 > R * R * 1.000001f + 1e-20f) continue;
                        const uint2 oa = sm.aux[r][q - pos[r]];
                        if (oa.x == myAux.x) continue;  // same owner
                        uint4 om = make_uint4(0, 0, 0, 0);
                        if (fam_on || FILL) om = __ldg(&C.sortedMeta[q]);
                        if (fam_on) {
                            if (C.any_mask && P.familyMasks[mask_pair(myFam, om.w)] != 0) continue;
                            const float Rt = R - fminf(extraA, P.familyExtraMargin[om.w]);
                            if (d2 > Rt * Rt * 1.000001f + 1e-20f) continue;
                        }
                        // do the un-inflated spheres overlap right now? (only decides which list the pair goes to)
                        const float Rtrue = __uint_as_float(myAux.y) + __uint_as_float(oa.y);
                        const bool touching = d2 < Rtrue * Rtrue;
                        if (!FILL) {
                            if (touching) nT++; else nN++;
                        } else {
                            const ContactList& L = touching ? P.ss : P.sn;
                            const uint32_t slot = touching ? slotT++ : slotN++;
                            if (slot < C.capacity) {
                                uint32_t skip = 0;
                                if (!touching) {
                                    // A candidate that is clearly apart at these very positions (float positions: allow
                                    // for their rounding) would have its history destroyed by the force pass that follows
                                    // this rebuild (no overlap => wildcards zeroed, DEMCalcForceKernels.cu:258-261): it
                                    // carries none over, so k_history has nothing to look up or write for it.
                                    const float slack = 3e-7f * (fabsf(me.x) + fabsf(me.y) + fabsf(me.z)) + 1e-8f;
                                    if (L.hist && sqrtf(d2) * 0.999999f - Rtrue * 1.000001f - slack > 0.f) skip = CINFO_NO_HISTORY;
                                }
                                const uint32_t matpair = (myMeta_z >> 16) * P.nMat + (om.z >> 16);
                                L.idB[slot] = om.y;
                                (touching ? C.idA_ss : C.idA_sn)[slot] = sid;
                                L.cinfo[slot] = make_uint4(myAux.x, om.x, (myMeta_z & 0xffffu) | ((om.z & 0xffffu) << 16), matpair | skip);
                            }
                        }
                    }
                }
            }
            __syncthreads();  // every lane is done with the staged data before the next phase overwrites it
#pragma unroll
            for (int r = 0; r < 5; r++) pos[r] += cnt[r];
        }
        if (!FILL && valid) {
            P.ss.seg_count[sid] = nT;
            P.sn.seg_count[sid] = nN;
        }
    }
}

// count -> (scan by the caller) -> fill
void launch_sweep_count(const DevParams& P, const CdParams& C, const uint32_t* keys, int grid, cudaStream_t s) {
    k_sweep_tma<false><<<grid, SW_THREADS, 0, s>>>(P, C, keys);
}
void launch_sweep_fill(const DevParams& P, const CdParams& C, const uint32_t* keys, int grid, cudaStream_t s) {
    k_sweep_tma<true><<<grid, SW_THREADS, 0, s>>>(P, C, keys);
}

}  // namespace demb
