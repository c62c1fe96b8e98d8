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
//     DEMContactKernels_SphereSphere.cu:121-126).  The distance tests run ONCE: the count pass hands the candidates
//     it accepted to the fill pass through a 48-byte record per sphere;
//   * the count pass is bound by the instruction issue rate, so its inner loop is kept convergent: a thread walks its
//     five runs as ONE loop (the warp pays max-over-lanes of the total, not the sum over runs of the per-run maxima)
//     that does nothing but the distance test and notes the hits in shared memory; the acceptance rule (owner, family
//     mask, extra margin, touching or not, first step due) runs in a second, short loop over the hits.
#include "dem_kernels.h"

namespace demb {

constexpr int SW_THREADS = 128;  // sorted positions per work item
constexpr int SW_CH = 160;       // staged entries per run and phase (even: the 8-byte stream stays 16-byte aligned)

constexpr int SW_HITS = 16;      // distance-test hits a thread notes before it stops to judge them

struct __align__(16) SweepSmem {
    float4 sph[5 * SW_CH];
    uint2 aux[5 * SW_CH];
    unsigned long long mbar;
    uint32_t rb[5], re[5];
    uint32_t pb[5];                          // sorted position of staged entry i of run r = i + pb[r] (i = flat index)
    uint32_t rng[5][SW_THREADS];             // per thread and run: its stretch as flat indices, first | end << 16
    unsigned short hit[SW_HITS][SW_THREADS]; // flat indices of the staged spheres that passed the distance test
};
static_assert(5 * SW_CH < 65536, "flat indices are 16 bits");

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

// Per sorted position j the count pass leaves the candidates it accepted in cand[j * SW_REC .. ): the candidate's sorted
// position | CAND_TOUCH | CAND_NOHIST.  The fill pass only has to look those up -- it never repeats a distance test.  A
// sphere with more than SW_K accepted candidates (dense polydisperse packings, many-component clumps) is rare; the fill
// pass finds its candidates again with plain loads (fill_slow).
constexpr uint32_t SW_K = 8;
constexpr uint32_t SW_REC = 12;  // words per record: SW_K candidates, SW_K "first step" bytes, nT | nN << 16, spare
constexpr uint32_t CAND_TOUCH = 0x80000000u;
constexpr uint32_t CAND_NOHIST = 0x40000000u;
constexpr uint32_t CAND_POS = 0x3fffffffu;

// the five forward runs [qb, qe) of the sphere at sorted position j in cell `key`
__device__ __forceinline__ void forward_runs(const CdParams& C, const GridInfo& g, uint32_t j, uint32_t key, uint32_t qb[5],
                                             uint32_t qe[5]) {
    const int cx = (int)(key % g.nbx);
    const int cy = (int)((key / g.nbx) % g.nby);
    const int cz = (int)(key / (g.nbx * g.nby));
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, (int)g.nbx - 1);
#pragma unroll
    for (int r = 0; r < 5; r++) {
        qb[r] = 0xffffffffu; qe[r] = 0u;
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

// Acceptance of the pair (me at j, ot at q): 0 = no candidate, else CAND_* flags | 1.  Superset of the double-precision
// test d2 <= R^2 && R - d > min(extraA, extraB) (DEMContactKernels_SphereSphere.cu:57-89).
__device__ __forceinline__ bool pair_near(const float4 me, const float4 ot, float& d2) {
    const float dx = me.x - ot.x, dy = me.y - ot.y, dz = me.z - ot.z;
    d2 = dx * dx + dy * dy + dz * dz;
    const float R = me.w + ot.w;
    return d2 <= R * R * 1.000001f + 1e-20f;
}
__device__ __forceinline__ uint32_t pair_verdict_near(const DevParams& P, const CdParams& C, const float4 me, const uint2 myAux,
                                                      const float4 ot, const uint2 oa, float d2, uint32_t q, bool fam_on,
                                                      uint32_t myFam, float extraA, bool want_hist, uint32_t& first) {
    first = 0u;
    if (oa.x == myAux.x) return 0u;  // same owner
    if (fam_on) {
        const uint32_t famB = __ldg(&C.sortedMeta[q]).w;
        if (C.any_mask && P.familyMasks[mask_pair(myFam, famB)] != 0) return 0u;
        const float Rt = me.w + ot.w - fminf(extraA, P.familyExtraMargin[famB]);
        if (d2 > Rt * Rt * 1.000001f + 1e-20f) return 0u;
    }
    // do the un-inflated spheres overlap right now? (only decides which list the pair goes to)
    const float Rtrue = __uint_as_float(myAux.y) + __uint_as_float(oa.y);
    if (d2 < Rtrue * Rtrue) return CAND_TOUCH | 1u;
    // A candidate that is clearly apart at these very positions (float positions: allow for their rounding) would have
    // its history destroyed by the force pass that follows this rebuild (no overlap => wildcards zeroed,
    // DEMCalcForceKernels.cu:258-261): it carries none over, so k_history has nothing to look up or write for it.
    const float slack = 3e-7f * (fabsf(me.x) + fabsf(me.y) + fabsf(me.z)) + 1e-8f;
    const float gap = sqrtf(d2) * 0.999999f - Rtrue * 1.000001f - slack;
    if (gap <= 0.f) return 1u;
    // The first step of the coming cycle at which this pair can possibly touch.  The margin of an owner IS the bound on
    // how far it travels during the maxDrift steps a list is used (DEMMiscKernels.cu:37-61), so after k integrations
    // the gap has closed by at most k (marginA + marginB) / maxDrift; until then the force kernel skips the pair after
    // reading its 16-byte record -- it cannot overlap, so it contributes nothing (margins include the family extra
    // margin, which only makes the bound more cautious).  A fixed expand factor carries no such promise: always test.
    if (P.beta < 0.f && (P.force_opts & 1u)) {
        const float closing = ((me.w - __uint_as_float(myAux.y)) + (ot.w - __uint_as_float(oa.y))) / (float)P.maxDrift;
        const float f = floorf(gap / fmaxf(closing, 1e-30f));
        first = (uint32_t)fminf(fmaxf(f, 0.f), 255.f);
    }
    return want_hist ? (CAND_NOHIST | 1u) : 1u;
}
__device__ __forceinline__ uint32_t pair_verdict(const DevParams& P, const CdParams& C, const float4 me, const uint2 myAux,
                                                 const float4 ot, const uint2 oa, uint32_t q, bool fam_on, uint32_t myFam,
                                                 float extraA, bool want_hist, uint32_t& first) {
    float d2;
    first = 0u;
    if (!pair_near(me, ot, d2)) return 0u;
    return pair_verdict_near(P, C, me, myAux, ot, oa, d2, q, fam_on, myFam, extraA, want_hist, first);
}

// COUNT pass: per-sphere counts into ss.seg_count / sn.seg_count, accepted candidates into C.cand.
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
    const bool want_hist = P.sn.hist != nullptr;
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
            forward_runs(C, g, j, keys[j], qb, qe);
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
        uint32_t nT = 0, nN = 0, myFam = 0;
        float extraA = 0.f;
        if (valid && fam_on) {
            myFam = C.sortedMeta[j].w;
            extraA = P.familyExtraMargin[myFam];
        }
        uint32_t* mycand = C.cand + (size_t)j * SW_REC;
        uint32_t first_lo = 0u, first_hi = 0u;  // the "first step" bytes of candidates 0-3 / 4-7
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
                for (int r = 0; r < 5; r++) {
                    if (cnt[r]) {
                        bulk_g2s(&sm.sph[r * SW_CH], C.sortedSph + pos[r], cnt[r] * 16u, &sm.mbar);
                        bulk_g2s(&sm.aux[r * SW_CH], C.sortedAux + pos[r], cnt[r] * 8u, &sm.mbar);
                    }
                    sm.pb[r] = pos[r] - (uint32_t)(r * SW_CH);
                }
            }
            // my stretch of every run in this phase, as flat indices into the staged arrays
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t lo = max(qb[r], pos[r]);
                const uint32_t hi = min(qe[r], pos[r] + cnt[r]);
                uint32_t w = 0u;
                if (valid && lo < hi) w = ((uint32_t)(r * SW_CH) + lo - pos[r]) | (((uint32_t)(r * SW_CH) + hi - pos[r]) << 16);
                sm.rng[r][tid] = w;
            }
            // one thread waits for the bulk copies (try_wait spins: 127 other threads would only burn issue slots),
            // the CTA barrier releases the rest
            if (tid == 0) mbar_wait(&sm.mbar, parity);
            parity ^= 1u;
            __syncthreads();
            {
                int r = -1;
                uint32_t a = 0u, b = 0u;
                bool done = false;
                while (!done) {
                    // (1) distance tests only, all five runs as one loop
                    uint32_t nh = 0u;
                    for (;;) {
                        if (a >= b) {
                            do {
                                r++;
                                if (r < 5) {
                                    const uint32_t w = sm.rng[r][tid];
                                    a = w & 0xffffu;
                                    b = w >> 16;
                                }
                            } while (r < 5 && a >= b);
                            if (r >= 5) { done = true; break; }
                        }
                        float d2;
                        if (pair_near(me, sm.sph[a], d2)) {
                            sm.hit[nh][tid] = (unsigned short)a;
                            nh++;
                        }
                        a++;
                        if (nh == (uint32_t)SW_HITS) break;
                    }
                    // (2) the acceptance rule for the hits
                    for (uint32_t k = 0; k < nh; k++) {
                        const uint32_t i = sm.hit[k][tid];
                        const uint32_t q = i + sm.pb[i / (uint32_t)SW_CH];
                        const float4 ot = sm.sph[i];
                        float d2;
                        pair_near(me, ot, d2);
                        uint32_t first;
                        const uint32_t v = pair_verdict_near(P, C, me, myAux, ot, sm.aux[i], d2, q, fam_on, myFam, extraA, want_hist, first);
                        if (v == 0u) continue;
                        const uint32_t kk = nT + nN;
                        if (kk < SW_K) {
                            mycand[kk] = q | (v & ~1u);
                            if (kk < 4u) first_lo |= first << (8u * kk); else first_hi |= first << (8u * (kk - 4u));
                        }
                        if (v & CAND_TOUCH) nT++; else nN++;
                    }
                }
            }
            __syncthreads();  // every lane is done with the staged data before the next phase overwrites it
#pragma unroll
            for (int r = 0; r < 5; r++) pos[r] += cnt[r];
        }
        if (valid) {
            const uint32_t sid = C.sortedMeta[j].y;
            P.ss.seg_count[sid] = nT;
            P.sn.seg_count[sid] = nN;
            reinterpret_cast<uint4*>(mycand)[2] = make_uint4(first_lo, first_hi, min(nT, 0xffffu) | (min(nN, 0xffffu) << 16), 0u);
        }
    }
}

// write contact (A at sorted position j, B described by om = sortedMeta[q]) of the new lists
__device__ __forceinline__ void emit_contact(const DevParams& P, const CdParams& C, uint32_t flags, uint32_t first, const uint4 om,
                                             uint32_t sid, uint32_t ownerA, uint32_t metaA_z, uint32_t& slotT, uint32_t& slotN) {
    const bool touching = (flags & CAND_TOUCH) != 0u;
    const ContactList& L = touching ? P.ss : P.sn;
    const uint32_t slot = touching ? slotT++ : slotN++;
    if (slot >= C.capacity) return;
    const uint32_t matpair = (metaA_z >> 16) * P.nMat + (om.z >> 16);
    L.idB[slot] = om.y;
    (touching ? C.idA_ss : C.idA_sn)[slot] = sid;
    L.cinfo[slot] = make_uint4(ownerA, om.x, (metaA_z & 0xffffu) | ((om.z & 0xffffu) << 16),
                               matpair | ((flags & CAND_NOHIST) ? CINFO_NO_HISTORY : 0u));
    if (!touching) L.due[slot] = (uint8_t)first;
}

// FILL pass: one thread per sorted position; the compiled records go to seg_start[sphere] + k (seg_start = exclusive scan
// of the counts over sphere ids: the lists are sphere-major, hence owner-major, and deterministic).
__global__ void __launch_bounds__(256) k_sweep_fill(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C,
                                                    const uint32_t* __restrict__ keys) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const GridInfo g = *C.grid;
    const uint32_t nSorted = C.cellStart[g.ncells];
    const bool fam_on = (C.any_mask != 0) || (C.max_extra > 0.f);
    const bool want_hist = P.sn.hist != nullptr;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nSorted; j += gridDim.x * blockDim.x) {
        const uint4* cp = reinterpret_cast<const uint4*>(C.cand + (size_t)j * SW_REC);
        const uint4 tail = cp[2];  // {first bytes 0-3, first bytes 4-7, nT | nN << 16, -}
        const uint32_t n = (tail.z & 0xffffu) + (tail.z >> 16);
        if (n == 0u) continue;
        const uint4 meta = C.sortedMeta[j];
        const uint32_t sid = meta.y;
        uint32_t slotT = P.ss.seg_start[sid], slotN = P.sn.seg_start[sid];
        if (n <= SW_K) {
            const uint32_t fw[2] = {tail.x, tail.y};
#pragma unroll
            for (uint32_t k0 = 0; k0 < SW_K; k0 += 4) {
                if (k0 >= n) break;
                const uint4 c4 = cp[k0 >> 2];
                const uint32_t c[4] = {c4.x, c4.y, c4.z, c4.w};
                const uint32_t f = fw[k0 >> 2];
                // (the four gathers are in flight together: the loop is otherwise a chain of dependent loads per thread)
                uint4 om[4];
#pragma unroll
                for (int t = 0; t < 4; t++) om[t] = __ldg(&C.sortedMeta[(k0 + t < n) ? (c[t] & CAND_POS) : j]);
#pragma unroll
                for (int t = 0; t < 4; t++)
                    if (k0 + t < n) emit_contact(P, C, c[t], (f >> (8 * t)) & 0xffu, om[t], sid, meta.x, meta.z, slotT, slotN);
            }
        } else {
            // more accepted candidates than the hand-over buffer holds: find them again (plain loads)
            const float4 me = C.sortedSph[j];
            const uint2 myAux = C.sortedAux[j];
            const float extraA = fam_on ? P.familyExtraMargin[meta.w] : 0.f;
            uint32_t qb[5], qe[5];
            forward_runs(C, g, j, keys[j], qb, qe);
            for (int r = 0; r < 5; r++)
                for (uint32_t q = qb[r]; q < qe[r]; q++) {
                    uint32_t first;
                    const uint32_t v = pair_verdict(P, C, me, myAux, __ldg(&C.sortedSph[q]), __ldg(&C.sortedAux[q]), q, fam_on,
                                                    meta.w, extraA, want_hist, first);
                    if (v) emit_contact(P, C, v, first, __ldg(&C.sortedMeta[q]), sid, meta.x, meta.z, slotT, slotN);
                }
        }
    }
}

// count -> (scan by the caller) -> fill
void launch_sweep_count(const DevParams& P, const CdParams& C, const uint32_t* keys, int grid, cudaStream_t s) {
    k_sweep_tma<<<grid, SW_THREADS, 0, s>>>(P, C, keys);
}
void launch_sweep_fill(const DevParams& P, const CdParams& C, const uint32_t* keys, int grid, cudaStream_t s) {
    k_sweep_fill<<<grid, 256, 0, s>>>(P, C, keys);
}

}  // namespace demb
