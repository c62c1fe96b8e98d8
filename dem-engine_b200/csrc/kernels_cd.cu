// kernels_cd.cu -- contact-list rebuild ("kinematic" work) for sm_100a, fully device-driven: no host round trip between
// the stages and none at the end -- every count, the grid and every flag stay on the device; the last kernel leaves a
// status record in pinned host memory that the host reads once it knows the rebuild to be complete.  All kernels are
// grid-stride over DEVICE-resident counts, so the same launch parameters -- hence the same CUDA graph -- serve every
// rebuild.  A rebuild that cannot hold its result (a list or table overflowed) sets the poison word: every later kernel
// of this context returns at once, so the state freezes at the failed rebuild until the host has grown the arrays.
//   k_maxvel / k_grid_setup      max |v| (all-gathered over the ranks) -> margin -> broad-phase cell size and grid
//   k_anal_prep                  resolves the analytical components (planes, cylinders) to world space once
//   k_sphere_prep                per sphere: world position (fixed-point decode + rotated offset), inflated radius,
//                                cell key + cell histogram; emits the sphere--analytical list (+history) directly
//                                with warp-aggregated atomics
//   k_scan_lookback              single-pass exclusive prefix sums with decoupled look-back (cell table, per-sphere
//                                contact offsets)
//   k_cs_scatter / k_gather_sorted_cs   counting sort into (cell, sphere id) order; radix sort (k_rs_*) as alternative
//   k_sweep_tma<COUNT|FILL>      kernels_sweep.cu: TMA-staged pair sweep over the upper half of the 27-cell stencil
//   k_history                    Hertz-Mindlin history carried over from the previous lists
//   k_tri_cells / k_st_emit      sphere--triangle broad phase
//   k_finish_counts              clamps, overflow -> poison (agreed over the ranks), status record
// Reference behaviour being reproduced: contactDetection(), src/algorithms/DEMCubContactDetection.cu:38-1123;
// acceptance rule of src/kernel/DEMContactKernels_SphereSphere.cu:57-89,172-214 and DEMBinSphereKernels.cu:78-128;
// margin of src/kernel/DEMMiscKernels.cu:37-69; history map of src/kernel/DEMHistoryMappingKernels.cu.
// The candidate list is a SUPERSET of the reference's (the force kernel re-tests true overlap), so physics is identical.
#include <algorithm>
#include <cstring>

#include "dem_kernels.h"

namespace demb {

// ---------------------------------------------------------------------------------------------------------------
// margin of one owner, computeMarginFromAbsv / fillMarginValues (DEMMiscKernels.cu:37-69)
__device__ __forceinline__ float owner_margin(const DevParams& P, float absv, uint32_t family) {
    const float extra = P.familyExtraMargin[family];
    if (P.beta >= 0.f) return P.beta + extra;
    if (absv > P.approxMaxVel) absv = P.approxMaxVel;
    return (float)((double)(absv * P.expSafetyMulti + P.expSafetyAdder) * (double)P.h * (double)P.maxDrift +
                   (double)extra);
}

__global__ void k_maxvel(const __grid_constant__ DevParams P, float errOutVel) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t n = P.active_list ? *P.nActivePtr : P.nOwners;
    float a = 0.f;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t o = P.active_list ? P.active_list[t] : t;
        const float4 v = P.state[o].vel;
        float b = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
        if (!isfinite(b) || b > errOutVel) atomicOr(&P.flags[DEM_FLAG_VELOCITY], 1u);
        if (!isfinite(b)) b = 0.f;
        a = fmaxf(a, b);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, off));
    if ((threadIdx.x & 31) == 0 && a > 0.f) atomicMax(reinterpret_cast<int*>(P.maxvel), __float_as_int(a));
}

// one warp: all-gather of max |v| over the ranks (the cell grid, hence the sorted order and the A/B roles of a pair,
// must be the same on every rank), then lane 0 decides the grid
__global__ void __launch_bounds__(32) k_grid_setup(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C,
                                                   const __grid_constant__ MgDev M) {
    if (P.flags[DEM_FLAG_POISON]) return;
    float vmax = *P.maxvel;
    if (M.world > 1) {
        uint32_t o0, o1;
        mg_allgather(M, __float_as_uint(vmax), 0u, o0, o1, P.flags);
        vmax = __uint_as_float(o0);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
        if (threadIdx.x == 0) *P.maxvel = vmax;
    }
    if (threadIdx.x != 0) return;
    float margin;
    if (P.beta >= 0.f) {
        margin = P.beta + C.max_extra;
    } else {
        const float a = fminf(vmax, P.approxMaxVel);
        margin = (float)((double)(a * P.expSafetyMulti + P.expSafetyAdder) * (double)P.h * (double)P.maxDrift +
                         (double)C.max_extra);
    }
    float cs = 2.f * (C.rmax + margin) * 1.0005f + 1e-30f;
    uint32_t nbx, nby, nbz;
    for (;;) {
        nbx = (uint32_t)fmaxf(1.f, ceilf(C.ext[0] / cs));
        nby = (uint32_t)fmaxf(1.f, ceilf(C.ext[1] / cs));
        nbz = (uint32_t)fmaxf(1.f, ceilf(C.ext[2] / cs));
        if ((double)nbx * (double)nby * (double)nbz <= (double)C.max_cells) break;
        cs *= 1.1f;
    }
    GridInfo g;
    g.cs = cs;
    g.inv_cs = 1.f / cs;
    g.nbx = nbx; g.nby = nby; g.nbz = nbz;
    g.ncells = nbx * nby * nbz;
    g.max_margin = margin;
    g.maxvel = vmax;
    // the most by which the gap of any pair can shrink in one step is 2 margin / maxDrift (k_force_ss leaves candidates
    // alone until they can touch); a fixed expand factor promises no such bound
    P.flags[DEM_FLAG_INV_CLOSING] = (P.beta < 0.f && margin > 0.f) ? __float_as_uint((float)P.maxDrift / (2.f * margin)) : 0u;
    // ghost layer: two clumps can touch up to 2 (R_clump + margin) apart; one more margin on each side covers the
    // distance an owner can travel before the next rebuild (that bound is what the margin is made of)
    g.halo = 2.f * (C.rclump + margin) + 2.f * margin;
    g.x0 = 0;
    if (C.slab_on) {
        // This rank only bins the cell columns its active spheres can fall into: the slab, the ghost layer on either
        // side, the reach of a ghost clump's spheres beyond its centre, and one spare column.  Same cell size and
        // origin on every rank, so the (cell, sphere id) order of any two spheres -- hence their A/B roles -- is the
        // same wherever the pair is evaluated; spheres outside the range are clamped into the edge columns, which
        // keeps true neighbours within one column of each other.
        const float reach = g.halo + C.rclump + margin + cs;
        const int lo = max(0, (int)floorf((C.slab_lo - reach) * g.inv_cs));
        const int hi = min((int)nbx - 1, (int)floorf((C.slab_hi + reach) * g.inv_cs));
        g.x0 = lo;
        g.nbx = (uint32_t)max(1, hi - lo + 1);
        g.ncells = g.nbx * nby * nbz;
    }
    *C.grid = g;
}

// World-space analytical components of this rebuild (plane point / cylinder centre, direction, owner margin, family)
__global__ void k_anal_prep(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.nAnal) return;
    const AnalObj ob = P.anal[k];
    const OwnerState* sb = P.state + ob.owner;
    const OwnerPos pB = sb->pos;
    const float4 qB = sb->quat;
    const float4 vB = sb->vel;
    double X, Y, Z;
    pos_decode(pB, P, X, Y, Z);
    const float3 rel = rotate(f3(ob.relx, ob.rely, ob.relz), qB);
    const float3 dir = rotate(f3(ob.rotx, ob.roty, ob.rotz), qB);
    AnalWorld w;
    w.px = (float)(X + (double)rel.x); w.py = (float)(Y + (double)rel.y); w.pz = (float)(Z + (double)rel.z);
    w.dx = dir.x; w.dy = dir.y; w.dz = dir.z;
    w.margin = owner_margin(P, sqrtf(vB.x * vB.x + vB.y * vB.y + vB.z * vB.z), pB.family);
    w.size1 = ob.size1;
    w.normal_sign = ob.normal_sign;
    w.type = ob.type;
    w.family = pB.family;
    w.material = ob.material;
    C.analw[k] = w;
}

// sphere--analytical candidate test with inflated geometry (DEMBinSphereKernels.cu:78-128). Conservative in float.
__device__ __forceinline__ bool sa_candidate(const DevParams& P, const AnalWorld& w, float3 sp /*LBF-rel*/, float rInfl,
                                             uint32_t famS, bool any_mask) {
    if (any_mask && P.familyMasks[mask_pair(famS, w.family)] != 0) return false;
    const float3 d = f3(sp.x - w.px, sp.y - w.py, sp.z - w.pz);
    const float3 dir = f3(w.dx, w.dy, w.dz);
    const float thr = fminf(P.familyExtraMargin[famS], P.familyExtraMargin[w.family]);
    const float slack = 1e-6f * (fabsf(sp.x) + fabsf(sp.y) + fabsf(sp.z) + 1.f);
    float depth;
    if (w.type == DEM_ANAL_PLANE) {
        depth = rInfl + w.margin - dot(d, dir);
    } else if (w.type == DEM_ANAL_CYL_INF) {
        const float3 s2c = f3(-d.x, -d.y, -d.z);
        const float proj = dot(s2c, dir);
        const float3 radial = s2c - proj * dir;
        const float cyl_rad = w.size1 - w.normal_sign * w.margin;
        depth = rInfl - w.normal_sign * (cyl_rad - length(radial));
    } else {
        return false;
    }
    return depth + slack > thr;
}

// exclusive warp scan of a count + one atomic per warp on a global cursor: returns this lane's first slot
__device__ __forceinline__ uint32_t warp_claim(uint32_t count, uint32_t* cursor) {
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

__global__ void __launch_bounds__(256) k_sphere_prep(const __grid_constant__ DevParams P,
                                                     const __grid_constant__ CdParams C) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t n = C.act_sph ? *C.act_count : P.nSpheres;
    const uint32_t nround = (n + 31u) & ~31u;  // whole warps take part in the slot claim
    const GridInfo g = *C.grid;
    for (uint32_t t_ = blockIdx.x * blockDim.x + threadIdx.x; t_ < nround; t_ += gridDim.x * blockDim.x) {
        const bool valid = t_ < n;
        const uint32_t i = (valid && C.act_sph) ? C.act_sph[t_] : t_;
        uint32_t nsa = 0, samask = 0;
        float3 sp = f3(0.f, 0.f, 0.f);
        uint2 s = make_uint2(0, 0);
        uint32_t family = 0;
        float rInfl = 0.f;
        if (valid) {
            s = P.sph[i];
            OwnerPos pos;
            float4 q, v;
            {
                const float* base = reinterpret_cast<const float*>(P.state + s.x);
                uint32_t a0, a1, a2, a3;
                asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                             : "l"(base));
                pos.voxel = ((unsigned long long)a1 << 32) | a0;
                pos.lx = (unsigned short)(a2 & 0xffffu); pos.ly = (unsigned short)(a2 >> 16);
                pos.lz = (unsigned short)(a3 & 0xffffu); pos.family = (unsigned char)((a3 >> 16) & 0xffu);
                pos.flags = 0;
                v = __ldg(&P.state[s.x].vel);
            }
            family = pos.family;
            const float4 comp = __ldg(&P.comp[s.y & 0xffffu]);
            const float margin = owner_margin(P, sqrtf(v.x * v.x + v.y * v.y + v.z * v.z), pos.family);
            double X, Y, Z;
            pos_decode(pos, P, X, Y, Z);
            const float3 rel = rotate(f3(comp.x, comp.y, comp.z), q);
            sp = f3((float)(X + (double)rel.x), (float)(Y + (double)rel.y), (float)(Z + (double)rel.z));
            rInfl = comp.w + margin;
            if (P.beta < 0.f) {
                // the owner's margin relative to the largest one, as a byte in its state record (dem_device.cuh): the force
                // kernel turns the two codes of a candidate pair into the bound on how fast their gap can close.  Every
                // sphere of the owner stores the same value; nobody in this kernel looks at the byte.
                uint32_t code = 255u;
                if (g.max_margin > 0.f) code = min(255u, (uint32_t)(256.f * (margin / g.max_margin) * 1.000001f));  // (c + 1) / 256 >= m / max
                reinterpret_cast<unsigned char*>(P.state + s.x)[15] = (unsigned char)(255u - code);
            }
            C.sphF[i] = make_float4(sp.x, sp.y, sp.z, rInfl);
            int cx = (int)floorf(sp.x * g.inv_cs) - g.x0, cy = (int)floorf(sp.y * g.inv_cs), cz = (int)floorf(sp.z * g.inv_cs);
            cx = min(max(cx, 0), (int)g.nbx - 1);
            cy = min(max(cy, 0), (int)g.nby - 1);
            cz = min(max(cz, 0), (int)g.nbz - 1);
            const uint32_t key = (uint32_t)cx + g.nbx * ((uint32_t)cy + g.nby * (uint32_t)cz);
            C.keys[0][i] = key;
            C.vals[0][i] = i;
            // cell histogram; the arrival rank doubles as the slot of the counting sort (made deterministic afterwards)
            C.vals[1][i] = atomicAdd(&C.cellStart[key], 1u);
            // analytical candidates (at most 32 components are tracked per sphere in the bit mask; more fall back below)
            for (uint32_t k = 0; k < P.nAnal; k++)
                if (sa_candidate(P, C.analw[k], sp, rInfl, family, C.any_mask != 0)) {
                    if (k < 32) samask |= 1u << k;
                    nsa++;
                }
        }
        if (P.nAnal == 0) continue;
        // ---- emit the sphere--analytical contacts of this warp into one contiguous run ----
        uint32_t slot = warp_claim(nsa, P.sa.count);
        if (!valid) continue;
        P.sa.seg_start[i] = slot;
        P.sa.seg_count[i] = (slot + nsa <= C.capacity) ? nsa : (slot < C.capacity ? C.capacity - slot : 0u);
        if (nsa == 0) continue;
        if (slot + nsa > C.capacity) atomicOr(&P.flags[DEM_FLAG_CAPACITY], 2u);
        const uint32_t oldStart = C.oldsa.seg_start[i], oldCount = C.oldsa.seg_count[i];
        for (uint32_t k = 0; k < P.nAnal; k++) {
            bool hit;
            if (k < 32) hit = (samask >> k) & 1u;
            else hit = sa_candidate(P, C.analw[k], sp, rInfl, family, C.any_mask != 0);
            if (!hit) continue;
            if (slot < C.capacity) {
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t alive = 0;
                for (uint32_t t = 0; t < oldCount; t++) {
                    if (C.oldsa.idB[oldStart + t] == k) {
                        alive = C.oldsa.cinfo[oldStart + t].w & 0x80000000u;
                        if (alive && C.oldsa.hist) h = C.oldsa.hist[oldStart + t];
                        break;
                    }
                }
                const uint32_t matpair = (s.y >> 16) * P.nMat + C.analw[k].material;
                P.sa.idB[slot] = k;
                P.sa.cinfo[slot] = make_uint4(s.x, k, s.y & 0xffffu, matpair | alive);
                if (P.sa.hist) P.sa.hist[slot] = h;
            }
            slot++;
        }
    }
}

// zero a u32 range unless the context is poisoned (a memset node cannot be made conditional)
__global__ void __launch_bounds__(256) k_zero_u32(uint32_t* __restrict__ p, size_t n, const uint32_t* __restrict__ flags) {
    if (flags[DEM_FLAG_POISON]) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}
int launch_zero_u32(uint32_t* p, size_t n, const uint32_t* flags, int num_sms, cudaStream_t s) {
    if (n == 0) return 0;
    const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)num_sms * 8);
    k_zero_u32<<<grid, 256, 0, s>>>(p, n, flags);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan (u32): block sums -> top-level scan -> apply
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem /*>=32*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        uint32_t w = (lane < nw) ? smem[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += t;
        }
        smem[lane] = winc - w;          // exclusive warp offsets
        if (lane == 31) smem[32] = winc;  // total
    }
    __syncthreads();
    total = smem[32];
    const uint32_t r = smem[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_sums(const uint32_t* __restrict__ in, uint32_t n,
                                                          uint32_t* __restrict__ sums) {
    __shared__ uint32_t sm[33];
    const uint32_t base = blockIdx.x * SC_TILE;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        const uint32_t idx = base + k * SC_THREADS + threadIdx.x;
        if (idx < n) acc += in[idx];
    }
    uint32_t total;
    block_exclusive_scan(acc, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_top(uint32_t* __restrict__ sums, uint32_t nblk,
                                                   uint32_t* __restrict__ total_out) {
    __shared__ uint32_t sm[33];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblk; base += 1024) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = (idx < nblk) ? sums[idx] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, sm, total);
        if (idx < nblk) sums[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_apply(uint32_t* __restrict__ data, uint32_t n,
                                                           const uint32_t* __restrict__ sums) {
    __shared__ uint32_t sm[33];
    // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS];
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        v[k] = (base + k < n) ? data[base + k] : 0u;
        acc += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(acc, sm, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        if (base + k < n) data[base + k] = ex;
        ex += v[k];
    }
}

// in-place exclusive scan of data[0..n); *total receives the sum (may be nullptr). tmp holds ceil(n/4096) words.
int launch_scan_exclusive(uint32_t* data, uint32_t n, uint32_t* tmp, uint32_t* total, cudaStream_t s) {
    if (n == 0) return 0;
    const uint32_t nblk = (n + SC_TILE - 1) / SC_TILE;
    k_scan_sums<<<nblk, SC_THREADS, 0, s>>>(data, n, tmp);
    k_scan_top<<<1, 1024, 0, s>>>(tmp, nblk, total);
    k_scan_apply<<<nblk, SC_THREADS, 0, s>>>(data, n, tmp);
    return 3;
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass exclusive scan with decoupled look-back (Merrill & Garland): tiles are handed out by an atomic counter
// (so a tile's predecessors are always already running), each tile publishes its aggregate, resolves its prefix from
// the descriptors behind it and publishes the inclusive prefix.  One pass over the data instead of the three of the
// sums / top / apply scheme; the length is read from device memory.
//   descriptor = value | state << 32   (state 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix)
constexpr int LB_THREADS = 256;
constexpr int LB_ITEMS = 16;
constexpr int LB_TILE = LB_THREADS * LB_ITEMS;

// idx != nullptr: element i of the scan is in[idx[i]] and its prefix goes to out[idx[i]] (the decomposed rebuild scans the
// per-sphere counts in the order of its active-sphere list instead of walking all spheres of the system).
__global__ void __launch_bounds__(LB_THREADS) k_scan_lookback(const uint32_t* in, uint32_t* out,
                                                              const uint32_t* __restrict__ n_ptr, uint32_t n_add,
                                                              unsigned long long* desc, uint32_t* total_out,
                                                              const uint32_t* __restrict__ flags,
                                                              const uint32_t* __restrict__ idx) {
    if (flags[DEM_FLAG_POISON]) return;
    __shared__ uint32_t sm[33];
    __shared__ uint32_t s_tile, s_prefix;
    const uint32_t n = (n_ptr ? *n_ptr : 0u) + n_add;
    const uint32_t ntiles = (n + LB_TILE - 1) / LB_TILE;
    uint32_t* tile_ctr = reinterpret_cast<uint32_t*>(desc);
    volatile unsigned long long* D = desc + 1;
    if (n == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && total_out) *total_out = 0u;
        return;
    }
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_ctr, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint32_t base = tile * LB_TILE + threadIdx.x * LB_ITEMS;
        uint32_t v[LB_ITEMS];
        uint32_t acc = 0;
        uint32_t id[LB_ITEMS];
        if (idx) {
#pragma unroll
            for (int k = 0; k < LB_ITEMS; k++) {
                id[k] = (base + k < n) ? idx[base + k] : 0xffffffffu;
                v[k] = (base + k < n) ? in[id[k]] : 0u;
            }
        } else if (base + LB_ITEMS <= n) {
            const uint4* src = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
            for (int k = 0; k < LB_ITEMS / 4; k++) {
                const uint4 q = src[k];
                v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < LB_ITEMS; k++) v[k] = (base + k < n) ? in[base + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < LB_ITEMS; k++) acc += v[k];
        uint32_t total;
        uint32_t ex = block_exclusive_scan(acc, sm, total);
        if (threadIdx.x < 32) {
            // look-back by one WARP, 32 descriptors at a time (a single thread walking back one descriptor per L2 round
            // trip made the chain of tiles the critical path: 24 us for 1200 tiles)
            const int lane = threadIdx.x;
            uint32_t prefix = 0;
            if (tile == 0) {
                if (lane == 0) D[0] = (unsigned long long)total | (2ull << 32);
            } else {
                if (lane == 0) {
                    D[tile] = (unsigned long long)total | (1ull << 32);
                    __threadfence();
                }
                for (int p = (int)tile - 1;; p -= 32) {
                    const int q = p - lane;  // lane l looks at the l-th descriptor behind
                    unsigned long long d = 2ull << 32;  // before the first tile: an inclusive prefix of zero
                    if (q >= 0) {
                        do { d = D[q]; } while ((d >> 32) == 0ull);
                    }
                    const uint32_t incl = __ballot_sync(0xffffffffu, (d >> 32) == 2ull);
                    const int first = __ffs(incl) - 1;  // the nearest inclusive prefix, -1 = none among these 32
                    uint32_t v = (first < 0 || lane <= first) ? (uint32_t)d : 0u;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    prefix += v;
                    if (first >= 0) break;
                }
                if (lane == 0) D[tile] = (unsigned long long)(prefix + total) | (2ull << 32);
            }
            if (lane == 0) {
                s_prefix = prefix;
                if (tile == ntiles - 1 && total_out) *total_out = prefix + total;
            }
        }
        __syncthreads();
        ex += s_prefix;
        if (idx) {
#pragma unroll
            for (int k = 0; k < LB_ITEMS; k++) {
                if (base + k < n) out[id[k]] = ex;
                ex += v[k];
            }
        } else if (base + LB_ITEMS <= n) {
            uint4* dst = reinterpret_cast<uint4*>(out + base);
#pragma unroll
            for (int k = 0; k < LB_ITEMS / 4; k++) {
                uint4 q;
                q.x = ex; ex += v[4 * k];
                q.y = ex; ex += v[4 * k + 1];
                q.z = ex; ex += v[4 * k + 2];
                q.w = ex; ex += v[4 * k + 3];
                dst[k] = q;
            }
        } else {
#pragma unroll
            for (int k = 0; k < LB_ITEMS; k++) {
                if (base + k < n) out[base + k] = ex;
                ex += v[k];
            }
        }
        __syncthreads();  // s_tile / s_prefix are rewritten by the next round
    }
}

int launch_scan_lookback(const uint32_t* in, uint32_t* out, const uint32_t* n_ptr, uint32_t n_add, uint32_t n_max,
                         unsigned long long* desc, uint32_t* total, const uint32_t* flags, int num_sms, cudaStream_t s,
                         const uint32_t* idx) {
    const uint32_t max_tiles = (n_max + LB_TILE - 1) / LB_TILE;
    cudaMemsetAsync(desc, 0, sizeof(unsigned long long) * ((size_t)max_tiles + 2), s);
    const int grid = (int)std::max<uint32_t>(1u, std::min<uint32_t>(max_tiles, (uint32_t)num_sms * 4u));
    k_scan_lookback<<<grid, LB_THREADS, 0, s>>>(in, out, n_ptr, n_add, desc, total, flags, idx);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// LSD radix sort of (cell key, sphere index) pairs, 8 bits per pass.
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift,
                                                        uint32_t* __restrict__ hist, uint32_t nblk) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t idx = base + k * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&sh[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblk + blockIdx.x] = sh[threadIdx.x];  // digit-major so one scan yields global offsets
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t* __restrict__ keys_in,
                                                           const uint32_t* __restrict__ vals_in,
                                                           uint32_t* __restrict__ keys_out,
                                                           uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                           const uint32_t* __restrict__ hist, uint32_t nblk) {
    __shared__ uint32_t warpHist[RS_WARPS][256];
    __shared__ uint32_t tileKeys[RS_TILE];
    __shared__ uint32_t tileVals[RS_TILE];
    __shared__ uint32_t binBase[256];
    __shared__ uint32_t globalBase[256];
    __shared__ uint32_t scanTmp[33];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int w = 0; w < RS_WARPS; w++) warpHist[w][t] = 0;
    __syncthreads();
    const uint32_t tileBase = blockIdx.x * RS_TILE;
    const uint32_t warpBase = tileBase + warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t idx = warpBase + k * 32 + lane;
        const bool valid = idx < n;
        key[k] = valid ? keys_in[idx] : 0xffffffffu;
        val[k] = valid ? vals_in[idx] : 0u;
        const uint32_t d = (key[k] >> shift) & 255u;
        const uint32_t mv = valid ? d : (0x100u | (uint32_t)lane);  // invalid lanes never match anyone
        const uint32_t peers = __match_any_sync(0xffffffffu, mv);
        const uint32_t r = __popc(peers & lt);
        const uint32_t prior = valid ? warpHist[warp][d] : 0u;
        __syncwarp();
        if (valid && r == 0) warpHist[warp][d] = prior + __popc(peers);
        __syncwarp();
        rank[k] = prior + r;
    }
    __syncthreads();
    // per digit: exclusive offsets of the warps, then block-exclusive offsets of the digits
    uint32_t running = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) {
        const uint32_t c = warpHist[w][t];
        warpHist[w][t] = running;
        running += c;
    }
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(running, scanTmp, total);
    binBase[t] = ex;
    globalBase[t] = hist[t * nblk + blockIdx.x];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t idx = warpBase + k * 32 + lane;
        if (idx < n) {
            const uint32_t d = (key[k] >> shift) & 255u;
            const uint32_t p = binBase[d] + warpHist[warp][d] + rank[k];
            tileKeys[p] = key[k];
            tileVals[p] = val[k];
        }
    }
    __syncthreads();
    const uint32_t tileCount = min((uint32_t)RS_TILE, n - tileBase);
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t p = k * RS_THREADS + t;
        if (p < tileCount) {
            const uint32_t kk = tileKeys[p];
            const uint32_t d = (kk >> shift) & 255u;
            const uint32_t dst = globalBase[d] + (p - binBase[d]);
            keys_out[dst] = kk;
            vals_out[dst] = tileVals[p];
        }
    }
}

int launch_cd_sort(const DevParams& P, const CdParams& C, int key_bits, cudaStream_t s, int* out_buf) {
    const uint32_t n = P.nSpheres;
    int launches = 0, cur = 0;
    if (n == 0) { *out_buf = 0; return 0; }
    const uint32_t nblk = (n + RS_TILE - 1) / RS_TILE;
    for (int shift = 0; shift < key_bits; shift += 8) {
        k_rs_hist<<<nblk, RS_THREADS, 0, s>>>(C.keys[cur], n, shift, C.rs_hist, nblk);
        launches += 1 + launch_scan_exclusive(C.rs_hist, 256u * nblk, C.scan_tmp, nullptr, s);
        k_rs_scatter<<<nblk, RS_THREADS, 0, s>>>(C.keys[cur], C.vals[cur], C.keys[cur ^ 1], C.vals[cur ^ 1], n, shift,
                                                 C.rs_hist, nblk);
        launches += 1;
        cur ^= 1;
    }
    *out_buf = cur;
    return launches;
}

// ---------------------------------------------------------------------------------------------------------------
// radix-sort path (single GPU, sort_mode 0): gather into the sorted order the sort produced
__global__ void __launch_bounds__(256) k_gather_sorted(const __grid_constant__ DevParams P,
                                                       const __grid_constant__ CdParams C,
                                                       const uint32_t* __restrict__ vals) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.nSpheres) return;
    const uint32_t i = vals[j];
    C.sortedSph[j] = C.sphF[i];
    const uint2 s = P.sph[i];
    uint32_t fam = 0;
    if (C.any_mask != 0 || C.max_extra > 0.f) fam = P.state[s.x].pos.family;
    C.sortedMeta[j] = make_uint4(s.x, i, s.y, fam);  // {owner, sphere id, comp | material<<16, family}
    C.sortedAux[j] = make_uint2(s.x, __float_as_uint(__ldg(&P.comp[s.y & 0xffffu]).w));
}

// ---- counting sort by cell (sort_mode 1): the histogram and its prefix exist anyway for the sweep ----
__global__ void __launch_bounds__(256) k_cs_scatter(const __grid_constant__ DevParams P,
                                                    const __grid_constant__ CdParams C) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t n = C.act_sph ? *C.act_count : P.nSpheres;
    for (uint32_t t_ = blockIdx.x * blockDim.x + threadIdx.x; t_ < n; t_ += gridDim.x * blockDim.x) {
        const uint32_t i = C.act_sph ? C.act_sph[t_] : t_;
        const uint32_t key = C.keys[0][i];
        C.keys[1][C.cellStart[key] + C.vals[1][i]] = i;  // arrival order inside the cell (non-deterministic)
    }
}

// gather into cell order; inside a cell the spheres are ranked by sphere id, which makes the result identical to a
// stable radix sort of (cell key, sphere id) no matter in which order the atomics arrived
__global__ void __launch_bounds__(256) k_gather_sorted_cs(const __grid_constant__ DevParams P,
                                                          const __grid_constant__ CdParams C) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t n = C.cellStart[C.grid->ncells];  // number of active spheres (== nSpheres on a single GPU)
    const bool need_fam = C.any_mask != 0 || C.max_extra > 0.f;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = C.keys[1][j];
        const uint32_t key = C.keys[0][i];
        const uint32_t sb = C.cellStart[key], se = C.cellStart[key + 1];
        uint32_t rank = 0;
        for (uint32_t t = sb; t < se; t++) rank += (C.keys[1][t] < i) ? 1u : 0u;
        const uint32_t dst = sb + rank;
        C.sortedSph[dst] = C.sphF[i];
        const uint2 s = P.sph[i];
        uint32_t fam = 0;
        if (need_fam) fam = P.state[s.x].pos.family;
        C.sortedMeta[dst] = make_uint4(s.x, i, s.y, fam);
        C.sortedAux[dst] = make_uint2(s.x, __float_as_uint(__ldg(&P.comp[s.y & 0xffffu]).w));
        C.vals[0][dst] = key;  // sorted keys for the sweep
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sphere--triangle broad phase (replaces makeTriangleSandwich / bin--triangle pairs / per-bin sphere--triangle sweep,
// DEMBinTriangleKernels.cu:22-221, DEMContactKernels_SphereTriangle.cu:116-427, and the host merge of
// HostSideHelpers.hpp:176-193).  A triangle is registered in every cell that its bounding box, grown by the largest
// inflated sphere radius plus its own margin, overlaps; a sphere then only looks at the triangles of its own cell.
// Returns false when the facet lies entirely outside the cell columns this rank bins (domain decomposition: every rank
// holds all facets of the replicated mesh owners but registers only those that can meet its slab).
__device__ __forceinline__ bool tri_cell_range(const GridInfo& g, const CdParams& C, float4 a, float4 b, float4 c,
                                               int lo[3], int hi[3]) {
    const float grow = C.rmax + g.max_margin + a.w + 1e-6f * (fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + 1.f);
    const float mn[3] = {fminf(a.x, fminf(b.x, c.x)) - grow, fminf(a.y, fminf(b.y, c.y)) - grow, fminf(a.z, fminf(b.z, c.z)) - grow};
    const float mx[3] = {fmaxf(a.x, fmaxf(b.x, c.x)) + grow, fmaxf(a.y, fmaxf(b.y, c.y)) + grow, fmaxf(a.z, fmaxf(b.z, c.z)) + grow};
    const int nb[3] = {(int)g.nbx, (int)g.nby, (int)g.nbz};
    if (C.slab_on) {
        const int xl = (int)floorf(mn[0] * g.inv_cs) - g.x0, xh = (int)floorf(mx[0] * g.inv_cs) - g.x0;
        if (xh < 0 || xl > nb[0] - 1) return false;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int off = (k == 0) ? g.x0 : 0;
        lo[k] = min(max((int)floorf(mn[k] * g.inv_cs) - off, 0), nb[k] - 1);
        hi[k] = min(max((int)floorf(mx[k] * g.inv_cs) - off, 0), nb[k] - 1);
    }
    return true;
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_tri_cells(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nTri) return;
    const GridInfo g = *C.grid;
    float4 a, b, c;
    if (!FILL) {
        const uint2 info = P.tri_info[t];
        const OwnerState* sb = P.state + info.x;
        const float4 q = sb->quat;
        const float4 v = sb->vel;
        double X, Y, Z;
        pos_decode(sb->pos, P, X, Y, Z);
        const float m = owner_margin(P, sqrtf(v.x * v.x + v.y * v.y + v.z * v.z), sb->pos.family);
        const float4 n1 = P.tri_n1[t], n2 = P.tri_n2[t], n3 = P.tri_n3[t];
        const float3 r1 = rotate(f3(n1.x, n1.y, n1.z), q), r2 = rotate(f3(n2.x, n2.y, n2.z), q), r3 = rotate(f3(n3.x, n3.y, n3.z), q);
        a = make_float4((float)(X + (double)r1.x), (float)(Y + (double)r1.y), (float)(Z + (double)r1.z), m);
        b = make_float4((float)(X + (double)r2.x), (float)(Y + (double)r2.y), (float)(Z + (double)r2.z), 0.f);
        c = make_float4((float)(X + (double)r3.x), (float)(Y + (double)r3.y), (float)(Z + (double)r3.z), 0.f);
        C.triW1[t] = a; C.triW2[t] = b; C.triW3[t] = c;
    } else {
        a = C.triW1[t]; b = C.triW2[t]; c = C.triW3[t];
    }
    int lo[3], hi[3];
    if (!tri_cell_range(g, C, a, b, c, lo, hi)) return;
    for (int z = lo[2]; z <= hi[2]; z++)
        for (int y = lo[1]; y <= hi[1]; y++)
            for (int x = lo[0]; x <= hi[0]; x++) {
                const uint32_t cell = (uint32_t)x + g.nbx * ((uint32_t)y + g.nby * (uint32_t)z);
                if (!FILL) {
                    atomicAdd(&C.triCellStart[cell], 1u);
                } else {
                    const uint32_t slot = C.triCellStart[cell] + atomicAdd(&C.triCellFill[cell], 1u);
                    if (slot < C.tri_pair_cap) C.triCellList[slot] = t; else atomicOr(&P.flags[DEM_FLAG_CAPACITY], 16u);
                }
            }
}

// squared distance from point p to triangle (a,b,c): Ericson, Real-Time Collision Detection, p.141 (float, conservative use)
__device__ __forceinline__ float tri_point_dist2(float3 a, float3 b, float3 c, float3 p) {
    const float3 ab = b - a, ac = c - a, ap = p - a;
    const float d1 = dot(ab, ap), d2 = dot(ac, ap);
    float3 q;
    if (d1 <= 0.f && d2 <= 0.f) q = a;
    else {
        const float3 bp = p - b;
        const float d3 = dot(ab, bp), d4 = dot(ac, bp);
        if (d3 >= 0.f && d4 <= d3) q = b;
        else {
            const float vc = d1 * d4 - d3 * d2;
            if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) q = a + (d1 / (d1 - d3)) * ab;
            else {
                const float3 cp = p - c;
                const float d5 = dot(ab, cp), d6 = dot(ac, cp);
                if (d6 >= 0.f && d5 <= d6) q = c;
                else {
                    const float vb = d5 * d2 - d1 * d6;
                    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) q = a + (d2 / (d2 - d6)) * ac;
                    else {
                        const float va = d3 * d6 - d5 * d4;
                        if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) q = b + ((d4 - d3) / ((d4 - d3) + (d5 - d6))) * (c - b);
                        else {
                            const float denom = 1.f / (va + vb + vc);
                            q = a + (vb * denom) * ab + (vc * denom) * ac;
                        }
                    }
                }
            }
        }
    }
    const float3 d = p - q;
    return dot(d, d);
}

// is triangle t a contact candidate of the sphere `me` (inflated radius in .w, family famS)?
__device__ __forceinline__ bool st_candidate(const DevParams& P, const CdParams& C, float4 me, uint32_t famS, uint32_t t) {
    const float4 a = C.triW1[t], b = C.triW2[t], c = C.triW3[t];
    const float R = me.w + a.w;
    const float d2 = tri_point_dist2(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), f3(me.x, me.y, me.z));
    if (d2 > R * R * 1.00001f + 1e-18f) return false;
    if (C.any_mask) {
        const uint32_t famT = P.state[P.tri_info[t].x].pos.family;
        if (P.familyMasks[mask_pair(famS, famT)] != 0) return false;
    }
    return true;
}

// per sphere (sphere-id order): candidates among the triangles registered in the sphere's own cell.  Two passes over
// the cell's triangle list -- count, claim a run of slots with one warp-aggregated atomic, test again and write -- so
// there is no cap on the number of facets a sphere may touch (a mesh finer than the grains is legal).
__global__ void __launch_bounds__(128) k_st_emit(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t n = C.act_sph ? *C.act_count : P.nSpheres;
    const uint32_t nround = (n + 31u) & ~31u;
    for (uint32_t t_ = blockIdx.x * blockDim.x + threadIdx.x; t_ < nround; t_ += gridDim.x * blockDim.x) {
        const bool valid = t_ < n;
        const uint32_t sid = (valid && C.act_sph) ? C.act_sph[t_] : t_;
        uint32_t count = 0, tb = 0, te = 0, famS = 0;
        uint2 s = make_uint2(0, 0);
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const uint32_t key = C.keys[0][sid];
            s = P.sph[sid];
            me = C.sphF[sid];
            famS = P.state[s.x].pos.family;
            tb = C.triCellStart[key];
            te = min(C.triCellStart[key + 1], C.tri_pair_cap);
            for (uint32_t k = tb; k < te; k++) count += st_candidate(P, C, me, famS, C.triCellList[k]) ? 1u : 0u;
        }
        uint32_t slot = warp_claim(count, P.st.count);
        if (!valid) continue;
        P.st.seg_start[sid] = slot;
        P.st.seg_count[sid] = (slot + count <= C.capacity) ? count : (slot < C.capacity ? C.capacity - slot : 0u);
        if (count == 0) continue;
        if (slot + count > C.capacity) atomicOr(&P.flags[DEM_FLAG_CAPACITY], 32u);
        const uint32_t oldStart = C.oldst.seg_start[sid], oldCount = C.oldst.seg_count[sid];
        for (uint32_t k = tb; k < te; k++) {
            const uint32_t t = C.triCellList[k];
            if (!st_candidate(P, C, me, famS, t)) continue;
            if (slot < C.capacity) {
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t alive = 0;
                for (uint32_t j = 0; j < oldCount; j++) {
                    if (C.oldst.idB[oldStart + j] == t) {
                        alive = C.oldst.cinfo[oldStart + j].w & 0x80000000u;
                        if (alive && C.oldst.hist) h = C.oldst.hist[oldStart + j];
                        break;
                    }
                }
                const uint32_t matpair = (s.y >> 16) * P.nMat + P.tri_info[t].y;
                P.st.idB[slot] = t;
                P.st.cinfo[slot] = make_uint4(s.x, t, s.y & 0xffffu, matpair | alive);
                if (P.st.hist) P.st.hist[slot] = h;
            }
            slot++;
        }
    }
}

// stage 0: world nodes + per-cell counts + fill of the per-cell triangle lists; stage 1: the sphere--triangle list
int launch_cd_triangles(const DevParams& P, const CdParams& C, int stage, int num_sms, cudaStream_t s) {
    if (P.nTri == 0) {
        if (stage == 0) cudaMemsetAsync(P.st.count, 0, sizeof(uint32_t) * 4, s);
        return 0;
    }
    int launches = 0;
    if (stage == 0) {
        cudaMemsetAsync(P.st.count, 0, sizeof(uint32_t) * 4, s);
        cudaMemsetAsync(C.triCellStart, 0, sizeof(uint32_t) * ((size_t)C.max_cells + 2), s);
        cudaMemsetAsync(C.triCellFill, 0, sizeof(uint32_t) * ((size_t)C.max_cells + 2), s);
        k_tri_cells<false><<<(P.nTri + 127) / 128, 128, 0, s>>>(P, C);
        launches += 1 + launch_scan_lookback(C.triCellStart, C.triCellStart, &C.grid->ncells, 1u, C.max_cells + 1, C.scan_desc,
                                             nullptr, P.flags, num_sms, s);
        k_tri_cells<true><<<(P.nTri + 127) / 128, 128, 0, s>>>(P, C);
        launches++;
    } else if (P.nSpheres) {
        const int grid = (int)std::min<uint32_t>((P.nSpheres + 127) / 128, (uint32_t)num_sms * 16u);
        k_st_emit<<<grid, 128, 0, s>>>(P, C);
        launches++;
    }
    return launches;
}

// History carry-over (DEMHistoryMappingKernels.cu), one thread per contact of the two new sphere--sphere lists: the pair
// may sit in either previous list as (A,B) or -- when the two spheres swapped their order in the sorted array -- as
// (B,A); then delta_tan changes sign.  Adjacent threads hold the contacts of the same sphere A (the lists are
// sphere-major), so their segment look-ups coalesce.  Sphere A of a contact comes from the scratch array the fill pass
// of the sweep left behind (the lists themselves only keep geometry B: A is implied by the segment table).
__global__ void __launch_bounds__(256) k_history(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t nT = min(P.ss.count[0], C.capacity);
    const uint32_t n = nT + min(P.sn.count[0], C.capacity);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
        const bool touching = c < nT;
        const ContactList& L = touching ? P.ss : P.sn;
        const uint32_t slot = touching ? c : c - nT;
        const uint32_t w = L.cinfo[slot].w;
        if (w & CINFO_NO_HISTORY) {  // (a contact that is not alive never has its history word read)
            L.cinfo[slot].w = w & ~CINFO_NO_HISTORY;
            continue;
        }
        const uint32_t sidA = (touching ? C.idA_ss : C.idA_sn)[slot];
        const uint32_t sidB = L.idB[slot];
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t alive = 0;
        bool found = false;
#pragma unroll 1
        for (int pass = 0; pass < 4 && !found; pass++) {
            const ContactList& O = ((pass & 1) == (touching ? 0 : 1)) ? C.oldss : C.oldsn;  // likelier list first
            const bool flipped = pass >= 2;
            const uint32_t a = flipped ? sidB : sidA, b = flipped ? sidA : sidB;
            const uint32_t os = O.seg_start[a], oc = O.seg_count[a];
            for (uint32_t t = 0; t < oc; t++) {
                if (O.idB[os + t] == b) {
                    alive = O.cinfo[os + t].w & 0x80000000u;
                    if (alive && O.hist) {
                        h = O.hist[os + t];
                        if (flipped) { h.x = -h.x; h.y = -h.y; h.z = -h.z; }
                    }
                    found = true;
                    break;
                }
            }
        }
        if (alive) L.cinfo[slot].w = w | alive;
        if (L.hist) L.hist[slot] = h;
    }
}

// Last kernel of a rebuild (one warp): clamp the counts, agree on the verdict over the ranks, poison the context when a
// list or table overflowed anywhere, and leave the status record for the host.
__global__ void __launch_bounds__(32) k_finish_counts(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C,
                                                      const __grid_constant__ MgDev M, int par) {
    const int lane = threadIdx.x;
    const uint32_t seq = P.flags[DEM_FLAG_SEQ] + 1u;
    uint32_t poison = P.flags[DEM_FLAG_POISON];
    uint32_t cnt[4] = {0, 0, 0, 0}, dem[4] = {0, 0, 0, 0};
    uint32_t tri_demand = 0;
    if (poison == 0u) {
        if (lane == 0) {
            const ContactList* L[4] = {&P.ss, &P.sn, &P.sa, &P.st};
            const uint32_t bit[4] = {1u, 4u, 2u, 32u};
            for (int k = 0; k < 4; k++) {
                dem[k] = L[k]->count[0];
                cnt[k] = min(dem[k], C.capacity);
                if (dem[k] > C.capacity) atomicOr(&P.flags[DEM_FLAG_CAPACITY], bit[k]);
                L[k]->count[0] = cnt[k];
                L[k]->count[1] = dem[k];
            }
            if (P.nTri) {
                tri_demand = C.triCellStart[C.grid->ncells];
                if (tri_demand > C.tri_pair_cap) atomicOr(&P.flags[DEM_FLAG_CAPACITY], 16u);
                P.flags[DEM_FLAG_TRI_DEMAND] = tri_demand;
            }
            __threadfence();
        }
        __syncwarp();
        uint32_t cap = P.flags[DEM_FLAG_CAPACITY] | ((P.flags[DEM_FLAG_HALO] & 8u) ? 8u : 0u);
        uint32_t vel = P.flags[DEM_FLAG_VELOCITY];
        if (M.world > 1) {
            // every rank must reach the same verdict, or they fall out of step with each other
            uint32_t o0, o1;
            mg_allgather(M, cap, vel, o0, o1, P.flags);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                o0 |= __shfl_xor_sync(0xffffffffu, o0, off);
                o1 |= __shfl_xor_sync(0xffffffffu, o1, off);
            }
            cap = o0;
            vel = o1;
            if (lane == 0 && vel) P.flags[DEM_FLAG_VELOCITY] = vel;
        }
        if (cap) poison = seq;
        if (lane == 0) {
            if (poison) P.flags[DEM_FLAG_POISON] = poison;
            else P.flags[DEM_FLAG_CYCLE_STEP] = 0u;  // the new lists are in use from here on
            RebuildStatus* st = C.status + (seq % REBUILD_STATUS_SLOTS);
            st->poison = poison;
            st->capflags = cap;
            st->velflag = vel;
            st->haloflags = P.flags[DEM_FLAG_HALO];
            st->tri_demand = tri_demand;
            for (int k = 0; k < 4; k++) { st->count[k] = cnt[k]; st->demand[k] = dem[k]; }
            st->grid = *C.grid;
            for (int k = 0; k < 8; k++) st->mg[k] = (M.world > 1) ? M.counts[par][k] : 0u;
            if (M.world > 1) {
                const uint32_t* rc = reinterpret_cast<const uint32_t*>(M.my_block + MG_HDR_RECV_COUNT) + par * 2;
                st->mg[5] = rc[0]; st->mg[6] = rc[1];
            }
            __threadfence_system();
            st->seq = seq;
        }
    } else if (lane == 0) {
        RebuildStatus* st = C.status + (seq % REBUILD_STATUS_SLOTS);
        st->poison = poison;
        st->capflags = 0; st->velflag = P.flags[DEM_FLAG_VELOCITY]; st->haloflags = P.flags[DEM_FLAG_HALO];
        __threadfence_system();
        st->seq = seq;
    }
    if (lane == 0) P.flags[DEM_FLAG_SEQ] = seq;
}

// ---------------------------------------------------------------------------------------------------------------
static int gs_grid(uint32_t n, int threads, int num_sms, int per_sm) {
    return (int)std::max<uint32_t>(1u, std::min<uint32_t>((n + threads - 1) / threads, (uint32_t)(num_sms * per_sm)));
}

int launch_cd_prepare(const DevParams& P, const CdParams& C, const MgDev* M, bool need_maxvel, int stage, int num_sms,
                      cudaStream_t s) {
    int launches = 0;
    if (stage == 0) {
        if (need_maxvel) {
            // velocities changed outside the integrator (initial state / host upload): recompute max |v|
            cudaMemsetAsync(P.maxvel, 0, sizeof(float), s);
            if (P.nOwners) {
                k_maxvel<<<gs_grid(P.nOwners, 256, num_sms, 8), 256, 0, s>>>(P, P.errOutVel);
                launches++;
            }
        }
    } else if (stage == 1) {
        MgDev single;
        memset(&single, 0, sizeof(single));
        single.world = 1;
        k_grid_setup<<<1, 32, 0, s>>>(P, C, M ? *M : single);
        launches++;
    } else {
        if (P.nAnal) {
            k_anal_prep<<<(P.nAnal + 63) / 64, 64, 0, s>>>(P, C);
            launches++;
        }
        cudaMemsetAsync(C.cellStart, 0, sizeof(uint32_t) * ((size_t)C.max_cells + 2), s);  // scratch
        // (the count words of the lists being built; under poison the lists they belong to are not looked at)
        cudaMemsetAsync(P.ss.count, 0, sizeof(uint32_t) * 4, s);
        cudaMemsetAsync(P.sn.count, 0, sizeof(uint32_t) * 4, s);
        cudaMemsetAsync(P.sa.count, 0, sizeof(uint32_t) * 4, s);
        // (decomposed: spheres this rank does not hold are not visited; k_mg_clear_segs has emptied the segments their
        // last visit to these buffers left behind)
        if (P.nSpheres) {
            k_sphere_prep<<<gs_grid(P.nSpheres, 256, num_sms, 8), 256, 0, s>>>(P, C);
            launches++;
        }
    }
    return launches;
}

int launch_cd_sweep(const DevParams& P, const CdParams& C, const MgDev* M, int par, int sorted_buf, int num_sms,
                    cudaStream_t s, cudaEvent_t* ev, bool sort_only) {
    int launches = 0;
    const uint32_t n = P.nSpheres;
    // cell histogram -> exclusive prefix over the ncells+1 entries of THIS rebuild's grid (device-resident length)
    launches += launch_scan_lookback(C.cellStart, C.cellStart, &C.grid->ncells, 1u, C.max_cells + 1, C.scan_desc, nullptr,
                                     P.flags, num_sms, s);
    if (ev) cudaEventRecord(ev[0], s);
    if (n) {
        if (sorted_buf < 0) {  // counting sort
            k_cs_scatter<<<gs_grid(n, 256, num_sms, 8), 256, 0, s>>>(P, C);
            k_gather_sorted_cs<<<gs_grid(n, 256, num_sms, 8), 256, 0, s>>>(P, C);
            launches += 2;
        } else {
            k_gather_sorted<<<(n + 255) / 256, 256, 0, s>>>(P, C, C.vals[sorted_buf]);
            launches++;
        }
        if (ev) cudaEventRecord(ev[1], s);
        if (!sort_only) {
            const uint32_t* keys = sorted_buf < 0 ? C.vals[0] : C.keys[sorted_buf];
            const int grid = gs_grid(n, 128, num_sms, 8);
            launch_sweep_count(P, C, keys, grid, s);
            // per-sphere counts -> first slots: over all spheres in sphere-id order, or (decomposed) over the active
            // spheres in the order of their list -- owner-major either way
            const uint32_t* np = C.act_sph ? C.act_count : nullptr;
            const uint32_t na = C.act_sph ? 0u : n;
            launches += 1 + launch_scan_lookback(P.ss.seg_count, P.ss.seg_start, np, na, n, C.scan_desc, P.ss.count, P.flags, num_sms, s, C.act_sph);
            launches += launch_scan_lookback(P.sn.seg_count, P.sn.seg_start, np, na, n, C.scan_desc, P.sn.count, P.flags, num_sms, s, C.act_sph);
            launch_sweep_fill(P, C, keys, gs_grid(n, 256, num_sms, 8), s);
            launches++;
            if (P.ss.hist) {  // (a history-less force model has nothing to carry over)
                k_history<<<num_sms * 8, 256, 0, s>>>(P, C);
                launches++;
            }
        }
    } else if (ev) {
        cudaEventRecord(ev[1], s);
    }
    if (ev) cudaEventRecord(ev[2], s);
    if (!sort_only) {
        MgDev single;
        memset(&single, 0, sizeof(single));
        single.world = 1;
        k_finish_counts<<<1, 32, 0, s>>>(P, C, M ? *M : single, par);
        launches++;
    }
    if (ev) cudaEventRecord(ev[3], s);
    return launches;
}

// ---------------------------------------------------------------------------------------------------------------
// reductions over clump owners (DEMInspector built-ins, AuxClasses.cpp:88-164)
__global__ void k_reduce(const __grid_constant__ DevParams P, int kind, uint32_t nClumps, double* out) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    double v = (kind == DEM_REDUCE_MIN_Z) ? 1e300 : ((kind == DEM_REDUCE_MAX_Z) ? -1e300 : 0.0);
    if (o < nClumps && (!P.active || P.active[o] == 1 || P.active[o] >= 3)) {
        const OwnerState st = P.state[o];
        if (kind == DEM_REDUCE_MAX_ABSV) {
            v = sqrtf(st.vel.x * st.vel.x + st.vel.y * st.vel.y + st.vel.z * st.vel.z);
        } else if (kind == DEM_REDUCE_MAX_Z || kind == DEM_REDUCE_MIN_Z) {
            double X, Y, Z;
            pos_decode(st.pos, P, X, Y, Z);
            v = Z + (double)P.LBF[2];
        } else if (kind == DEM_REDUCE_KINETIC_ENERGY) {
            const float4 sp = P.spin[o];
            const float4 mp = P.massprop[__float_as_uint(sp.w)];
            v = 0.5 * (double)st.vel.w * ((double)st.vel.x * st.vel.x + (double)st.vel.y * st.vel.y +
                                         (double)st.vel.z * st.vel.z) +
                0.5 * ((double)mp.y * sp.x * sp.x + (double)mp.z * sp.y * sp.y + (double)mp.w * sp.z * sp.z);
        } else if (kind == DEM_REDUCE_TOTAL_MASS) {
            v = st.vel.w;
        }
    }
    const bool is_max = (kind == DEM_REDUCE_MAX_ABSV || kind == DEM_REDUCE_MAX_Z);
    const bool is_min = (kind == DEM_REDUCE_MIN_Z);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, v, off);
        v = is_max ? fmax(v, t) : (is_min ? fmin(v, t) : v + t);
    }
    if ((threadIdx.x & 31) == 0) {
        if (is_max || is_min) {
            // CAS loop on the double
            unsigned long long* addr = reinterpret_cast<unsigned long long*>(out);
            unsigned long long old = *addr, assumed;
            do {
                assumed = old;
                const double cur = __longlong_as_double((long long)assumed);
                const double nv = is_max ? fmax(cur, v) : fmin(cur, v);
                if (nv == cur) break;
                old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(nv));
            } while (assumed != old);
        } else {
            atomicAdd(out, v);
        }
    }
}

// All five reductions in ONE pass over the owners (a caller that polls several inspectors per frame pays for the 80-byte
// owner stream once): per-thread values -> warp shuffle -> shared memory -> one atomic per CTA and quantity.
// out[kind] must hold the neutral element of each reduction on entry.
__device__ __forceinline__ void atomic_minmax_double(double* out, double v, bool is_max) {
    unsigned long long* addr = reinterpret_cast<unsigned long long*>(out);
    unsigned long long old = *addr, assumed;
    do {
        assumed = old;
        const double cur = __longlong_as_double((long long)assumed);
        const double nv = is_max ? fmax(cur, v) : fmin(cur, v);
        if (nv == cur) break;
        old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(nv));
    } while (assumed != old);
}

__global__ void __launch_bounds__(256) k_reduce_many(const __grid_constant__ DevParams P, uint32_t mask, uint32_t nClumps,
                                                     double* out) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    double v[5] = {0.0, -1e300, 1e300, 0.0, 0.0};
    if (o < nClumps && (!P.active || P.active[o] == 1 || P.active[o] >= 3)) {
        const OwnerState st = P.state[o];
        v[DEM_REDUCE_MAX_ABSV] = sqrtf(st.vel.x * st.vel.x + st.vel.y * st.vel.y + st.vel.z * st.vel.z);
        if (mask & ((1u << DEM_REDUCE_MAX_Z) | (1u << DEM_REDUCE_MIN_Z))) {
            double X, Y, Z;
            pos_decode(st.pos, P, X, Y, Z);
            v[DEM_REDUCE_MAX_Z] = v[DEM_REDUCE_MIN_Z] = Z + (double)P.LBF[2];
        }
        if (mask & (1u << DEM_REDUCE_KINETIC_ENERGY)) {
            const float4 sp = P.spin[o];
            const float4 mp = P.massprop[__float_as_uint(sp.w)];
            v[DEM_REDUCE_KINETIC_ENERGY] =
                0.5 * (double)st.vel.w * ((double)st.vel.x * st.vel.x + (double)st.vel.y * st.vel.y + (double)st.vel.z * st.vel.z) +
                0.5 * ((double)mp.y * sp.x * sp.x + (double)mp.z * sp.y * sp.y + (double)mp.w * sp.z * sp.z);
        }
        v[DEM_REDUCE_TOTAL_MASS] = st.vel.w;
    }
    __shared__ double sh[5][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        if (!(mask & (1u << k))) continue;
        const bool is_max = (k == DEM_REDUCE_MAX_ABSV || k == DEM_REDUCE_MAX_Z), is_min = (k == DEM_REDUCE_MIN_Z);
        double x = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, x, off);
            x = is_max ? fmax(x, t) : (is_min ? fmin(x, t) : x + t);
        }
        if (lane == 0) sh[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < 5 && (mask & (1u << threadIdx.x))) {
        const int k = threadIdx.x;
        const bool is_max = (k == DEM_REDUCE_MAX_ABSV || k == DEM_REDUCE_MAX_Z), is_min = (k == DEM_REDUCE_MIN_Z);
        double x = sh[k][0];
        for (int w = 1; w < 8; w++) x = is_max ? fmax(x, sh[k][w]) : (is_min ? fmin(x, sh[k][w]) : x + sh[k][w]);
        if (is_max || is_min) atomic_minmax_double(out + k, x, is_max); else atomicAdd(out + k, x);
    }
}

// Sphere-level inspector quantities (AuxClasses.cpp:19-50, DEMSphereQueryKernels.cu:13-52 of the reference): its
// clump_max_z / clump_min_z / clump_max_absv look at every SPHERE -- the top / bottom of the sphere, the velocity of its
// centre v + omega x r -- not at the owner's centre.  One thread per sphere; *out holds the neutral element on entry.
__global__ void __launch_bounds__(256) k_reduce_spheres(const __grid_constant__ DevParams P, int kind, double* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool is_min = (kind == DEM_REDUCE_SPHERE_MIN_Z);
    double v = is_min ? 1e300 : ((kind == DEM_REDUCE_SPHERE_MAX_ABSV) ? 0.0 : -1e300);
    if (i < P.nSpheres) {
        const uint2 s = P.sph[i];
        if (!P.active || P.active[s.x] == 1 || P.active[s.x] >= 3) {
            const OwnerState st = P.state[s.x];
            const float4 comp = P.comp[s.y & 0xffffu];
            const float3 rel = rotate(f3(comp.x, comp.y, comp.z), st.quat);
            if (kind == DEM_REDUCE_SPHERE_MAX_ABSV) {
                const float3 u = cross(f3(st.omg.x, st.omg.y, st.omg.z), rel) + f3(st.vel.x, st.vel.y, st.vel.z);
                v = sqrtf(dot(u, u));
            } else {
                double X, Y, Z;
                pos_decode(st.pos, P, X, Y, Z);
                const float z = (float)(Z + (double)rel.z + (double)P.LBF[2]);
                v = is_min ? z - comp.w : z + comp.w;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double t = __shfl_xor_sync(0xffffffffu, v, off);
        v = is_min ? fmin(v, t) : fmax(v, t);
    }
    if ((threadIdx.x & 31) == 0) atomic_minmax_double(out, v, !is_min);
}

int launch_reduce_spheres(const DevParams& P, int kind, double* d_out, cudaStream_t s) {
    if (P.nSpheres == 0) return 0;
    k_reduce_spheres<<<(P.nSpheres + 255) / 256, 256, 0, s>>>(P, kind, d_out);
    return 1;
}

int launch_reduce_many(const DevParams& P, uint32_t mask, double* d_out, cudaStream_t s) {
    const uint32_t n = P.nOwners;
    if (n == 0) return 0;
    k_reduce_many<<<(n + 255) / 256, 256, 0, s>>>(P, mask, n, d_out);
    return 1;
}

int launch_reduce(const DevParams& P, int kind, double* d_out, cudaStream_t s) {
    // clump owners are the owners that have spheres: the caller passes nOwners restricted to clumps via P.nOwners
    const uint32_t n = P.nOwners;
    if (n == 0) return 0;
    k_reduce<<<(n + 255) / 256, 256, 0, s>>>(P, kind, n, d_out);
    return 1;
}

}  // namespace demb
