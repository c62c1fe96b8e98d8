// kernels_cd.cu -- contact-list rebuild ("kinematic" work) for sm_100a, fully device-driven (no host round trips
// between the stages):
//   k_maxvel / k_grid_setup      max |v| -> margin -> broad-phase cell size and grid, decided ON the device; also
//                                resolves the analytical components (planes, cylinders) to world space once
//   k_sphere_prep                per sphere: world position (fixed-point decode + rotated offset), inflated radius,
//                                cell key + cell histogram; emits the sphere--analytical list (+history) directly
//                                with warp-aggregated atomics
//   radix sort (k_rs_*)          LSD, 8-bit digits, tile ranking with warp match + shared-memory staged coalesced scatter
//   scan (k_scan_*)              exclusive prefix sums (cell table, per-sphere contact offsets, sort histograms)
//   k_gather_sorted              cell-ordered float4 {x,y,z,r'} + {owner,id}
//   k_sweep                      ONE pass over the upper half of the 27-cell stencil (5 contiguous runs of the sorted array
//                                instead of 9): distance test on the float4 stream first, accepted candidates staged per
//                                thread, slots claimed with one warp-aggregated atomic, then the compiled per-contact
//                                record is written and the Hertz-Mindlin history carried over from the previous list
// Reference behaviour being reproduced: contactDetection(), src/algorithms/DEMCubContactDetection.cu:38-1123;
// acceptance rule of src/kernel/DEMContactKernels_SphereSphere.cu:57-89,172-214 and DEMBinSphereKernels.cu:78-128;
// margin of src/kernel/DEMMiscKernels.cu:37-69; history map of src/kernel/DEMHistoryMappingKernels.cu.
// The candidate list is a SUPERSET of the reference's (the force kernel re-tests true overlap), so physics is identical.
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
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    float a = 0.f;
    if (o < P.nOwners && (!P.active || P.active[o] != 0)) {
        const float4 v = P.state[o].vel;
        a = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
        if (!isfinite(a) || a > errOutVel) atomicOr(&P.flags[3], 1u);
        if (!isfinite(a)) a = 0.f;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, off));
    if ((threadIdx.x & 31) == 0 && a > 0.f) atomicMax(reinterpret_cast<int*>(P.maxvel), __float_as_int(a));
}

__global__ void k_grid_setup(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float vmax = *P.maxvel;
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
    const uint32_t t_ = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = t_ < (C.act_sph ? C.nActSph : P.nSpheres);
    const uint32_t i = (valid && C.act_sph) ? C.act_sph[t_] : t_;
    uint32_t nsa = 0, samask = 0;
    float3 sp = f3(0.f, 0.f, 0.f);
    uint2 s = make_uint2(0, 0);
    uint32_t family = 0;
    if (valid) {
        s = P.sph[i];
        if (P.active && P.active[s.x] == 0) {
            // owner not held by this rank (domain decomposition): the sphere takes no part in this rebuild
            C.keys[0][i] = 0xffffffffu;
            P.sa.seg_start[i] = 0;
            P.sa.seg_count[i] = 0;
            valid = false;
        }
    }
    if (valid) {
        const GridInfo g = *C.grid;
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
        const float rInfl = comp.w + margin;
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
    if (P.nAnal == 0) return;
    // ---- emit the sphere--analytical contacts of this warp into one contiguous run ----
    uint32_t slot = warp_claim(nsa, P.sa.count);
    if (!valid) return;
    P.sa.seg_start[i] = slot;
    P.sa.seg_count[i] = (slot + nsa <= C.capacity) ? nsa : (slot < C.capacity ? C.capacity - slot : 0u);
    if (nsa == 0) return;
    if (slot + nsa > C.capacity) atomicOr(&P.flags[0], 2u);
    const uint32_t oldStart = C.oldsa.seg_start[i], oldCount = C.oldsa.seg_count[i];
    for (uint32_t k = 0; k < P.nAnal; k++) {
        bool hit;
        if (k < 32) hit = (samask >> k) & 1u;
        else hit = sa_candidate(P, C.analw[k], sp, C.sphF[i].w, family, C.any_mask != 0);
        if (!hit) continue;
        if (slot < C.capacity) {
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t alive = 0;
            for (uint32_t t = 0; t < oldCount; t++) {
                if (C.oldsa.pair[oldStart + t].y == k) {
                    alive = C.oldsa.cinfo[oldStart + t].w & 0x80000000u;
                    if (alive && C.oldsa.hist) h = C.oldsa.hist[oldStart + t];
                    break;
                }
            }
            const uint32_t matpair = (s.y >> 16) * P.nMat + C.analw[k].material;
            P.sa.pair[slot] = make_uint2(i, k);
            P.sa.cinfo[slot] = make_uint4(s.x, k, s.y & 0xffffu, matpair | alive);
            if (P.sa.hist) P.sa.hist[slot] = h;
        }
        slot++;
    }
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
__global__ void __launch_bounds__(256) k_gather_sorted(const __grid_constant__ DevParams P,
                                                       const __grid_constant__ CdParams C,
                                                       const uint32_t* __restrict__ vals) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.nSpheres) return;
    const uint32_t i = vals[j];
    C.sortedSph[j] = C.sphF[i];
    const uint2 s = P.sph[i];
    uint32_t fam = 0;
    if (C.any_mask != 0 || C.max_extra > 0.f) fam = P.state[s.x].pos.family;
    C.sortedMeta[j] = make_uint4(s.x, i, s.y, fam);  // {owner, sphere id, comp | material<<16, family}
    C.sortedPos[i] = j;
}

// ---- counting sort by cell (sort_mode 1): the histogram and its prefix exist anyway for the sweep ----
__global__ void __launch_bounds__(256) k_cs_scatter(const __grid_constant__ DevParams P,
                                                    const __grid_constant__ CdParams C) {
    const uint32_t t_ = blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= (C.act_sph ? C.nActSph : P.nSpheres)) return;
    const uint32_t i = C.act_sph ? C.act_sph[t_] : t_;
    const uint32_t key = C.keys[0][i];
    if (key == 0xffffffffu) return;  // inactive on this rank
    C.keys[1][C.cellStart[key] + C.vals[1][i]] = i;  // arrival order inside the cell (non-deterministic)
}

// gather into cell order; inside a cell the spheres are ranked by sphere id, which makes the result identical to a
// stable radix sort of (cell key, sphere id) no matter in which order the atomics arrived
__global__ void __launch_bounds__(256) k_gather_sorted_cs(const __grid_constant__ DevParams P,
                                                          const __grid_constant__ CdParams C) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C.cellStart[C.grid->ncells]) return;  // number of active spheres (== nSpheres on a single GPU)
    const uint32_t i = C.keys[1][j];
    const uint32_t key = C.keys[0][i];
    const uint32_t sb = C.cellStart[key], se = C.cellStart[key + 1];
    uint32_t rank = 0;
    for (uint32_t t = sb; t < se; t++) rank += (C.keys[1][t] < i) ? 1u : 0u;
    const uint32_t dst = sb + rank;
    C.sortedSph[dst] = C.sphF[i];
    const uint2 s = P.sph[i];
    uint32_t fam = 0;
    if (C.any_mask != 0 || C.max_extra > 0.f) fam = P.state[s.x].pos.family;
    C.sortedMeta[dst] = make_uint4(s.x, i, s.y, fam);
    C.vals[0][dst] = key;  // sorted keys for the sweep
    C.sortedPos[i] = dst;
}

// ---------------------------------------------------------------------------------------------------------------
// sphere--triangle broad phase (replaces makeTriangleSandwich / bin--triangle pairs / per-bin sphere--triangle sweep,
// DEMBinTriangleKernels.cu:22-221, DEMContactKernels_SphereTriangle.cu:116-427, and the host merge of
// HostSideHelpers.hpp:176-193).  A triangle is registered in every cell that its bounding box, grown by the largest
// inflated sphere radius plus its own margin, overlaps; a sphere then only looks at the triangles of its own cell.
__device__ __forceinline__ void tri_cell_range(const GridInfo& g, const CdParams& C, float4 a, float4 b, float4 c,
                                               int lo[3], int hi[3]) {
    const float grow = C.rmax + g.max_margin + a.w + 1e-6f * (fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + 1.f);
    const float mn[3] = {fminf(a.x, fminf(b.x, c.x)) - grow, fminf(a.y, fminf(b.y, c.y)) - grow, fminf(a.z, fminf(b.z, c.z)) - grow};
    const float mx[3] = {fmaxf(a.x, fmaxf(b.x, c.x)) + grow, fmaxf(a.y, fmaxf(b.y, c.y)) + grow, fmaxf(a.z, fmaxf(b.z, c.z)) + grow};
    const int nb[3] = {(int)g.nbx, (int)g.nby, (int)g.nbz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int off = (k == 0) ? g.x0 : 0;
        lo[k] = min(max((int)floorf(mn[k] * g.inv_cs) - off, 0), nb[k] - 1);
        hi[k] = min(max((int)floorf(mx[k] * g.inv_cs) - off, 0), nb[k] - 1);
    }
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_tri_cells(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nTri) return;
    const GridInfo g = *C.grid;
    float4 a, b, c;
    if (!FILL) {
        const uint2 info = P.tri_info[t];
        if (P.active && P.active[info.x] == 0) {  // mesh owner not held by this rank
            C.triW1[t] = make_float4(0.f, 0.f, 0.f, -1.f);
            return;
        }
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
        if (a.w < 0.f) return;
    }
    int lo[3], hi[3];
    tri_cell_range(g, C, a, b, c, lo, hi);
    for (int z = lo[2]; z <= hi[2]; z++)
        for (int y = lo[1]; y <= hi[1]; y++)
            for (int x = lo[0]; x <= hi[0]; x++) {
                const uint32_t cell = (uint32_t)x + g.nbx * ((uint32_t)y + g.nby * (uint32_t)z);
                if (!FILL) {
                    atomicAdd(&C.triCellStart[cell], 1u);
                } else {
                    const uint32_t slot = C.triCellStart[cell] + atomicAdd(&C.triCellFill[cell], 1u);
                    if (slot < C.tri_pair_cap) C.triCellList[slot] = t; else atomicOr(&P.flags[0], 16u);
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

constexpr int ST_MAXC = 24;

// per sphere (sphere-id order): candidates among the triangles registered in the sphere's own cell
__global__ void __launch_bounds__(128) k_st_emit(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    const uint32_t sid = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = sid < P.nSpheres;
    uint32_t key = 0;
    if (valid) {
        key = C.keys[0][sid];
        if (key == 0xffffffffu) {
            P.st.seg_start[sid] = 0; P.st.seg_count[sid] = 0;
            valid = false;
        }
    }
    uint32_t acc[ST_MAXC];
    uint32_t count = 0;
    uint2 s = make_uint2(0, 0);
    if (valid) {
        s = P.sph[sid];
        const float4 me = C.sphF[sid];
        const uint32_t famS = P.state[s.x].pos.family;
        const uint32_t tb = C.triCellStart[key], te = min(C.triCellStart[key + 1], C.tri_pair_cap);
        for (uint32_t k = tb; k < te; k++) {
            const uint32_t t = C.triCellList[k];
            const float4 a = C.triW1[t], b = C.triW2[t], c = C.triW3[t];
            const float R = me.w + a.w;
            const float d2 = tri_point_dist2(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), f3(me.x, me.y, me.z));
            if (d2 > R * R * 1.00001f + 1e-18f) continue;
            if (C.any_mask) {
                const uint32_t famT = P.state[P.tri_info[t].x].pos.family;
                if (P.familyMasks[mask_pair(famS, famT)] != 0) continue;
            }
            if (count < ST_MAXC) acc[count] = t;
            count++;
        }
        if (count > ST_MAXC) { atomicOr(&P.flags[2], 2u); count = ST_MAXC; }
    }
    uint32_t slot = warp_claim(count, P.st.count);
    if (!valid) return;
    P.st.seg_start[sid] = slot;
    P.st.seg_count[sid] = (slot + count <= C.capacity) ? count : (slot < C.capacity ? C.capacity - slot : 0u);
    if (count == 0) return;
    if (slot + count > C.capacity) atomicOr(&P.flags[0], 32u);
    const uint32_t oldStart = C.oldst.seg_start[sid], oldCount = C.oldst.seg_count[sid];
    for (uint32_t k = 0; k < count; k++, slot++) {
        if (slot >= C.capacity) break;
        const uint32_t t = acc[k];
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t alive = 0;
        for (uint32_t j = 0; j < oldCount; j++) {
            if (C.oldst.pair[oldStart + j].y == t) {
                alive = C.oldst.cinfo[oldStart + j].w & 0x80000000u;
                if (alive && C.oldst.hist) h = C.oldst.hist[oldStart + j];
                break;
            }
        }
        const uint32_t matpair = (s.y >> 16) * P.nMat + P.tri_info[t].y;
        P.st.pair[slot] = make_uint2(sid, t);
        P.st.cinfo[slot] = make_uint4(s.x, t, s.y & 0xffffu, matpair | alive);
        if (P.st.hist) P.st.hist[slot] = h;
    }
}

// stage 0: world nodes + per-cell counts; stage 1: fill the per-cell triangle lists (after the scan) and emit the list
int launch_cd_triangles(const DevParams& P, const CdParams& C, int stage, cudaStream_t s) {
    if (P.nTri == 0) {
        if (stage == 0) cudaMemsetAsync(P.st.count, 0, sizeof(uint32_t) * 4, s);
        return 0;
    }
    int launches = 0;
    if (stage == 0) {
        cudaMemsetAsync(P.st.count, 0, sizeof(uint32_t) * 4, s);
        cudaMemsetAsync(C.triCellStart, 0, sizeof(uint32_t) * ((size_t)C.max_cells + 1), s);
        cudaMemsetAsync(C.triCellFill, 0, sizeof(uint32_t) * ((size_t)C.max_cells + 1), s);
        k_tri_cells<false><<<(P.nTri + 127) / 128, 128, 0, s>>>(P, C);
        launches += 1 + launch_scan_exclusive(C.triCellStart, C.max_cells + 1, C.scan_tmp, nullptr, s);
        k_tri_cells<true><<<(P.nTri + 127) / 128, 128, 0, s>>>(P, C);
        launches++;
    } else if (P.nSpheres) {
        k_st_emit<<<(P.nSpheres + 127) / 128, 128, 0, s>>>(P, C);
        launches++;
    }
    return launches;
}

constexpr int SWEEP_MAXC = 40;  // accepted candidates staged per sphere (half stencil)
constexpr uint32_t CINFO_NO_HISTORY = 0x40000000u;  // sweep -> k_history: this contact carries no history over

// The sweep.  Cells are x-fastest, so the x-neighbours of a row are ONE contiguous run of the cell-sorted array.
// Each sphere looks only "forward" (upper half of the 27-cell stencil => 5 runs, the own row starting right behind
// itself), so every pair is found exactly once by the sphere that comes first in the sorted order; that sphere is
// geometry A of the contact.
__global__ void __launch_bounds__(128) k_sweep(const __grid_constant__ DevParams P,
                                               const __grid_constant__ CdParams C,
                                               const uint32_t* __restrict__ keys) {
    // One thread per sphere IN SPHERE-ID ORDER (clump by clump), so that the slots claimed below make the contact list
    // owner-major: the force kernel then streams the A side and reduces it inside the warp.
    const uint32_t t_ = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = t_ < (C.act_sph ? C.nActSph : P.nSpheres);
    const uint32_t sid = (valid && C.act_sph) ? C.act_sph[t_] : t_;
    if (valid && C.keys[0][sid] == 0xffffffffu) {
        // inactive on this rank: leave empty segments behind so that later history look-ups find nothing stale
        P.ss.seg_start[sid] = 0; P.ss.seg_count[sid] = 0;
        P.sn.seg_start[sid] = 0; P.sn.seg_count[sid] = 0;
        valid = false;
    }
    const uint32_t j = valid ? C.sortedPos[sid] : 0u;
    uint32_t acc[SWEEP_MAXC];  // staged candidates: sorted index, bit 31 = spheres overlap right now
    uint32_t count = 0, countT = 0;
    uint4 meta = make_uint4(0, 0, 0, 0);
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    float myMargin = 0.f;
    if (valid) {
        const GridInfo g = *C.grid;
        me = C.sortedSph[j];
        meta = C.sortedMeta[j];
        myMargin = me.w - __ldg(&P.comp[meta.z & 0xffffu]).w;
        const uint32_t key = keys[j];
        const int cx = (int)(key % g.nbx);
        const int cy = (int)((key / g.nbx) % g.nby);
        const int cz = (int)(key / (g.nbx * g.nby));
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, (int)g.nbx - 1);
        const bool fam_on = (C.any_mask != 0) || (C.max_extra > 0.f);
        const float extraA = fam_on ? P.familyExtraMargin[meta.w] : 0.f;
        // ---- the five runs of the half stencil: all ten bounds are fetched before any of them is used ----
        uint32_t qb[5], qe[5];
#pragma unroll
        for (int r = 0; r < 5; r++) {
            // r=0: own row behind me; r=1: (dy=+1,dz=0); r=2..4: (dy=-1,0,+1; dz=+1)
            const int dy = (r == 0) ? 0 : (r == 1 ? 1 : r - 3);
            const int dz = (r < 2) ? 0 : 1;
            const int y = cy + dy, z = cz + dz;
            const bool ok = !(y < 0 || y >= (int)g.nby || z >= (int)g.nbz);
            const uint32_t row = ok ? g.nbx * ((uint32_t)y + g.nby * (uint32_t)z) : 0u;
            qb[r] = (r == 0) ? j + 1 : (ok ? __ldg(&C.cellStart[row + x0]) : 0u);
            qe[r] = ok ? __ldg(&C.cellStart[row + x1 + 1]) : 0u;
        }
        // ---- distance test, four candidates in flight at a time (independent 16-byte loads of the sorted stream) ----
#pragma unroll
        for (int r = 0; r < 5; r++) {
            for (uint32_t q = qb[r]; q < qe[r]; q += 4) {
                float4 ot[4];
#pragma unroll
                for (int u = 0; u < 4; u++) ot[u] = __ldg(&C.sortedSph[min(q + u, qe[r] - 1u)]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (q + u >= qe[r]) break;
                    const float dx = me.x - ot[u].x, dy2 = me.y - ot[u].y, dz2 = me.z - ot[u].z;
                    const float d2 = dx * dx + dy2 * dy2 + dz2 * dz2;
                    const float R = me.w + ot[u].w;
                    // superset of the double-precision test d2 <= R^2 && R - d > min(extraA, extraB)
                    if (d2 > R * R * 1.000001f + 1e-20f) continue;
                    if (count < SWEEP_MAXC) acc[count] = q + u;
                    count++;
                }
            }
        }
        if (count > SWEEP_MAXC) {
            atomicOr(&P.flags[2], 1u);  // more neighbours within reach of one sphere than can be staged
            count = SWEEP_MAXC;
        }
        // ---- owner / family filter and "in touch right now?" on the staged few (their records, four at a time) ----
        uint32_t kept = 0;
        for (uint32_t k0 = 0; k0 < count; k0 += 4) {
            uint4 om[4];
            float4 ot[4];
            uint32_t qq[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                qq[u] = acc[min(k0 + u, count - 1u)];
                om[u] = __ldg(&C.sortedMeta[qq[u]]);
                ot[u] = __ldg(&C.sortedSph[qq[u]]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (k0 + u >= count) break;
                if (om[u].x == meta.x) continue;  // same owner
                const float dx = me.x - ot[u].x, dy2 = me.y - ot[u].y, dz2 = me.z - ot[u].z;
                const float d2 = dx * dx + dy2 * dy2 + dz2 * dz2;
                const float R = me.w + ot[u].w;
                if (fam_on) {
                    if (C.any_mask && P.familyMasks[mask_pair(meta.w, om[u].w)] != 0) continue;
                    const float Rt = R - fminf(extraA, P.familyExtraMargin[om[u].w]);
                    if (d2 > Rt * Rt * 1.000001f + 1e-20f) continue;
                }
                // do the un-inflated spheres overlap right now? (only decides which list the pair goes to)
                const float Rtrue = R - myMargin - (ot[u].w - __ldg(&P.comp[om[u].z & 0xffffu]).w);
                const uint32_t touching = (d2 < Rtrue * Rtrue) ? 0x80000000u : 0u;
                acc[kept++] = qq[u] | touching;  // (kept <= k0 + u: never overtakes the read position)
                countT += touching >> 31;
            }
        }
        count = kept;
    }
    const uint32_t countN = count - countT;
    uint32_t slotT = warp_claim(countT, P.ss.count);
    uint32_t slotN = warp_claim(countN, P.sn.count);
    if (!valid) return;
    P.ss.seg_start[meta.y] = slotT;
    P.ss.seg_count[meta.y] = (slotT + countT <= C.capacity) ? countT : (slotT < C.capacity ? C.capacity - slotT : 0u);
    P.sn.seg_start[meta.y] = slotN;
    P.sn.seg_count[meta.y] = (slotN + countN <= C.capacity) ? countN : (slotN < C.capacity ? C.capacity - slotN : 0u);
    if (count == 0) return;
    if (slotT + countT > C.capacity) atomicOr(&P.flags[0], 1u);
    if (slotN + countN > C.capacity) atomicOr(&P.flags[0], 4u);
    const uint32_t nM = P.nMat;
    for (uint32_t k = 0; k < count; k++) {
        const bool touching = (acc[k] >> 31) != 0u;
        const uint4 om = __ldg(&C.sortedMeta[acc[k] & 0x7fffffffu]);
        const ContactList& L = touching ? P.ss : P.sn;
        const uint32_t slot = touching ? slotT++ : slotN++;
        if (slot >= C.capacity) continue;
        // The history carry-over is NOT done here: per sphere it is a doubly nested search of very uneven length (the
        // warp would run at 5-9 active lanes of 32); k_history does it with one thread per emitted contact instead.
        uint32_t skip = 0;
        if (!touching) {
            // A candidate that is clearly apart at these very positions (float positions: allow for their rounding)
            // would have its history destroyed by the force pass that follows this rebuild (no overlap => wildcards
            // zeroed, DEMCalcForceKernels.cu:258-261): it carries none over, so there is nothing to look up or write.
            const float4 ot = __ldg(&C.sortedSph[acc[k] & 0x7fffffffu]);
            const float dx = me.x - ot.x, dy = me.y - ot.y, dz = me.z - ot.z;
            const float Rtrue = (me.w - myMargin) + __ldg(&P.comp[om.z & 0xffffu]).w;
            const float slack = 3e-7f * (fabsf(me.x) + fabsf(me.y) + fabsf(me.z)) + 1e-8f;
            if (sqrtf(dx * dx + dy * dy + dz * dz) * 0.999999f - Rtrue * 1.000001f - slack > 0.f) skip = CINFO_NO_HISTORY;
        }
        const uint32_t matpair = (meta.z >> 16) * nM + (om.z >> 16);
        L.pair[slot] = make_uint2(meta.y, om.y);
        L.cinfo[slot] = make_uint4(meta.x, om.x, (meta.z & 0xffffu) | ((om.z & 0xffffu) << 16), matpair | skip);
    }
}

// History carry-over (DEMHistoryMappingKernels.cu), one thread per contact of the two new sphere--sphere lists: the pair
// may sit in either previous list as (A,B) or -- when the two spheres swapped their order in the sorted array -- as
// (B,A); then delta_tan changes sign.  Adjacent threads hold the contacts of the same sphere A (the lists are
// sphere-major), so their segment look-ups coalesce.
__global__ void __launch_bounds__(256) k_history(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    const uint32_t nT = min(*P.ss.count, C.capacity);
    const uint32_t n = nT + min(*P.sn.count, C.capacity);
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
        const uint2 pr = L.pair[slot];
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t alive = 0;
        bool found = false;
#pragma unroll 1
        for (int pass = 0; pass < 4 && !found; pass++) {
            const ContactList& O = ((pass & 1) == (touching ? 0 : 1)) ? C.oldss : C.oldsn;  // likelier list first
            const bool flipped = pass >= 2;
            const uint32_t a = flipped ? pr.y : pr.x, b = flipped ? pr.x : pr.y;
            const uint32_t os = O.seg_start[a], oc = O.seg_count[a];
            for (uint32_t t = 0; t < oc; t++) {
                if (O.pair[os + t].y == b) {
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

__global__ void k_finish_counts(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t a = *P.ss.count, b = *P.sa.count, c = *P.sn.count, d = *P.st.count;
        P.ss.count[1] = a;  // the unclamped demand, read back by the host to size a regrow
        P.sa.count[1] = b;
        P.sn.count[1] = c;
        P.st.count[1] = d;
        if (d > C.capacity) { atomicOr(&P.flags[0], 32u); d = C.capacity; }
        *P.st.count = d;
        if (a > C.capacity) { atomicOr(&P.flags[0], 1u); a = C.capacity; }
        if (b > C.capacity) { atomicOr(&P.flags[0], 2u); b = C.capacity; }
        if (c > C.capacity) { atomicOr(&P.flags[0], 4u); c = C.capacity; }
        *P.ss.count = a;
        *P.sa.count = b;
        *P.sn.count = c;
    }
}

// ---------------------------------------------------------------------------------------------------------------
int launch_cd_prepare(const DevParams& P, const CdParams& C, bool need_maxvel, int stage, cudaStream_t s) {
    int launches = 0;
    if (stage == 0) {
        if (need_maxvel) {
            // velocities changed outside the integrator (initial state / host upload): recompute max |v|
            cudaMemsetAsync(P.maxvel, 0, sizeof(float), s);
            if (P.nOwners) {
                k_maxvel<<<(P.nOwners + 255) / 256, 256, 0, s>>>(P, P.errOutVel);
                launches++;
            }
        }
    } else if (stage == 1) {
        k_grid_setup<<<1, 32, 0, s>>>(P, C);
        launches++;
    } else {
        if (P.nAnal) {
            k_anal_prep<<<(P.nAnal + 63) / 64, 64, 0, s>>>(P, C);
            launches++;
        }
        cudaMemsetAsync(C.cellStart, 0, sizeof(uint32_t) * (size_t)C.scan_cells, s);
        cudaMemsetAsync(P.ss.count, 0, sizeof(uint32_t) * 4, s);
        cudaMemsetAsync(P.sn.count, 0, sizeof(uint32_t) * 4, s);
        cudaMemsetAsync(P.sa.count, 0, sizeof(uint32_t) * 4, s);
        if (C.act_sph && P.nTri)  // the per-sphere triangle pass walks all spheres and skips those without a key
            cudaMemsetAsync(C.keys[0], 0xff, sizeof(uint32_t) * (size_t)P.nSpheres, s);
        if (C.act_sph) {
            // spheres this rank does not hold are not visited: leave empty segments behind for them, so that later
            // history look-ups find nothing stale
            cudaMemsetAsync(P.ss.seg_count, 0, sizeof(uint32_t) * ((size_t)P.nSpheres + 1), s);
            cudaMemsetAsync(P.sn.seg_count, 0, sizeof(uint32_t) * ((size_t)P.nSpheres + 1), s);
            cudaMemsetAsync(P.sa.seg_count, 0, sizeof(uint32_t) * ((size_t)P.nSpheres + 1), s);
        }
        const uint32_t n = C.act_sph ? C.nActSph : P.nSpheres;
        if (n) {
            k_sphere_prep<<<(n + 255) / 256, 256, 0, s>>>(P, C);
            launches++;
        }
    }
    return launches;
}

int launch_cd_sweep(const DevParams& P, const CdParams& C, int sorted_buf, cudaStream_t s, cudaEvent_t* ev, bool sort_only) {
    int launches = 0;
    const uint32_t n = C.act_sph ? C.nActSph : P.nSpheres;
    // cell histogram -> exclusive prefix (ncells+1 entries; on a single GPU the host does not know the grid of this
    // rebuild yet and scans the full capacity, which keeps the launch shape static)
    launches += launch_scan_exclusive(C.cellStart, C.scan_cells, C.scan_tmp, nullptr, s);
    if (ev) cudaEventRecord(ev[0], s);
    if (n) {
        if (sorted_buf < 0) {  // counting sort
            k_cs_scatter<<<(n + 255) / 256, 256, 0, s>>>(P, C);
            k_gather_sorted_cs<<<(n + 255) / 256, 256, 0, s>>>(P, C);
            launches++;
        } else {
            k_gather_sorted<<<(n + 255) / 256, 256, 0, s>>>(P, C, C.vals[sorted_buf]);
        }
        if (ev) cudaEventRecord(ev[1], s);
        if (sort_only) return launches + 1;
        k_sweep<<<(n + 127) / 128, 128, 0, s>>>(P, C, sorted_buf < 0 ? C.vals[0] : C.keys[sorted_buf]);
        k_history<<<148 * 8, 256, 0, s>>>(P, C);
        launches += 3;
    } else if (ev) {
        cudaEventRecord(ev[1], s);
    }
    if (ev) cudaEventRecord(ev[2], s);
    k_finish_counts<<<1, 32, 0, s>>>(P, C);
    launches++;
    if (ev) cudaEventRecord(ev[3], s);
    return launches;
}

// ---------------------------------------------------------------------------------------------------------------
// reductions over clump owners (DEMInspector built-ins, AuxClasses.cpp:88-164)
__global__ void k_reduce(const __grid_constant__ DevParams P, int kind, uint32_t nClumps, double* out) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    double v = (kind == DEM_REDUCE_MIN_Z) ? 1e300 : ((kind == DEM_REDUCE_MAX_Z) ? -1e300 : 0.0);
    if (o < nClumps && (!P.active || P.active[o] == 1)) {
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
    if (o < nClumps && (!P.active || P.active[o] == 1)) {
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
