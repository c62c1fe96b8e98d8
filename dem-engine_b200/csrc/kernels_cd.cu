// kernels_cd.cu -- contact-list rebuild ("kinematic" work) for sm_100a, fully device-driven (no host round trips
// between the stages):
//   k_maxvel / k_grid_setup      max |v| -> margin -> broad-phase cell size and grid, decided ON the device
//   k_sphere_prep                per sphere: world position (fixed-point decode + rotated offset), inflated radius,
//                                cell key + cell histogram, sphere--analytical candidate count
//   radix sort (k_rs_*)          LSD, 8-bit digits, tile ranking with warp match + shared-memory staged coalesced scatter
//   scan (k_scan_*)              exclusive prefix sums (cell table, per-sphere contact offsets, sort histograms)
//   k_gather_sorted              cell-ordered float4 {x,y,z,r'} + {owner,id}
//   k_sweep<FILL>                27-cell sweep over 9 contiguous rows; count pass + fill pass; the fill pass compiles the
//                                per-contact record and carries the Hertz-Mindlin history over from the previous list
//   k_sa_fill                    sphere--analytical list + history
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
    if (o < P.nOwners) {
        const float4 v = P.state[o].vel;
        a = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
        if (!isfinite(a) || a > errOutVel) atomicOr(&P.flags[1], 1u);
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
    *C.grid = g;
}

// sphere--analytical candidate test with inflated geometry (DEMBinSphereKernels.cu:78-128). Conservative in float.
__device__ __forceinline__ bool sa_candidate(const DevParams& P, const AnalObj& ob, float3 sp /*LBF-rel*/, float rInfl,
                                             uint32_t famS, bool any_mask) {
    const OwnerState* sb = P.state + ob.owner;
    const OwnerPos pB = sb->pos;
    if (any_mask && P.familyMasks[mask_pair(famS, pB.family)] != 0) return false;
    const float4 qB = sb->quat;
    const float4 vB = sb->vel;
    const float mB = owner_margin(P, sqrtf(vB.x * vB.x + vB.y * vB.y + vB.z * vB.z), pB.family);
    double X, Y, Z;
    pos_decode(pB, P, X, Y, Z);
    const float3 rel = rotate(f3(ob.relx, ob.rely, ob.relz), qB);
    const float3 dir = rotate(f3(ob.rotx, ob.roty, ob.rotz), qB);
    const float3 d = f3(sp.x - (float)(X + (double)rel.x), sp.y - (float)(Y + (double)rel.y),
                        sp.z - (float)(Z + (double)rel.z));
    const float thr = fminf(P.familyExtraMargin[famS], P.familyExtraMargin[pB.family]);
    const float slack = 1e-6f * (fabsf(sp.x) + fabsf(sp.y) + fabsf(sp.z) + 1.f);
    float depth;
    if (ob.type == DEM_ANAL_PLANE) {
        depth = rInfl + mB - dot(d, dir);
    } else if (ob.type == DEM_ANAL_CYL_INF) {
        const float3 s2c = f3(-d.x, -d.y, -d.z);
        const float proj = dot(s2c, dir);
        const float3 radial = s2c - proj * dir;
        const float cyl_rad = ob.size1 - ob.normal_sign * mB;
        depth = rInfl - ob.normal_sign * (cyl_rad - length(radial));
    } else {
        return false;
    }
    return depth + slack > thr;
}

__global__ void __launch_bounds__(256) k_sphere_prep(const __grid_constant__ DevParams P,
                                                     const __grid_constant__ CdParams C) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nSpheres) return;
    const GridInfo g = *C.grid;
    const uint2 s = P.sph[i];
    const OwnerState* st = P.state + s.x;
    const OwnerPos pos = st->pos;
    const float4 q = st->quat;
    const float4 v = st->vel;
    const float4 comp = __ldg(&P.comp[s.y & 0xffffu]);
    const float margin = owner_margin(P, sqrtf(v.x * v.x + v.y * v.y + v.z * v.z), pos.family);
    double X, Y, Z;
    pos_decode(pos, P, X, Y, Z);
    const float3 rel = rotate(f3(comp.x, comp.y, comp.z), q);
    const float3 sp = f3((float)(X + (double)rel.x), (float)(Y + (double)rel.y), (float)(Z + (double)rel.z));
    const float rInfl = comp.w + margin;
    C.sphF[i] = make_float4(sp.x, sp.y, sp.z, rInfl);
    int cx = (int)floorf(sp.x * g.inv_cs), cy = (int)floorf(sp.y * g.inv_cs), cz = (int)floorf(sp.z * g.inv_cs);
    cx = min(max(cx, 0), (int)g.nbx - 1);
    cy = min(max(cy, 0), (int)g.nby - 1);
    cz = min(max(cz, 0), (int)g.nbz - 1);
    const uint32_t key = (uint32_t)cx + g.nbx * ((uint32_t)cy + g.nby * (uint32_t)cz);
    C.keys[0][i] = key;
    C.vals[0][i] = i;
    atomicAdd(&C.cellStart[key], 1u);
    uint32_t nsa = 0;
    for (uint32_t k = 0; k < P.nAnal; k++)
        if (sa_candidate(P, P.anal[k], sp, rInfl, pos.family, C.any_mask != 0)) nsa++;
    C.saCnt[i] = nsa;
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
    C.sortedMeta[j] = make_uint2(P.sph[i].x, i);
}

// The 27-cell sweep.  Cells are x-fastest, so the 3 x-neighbours of a row are ONE contiguous run of the sorted
// array: 9 runs per sphere, each delimited by two reads of the exclusive-prefix cell table.
template <bool FILL>
__global__ void __launch_bounds__(128) k_sweep(const __grid_constant__ DevParams P,
                                               const __grid_constant__ CdParams C,
                                               const uint32_t* __restrict__ keys) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.nSpheres) return;
    const GridInfo g = *C.grid;
    const float4 me = C.sortedSph[j];
    const uint2 meta = C.sortedMeta[j];
    const uint32_t key = keys[j];
    const int cx = (int)(key % g.nbx);
    const int cy = (int)((key / g.nbx) % g.nby);
    const int cz = (int)(key / (g.nbx * g.nby));
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, (int)g.nbx - 1);
    const bool any_mask = C.any_mask != 0;
    uint32_t famA = 0;
    float extraA = 0.f;
    if (any_mask || C.max_extra > 0.f) {
        famA = P.state[meta.x].pos.family;
        extraA = P.familyExtraMargin[famA];
    }
    uint32_t count = 0;
    uint32_t out = 0, oldStart = 0, oldCount = 0, myComp = 0;
    if (FILL) {
        out = C.cnt[j];
        oldStart = C.oldss.seg_start[meta.y];
        oldCount = C.oldss.seg_count[meta.y];
        myComp = P.sph[meta.y].y;
    }
    for (int dz = -1; dz <= 1; dz++) {
        const int z = cz + dz;
        if (z < 0 || z >= (int)g.nbz) continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int y = cy + dy;
            if (y < 0 || y >= (int)g.nby) continue;
            const uint32_t row = g.nbx * ((uint32_t)y + g.nby * (uint32_t)z);
            const uint32_t qb = C.cellStart[row + x0], qe = C.cellStart[row + x1 + 1];
            for (uint32_t q = qb; q < qe; q++) {
                const uint2 om = C.sortedMeta[q];
                // A is the sphere with the smaller id (the reference emits ascending sphere ids within a bin)
                if (om.y <= meta.y || om.x == meta.x) continue;
                const float4 ot = C.sortedSph[q];
                const float dx = me.x - ot.x, dy2 = me.y - ot.y, dz2 = me.z - ot.z;
                const float d2 = dx * dx + dy2 * dy2 + dz2 * dz2;
                const float R = me.w + ot.w;
                // superset of the double-precision test d2 <= R^2 && R - d > min(extraA, extraB)
                float Rt = R;
                uint32_t famB = 0;
                if (any_mask || C.max_extra > 0.f) {
                    famB = P.state[om.x].pos.family;
                    if (any_mask && P.familyMasks[mask_pair(famA, famB)] != 0) continue;
                    Rt = R - fminf(extraA, P.familyExtraMargin[famB]);
                }
                if (d2 > Rt * Rt * 1.000001f + 1e-20f) continue;
                if (FILL) {
                    const uint32_t slot = out + count;
                    if (slot < C.capacity) {
                        const uint32_t compB = P.sph[om.y].y;
                        const uint32_t nM = P.nMat;
                        const uint32_t matpair = (myComp >> 16) * nM + (compB >> 16);
                        // history carry-over: look (A,B) up in A's segment of the previous list
                        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
                        uint32_t alive = 0;
                        for (uint32_t t = 0; t < oldCount; t++) {
                            if (C.oldss.pair[oldStart + t].y == om.y) {
                                alive = C.oldss.cinfo[oldStart + t].w & 0x80000000u;
                                if (alive && C.oldss.hist) h = C.oldss.hist[oldStart + t];
                                break;
                            }
                        }
                        P.ss.pair[slot] = make_uint2(meta.y, om.y);
                        P.ss.cinfo[slot] =
                            make_uint4(meta.x, om.x, (myComp & 0xffffu) | ((compB & 0xffffu) << 16), matpair | alive);
                        if (P.ss.hist) P.ss.hist[slot] = h;
                    }
                }
                count++;
            }
        }
    }
    if (FILL) {
        P.ss.seg_start[meta.y] = out;
        P.ss.seg_count[meta.y] = (out + count <= C.capacity) ? count : (out < C.capacity ? C.capacity - out : 0u);
    } else {
        C.cnt[j] = count;
    }
}

__global__ void __launch_bounds__(256) k_sa_fill(const __grid_constant__ DevParams P,
                                                 const __grid_constant__ CdParams C) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nSpheres) return;
    const uint32_t out = C.saCnt[i], end = C.saCnt[i + 1];
    P.sa.seg_start[i] = out;
    P.sa.seg_count[i] = (end <= C.capacity) ? end - out : (out < C.capacity ? C.capacity - out : 0u);
    if (end == out) return;
    const float4 me = C.sphF[i];
    const uint2 s = P.sph[i];
    const uint32_t fam = P.state[s.x].pos.family;
    const uint32_t oldStart = C.oldsa.seg_start[i], oldCount = C.oldsa.seg_count[i];
    uint32_t k = 0;
    for (uint32_t ob = 0; ob < P.nAnal; ob++) {
        const AnalObj a = P.anal[ob];
        if (!sa_candidate(P, a, f3(me.x, me.y, me.z), me.w, fam, C.any_mask != 0)) continue;
        const uint32_t slot = out + k;
        if (slot < C.capacity) {
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t alive = 0;
            for (uint32_t t = 0; t < oldCount; t++) {
                if (C.oldsa.pair[oldStart + t].y == ob) {
                    alive = C.oldsa.cinfo[oldStart + t].w & 0x80000000u;
                    if (alive && C.oldsa.hist) h = C.oldsa.hist[oldStart + t];
                    break;
                }
            }
            const uint32_t matpair = (s.y >> 16) * P.nMat + a.material;
            P.sa.pair[slot] = make_uint2(i, ob);
            P.sa.cinfo[slot] = make_uint4(s.x, ob, s.y & 0xffffu, matpair | alive);
            if (P.sa.hist) P.sa.hist[slot] = h;
        }
        k++;
    }
}

__global__ void k_finish_counts(const __grid_constant__ DevParams P, const __grid_constant__ CdParams C,
                                const uint32_t* __restrict__ ssTotal, const uint32_t* __restrict__ saTotal) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t a = *ssTotal, b = *saTotal;
        if (a > C.capacity) { atomicOr(&P.flags[0], 1u); a = C.capacity; }
        if (b > C.capacity) { atomicOr(&P.flags[0], 2u); b = C.capacity; }
        *P.ss.count = a;
        *P.sa.count = b;
    }
}

// ---------------------------------------------------------------------------------------------------------------
int launch_cd_prepare(const DevParams& P, const CdParams& C, cudaStream_t s) {
    int launches = 0;
    cudaMemsetAsync(P.maxvel, 0, sizeof(float), s);
    if (P.nOwners) {
        k_maxvel<<<(P.nOwners + 255) / 256, 256, 0, s>>>(P, 3.0e38f);
        launches++;
    }
    k_grid_setup<<<1, 32, 0, s>>>(P, C);
    launches++;
    cudaMemsetAsync(C.cellStart, 0, sizeof(uint32_t) * ((size_t)C.max_cells + 1), s);
    if (P.nSpheres) {
        k_sphere_prep<<<(P.nSpheres + 255) / 256, 256, 0, s>>>(P, C);
        launches++;
    }
    return launches;
}

int launch_cd_sweep(const DevParams& P, const CdParams& C, int sorted_buf, cudaStream_t s, cudaEvent_t* ev) {
    int launches = 0;
    const uint32_t n = P.nSpheres;
    // cell histogram -> exclusive prefix (ncells+1 entries; scanning the full capacity keeps the launch shape static)
    launches += launch_scan_exclusive(C.cellStart, C.max_cells + 1, C.scan_tmp, nullptr, s);
    // totals live right behind the per-sphere offsets: cnt[n], saCnt[n]
    if (n) {
        k_gather_sorted<<<(n + 255) / 256, 256, 0, s>>>(P, C, C.vals[sorted_buf]);
        if (ev) cudaEventRecord(ev[0], s);
        k_sweep<false><<<(n + 127) / 128, 128, 0, s>>>(P, C, C.keys[sorted_buf]);
        if (ev) cudaEventRecord(ev[1], s);
        launches += 2;
        launches += launch_scan_exclusive(C.cnt, n, C.scan_tmp, C.cnt + n, s);
        launches += launch_scan_exclusive(C.saCnt, n, C.scan_tmp, C.saCnt + n, s);
        if (ev) cudaEventRecord(ev[2], s);
        k_sweep<true><<<(n + 127) / 128, 128, 0, s>>>(P, C, C.keys[sorted_buf]);
        if (ev) cudaEventRecord(ev[3], s);
        k_sa_fill<<<(n + 255) / 256, 256, 0, s>>>(P, C);
        launches += 2;
    } else {
        cudaMemsetAsync(C.cnt, 0, sizeof(uint32_t), s);
        cudaMemsetAsync(C.saCnt, 0, sizeof(uint32_t), s);
        if (ev) for (int k = 0; k < 4; k++) cudaEventRecord(ev[k], s);
    }
    k_finish_counts<<<1, 32, 0, s>>>(P, C, C.cnt + n, C.saCnt + n);
    launches++;
    return launches;
}

// ---------------------------------------------------------------------------------------------------------------
// reductions over clump owners (DEMInspector built-ins, AuxClasses.cpp:88-164)
__global__ void k_reduce(const __grid_constant__ DevParams P, int kind, uint32_t nClumps, double* out) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    double v = (kind == DEM_REDUCE_MIN_Z) ? 1e300 : ((kind == DEM_REDUCE_MAX_Z) ? -1e300 : 0.0);
    if (o < nClumps) {
        const OwnerState st = P.state[o];
        if (kind == DEM_REDUCE_MAX_ABSV) {
            v = sqrtf(st.vel.x * st.vel.x + st.vel.y * st.vel.y + st.vel.z * st.vel.z);
        } else if (kind == DEM_REDUCE_MAX_Z || kind == DEM_REDUCE_MIN_Z) {
            double X, Y, Z;
            pos_decode(st.pos, P, X, Y, Z);
            v = Z + (double)P.LBF[2];
        } else if (kind == DEM_REDUCE_KINETIC_ENERGY) {
            const float4 mp = P.massprop[__float_as_uint(st.omg.w)];
            v = 0.5 * (double)st.vel.w * ((double)st.vel.x * st.vel.x + (double)st.vel.y * st.vel.y +
                                         (double)st.vel.z * st.vel.z) +
                0.5 * ((double)mp.y * st.omg.x * st.omg.x + (double)mp.z * st.omg.y * st.omg.y +
                       (double)mp.w * st.omg.z * st.omg.z);
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

int launch_reduce(const DevParams& P, int kind, double* d_out, cudaStream_t s) {
    // clump owners are the owners that have spheres: the caller passes nOwners restricted to clumps via P.nOwners
    const uint32_t n = P.nOwners;
    if (n == 0) return 0;
    k_reduce<<<(n + 255) / 256, 256, 0, s>>>(P, kind, n, d_out);
    return 1;
}

}  // namespace demb
