// dem_device.cuh -- device data layout and math helpers of the B200-native DEM core.
//
// HBM layout (see DESIGN.md "Data layout"):
//   owners  : one 64-byte record per owner, fetched with TWO 256-bit loads (one per 32-byte sector)
//             sector 0: pos  {u64 voxelID; u16 locX,locY,locZ; u8 family; u8 flags}   (the reference's voxelID / locX /
//                            locY / locZ / familyID arrays, src/DEM/Defines.h:272-280, packed) + quat {w,x,y,z};
//                            flags = 255 - c, c the owner's contact margin of the last rebuild in 256ths of the largest
//                            one (rounded up; 0 = "as large as the largest"): k_sphere_prep writes it, k_force_ss reads
//                            it with the record it gathers anyway
//             sector 1: vel {vx,vy,vz,mass} + omg {WORLD-frame angular velocity, unused}
//           + a 16-byte spin record {body-frame omgBar xyz, bits(inertiaPropOffset)} that only the integrator streams
//           + one 32-byte wrench accumulator {Fx,Fy,Fz,0 | Tx,Ty,Tz,0}, force AND torque in the world frame (the
//             integrator rotates the summed torque into the body frame once per owner), the target of 128-bit vector
//             reductions (red.global.add.v4.f32, sm_90+).
//   spheres : uint2 {owner, comp | material<<16}
//   contacts: four lists (sphere-sphere in touch, sphere-sphere candidates, sphere-analytical, sphere-triangle: no
//             per-contact type byte), each a compiled 16-byte record {ownerA, ownerB|obj, compA|compB<<16,
//             matpair|alive<<31} + float4 history {delta_tan_xyz, delta_time} (+ uint2 {geoA, geoB} for the rebuild, and
//             the optional force / contact-point record).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace demb {

struct __align__(16) OwnerPos {
    unsigned long long voxel;
    unsigned short lx, ly, lz;
    unsigned char family;
    unsigned char flags;
};
static_assert(sizeof(OwnerPos) == 16, "OwnerPos must be one 128-bit word");

// 64-byte owner record: one 128-byte line holds two owners; a gather costs exactly two 32-byte sectors
struct __align__(64) OwnerState {
    OwnerPos pos;
    float4 quat;  // w,x,y,z
    float4 vel;   // vx,vy,vz, mass
    float4 omg;   // WORLD-frame angular velocity R(q) omgBar (w unused); body-frame omgBar lives in DevParams::spin
};
static_assert(sizeof(OwnerState) == 64, "OwnerState must be 64 bytes");

struct __align__(32) Wrench {
    float4 f;  // world-frame force sum
    float4 t;  // world-frame torque sum
};

// per material pair, precomputed on the host exactly as matProxy2ContactParam<float> and the beta expression of
// FullHertzianForceModel.cu:59-60 evaluate them
struct __align__(16) MatPair {
    float E_cnt, G_cnt, beta, mu;
    float Crr, CoR, pad0, pad1;
};

struct __align__(16) AnalObj {
    float relx, rely, relz;  // component position in owner frame
    float rotx, roty, rotz;  // plane normal / cylinder axis in owner frame
    float size1, size2, size3;
    float normal_sign;  // objNormal
    float mass;
    uint32_t owner;
    uint32_t type;
    uint32_t material;
    uint32_t pad0, pad1;
};

struct Prescr {  // == DemPrescription (88 bytes)
    uint8_t used;
    uint8_t linVelPrescribed[3];
    uint8_t rotVelPrescribed[3];
    uint8_t linPosPrescribed[3];
    uint8_t rotPosPrescribed;
    uint8_t hasLinVel[3];
    uint8_t hasRotVel[3];
    uint8_t hasLinPos[3];
    uint8_t hasAcc[3];
    uint8_t hasAngAcc[3];
    uint8_t pad_[2];
    float linVel[3];
    float rotVel[3];
    float linPos[3];
    float acc[3];
    float angAcc[3];
};
static_assert(sizeof(Prescr) == 88, "Prescr must match DemPrescription");

// one contact list.  Geometry A of a contact is the sphere whose segment [seg_start, seg_start + seg_count) it lies in.
struct ContactList {
    uint32_t* idB;       // geometry B: sphere id (analytical component / triangle id for the other lists)
    uint4* cinfo;        // compiled record {ownerA, ownerB|objID, compA | compB<<16, matpair | alive<<31}
    float4* hist;        // delta_tan_xyz, delta_time
    uint32_t* seg_start; // per sphere A: first contact of A in this list
    uint32_t* seg_count; // per sphere A: number of contacts of A
    uint32_t* count;     // device-resident number of contacts: [0] clamped to the capacity, [1] demand
    float4* force;       // optional per-contact force record (xyz) -- nullptr when SetNoForceRecord
    float4* cpoint;      // ... and the contact point (world frame, LBF-relative) the force acts at
    uint8_t* due;        // sphere--sphere candidate list only: the step of the list's cycle from which on the pair has to
                         // be looked at (set by the sweep, refreshed by k_force_ss); nullptr for the other lists
};

// status words of a context (DevParams::flags)
enum {
    DEM_FLAG_CAPACITY = 0,  // bit mask: which list / table overflowed in the last rebuild
    DEM_FLAG_TRI_DEMAND = 1,  // entries the triangle--cell table would have needed
    DEM_FLAG_HALO = 2,      // multi-GPU: a halo buffer overflowed / a neighbour stopped answering
    DEM_FLAG_VELOCITY = 3,  // an owner has a non-finite or too large velocity
    DEM_FLAG_POISON = 4,    // != 0: a rebuild failed; every kernel of every later step / rebuild is a no-op until the
                            // host has grown the lists and cleared it (value = sequence number of the failed rebuild)
    DEM_FLAG_SEQ = 5,       // rebuilds finished so far (device-side counter: graph replays carry no host arguments)
    DEM_FLAG_CYCLE_STEP = 6,  // integrations since the last rebuild (zeroed by the rebuild, advanced by the integrator)
    DEM_FLAG_INV_CLOSING = 7,  // float bits: maxDrift / (2 * largest margin of the last rebuild) = 1 / the most by which
                               // the gap of any pair can shrink in one step (0: fixed expand factor, no such bound)
    DEM_NUM_FLAGS = 8
};
constexpr uint32_t CINFO_NO_HISTORY = 0x40000000u;  // sweep -> k_history: this contact carries no history over
// cinfo.w: bits 0-15 material pair, bit 30 CINFO_NO_HISTORY (rebuild only), bit 31 alive

// Multi-GPU (slab decomposition) state as the kernels see it.  Everything that changes from step to step or rebuild
// to rebuild lives in DEVICE memory (epoch counters, counts, lists), so that the same parameter block -- hence the same
// CUDA graph -- serves every step and every rebuild.  Layout of a rank's peer-visible block (mapped by all ranks):
//   [0, 1024)                      header: see the MG_* offsets below
//   MG_OFF_GID  + ((par*2 + dir) * cap) * 4      halo membership lists written by the neighbour in direction dir
//   MG_OFF_REC  + ((dir*2 + half) * cap) * 80    {state, spin} records written by the neighbour in direction dir
constexpr uint32_t MG_MAX_WORLD = 8;
constexpr uint32_t MG_HDR_STEP_FLAG = 0;      // u64[2]: neighbour `dir` has completed exchange number e
constexpr uint32_t MG_HDR_RECV_COUNT = 16;    // u32[2 par][2 dir]: owners the neighbour sends me during this cycle
constexpr uint32_t MG_HDR_MAIL = 64;          // all-gather mailbox: [2 parity][8 ranks] x {u64 seq, u32 val[2]}
constexpr uint32_t MG_HDR_BYTES = 1024;
struct MgDev {
    int rank, world;
    int has[2];                       // neighbour present on the left / right
    char* my_block;
    char* peer_block[MG_MAX_WORLD];   // every rank's block (peer_block[rank] == my_block)
    uint32_t cap;                     // owners per halo list
    unsigned long long* epoch;        // exchanges (per-step and per-rebuild alike) completed so far
    unsigned long long* mail_ctr;     // all-gathers completed so far
    uint32_t* block_ctr;              // [0] pull kernel, [1] push kernel: last-block detection
    uint8_t* flag;                    // per global owner: 0 unknown here, 1 own, 2 ghost, 3 / 4 own + halo (see DevParams::active)
    uint32_t* active_list[2];         // [par] compact list of active owners (own + ghost) of the cycle with parity par
    uint32_t* counts[2];              // [par] {own, send-left, send-right, active, active spheres, -, -, -}
    uint32_t* send_gid[2][2];         // [par][dir] own owners inside the halo of the left / right cut
    uint32_t* act_sph[2];             // [par] spheres of the active owners of that cycle (the rebuild walks these only)
    const uint2* owner_sph;           // per owner {first sphere, number of spheres} (nullptr: not contiguous)
    float cut_lo, cut_hi;             // my slab in LBF-relative x
    uint32_t nClumpOwners;            // owners >= this index are analytical / mesh owners, replicated on every rank
};
__host__ __device__ __forceinline__ size_t mg_off_gid(uint32_t cap, int par, int dir) {
    return (size_t)MG_HDR_BYTES + ((size_t)(par * 2 + dir) * cap) * 4u;
}
__host__ __device__ __forceinline__ size_t mg_off_rec(uint32_t cap, int dir, int half) {
    return (size_t)MG_HDR_BYTES + (size_t)4u * cap * 4u + ((size_t)(dir * 2 + half) * cap) * 80u;
}
__host__ __device__ __forceinline__ size_t mg_block_bytes(uint32_t cap) { return mg_off_rec(cap, 2, 0); }

// Everything a kernel needs, passed by value (__grid_constant__)
struct DevParams {
    // position code
    uint32_t nvXp2, nvYp2;
    double l, voxelSize, inv_l;
    float LBF[3];
    float G[3];
    float h;
    float half_h;  // (float)(0.5*h)
    uint32_t integrator;
    uint32_t nOwners, nSpheres, nAnal, nTri, nMat;
    // margin policy
    float beta, approxMaxVel, expSafetyMulti, expSafetyAdder;
    float errOutVel;
    uint32_t maxDrift;
    uint32_t fast_encode;        // integrator: division-free position encode
    uint32_t force_opts;         // sphere--sphere force kernels: see DemCtx::force_opts (dem_core.cu)
    double inv_voxelSize;
    // owners
    OwnerState* state;
    float4* spin;     // body-frame angular velocity omgBar xyz, w = bits(inertiaPropOffset)
    Wrench* wrench;
    Wrench* acc_out;  // optional per-owner {a, alpha} read-out (nullptr = off)
    // domain decomposition (nullptr on a single GPU): per-owner activity flag (0 unknown, 1 own, 2 ghost), the
    // compact list of active owners the integrator walks and its DEVICE-resident length
    const uint8_t* active;       // 0 unknown here, 1 own, 2 ghost, 3 own + in the left halo list, 4 own + in the right one only
    const uint32_t* active_list;
    const uint32_t* nActivePtr;
    // the owners this rank sends to its left / right neighbour every step (integrated first, so that the exchange can
    // run beside the integration of the rest) and their DEVICE-resident counts {-, left, right}
    const uint32_t* halo_gid[2];
    const uint32_t* halo_counts;
    // spheres / templates
    const uint2* sph;
    const float4* comp;      // {relx, rely, relz, radius}
    const float4* massprop;  // {mass, moiX, moiY, moiZ}
    const MatPair* matpair;  // nMat x nMat
    const AnalObj* anal;
    const float4* tri_n1;  // triangle nodes in owner frame (xyz, w unused)
    const float4* tri_n2;
    const float4* tri_n3;
    const uint2* tri_info;  // {ownerMesh, material}
    const uint8_t* familyMasks;
    const float* familyExtraMargin;
    const Prescr* presc;
    // contact lists: ss = sphere-sphere pairs in touch at the last rebuild, sn = the remaining sphere-sphere
    // candidates, sa = sphere-analytical, st = sphere-triangle
    ContactList ss, sn, sa, st;
    // status words (device), indexed by DEM_FLAG_*
    uint32_t* flags;
    float* maxvel;       // device float: max |v| of the current state (kept up to date by the integrator)
    float* maxvel_next;  // the slot the NEXT step will accumulate into (zeroed by this step's integrator)
};

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length(float3 a) { return sqrtf(dot(a, a)); }

// applyOriQToVector3 (reference src/kernel/DEMHelperKernels.cuh:161-173); q = {w,x,y,z}
__device__ __forceinline__ float3 rotate(float3 v, float4 q) {
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    float3 r;
    r.x = (2.0f * (w * w + x * x) - 1.0f) * v.x + (2.0f * (x * y - w * z)) * v.y + (2.0f * (x * z + w * y)) * v.z;
    r.y = (2.0f * (x * y + w * z)) * v.x + (2.0f * (w * w + y * y) - 1.0f) * v.y + (2.0f * (y * z - w * x)) * v.z;
    r.z = (2.0f * (x * z - w * y)) * v.x + (2.0f * (y * z + w * x)) * v.y + (2.0f * (w * w + z * z) - 1.0f) * v.z;
    return r;
}
__device__ __forceinline__ float3 rotate_inv(float3 v, float4 q) {
    return rotate(v, make_float4(q.x, -q.y, -q.z, -q.w));
}

// integer world coordinates in units of l (voxel index * 2^16 + sub-voxel), reference DEMHelperKernels.cuh:91-134
__device__ __forceinline__ void pos_ints(const OwnerPos& p, uint32_t nvXp2, uint32_t nvYp2, long long& ix,
                                         long long& iy, long long& iz) {
    const unsigned long long vx = p.voxel & ((1ull << nvXp2) - 1ull);
    const unsigned long long vy = (p.voxel >> nvXp2) & ((1ull << nvYp2) - 1ull);
    const unsigned long long vz = p.voxel >> (nvXp2 + nvYp2);
    ix = (long long)((vx << 16) | p.lx);
    iy = (long long)((vy << 16) | p.ly);
    iz = (long long)((vz << 16) | p.lz);
}
// decode exactly as voxelIDToPosition<double>: X = vx*voxelSize + sub*l (LBF-relative)
__device__ __forceinline__ void pos_decode(const OwnerPos& p, const DevParams& P, double& X, double& Y, double& Z) {
    const unsigned long long vx = p.voxel & ((1ull << P.nvXp2) - 1ull);
    const unsigned long long vy = (p.voxel >> P.nvXp2) & ((1ull << P.nvYp2) - 1ull);
    const unsigned long long vz = p.voxel >> (P.nvXp2 + P.nvYp2);
    X = (double)vx * P.voxelSize + (double)p.lx * P.l;
    Y = (double)vy * P.voxelSize + (double)p.ly * P.l;
    Z = (double)vz * P.voxelSize + (double)p.lz * P.l;
}
// positionToVoxelID (reference DEMHelperKernels.cuh:137-159): truncating encode
__device__ __forceinline__ void pos_encode(OwnerPos& p, const DevParams& P, double X, double Y, double Z) {
    const unsigned long long nx = (unsigned long long)(X / P.voxelSize);
    const unsigned long long ny = (unsigned long long)(Y / P.voxelSize);
    const unsigned long long nz = (unsigned long long)(Z / P.voxelSize);
    p.lx = (unsigned short)((X - (double)nx * P.voxelSize) / P.l);
    p.ly = (unsigned short)((Y - (double)ny * P.voxelSize) / P.l);
    p.lz = (unsigned short)((Z - (double)nz * P.voxelSize) / P.l);
    p.voxel = nx + (ny << P.nvXp2) + (nz << (P.nvXp2 + P.nvYp2));
}

// Same truncating encode without the six double-precision divisions: quotient by reciprocal multiply, then an exact
// fused-multiply-add remainder corrects it to floor(X / voxelSize) and floor(rem / l). Identical to pos_encode except
// when a quotient lies within one rounding error of an integer (where the reference's own rounded division decides).
__device__ __forceinline__ void pos_encode_fast(OwnerPos& p, const DevParams& P, double X, double Y, double Z) {
    const double c[3] = {X, Y, Z};
    unsigned long long n[3];
    unsigned int sub[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        long long q = (long long)(c[k] * P.inv_voxelSize);
        double r = fma(-(double)q, P.voxelSize, c[k]);
        if (r < 0.0) { q--; r += P.voxelSize; } else if (r >= P.voxelSize) { q++; r -= P.voxelSize; }
        int sq = (int)(r * P.inv_l);
        const double rr = fma(-(double)sq, P.l, r);
        if (rr < 0.0) sq--; else if (rr >= P.l) sq++;
        n[k] = (unsigned long long)q;
        sub[k] = (unsigned int)min(max(sq, 0), 65535);
    }
    p.lx = (unsigned short)sub[0]; p.ly = (unsigned short)sub[1]; p.lz = (unsigned short)sub[2];
    p.voxel = n[0] + (n[1] << P.nvXp2) + (n[2] << (P.nvXp2 + P.nvYp2));
}

__device__ __forceinline__ uint32_t mask_pair(uint32_t i, uint32_t j) {
    const uint32_t a = min(i, j), b = max(i, j);
    return (1u + b) * b / 2u + a;
}

// 128-bit vector reduction into global memory (sm_90+): SASS REDG.E.ADD.F32x4
__device__ __forceinline__ void red_add_v4(float4* addr, float x, float y, float z) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(0.0f)
                 : "memory");
}

// ---- system-scope flag words (peer memory over NVLink) ----
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
constexpr long long MG_SPIN_TIMEOUT_CYCLES = 20000000000ll;  // ~10 s: the peer is gone; report instead of hanging

// All-gather of two 32-bit words over the ranks through the peer-mapped mailboxes, called by ONE full warp: lane r
// (r < world) stores this rank's words into rank r's mailbox and then waits for rank r's words in its own.  On return
// lane r holds rank r's words in (o0, o1); lanes >= world hold this rank's own.  Two mailbox sets alternate with the
// all-gather number: a rank cannot start all-gather m+2 before every rank has finished reading all-gather m.
__device__ __forceinline__ void mg_allgather(const MgDev& M, uint32_t v0, uint32_t v1, uint32_t& o0, uint32_t& o1,
                                             uint32_t* flags) {
    const int lane = threadIdx.x & 31;
    const unsigned long long m = *M.mail_ctr + 1ull;
    __syncwarp();
    o0 = v0; o1 = v1;
    if (lane < M.world) {
        char* slot = M.peer_block[lane] + MG_HDR_MAIL + ((size_t)(m & 1ull) * MG_MAX_WORLD + (size_t)M.rank) * 16u;
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(slot + 8), "r"(v0), "r"(v1) : "memory");
        st_release_sys(reinterpret_cast<unsigned long long*>(slot), m);
        const char* mine = M.my_block + MG_HDR_MAIL + ((size_t)(m & 1ull) * MG_MAX_WORLD + (size_t)lane) * 16u;
        const long long t0 = clock64();
        bool ok = true;
        while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(mine)) < m) {
            if (clock64() - t0 > MG_SPIN_TIMEOUT_CYCLES) { atomicOr(&flags[DEM_FLAG_HALO], 64u); ok = false; break; }
        }
        if (ok) {
            const uint2 r = __ldcg(reinterpret_cast<const uint2*>(mine + 8));
            o0 = r.x; o1 = r.y;
        }
    }
    __syncwarp();
    if (lane == 0) *M.mail_ctr = m;
}

template <typename T>
__device__ __forceinline__ T ldg128(const T* p) {
    static_assert(sizeof(T) == 16, "128-bit load");
    int4 r = __ldg(reinterpret_cast<const int4*>(p));
    return *reinterpret_cast<T*>(&r);
}

}  // namespace demb
