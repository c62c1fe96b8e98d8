// kernels_step.cu -- per-step hot kernels: contact force (Hertz-Mindlin with history / frictionless Hertz) and
// owner integration.  Hand-written for sm_100a; HBM-bound by design:
//   * every owner gather is two 256-bit loads of one 64-byte record (LDG.E.ENL2.256),
//   * the contact stream is a 128-bit compiled record + a 128-bit history word, the history word being touched only
//     for contacts that are (or just stopped being) in physical touch,
//   * owner wrenches are scattered with 128-bit vector reductions (REDG.E.ADD.F32x4) into a 32-byte accumulator that
//     stays L2-resident,
//   * the owner record carries the WORLD-frame angular velocity and the wrench accumulates the WORLD-frame torque, so a
//     contact needs two quaternion rotations (the sphere offsets) instead of the reference's eight; the integrator
//     rotates once per owner (R(w x c) = (Rw) x (Rc), sum_i R^T t_i = R^T sum_i t_i).
// Reference behaviour being reproduced: src/kernel/DEMCalcForceKernels.cu:44-267 (calculateContactForces),
// DEMCustomizablePolicies/FullHertzianForceModel.cu, FrictionlessHertzianForceModel.cu,
// src/kernel/DEMCollectForceKernels_Compact.cu:13-102 (forceToAcc), src/kernel/DEMIntegrationKernels.cu:100-264.
#include <algorithm>

#include "dem_kernels.h"

namespace demb {

// ---------------------------------------------------------------------------------------------------------------
// gathered end point
struct End {
    float4 q;   // w,x,y,z
    float3 v;   // linear velocity
    float3 ww;  // WORLD-frame angular velocity
    float mass;
};

// the owner record is two 32-byte sectors, one 256-bit load each: geometry {pos, quat} and kinematics {vel+mass, omg}
__device__ __forceinline__ void load_owner_geom(const OwnerState* __restrict__ st, uint32_t o, OwnerPos& pos, End& e) {
    const float* base = reinterpret_cast<const float*>(st + o);
    uint32_t a0, a1, a2, a3;
    float q0, q1, q2, q3;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=f"(q0), "=f"(q1), "=f"(q2), "=f"(q3)
                 : "l"(base));
    pos.voxel = ((unsigned long long)a1 << 32) | a0;
    pos.lx = (unsigned short)(a2 & 0xffffu);
    pos.ly = (unsigned short)(a2 >> 16);
    pos.lz = (unsigned short)(a3 & 0xffffu);
    pos.family = (unsigned char)((a3 >> 16) & 0xffu);
    pos.flags = (unsigned char)(a3 >> 24);
    e.q = make_float4(q0, q1, q2, q3);
}
__device__ __forceinline__ void load_owner_kin(const OwnerState* __restrict__ st, uint32_t o, End& e) {
    const float* base = reinterpret_cast<const float*>(st + o);
    float v0, v1, v2, v3, w0, w1, w2, w3;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(w0), "=f"(w1), "=f"(w2), "=f"(w3)
                 : "l"(base + 8));
    e.v = f3(v0, v1, v2);
    e.mass = v3;
    e.ww = f3(w0, w1, w2);
}
__device__ __forceinline__ void load_owner(const OwnerState* __restrict__ st, uint32_t o, OwnerPos& pos, End& e) {
    load_owner_geom(st, o, pos, e);
    load_owner_kin(st, o, e);
}

// contact point in the world frame (LBF-relative), for the per-contact record: owner A's position + lever arm
__device__ __forceinline__ float4 contact_point_world(const DevParams& P, const OwnerPos& pA, float3 armA) {
    double X, Y, Z;
    pos_decode(pA, P, X, Y, Z);
    return make_float4((float)(X + (double)armA.x), (float)(Y + (double)armA.y), (float)(Z + (double)armA.z), 0.f);
}

// Division / square root of the force model.  FAST = one MUFU (reciprocal / reciprocal square root, 1 ulp) and a
// multiply instead of the IEEE-rounded sequences (8-12 instructions and a branch each): -100 of ~1100 SASS
// instructions, -3 % kernel time; the difference (<= 2 ulp of a force term) is far inside the fp32 parity tolerance
// (tests/test_gpu_parity.py runs both).  FAST = false reproduces the reference's correctly rounded operators.
template <bool FAST>
__device__ __forceinline__ float fdiv(float a, float b) { return FAST ? __fdividef(a, b) : a / b; }
template <bool FAST>
__device__ __forceinline__ float fsqrt(float a) { return FAST ? a * rsqrtf(fmaxf(a, 1e-37f)) : sqrtf(a); }
template <bool FAST>
__device__ __forceinline__ float flength(float3 a) { return fsqrt<FAST>(dot(a, a)); }

// Hertz-Mindlin with history (MODEL 0) or frictionless Hertz (MODEL 1).
//   depth > 0, n = unit normal B->A, armA/armB = contact point minus owner position (world frame).
// Returns force on A (world) and the torque-only rolling-resistance pseudo force.
template <int MODEL, bool FAST = false>
__device__ __forceinline__ void contact_model(const MatPair& mp, float h, float depth, float3 n, float3 armA,
                                              float3 armB, const End& A, const End& B, float rA, float rB,
                                              float4& hist, float3& force, float3& troll) {
    // velocity of the contact point on each body: v + w x r  ( == v + R(omgBar x c_local) of the reference)
    const float3 rotVelCPA = cross(A.ww, armA);
    const float3 rotVelCPB = cross(B.ww, armB);
    const float3 velB2A = (A.v + rotVelCPA) - (B.v + rotVelCPB);
    const float projection = dot(velB2A, n);
    const float mass_eff = fdiv<FAST>(A.mass * B.mass, A.mass + B.mass);
    const float sqrt_Rd = fsqrt<FAST>(depth * fdiv<FAST>(rA * rB, rA + rB));
    const float Sn = 2.f * mp.E_cnt * sqrt_Rd;
    const float k_n = 0.6666666666666666f * Sn;
    const float gamma_n = 1.825741858350554f * mp.beta * fsqrt<FAST>(Sn * mass_eff);
    force = (k_n * depth + gamma_n * projection) * n;
    troll = f3(0.f, 0.f, 0.f);
    if (MODEL == 0) {
        const float3 vrel_tan = velB2A - projection * n;
        float3 delta_tan = f3(hist.x, hist.y, hist.z);
        delta_tan = delta_tan + h * vrel_tan;
        delta_tan = delta_tan - dot(delta_tan, n) * n;
        float delta_time = hist.w + h;
        if (mp.Crr > 0.f) {
            bool add = true;
            const float R_eff = sqrtf((rA * rB) / (rA + rB));
            const float kn_simple = 1.3333333333333333f * mp.E_cnt * sqrtf(R_eff);
            const float gn_simple = -2.f * sqrtf(1.6666666666666667f * mass_eff * mp.E_cnt) * mp.beta * powf(R_eff, 0.25f);
            const float d_coeff = gn_simple / (2.f * sqrtf(kn_simple * mass_eff));
            if (d_coeff < 1.0f) {
                const float t_collision = 3.1415926535897932f * sqrtf(mass_eff / (kn_simple * (1.f - d_coeff * d_coeff)));
                if (delta_time <= t_collision) add = false;
            }
            if (add) {
                const float3 v_rot = rotVelCPB - rotVelCPA;
                const float v_rot_mag = length(v_rot);
                if (v_rot_mag > 1e-12f) troll = (v_rot * (1.f / v_rot_mag)) * (mp.Crr * length(force));
            }
        }
        if (mp.mu > 0.f) {
            const float kt = 8.f * mp.G_cnt * sqrt_Rd;
            const float gt = -1.825741858350554f * mp.beta * fsqrt<FAST>(mass_eff * kt);
            float3 tf = (-kt) * delta_tan - gt * vrel_tan;
            const float ft = flength<FAST>(tf);
            if (ft > 1e-12f) {
                const float ft_max = flength<FAST>(force) * mp.mu;
                if (ft > ft_max) {
                    tf = fdiv<FAST>(ft_max, ft) * tf;
                    delta_tan = (tf + gt * vrel_tan) * fdiv<FAST>(1.f, -kt);
                }
            } else {
                tf = f3(0.f, 0.f, 0.f);
            }
            force = force + tf;
        }
        hist = make_float4(delta_tan.x, delta_tan.y, delta_tan.z, delta_time);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sphere--sphere contacts, PLAIN kernel.  One thread per contact over the virtual concatenation of the two
// sphere--sphere lists (contacts in touch at the last rebuild first, mere candidates after them: warps are then
// homogeneous and the candidates' warps skip the force model).  Counts are DEVICE-resident (no host sync).  No shared
// memory: the whole L1 serves the owner gathers (every KB of shared memory taken from it costs this kernel time: 118 us
// for the touching list with none, 121 with 8 KB, 125 with 36 KB per CTA).  This is the kernel of choice while the
// lists are short-lived (cd_update_freq < 32): every candidate is looked at every step, but its 16-byte record is
// streamed, whereas the kernel below, which skips the candidates that cannot touch yet, has to fetch the records of
// those that are due as scattered sectors -- on the settled 1M-clump bed at 20 steps per list both take the same 37-40 us
// for the candidates (tools/tune2.py, gpurun_out/r2b_tune*.log); at 60 steps per list it is 93 us against 64.
template <int MODEL, bool RECORD, int MINB, bool FAST>
__global__ void __launch_bounds__(256, MINB) k_force_ss(const __grid_constant__ DevParams P) {
    if (P.flags[DEM_FLAG_POISON]) return;
    const uint32_t nT = *P.ss.count;
    // (force_opts bits 2 / 3: measurement only -- leave the candidate list / the touching list out)
    const uint32_t n = (P.force_opts & 4u) ? nT : nT + *P.sn.count;
    const uint32_t step = gridDim.x * blockDim.x;
    const uint32_t nround = (n + 31u) & ~31u;  // whole warps take part in the A-side reduction
    const uint32_t cfirst = (P.force_opts & 8u) ? (nT & ~31u) : 0u;
    const int lane = threadIdx.x & 31;
    const bool lazy = (P.force_opts & 2u) != 0u;
    for (uint32_t c = cfirst + blockIdx.x * blockDim.x + threadIdx.x; c < nround; c += step) {
        float wA[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t keyA = 0xffffffffu - (uint32_t)lane;  // never equal to a neighbour's key
        bool touch = false;
        bool work = c < n;
        const bool inT = c < nT;
        const uint32_t idx = inT ? c : c - nT;
        uint4* const cinfo = inT ? P.ss.cinfo : P.sn.cinfo;
        float4* const histp = inT ? P.ss.hist : P.sn.hist;
        uint4 ci = make_uint4(0u, 0u, 0u, 0u);
        if (work) {
            ci = __ldcs(&cinfo[idx]);  // streaming: evict-first
            keyA = ci.x;
        }
        if (work) {
        const uint32_t oA = ci.x, oB = ci.y;
        const bool alive = (ci.w >> 31) != 0u;
        float4 hist = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODEL == 0 && alive) hist = __ldcs(&histp[idx]);
        const float4 compA = __ldg(&P.comp[ci.z & 0xffffu]);
        const float4 compB = __ldg(&P.comp[ci.z >> 16]);
        OwnerPos pA, pB;
        End A, B;
        load_owner_geom(P.state, oA, pA, A);
        load_owner_geom(P.state, oB, pB, B);
        // velocities are only needed for pairs in touch: the list of pairs that overlapped at the rebuild fetches them
        // up front (independent loads in flight together), the candidate list only once the narrow phase says so
        if (inT || !lazy) {
            load_owner_kin(P.state, oA, A);
            load_owner_kin(P.state, oB, B);
        }

        // ---- narrow phase (checkSpheresOverlap<double,float>, DEMHelperKernels.cuh:292-326) ----
        const float3 relA = rotate(f3(compA.x, compA.y, compA.z), A.q);
        const float3 relB = rotate(f3(compB.x, compB.y, compB.z), B.q);
        long long ax, ay, az, bx, by, bz;
        pos_ints(pA, P.nvXp2, P.nvYp2, ax, ay, az);
        pos_ints(pB, P.nvXp2, P.nvYp2, bx, by, bz);
        // centre(A) - centre(B): exact integer owner difference (units of l) + rotated offsets
        const double dx = (double)(ax - bx) * P.l + ((double)relA.x - (double)relB.x);
        const double dy = (double)(ay - by) * P.l + ((double)relA.y - (double)relB.y);
        const double dz = (double)(az - bz) * P.l + ((double)relA.z - (double)relB.z);
        const double d2 = dx * dx + dy * dy + dz * dz;
        const float rA = compA.w, rB = compB.w;
        const double R = (double)rA + (double)rB;
        float3 nrm = f3((float)dx, (float)dy, (float)dz);
        const float mag2 = dot(nrm, nrm);
        const float imag = FAST ? rsqrtf(fmaxf(mag2, 1e-37f)) : 0.f;
        const float mag = FAST ? mag2 * imag : sqrtf(mag2);
        // overlap = R - |d| = (R^2 - d^2) / (R + |d|): numerator in double, the rest in float
        const float depth = fdiv<FAST>((float)(R * R - d2), (float)R + mag);

        if (depth > 0.f) {
            if (!inT && lazy) {
                load_owner_kin(P.state, oA, A);
                load_owner_kin(P.state, oB, B);
            }
            nrm = nrm * (FAST ? imag : 1.f / mag);
            // contact point = centre(B) + (rB - depth/2) n ; lever arms from each owner (world frame)
            const float s = rB - 0.5f * depth;
            const float3 armB = relB + s * nrm;
            const float3 armA = f3(relA.x - (float)dx, relA.y - (float)dy, relA.z - (float)dz) + s * nrm;
            const MatPair mp = P.matpair[ci.w & 0xffffu];
            float3 force, troll;
            contact_model<MODEL, FAST>(mp, P.h, depth, nrm, armA, armB, A, B, rA, rB, hist, force, troll);
            // wrench scatter (forceToAcc semantics; force and WORLD-frame torque sums, divided by mass / rotated and
            // divided by MOI once per owner in the integrator)
            const float3 Ft = force + troll;
            const float3 TA = cross(armA, Ft);
            const float3 TB = cross(Ft, armB);  // armB x (-Ft)
            wA[0] = force.x; wA[1] = force.y; wA[2] = force.z;
            wA[3] = TA.x; wA[4] = TA.y; wA[5] = TA.z;
            touch = true;
            red_add_v4(&P.wrench[oB].f, -force.x, -force.y, -force.z);
            red_add_v4(&P.wrench[oB].t, TB.x, TB.y, TB.z);
            if (MODEL == 0) {
                __stcs(&histp[idx], hist);
                if (!alive) cinfo[idx].w = ci.w | 0x80000000u;
            }
            if (RECORD) {
                float4* const frc = inT ? P.ss.force : P.sn.force;
                float4* const cpt = inT ? P.ss.cpoint : P.sn.cpoint;
                frc[idx] = make_float4(force.x, force.y, force.z, 0.f);
                cpt[idx] = contact_point_world(P, pA, armA);
            }
        } else {
            // not in touch: destroy history (FullHertzianForceModel.cu:129-136, DEMCalcForceKernels.cu:258-261)
            if (MODEL == 0 && alive) {
                __stcs(&histp[idx], make_float4(0.f, 0.f, 0.f, 0.f));
                cinfo[idx].w = ci.w & 0x7fffffffu;
            }
            if (RECORD) {
                float4* const frc = inT ? P.ss.force : P.sn.force;
                frc[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        }  // work
        // A side: the list is owner-major, so the contacts of one owner sit in adjacent lanes. Segmented suffix sum over
        // runs of equal owner, then ONE pair of vector reductions per run instead of one per contact.
        if (__any_sync(0xffffffffu, touch)) {
            // runs = maximal stretches of adjacent lanes with the same owner (robust to any key sequence)
            const uint32_t kprev = __shfl_up_sync(0xffffffffu, keyA, 1);
            const bool head = (lane == 0) || (kprev != keyA);
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            const uint32_t above = (lane == 31) ? 0u : (heads >> (lane + 1));
            const int run_end = above ? lane + __ffs(above) : 32;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const bool take = lane + off < run_end;
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const float up = __shfl_down_sync(0xffffffffu, wA[k], off);
                    if (take) wA[k] += up;
                }
            }
            const bool any_force = (wA[0] != 0.f) | (wA[1] != 0.f) | (wA[2] != 0.f) | (wA[3] != 0.f) | (wA[4] != 0.f) | (wA[5] != 0.f);
            if (head && any_force) {
                red_add_v4(&P.wrench[keyA].f, wA[0], wA[1], wA[2]);
                red_add_v4(&P.wrench[keyA].t, wA[3], wA[4], wA[5]);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// sphere--sphere contacts, kernel for LONG-LIVED lists.  One thread per contact: first the list of the pairs that were in touch at the last rebuild,
// then the mere candidates (warps are homogeneous: candidate warps hardly ever run the force model).  Counts are
// DEVICE-resident (no host sync).
//
// Candidates are the larger half of the work items of a settled bed and almost all of them are apart, so each carries
// the next step of the list's cycle at which it can possibly touch (one byte, ContactList::due): the margin of an owner
// IS the bound on how far it travels during the maxDrift steps a list is used (DEMMiscKernels.cu:37-61), so a gap g
// cannot close in fewer than g / c steps, c = (marginA + marginB) / maxDrift (the margins ride along in the owner
// records as a byte each, relative to the largest one: no extra gather).  The sweep sets the first value
// (kernels_sweep.cu), and every evaluation that finds the pair still apart REFRESHES it from the exact gap it just
// computed: a pair that stays g apart is looked at every g / c steps, so a cycle of D steps costs about ln D
// evaluations per candidate instead of D -- and lengthening the cycle costs the force kernel next to nothing.  A pair
// that is not evaluated contributes nothing (it cannot overlap), hence identical results.
// A warp that skipped some of its lanes would still wait for the owner gathers of the others, so the candidate phase
// COMPACTS per warp: each warp streams its contiguous share of the due bytes (one 32-bit load = 4 candidates per lane),
// queues the indices of those that are due (order preserved: the list stays owner-major for the A-side reduction) in a
// small ring in shared memory and pops 32 at a time.  No CTA barrier: the warps stay independent, which is what hides
// the latency of the dependent gathers.

// The step of the cycle at which a candidate that is `gap` apart now has to be looked at again: the gap closes by at
// most (marginA + marginB) / maxDrift per step; inv_closing = maxDrift / (2 max margin), the owners' flag bytes hold
// 255 - c with margin <= max margin (c + 1) / 256 (0.999, -1e-9 m: rounding of the float gap).
__device__ __forceinline__ uint32_t cand_due_step(float gap, float inv_closing, uint32_t flagsA, uint32_t flagsB, uint32_t cyc) {
    const float parts = (float)(512u - flagsA - flagsB);  // (cA + 1) + (cB + 1) of 512
    const float steps = fminf(fmaxf(__fdividef((gap - 1e-9f) * inv_closing * (0.999f * 512.f), parts), 0.f), 250.f);
    return min(cyc + (uint32_t)steps, 255u);
}

// One contact (lane) of a warp: narrow phase, force model, B-side reductions; then the warp's A-side reduction.  Called
// by whole warps (lanes without a contact pass work = false).  IN_T: the list of pairs in touch at the rebuild.
template <int MODEL, bool RECORD, bool FAST, bool IN_T>
__device__ __forceinline__ void ss_contact(const DevParams& P, bool work, uint32_t idx, int lane, bool lazy, uint32_t cyc,
                                           float inv_closing) {
    float wA[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t keyA = 0xffffffffu - (uint32_t)lane;  // never equal to a neighbour's key
    bool touch = false;
    uint4* const cinfo = IN_T ? P.ss.cinfo : P.sn.cinfo;
    float4* const histp = IN_T ? P.ss.hist : P.sn.hist;
    if (work) {
        const uint4 ci = __ldcs(&cinfo[idx]);  // streaming: evict-first
        keyA = ci.x;
        const uint32_t oA = ci.x, oB = ci.y;
        const bool alive = (ci.w >> 31) != 0u;
        float4 hist = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODEL == 0 && alive) hist = __ldcs(&histp[idx]);
        const float4 compA = __ldg(&P.comp[ci.z & 0xffffu]);
        const float4 compB = __ldg(&P.comp[ci.z >> 16]);
        OwnerPos pA, pB;
        End A, B;
        load_owner_geom(P.state, oA, pA, A);
        load_owner_geom(P.state, oB, pB, B);
        // velocities are only needed for pairs in touch: the list of pairs that overlapped at the rebuild fetches them
        // up front (independent loads in flight together), the candidate list only once the narrow phase says so
        if (IN_T || !lazy) {
            load_owner_kin(P.state, oA, A);
            load_owner_kin(P.state, oB, B);
        }

        // ---- narrow phase (checkSpheresOverlap<double,float>, DEMHelperKernels.cuh:292-326) ----
        const float3 relA = rotate(f3(compA.x, compA.y, compA.z), A.q);
        const float3 relB = rotate(f3(compB.x, compB.y, compB.z), B.q);
        long long ax, ay, az, bx, by, bz;
        pos_ints(pA, P.nvXp2, P.nvYp2, ax, ay, az);
        pos_ints(pB, P.nvXp2, P.nvYp2, bx, by, bz);
        // centre(A) - centre(B): exact integer owner difference (units of l) + rotated offsets
        const double dx = (double)(ax - bx) * P.l + ((double)relA.x - (double)relB.x);
        const double dy = (double)(ay - by) * P.l + ((double)relA.y - (double)relB.y);
        const double dz = (double)(az - bz) * P.l + ((double)relA.z - (double)relB.z);
        const double d2 = dx * dx + dy * dy + dz * dz;
        const float rA = compA.w, rB = compB.w;
        const double R = (double)rA + (double)rB;
        float3 nrm = f3((float)dx, (float)dy, (float)dz);
        const float mag2 = dot(nrm, nrm);
        const float imag = FAST ? rsqrtf(fmaxf(mag2, 1e-37f)) : 0.f;
        const float mag = FAST ? mag2 * imag : sqrtf(mag2);
        // overlap = R - |d| = (R^2 - d^2) / (R + |d|): numerator in double, the rest in float
        const float depth = fdiv<FAST>((float)(R * R - d2), (float)R + mag);

        if (depth > 0.f) {
            if (!IN_T && lazy) {
                load_owner_kin(P.state, oA, A);
                load_owner_kin(P.state, oB, B);
            }
            nrm = nrm * (FAST ? imag : 1.f / mag);
            // contact point = centre(B) + (rB - depth/2) n ; lever arms from each owner (world frame)
            const float s = rB - 0.5f * depth;
            const float3 armB = relB + s * nrm;
            const float3 armA = f3(relA.x - (float)dx, relA.y - (float)dy, relA.z - (float)dz) + s * nrm;
            const MatPair mp = P.matpair[ci.w & 0xffffu];
            float3 force, troll;
            contact_model<MODEL, FAST>(mp, P.h, depth, nrm, armA, armB, A, B, rA, rB, hist, force, troll);
            // wrench scatter (forceToAcc semantics; force and WORLD-frame torque sums, divided by mass / rotated and
            // divided by MOI once per owner in the integrator)
            const float3 Ft = force + troll;
            const float3 TA = cross(armA, Ft);
            const float3 TB = cross(Ft, armB);  // armB x (-Ft)
            wA[0] = force.x; wA[1] = force.y; wA[2] = force.z;
            wA[3] = TA.x; wA[4] = TA.y; wA[5] = TA.z;
            touch = true;
            red_add_v4(&P.wrench[oB].f, -force.x, -force.y, -force.z);
            red_add_v4(&P.wrench[oB].t, TB.x, TB.y, TB.z);
            if (MODEL == 0) {
                __stcs(&histp[idx], hist);
                if (!alive) cinfo[idx].w = ci.w | 0x80000000u;
            }
            if (RECORD) {
                float4* const frc = IN_T ? P.ss.force : P.sn.force;
                float4* const cpt = IN_T ? P.ss.cpoint : P.sn.cpoint;
                frc[idx] = make_float4(force.x, force.y, force.z, 0.f);
                cpt[idx] = contact_point_world(P, pA, armA);
            }
        } else {
            // not in touch: destroy history (FullHertzianForceModel.cu:129-136, DEMCalcForceKernels.cu:258-261)
            if (MODEL == 0 && alive) {
                __stcs(&histp[idx], make_float4(0.f, 0.f, 0.f, 0.f));
                cinfo[idx].w = ci.w & 0x7fffffffu;
            }
            // a candidate also learns when it has to be looked at again
            if (!IN_T && inv_closing > 0.f) {
                const uint32_t due = cand_due_step(-depth, inv_closing, pA.flags, pB.flags, cyc);
                if (due > cyc) P.sn.due[idx] = (uint8_t)due;
            }
            if (RECORD) {
                float4* const frc = IN_T ? P.ss.force : P.sn.force;
                frc[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    // A side: the list is owner-major, so the contacts of one owner sit in adjacent lanes. Segmented suffix sum over
    // runs of equal owner, then ONE pair of vector reductions per run instead of one per contact.
    if (__any_sync(0xffffffffu, touch)) {
        // runs = maximal stretches of adjacent lanes with the same owner (robust to any key sequence)
        const uint32_t kprev = __shfl_up_sync(0xffffffffu, keyA, 1);
        const bool head = (lane == 0) || (kprev != keyA);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t above = (lane == 31) ? 0u : (heads >> (lane + 1));
        const int run_end = above ? lane + __ffs(above) : 32;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const bool take = lane + off < run_end;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const float up = __shfl_down_sync(0xffffffffu, wA[k], off);
                if (take) wA[k] += up;
            }
        }
        const bool any_force = (wA[0] != 0.f) | (wA[1] != 0.f) | (wA[2] != 0.f) | (wA[3] != 0.f) | (wA[4] != 0.f) | (wA[5] != 0.f);
        if (head && any_force) {
            red_add_v4(&P.wrench[keyA].f, wA[0], wA[1], wA[2]);
            red_add_v4(&P.wrench[keyA].t, wA[3], wA[4], wA[5]);
        }
    }
}

// Screen of ONE candidate that is due: narrow phase only (same expressions as ss_contact, so both agree on who is in
// touch).  Split in a load half and a compute half so that a lane can have the gathers of two candidates in flight
// together: the candidate phase runs at the speed of a warp's chain of dependent loads (due bytes -> compiled record ->
// owner records), not of any pipe.
struct CandScreen {
    uint4 ci;
    OwnerPos pA, pB;
    float4 qA, qB;
};
__device__ __forceinline__ void cand_owners(const DevParams& P, CandScreen& c) {
    End e;
    load_owner_geom(P.state, c.ci.x, c.pA, e);
    c.qA = e.q;
    load_owner_geom(P.state, c.ci.y, c.pB, e);
    c.qB = e.q;
}
// true: the pair needs the full treatment (it is in touch, or it carries live history that has to be destroyed)
template <bool FAST>
__device__ __forceinline__ bool cand_screen(const DevParams& P, const CandScreen& c, uint32_t idx, uint32_t cyc,
                                            float inv_closing) {
    const float4 compA = __ldg(&P.comp[c.ci.z & 0xffffu]);
    const float4 compB = __ldg(&P.comp[c.ci.z >> 16]);
    const float3 relA = rotate(f3(compA.x, compA.y, compA.z), c.qA);
    const float3 relB = rotate(f3(compB.x, compB.y, compB.z), c.qB);
    long long ax, ay, az, bx, by, bz;
    pos_ints(c.pA, P.nvXp2, P.nvYp2, ax, ay, az);
    pos_ints(c.pB, P.nvXp2, P.nvYp2, bx, by, bz);
    const double dx = (double)(ax - bx) * P.l + ((double)relA.x - (double)relB.x);
    const double dy = (double)(ay - by) * P.l + ((double)relA.y - (double)relB.y);
    const double dz = (double)(az - bz) * P.l + ((double)relA.z - (double)relB.z);
    const double d2 = dx * dx + dy * dy + dz * dz;
    const double R = (double)compA.w + (double)compB.w;
    const float3 nrm = f3((float)dx, (float)dy, (float)dz);
    const float mag2 = dot(nrm, nrm);
    const float mag = FAST ? mag2 * rsqrtf(fmaxf(mag2, 1e-37f)) : sqrtf(mag2);
    const float depth = fdiv<FAST>((float)(R * R - d2), (float)R + mag);
    if (depth > 0.f || (c.ci.w >> 31) != 0u) return true;
    if (inv_closing > 0.f) {
        const uint32_t due = cand_due_step(-depth, inv_closing, c.pA.flags, c.pB.flags, cyc);
        if (due > cyc) P.sn.due[idx] = (uint8_t)due;
    }
    return false;
}

constexpr int FQ_LOADS = 6;      // 32-bit words of due bytes filtered per lane and round (4 candidates each)
constexpr int FQ_SLOTS = 1024;   // per-warp ring of due candidate indices: < 64 left over + 128 * FQ_LOADS new ones
constexpr int HQ_SLOTS = 128;    // per-warp ring of the screened candidates that need the force model
static_assert(63 + 128 * FQ_LOADS <= FQ_SLOTS, "ring too small");

template <int MODEL, bool RECORD, int MINB, bool FAST>
__global__ void __launch_bounds__(256, MINB) k_force_ss_due(const __grid_constant__ DevParams P) {
    if (P.flags[DEM_FLAG_POISON]) return;
    __shared__ uint32_t fq[8][FQ_SLOTS];
    __shared__ uint32_t hq[8][HQ_SLOTS];
    const int lane = threadIdx.x & 31;
    // ---- the pairs that were in touch at the rebuild (force_opts bits 2 / 3: measurement only -- leave the candidate
    //      list / this list out) ----
    {
        const uint32_t nT = (P.force_opts & 8u) ? 0u : *P.ss.count;
        const uint32_t step = gridDim.x * blockDim.x;
        const uint32_t nround = (nT + 31u) & ~31u;  // whole warps take part in the A-side reduction
        for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nround; c += step)
            ss_contact<MODEL, RECORD, FAST, true>(P, c < nT, c, lane, false, 0u, 0.f);
    }
    // ---- the candidates ----
    const uint32_t nN = (P.force_opts & 4u) ? 0u : *P.sn.count;
    // a fixed expand factor (beta >= 0) is no bound on anybody's motion: then every candidate is looked at every step
    const bool skip = (P.force_opts & 1u) != 0u && P.beta < 0.f && P.sn.due != nullptr;
    const bool lazy = (P.force_opts & 2u) != 0u;
    // integrations since the rebuild that made these lists
    const uint32_t cyc = skip ? P.flags[DEM_FLAG_CYCLE_STEP] : 0xffffffffu;
    const float inv_closing = (skip && (P.force_opts & 32u)) ? __uint_as_float(P.flags[DEM_FLAG_INV_CLOSING]) : 0.f;
    // compacting: every warp owns a contiguous share of the list (a multiple of 128 candidates: aligned 32-bit loads)
    uint32_t* const ring = fq[threadIdx.x >> 5];
    uint32_t* const heavy = hq[threadIdx.x >> 5];
    const uint32_t nwarps = gridDim.x * 8u;
    const uint32_t share = ((nN + nwarps - 1u) / nwarps + 127u) & ~127u;
    uint32_t next = min(nN, (blockIdx.x * 8u + (threadIdx.x >> 5)) * share);  // next unfiltered candidate of this warp
    const uint32_t last = min(nN, next + share);
    const uint32_t* const due4 = reinterpret_cast<const uint32_t*>(P.sn.due);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t qhead = 0u, qcnt = 0u, hhead = 0u, hcnt = 0u;  // warp-uniform ring state
    for (;;) {
        // refill the ring until enough due candidates are queued or the share is used up
        while (qcnt < 64u && next < last) {
            uint32_t w[FQ_LOADS];
#pragma unroll
            for (int u = 0; u < FQ_LOADS; u++) {
                const uint32_t k = next + (uint32_t)(u * 128 + lane * 4);
                w[u] = (k < last) ? due4[k >> 2] : 0xffffffffu;  // (plain load: this kernel also stores due bytes)
            }
#pragma unroll
            for (int u = 0; u < FQ_LOADS; u++) {
                const uint32_t k = next + (uint32_t)(u * 128 + lane * 4);
                uint32_t m[4];
                bool d[4];
                uint32_t below = 0u;  // due candidates of the lanes below me
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    d[b] = (k + b < last) && ((w[u] >> (8 * b)) & 0xffu) <= cyc;
                    if (RECORD && (k + b < last) && !d[b]) P.sn.force[k + b] = make_float4(0.f, 0.f, 0.f, 0.f);
                    m[b] = __ballot_sync(0xffffffffu, d[b]);
                    below += (uint32_t)__popc(m[b] & lt);
                }
                uint32_t slot = qhead + qcnt + below;
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (d[b]) ring[(slot++) & (FQ_SLOTS - 1)] = k + b;
                qcnt += (uint32_t)(__popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]));
            }
            next += 128u * FQ_LOADS;
            __syncwarp();
        }
        if (qcnt == 0u && hcnt == 0u) break;
        // ---- screen up to 64 due candidates: both compiled records, then all four owner records, in flight together ----
        if (qcnt) {
            const uint32_t take = min(qcnt, 64u);
            const bool w0 = (uint32_t)lane < take, w1 = (uint32_t)lane + 32u < take;
            uint32_t i0 = 0u, i1 = 0u;
            if (w0) i0 = ring[(qhead + (uint32_t)lane) & (FQ_SLOTS - 1)];
            if (w1) i1 = ring[(qhead + 32u + (uint32_t)lane) & (FQ_SLOTS - 1)];
            __syncwarp();
            qhead += take;
            qcnt -= take;
            CandScreen c0, c1;
            c0.ci = c1.ci = make_uint4(0u, 0u, 0u, 0u);
            if (w0) c0.ci = __ldcs(&P.sn.cinfo[i0]);
            if (w1) c1.ci = __ldcs(&P.sn.cinfo[i1]);
            if (w0) cand_owners(P, c0);
            if (w1) cand_owners(P, c1);
            const bool h0 = w0 && cand_screen<FAST>(P, c0, i0, cyc, inv_closing);
            const bool h1 = w1 && cand_screen<FAST>(P, c1, i1, cyc, inv_closing);
            if (RECORD) {
                if (w0 && !h0) P.sn.force[i0] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (w1 && !h1) P.sn.force[i1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const uint32_t m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
            if (h0) heavy[(hhead + hcnt + (uint32_t)__popc(m0 & lt)) & (HQ_SLOTS - 1)] = i0;
            if (h1) heavy[(hhead + hcnt + (uint32_t)__popc(m0) + (uint32_t)__popc(m1 & lt)) & (HQ_SLOTS - 1)] = i1;
            hcnt += (uint32_t)(__popc(m0) + __popc(m1));
            __syncwarp();
        }
        // ---- the force model for those in touch: full warps, or whatever is left once the share is used up ----
        while (hcnt >= 32u || (hcnt && qcnt == 0u && next >= last)) {
            const uint32_t take = min(hcnt, 32u);
            const bool work = (uint32_t)lane < take;
            uint32_t idx = 0u;
            if (work) idx = heavy[(hhead + (uint32_t)lane) & (HQ_SLOTS - 1)];
            __syncwarp();
            hhead += take;
            hcnt -= take;
            ss_contact<MODEL, RECORD, FAST, false>(P, work, idx, lane, lazy, cyc, inv_closing);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Wrench of the few "wall" owners (analytical boundaries, meshes).  The contacts of a warp almost always share ONE such
// owner, so a per-contact or even per-warp reduction would serialise on the single L2 line of its wrench: reduce across
// the warp first, then into a per-CTA shared-memory table, and flush one vector reduction pair per owner per CTA.
struct CtaWallAcc {
    uint32_t key[8];
    float val[8][6];
    int count;
};
__device__ __forceinline__ void wall_init(CtaWallAcc& w) {
    if (threadIdx.x < 8) {
        w.key[threadIdx.x] = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 6; k++) w.val[threadIdx.x][k] = 0.f;
    }
    if (threadIdx.x == 0) w.count = 0;
    __syncthreads();
}
// called by whole warps (inactive lanes pass touchB = false)
__device__ __forceinline__ void wall_add(CtaWallAcc& w, const DevParams& P, bool touchB, uint32_t oBkey, const float wB[6]) {
    uint32_t todo = __ballot_sync(0xffffffffu, touchB);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const uint32_t key = __shfl_sync(0xffffffffu, oBkey, leader);
        const bool mine = touchB && (oBkey == key);
        float v[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
            v[k] = mine ? wB[k] : 0.f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        }
        if ((int)(threadIdx.x & 31) == leader) {
            int slot = -1;
            const int cnt = min(w.count, 8);
            for (int t = 0; t < cnt; t++)
                if (w.key[t] == key) { slot = t; break; }
            if (slot < 0) {
                slot = atomicAdd(&w.count, 1);
                if (slot < 8) w.key[slot] = key;  // (a racing warp may append the same owner twice: harmless)
            }
            if (slot < 8) {
#pragma unroll
                for (int k = 0; k < 6; k++) atomicAdd(&w.val[slot][k], v[k]);
            } else {
                red_add_v4(&P.wrench[key].f, v[0], v[1], v[2]);
                red_add_v4(&P.wrench[key].t, v[3], v[4], v[5]);
            }
        }
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}
__device__ __forceinline__ void wall_flush(CtaWallAcc& w, const DevParams& P) {
    __syncthreads();
    if (threadIdx.x < 8 && threadIdx.x < (unsigned)min(w.count, 8) && w.key[threadIdx.x] != 0xffffffffu) {
        const float* v = w.val[threadIdx.x];
        red_add_v4(&P.wrench[w.key[threadIdx.x]].f, v[0], v[1], v[2]);
        red_add_v4(&P.wrench[w.key[threadIdx.x]].t, v[3], v[4], v[5]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sphere--analytical contacts (planes, infinite cylinders): checkSphereEntityOverlap, DEMHelperKernels.cuh:459-521
template <int MODEL, bool RECORD>
__global__ void __launch_bounds__(256) k_force_sa(const __grid_constant__ DevParams P) {
    if (P.flags[DEM_FLAG_POISON]) return;
    __shared__ CtaWallAcc wall;
    wall_init(wall);
    const uint32_t n = *P.sa.count;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t nround = (n + 31u) & ~31u;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nround; c += stride) {
        // every lane of the warp takes part in the B-side aggregation below, active or not
        const bool active = c < n;
        float wB[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool touchB = false;
        uint32_t oBkey = 0xffffffffu;
        if (active) {
            const uint4 ci = P.sa.cinfo[c];
            const uint32_t oA = ci.x;
            const AnalObj ob = P.anal[ci.y];
            const uint32_t oB = ob.owner;
            oBkey = oB;
            const bool alive = (ci.w >> 31) != 0u;
            float4 hist = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODEL == 0 && alive) hist = P.sa.hist[c];
            const float4 compA = __ldg(&P.comp[ci.z & 0xffffu]);
            OwnerPos pA, pB;
            End A, B;
            load_owner(P.state, oA, pA, A);
            load_owner(P.state, oB, pB, B);
            B.mass = ob.mass;  // objMass (DEMCalcForceKernels.cu:198)
            const float3 relA = rotate(f3(compA.x, compA.y, compA.z), A.q);
            const float3 relB = rotate(f3(ob.relx, ob.rely, ob.relz), B.q);
            const float3 dirB = rotate(f3(ob.rotx, ob.roty, ob.rotz), B.q);
            long long ax, ay, az, bx, by, bz;
            pos_ints(pA, P.nvXp2, P.nvYp2, ax, ay, az);
            pos_ints(pB, P.nvXp2, P.nvYp2, bx, by, bz);
            // sphere centre minus entity point
            const double dx = (double)(ax - bx) * P.l + ((double)relA.x - (double)relB.x);
            const double dy = (double)(ay - by) * P.l + ((double)relA.y - (double)relB.y);
            const double dz = (double)(az - bz) * P.l + ((double)relA.z - (double)relB.z);
            const float rA = compA.w;
            float depth;
            float3 nrm;
            float3 armA;  // CP - ownerA
            if (ob.type == DEM_ANAL_PLANE) {
                const float dist = (float)(dx * (double)dirB.x + dy * (double)dirB.y + dz * (double)dirB.z);
                depth = (float)((double)rA - (double)dist);
                const float s = (float)((double)dist + ((double)rA - (double)dist) / 2.0);
                nrm = dirB;
                armA = relA - s * dirB;
            } else {  // DEM_ANAL_CYL_INF
                // sph2cyl = B - A, minus its axial projection
                const float proj = (float)(-(dx * (double)dirB.x + dy * (double)dirB.y + dz * (double)dirB.z));
                const double sx = -dx - (double)(proj * dirB.x);
                const double sy = -dy - (double)(proj * dirB.y);
                const double sz = -dz - (double)(proj * dirB.z);
                const double dr = sqrt(sx * sx + sy * sy + sz * sz);
                const float cyl_rad = ob.size1;
                const double dep = (double)rA - (double)ob.normal_sign * ((double)cyl_rad - dr);
                depth = (float)dep;
                if (dr >= 1e-12) {
                    const double k = (double)ob.normal_sign / dr;
                    nrm = f3((float)(k * sx), (float)(k * sy), (float)(k * sz));
                    const float s = (float)((double)rA - dep / 2.0);
                    armA = relA - s * nrm;
                } else {
                    nrm = dirB;
                    armA = relA;
                }
            }
            if (depth > 0.f) {
                // CP - ownerB = (CP - ownerA) + (ownerA - ownerB)
                const float3 armB = f3((float)((double)armA.x + (double)(ax - bx) * P.l),
                                       (float)((double)armA.y + (double)(ay - by) * P.l),
                                       (float)((double)armA.z + (double)(az - bz) * P.l));
                const MatPair mp = P.matpair[ci.w & 0xffffu];
                float3 force, troll;
                contact_model<MODEL>(mp, P.h, depth, nrm, armA, armB, A, B, rA, 1e15f, hist, force, troll);
                const float3 Ft = force + troll;
                const float3 TA = cross(armA, Ft);
                const float3 TB = cross(Ft, armB);
                red_add_v4(&P.wrench[oA].f, force.x, force.y, force.z);
                red_add_v4(&P.wrench[oA].t, TA.x, TA.y, TA.z);
                wB[0] = -force.x; wB[1] = -force.y; wB[2] = -force.z;
                wB[3] = TB.x; wB[4] = TB.y; wB[5] = TB.z;
                touchB = true;
                if (MODEL == 0) {
                    P.sa.hist[c] = hist;
                    if (!alive) P.sa.cinfo[c].w = ci.w | 0x80000000u;
                }
                if (RECORD) {
                    P.sa.force[c] = make_float4(force.x, force.y, force.z, 0.f);
                    P.sa.cpoint[c] = contact_point_world(P, pA, armA);
                }
            } else {
                if (MODEL == 0 && alive) {
                    P.sa.hist[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    P.sa.cinfo[c].w = ci.w & 0x7fffffffu;
                }
                if (RECORD) P.sa.force[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        wall_add(wall, P, touchB, oBkey, wB);
    }
    wall_flush(wall, P);
}

// ---------------------------------------------------------------------------------------------------------------
// sphere--triangle contacts: triangle_sphere_CD<double3,double> (DEMCollisionKernels.cu:15-156) inside
// calculateContactForces (DEMCalcForceKernels.cu:134-177).  Everything is evaluated in the frame of owner A's position
// (the integer owner difference is exact), the nodes being rotated in double by the float quaternion as the
// reference's equipOwnerPosRot<double3> does.
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 dcross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ D3 rotate_d(float4 n, float4 q) {
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    const double X = (double)n.x, Y = (double)n.y, Z = (double)n.z;
    D3 r;
    r.x = (double)(2.0f * (w * w + x * x) - 1.0f) * X + (double)(2.0f * (x * y - w * z)) * Y + (double)(2.0f * (x * z + w * y)) * Z;
    r.y = (double)(2.0f * (x * y + w * z)) * X + (double)(2.0f * (w * w + y * y) - 1.0f) * Y + (double)(2.0f * (y * z - w * x)) * Z;
    r.z = (double)(2.0f * (x * z - w * y)) * X + (double)(2.0f * (y * z + w * x)) * Y + (double)(2.0f * (w * w + z * z) - 1.0f) * Z;
    return r;
}
// closest point of triangle ABC to P; returns true when it lies on an edge or vertex (snap_to_face)
__device__ __forceinline__ bool snap_to_face(D3 A, D3 B, D3 C, D3 Pt, D3& res) {
    const D3 AB = B - A, AC = C - A, AP = Pt - A;
    const double d1 = ddot(AB, AP), d2 = ddot(AC, AP);
    if (d1 <= 0. && d2 <= 0.) { res = A; return true; }
    const D3 BP = Pt - B;
    const double d3 = ddot(AB, BP), d4 = ddot(AC, BP);
    if (d3 >= 0. && d4 <= d3) { res = B; return true; }
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0. && d1 >= 0. && d3 <= 0.) { res = A + AB * (d1 / (d1 - d3)); return true; }
    const D3 CP = Pt - C;
    const double d5 = ddot(AB, CP), d6 = ddot(AC, CP);
    if (d6 >= 0. && d5 <= d6) { res = C; return true; }
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0. && d2 >= 0. && d6 <= 0.) { res = A + AC * (d2 / (d2 - d6)); return true; }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0. && (d4 - d3) >= 0. && (d5 - d6) >= 0.) { res = B + (C - B) * ((d4 - d3) / ((d4 - d3) + (d5 - d6))); return true; }
    const double denom = 1.0 / (va + vb + vc);
    res = (A + AB * (vb * denom)) + AC * (vc * denom);
    return false;
}

template <int MODEL, bool RECORD>
__global__ void __launch_bounds__(256) k_force_st(const __grid_constant__ DevParams P) {
    if (P.flags[DEM_FLAG_POISON]) return;
    __shared__ CtaWallAcc wall;
    wall_init(wall);
    const uint32_t n = *P.st.count;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t nround = (n + 31u) & ~31u;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nround; c += stride) {
        const bool active = c < n;
        float wB[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool touchB = false;
        uint32_t oBkey = 0xffffffffu;
        if (active) {
            const uint4 ci = P.st.cinfo[c];
            const uint32_t oA = ci.x, tri = ci.y;
            const uint32_t oB = P.tri_info[tri].x;
            oBkey = oB;
            const bool alive = (ci.w >> 31) != 0u;
            float4 hist = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODEL == 0 && alive) hist = P.st.hist[c];
            const float4 compA = __ldg(&P.comp[ci.z & 0xffffu]);
            OwnerPos pA, pB;
            End A, B;
            load_owner(P.state, oA, pA, A);
            load_owner(P.state, oB, pB, B);
            const float3 relA = rotate(f3(compA.x, compA.y, compA.z), A.q);
            long long ax, ay, az, bx, by, bz;
            pos_ints(pA, P.nvXp2, P.nvYp2, ax, ay, az);
            pos_ints(pB, P.nvXp2, P.nvYp2, bx, by, bz);
            const D3 ownB = {(double)(bx - ax) * P.l, (double)(by - ay) * P.l, (double)(bz - az) * P.l};  // ownerB - ownerA
            const D3 n1 = rotate_d(__ldg(&P.tri_n1[tri]), B.q) + ownB;
            const D3 n2 = rotate_d(__ldg(&P.tri_n2[tri]), B.q) + ownB;
            const D3 n3 = rotate_d(__ldg(&P.tri_n3[tri]), B.q) + ownB;
            const D3 sp = {(double)relA.x, (double)relA.y, (double)relA.z};
            const double radius = (double)compA.w;
            D3 fn = dcross(n2 - n1, n3 - n1);
            const double invLen = (double)(1.0f / sqrtf((float)ddot(fn, fn)));  // normalize(double3) is float precision
            fn = fn * invLen;
            const double hgt = ddot(sp - n1, fn);
            D3 cp, nd;
            double dpen;  // signed distance minus radius (negative = overlap)
            bool in_contact;
            if (!snap_to_face(n1, n2, n3, sp, cp)) {
                dpen = hgt - radius;
                nd = fn;
                in_contact = !(hgt >= radius || hgt <= -radius);
            } else {
                const D3 d = sp - cp;
                const double dist = sqrt(ddot(d, d));
                dpen = dist - radius;
                nd = d * (1.0 / dist);
                in_contact = !(dpen >= 0. || hgt >= radius || hgt <= -radius);
            }
            if (in_contact && dpen < 0.) {
                const float3 nrm = f3((float)nd.x, (float)nd.y, (float)nd.z);
                const float3 armA = f3((float)cp.x, (float)cp.y, (float)cp.z);
                const float3 armB = f3((float)(cp.x - ownB.x), (float)(cp.y - ownB.y), (float)(cp.z - ownB.z));
                const MatPair mp = P.matpair[ci.w & 0xffffu];
                float3 force, troll;
                contact_model<MODEL>(mp, P.h, (float)(-dpen), nrm, armA, armB, A, B, compA.w, 1e15f, hist, force, troll);
                const float3 Ft = force + troll;
                const float3 TA = cross(armA, Ft);
                const float3 TB = cross(Ft, armB);
                red_add_v4(&P.wrench[oA].f, force.x, force.y, force.z);
                red_add_v4(&P.wrench[oA].t, TA.x, TA.y, TA.z);
                wB[0] = -force.x; wB[1] = -force.y; wB[2] = -force.z;
                wB[3] = TB.x; wB[4] = TB.y; wB[5] = TB.z;
                touchB = true;
                if (MODEL == 0) {
                    P.st.hist[c] = hist;
                    if (!alive) P.st.cinfo[c].w = ci.w | 0x80000000u;
                }
                if (RECORD) {
                    P.st.force[c] = make_float4(force.x, force.y, force.z, 0.f);
                    P.st.cpoint[c] = contact_point_world(P, pA, armA);
                }
            } else {
                if (MODEL == 0 && alive) {
                    P.st.hist[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    P.st.cinfo[c].w = ci.w & 0x7fffffffu;
                }
                if (RECORD) P.st.force[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        wall_add(wall, P, touchB, oBkey, wB);
    }
    wall_flush(wall, P);
}

// ---------------------------------------------------------------------------------------------------------------
// integrateOwners (DEMIntegrationKernels.cu:100-264). One thread per owner: 64-byte state + 16-byte body-frame spin
// + 32-byte wrench in, state + spin out, wrench zeroed (prepareAccArrays, DEMPrepForceKernels.cu:14-37, fused).
__device__ __forceinline__ void st_v8(void* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e),
                 "f"(f), "f"(g), "f"(h)
                 : "memory");
}

__device__ __forceinline__ float integrate_one(const DevParams& P, const uint32_t o) {
    OwnerPos pos;
    End e;
    load_owner(P.state, o, pos, e);
    const float4 spin = P.spin[o];  // body-frame angular velocity, w = bits(inertiaPropOffset)
    const float4 mp = __ldg(&P.massprop[__float_as_uint(spin.w)]);
    float wf[8];
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(wf[0]), "=f"(wf[1]), "=f"(wf[2]), "=f"(wf[3]), "=f"(wf[4]), "=f"(wf[5]), "=f"(wf[6]), "=f"(wf[7])
                 : "l"(P.wrench + o));
    const float h = P.h;

    bool LinVelP[3] = {false, false, false}, RotVelP[3] = {false, false, false}, LinP[3] = {false, false, false};
    bool RotP = false;
    double X[3];
    pos_decode(pos, P, X[0], X[1], X[2]);
    X[0] += (double)P.LBF[0];
    X[1] += (double)P.LBF[1];
    X[2] += (double)P.LBF[2];
    float v[3] = {e.v.x, e.v.y, e.v.z}, w[3] = {spin.x, spin.y, spin.z};
    float oldv[3] = {v[0], v[1], v[2]}, oldw[3] = {w[0], w[1], w[2]};
    float extra_acc[3] = {0.f, 0.f, 0.f}, extra_ang[3] = {0.f, 0.f, 0.f};
    const Prescr* pr = P.presc + pos.family;
    if (pr->used) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (pr->hasLinVel[k]) v[k] = pr->linVel[k];
            if (pr->hasRotVel[k]) w[k] = pr->rotVel[k];
            LinVelP[k] = pr->linVelPrescribed[k];
            RotVelP[k] = pr->rotVelPrescribed[k];
            if (pr->hasLinPos[k]) X[k] = pr->linPos[k];
            LinP[k] = pr->linPosPrescribed[k];
            if (pr->hasAcc[k]) extra_acc[k] = pr->acc[k];
            if (pr->hasAngAcc[k]) extra_ang[k] = pr->angAcc[k];
        }
        RotP = pr->rotPosPrescribed;
    }
    // a = F/m ; alpha = R^T T_world / I  (the reference accumulates F_i/m and R^T(c x F_i)/I per contact; same sums up
    // to rounding)
    const float acc[3] = {wf[0] / mp.x, wf[1] / mp.x, wf[2] / mp.x};
    const float3 Tb = rotate_inv(f3(wf[4], wf[5], wf[6]), e.q);
    const float ang[3] = {Tb.x / mp.y, Tb.y / mp.z, Tb.z / mp.w};
    float vup[3] = {0.f, 0.f, 0.f}, wup[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!LinVelP[k]) {
            vup[k] = (acc[k] + extra_acc[k] + P.G[k]) * h;
            v[k] += vup[k];
        } else {
            oldv[k] = v[k];
        }
        if (!RotVelP[k]) {
            wup[k] = (ang[k] + extra_ang[k]) * h;
            w[k] += wup[k];
        } else {
            oldw[k] = w[k];
        }
    }
    float vp[3], wp[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (P.integrator == DEM_EXTENDED_TAYLOR) {
            vp[k] = oldv[k] + vup[k] * 0.5f;
            wp[k] = oldw[k] + wup[k] * 0.5f;
        } else if (P.integrator == DEM_CENTERED_DIFFERENCE) {
            vp[k] = oldv[k] + vup[k];
            wp[k] = oldw[k] + wup[k];
        } else {
            vp[k] = oldv[k];
            wp[k] = oldw[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!LinP[k]) X[k] = __dadd_rn(X[k], __dmul_rn((double)vp[k], (double)h));
        X[k] -= (double)P.LBF[k];
    }
    if (P.fast_encode) pos_encode_fast(pos, P, X[0], X[1], X[2]); else pos_encode(pos, P, X[0], X[1], X[2]);
    float4 q = e.q;
    if (!RotP) {
        const float hh = P.half_h;
        const float b2 = hh * wp[0], c2 = hh * wp[1], d2 = hh * wp[2];
        const float a1 = q.x, b1 = q.y, c1 = q.z, d1 = q.w;
        // Hamilton product q * (1, ha) and renormalisation with every product and sum rounded separately (no FMA
        // contraction), in the reference's association order (DEMHelperKernels.cuh:228-245): an owner with a
        // prescribed spin then carries bit for bit the orientation the reference's arithmetic gives it -- one ulp of a
        // float quaternion moves a facet of a 0.1 m drum by 1e-8 m, 0.1 % of a typical contact overlap.
        const float Aq = __fsub_rn(__fsub_rn(__fsub_rn(a1, __fmul_rn(b1, b2)), __fmul_rn(c1, c2)), __fmul_rn(d1, d2));
        const float Bq = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(a1, b2), b1), __fmul_rn(c1, d2)), __fmul_rn(d1, c2));
        const float Cq = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(a1, c2), __fmul_rn(b1, d2)), c1), __fmul_rn(d1, b2));
        const float Dq = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(a1, d2), __fmul_rn(b1, c2)), __fmul_rn(c1, b2)), d1);
        const float len = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Bq, Bq), __fmul_rn(Cq, Cq)), __fmul_rn(Dq, Dq)), __fmul_rn(Aq, Aq)));
        q = make_float4(Aq / len, Bq / len, Cq / len, Dq / len);
    }
    // world-frame angular velocity of the NEW state for the next force evaluation
    const float3 ww = rotate(f3(w[0], w[1], w[2]), q);
    float* dst = reinterpret_cast<float*>(P.state + o);
    const uint32_t p2 = (uint32_t)pos.lx | ((uint32_t)pos.ly << 16);
    const uint32_t p3 = (uint32_t)pos.lz | ((uint32_t)pos.family << 16) | ((uint32_t)pos.flags << 24);
    st_v8(dst, __uint_as_float((uint32_t)(pos.voxel & 0xffffffffull)), __uint_as_float((uint32_t)(pos.voxel >> 32)),
          __uint_as_float(p2), __uint_as_float(p3), q.x, q.y, q.z, q.w);
    st_v8(dst + 8, v[0], v[1], v[2], e.mass, ww.x, ww.y, ww.z, 0.f);
    P.spin[o] = make_float4(w[0], w[1], w[2], spin.w);
    // per-owner acceleration read-out (ContactAcc / ContactAngAccLocal trackers), only when requested
    if (P.acc_out) st_v8(P.acc_out + o, acc[0], acc[1], acc[2], 0.f, ang[0], ang[1], ang[2], 0.f);
    // consume the wrench: the accumulator is zero again for the next step's reductions
    st_v8(P.wrench + o, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
    return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

// MODE 0: every owner (single GPU).  Decomposed runs split the step's integration in two so that the halo exchange can
// run beside the bulk of it: MODE 1 = the owners in this rank's halo send lists (first), MODE 2 = all other active owners
// (ghosts only have their unused wrench consumed).
template <int MODE>
__global__ void __launch_bounds__(256, 4) k_integrate(const __grid_constant__ DevParams P) {
    if (P.flags[DEM_FLAG_POISON]) return;
    // max |v| bookkeeping for the contact margin (replaces the absv inspector + cub max of kT.cpp:125-149):
    // this step accumulates into maxvel_next; the slot of the state being left behind is zeroed for the step after.
    if (MODE != 1 && blockIdx.x == 0 && threadIdx.x == 0) {
        *P.maxvel = 0.f;
        P.flags[DEM_FLAG_CYCLE_STEP] += 1u;  // (read by the next step's force kernel: stream order)
    }
    float absv = 0.f;
    if (MODE == 1) {
        const uint32_t nL = P.halo_counts[1], nR = P.halo_counts[2];
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nL + nR; t += gridDim.x * blockDim.x) {
            const uint32_t o = (t < nL) ? P.halo_gid[0][t] : P.halo_gid[1][t - nL];
            if (t >= nL && P.active[o] == 3) continue;  // in both lists: done as a member of the left one
            const float a = integrate_one(P, o);
            if (!isfinite(a) || a > P.errOutVel) atomicOr(&P.flags[DEM_FLAG_VELOCITY], 1u);
            if (isfinite(a)) absv = fmaxf(absv, a);
        }
    } else {
        const uint32_t n = P.active_list ? *P.nActivePtr : P.nOwners;
        auto one = [&](uint32_t t) {
            const uint32_t o = P.active_list ? P.active_list[t] : t;
            float a = 0.f;
            const uint32_t role = P.active ? P.active[o] : 1u;
            if (role == 2u) {
                // ghost: its state arrives from the owning rank; only consume the (unused) wrench
                st_v8(P.wrench + o, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
            } else if (MODE == 0 || role == 1u) {
                a = integrate_one(P, o);
            }
            if (!isfinite(a) || a > P.errOutVel) atomicOr(&P.flags[DEM_FLAG_VELOCITY], 1u);
            if (isfinite(a)) absv = fmaxf(absv, a);
        };
        const uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (MODE == 0) {
            if (t0 < n) one(t0);  // launched with one thread per owner
        } else {
            for (uint32_t t = t0; t < n; t += gridDim.x * blockDim.x) one(t);
        }
    }
    // non-negative floats order like their bit patterns: integer max in the warp (REDUX), then in the CTA, then ONE
    // atomic per CTA (a per-warp atomic on a single address serialises 30k updates in L2)
    __shared__ int s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const int wv = __reduce_max_sync(0xffffffffu, __float_as_int(absv));
    if ((threadIdx.x & 31) == 0 && wv > 0) atomicMax(&s_max, wv);
    __syncthreads();
    if (threadIdx.x == 0 && s_max > 0) atomicMax(reinterpret_cast<int*>(P.maxvel_next), s_max);
}

// ---------------------------------------------------------------------------------------------------------------
// Which sphere--sphere kernel: the one that leaves candidates alone until they are due needs motion-bounded margins
// (beta < 0) and lists that live long enough for the skipping to pay for the scattered record fetches (see the comment
// on k_force_ss); force_opts bit 6 forces it (tests, measurements).
static bool use_due_kernel(const DevParams& P) {
    const bool can = (P.force_opts & 1u) && (P.force_opts & 16u) && P.beta < 0.f && P.sn.due != nullptr;
    return can && (P.maxDrift >= 32u || (P.force_opts & 64u));
}

template <int MINB, bool FAST>
static void launch_force_ss_t(const DevParams& P, int model, bool record, int grid, cudaStream_t s) {
    const int block = 256;
    if (use_due_kernel(P)) {
        if (model == DEM_HERTZIAN) {
            if (record) k_force_ss_due<0, true, MINB, FAST><<<grid, block, 0, s>>>(P); else k_force_ss_due<0, false, MINB, FAST><<<grid, block, 0, s>>>(P);
        } else {
            if (record) k_force_ss_due<1, true, MINB, FAST><<<grid, block, 0, s>>>(P); else k_force_ss_due<1, false, MINB, FAST><<<grid, block, 0, s>>>(P);
        }
        return;
    }
    if (model == DEM_HERTZIAN) {
        if (record) k_force_ss<0, true, MINB, FAST><<<grid, block, 0, s>>>(P); else k_force_ss<0, false, MINB, FAST><<<grid, block, 0, s>>>(P);
    } else {
        if (record) k_force_ss<1, true, MINB, FAST><<<grid, block, 0, s>>>(P); else k_force_ss<1, false, MINB, FAST><<<grid, block, 0, s>>>(P);
    }
}

// ctas_per_sm CTAs per SM (2, 3 or 4 -- selects the register budget the kernel was compiled for)
void launch_force_ss(const DevParams& P, int model, bool record, int num_sms, int ctas_per_sm, bool fast, cudaStream_t s) {
    const int grid = num_sms * ctas_per_sm;
    if (fast) {
        if (ctas_per_sm >= 4) launch_force_ss_t<4, true>(P, model, record, grid, s);
        else if (ctas_per_sm == 3) launch_force_ss_t<3, true>(P, model, record, grid, s);
        else launch_force_ss_t<2, true>(P, model, record, grid, s);
    } else {
        if (ctas_per_sm >= 4) launch_force_ss_t<4, false>(P, model, record, grid, s);
        else if (ctas_per_sm == 3) launch_force_ss_t<3, false>(P, model, record, grid, s);
        else launch_force_ss_t<2, false>(P, model, record, grid, s);
    }
}

void launch_force_sa(const DevParams& P, int model, bool record, int grid, cudaStream_t s) {
    const int block = 256;
    if (model == DEM_HERTZIAN) {
        if (record) k_force_sa<0, true><<<grid, block, 0, s>>>(P); else k_force_sa<0, false><<<grid, block, 0, s>>>(P);
    } else {
        if (record) k_force_sa<1, true><<<grid, block, 0, s>>>(P); else k_force_sa<1, false><<<grid, block, 0, s>>>(P);
    }
}

void launch_force_st(const DevParams& P, int model, bool record, int grid, cudaStream_t s) {
    const int block = 256;
    if (model == DEM_HERTZIAN) {
        if (record) k_force_st<0, true><<<grid, block, 0, s>>>(P); else k_force_st<0, false><<<grid, block, 0, s>>>(P);
    } else {
        if (record) k_force_st<1, true><<<grid, block, 0, s>>>(P); else k_force_st<1, false><<<grid, block, 0, s>>>(P);
    }
}

void launch_integrate(const DevParams& P, int grid_hint, cudaStream_t s) {
    const int block = 256;
    // single GPU: one owner per thread.  Decomposed: the number of active owners lives on the device and the kernel
    // strides over it; the host sizes the grid from the count the last confirmed rebuild reported (grid_hint, with slack,
    // in coarse steps so that a captured cycle stays valid) -- one owner per thread then, because the loads of one
    // owner form a dependent chain, without the thousands of empty CTAs a full-size grid would launch.
    int grid = (int)((P.nOwners + block - 1) / block);
    if (P.active_list && grid_hint > 0) grid = std::min(grid, grid_hint);
    if (grid <= 0) return;
    if (P.active_list && P.halo_counts) k_integrate<2><<<grid, block, 0, s>>>(P);
    else k_integrate<0><<<grid, block, 0, s>>>(P);
}
// decomposed runs, first half of the integration: the owners in the halo send lists
void launch_integrate_halo(const DevParams& P, int grid_hint, cudaStream_t s) {
    if (!P.halo_counts) return;
    k_integrate<1><<<std::max(1, std::min(grid_hint, 1024)), 256, 0, s>>>(P);
}

}  // namespace demb
