/*
 * dem_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; see dem_oracle.h).
 *
 * Restates, in plain C, the arithmetic of the reference DEM hot path.  Every function
 * cites the reference file:line (relative to /root/reference) it follows.  Float/double
 * mixing follows the C++ overload resolution of the reference kernels exactly, so that the
 * results are bit-identical to the host-compiled reference text (oracle/_ref) when both are
 * built with -ffp-contract=off.
 */
#include "dem_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TINY_FLOAT 1e-12 /* DEME_TINY_FLOAT (double literal), src/DEM/Defines.h:29 */
#define HUGE_FLOAT 1e15  /* DEME_HUGE_FLOAT, src/DEM/Defines.h:30 */
static const double TWO_OVER_THREE = 2. / 3.;
static const double FOUR_OVER_THREE = 4. / 3.;
static const double FIVE_OVER_THREE = 5. / 3.;
static const double TWO_TIMES_SQRT_FIVE_OVER_SIX = 1.825741858350554;
static const double PI_ = 3.1415926535897932385;
static const double PI_SQUARED = 9.869604401089358; /* src/DEM/Defines.h:39-44 */

typedef struct { float x, y, z; } f3;
typedef struct { double x, y, z; } d3;

size_t orc_sizeof_world(void) { return sizeof(OrcWorld); }
size_t orc_sizeof_prescription(void) { return sizeof(OrcPrescription); }

/* ---------- small float3 helpers (src/kernel/CUDAMathHelpers.cuh:1032,1066,1182) ---------- */
static inline f3 f3_make(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 f3_add(f3 a, f3 b) { return f3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 f3_sub(f3 a, f3 b) { return f3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 f3_scale(f3 a, float s) { return f3_make(a.x * s, a.y * s, a.z * s); }
static inline float f3_dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float f3_len(f3 a) { return sqrtf(f3_dot(a, a)); }
static inline f3 f3_cross(f3 a, f3 b) {
    return f3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* applyOriQToVector3<float,float>, src/kernel/DEMHelperKernels.cuh:161-173 */
static inline void quat_rotate_f(float* X, float* Y, float* Z, float Qw, float Qx, float Qy, float Qz) {
    float oldX = *X, oldY = *Y, oldZ = *Z;
    *X = (2.0f * (Qw * Qw + Qx * Qx) - 1.0f) * oldX + (2.0f * (Qx * Qy - Qw * Qz)) * oldY +
         (2.0f * (Qx * Qz + Qw * Qy)) * oldZ;
    *Y = (2.0f * (Qx * Qy + Qw * Qz)) * oldX + (2.0f * (Qw * Qw + Qy * Qy) - 1.0f) * oldY +
         (2.0f * (Qy * Qz - Qw * Qx)) * oldZ;
    *Z = (2.0f * (Qx * Qz - Qw * Qy)) * oldX + (2.0f * (Qy * Qz + Qw * Qx)) * oldY +
         (2.0f * (Qw * Qw + Qz * Qz) - 1.0f) * oldZ;
}
/* applyOriQToVector3<double,float>: vector double, quaternion float (T2)2.0 constants are float
 * (src/kernel/DEMCalcForceKernels.cu:171-174 instantiates it with double vectors). */
static inline void quat_rotate_d(double* X, double* Y, double* Z, float Qw, float Qx, float Qy, float Qz) {
    double oldX = *X, oldY = *Y, oldZ = *Z;
    *X = (2.0f * (Qw * Qw + Qx * Qx) - 1.0f) * oldX + (2.0f * (Qx * Qy - Qw * Qz)) * oldY +
         (2.0f * (Qx * Qz + Qw * Qy)) * oldZ;
    *Y = (2.0f * (Qx * Qy + Qw * Qz)) * oldX + (2.0f * (Qw * Qw + Qy * Qy) - 1.0f) * oldY +
         (2.0f * (Qy * Qz - Qw * Qx)) * oldZ;
    *Z = (2.0f * (Qx * Qz - Qw * Qy)) * oldX + (2.0f * (Qy * Qz + Qw * Qx)) * oldY +
         (2.0f * (Qw * Qw + Qz * Qz) - 1.0f) * oldZ;
}

/* ---------- A.1 position codec (src/kernel/DEMHelperKernels.cuh:91-159) ---------- */
static inline void decode_raw(const OrcWorld* w, uint64_t ID, uint16_t sx, uint16_t sy, uint16_t sz, double* X,
                              double* Y, double* Z) {
    uint64_t vx = ID & (((uint64_t)1 << w->nvXp2) - 1);
    uint64_t vy = (ID >> w->nvXp2) & (((uint64_t)1 << w->nvYp2) - 1);
    uint64_t vz = ID >> (w->nvXp2 + w->nvYp2);
    *X = (double)vx * w->voxelSize + (double)sx * w->l;
    *Y = (double)vy * w->voxelSize + (double)sy * w->l;
    *Z = (double)vz * w->voxelSize + (double)sz * w->l;
}
void orc_voxel_decode(const OrcWorld* w, uint32_t o, double xyz[3]) {
    decode_raw(w, w->voxelID[o], w->locX[o], w->locY[o], w->locZ[o], &xyz[0], &xyz[1], &xyz[2]);
}
void orc_voxel_encode(const OrcWorld* w, const double xyz[3], uint64_t* voxel, uint16_t loc[3]) {
    uint64_t nx = (uint64_t)(xyz[0] / w->voxelSize);
    uint64_t ny = (uint64_t)(xyz[1] / w->voxelSize);
    uint64_t nz = (uint64_t)(xyz[2] / w->voxelSize);
    loc[0] = (uint16_t)((xyz[0] - (double)nx * w->voxelSize) / w->l);
    loc[1] = (uint16_t)((xyz[1] - (double)ny * w->voxelSize) / w->l);
    loc[2] = (uint16_t)((xyz[2] - (double)nz * w->voxelSize) / w->l);
    uint64_t ID = nx;
    ID += ny << w->nvXp2;
    ID += nz << (w->nvXp2 + w->nvYp2);
    *voxel = ID;
}
/* Initial placement: float world position minus float LBF, then encode (src/DEM/dT.cpp populateEntityArrays:
 * "this_CoM_coord = this_clump_xyz - LBF" in float, cast to double). */
void orc_encode_positions(OrcWorld* w, const float* xyz, uint32_t first, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        double p[3];
        for (int k = 0; k < 3; k++) {
            float rel = xyz[3 * i + k] - w->LBF[k];
            p[k] = (double)rel;
        }
        uint16_t loc[3];
        orc_voxel_encode(w, p, &w->voxelID[first + i], loc);
        w->locX[first + i] = loc[0];
        w->locY[first + i] = loc[1];
        w->locZ[first + i] = loc[2];
    }
}
/* Read-out as the reference reports it: float decode + LBF (src/DEM/dT.cpp:3062-3076). */
void orc_decode_positions(const OrcWorld* w, float* xyz, uint32_t first, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        double p[3];
        orc_voxel_decode(w, first + i, p);
        for (int k = 0; k < 3; k++)
            xyz[3 * i + k] = (float)(p[k] + (double)w->LBF[k]);
    }
}

/* locateMaskPair, src/kernel/DEMHelperKernels.cuh:57-62 */
static inline unsigned int mask_pair(unsigned int i, unsigned int j) {
    if (i > j) { unsigned int t = i; i = j; j = t; }
    return (1 + j) * j / 2 + i;
}

/* ---------- A.8 margin (src/kernel/DEMMiscKernels.cu:37-69; absv = |v|, src/DEM/AuxClasses.cpp:54-61) ---------- */
void orc_compute_margins(OrcWorld* w, uint32_t maxDrift) {
    for (uint32_t o = 0; o < w->nOwners; o++) {
        unsigned int fam = w->familyID[o];
        if (w->beta >= 0.f) { /* fillMarginValues */
            w->marginSize[o] = w->beta + w->familyExtraMarginSize[fam];
            continue;
        }
        float vx = w->vX[o], vy = w->vY[o], vz = w->vZ[o];
        float absv = sqrtf(vx * vx + vy * vy + vz * vz);
        if (absv > w->approxMaxVel) absv = w->approxMaxVel;
        w->marginSize[o] = (float)((double)(absv * w->expSafetyMulti + w->expSafetyAdder) * w->h * maxDrift +
                                   w->familyExtraMarginSize[fam]);
    }
}

/* ---------- A.2 sphere world position (src/kernel/DEMCalcForceKernels.cu:20-42) ---------- */
static inline void owner_pos_rot(const OrcWorld* w, uint32_t owner, f3* relPos, d3* ownerPos, d3* bodyPos, float q[4],
                                 int addLBF) {
    double p[3];
    orc_voxel_decode(w, owner, p);
    if (addLBF) {
        p[0] += w->LBF[0];
        p[1] += w->LBF[1];
        p[2] += w->LBF[2];
    }
    ownerPos->x = p[0]; ownerPos->y = p[1]; ownerPos->z = p[2];
    q[0] = w->oriQw[owner]; q[1] = w->oriQx[owner]; q[2] = w->oriQy[owner]; q[3] = w->oriQz[owner];
    quat_rotate_f(&relPos->x, &relPos->y, &relPos->z, q[0], q[1], q[2], q[3]);
    bodyPos->x = ownerPos->x + (double)relPos->x;
    bodyPos->y = ownerPos->y + (double)relPos->y;
    bodyPos->z = ownerPos->z + (double)relPos->z;
}

static int g_orc_threads = 1; /* host threads of the broad phase, see orc_set_threads */
void orc_sphere_positions(const OrcWorld* w, double* xyz, float* radius) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(g_orc_threads > 64 ? 64 : (g_orc_threads < 1 ? 1 : g_orc_threads))
#endif
    for (uint32_t s = 0; s < w->nSpheres; s++) {
        uint32_t o = w->ownerClumpBody[s];
        unsigned c = w->clumpComponentOffset[s];
        f3 rel = f3_make(w->CDRelPosX[c], w->CDRelPosY[c], w->CDRelPosZ[c]);
        d3 op, bp; float q[4];
        owner_pos_rot(w, o, &rel, &op, &bp, q, 0);
        xyz[3 * s + 0] = bp.x; xyz[3 * s + 1] = bp.y; xyz[3 * s + 2] = bp.z;
        if (radius) radius[s] = w->Radii[c];
    }
}

/* ---------- A.3 narrow phase ---------- */
/* checkSpheresOverlap<double,float>, src/kernel/DEMHelperKernels.cuh:292-326 */
static inline int spheres_overlap(double XA, double YA, double ZA, double radA, double XB, double YB, double ZB,
                                  double radB, d3* CP, f3* n, double* overlapDepth) {
    double centerDist2 = (XA - XB) * (XA - XB) + (YA - YB) * (YA - YB) + (ZA - ZB) * (ZA - ZB);
    int contact = !(centerDist2 > (radA + radB) * (radA + radB));
    float nx = (float)(XA - XB), ny = (float)(YA - YB), nz = (float)(ZA - ZB);
    float mag = sqrtf(nx * nx + ny * ny + nz * nz);
    nx /= mag; ny /= mag; nz /= mag;
    n->x = nx; n->y = ny; n->z = nz;
    *overlapDepth = radA + radB - sqrt(centerDist2);
    CP->x = XB + (radB - *overlapDepth / 2.0) * nx;
    CP->y = YB + (radB - *overlapDepth / 2.0) * ny;
    CP->z = ZB + (radB - *overlapDepth / 2.0) * nz;
    return contact;
}

/* checkSphereEntityOverlap<double3,float,double>, src/kernel/DEMHelperKernels.cuh:459-521.
 * radA is T2: float on the dT side, but the kT side passes a double (template deduces T2=double there);
 * both are served by passing the already-summed (radA + beta4Entity) term in the precision of the caller. */
static inline int sphere_entity_overlap(d3 A, double radA_plus_beta, double radA, uint8_t typeB, d3 B, f3 dirB,
                                        float size1B, float normal_sign, float beta4Entity, d3* CP, f3* cntNormal,
                                        double* overlapDepth, int radA_is_float) {
    switch (typeB) {
        case ORC_ANAL_PLANE: {
            d3 p2s = {A.x - B.x, A.y - B.y, A.z - B.z};
            /* dot(double3,float3) returns float, src/kernel/CUDAMathHelpers.cuh:1221 */
            float distf = (float)(p2s.x * dirB.x + p2s.y * dirB.y + p2s.z * dirB.z);
            double dist = distf;
            *overlapDepth = radA_plus_beta - dist;
            int type = (*overlapDepth < 0.0) ? ORC_NOT_A_CONTACT : ORC_SPHERE_PLANE;
            float s = (float)(dist + *overlapDepth / 2.0);
            CP->x = A.x - (double)(dirB.x * s);
            CP->y = A.y - (double)(dirB.y * s);
            CP->z = A.z - (double)(dirB.z * s);
            *cntNormal = dirB;
            return type;
        }
        case ORC_ANAL_CYL_INF: {
            d3 s2c = {B.x - A.x, B.y - A.y, B.z - A.z};
            float projf = (float)(s2c.x * dirB.x + s2c.y * dirB.y + s2c.z * dirB.z);
            double proj = projf;
            /* proj_dist * dirB : double scalar converts to float for operator*(float,float3);
             * double3 -= float3 (src/kernel/CUDAMathHelpers.cuh:1308) */
            float pf = (float)proj;
            s2c.x -= (double)(pf * dirB.x);
            s2c.y -= (double)(pf * dirB.y);
            s2c.z -= (double)(pf * dirB.z);
            double dist_delta_r = sqrt(s2c.x * s2c.x + s2c.y * s2c.y + s2c.z * s2c.z);
            float cyl_rad = size1B - normal_sign * beta4Entity;
            if (radA_is_float)
                *overlapDepth = (float)radA - normal_sign * (cyl_rad - dist_delta_r);
            else
                *overlapDepth = radA - normal_sign * (cyl_rad - dist_delta_r);
            int type = (*overlapDepth < 0.0) ? ORC_NOT_A_CONTACT : ORC_SPHERE_CYL;
            if (dist_delta_r >= TINY_FLOAT) {
                double k = normal_sign / dist_delta_r;
                cntNormal->x = (float)(k * s2c.x);
                cntNormal->y = (float)(k * s2c.y);
                cntNormal->z = (float)(k * s2c.z);
                float s = (float)(radA - *overlapDepth / 2.0);
                CP->x = A.x - (double)(cntNormal->x * s);
                CP->y = A.y - (double)(cntNormal->y * s);
                CP->z = A.z - (double)(cntNormal->z * s);
            } else {
                *cntNormal = dirB;
                *CP = A;
            }
            return type;
        }
        default:
            return ORC_NOT_A_CONTACT;
    }
}

/* snap_to_face<double3,double>, src/kernel/DEMCollisionKernels.cu:15-81 (Ericson p.141) */
static inline d3 d3_sub(d3 a, d3 b) { d3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline d3 d3_add(d3 a, d3 b) { d3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline d3 d3_scale(d3 a, double s) { d3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline double d3_dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline d3 d3_cross(d3 a, d3 b) {
    d3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static int snap_to_face(d3 A, d3 B, d3 C, d3 P, d3* res) {
    d3 AB = d3_sub(B, A), AC = d3_sub(C, A), AP = d3_sub(P, A);
    double d1 = d3_dot(AB, AP), d2 = d3_dot(AC, AP);
    if (d1 <= 0 && d2 <= 0) { *res = A; return 1; }
    d3 BP = d3_sub(P, B);
    double d3v = d3_dot(AB, BP), d4 = d3_dot(AC, BP);
    if (d3v >= 0 && d4 <= d3v) { *res = B; return 1; }
    double vc = d1 * d4 - d3v * d2;
    if (vc <= 0 && d1 >= 0 && d3v <= 0) {
        double v = d1 / (d1 - d3v);
        *res = d3_add(A, d3_scale(AB, v));
        return 1;
    }
    d3 CPv = d3_sub(P, C);
    double d5 = d3_dot(AB, CPv), d6 = d3_dot(AC, CPv);
    if (d6 >= 0 && d5 <= d6) { *res = C; return 1; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        double wv = d2 / (d2 - d6);
        *res = d3_add(A, d3_scale(AC, wv));
        return 1;
    }
    double va = d3v * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3v) >= 0 && (d5 - d6) >= 0) {
        double wv = (d4 - d3v) / ((d4 - d3v) + (d5 - d6));
        *res = d3_add(B, d3_scale(d3_sub(C, B), wv));
        return 1;
    }
    /* __drcp_ru / __dmul_ru are round-up device intrinsics; round-to-nearest here (<=1 ulp apart) */
    double denom = 1.0 / (va + vb + vc);
    double v = vb * denom, wv = vc * denom;
    *res = d3_add(d3_add(A, d3_scale(AB, v)), d3_scale(AC, wv));
    return 0;
}
/* triangle_sphere_CD<double3,double>, src/kernel/DEMCollisionKernels.cu:98-156 */
static int triangle_sphere_CD(d3 A, d3 B, d3 C, d3 sp, double radius, d3* normal, double* depth, d3* pt1) {
    d3 fn = d3_cross(d3_sub(B, A), d3_sub(C, A));
    /* normalize(double3) uses rsqrtf (float!), src/kernel/CUDAMathHelpers.cuh:1402-1405 */
    double invLen = 1.0f / sqrtf((float)d3_dot(fn, fn));
    fn = d3_scale(fn, invLen);
    double h = d3_dot(d3_sub(sp, A), fn);
    d3 faceLoc;
    int in_contact;
    if (!snap_to_face(A, B, C, sp, &faceLoc)) {
        *depth = h - radius;
        *normal = fn;
        *pt1 = faceLoc;
        in_contact = !(h >= radius || h <= -radius);
    } else {
        d3 nd = d3_sub(sp, faceLoc);
        double dist = sqrt(d3_dot(nd, nd));
        *depth = dist - radius;
        *normal = d3_scale(nd, 1.0 / dist);
        *pt1 = faceLoc;
        in_contact = !(*depth >= 0. || h >= radius || h <= -radius);
    }
    return in_contact;
}

/* matProxy2ContactParam<float>, src/kernel/DEMHelperKernels.cuh:433-455 */
static inline void mat_proxy(float* E_eff, float* G_eff, float Y1, float nu1, float Y2, float nu2) {
    float invE = (1.0f - nu1 * nu1) / Y1 + (1.0f - nu2 * nu2) / Y2;
    *E_eff = 1.0f / invE;
    if (G_eff) {
        float invG = 2.0f * (2.0f - nu1) * (1.0f + nu1) / Y1 + 2.0f * (2.0f - nu2) * (1.0f + nu2) / Y2;
        *G_eff = 1.0f / invG;
    }
}

/* ---------- prepareAccArrays, src/kernel/DEMPrepForceKernels.cu:14-37 ---------- */
void orc_prepare_acc(OrcWorld* w) {
    for (uint32_t o = 0; o < w->nOwners; o++) {
        if (w->accSpecified[o]) {
            w->accSpecified[o] = 0;
        } else {
            w->aX[o] = 0; w->aY[o] = 0; w->aZ[o] = 0;
        }
        if (w->angAccSpecified[o]) {
            w->angAccSpecified[o] = 0;
        } else {
            w->alphaX[o] = 0; w->alphaY[o] = 0; w->alphaZ[o] = 0;
        }
    }
}

/* ---------- calculateContactForces, src/kernel/DEMCalcForceKernels.cu:44-267
 *            with FullHertzianForceModel.cu / FrictionlessHertzianForceModel.cu pasted at _DEMForceModel_ ---------- */
static void calc_one_contact(OrcWorld* w, uint64_t cid) {
    uint8_t ContactType = w->contactType[cid];
    d3 contactPnt = {0, 0, 0};
    f3 B2A = {0, 0, 0};
    double overlapDepth = 0;
    d3 AOwnerPos, bodyAPos, BOwnerPos = {0, 0, 0}, bodyBPos;
    float AOwnerMass, ARadius, BOwnerMass = 0, BRadius = 0;
    float AOriQ[4], BOriQ[4] = {1, 0, 0, 0};
    unsigned bodyAMatType, bodyBMatType = 0;
    float extraMarginSize;
    f3 ALinVel, ARotVel, BLinVel = {0, 0, 0}, BRotVel = {0, 0, 0};
    unsigned AOwnerFamily, BOwnerFamily;
    const float ts = w->h;
    {
        uint32_t sphereID = w->idGeometryA[cid];
        uint32_t myOwner = w->ownerClumpBody[sphereID];
        unsigned c = w->clumpComponentOffset[sphereID];
        f3 myRelPos = f3_make(w->CDRelPosX[c], w->CDRelPosY[c], w->CDRelPosZ[c]);
        ARadius = w->Radii[c];
        AOwnerMass = w->MassProperties[w->inertiaPropOffsets[myOwner]];
        AOwnerFamily = w->familyID[myOwner];
        ALinVel = f3_make(w->vX[myOwner], w->vY[myOwner], w->vZ[myOwner]);
        ARotVel = f3_make(w->omgBarX[myOwner], w->omgBarY[myOwner], w->omgBarZ[myOwner]);
        owner_pos_rot(w, myOwner, &myRelPos, &AOwnerPos, &bodyAPos, AOriQ, 1);
        bodyAMatType = w->sphereMaterialOffset[sphereID];
        extraMarginSize = w->familyExtraMarginSize[AOwnerFamily];
    }
    if (ContactType == ORC_SPHERE_SPHERE) {
        uint32_t sphereID = w->idGeometryB[cid];
        uint32_t myOwner = w->ownerClumpBody[sphereID];
        unsigned c = w->clumpComponentOffset[sphereID];
        f3 myRelPos = f3_make(w->CDRelPosX[c], w->CDRelPosY[c], w->CDRelPosZ[c]);
        BRadius = w->Radii[c];
        BOwnerMass = w->MassProperties[w->inertiaPropOffsets[myOwner]];
        BOwnerFamily = w->familyID[myOwner];
        BLinVel = f3_make(w->vX[myOwner], w->vY[myOwner], w->vZ[myOwner]);
        BRotVel = f3_make(w->omgBarX[myOwner], w->omgBarY[myOwner], w->omgBarZ[myOwner]);
        owner_pos_rot(w, myOwner, &myRelPos, &BOwnerPos, &bodyBPos, BOriQ, 1);
        bodyBMatType = w->sphereMaterialOffset[sphereID];
        if (!(extraMarginSize > w->familyExtraMarginSize[BOwnerFamily]))
            extraMarginSize = w->familyExtraMarginSize[BOwnerFamily];
        spheres_overlap(bodyAPos.x, bodyAPos.y, bodyAPos.z, ARadius, bodyBPos.x, bodyBPos.y, bodyBPos.z, BRadius,
                        &contactPnt, &B2A, &overlapDepth);
        if (overlapDepth < -extraMarginSize) ContactType = ORC_NOT_A_CONTACT;
    } else if (ContactType == ORC_SPHERE_MESH) {
        uint32_t triID = w->idGeometryB[cid];
        uint32_t myOwner = w->ownerMesh[triID];
        BRadius = (float)HUGE_FLOAT;
        bodyBMatType = w->triMaterialOffset[triID];
        BOwnerFamily = w->familyID[myOwner];
        if (!(extraMarginSize > w->familyExtraMarginSize[BOwnerFamily]))
            extraMarginSize = w->familyExtraMarginSize[BOwnerFamily];
        d3 n1 = {w->relPosNode1[3 * triID], w->relPosNode1[3 * triID + 1], w->relPosNode1[3 * triID + 2]};
        d3 n2 = {w->relPosNode2[3 * triID], w->relPosNode2[3 * triID + 1], w->relPosNode2[3 * triID + 2]};
        d3 n3 = {w->relPosNode3[3 * triID], w->relPosNode3[3 * triID + 1], w->relPosNode3[3 * triID + 2]};
        BOwnerMass = w->MassProperties[w->inertiaPropOffsets[myOwner]];
        BLinVel = f3_make(w->vX[myOwner], w->vY[myOwner], w->vZ[myOwner]);
        BRotVel = f3_make(w->omgBarX[myOwner], w->omgBarY[myOwner], w->omgBarZ[myOwner]);
        /* equipOwnerPosRot<double3>: relPos is double3 here, rotation on doubles with float quaternion */
        double p[3];
        orc_voxel_decode(w, myOwner, p);
        BOwnerPos.x = p[0] + w->LBF[0]; BOwnerPos.y = p[1] + w->LBF[1]; BOwnerPos.z = p[2] + w->LBF[2];
        BOriQ[0] = w->oriQw[myOwner]; BOriQ[1] = w->oriQx[myOwner];
        BOriQ[2] = w->oriQy[myOwner]; BOriQ[3] = w->oriQz[myOwner];
        quat_rotate_d(&n1.x, &n1.y, &n1.z, BOriQ[0], BOriQ[1], BOriQ[2], BOriQ[3]);
        n1 = d3_add(BOwnerPos, n1);
        quat_rotate_d(&n2.x, &n2.y, &n2.z, BOriQ[0], BOriQ[1], BOriQ[2], BOriQ[3]);
        n2 = d3_add(n2, BOwnerPos);
        quat_rotate_d(&n3.x, &n3.y, &n3.z, BOriQ[0], BOriQ[1], BOriQ[2], BOriQ[3]);
        n3 = d3_add(n3, BOwnerPos);
        d3 cn;
        int in_contact = triangle_sphere_CD(n1, n2, n3, bodyAPos, ARadius, &cn, &overlapDepth, &contactPnt);
        B2A = f3_make((float)cn.x, (float)cn.y, (float)cn.z);
        if ((overlapDepth > extraMarginSize) || (!in_contact && overlapDepth < 0.)) ContactType = ORC_NOT_A_CONTACT;
        overlapDepth = -overlapDepth;
    } else if (ContactType > 10) {
        unsigned objID = w->idGeometryB[cid] & 0xFF; /* objID_t is uint8_t */
        uint32_t myOwner = w->objOwner[objID];
        bodyBMatType = w->objMaterial[objID];
        BOwnerMass = w->objMass[objID];
        BRadius = (float)HUGE_FLOAT;
        f3 myRelPos = f3_make(w->objRelPosX[objID], w->objRelPosY[objID], w->objRelPosZ[objID]);
        BOwnerFamily = w->familyID[myOwner];
        BLinVel = f3_make(w->vX[myOwner], w->vY[myOwner], w->vZ[myOwner]);
        BRotVel = f3_make(w->omgBarX[myOwner], w->omgBarY[myOwner], w->omgBarZ[myOwner]);
        owner_pos_rot(w, myOwner, &myRelPos, &BOwnerPos, &bodyBPos, BOriQ, 1);
        if (!(extraMarginSize > w->familyExtraMarginSize[BOwnerFamily]))
            extraMarginSize = w->familyExtraMarginSize[BOwnerFamily];
        f3 rot = f3_make(w->objRotX[objID], w->objRotY[objID], w->objRotZ[objID]);
        quat_rotate_f(&rot.x, &rot.y, &rot.z, BOriQ[0], BOriQ[1], BOriQ[2], BOriQ[3]);
        /* radA (float) + beta4Entity (0.0 -> float param) is a float sum */
        float rpb = ARadius + 0.0f;
        sphere_entity_overlap(bodyAPos, (double)rpb, (double)ARadius, w->objType[objID], bodyBPos, rot,
                              w->objSize1[objID], w->objNormal[objID], 0.0f, &contactPnt, &B2A, &overlapDepth, 1);
        if (overlapDepth < -extraMarginSize) ContactType = ORC_NOT_A_CONTACT;
    }

    float delta_tan_x = 0, delta_tan_y = 0, delta_tan_z = 0, delta_time = 0;
    const int history = (w->force_model == ORC_HERTZIAN);
    if (history) {
        delta_tan_x = w->contactWildcards[0][cid];
        delta_tan_y = w->contactWildcards[1][cid];
        delta_tan_z = w->contactWildcards[2][cid];
        delta_time = w->contactWildcards[3][cid];
    }
    if (ContactType != ORC_NOT_A_CONTACT) {
        f3 force = {0, 0, 0}, torque_only_force = {0, 0, 0};
        f3 locCPA = f3_make((float)(contactPnt.x - AOwnerPos.x), (float)(contactPnt.y - AOwnerPos.y),
                            (float)(contactPnt.z - AOwnerPos.z));
        f3 locCPB = f3_make((float)(contactPnt.x - BOwnerPos.x), (float)(contactPnt.y - BOwnerPos.y),
                            (float)(contactPnt.z - BOwnerPos.z));
        quat_rotate_f(&locCPA.x, &locCPA.y, &locCPA.z, AOriQ[0], -AOriQ[1], -AOriQ[2], -AOriQ[3]);
        quat_rotate_f(&locCPB.x, &locCPB.y, &locCPB.z, BOriQ[0], -BOriQ[1], -BOriQ[2], -BOriQ[3]);

        if (overlapDepth > 0) {
            float E_cnt, G_cnt = 0, CoR_cnt, mu_cnt = 0, Crr_cnt = 0;
            const unsigned nM = w->nMat;
            mat_proxy(&E_cnt, history ? &G_cnt : NULL, w->E[bodyAMatType], w->nu[bodyAMatType], w->E[bodyBMatType],
                      w->nu[bodyBMatType]);
            CoR_cnt = w->CoR[bodyAMatType * nM + bodyBMatType];
            if (history) {
                mu_cnt = w->mu[bodyAMatType * nM + bodyBMatType];
                Crr_cnt = w->Crr[bodyAMatType * nM + bodyBMatType];
            }
            f3 rotVelCPA = f3_cross(ARotVel, locCPA);
            f3 rotVelCPB = f3_cross(BRotVel, locCPB);
            quat_rotate_f(&rotVelCPA.x, &rotVelCPA.y, &rotVelCPA.z, AOriQ[0], AOriQ[1], AOriQ[2], AOriQ[3]);
            quat_rotate_f(&rotVelCPB.x, &rotVelCPB.y, &rotVelCPB.z, BOriQ[0], BOriQ[1], BOriQ[2], BOriQ[3]);

            const f3 velB2A = f3_sub(f3_add(ALinVel, rotVelCPA), f3_add(BLinVel, rotVelCPB));
            const float projection = f3_dot(velB2A, B2A);
            f3 vrel_tan = f3_sub(velB2A, f3_scale(B2A, projection));
            f3 delta_tan = f3_make(delta_tan_x, delta_tan_y, delta_tan_z);
            if (history) {
                delta_tan = f3_add(delta_tan, f3_scale(vrel_tan, ts));
                const float disp_proj = f3_dot(delta_tan, B2A);
                delta_tan = f3_sub(delta_tan, f3_scale(B2A, disp_proj));
                delta_time += ts;
            }
            const float mass_eff = (AOwnerMass * BOwnerMass) / (AOwnerMass + BOwnerMass);
            const float sqrt_Rd =
                (float)sqrt(overlapDepth * (double)(ARadius * BRadius) / (double)(ARadius + BRadius));
            const float Sn = (float)(2. * E_cnt * sqrt_Rd);
            const float loge = (float)((CoR_cnt < TINY_FLOAT) ? log(TINY_FLOAT) : (double)logf(CoR_cnt));
            const float beta = (float)(loge / sqrt(loge * loge + PI_SQUARED));
            const float k_n = (float)(TWO_OVER_THREE * Sn);
            const float gamma_n = (float)(TWO_TIMES_SQRT_FIVE_OVER_SIX * beta * sqrtf(Sn * mass_eff));
            force = f3_add(force, f3_scale(B2A, (float)(k_n * overlapDepth + gamma_n * projection)));

            if (history && Crr_cnt > 0.0) {
                int should_add = 1;
                {
                    const float R_eff = sqrtf((ARadius * BRadius) / (ARadius + BRadius));
                    const float kn_simple = (float)(FOUR_OVER_THREE * E_cnt * sqrtf(R_eff));
                    const float gn_simple =
                        -2.f * sqrtf((float)(FIVE_OVER_THREE * mass_eff * E_cnt)) * beta * powf(R_eff, 0.25f);
                    const float d_coeff = gn_simple / (2.f * sqrtf(kn_simple * mass_eff));
                    if (d_coeff < 1.0) {
                        float t_collision =
                            (float)(PI_ * sqrtf(mass_eff / (kn_simple * (1.f - d_coeff * d_coeff))));
                        if (delta_time <= t_collision) should_add = 0;
                    }
                }
                if (should_add) {
                    const f3 v_rot = f3_sub(rotVelCPB, rotVelCPA);
                    const float v_rot_mag = f3_len(v_rot);
                    if (v_rot_mag > TINY_FLOAT) {
                        /* (v_rot / v_rot_mag) * (Crr * |force|): float3/float is a*(1/b) in
                         * CUDAMathHelpers? -> it is component-wise division; see ref check */
                        f3 dir = f3_make(v_rot.x / v_rot_mag, v_rot.y / v_rot_mag, v_rot.z / v_rot_mag);
                        torque_only_force = f3_scale(dir, Crr_cnt * f3_len(force));
                    }
                }
            }
            if (history && mu_cnt > 0.0) {
                const float kt = (float)(8. * G_cnt * sqrt_Rd);
                const float gt = (float)(-TWO_TIMES_SQRT_FIVE_OVER_SIX * beta * sqrtf(mass_eff * kt));
                f3 tangent_force = f3_sub(f3_scale(delta_tan, -kt), f3_scale(vrel_tan, gt));
                const float ft = f3_len(tangent_force);
                if (ft > TINY_FLOAT) {
                    const float ft_max = f3_len(force) * mu_cnt;
                    if (ft > ft_max) {
                        tangent_force = f3_scale(tangent_force, ft_max / ft);
                        f3 num = f3_add(tangent_force, f3_scale(vrel_tan, gt));
                        float d = -kt;
                        delta_tan = f3_make(num.x / d, num.y / d, num.z / d);
                    }
                } else {
                    tangent_force = f3_make(0, 0, 0);
                }
                force = f3_add(force, tangent_force);
            }
            delta_tan_x = delta_tan.x; delta_tan_y = delta_tan.y; delta_tan_z = delta_tan.z;
        } else if (history) {
            delta_time = 0; delta_tan_x = 0; delta_tan_y = 0; delta_tan_z = 0;
        }
        /* _contactInfoWrite_, DEMCustomizablePolicies/ContactInfoWriteBack.cu */
        w->contactPointGeometryA[3 * cid + 0] = locCPA.x;
        w->contactPointGeometryA[3 * cid + 1] = locCPA.y;
        w->contactPointGeometryA[3 * cid + 2] = locCPA.z;
        w->contactPointGeometryB[3 * cid + 0] = locCPB.x;
        w->contactPointGeometryB[3 * cid + 1] = locCPB.y;
        w->contactPointGeometryB[3 * cid + 2] = locCPB.z;
        w->contactForces[3 * cid + 0] = force.x;
        w->contactForces[3 * cid + 1] = force.y;
        w->contactForces[3 * cid + 2] = force.z;
        w->contactTorque_convToForce[3 * cid + 0] = torque_only_force.x;
        w->contactTorque_convToForce[3 * cid + 1] = torque_only_force.y;
        w->contactTorque_convToForce[3 * cid + 2] = torque_only_force.z;
    } else {
        delta_tan_x = 0; delta_tan_y = 0; delta_tan_z = 0; delta_time = 0;
    }
    if (history) {
        w->contactWildcards[0][cid] = delta_tan_x;
        w->contactWildcards[1][cid] = delta_tan_y;
        w->contactWildcards[2][cid] = delta_tan_z;
        w->contactWildcards[3][cid] = delta_time;
    }
}

void orc_calc_forces(OrcWorld* w) {
    /* prepareForceArrays, src/kernel/DEMPrepForceKernels.cu:39-44 */
    memset(w->contactForces, 0, sizeof(float) * 3 * w->nContacts);
    memset(w->contactTorque_convToForce, 0, sizeof(float) * 3 * w->nContacts);
    for (uint64_t c = 0; c < w->nContacts; c++)
        calc_one_contact(w, c);
}

/* ---------- forceToAcc, src/kernel/DEMCollectForceKernels_Compact.cu:13-102 (sequential => deterministic order) ---------- */
static inline void add_wrench(OrcWorld* w, uint32_t owner, f3 F, f3 Ftot, f3 cp) {
    unsigned ip = w->inertiaPropOffsets[owner];
    float m = w->MassProperties[ip];
    w->aX[owner] += F.x / m;
    w->aY[owner] += F.y / m;
    w->aZ[owner] += F.z / m;
    f3 myF = Ftot;
    quat_rotate_f(&myF.x, &myF.y, &myF.z, w->oriQw[owner], -w->oriQx[owner], -w->oriQy[owner], -w->oriQz[owner]);
    f3 t = f3_cross(cp, myF);
    w->alphaX[owner] += t.x / w->moiX[ip];
    w->alphaY[owner] += t.y / w->moiY[ip];
    w->alphaZ[owner] += t.z / w->moiZ[ip];
}
void orc_force_to_acc(OrcWorld* w) {
    for (uint64_t c = 0; c < w->nContacts; c++) {
        uint8_t type = w->contactType[c];
        f3 F = f3_make(w->contactForces[3 * c], w->contactForces[3 * c + 1], w->contactForces[3 * c + 2]);
        f3 T = f3_make(w->contactTorque_convToForce[3 * c], w->contactTorque_convToForce[3 * c + 1],
                       w->contactTorque_convToForce[3 * c + 2]);
        f3 cpA = f3_make(w->contactPointGeometryA[3 * c], w->contactPointGeometryA[3 * c + 1],
                         w->contactPointGeometryA[3 * c + 2]);
        f3 cpB = f3_make(w->contactPointGeometryB[3 * c], w->contactPointGeometryB[3 * c + 1],
                         w->contactPointGeometryB[3 * c + 2]);
        uint32_t ownerA = w->ownerClumpBody[w->idGeometryA[c]];
        add_wrench(w, ownerA, F, f3_add(F, T), cpA);
        uint32_t gB = w->idGeometryB[c];
        uint32_t ownerB = (type == ORC_SPHERE_SPHERE) ? w->ownerClumpBody[gB]
                          : (type == ORC_SPHERE_MESH) ? w->ownerMesh[gB]
                                                      : w->objOwner[gB];
        f3 nF = f3_make(-F.x, -F.y, -F.z);
        f3 nFT = f3_scale(f3_add(F, T), -1.f);
        add_wrench(w, ownerB, nF, nFT, cpB);
    }
}

/* ---------- integrateOwners, src/kernel/DEMIntegrationKernels.cu:100-264 ---------- */
void orc_integrate(OrcWorld* w) {
    const float h = w->h;
    for (uint32_t o = 0; o < w->nOwners; o++) {
        unsigned fam = w->familyID[o];
        const OrcPrescription* P = &w->prescriptions[fam];
        int LinVelP[3] = {0, 0, 0}, RotVelP[3] = {0, 0, 0}, LinP[3] = {0, 0, 0}, RotP = 0;
        double X[3];
        f3 old_v = f3_make(w->vX[o], w->vY[o], w->vZ[o]);
        f3 old_omg = f3_make(w->omgBarX[o], w->omgBarY[o], w->omgBarZ[o]);
        orc_voxel_decode(w, o, X);
        X[0] += (double)w->LBF[0]; X[1] += (double)w->LBF[1]; X[2] += (double)w->LBF[2];
        float* vp[3] = {&w->vX[o], &w->vY[o], &w->vZ[o]};
        float* op[3] = {&w->omgBarX[o], &w->omgBarY[o], &w->omgBarZ[o]};
        float* ap[3] = {&w->aX[o], &w->aY[o], &w->aZ[o]};
        float* alp[3] = {&w->alphaX[o], &w->alphaY[o], &w->alphaZ[o]};
        float extra_acc[3] = {0, 0, 0}, extra_angAcc[3] = {0, 0, 0};
        if (P->used) {
            for (int k = 0; k < 3; k++) {
                if (P->hasLinVel[k]) *vp[k] = P->linVel[k];
                if (P->hasRotVel[k]) *op[k] = P->rotVel[k];
                LinVelP[k] = P->linVelPrescribed[k];
                RotVelP[k] = P->rotVelPrescribed[k];
                if (P->hasLinPos[k]) X[k] = P->linPos[k];
                LinP[k] = P->linPosPrescribed[k];
                if (P->hasAcc[k]) extra_acc[k] = P->acc[k];
                if (P->hasAngAcc[k]) extra_angAcc[k] = P->angAcc[k];
            }
            RotP = P->rotPosPrescribed;
        }
        float v_update[3] = {0, 0, 0}, omg_update[3] = {0, 0, 0};
        float oldv[3] = {old_v.x, old_v.y, old_v.z}, oldo[3] = {old_omg.x, old_omg.y, old_omg.z};
        for (int k = 0; k < 3; k++) {
            if (!LinVelP[k]) {
                v_update[k] = (*ap[k] + extra_acc[k] + w->G[k]) * h;
                *vp[k] += v_update[k];
            } else {
                oldv[k] = *vp[k];
            }
        }
        for (int k = 0; k < 3; k++) {
            if (!RotVelP[k]) {
                omg_update[k] = (*alp[k] + extra_angAcc[k]) * h;
                *op[k] += omg_update[k];
            } else {
                oldo[k] = *op[k];
            }
        }
        /* _integrationVelocityPassOnStrategy_ (DEMCustomizablePolicies/IntegrationVelPassOn*.cu):
         * "v = old_v + v_update * 0.5" -> float3 * double resolves to float3*float */
        float v[3], omg[3];
        for (int k = 0; k < 3; k++) {
            if (w->integrator == ORC_EXTENDED_TAYLOR) {
                v[k] = oldv[k] + v_update[k] * 0.5f;
                omg[k] = oldo[k] + omg_update[k] * 0.5f;
            } else if (w->integrator == ORC_CENTERED_DIFFERENCE) {
                v[k] = oldv[k] + v_update[k];
                omg[k] = oldo[k] + omg_update[k];
            } else {
                v[k] = oldv[k];
                omg[k] = oldo[k];
            }
        }
        for (int k = 0; k < 3; k++) {
            if (!LinP[k]) X[k] += (double)v[k] * h;
            X[k] -= (double)w->LBF[k];
        }
        uint16_t loc[3];
        orc_voxel_encode(w, X, &w->voxelID[o], loc);
        w->locX[o] = loc[0]; w->locY[o] = loc[1]; w->locZ[o] = loc[2];
        if (!RotP) {
            /* ha = 0.5 * h * omgBar : (float)(0.5*h) * float3 */
            const float hh = (float)(0.5 * h);
            const float hax = hh * omg[0], hay = hh * omg[1], haz = hh * omg[2];
            /* HamiltonProduct(q, (1,ha)), src/kernel/DEMHelperKernels.cuh:228-245 */
            const float a1 = w->oriQw[o], b1 = w->oriQx[o], c1 = w->oriQy[o], d1 = w->oriQz[o];
            const float a2 = 1.0f, b2 = hax, c2 = hay, d2 = haz;
            float A = a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2;
            float B = a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2;
            float C = a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2;
            float D = a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2;
            /* oriQ /= length(oriQ): float4 dot order x,y,z,w (CUDAMathHelpers.cuh:1035) */
            float len = sqrtf(B * B + C * C + D * D + A * A);
            w->oriQw[o] = A / len; w->oriQx[o] = B / len; w->oriQy[o] = C / len; w->oriQz[o] = D / len;
        }
    }
}

/* ---------- broad phase (acceptance rule of src/kernel/DEMContactKernels_SphereSphere.cu:57-89,172-214 and
 *            src/kernel/DEMBinSphereKernels.cu:78-128) + history map (src/kernel/DEMHistoryMappingKernels.cu) ---------- */
/* Host threads of the broad phase.  Only a build with OpenMP (oracle/_ref/libdemref.so, the timed reference arm) spreads
 * the search; liboracle.so is compiled without it and stays serial.  The candidate SET does not depend on the number of
 * threads, and the list is sorted by (type, A, B) before it is used, so neither does any result. */
void orc_set_threads(int n) { g_orc_threads = n < 1 ? 1 : n; }
#ifdef _OPENMP
#define ORC_CHUNKS (g_orc_threads > 64 ? 64 : g_orc_threads)
#else
#define ORC_CHUNKS 1
#endif

typedef struct { uint32_t a, b; uint8_t type; } CKey;
typedef struct { CKey* keys; size_t n, cap; } CKeyBuf;
static inline void ckey_push(CKeyBuf* kb, uint32_t a, uint32_t b, uint8_t type) {
    if (kb->n == kb->cap) {
        kb->cap = kb->cap ? kb->cap * 2 : 1024;
        kb->keys = (CKey*)realloc(kb->keys, sizeof(CKey) * kb->cap);
    }
    kb->keys[kb->n].a = a; kb->keys[kb->n].b = b; kb->keys[kb->n].type = type;
    kb->n++;
}
static int ckey_cmp(const void* pa, const void* pb) {
    const CKey* x = (const CKey*)pa; const CKey* y = (const CKey*)pb;
    if (x->type != y->type) return x->type < y->type ? -1 : 1;
    if (x->a != y->a) return x->a < y->a ? -1 : 1;
    if (x->b != y->b) return x->b < y->b ? -1 : 1;
    return 0;
}

int orc_detect_contacts(OrcWorld* w) {
    const uint32_t nS = w->nSpheres;
    double* pos = (double*)malloc(sizeof(double) * 3 * (nS ? nS : 1));
    float* rad = (float*)malloc(sizeof(float) * (nS ? nS : 1));
    orc_sphere_positions(w, pos, rad);
    float rmax = 0;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    const int nchunk = ORC_CHUNKS;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nchunk) reduction(max : rmax) reduction(min : lo[:3]) reduction(max : hi[:3])
#endif
    for (uint32_t s = 0; s < nS; s++) {
        rad[s] = rad[s] + w->marginSize[w->ownerClumpBody[s]]; /* float add, fillSharedMemSpheres */
        if (rad[s] > rmax) rmax = rad[s];
        for (int k = 0; k < 3; k++) {
            if (pos[3 * s + k] < lo[k]) lo[k] = pos[3 * s + k];
            if (pos[3 * s + k] > hi[k]) hi[k] = pos[3 * s + k];
        }
    }
    /* every chunk of the sphere range collects into its own buffer (one chunk, i.e. the plain serial search, without OpenMP) */
    CKeyBuf* kb = (CKeyBuf*)calloc((size_t)nchunk + 1, sizeof(CKeyBuf));
    for (int b = 0; b < nchunk; b++) {
        kb[b].cap = 16 + (size_t)nS * 8 / (size_t)nchunk;
        kb[b].keys = (CKey*)malloc(sizeof(CKey) * kb[b].cap);
    }
#define PUSH(A_, B_, T_) ckey_push(mine, (A_), (B_), (T_))

    if (nS > 0) {
        double cs = 2.0 * (double)rmax * 1.0001 + 1e-30;
        long nb[3];
        for (int k = 0; k < 3; k++) {
            nb[k] = (long)((hi[k] - lo[k]) / cs) + 1;
            if (nb[k] < 1) nb[k] = 1;
        }
        /* cap the grid so memory stays bounded */
        while ((double)nb[0] * nb[1] * nb[2] > 6.4e7) {
            cs *= 1.26;
            for (int k = 0; k < 3; k++) nb[k] = (long)((hi[k] - lo[k]) / cs) + 1;
        }
        size_t ncell = (size_t)nb[0] * nb[1] * nb[2];
        uint32_t* cstart = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
        uint32_t* cellOf = (uint32_t*)malloc(sizeof(uint32_t) * nS);
        uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * nS);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nchunk)
#endif
        for (uint32_t s = 0; s < nS; s++) {
            long c[3];
            for (int k = 0; k < 3; k++) {
                c[k] = (long)((pos[3 * s + k] - lo[k]) / cs);
                if (c[k] >= nb[k]) c[k] = nb[k] - 1;
            }
            cellOf[s] = (uint32_t)(c[0] + nb[0] * (c[1] + nb[1] * c[2]));
        }
        for (uint32_t s = 0; s < nS; s++) cstart[cellOf[s] + 1]++;
        for (size_t c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
        uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * ncell);
        memcpy(fill, cstart, sizeof(uint32_t) * ncell);
        for (uint32_t s = 0; s < nS; s++) order[fill[cellOf[s]]++] = s;
        free(fill);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nchunk)
#endif
        for (int ch = 0; ch < nchunk; ch++) {
        CKeyBuf* mine = &kb[ch];
        const uint32_t A0 = (uint32_t)((uint64_t)nS * (uint64_t)ch / (uint64_t)nchunk);
        const uint32_t A1 = (uint32_t)((uint64_t)nS * (uint64_t)(ch + 1) / (uint64_t)nchunk);
        for (uint32_t A = A0; A < A1; A++) {
            long cx = cellOf[A] % nb[0], cy = (cellOf[A] / nb[0]) % nb[1], cz = cellOf[A] / (nb[0] * nb[1]);
            uint32_t oA = w->ownerClumpBody[A];
            unsigned famA = w->familyID[oA];
            for (long dz = -1; dz <= 1; dz++)
                for (long dy = -1; dy <= 1; dy++)
                    for (long dx = -1; dx <= 1; dx++) {
                        long x = cx + dx, y = cy + dy, z = cz + dz;
                        if (x < 0 || y < 0 || z < 0 || x >= nb[0] || y >= nb[1] || z >= nb[2]) continue;
                        size_t cc = (size_t)(x + nb[0] * (y + nb[1] * z));
                        for (uint32_t q = cstart[cc]; q < cstart[cc + 1]; q++) {
                            uint32_t B = order[q];
                            if (B <= A) continue;
                            uint32_t oB = w->ownerClumpBody[B];
                            if (oA == oB) continue;
                            unsigned famB = w->familyID[oB];
                            if (w->familyMasks[mask_pair(famA, famB)] != 0) continue;
                            d3 cp; f3 nn; double depth;
                            int hit = spheres_overlap(pos[3 * A], pos[3 * A + 1], pos[3 * A + 2], rad[A], pos[3 * B],
                                                      pos[3 * B + 1], pos[3 * B + 2], rad[B], &cp, &nn, &depth);
                            float mA = w->familyExtraMarginSize[famA], mB = w->familyExtraMarginSize[famB];
                            float am = (mA < mB) ? mA : mB;
                            if (hit && depth > (double)am) PUSH(A, B, ORC_SPHERE_SPHERE);
                        }
                    }
        }
        }
        free(cstart); free(cellOf); free(order);
    }
    /* sphere--analytical, src/kernel/DEMBinSphereKernels.cu:78-128 */
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nchunk)
#endif
    for (int ch = 0; ch < nchunk; ch++) {
    CKeyBuf* mine = &kb[ch];
    const uint32_t s0 = (uint32_t)((uint64_t)nS * (uint64_t)ch / (uint64_t)nchunk);
    const uint32_t s1 = (uint32_t)((uint64_t)nS * (uint64_t)(ch + 1) / (uint64_t)nchunk);
    for (uint32_t s = s0; s < s1 && w->nAnal > 0; s++) {
        uint32_t oS = w->ownerClumpBody[s];
        unsigned famS = w->familyID[oS];
        unsigned c = w->clumpComponentOffset[s];
        double myRadius = (double)w->Radii[c];
        myRadius += w->marginSize[oS];
        d3 P = {pos[3 * s], pos[3 * s + 1], pos[3 * s + 2]};
        for (uint32_t ob = 0; ob < w->nAnal; ob++) {
            uint32_t oB = w->objOwner[ob];
            unsigned famB = w->familyID[oB];
            if (w->familyMasks[mask_pair(famS, famB)] != 0) continue;
            f3 rel = f3_make(w->objRelPosX[ob], w->objRelPosY[ob], w->objRelPosZ[ob]);
            f3 rot = f3_make(w->objRotX[ob], w->objRotY[ob], w->objRotZ[ob]);
            double op[3];
            orc_voxel_decode(w, oB, op);
            float qw = w->oriQw[oB], qx = w->oriQx[oB], qy = w->oriQy[oB], qz = w->oriQz[oB];
            quat_rotate_f(&rel.x, &rel.y, &rel.z, qw, qx, qy, qz);
            quat_rotate_f(&rot.x, &rot.y, &rot.z, qw, qx, qy, qz);
            d3 Bp = {op[0] + (double)rel.x, op[1] + (double)rel.y, op[2] + (double)rel.z};
            d3 cp; f3 nn; double depth;
            float betaE = w->marginSize[oB];
            /* radA is double here: radA + beta4Entity is a double sum */
            int t = sphere_entity_overlap(P, myRadius + (double)betaE, myRadius, w->objType[ob], Bp, rot,
                                          w->objSize1[ob], w->objNormal[ob], betaE, &cp, &nn, &depth, 0);
            float mA = w->familyExtraMarginSize[famS], mB = w->familyExtraMarginSize[famB];
            double thr = (mA < mB) ? mA : mB;
            if (t && depth > thr) PUSH(s, ob, (uint8_t)t);
        }
    }
    }
    /* sphere--triangle.  The reference sandwiches each facet between two offset copies, bins them and tests the
     * inflated sphere against both copies with a one-sided test (src/kernel/DEMBinTriangleKernels.cu:22-221,
     * src/kernel/DEMContactKernels_SphereTriangle.cu:196-262); which non-touching pairs that admits depends on its bin
     * size.  What the force pass needs is every pair that can come within one radius of the facet before the next
     * rebuild, so the restated rule is geometric: distance(centre, facet) < r + margin(sphere) + margin(mesh), less the
     * smaller family extra margin (same shallow-contact drop as :238-247). */
    CKeyBuf* mine = &kb[nchunk];
    for (uint32_t t = 0; t < w->nTri && nS > 0; t++) {
        uint32_t oT = w->ownerMesh[t];
        unsigned famT = w->familyID[oT];
        double op[3];
        orc_voxel_decode(w, oT, op);
        float qw = w->oriQw[oT], qx = w->oriQx[oT], qy = w->oriQy[oT], qz = w->oriQz[oT];
        d3 nd[3] = {{w->relPosNode1[3 * t], w->relPosNode1[3 * t + 1], w->relPosNode1[3 * t + 2]},
                    {w->relPosNode2[3 * t], w->relPosNode2[3 * t + 1], w->relPosNode2[3 * t + 2]},
                    {w->relPosNode3[3 * t], w->relPosNode3[3 * t + 1], w->relPosNode3[3 * t + 2]}};
        double bl[3] = {1e300, 1e300, 1e300}, bh[3] = {-1e300, -1e300, -1e300};
        for (int k = 0; k < 3; k++) {
            quat_rotate_d(&nd[k].x, &nd[k].y, &nd[k].z, qw, qx, qy, qz);
            nd[k].x += op[0]; nd[k].y += op[1]; nd[k].z += op[2];
            const double c[3] = {nd[k].x, nd[k].y, nd[k].z};
            for (int a = 0; a < 3; a++) {
                if (c[a] < bl[a]) bl[a] = c[a];
                if (c[a] > bh[a]) bh[a] = c[a];
            }
        }
        const double mT = (double)w->marginSize[oT];
        const double grow = (double)rmax + mT;
        for (uint32_t sph = 0; sph < nS; sph++) {
            const double* P = pos + 3 * sph;
            if (P[0] < bl[0] - grow || P[0] > bh[0] + grow || P[1] < bl[1] - grow || P[1] > bh[1] + grow ||
                P[2] < bl[2] - grow || P[2] > bh[2] + grow)
                continue;
            uint32_t oS = w->ownerClumpBody[sph];
            if (oS == oT) continue;
            unsigned famS = w->familyID[oS];
            if (w->familyMasks[mask_pair(famS, famT)] != 0) continue;
            d3 sp = {P[0], P[1], P[2]}, q;
            snap_to_face(nd[0], nd[1], nd[2], sp, &q);
            d3 dd = d3_sub(sp, q);
            const double dist = sqrt(d3_dot(dd, dd));
            float mA = w->familyExtraMarginSize[famS], mB = w->familyExtraMarginSize[famT];
            const double am = (mA < mB) ? mA : mB;
            if (((double)rad[sph] + mT) - dist > am) PUSH(sph, t, ORC_SPHERE_MESH);
        }
    }
#undef PUSH
    free(pos); free(rad);
    /* one list sorted by (type, A, B): every buffer is sorted on its own (in parallel where there are threads), then merged */
    size_t n = 0;
    for (int b = 0; b <= nchunk; b++) n += kb[b].n;
    CKey* keys = (CKey*)malloc(sizeof(CKey) * (n ? n : 1));
    if (nchunk == 1) {
        n = 0;
        for (int b = 0; b <= nchunk; b++) {
            if (kb[b].n) memcpy(keys + n, kb[b].keys, sizeof(CKey) * kb[b].n);
            n += kb[b].n;
        }
        qsort(keys, n, sizeof(CKey), ckey_cmp);  /* (the serial search: one sort of everything, as before) */
    } else {
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nchunk)
#endif
        for (int b = 0; b <= nchunk; b++)
            if (kb[b].n) qsort(kb[b].keys, kb[b].n, sizeof(CKey), ckey_cmp);
        /* chunk b holds spheres below those of chunk b + 1 (and the facet pairs sit in a buffer of their own type), so the
         * merged list is, type after type in ascending order, every buffer's run of that type in buffer order */
        size_t* head = (size_t*)calloc((size_t)nchunk + 1, sizeof(size_t));
        size_t i = 0;
        while (i < n) {
            int t = 256;
            for (int b = 0; b <= nchunk; b++)
                if (head[b] < kb[b].n && (int)kb[b].keys[head[b]].type < t) t = (int)kb[b].keys[head[b]].type;
            for (int b = 0; b <= nchunk; b++) {
                size_t e = head[b];
                while (e < kb[b].n && (int)kb[b].keys[e].type == t) e++;
                if (e > head[b]) memcpy(keys + i, kb[b].keys + head[b], sizeof(CKey) * (e - head[b]));
                i += e - head[b];
                head[b] = e;
            }
        }
        free(head);
    }
    for (int b = 0; b <= nchunk; b++) free(kb[b].keys);
    free(kb);
    if (n > w->contactCapacity) { free(keys); return -1; }

    /* history carry-over: both lists sorted by (type,A,B) => merge */
    const int history = (w->force_model == ORC_HERTZIAN);
    float* nw[4] = {0, 0, 0, 0};
    if (history)
        for (int k = 0; k < 4; k++) nw[k] = (float*)calloc(n ? n : 1, sizeof(float));
    if (history) {
        uint64_t j = 0;
        for (size_t i = 0; i < n; i++) {
            while (j < w->nContacts) {
                CKey ok = {w->idGeometryA[j], w->idGeometryB[j], w->contactType[j]};
                int cmp = ckey_cmp(&ok, &keys[i]);
                if (cmp < 0) { j++; continue; }
                if (cmp == 0)
                    for (int k = 0; k < 4; k++) nw[k][i] = w->contactWildcards[k][j];
                break;
            }
        }
    }
    for (size_t i = 0; i < n; i++) {
        w->idGeometryA[i] = keys[i].a;
        w->idGeometryB[i] = keys[i].b;
        w->contactType[i] = keys[i].type;
        if (history)
            for (int k = 0; k < 4; k++) w->contactWildcards[k][i] = nw[k][i];
    }
    if (history)
        for (int k = 0; k < 4; k++) free(nw[k]);
    w->nContacts = n;
    free(keys);
    return 0;
}

/* ---------- the hot loop (dT workerThread, src/DEM/dT.cpp:2401-2466, with a synchronous rebuild) ---------- */
int orc_step(OrcWorld* w, uint32_t nsteps, uint32_t cd_every, uint64_t* step_counter) {
    if (cd_every < 1) cd_every = 1;
    for (uint32_t s = 0; s < nsteps; s++) {
        if ((*step_counter) % cd_every == 0) {
            orc_compute_margins(w, cd_every);
            int rc = orc_detect_contacts(w);
            if (rc) return rc;
        }
        orc_prepare_acc(w);
        orc_calc_forces(w);
        orc_force_to_acc(w);
        orc_integrate(w);
        w->timeElapsed += (double)w->h;
        (*step_counter)++;
    }
    return 0;
}
