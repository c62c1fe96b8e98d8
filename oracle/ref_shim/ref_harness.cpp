/*
 * ref_harness.cpp -- ORACLE BUILD INFRASTRUCTURE (not product code).
 *
 * Executes the reference's own kernel text (src/kernel/*.cu of /root/reference, placeholder-substituted by
 * gen_ref_src.py into a scratch directory) on the CPU, one CUDA "thread" at a time, over the OrcWorld arrays
 * of oracle/dem_oracle.h.  Built into oracle/_ref/libdemref.so by oracle/Makefile; used by tests to pin the
 * C restatement (dem_oracle.c) and to generate tests/golden/*.npz, and by bench.py as the
 * cpu_baseline kind "reference".
 *
 * Kernels executed through this harness (all strictly per-thread, no intra-block cooperation):
 *   calculateContactForces   src/kernel/DEMCalcForceKernels.cu:44
 *   forceToAcc               src/kernel/DEMCollectForceKernels_Compact.cu:13
 *   prepareAccArrays / prepareForceArrays   src/kernel/DEMPrepForceKernels.cu:32,39
 *   integrateOwners          src/kernel/DEMIntegrationKernels.cu:256
 *   computeMarginFromAbsv / fillMarginValues   src/kernel/DEMMiscKernels.cu:37,63
 *   getNumberOfBinsEachSphereTouches / populateBinSphereTouchingPairs   src/kernel/DEMBinSphereKernels.cu:11,133
 * plus the device function calcContactPoint (src/kernel/DEMContactKernels_SphereSphere.cu:57).
 */
#include "cuda_host_shim.h"

#include <vector>
#include <cstring>

#include "../dem_oracle.h"

/* Pre-include the reference headers at global scope (include guards make the nested includes no-ops). */
#include <DEM/Defines.h>
#include <DEMHelperKernels.cuh>
#undef DEME_ABORT_KERNEL
#define DEME_ABORT_KERNEL(...)                     \
    {                                              \
        fprintf(stderr, __VA_ARGS__);              \
        abort();                                   \
    }
#include <DEMCollisionKernels.cu>

/* ---- runtime-bound stand-ins for the jitified __constant__ tables (ClumpCompDefJitify.cu etc.) ---- */
static unsigned char g_nvXp2, g_nvYp2;
static double g_voxelSize, g_l;
static const OrcPrescription* g_presc;
static const float *Radii, *CDRelPosX, *CDRelPosY, *CDRelPosZ;
static const float *MassProperties, *moiX, *moiY, *moiZ;
static const deme::objType_t* objType;
static const deme::bodyID_t* objOwner;
static const float* objNormal;
static const deme::materialsOffset_t* objMaterial;
static const float *objRelPosX, *objRelPosY, *objRelPosZ, *objRotX, *objRotY, *objRotZ, *objSize1, *objSize2,
    *objSize3, *objMass;
static const float *E, *nu;
struct Mat2 {
    const float* p = nullptr;
    unsigned n = 0;
    const float* operator[](unsigned i) const { return p + (size_t)i * n; }
};
static Mat2 CoR, mu, Crr;

namespace ref_full {
#include "calcforce_full.inc"
}
namespace ref_frictionless {
#include "calcforce_frictionless.inc"
}
namespace ref_collect {
#include "collect_compact.inc"
}
namespace ref_prep {
#include "prepforce.inc"
}
namespace ref_misc {
#include "misc.inc"
}
namespace ref_bin {
#include "binsphere.inc"
}
namespace ref_css {
#include "contact_ss.inc"
}
namespace ref_euler {
#include "integrate_euler.inc"
}
namespace ref_centered {
#include "integrate_centered.inc"
}
namespace ref_taylor {
#include "integrate_taylor.inc"
}

namespace {

struct Bound {
    deme::DEMSimParams sp;
    deme::DEMDataDT dt;
    deme::DEMDataKT kt;
    std::vector<deme::clumpComponentOffset_t> comp8;
};

void bind(OrcWorld* w, Bound& b) {
    g_nvXp2 = (unsigned char)w->nvXp2;
    g_nvYp2 = (unsigned char)w->nvYp2;
    g_voxelSize = w->voxelSize;
    g_l = w->l;
    g_presc = w->prescriptions;
    Radii = w->Radii; CDRelPosX = w->CDRelPosX; CDRelPosY = w->CDRelPosY; CDRelPosZ = w->CDRelPosZ;
    MassProperties = w->MassProperties; moiX = w->moiX; moiY = w->moiY; moiZ = w->moiZ;
    objType = w->objType; objOwner = w->objOwner; objNormal = w->objNormal; objMaterial = w->objMaterial;
    objRelPosX = w->objRelPosX; objRelPosY = w->objRelPosY; objRelPosZ = w->objRelPosZ;
    objRotX = w->objRotX; objRotY = w->objRotY; objRotZ = w->objRotZ;
    objSize1 = w->objSize1; objSize2 = w->objSize2; objSize3 = w->objSize3; objMass = w->objMass;
    E = w->E; nu = w->nu;
    CoR.p = w->CoR; CoR.n = w->nMat;
    mu.p = w->mu; mu.n = w->nMat;
    Crr.p = w->Crr; Crr.n = w->nMat;

    deme::DEMSimParams& s = b.sp;
    memset(&s, 0, sizeof(s));
    s.nvXp2 = w->nvXp2; s.nvYp2 = w->nvYp2; s.nvZp2 = w->nvZp2;
    s.l = w->l; s.voxelSize = w->voxelSize;
    s.nSpheresGM = w->nSpheres; s.nTriGM = w->nTri; s.nAnalGM = (deme::objID_t)w->nAnal;
    s.nOwnerBodies = w->nOwners;
    s.LBFX = w->LBF[0]; s.LBFY = w->LBF[1]; s.LBFZ = w->LBF[2];
    s.Gx = w->G[0]; s.Gy = w->G[1]; s.Gz = w->G[2];
    s.h = w->h; s.timeElapsed = w->timeElapsed;
    s.beta = w->beta; s.approxMaxVel = w->approxMaxVel;
    s.expSafetyMulti = w->expSafetyMulti; s.expSafetyAdder = w->expSafetyAdder;
    s.errOutBinSphNum = 32768; s.errOutBinTriNum = 32768;

    b.comp8.resize(w->nSpheres);
    for (uint32_t i = 0; i < w->nSpheres; i++) b.comp8[i] = (deme::clumpComponentOffset_t)w->clumpComponentOffset[i];

    deme::DEMDataDT& d = b.dt;
    memset((void*)&d, 0, sizeof(d));
    d.inertiaPropOffsets = w->inertiaPropOffsets; d.familyID = w->familyID; d.voxelID = w->voxelID;
    d.locX = w->locX; d.locY = w->locY; d.locZ = w->locZ;
    d.oriQw = w->oriQw; d.oriQx = w->oriQx; d.oriQy = w->oriQy; d.oriQz = w->oriQz;
    d.vX = w->vX; d.vY = w->vY; d.vZ = w->vZ;
    d.omgBarX = w->omgBarX; d.omgBarY = w->omgBarY; d.omgBarZ = w->omgBarZ;
    d.aX = w->aX; d.aY = w->aY; d.aZ = w->aZ;
    d.alphaX = w->alphaX; d.alphaY = w->alphaY; d.alphaZ = w->alphaZ;
    d.accSpecified = w->accSpecified; d.angAccSpecified = w->angAccSpecified;
    d.idGeometryA = w->idGeometryA; d.idGeometryB = w->idGeometryB; d.contactType = w->contactType;
    d.familyMasks = w->familyMasks; d.familyExtraMarginSize = w->familyExtraMarginSize;
    d.contactForces = reinterpret_cast<float3*>(w->contactForces);
    d.contactTorque_convToForce = reinterpret_cast<float3*>(w->contactTorque_convToForce);
    d.contactPointGeometryA = reinterpret_cast<float3*>(w->contactPointGeometryA);
    d.contactPointGeometryB = reinterpret_cast<float3*>(w->contactPointGeometryB);
    d.ownerClumpBody = w->ownerClumpBody; d.clumpComponentOffset = b.comp8.data();
    d.sphereMaterialOffset = w->sphereMaterialOffset;
    d.ownerMesh = w->ownerMesh;
    d.relPosNode1 = reinterpret_cast<float3*>(w->relPosNode1);
    d.relPosNode2 = reinterpret_cast<float3*>(w->relPosNode2);
    d.relPosNode3 = reinterpret_cast<float3*>(w->relPosNode3);
    d.triMaterialOffset = w->triMaterialOffset;
    for (int k = 0; k < 4; k++) d.contactWildcards[k] = w->contactWildcards[k];

    deme::DEMDataKT& k = b.kt;
    memset((void*)&k, 0, sizeof(k));
    k.familyID = w->familyID; k.voxelID = w->voxelID; k.locX = w->locX; k.locY = w->locY; k.locZ = w->locZ;
    k.oriQw = w->oriQw; k.oriQx = w->oriQx; k.oriQy = w->oriQy; k.oriQz = w->oriQz;
    k.marginSize = w->marginSize; k.familyMasks = w->familyMasks;
    k.familyExtraMarginSize = w->familyExtraMarginSize;
    k.ownerClumpBody = w->ownerClumpBody; k.clumpComponentOffset = b.comp8.data();
}

int g_threads = 1;

/* one CUDA "thread" per index; indices are spread over g_threads host threads (OpenMP) */
template <typename F>
void launch(size_t n, unsigned block, F&& body) {
    if (g_threads <= 1) {
        shim_blockDim.x = block;
        for (size_t i = 0; i < n; i++) {
            shim_blockIdx.x = (unsigned)(i / block);
            shim_threadIdx.x = (unsigned)(i % block);
            body();
        }
        return;
    }
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (long long i = 0; i < (long long)n; i++) {
        shim_blockDim.x = block;
        shim_blockIdx.x = (unsigned)(i / block);
        shim_threadIdx.x = (unsigned)(i % block);
        body();
    }
}

}  // namespace

extern "C" {

/* number of host threads the kernels are spread over (1 = serial and deterministic) */
void ref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

void ref_prepare_acc(OrcWorld* w) {
    Bound b; bind(w, b);
    launch(w->nOwners, 1024, [&] { ref_prep::prepareAccArrays(&b.sp, &b.dt); });
}

void ref_calc_forces(OrcWorld* w) {
    Bound b; bind(w, b);
    size_t n = w->nContacts;
    launch(n, 1024, [&] { ref_prep::prepareForceArrays(&b.sp, &b.dt, n); });
    if (w->force_model == ORC_HERTZIAN)
        launch(n, 256, [&] { ref_full::calculateContactForces(&b.sp, &b.dt, n); });
    else
        launch(n, 256, [&] { ref_frictionless::calculateContactForces(&b.sp, &b.dt, n); });
}

void ref_force_to_acc(OrcWorld* w) {
    Bound b; bind(w, b);
    size_t n = w->nContacts;
    launch(n, 1024, [&] { ref_collect::forceToAcc(&b.dt, n); });
}

void ref_integrate(OrcWorld* w) {
    Bound b; bind(w, b);
    if (w->integrator == ORC_EXTENDED_TAYLOR)
        launch(w->nOwners, 1024, [&] { ref_taylor::integrateOwners(&b.sp, &b.dt); });
    else if (w->integrator == ORC_CENTERED_DIFFERENCE)
        launch(w->nOwners, 1024, [&] { ref_centered::integrateOwners(&b.sp, &b.dt); });
    else
        launch(w->nOwners, 1024, [&] { ref_euler::integrateOwners(&b.sp, &b.dt); });
}

/* kT.cpp:125-168: marginSize first receives |v| (the absv inspector output), then the margin kernel runs */
void ref_compute_margins(OrcWorld* w, uint32_t maxDrift) {
    Bound b; bind(w, b);
    if (w->beta >= 0.f) {
        launch(w->nOwners, 1024, [&] { ref_misc::fillMarginValues(&b.sp, &b.kt, w->nOwners); });
        return;
    }
    for (uint32_t o = 0; o < w->nOwners; o++) {
        /* inspectOwnerProperty with the absv code: sqrt(vX^2+vY^2+vZ^2) in float (AuxClasses.cpp:54-61) */
        float vx = w->vX[o], vy = w->vY[o], vz = w->vZ[o];
        w->marginSize[o] = sqrt(vx * vx + vy * vy + vz * vz);
    }
    float ts = w->h;
    unsigned int md = maxDrift;
    launch(w->nOwners, 1024, [&] { ref_misc::computeMarginFromAbsv(&b.sp, &b.kt, &ts, &md, w->nOwners); });
}

/* Sphere--analytical contacts through the reference's own two-pass bin kernels.
 * Returns the number of contacts written (sphere, obj, type triples), or -1 if cap is too small.
 * Also returns in nBinsTouched[] (nSpheres entries, may be NULL) the per-sphere bin-touch counts for the
 * given bin size / grid (nbX,nbY,nbZ). */
long ref_sphere_anal_contacts(OrcWorld* w, double binSize, uint32_t nbX, uint32_t nbY, uint32_t nbZ,
                              uint32_t* outSphere, uint32_t* outObj, uint8_t* outType, long cap,
                              uint16_t* nBinsTouched) {
    Bound b; bind(w, b);
    b.sp.binSize = binSize; b.sp.nbX = nbX; b.sp.nbY = nbY; b.sp.nbZ = nbZ;
    const uint32_t n = w->nSpheres;
    std::vector<deme::binsSphereTouches_t> nb(n + 1, 0);
    std::vector<deme::objID_t> na(n + 1, 0);
    launch(n, 1024, [&] { ref_bin::getNumberOfBinsEachSphereTouches(&b.sp, &b.kt, nb.data(), na.data()); });
    std::vector<deme::binSphereTouchPairs_t> nbScan(n + 1, 0), naScan(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        nbScan[i + 1] = nbScan[i] + nb[i];
        naScan[i + 1] = naScan[i] + na[i];
        if (nBinsTouched) nBinsTouched[i] = nb[i];
    }
    std::vector<deme::binID_t> binIDs(nbScan[n] + 1);
    std::vector<deme::bodyID_t> sphIDs(nbScan[n] + 1);
    std::vector<deme::bodyID_t> idA(naScan[n] + 1), idB(naScan[n] + 1);
    std::vector<deme::contact_t> ct(naScan[n] + 1);
    launch(n, 1024, [&] {
        ref_bin::populateBinSphereTouchingPairs(&b.sp, &b.kt, nbScan.data(), naScan.data(), binIDs.data(),
                                                sphIDs.data(), idA.data(), idB.data(), ct.data());
    });
    long cnt = naScan[n];
    if (cnt > cap) return -1;
    for (long i = 0; i < cnt; i++) {
        outSphere[i] = idA[i]; outObj[i] = idB[i]; outType[i] = ct[i];
    }
    return cnt;
}

/* calcContactPoint (src/kernel/DEMContactKernels_SphereSphere.cu:57-89) on explicit inputs */
int ref_calc_contact_point(double binSize, uint32_t nbX, uint32_t nbY, const double A[3], float rA, const double B[3],
                           float rB, float marginA, float marginB, uint32_t* binID) {
    deme::DEMSimParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.binSize = binSize; sp.nbX = nbX; sp.nbY = nbY;
    deme::binID_t bin;
    bool hit = ref_css::calcContactPoint(&sp, A[0], A[1], A[2], rA, B[0], B[1], B[2], rB, bin, marginA, marginB);
    *binID = bin;
    return hit ? 1 : 0;
}

/* position codec straight from the reference templates (src/kernel/DEMHelperKernels.cuh:117-159) */
void ref_voxel_decode(OrcWorld* w, uint32_t o, double xyz[3]) {
    voxelIDToPosition<double, deme::voxelID_t, deme::subVoxelPos_t>(
        xyz[0], xyz[1], xyz[2], w->voxelID[o], w->locX[o], w->locY[o], w->locZ[o], (unsigned char)w->nvXp2,
        (unsigned char)w->nvYp2, w->voxelSize, w->l);
}
void ref_voxel_encode(OrcWorld* w, const double xyz[3], uint64_t* voxel, uint16_t loc[3]) {
    deme::voxelID_t id;
    positionToVoxelID<deme::voxelID_t, deme::subVoxelPos_t, double>(id, loc[0], loc[1], loc[2], xyz[0], xyz[1],
                                                                    xyz[2], (unsigned char)w->nvXp2,
                                                                    (unsigned char)w->nvYp2, w->voxelSize, w->l);
    *voxel = id;
}

/* hot loop: reference kernels for force/accumulate/integrate, oracle's broad phase for the rebuild
 * (the reference's sweep kernels are block-cooperative and cannot run through the shim). */
int ref_step(OrcWorld* w, uint32_t nsteps, uint32_t cd_every, uint64_t* step_counter) {
    if (cd_every < 1) cd_every = 1;
    for (uint32_t s = 0; s < nsteps; s++) {
        if ((*step_counter) % cd_every == 0) {
            ref_compute_margins(w, cd_every);
            int rc = orc_detect_contacts(w);
            if (rc) return rc;
        }
        ref_prepare_acc(w);
        ref_calc_forces(w);
        ref_force_to_acc(w);
        ref_integrate(w);
        w->timeElapsed += (double)w->h;
        (*step_counter)++;
    }
    return 0;
}

}  // extern "C"
