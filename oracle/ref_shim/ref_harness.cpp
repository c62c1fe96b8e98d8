/*
 * ref_harness.cpp -- ORACLE BUILD INFRASTRUCTURE (not product code).
 *
 * Executes the reference's own kernel text (src/kernel/*.cu of /root/reference, placeholder-substituted by
 * gen_ref_src.py into a scratch directory) on the CPU, one CUDA "thread" at a time, over the OrcWorld arrays
 * of oracle/dem_oracle.h.  Built into oracle/_ref/libdemref.so by oracle/Makefile; used by tests to pin the
 * C restatement (dem_oracle.c) and to generate tests/golden/*.npz, and by bench.py as the
 * cpu_baseline kind "reference".
 *
 * Kernels executed through this harness, per thread:
 *   calculateContactForces   src/kernel/DEMCalcForceKernels.cu:44
 *   forceToAcc               src/kernel/DEMCollectForceKernels_Compact.cu:13
 *   prepareAccArrays / prepareForceArrays   src/kernel/DEMPrepForceKernels.cu:32,39
 *   integrateOwners          src/kernel/DEMIntegrationKernels.cu:256
 *   computeMarginFromAbsv / fillMarginValues   src/kernel/DEMMiscKernels.cu:37,63
 *   getNumberOfBinsEachSphereTouches / populateBinSphereTouchingPairs   src/kernel/DEMBinSphereKernels.cu:11,133
 *   makeTriangleSandwich / getNumberOfBinsEachTriangleTouches / populateBinTriangleTouchingPairs
 *                            src/kernel/DEMBinTriangleKernels.cu:22,87,139 (with DEMTriangleBoxIntersect.cu)
 *   buildPersistentMap       src/kernel/DEMHistoryMappingKernels.cu:17
 * block-cooperative (every CUDA thread of a block on its own fiber, launch_coop below):
 *   getNumberOfSphereContactsEachBin / populateSphSphContactPairsEachBin   src/kernel/DEMContactKernels_SphereSphere.cu:91,267
 *   getNumberOfSphTriContactsEachBin / populateTriSphContactsEachBin       src/kernel/DEMContactKernels_SphereTriangle.cu:116,272
 * plus the device functions calcContactPoint (src/kernel/DEMContactKernels_SphereSphere.cu:57), fillSharedMemSpheres /
 * fillSharedMemTriangles (src/kernel/DEMContactKernels_SphereTriangle.cu:16,75), triangle_sphere_CD_directional and
 * snap_to_face (src/kernel/DEMCollisionKernels.cu).
 */
#include "cuda_host_shim.h"

#include <algorithm>
#include <utility>
#include <vector>
#include <cstring>
#include <functional>
#include <ucontext.h>

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
namespace ref_hist {
#include "history.inc"
}
namespace ref_bintri {
#include "bintriangle.inc"
}
namespace ref_cst {
#include "contact_st.inc"
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
    k.ownerMesh = w->ownerMesh;
    k.relPosNode1 = reinterpret_cast<float3*>(w->relPosNode1);
    k.relPosNode2 = reinterpret_cast<float3*>(w->relPosNode2);
    k.relPosNode3 = reinterpret_cast<float3*>(w->relPosNode3);
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

/* A block-cooperative launch: the blocks run one after the other; the `block` CUDA threads of a block are fibers
 * (ucontext) on this host thread, run round-robin from barrier to barrier -- __syncthreads (cuda_host_shim.h) parks the
 * calling fiber, and the next round starts only when every fiber that has not returned is parked, which is the CUDA rule.
 * Deterministic: fiber t always runs before fiber t + 1 between two barriers. */
struct Coop {
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    unsigned cur = 0;
    const std::function<void()>* body = nullptr;
};
thread_local Coop* g_coop = nullptr;
void coop_barrier() {
    Coop* c = g_coop;
    swapcontext(&c->ctx[c->cur], &c->sched);
}
void coop_entry() {
    Coop* c = g_coop;
    (*c->body)();
    c->done[c->cur] = 1;  // (uc_link takes the fiber back to the scheduler)
}
void launch_coop(size_t nBlocks, unsigned block, const std::function<void()>& body) {
    const size_t STACK = 128 * 1024;
    Coop c;
    c.body = &body;
    c.ctx.resize(block);
    c.done.assign(block, 0);
    std::vector<char> stacks(STACK * block);
    g_coop = &c;
    shim_sync_hook = coop_barrier;
    shim_blockDim.x = block;
    for (size_t b = 0; b < nBlocks; b++) {
        shim_blockIdx.x = (unsigned)b;
        for (unsigned t = 0; t < block; t++) {
            getcontext(&c.ctx[t]);
            c.ctx[t].uc_stack.ss_sp = stacks.data() + STACK * t;
            c.ctx[t].uc_stack.ss_size = STACK;
            c.ctx[t].uc_link = &c.sched;
            makecontext(&c.ctx[t], coop_entry, 0);
            c.done[t] = 0;
        }
        unsigned remaining = block;
        while (remaining)
            for (unsigned t = 0; t < block; t++)
                if (!c.done[t]) {
                    c.cur = t;
                    shim_threadIdx.x = t;
                    swapcontext(&c.sched, &c.ctx[t]);
                    if (c.done[t]) remaining--;
                }
    }
    shim_sync_hook = nullptr;
    g_coop = nullptr;
    shim_threadIdx.x = 0;
}

}  // namespace

extern "C" {

/* number of host threads the kernels are spread over (1 = serial and deterministic) */
void ref_set_threads(int n) { g_threads = n < 1 ? 1 : n; orc_set_threads(g_threads); }

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

/* Contact-history map through the reference's own buildPersistentMap (src/kernel/DEMHistoryMappingKernels.cu:17-61; the
 * call site is DEMCubContactDetection.cu:860-960): both lists sorted by idA, per-sphere run lengths and their scans (the
 * reference gets them from cub run-length encode + fillRunLengthArray + prefix scan; restated here), then one thread per
 * sphere looks every new contact up among the sphere's old ones by (idB, type).  mapping[i] = index of new contact i's
 * partner in the old list, or 0xFFFFFFFF for a new contact.  Returns 0, or -1 if a list is not sorted by idA. */
int ref_history_map(uint32_t nSpheres, uint32_t nNew, const uint32_t* newA, const uint32_t* newB, const uint8_t* newT,
                    uint32_t nOld, const uint32_t* oldA, const uint32_t* oldB, const uint8_t* oldT, uint32_t* mapping) {
    for (uint32_t i = 1; i < nNew; i++) if (newA[i] < newA[i - 1]) return -1;
    for (uint32_t i = 1; i < nOld; i++) if (oldA[i] < oldA[i - 1]) return -1;
    std::vector<deme::geoSphereTouches_t> runNew(nSpheres + 1, 0), runOld(nSpheres + 1, 0);
    for (uint32_t i = 0; i < nNew; i++) runNew[newA[i]]++;
    for (uint32_t i = 0; i < nOld; i++) runOld[oldA[i]]++;
    std::vector<deme::contactPairs_t> scanNew(nSpheres + 1, 0), scanOld(nSpheres + 1, 0);
    for (uint32_t s = 0; s < nSpheres; s++) {
        scanNew[s + 1] = scanNew[s] + runNew[s];
        scanOld[s + 1] = scanOld[s] + runOld[s];
    }
    deme::DEMDataKT kt;
    memset((void*)&kt, 0, sizeof(kt));
    std::vector<deme::bodyID_t> nB(newB, newB + nNew), oB(oldB, oldB + nOld);
    std::vector<deme::contact_t> nT(newT, newT + nNew), oT(oldT, oldT + nOld);
    nB.push_back(0); oB.push_back(0); nT.push_back(0); oT.push_back(0);
    kt.idGeometryB = nB.data(); kt.contactType = nT.data();
    kt.previous_idGeometryB = oB.data(); kt.previous_contactType = oT.data();
    std::vector<deme::contactPairs_t> map(nNew + 1, deme::NULL_MAPPING_PARTNER);
    launch(nSpheres, 256, [&] {
        ref_hist::buildPersistentMap(runNew.data(), runOld.data(), scanNew.data(), scanOld.data(), map.data(), &kt, nSpheres);
    });
    for (uint32_t i = 0; i < nNew; i++) mapping[i] = map[i];
    return 0;
}

/* Sphere--sphere contact pairs as the reference finds them (contactDetection(), src/algorithms/
 * DEMCubContactDetection.cu:95-250 and :455-566): sphere -> bin registration through the reference's two per-thread
 * kernels, the glue its CUB calls do (stable sort of the (bin, sphere) pairs by bin, run-length encode, scans) restated
 * with the standard library, then the reference's OWN block-cooperative per-bin sweeps
 * getNumberOfSphereContactsEachBin / populateSphSphContactPairsEachBin (src/kernel/DEMContactKernels_SphereSphere.cu:
 * 91-265, 267-440) with DEME_KT_CD_NTHREADS_PER_BLOCK threads per bin, on fibers (launch_coop).
 * Returns the number of pairs written, or -1 if cap is too small; stats[0] = active bins, stats[1] = most spheres in a
 * bin, stats[2] = contact slots the count pass reserved (>= pairs written). */
long ref_sphere_sphere_contacts(OrcWorld* w, double binSize, uint32_t nbX, uint32_t nbY, uint32_t nbZ, uint32_t* outA,
                                uint32_t* outB, long cap, uint64_t* stats) {
    Bound b; bind(w, b);
    b.sp.binSize = binSize; b.sp.nbX = nbX; b.sp.nbY = nbY; b.sp.nbZ = nbZ;
    b.sp.errOutBinSphNum = 32768;  // the reference's default allowance (SetMaxSphereInBin)
    const uint32_t n = w->nSpheres;
    std::vector<deme::binsSphereTouches_t> nb(n + 1, 0);
    std::vector<deme::objID_t> na(n + 1, 0);
    launch(n, 1024, [&] { ref_bin::getNumberOfBinsEachSphereTouches(&b.sp, &b.kt, nb.data(), na.data()); });
    std::vector<deme::binSphereTouchPairs_t> nbScan(n + 1, 0), naScan(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        nbScan[i + 1] = nbScan[i] + nb[i];
        naScan[i + 1] = naScan[i] + na[i];
    }
    const size_t nPairs = nbScan[n];
    std::vector<deme::binID_t> binIDs(nPairs + 1);
    std::vector<deme::bodyID_t> sphIDs(nPairs + 1);
    std::vector<deme::bodyID_t> idA(naScan[n] + 1), idB(naScan[n] + 1);
    std::vector<deme::contact_t> ct(naScan[n] + 1);
    launch(n, 1024, [&] {
        ref_bin::populateBinSphereTouchingPairs(&b.sp, &b.kt, nbScan.data(), naScan.data(), binIDs.data(),
                                                sphIDs.data(), idA.data(), idB.data(), ct.data());
    });
    // cubDEMSortByKeys (a stable radix sort): by bin, spheres of a bin in the order they were written
    std::vector<size_t> order(nPairs);
    for (size_t i = 0; i < nPairs; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return binIDs[x] < binIDs[y]; });
    std::vector<deme::bodyID_t> sphSorted(nPairs + 1);
    std::vector<deme::binID_t> activeBinIDs;
    std::vector<deme::spheresBinTouches_t> numSpheresBinTouches;
    for (size_t i = 0; i < nPairs; i++) {
        sphSorted[i] = sphIDs[order[i]];
        const deme::binID_t bin = binIDs[order[i]];
        if (activeBinIDs.empty() || activeBinIDs.back() != bin) {  // cubDEMRunLengthEncode
            activeBinIDs.push_back(bin);
            numSpheresBinTouches.push_back(0);
        }
        numSpheresBinTouches.back()++;
    }
    const size_t nActive = activeBinIDs.size();
    std::vector<deme::binSphereTouchPairs_t> lookUp(nActive + 1, 0);  // cubDEMPrefixScan (exclusive)
    uint64_t most = 0;
    for (size_t i = 0; i < nActive; i++) {
        lookUp[i + 1] = lookUp[i] + numSpheresBinTouches[i];
        if (numSpheresBinTouches[i] > most) most = numSpheresBinTouches[i];
    }
    std::vector<deme::binContactPairs_t> numCnt(nActive + 1, 0);
    launch_coop(nActive, DEME_KT_CD_NTHREADS_PER_BLOCK, [&] {
        ref_css::getNumberOfSphereContactsEachBin(&b.sp, &b.kt, sphSorted.data(), activeBinIDs.data(),
                                                  numSpheresBinTouches.data(), lookUp.data(), numCnt.data(), nActive);
    });
    std::vector<deme::contactPairs_t> offsets(nActive + 1, 0);
    for (size_t i = 0; i < nActive; i++) offsets[i + 1] = offsets[i] + numCnt[i];
    const size_t nSlots = offsets[nActive];
    std::vector<deme::bodyID_t> cA(nSlots + 1), cB(nSlots + 1);
    std::vector<deme::contact_t> cT(nSlots + 1, deme::NOT_A_CONTACT);
    launch_coop(nActive, DEME_KT_CD_NTHREADS_PER_BLOCK, [&] {
        ref_css::populateSphSphContactPairsEachBin(&b.sp, &b.kt, sphSorted.data(), activeBinIDs.data(),
                                                   numSpheresBinTouches.data(), lookUp.data(), offsets.data(), cA.data(),
                                                   cB.data(), cT.data(), nActive);
    });
    if (stats) { stats[0] = nActive; stats[1] = most; stats[2] = nSlots; }
    long cnt = 0;
    for (size_t i = 0; i < nSlots; i++) {
        if (cT[i] == deme::NOT_A_CONTACT) continue;
        if (cnt >= cap) return -1;
        outA[cnt] = cA[i]; outB[cnt] = cB[i];
        cnt++;
    }
    return cnt;
}

/* Sphere--triangle pairs through the reference's OWN block-cooperative per-bin kernels getNumberOfSphTriContactsEachBin /
 * populateTriSphContactsEachBin (src/kernel/DEMContactKernels_SphereTriangle.cu:116-270, 272-440) on fibers, behind the same
 * per-thread registration kernels as ref_sphere_tri_contacts below and the glue of contactDetection()
 * (DEMCubContactDetection.cu:300-450: stable sort by bin, run-length encode, scans, hostMergeSearchMapGen).  Same outputs as
 * ref_sphere_tri_contacts; the test holds the two against each other. */
long ref_sphere_tri_contacts_coop(OrcWorld* w, double binSize, uint32_t nbX, uint32_t nbY, uint32_t nbZ, uint32_t* outSphere,
                                  uint32_t* outTri, long cap) {
    Bound b; bind(w, b);
    b.sp.binSize = binSize; b.sp.nbX = nbX; b.sp.nbY = nbY; b.sp.nbZ = nbZ;
    b.sp.errOutBinSphNum = 32768; b.sp.errOutBinTriNum = 32768;
    const uint32_t nT = w->nTri, nS = w->nSpheres;
    if (nT == 0 || nS == 0) return 0;
    std::vector<float3> sA1(nT), sA2(nT), sA3(nT), sB1(nT), sB2(nT), sB3(nT);
    launch(nT, 128, [&] {
        ref_bintri::makeTriangleSandwich(&b.sp, &b.kt, sA1.data(), sA2.data(), sA3.data(), sB1.data(), sB2.data(), sB3.data());
    });
    std::vector<deme::binsTriangleTouches_t> ntb(nT + 1, 0);
    launch(nT, 128, [&] {
        ref_bintri::getNumberOfBinsEachTriangleTouches(&b.sp, &b.kt, ntb.data(), sA1.data(), sA2.data(), sA3.data(), sB1.data(),
                                                       sB2.data(), sB3.data());
    });
    std::vector<deme::binsTriangleTouchPairs_t> tscan(nT + 1, 0);
    for (uint32_t t = 0; t < nT; t++) tscan[t + 1] = tscan[t] + ntb[t];
    std::vector<deme::binID_t> tbin(tscan[nT] + 1);
    std::vector<deme::bodyID_t> ttri(tscan[nT] + 1);
    launch(nT, 128, [&] {
        ref_bintri::populateBinTriangleTouchingPairs(&b.sp, &b.kt, tscan.data(), tbin.data(), ttri.data(), sA1.data(), sA2.data(),
                                                     sA3.data(), sB1.data(), sB2.data(), sB3.data());
    });
    std::vector<deme::binsSphereTouches_t> nsb(nS + 1, 0);
    std::vector<deme::objID_t> na(nS + 1, 0);
    launch(nS, 1024, [&] { ref_bin::getNumberOfBinsEachSphereTouches(&b.sp, &b.kt, nsb.data(), na.data()); });
    std::vector<deme::binSphereTouchPairs_t> sscan(nS + 1, 0), ascan(nS + 1, 0);
    for (uint32_t i = 0; i < nS; i++) {
        sscan[i + 1] = sscan[i] + nsb[i];
        ascan[i + 1] = ascan[i] + na[i];
    }
    std::vector<deme::binID_t> sbin(sscan[nS] + 1);
    std::vector<deme::bodyID_t> ssph(sscan[nS] + 1);
    std::vector<deme::bodyID_t> idA(ascan[nS] + 1), idB(ascan[nS] + 1);
    std::vector<deme::contact_t> ct(ascan[nS] + 1);
    launch(nS, 1024, [&] {
        ref_bin::populateBinSphereTouchingPairs(&b.sp, &b.kt, sscan.data(), ascan.data(), sbin.data(), ssph.data(), idA.data(),
                                                idB.data(), ct.data());
    });
    // stable sort by bin + run-length encode + exclusive scan, for spheres and for facets
    auto group = [](const std::vector<deme::binID_t>& bins, const std::vector<deme::bodyID_t>& ids, size_t n,
                    std::vector<deme::bodyID_t>& sorted, std::vector<deme::binID_t>& active, std::vector<uint32_t>& count,
                    std::vector<uint32_t>& start) {
        std::vector<size_t> order(n);
        for (size_t i = 0; i < n; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return bins[x] < bins[y]; });
        sorted.assign(n + 1, 0);
        for (size_t i = 0; i < n; i++) {
            sorted[i] = ids[order[i]];
            if (active.empty() || active.back() != bins[order[i]]) { active.push_back(bins[order[i]]); count.push_back(0); }
            count.back()++;
        }
        start.assign(active.size() + 1, 0);
        for (size_t i = 0; i < active.size(); i++) start[i + 1] = start[i] + count[i];
    };
    std::vector<deme::bodyID_t> sphSorted, triSorted;
    std::vector<deme::binID_t> actS, actT;
    std::vector<uint32_t> cntS32, cntT32, startS, startT;
    group(sbin, ssph, sscan[nS], sphSorted, actS, cntS32, startS);
    group(tbin, ttri, tscan[nT], triSorted, actT, cntT32, startT);
    std::vector<deme::spheresBinTouches_t> cntS(cntS32.begin(), cntS32.end());
    std::vector<deme::trianglesBinTouches_t> cntT(cntT32.begin(), cntT32.end());
    cntS.push_back(0); cntT.push_back(0);
    std::vector<deme::binSphereTouchPairs_t> lookS(startS.begin(), startS.end());
    std::vector<deme::binsTriangleTouchPairs_t> lookT(startT.begin(), startT.end());
    const size_t nActT = actT.size();
    std::vector<deme::binID_t> map(nActT + 1, deme::NULL_BINID);  // hostMergeSearchMapGen, HostSideHelpers.hpp:177-193
    {
        size_t i2 = 0;
        for (size_t i1 = 0; i1 < nActT; i1++) {
            while (i2 < actS.size() && actS[i2] < actT[i1]) i2++;
            if (i2 < actS.size() && actS[i2] == actT[i1]) map[i1] = (deme::binID_t)i2;
        }
    }
    actS.push_back(deme::NULL_BINID); actT.push_back(deme::NULL_BINID);
    std::vector<deme::binContactPairs_t> numCnt(nActT + 1, 0);
    launch_coop(nActT, DEME_KT_CD_NTHREADS_PER_BLOCK, [&] {
        ref_cst::getNumberOfSphTriContactsEachBin(&b.sp, &b.kt, sphSorted.data(), actS.data(), cntS.data(), lookS.data(), map.data(),
                                                  triSorted.data(), actT.data(), cntT.data(), lookT.data(), numCnt.data(),
                                                  sA1.data(), sA2.data(), sA3.data(), sB1.data(), sB2.data(), sB3.data(), nActT);
    });
    std::vector<deme::contactPairs_t> offsets(nActT + 1, 0);
    for (size_t i = 0; i < nActT; i++) offsets[i + 1] = offsets[i] + numCnt[i];
    const size_t nSlots = offsets[nActT];
    std::vector<deme::bodyID_t> cA(nSlots + 1), cB(nSlots + 1);
    std::vector<deme::contact_t> cT(nSlots + 1, deme::NOT_A_CONTACT);
    launch_coop(nActT, DEME_KT_CD_NTHREADS_PER_BLOCK, [&] {
        ref_cst::populateTriSphContactsEachBin(&b.sp, &b.kt, sphSorted.data(), actS.data(), cntS.data(), lookS.data(), map.data(),
                                               triSorted.data(), actT.data(), cntT.data(), lookT.data(), offsets.data(), cA.data(),
                                               cB.data(), cT.data(), sA1.data(), sA2.data(), sA3.data(), sB1.data(), sB2.data(),
                                               sB3.data(), nActT);
    });
    long cnt = 0;
    for (size_t i = 0; i < nSlots; i++) {
        if (cT[i] == deme::NOT_A_CONTACT) continue;
        if (cnt >= cap) return -1;
        outSphere[cnt] = cA[i]; outTri[cnt] = cB[i];
        cnt++;
    }
    return cnt;
}

/* Sphere--triangle contact candidates as the reference finds them (contactDetection(), src/algorithms/
 * DEMCubContactDetection.cu:262-470): facet sandwich, facet -> bin and sphere -> bin registration through the reference's
 * own per-thread kernels; then, for every bin that holds both, the pair test of getNumberOfSphTriContactsEachBin
 * (src/kernel/DEMContactKernels_SphereTriangle.cu:196-262) restated around the reference's own device functions (fast; the
 * kernel itself runs in ref_sphere_tri_contacts_coop above, and a test holds the two lists against each other).
 * Returns the number of (sphere, triangle) pairs written, or -1 if cap is too small; triBins (nTri entries, may be
 * NULL) receives the number of bins each facet registered in. */
long ref_sphere_tri_contacts(OrcWorld* w, double binSize, uint32_t nbX, uint32_t nbY, uint32_t nbZ, uint32_t* outSphere,
                             uint32_t* outTri, long cap, uint32_t* triBins) {
    Bound b; bind(w, b);
    b.sp.binSize = binSize; b.sp.nbX = nbX; b.sp.nbY = nbY; b.sp.nbZ = nbZ;
    const uint32_t nT = w->nTri, nS = w->nSpheres;
    if (nT == 0 || nS == 0) return 0;
    std::vector<float3> sA1(nT), sA2(nT), sA3(nT), sB1(nT), sB2(nT), sB3(nT);
    launch(nT, 128, [&] {
        ref_bintri::makeTriangleSandwich(&b.sp, &b.kt, sA1.data(), sA2.data(), sA3.data(), sB1.data(), sB2.data(), sB3.data());
    });
    std::vector<deme::binsTriangleTouches_t> ntb(nT + 1, 0);
    launch(nT, 128, [&] {
        ref_bintri::getNumberOfBinsEachTriangleTouches(&b.sp, &b.kt, ntb.data(), sA1.data(), sA2.data(), sA3.data(), sB1.data(),
                                                       sB2.data(), sB3.data());
    });
    std::vector<deme::binsTriangleTouchPairs_t> tscan(nT + 1, 0);
    for (uint32_t t = 0; t < nT; t++) {
        tscan[t + 1] = tscan[t] + ntb[t];
        if (triBins) triBins[t] = ntb[t];
    }
    std::vector<deme::binID_t> tbin(tscan[nT] + 1);
    std::vector<deme::bodyID_t> ttri(tscan[nT] + 1);
    launch(nT, 128, [&] {
        ref_bintri::populateBinTriangleTouchingPairs(&b.sp, &b.kt, tscan.data(), tbin.data(), ttri.data(), sA1.data(), sA2.data(),
                                                     sA3.data(), sB1.data(), sB2.data(), sB3.data());
    });
    /* spheres -> bins */
    std::vector<deme::binsSphereTouches_t> nsb(nS + 1, 0);
    std::vector<deme::objID_t> na(nS + 1, 0);
    launch(nS, 1024, [&] { ref_bin::getNumberOfBinsEachSphereTouches(&b.sp, &b.kt, nsb.data(), na.data()); });
    std::vector<deme::binSphereTouchPairs_t> sscan(nS + 1, 0), ascan(nS + 1, 0);
    for (uint32_t i = 0; i < nS; i++) {
        sscan[i + 1] = sscan[i] + nsb[i];
        ascan[i + 1] = ascan[i] + na[i];
    }
    std::vector<deme::binID_t> sbin(sscan[nS] + 1);
    std::vector<deme::bodyID_t> ssph(sscan[nS] + 1);
    std::vector<deme::bodyID_t> idA(ascan[nS] + 1), idB(ascan[nS] + 1);
    std::vector<deme::contact_t> ct(ascan[nS] + 1);
    launch(nS, 1024, [&] {
        ref_bin::populateBinSphereTouchingPairs(&b.sp, &b.kt, sscan.data(), ascan.data(), sbin.data(), ssph.data(), idA.data(),
                                                idB.data(), ct.data());
    });
    /* group by bin (what the reference does with a radix sort + run-length encode, :340-420) */
    std::vector<std::pair<deme::binID_t, deme::bodyID_t>> tb, sb;
    for (size_t i = 0; i < (size_t)tscan[nT]; i++)
        if (tbin[i] != deme::NULL_BINID) tb.emplace_back(tbin[i], ttri[i]);
    for (size_t i = 0; i < (size_t)sscan[nS]; i++)
        if (sbin[i] != deme::NULL_BINID) sb.emplace_back(sbin[i], ssph[i]);
    std::sort(tb.begin(), tb.end());
    std::sort(sb.begin(), sb.end());
    long cnt = 0;
    size_t it = 0, is = 0;
    while (it < tb.size() && is < sb.size()) {
        if (tb[it].first < sb[is].first) { it++; continue; }
        if (sb[is].first < tb[it].first) { is++; continue; }
        const deme::binID_t binID = tb[it].first;
        size_t te = it, se = is;
        while (te < tb.size() && tb[te].first == binID) te++;
        while (se < sb.size() && sb[se].first == binID) se++;
        for (size_t a = it; a < te; a++) {
            deme::bodyID_t triOwner, triID;
            deme::family_t triFam;
            float3 A1, A2, A3, B1, B2, B3;
            ref_cst::fillSharedMemTriangles(&b.sp, &b.kt, 0, tb[a].second, &triOwner, &triID, &triFam, sA1.data(), sA2.data(),
                                            sA3.data(), sB1.data(), sB2.data(), sB3.data(), &A1, &A2, &A3, &B1, &B2, &B3);
            for (size_t c = is; c < se; c++) {
                deme::bodyID_t sphereID = sb[c].second, ownerID;
                deme::family_t ownerFamily;
                float myRadius;
                float3 sphXYZ;
                ref_cst::fillSharedMemSpheres<float, float>(&b.sp, &b.kt, 0, sphereID, &ownerID, &sphereID, &ownerFamily, &myRadius,
                                                            &sphXYZ.x, &sphXYZ.y, &sphXYZ.z);
                if (ownerID == triOwner) continue;
                unsigned int maskMatID = locateMaskPair<unsigned int>(ownerFamily, triFam);
                if (b.kt.familyMasks[maskMatID] != deme::DONT_PREVENT_CONTACT) continue;
                float artificialMargin = (b.kt.familyExtraMarginSize[ownerFamily] < b.kt.familyExtraMarginSize[triFam])
                                             ? b.kt.familyExtraMarginSize[ownerFamily]
                                             : b.kt.familyExtraMarginSize[triFam];
                float3 cntPnt, normal;
                float depth;
                bool inA = triangle_sphere_CD_directional<float3, float>(A1, A2, A3, sphXYZ, myRadius, normal, depth, cntPnt);
                inA = inA && (-depth > artificialMargin);
                bool inB = triangle_sphere_CD_directional<float3, float>(B1, B2, B3, sphXYZ, myRadius, normal, depth, cntPnt);
                inB = inB && (-depth > artificialMargin);
                if (inA || inB) {
                    snap_to_face(A1, A2, A3, sphXYZ, cntPnt);
                    deme::binID_t contactPntBin = getPointBinID<deme::binID_t>(cntPnt.x, cntPnt.y, cntPnt.z, b.sp.binSize,
                                                                               b.sp.nbX, b.sp.nbY);
                    if (contactPntBin == binID) {
                        if (cnt >= cap) return -1;
                        outSphere[cnt] = sphereID;
                        outTri[cnt] = triID;
                        cnt++;
                    }
                }
            }
        }
        it = te;
        is = se;
    }
    return cnt;
}

/* One sphere against one facet with the reference's narrow-phase functions on the double-precision nodes the force kernel
 * builds (src/kernel/DEMCalcForceKernels.cu:150-181): out[0] = distance from the sphere centre to the facet (snap_to_face),
 * out[1] = the sphere's radius, out[2] = penetration reported by triangle_sphere_CD<double3,double> (0 when not in
 * contact).  Returns 1 when triangle_sphere_CD reports contact. */
int ref_tri_sphere_gap(OrcWorld* w, uint32_t sphereID, uint32_t triID, double out[3]) {
    Bound b; bind(w, b);
    const uint32_t oS = w->ownerClumpBody[sphereID], oT = w->ownerMesh[triID];
    double3 ownS, ownT;
    voxelIDToPosition<double, deme::voxelID_t, deme::subVoxelPos_t>(ownS.x, ownS.y, ownS.z, w->voxelID[oS], w->locX[oS],
                                                                    w->locY[oS], w->locZ[oS], g_nvXp2, g_nvYp2, g_voxelSize, g_l);
    voxelIDToPosition<double, deme::voxelID_t, deme::subVoxelPos_t>(ownT.x, ownT.y, ownT.z, w->voxelID[oT], w->locX[oT],
                                                                    w->locY[oT], w->locZ[oT], g_nvXp2, g_nvYp2, g_voxelSize, g_l);
    const uint32_t comp = w->clumpComponentOffset[sphereID];
    float3 rel = make_float3(CDRelPosX[comp], CDRelPosY[comp], CDRelPosZ[comp]);
    applyOriQToVector3<float, deme::oriQ_t>(rel.x, rel.y, rel.z, w->oriQw[oS], w->oriQx[oS], w->oriQy[oS], w->oriQz[oS]);
    const double3 P = make_double3(ownS.x + (double)rel.x, ownS.y + (double)rel.y, ownS.z + (double)rel.z);
    double3 nd[3];
    const float* src[3] = {w->relPosNode1 + 3 * (size_t)triID, w->relPosNode2 + 3 * (size_t)triID, w->relPosNode3 + 3 * (size_t)triID};
    for (int k = 0; k < 3; k++) {
        nd[k] = make_double3((double)src[k][0], (double)src[k][1], (double)src[k][2]);
        applyOriQToVector3<double, deme::oriQ_t>(nd[k].x, nd[k].y, nd[k].z, w->oriQw[oT], w->oriQx[oT], w->oriQy[oT], w->oriQz[oT]);
        nd[k] = make_double3(nd[k].x + ownT.x, nd[k].y + ownT.y, nd[k].z + ownT.z);
    }
    double3 q;
    snap_to_face<double3, double>(nd[0], nd[1], nd[2], P, q);
    const double dx = P.x - q.x, dy = P.y - q.y, dz = P.z - q.z;
    out[0] = sqrt(dx * dx + dy * dy + dz * dz);
    out[1] = (double)Radii[comp];
    double3 cn, cp;
    double depth = 0.0;
    const bool hit = triangle_sphere_CD<double3, double>(nd[0], nd[1], nd[2], P, (double)Radii[comp], cn, depth, cp);
    out[2] = hit ? depth : 0.0;
    return hit ? 1 : 0;
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

/* The WHOLE contact rebuild through reference kernels (sphere scenes; slow -- fibers -- so for short runs of small scenes):
 * sphere--analytical pairs from the sphere -> bin kernels, sphere--sphere pairs from the per-bin sweeps, the list ordered
 * by (type, A, B) like the oracle's (the order fixes the order of the force accumulation), history words carried over through
 * buildPersistentMap.  Facet pairs are not taken from the reference here: its list legitimately differs from the oracle's
 * (facet broad phase), so a scene with a mesh is refused (-2).  Returns 0, -1 on capacity. */
static double g_full_cd_bin_mult = 0.0;  // > 0: ref_step rebuilds this way, bin size = mult * 2 * (largest radius + margin)
void ref_use_reference_rebuild(double bin_mult) { g_full_cd_bin_mult = bin_mult; }

int ref_detect_contacts_full(OrcWorld* w, double bin_mult) {
    if (w->nTri > 0) return -2;
    const uint32_t nS = w->nSpheres;
    float rmax = 0.f, mmax = 0.f;
    for (uint32_t c = 0; c < w->nComp; c++) rmax = std::max(rmax, w->Radii[c]);
    for (uint32_t o = 0; o < w->nOwners; o++) mmax = std::max(mmax, w->marginSize[o]);
    const double binSize = bin_mult * 2.0 * ((double)rmax + (double)mmax);
    const double ext[3] = {std::ldexp(w->voxelSize, w->nvXp2), std::ldexp(w->voxelSize, w->nvYp2), std::ldexp(w->voxelSize, w->nvZp2)};
    uint32_t nb[3];
    for (int k = 0; k < 3; k++) nb[k] = (uint32_t)std::ceil(ext[k] / binSize);
    struct Key { uint32_t a, b; uint8_t t; };
    std::vector<Key> keys;
    {
        const long cap = 64L * nS + 1024;
        std::vector<uint32_t> A(cap), B(cap);
        std::vector<uint8_t> T(cap);
        long n = ref_sphere_anal_contacts(w, binSize, nb[0], nb[1], nb[2], A.data(), B.data(), T.data(), cap, nullptr);
        if (n < 0) return -1;
        for (long i = 0; i < n; i++) keys.push_back({A[i], B[i], T[i]});
        n = ref_sphere_sphere_contacts(w, binSize, nb[0], nb[1], nb[2], A.data(), B.data(), cap, nullptr);
        if (n < 0) return -1;
        for (long i = 0; i < n; i++) keys.push_back({std::min(A[i], B[i]), std::max(A[i], B[i]), (uint8_t)ORC_SPHERE_SPHERE});
    }
    std::sort(keys.begin(), keys.end(), [](const Key& x, const Key& y) {
        return x.t != y.t ? x.t < y.t : (x.a != y.a ? x.a < y.a : x.b < y.b);
    });
    const uint32_t nNew = (uint32_t)keys.size(), nOld = (uint32_t)w->nContacts;
    if (nNew > w->contactCapacity) return -1;
    // buildPersistentMap wants both lists ordered by idA
    std::vector<uint32_t> pn(nNew), po(nOld);
    for (uint32_t i = 0; i < nNew; i++) pn[i] = i;
    for (uint32_t i = 0; i < nOld; i++) po[i] = i;
    std::stable_sort(pn.begin(), pn.end(), [&](uint32_t x, uint32_t y) { return keys[x].a < keys[y].a; });
    std::stable_sort(po.begin(), po.end(), [&](uint32_t x, uint32_t y) { return w->idGeometryA[x] < w->idGeometryA[y]; });
    std::vector<uint32_t> nA(nNew + 1), nB(nNew + 1), oA(nOld + 1), oB(nOld + 1), map(nNew + 1);
    std::vector<uint8_t> nT(nNew + 1), oT(nOld + 1);
    for (uint32_t i = 0; i < nNew; i++) { nA[i] = keys[pn[i]].a; nB[i] = keys[pn[i]].b; nT[i] = keys[pn[i]].t; }
    for (uint32_t i = 0; i < nOld; i++) { oA[i] = w->idGeometryA[po[i]]; oB[i] = w->idGeometryB[po[i]]; oT[i] = w->contactType[po[i]]; }
    if (ref_history_map(nS, nNew, nA.data(), nB.data(), nT.data(), nOld, oA.data(), oB.data(), oT.data(), map.data())) return -3;
    const bool history = (w->force_model == ORC_HERTZIAN);
    std::vector<float> wc[4];
    for (int k = 0; k < 4; k++) wc[k].assign(nNew + 1, 0.f);
    if (history)
        for (uint32_t i = 0; i < nNew; i++)
            if (map[i] != deme::NULL_MAPPING_PARTNER)
                for (int k = 0; k < 4; k++) wc[k][pn[i]] = w->contactWildcards[k][po[map[i]]];
    for (uint32_t i = 0; i < nNew; i++) {
        w->idGeometryA[i] = keys[i].a; w->idGeometryB[i] = keys[i].b; w->contactType[i] = keys[i].t;
        if (history)
            for (int k = 0; k < 4; k++) w->contactWildcards[k][i] = wc[k][i];
    }
    w->nContacts = nNew;
    return 0;
}

/* hot loop: reference kernels for force/accumulate/integrate, oracle's broad phase for the rebuild (its list is the one the
 * reference's own sweep kernels produce, pair for pair: ref_sphere_sphere_contacts and the test that compares them; the sweeps
 * themselves run on fibers, far too slowly for a loop like this one). */
int ref_step(OrcWorld* w, uint32_t nsteps, uint32_t cd_every, uint64_t* step_counter) {
    if (cd_every < 1) cd_every = 1;
    for (uint32_t s = 0; s < nsteps; s++) {
        if ((*step_counter) % cd_every == 0) {
            ref_compute_margins(w, cd_every);
            int rc = g_full_cd_bin_mult > 0.0 ? ref_detect_contacts_full(w, g_full_cd_bin_mult) : orc_detect_contacts(w);
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
