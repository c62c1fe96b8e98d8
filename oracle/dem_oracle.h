/*
 * dem_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the DEM stepping hot path of projectchrono/DEM-Engine
 * (reference tree: /root/reference, all file:line citations are relative to it).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this.  The product path (dem-engine_b200/csrc) never does.
 *
 * Parity status: PINNED.  Every arithmetic routine here is checked bit-for-bit
 * (tests/test_oracle_vs_ref.py) against oracle/_ref/libdemref.so, which is the
 * reference's own kernel text (src/kernel, .cu files) compiled for the host through the
 * shim in oracle/ref_shim/ (recipe: oracle/Makefile), and against the committed
 * golden vectors in tests/golden/ generated from that library.
 *
 * Data layout follows the reference SoA contract (src/DEM/Defines.h:269-373).
 */
#ifndef DEM_ORACLE_H
#define DEM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* contact type codes, src/DEM/Defines.h:74-82 */
enum {
    ORC_NOT_A_CONTACT = 0,
    ORC_SPHERE_SPHERE = 1,
    ORC_SPHERE_MESH = 2,
    ORC_SPHERE_PLANE = 11,
    ORC_SPHERE_PLATE = 12,
    ORC_SPHERE_CYL = 13
};
/* analytical component types, src/DEM/Defines.h:68-72 */
enum { ORC_ANAL_PLANE = 0, ORC_ANAL_PLATE = 1, ORC_ANAL_CYL_INF = 2 };
/* integrators, src/DEM/Defines.h:146 */
enum { ORC_FORWARD_EULER = 0, ORC_CENTERED_DIFFERENCE = 1, ORC_EXTENDED_TAYLOR = 2 };
/* force models, src/DEM/Defines.h:150 */
enum { ORC_HERTZIAN = 0, ORC_HERTZIAN_FRICTIONLESS = 1 };

#define ORC_NULL_MAPPING 0xFFFFFFFFu /* src/DEM/Defines.h:99 */
#define ORC_NUM_FAMILIES 256
#define ORC_NUM_MASKS 32896 /* (256*257)/2, src/kernel/DEMHelperKernels.cuh:57-62 */

/* Per-family motion prescription with numeric constants only
 * (what equipFamilyPrescribedMotions, src/DEM/APIPrivate.cpp:1601-1708, generates for constant strings). */
typedef struct {
    uint8_t used;
    uint8_t linVelPrescribed[3]; /* LinVel{X,Y,Z}Prescribed */
    uint8_t rotVelPrescribed[3];
    uint8_t linPosPrescribed[3];
    uint8_t rotPosPrescribed;
    uint8_t hasLinVel[3]; /* "vX = <const>" present */
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
} OrcPrescription;

typedef struct {
    /* ---- DEMSimParams subset (src/DEM/Defines.h:194-265) ---- */
    uint32_t nvXp2, nvYp2, nvZp2;
    uint32_t integrator;
    uint32_t force_model;
    uint32_t pad0_;
    double l;
    double voxelSize;
    double timeElapsed;
    float LBF[3];
    float G[3];
    float h;
    float beta; /* fixed expand factor (SetExpandFactor(beta,true)) ; <0 => velocity based margin */
    float approxMaxVel;
    float expSafetyMulti;
    float expSafetyAdder;
    float pad1_;

    /* ---- counts ---- */
    uint32_t nOwners;
    uint32_t nSpheres;
    uint32_t nTri;
    uint32_t nAnal;
    uint32_t nMat;
    uint32_t nComp;
    uint32_t nMassProps;
    uint32_t pad2_;

    /* ---- owners (DEMDataDT) ---- */
    uint64_t* voxelID;
    uint16_t *locX, *locY, *locZ;
    float *oriQw, *oriQx, *oriQy, *oriQz;
    float *vX, *vY, *vZ;
    float *omgBarX, *omgBarY, *omgBarZ;
    float *aX, *aY, *aZ;
    float *alphaX, *alphaY, *alphaZ;
    uint8_t* familyID;
    uint16_t* inertiaPropOffsets;
    uint8_t* accSpecified;
    uint8_t* angAccSpecified;

    /* ---- spheres ---- */
    uint32_t* ownerClumpBody;
    uint16_t* clumpComponentOffset;
    uint16_t* sphereMaterialOffset;

    /* ---- templates (the reference's jitified __constant__ tables) ---- */
    float *Radii, *CDRelPosX, *CDRelPosY, *CDRelPosZ; /* nComp */
    float *MassProperties, *moiX, *moiY, *moiZ;        /* nMassProps */

    /* ---- materials: E,nu per material; CoR,mu,Crr nMat x nMat row-major ---- */
    float *E, *nu, *CoR, *mu, *Crr;

    /* ---- analytical components (nAnal) ---- */
    uint32_t* objOwner;
    uint8_t* objType;
    uint16_t* objMaterial;
    float* objNormal; /* 0 inward / 1 outward, stored as float like the jitified table */
    float *objRelPosX, *objRelPosY, *objRelPosZ;
    float *objRotX, *objRotY, *objRotZ;
    float *objSize1, *objSize2, *objSize3;
    float* objMass;

    /* ---- triangles (nTri), node coordinates in owner frame, xyz interleaved ---- */
    uint32_t* ownerMesh;
    float *relPosNode1, *relPosNode2, *relPosNode3;
    uint16_t* triMaterialOffset;

    /* ---- families ---- */
    uint8_t* familyMasks;         /* ORC_NUM_MASKS */
    float* familyExtraMarginSize; /* 256 */
    OrcPrescription* prescriptions; /* 256 */

    /* ---- contacts ---- */
    uint64_t nContacts;
    uint64_t contactCapacity;
    uint32_t *idGeometryA, *idGeometryB;
    uint8_t* contactType;
    float* contactWildcards[4]; /* delta_tan_x, delta_tan_y, delta_tan_z, delta_time */
    float *contactForces, *contactTorque_convToForce; /* xyz interleaved, 3*capacity */
    float *contactPointGeometryA, *contactPointGeometryB;

    /* ---- scratch owned by the oracle ---- */
    float* marginSize; /* nOwners */
} OrcWorld;

size_t orc_sizeof_world(void);
size_t orc_sizeof_prescription(void);

/* A.1 position codec */
void orc_voxel_decode(const OrcWorld* w, uint32_t owner, double xyz[3]);
void orc_voxel_encode(const OrcWorld* w, const double xyz[3], uint64_t* voxel, uint16_t loc[3]);
void orc_encode_positions(OrcWorld* w, const float* xyz_world, uint32_t first, uint32_t n);
void orc_decode_positions(const OrcWorld* w, float* xyz_world, uint32_t first, uint32_t n);

/* A.8 margin */
void orc_compute_margins(OrcWorld* w, uint32_t maxDrift);

/* broad phase + history: rebuilds w->contacts, carrying wildcards over */
int orc_detect_contacts(OrcWorld* w);
/* host threads of the broad phase (effective only in a build with OpenMP, i.e. oracle/_ref; results do not depend on it) */
void orc_set_threads(int n);

/* A.3-A.5 per-contact force;  A.6 accumulation;  A.7 integration */
void orc_prepare_acc(OrcWorld* w);
void orc_calc_forces(OrcWorld* w);
void orc_force_to_acc(OrcWorld* w);
void orc_integrate(OrcWorld* w);

/* nsteps of the hot loop with a contact rebuild every cd_every steps (>=1).
 * step_count is the number of steps already taken since the last rebuild phase origin. */
int orc_step(OrcWorld* w, uint32_t nsteps, uint32_t cd_every, uint64_t* step_counter);

/* sphere world positions (LBF-relative, as the kT kernels see them) */
void orc_sphere_positions(const OrcWorld* w, double* xyz /*3*nSpheres*/, float* radius);

#ifdef __cplusplus
}
#endif
#endif
