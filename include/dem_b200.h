/*
 * dem_b200.h -- C ABI of the B200-native DEM stepping core (libdemcore.so).
 *
 * This is the drop-in boundary for the hot path of projectchrono/DEM-Engine
 * (kinematic-thread contact detection -> dynamic-thread contact force -> owner integration).
 * In the reference that path is not behind a plugin ABI: it lives behind the C++ class
 * deme::DEMSolver (src/DEM/API.h:50) whose Initialize()/DoDynamics() drive two worker
 * threads (src/DEM/dT.cpp:2324, src/DEM/kT.cpp:218).  Each entry point below names the
 * reference interface it replaces (file:line, relative to the reference tree).
 *
 * Conventions: every function returns 0 on success and a negative DEM_ERR_* code on failure
 * (no exceptions cross the boundary); dem_last_error() returns a human readable message.
 * All pointers are HOST memory owned by the caller; arrays follow the reference's SoA
 * contract (src/DEM/Defines.h:269-373).  One context is driven by one CPU thread at a time.
 * No torch / C++ types appear in any signature.
 */
#ifndef DEM_B200_H
#define DEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEM_B200_ABI_VERSION 1

enum {
    DEM_OK = 0,
    DEM_ERR_INVALID = -1,  /* bad argument / call order */
    DEM_ERR_CUDA = -2,     /* CUDA runtime failure (message has file:line) */
    DEM_ERR_CAPACITY = -3, /* a device array overflowed and could not be grown */
    DEM_ERR_VELOCITY = -4, /* max velocity exceeded errOutVel / non-finite state (kT.cpp:136-149) */
    DEM_ERR_NO_GPU = -5    /* no CUDA device: there is NO CPU fallback */
};

/* contact type codes, src/DEM/Defines.h:74-82 */
enum { DEM_CNT_NONE = 0, DEM_CNT_SPHERE_SPHERE = 1, DEM_CNT_SPHERE_MESH = 2, DEM_CNT_SPHERE_PLANE = 11,
       DEM_CNT_SPHERE_PLATE = 12, DEM_CNT_SPHERE_CYL = 13 };
/* analytical component types, src/DEM/Defines.h:68-72 */
enum { DEM_ANAL_PLANE = 0, DEM_ANAL_PLATE = 1, DEM_ANAL_CYL_INF = 2 };
/* TIME_INTEGRATOR, src/DEM/Defines.h:146 */
enum { DEM_FORWARD_EULER = 0, DEM_CENTERED_DIFFERENCE = 1, DEM_EXTENDED_TAYLOR = 2 };
/* FORCE_MODEL, src/DEM/Defines.h:150 */
enum { DEM_HERTZIAN = 0, DEM_HERTZIAN_FRICTIONLESS = 1 };

#define DEM_NUM_FAMILIES 256
#define DEM_NUM_FAMILY_MASKS 32896 /* upper-triangular 256x256, src/kernel/DEMHelperKernels.cuh:57-62 */

typedef struct DemCtx DemCtx;

/* Replaces deme::DEMSimParams (src/DEM/Defines.h:194-265) + the SolverFlags that steer the hot path
 * (src/DEM/Structs.h:482-531). */
typedef struct DemSimParams {
    uint32_t nvXp2, nvYp2, nvZp2; /* voxel-count bits per axis (figureOutNV, APIPrivate.cpp:373-487) */
    uint32_t integrator;          /* DEM_FORWARD_EULER / CENTERED_DIFFERENCE / EXTENDED_TAYLOR */
    uint32_t force_model;         /* DEM_HERTZIAN / DEM_HERTZIAN_FRICTIONLESS */
    uint32_t cd_update_freq;      /* steps between contact-list rebuilds (SetCDUpdateFreq, API.h:107-113); >=1 */
    double l;                     /* smallest length unit */
    double voxelSize;             /* 2^16 * l */
    float LBF[3];                 /* left-bottom-front corner of the world */
    float G[3];                   /* gravitational acceleration */
    float userBoxMin[3];
    float userBoxMax[3];
    float h;                      /* step size (float-rounded, Defines.h:240) */
    float beta;                   /* >=0: fixed expand factor (SetExpandFactor(beta,true)); <0: velocity based margin */
    float approxMaxVel;           /* SetMaxVelocity */
    float expSafetyMulti;         /* SetExpandSafetyMultiplier */
    float expSafetyAdder;         /* SetExpandSafetyAdder */
    float errOutVel;              /* SetErrorOutVelocity */
    uint32_t record_contact_forces; /* 0 == SetNoForceRecord() */
    uint32_t pad_;
} DemSimParams;

/* Per-family motion prescription with numeric constants (SetFamilyFixed / SetFamilyPrescribedLinVel / AngVel /
 * Position / AddFamilyPrescribedAcc with constant strings; what equipFamilyPrescribedMotions,
 * APIPrivate.cpp:1601-1708, compiles into the integration kernel). */
typedef struct DemPrescription {
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
} DemPrescription;

/* Replaces the timers/collaboration stats of ShowTimingStats / ShowThreadCollaborationStats
 * (APIPublic.cpp:2215, :2481-2505) and GetNumContacts / GetAvgSphContacts. */
typedef struct DemStats {
    uint64_t n_steps;          /* steps taken since creation */
    uint64_t n_rebuilds;       /* contact-list rebuilds */
    uint64_t n_contacts_ss;    /* sphere-sphere candidates in the current list */
    uint64_t n_contacts_sa;    /* sphere-analytical */
    uint64_t n_contacts_st;    /* sphere-triangle */
    uint64_t contact_capacity; /* per-list device capacity */
    uint64_t kernel_launches;  /* kernels (own) launched or replayed from graphs since creation */
    uint64_t device_bytes;     /* device memory held */
    double sim_time;           /* GetSimTime */
    float max_margin;          /* largest contact margin of the last rebuild */
    float cell_size;           /* broad-phase cell edge of the last rebuild */
    uint32_t n_cells[3];
    uint32_t overflow;         /* !=0: a list overflowed (capacity was grown and the rebuild redone) */
    uint32_t cd_update_freq;   /* steps per contact-list cycle right now (GetUpdateFreq; moves when "adaptive_update_freq" is on) */
    uint64_t n_contacts_ss_touching; /* sphere-sphere pairs that overlapped at the last rebuild */
} DemStats;

/* ---- host-side set-up arithmetic ---------------------------------------------------------------------- */
/* DEMSolver::figureOutNV (APIPrivate.cpp:373-487): split 64 voxel bits, choose l. box_* is the target (enlarged) box. */
int dem_host_figure_out_nv(const float box_min[3], const float box_max[3], uint32_t nv_p2[3], double* l,
                           double* voxel_size);
/* ... with the length along one axis made exact (InstructBoxDomainDimension(..., dir_exact), APIPublic.cpp:855-868 and
 * APIPrivate.cpp:442-476): exact_dir = 0 / 1 / 2 for X / Y / Z, -1 for none.  The caller passes a target box that is
 * not enlarged along that axis. */
int dem_host_figure_out_nv_exact(const float box_min[3], const float box_max[3], int exact_dir, uint32_t nv_p2[3],
                                 double* l, double* voxel_size);
/* InstructBoxDomainDimension(x,y,z) (APIPublic.cpp:845-872). Outputs user box and 20%-enlarged target box. */
int dem_host_box_domain(float x, float y, float z, float user_min[3], float user_max[3], float target_min[3],
                        float target_max[3]);
/* positionToVoxelID on the host (DEMHelperKernels.cuh:137-159), as dT::populateEntityArrays uses it (dT.cpp:638-1024). */
int dem_host_encode_positions(const DemSimParams* p, const float* xyz_world, uint64_t n, uint64_t* voxelID,
                              uint16_t* locX, uint16_t* locY, uint16_t* locZ);

/* ---- life cycle: DEMSolver::DEMSolver / ~DEMSolver (APIPublic.cpp:23-92) ------------------------------- */
int dem_ctx_create(DemCtx** out, int device);
/* Number of CUDA devices (0 without a driver / device). */
int dem_device_count(void);
/* DEMSolver(nGPUs) / DEMSolver(std::vector<int>) (src/DEM/API.h:53,56; the reference caps nGPUs at 2, one for each of its
 * worker threads -- here up to the 8 GPUs of a box): ONE context that drives n devices.  Every entry point called on it
 * acts on the whole group: set-up calls reach every device, dem_initialize shards the scene into x-slabs (see the
 * multi-GPU section below) when it qualifies -- at least "group_min_owners" (dem_set_option, default 50 000) clump owners
 * per GPU and no free-moving wall / mesh owner; otherwise the scene runs on devices[0] alone -- stepping calls step all
 * ranks together, and read-back calls first merge the ranks' owners into the first device (peer copies).  devices may
 * be NULL (0 .. n-1). */
int dem_ctx_create_group(DemCtx** out, const int* devices, int n);
int dem_ctx_destroy(DemCtx* ctx);
const char* dem_last_error(const DemCtx* ctx);
int dem_abi_version(void);
/* run on a caller-provided cudaStream_t (e.g. a torch stream) instead of the context's own stream */
int dem_set_stream(DemCtx* ctx, void* cuda_stream);

/* ---- Initialize(): flattened user input (APIPublic.cpp:2161-2213; dT::populateEntityArrays dT.cpp:638-1024) ---- */
int dem_set_params(DemCtx* ctx, const DemSimParams* p); /* setSimParams, APIPrivate.cpp:1121 */
/* clump component table + mass-property table (equipClumpTemplates / equipMassMoiVolume, APIPrivate.cpp:2028,1795) */
int dem_upload_templates(DemCtx* ctx, uint32_t nComp, const float* radii, const float* relX, const float* relY,
                         const float* relZ, uint32_t nMassProps, const float* mass, const float* moiX,
                         const float* moiY, const float* moiZ);
/* equipMaterials (APIPrivate.cpp:1877-2026): E,nu per material; CoR,mu,Crr nMat x nMat row-major */
int dem_upload_materials(DemCtx* ctx, uint32_t nMat, const float* E, const float* nu, const float* CoR,
                         const float* mu, const float* Crr);
/* equipAnalGeoTemplates (APIPrivate.cpp:1724-1793) */
int dem_upload_analytical(DemCtx* ctx, uint32_t nAnal, const uint32_t* objOwner, const uint8_t* objType,
                          const uint16_t* objMaterial, const float* objNormal, const float* relPosX,
                          const float* relPosY, const float* relPosZ, const float* rotX, const float* rotY,
                          const float* rotZ, const float* size1, const float* size2, const float* size3,
                          const float* objMass);
/* family mask matrix (APIPrivate.cpp:815), extra margins, prescriptions (256 entries) */
int dem_upload_families(DemCtx* ctx, const uint8_t* masks, const float* extraMargin, const DemPrescription* presc);
/* owners: clumps, then analytical objects, then meshes (dT.cpp:638-1024) */
int dem_upload_owners(DemCtx* ctx, uint32_t nOwners, const uint64_t* voxelID, const uint16_t* locX,
                      const uint16_t* locY, const uint16_t* locZ, const float* oriQw, const float* oriQx,
                      const float* oriQy, const float* oriQz, const float* vX, const float* vY, const float* vZ,
                      const float* omgBarX, const float* omgBarY, const float* omgBarZ, const uint8_t* familyID,
                      const uint16_t* inertiaPropOffsets);
int dem_upload_spheres(DemCtx* ctx, uint32_t nSpheres, const uint32_t* ownerClumpBody,
                       const uint16_t* clumpComponentOffset, const uint16_t* sphereMaterialOffset);
/* triangles of mesh owners (preprocessTriangleObjs; nodes in owner frame, xyz interleaved) */
int dem_upload_triangles(DemCtx* ctx, uint32_t nTri, const uint32_t* ownerMesh, const float* node1,
                         const float* node2, const float* node3, const uint16_t* triMaterialOffset);
/* Deforming mesh: new owner-frame node positions of the facets [first, first+n) (3 floats per node, one array per
 * facet corner). Replaces SetTriNodeRelPos / UpdateTriNodeRelPos (src/DEM/API.h:489-491, dT.cpp:3135-3158), what
 * DEMTracker::UpdateMesh / UpdateMeshByIncrement call (AuxClasses.cpp:681-693). Stream-ordered; the contact list is
 * rebuilt before the next step. */
int dem_update_triangle_nodes(DemCtx* ctx, uint32_t first, uint32_t n, const float* node1, const float* node2,
                              const float* node3);

/* allocateGPUArrays + initGPUArrays (dT.cpp:409, kT.cpp:579). contact_capacity==0 -> automatic */
int dem_initialize(DemCtx* ctx, uint64_t contact_capacity);
/* restart: SetExistingContacts / SetExistingContactWildcards (Structs.h:857-882, dT.cpp:849-881) */
int dem_set_contacts(DemCtx* ctx, uint64_t n, const uint32_t* idA, const uint32_t* idB, const uint8_t* type,
                     const float* wildcards4 /* n x 4 row-major: delta_tan_xyz, delta_time */);

/* ---- the hot loop ------------------------------------------------------------------------------------- */
/* DoDynamics(t) / DoDynamicsThenSync(t) (APIPublic.cpp:2446-2479): takes the same number of steps as the reference's
 * loop "for (double cycle = 0; cycle < t; cycle += (double)h)" (dT.cpp:2401) and blocks until they are done.
 * t <= 0 only (re)builds the contact list (the dry run of DoDynamicsThenSync(0), dT.cpp:2393-2398). */
int dem_do_dynamics(DemCtx* ctx, double t);
/* DoStepDynamics() x n (API.h:1252-1263). Blocking. */
int dem_step(DemCtx* ctx, uint64_t n_steps);
/* enqueue n steps on the stream without waiting (dem_sync() later); lets callers time with their own events.  Contact-
 * list rebuilds are enqueued like steps: their verdict (counts, overflow, too-fast owners) is read one cycle later or at
 * dem_sync, which is where DEM_ERR_VELOCITY / DEM_ERR_CAPACITY surface.  A rebuild that overflowed freezes the state on
 * the device; the host grows the lists and replays from that rebuild on, so the result is the same as if nothing had
 * happened (family tables uploaded in between are replayed with their latest values). */
int dem_step_async(DemCtx* ctx, uint64_t n_steps);
int dem_sync(DemCtx* ctx);
/* force a contact-list rebuild now (kT contactDetection(), DEMCubContactDetection.cu:38-1123) */
int dem_rebuild_contacts(DemCtx* ctx);
/* UpdateStepSize (API.h) */
/* SetSimTime (API.h:117, dT.cpp:2709-2713): the simulated-time clock reported by dem_get_stats */
int dem_set_sim_time(DemCtx* ctx, double t);
int dem_update_step_size(DemCtx* ctx, float h);

/* ---- state access: trackers / writers read through these (dT.cpp:3062-3130) ---------------------------- */
/* any output pointer may be NULL */
int dem_download_owner_state(DemCtx* ctx, uint32_t first, uint32_t n, uint64_t* voxelID, uint16_t* locX,
                             uint16_t* locY, uint16_t* locZ, float* oriQ_wxyz, float* vel_xyz, float* omg_xyz,
                             float* acc_xyz, float* angacc_xyz, uint8_t* family);
/* decoded world positions as the reference reports them (float, LBF added; dT.cpp:3062-3076) or in double */
int dem_download_positions(DemCtx* ctx, uint32_t first, uint32_t n, float* xyz_f32, double* xyz_f64);
int dem_upload_owner_state(DemCtx* ctx, uint32_t first, uint32_t n, const float* pos_xyz_world,
                           const float* oriQ_wxyz, const float* vel_xyz, const float* omg_xyz,
                           const uint8_t* family);
/* AddOwnerNextStepAcc / AddOwnerNextStepAngAcc (src/DEM/API.h:477-486, dT.cpp:3160-3174; DEMTracker::AddAcc / AddAngAcc):
 * n consecutive owners get an extra acceleration (world frame) and / or angular acceleration (owner frame) for the NEXT
 * step only, on top of what that step's contacts give.  Either pointer may be NULL. */
int dem_add_owner_acc(DemCtx* ctx, uint32_t first, uint32_t n, const float* acc_xyz, const float* angacc_local_xyz);
/* SetFamilyClumpMaterial / SetFamilyMeshMaterial (src/DEM/APIPublic.cpp:1597-1604, dT.cpp:2719-2738): the spheres
 * (meshes != 0: the facets) of every owner in `family` get `material`; contact history is kept. */
int dem_set_family_material(DemCtx* ctx, uint32_t family, uint32_t material, int meshes);
/* contact list in the reference's order (type, idA, idB). Pass NULL arrays to only query the count. */
int dem_download_contacts(DemCtx* ctx, uint64_t capacity, uint64_t* n, uint32_t* idA, uint32_t* idB, uint8_t* type,
                          float* wildcards4, float* force_xyz);
/* the same rows plus the contact point (world frame) each force acts at -- the per-contact record behind
 * GetOwnerContactForces / DEMTracker::GetContactForces (dT.cpp:2740-2791, DEMDynamicMisc.cu:14-100) and the contact
 * file's point columns. Needs record_contact_forces; the point is the one of the last force evaluation. */
int dem_download_contact_records(DemCtx* ctx, uint64_t capacity, uint64_t* n, uint32_t* idA, uint32_t* idB,
                                 uint8_t* type, float* wildcards4, float* force_xyz, float* point_xyz);
int dem_get_stats(DemCtx* ctx, DemStats* out);

/* reductions over clump owners (DEMInspector built-ins, AuxClasses.cpp:88-164) */
enum { DEM_REDUCE_MAX_ABSV = 0, DEM_REDUCE_MAX_Z = 1, DEM_REDUCE_MIN_Z = 2, DEM_REDUCE_KINETIC_ENERGY = 3,
       DEM_REDUCE_TOTAL_MASS = 4,
       /* (dem_reduce only) the reference's sphere-level forms, AuxClasses.cpp:19-50: top / bottom of every sphere (centre
        * z +- radius) and the speed of every sphere centre (v + omega x r), where 0..2 look at the owners' centres */
       DEM_REDUCE_SPHERE_MAX_Z = 5, DEM_REDUCE_SPHERE_MIN_Z = 6, DEM_REDUCE_SPHERE_MAX_ABSV = 7 };
int dem_reduce(DemCtx* ctx, int kind, double* out);
/* Several reductions in one pass over the owners and one read-back: bit k of kind_mask selects DEM_REDUCE_<k>;
 * out[k] receives it (entries of unselected kinds are left untouched). What a caller polling several inspectors per
 * frame (DEMInspector::GetValue in a loop, e.g. DEMdemo_Mixer.cpp:130-140) should use. */
int dem_reduce_many(DemCtx* ctx, uint32_t kind_mask, double out[5]);

/* ---- multi-GPU: slab decomposition with ghost-owner halo exchange (no counterpart in the reference, whose
 * "multi-GPU" is the kT/dT thread pair of APIPublic.cpp:35-48; extends the <= 2 GPUs of src/DEM/API.h:53,56 to the 8 of a
 * box).  Every rank holds a context that was given the SAME complete input; from dem_mgpu_init* on each rank integrates
 * the owners whose centre lies in its x-slab and exchanges the halo owners with its neighbours each step.  All exchanges
 * -- per step and per rebuild -- are device-driven over peer-mapped memory (NVLink stores + system-scope flag words): no
 * library call and no host synchronisation on the path; NCCL only carries the cudaIpc handles at set-up in the
 * one-process-per-GPU form.  Wall and mesh owners are replicated (every rank registers the facets that can meet its
 * slab): exact while they are fixed or follow a fully prescribed motion -- anything else is refused (DEM_ERR_INVALID).
 *   dem_mgpu_init        one process per GPU (torchrun): every rank calls it with the id rank 0 created
 *   dem_mgpu_init_local  all ranks are contexts of the calling process, one per GPU (what DEMSolver(nGPUs) uses); such
 *                        contexts must be stepped together through dem_group_step_async / dem_group_sync, because
 *                        their kernels wait for each other on the device */
int dem_mgpu_unique_id(uint8_t out[128]);
int dem_mgpu_init(DemCtx* ctx, int rank, int world, const uint8_t unique_id[128]);
int dem_mgpu_init_local(DemCtx** ctxs, int world);
/* enqueue a device-side barrier over the ranks on the context's stream (no host wait) */
int dem_mgpu_barrier(DemCtx* ctx);
/* n steps on every context of a local group (enqueued cycle by cycle across the ranks); dem_group_sync blocks until all
 * are done and replays on all ranks what a failed (overflowed) rebuild dropped */
int dem_group_step_async(DemCtx** ctxs, int world, uint64_t n_steps);
int dem_group_sync(DemCtx** ctxs, int world);
/* after dem_group_sync: ctxs[0] receives the records of the clump owners the other ranks own, so that it holds the
 * merged state of the whole system for trackers / writers / inspectors */
int dem_group_gather(DemCtx** ctxs, int world);
/* out (as of the last confirmed rebuild): [0] owners owned [1] owners active (own + ghost) [2] sent left [3] sent right
 * [4] halo bytes sent per step [5] world size in the low 32 bits; bit 32 set when decomposed (peer-memory exchange) */
int dem_mgpu_info(DemCtx* ctx, uint64_t out[6]);
/* the x-slab (LBF-relative) of `rank` out of `world` */
int dem_host_slab_bounds(const DemSimParams* p, int world, int rank, float* lo, float* hi);
/* Host statement of the ownership rule the device applies at every rebuild: for each of the first n owners (clump
 * owners; position codes as uploaded) role[i] = 1 if `rank` owns it (centre inside its slab), 2 if it is a ghost here
 * (owned by a neighbour, within `halo` of the shared cut), 0 if this rank never sees it; send[i] bit 0 / bit 1 set when
 * an owned owner must be sent to the left / right neighbour.  Both sides of a cut derive the same answer from the same
 * position codes, so no negotiation is needed. */
int dem_host_partition_owners(const DemSimParams* p, int world, int rank, float halo, uint64_t n, const uint64_t* voxelID,
                              const uint16_t* locX, uint8_t* role, uint8_t* send);

/* "adaptive_update_freq" (0/1; UseAdaptiveUpdateFreq, src/DEM/API.h, dT.h:721-752): let the core pick the steps per
 * contact-list cycle that costs the least device time per step, starting from DemSimParams::cd_update_freq and staying
 * inside ["update_freq_min", "update_freq_max"] (SetCDMaxUpdateFreq).  Single-device contexts only: the ranks of a
 * decomposition must agree on the cycle length, so there it stays where it was set.
 * Execution knobs that do not change results: "ctas_per_sm" (2..4, register budget / occupancy of the force kernel),
 * "fast_encode" (0/1), "sort_mode" (0 radix sort, 1 counting sort; identical order), "keep_acc" (0/1: write per-owner accelerations every step for ContactAcc trackers) */
int dem_set_option(DemCtx* ctx, const char* name, double value);

/* ---- measurement hooks (bench.py / ncu) ---------------------------------------------------------------- */
/* Run n steps and return the mean device time per kernel in microseconds, measured with CUDA events on the
 * launching stream: [0]=sphere-sphere force kernel [1]=sphere-analytical force kernel [2]=integration kernel
 * [3]=contact rebuild amortised per step [4]=whole step [5]=ghost-owner halo exchange (multi-GPU; includes waiting for
 * the neighbour) [6..7] reserved.  The wall / mesh kernels run serialised here (no side stream), so [4] is an upper
 * bound of the production step. */
int dem_profile_steps(DemCtx* ctx, uint64_t n_steps, float out_us[8]);
/* One contact-list rebuild with CUDA events between its stages (microseconds): [0] margins + cell keys + histogram +
 * sphere-analytical list [1] sort [2] cell-table scan [3] gather [4] sweep (+ history carry-over) [5] counts
 * [6] multi-GPU: re-deciding ownership + halo lists (part of [0]) [7] whole rebuild */
int dem_profile_rebuild(DemCtx* ctx, float out_us[8]);

/* Binning + sort only (the neighbour-search stress case): `repeats` times { margins, sphere world positions, cell keys,
 * cell histogram ; sort into (cell, sphere id) order ; gather of the sorted sphere stream }, no sweep.  Mean device time
 * in microseconds: [0] positions + keys + histogram [1] sort (+ gather) [2] both.  Leaves the sorted arrays behind for
 * dem_debug_download. */
int dem_profile_binning(DemCtx* ctx, uint32_t repeats, float out_us[3]);

/* Raw views of device scratch of the LAST rebuild / step, for tests and tools (synchronises).  `what`:
 *   "sphere_keys" u32[nSpheres]  cell key of each sphere (0xffffffff = not held by this rank)
 *   "sorted_keys" u32[nSpheres]  cell keys in sorted order      "sorted_ids" u32[nSpheres]  sphere ids in sorted order
 *   "sphere_pos"  f32[4*nSpheres] LBF-relative centre + inflated radius
 * Copies min(n, available) elements and returns the number copied in *n_out. */
int dem_debug_download(DemCtx* ctx, const char* what, void* out, uint64_t n, uint64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* DEM_B200_H */
