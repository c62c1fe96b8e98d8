// dem_core.cu -- C ABI (include/dem_b200.h) of the B200-native DEM core: context, device arena, the step loop and the
// contact-list rebuild driver.  Replaces, for the hot path only, the two worker loops of the reference
// (dT workerThread src/DEM/dT.cpp:2324-2479, kT workerThread src/DEM/kT.cpp:218-320) and their
// cudaMemcpy mailbox hand-shake (dT.cpp:1989-2038, kT.cpp:193-216) with ONE in-order stream: the rebuild runs on the
// same GPU right before the force kernel that first uses the list, so the list is never stale by more than
// cd_update_freq steps and no peer copies exist.  The host never waits inside a rebuild: every count and flag stays on
// the device, the rebuild's last kernel leaves a status record in pinned memory, and the host reads that record one
// cycle later (right before it enqueues the next rebuild).  A rebuild that could not hold its result poisons the
// context on the device -- every later kernel returns at once -- and the host, when it finds out, grows the arrays,
// rolls its own bookkeeping back to that rebuild and replays.  There is no CPU fallback: without a device every entry
// point fails.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound with dlopen so that single-GPU use needs no NCCL at all

#include "dem_kernels.h"

using namespace demb;
static_assert(sizeof(DemSimParams) == 120, "DemSimParams layout is part of the C ABI");
static_assert(sizeof(DemPrescription) == 88 && sizeof(DemPrescription) == sizeof(Prescr), "DemPrescription layout");

namespace {

// NCCL entry points, resolved at run time from the NCCL already loaded in the process (torch's) or the system one.
// Used ONLY to bootstrap the multi-process decomposition (all-gather of the cudaIpc handles of the peer blocks); every
// exchange after that is device-driven over the peer-mapped blocks (kernels_mgpu.cu).
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) return false;
#define NCCL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(handle, name)); if (!field) return false;
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId") NCCL_SYM(CommInitRank, "ncclCommInitRank")
        NCCL_SYM(CommDestroy, "ncclCommDestroy") NCCL_SYM(AllGather, "ncclAllGather")
        NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;

// domain decomposition state of one rank (device side: MgDev, dem_device.cuh)
struct MgState {
    bool on = false;
    bool local = false;  // all ranks are contexts of THIS process (peer access) instead of one process per GPU (cudaIpc)
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    float cut_lo = 0.f, cut_hi = 0.f;
    uint32_t cap = 0;
    char* my_block = nullptr;
    char* peer_block[MG_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool peer_ipc[MG_MAX_WORLD] = {false, false, false, false, false, false, false, false};
    unsigned long long* d_ctrs = nullptr;  // [0] epoch [1] mail counter [2..3] block counters (as u32)
    uint8_t* d_flag = nullptr;
    uint32_t* d_active_list[2] = {nullptr, nullptr};
    uint32_t* d_counts = nullptr;  // [2 parities][8]
    uint32_t* d_send_gid[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    uint32_t* d_act_sph[2] = {nullptr, nullptr};
    uint2* d_owner_sph = nullptr;
    uint32_t last[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // counts of the last confirmed rebuild (RebuildStatus::mg)
};

struct ListBuf {
    uint32_t* idB = nullptr;
    uint4* cinfo = nullptr;
    float4* hist = nullptr;
    uint32_t* seg_start = nullptr;
    uint32_t* seg_count = nullptr;
    uint32_t* count = nullptr;
    float4* force = nullptr;
    float4* cpoint = nullptr;
    uint8_t* due = nullptr;  // sphere--sphere candidate list only
};

// host bookkeeping at the moment a rebuild was enqueued: what the host returns to when that rebuild turns out to
// have failed on the device
struct PendingRebuild {
    bool valid = false;
    uint32_t seq = 0;
    uint64_t n_steps = 0, steps_since = 0, n_rebuilds = 0;
    double sim_time = 0.0;
    int cur = 0, maxvel_slot = 0;
    bool need_maxvel = false;
};

}  // namespace

struct DemCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t side = nullptr;   // wall / mesh contact kernels run here, next to the sphere--sphere kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t side2 = nullptr;  // decomposed runs: the halo exchange runs here, next to the integration of the bulk
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
    int overlap_walls = 1;
    int num_sms = 148;
    std::string err;
    bool initialized = false;
    bool params_set = false;
    DemSimParams sp{};

    // host copies of the flattened input
    std::vector<float4> h_comp, h_massprop;
    std::vector<MatPair> h_matpair;
    uint32_t nMat = 0;
    std::vector<AnalObj> h_anal;
    std::vector<uint8_t> h_masks;
    std::vector<float> h_extra;
    std::vector<Prescr> h_presc;
    std::vector<OwnerState> h_state;
    std::vector<float4> h_spin;
    std::vector<uint2> h_sph;
    std::vector<float4> h_tri1, h_tri2, h_tri3;
    std::vector<uint2> h_tri_info;
    uint32_t nOwners = 0, nSpheres = 0, nAnal = 0, nTri = 0, nClumpOwners = 0;
    float rmax = 0.f, max_extra = 0.f;
    uint32_t any_mask = 0;

    // device arrays
    OwnerState* d_state = nullptr;
    float4* d_spin = nullptr;
    Wrench* d_wrench = nullptr;
    Wrench* d_acc = nullptr;
    uint2* d_sph = nullptr;
    float4* d_comp = nullptr;
    float4* d_massprop = nullptr;
    MatPair* d_matpair = nullptr;
    AnalObj* d_anal = nullptr;
    // family tables live in one device blob [masks | extra margins | prescriptions] refreshed with ONE copy from a
    // pinned staging buffer (a co-simulating caller re-sends them every step); two staging buffers alternate so that
    // the host only ever waits for the upload before the previous one
    char* d_famblob = nullptr;
    char* h_famblob[2] = {nullptr, nullptr};  // pinned
    cudaEvent_t ev_fam[2] = {nullptr, nullptr};
    int fam_slot = 0;
    uint8_t* d_masks = nullptr;
    float* d_extra = nullptr;
    Prescr* d_presc = nullptr;
    uint32_t* d_flags = nullptr;   // DEM_NUM_FLAGS status words (DEM_FLAG_*)
    float* d_maxvel = nullptr;
    double* d_reduce = nullptr;
    double* d_reduce_many = nullptr;
    // contact lists [kind][buffer]: kind 0 = sphere-sphere in touch at the last rebuild, 1 = other sphere-sphere
    // candidates, 2 = sphere-analytical, 3 = sphere-triangle; two buffers each (current / being rebuilt)
    ListBuf lists[4][2];
    int cur = 0;  // index of the current list buffers
    uint64_t capacity = 0;
    // rebuild scratch
    GridInfo* d_grid = nullptr;
    float4* d_sphF = nullptr;
    uint32_t* d_keys[2] = {nullptr, nullptr};
    uint32_t* d_vals[2] = {nullptr, nullptr};
    uint32_t* d_cellStart = nullptr;
    float4* d_sortedSph = nullptr;
    uint2* d_sortedAux = nullptr;
    uint4* d_sortedMeta = nullptr;
    AnalWorld* d_analw = nullptr;
    uint32_t* d_rs_hist = nullptr;
    uint32_t* d_scan_tmp = nullptr;
    unsigned long long* d_scan_desc = nullptr;
    uint32_t* d_cand = nullptr;               // candidates accepted by the sweep's count pass (SW_REC = 12 words per sphere)
    uint32_t* d_idA[2] = {nullptr, nullptr};  // sphere A of every new sphere--sphere contact (fill pass -> history pass)
    // triangles
    float4* d_tri[3] = {nullptr, nullptr, nullptr};   // owner-frame nodes
    uint2* d_tri_info = nullptr;
    float4* d_triW[3] = {nullptr, nullptr, nullptr};  // world nodes of the last rebuild
    uint32_t* d_triCellStart = nullptr;
    uint32_t* d_triCellFill = nullptr;
    uint32_t* d_triCellList = nullptr;
    uint64_t tri_pair_cap = 0;
    uint32_t max_cells = 0;
    int key_bits = 8;
    // pinned host memory
    uint32_t* h_pinned = nullptr;        // 256 words of read-back scratch (reductions at words 112..131)
    RebuildStatus* h_status = nullptr;   // ring of REBUILD_STATUS_SLOTS records written by k_finish_counts (mapped)
    RebuildStatus* d_status = nullptr;   // the device's view of it

    // bookkeeping
    uint64_t n_steps = 0, n_rebuilds = 0, launches = 0, device_bytes = 0;
    uint64_t steps_target = 0;  // steps asked for so far (n_steps falls behind it while a failed rebuild is replayed)
    uint64_t steps_since_rebuild = 0;
    uint32_t list_freq = 0;   // the drift (steps) the lists in use were built for: they are rebuilt after that many steps
    bool need_rebuild = true;
    bool need_maxvel = true;  // velocities changed outside the integrator
    int maxvel_slot = 0;
    double sim_time = 0.0;
    uint64_t n_list[4] = {0, 0, 0, 0};
    GridInfo last_grid{};
    uint32_t overflow_seen = 0;
    uint32_t seq_host = 0;    // rebuilds enqueued so far == DEM_FLAG_SEQ on the device once they have run
    PendingRebuild pending;
    // Adaptive update frequency (UseAdaptiveUpdateFreq; the reference's tuner is AccumStepUpdater, src/DEM/dT.h:721-752,
    // dT.cpp:2280-2297: it sizes the drift margin to how many steps dT gets through per kT update).  Here the list is
    // rebuilt in-stream, so the question is only which frequency costs the least device time per step -- a longer cycle
    // amortises the rebuild, a shorter one keeps the margin (hence the candidate list the force kernel walks) small.
    // A hill climb on the measured device time of whole cycles answers it: events recorded at the head of every
    // rebuild, read back when that rebuild is confirmed (no extra synchronisation).
    struct FreqTuner {
        int on = 0;
        int fmin = 4, fmax = 200;
        cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        uint64_t steps_at[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int freq_at[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        bool valid[8] = {false, false, false, false, false, false, false, false};
        double acc_us = 0.0;
        uint64_t acc_steps = 0;
        int acc_cycles = 0;
        double prev_us = -1.0;   // us per step measured at the previous frequency
        int prev_f = 0;
        int dir = +1;
        int settled = 0;         // consecutive probes that did not pay: hold the frequency for a while
        int hold_cycles = 0;
        uint64_t changes = 0;
    } tuner;
    // CUDA graph of one whole contact-list cycle (rebuild + cd_update_freq steps): [list buffer][max|v| slot]
    struct CycleGraph {
        cudaGraphExec_t exec = nullptr;
        DevParams P0;        // the parameters the captured kernels were launched with: validity check
        CdParams C0;
        uint32_t L = 0;      // steps in the graph
        int cfg[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t launches = 0;
    };
    CycleGraph graphs[2][2];
    int use_graph = 2;       // 0 off, 1 on, 2 auto (on for scenes small enough to be launch-bound and on several GPUs)
    uint64_t graph_launches = 0;
    int ctas_per_sm = 4;
    int fast_math = 1;  // sphere--sphere force kernel: MUFU reciprocal / rsqrt instead of IEEE division / sqrt
    int fast_encode = 1;
    // sphere--sphere force kernels (kernels_step.cu).  bit 0: leave candidates alone until the step at which they can
    // touch (k_force_ss_due, chosen for lists that live >= 32 steps); bit 1: fetch velocities only for pairs in touch;
    // bit 4: (with bit 0) compact the due candidates per warp; bit 5: every evaluation refreshes the step a candidate is
    // due at; bit 6: use k_force_ss_due whatever the list's life time; bits 2 / 3: measurement only
    int force_opts = 1 | 2 | 16 | 32;
    int sort_mode = 1;  // 0 radix sort, 1 counting sort + in-cell rank by sphere id (same order)
    bool keep_acc = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    MgState mg;
    // in-process GPU group (dem_ctx_create_group): this context is rank 0, peers are ranks 1..; every entry point called
    // on it acts on the whole group
    std::vector<DemCtx*> peers;
    bool group_on = false;     // the scene is sharded over the group (otherwise the peers idle)
    bool merged = true;        // this context holds the merged state of all ranks
    double group_min_owners = 50000.0;  // shard only when every GPU gets at least this many clump owners
    float rclump = 0.f;
    int sa_grid = 148;
    int last_sorted_buf = 0;
};

namespace {

// applyOriQToVector3 on the host (reference DEMHelperKernels.cuh:161-173); q = {w,x,y,z}
float3 host_rotate(float3 v, float4 q) {
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    float3 r;
    r.x = (2.0f * (w * w + x * x) - 1.0f) * v.x + (2.0f * (x * y - w * z)) * v.y + (2.0f * (x * z + w * y)) * v.z;
    r.y = (2.0f * (x * y + w * z)) * v.x + (2.0f * (w * w + y * y) - 1.0f) * v.y + (2.0f * (y * z - w * x)) * v.z;
    r.z = (2.0f * (x * z - w * y)) * v.x + (2.0f * (y * z + w * x)) * v.y + (2.0f * (w * w + z * z) - 1.0f) * v.z;
    return r;
}

int mg_reset_ownership(DemCtx* ctx);

int fail(DemCtx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(ctx, DEM_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
int dalloc(DemCtx* ctx, T** p, size_t n) {
    if (n == 0) n = 1;
    CK(cudaMalloc((void**)p, n * sizeof(T)));
    ctx->device_bytes += n * sizeof(T);
    return DEM_OK;
}
template <typename T>
void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

int alloc_list(DemCtx* ctx, ListBuf& L, uint64_t cap, uint32_t nSph, bool hist, bool force, bool due = false) {
    int rc;
    if (due) {  // (+ one filter round of k_force_ss beyond the last entry; 0 = look at the pair at every step)
        if ((rc = dalloc(ctx, &L.due, cap + 1024))) return rc;
        CK(cudaMemset(L.due, 0, cap + 1024));
    }
    if ((rc = dalloc(ctx, &L.idB, cap))) return rc;
    if ((rc = dalloc(ctx, &L.cinfo, cap))) return rc;
    if (hist && (rc = dalloc(ctx, &L.hist, cap))) return rc;
    if (force && (rc = dalloc(ctx, &L.force, cap))) return rc;
    if (force && (rc = dalloc(ctx, &L.cpoint, cap))) return rc;
    if ((rc = dalloc(ctx, &L.seg_start, (size_t)nSph + 1))) return rc;
    if ((rc = dalloc(ctx, &L.seg_count, (size_t)nSph + 1))) return rc;
    if ((rc = dalloc(ctx, &L.count, 4))) return rc;
    CK(cudaMemset(L.seg_start, 0, sizeof(uint32_t) * ((size_t)nSph + 1)));
    CK(cudaMemset(L.seg_count, 0, sizeof(uint32_t) * ((size_t)nSph + 1)));
    CK(cudaMemset(L.count, 0, sizeof(uint32_t) * 4));
    // the record of a contact that no force kernel has visited yet reads as zero force at the origin
    if (force) {
        CK(cudaMemset(L.force, 0, sizeof(float4) * std::max<uint64_t>(cap, 1)));
        CK(cudaMemset(L.cpoint, 0, sizeof(float4) * std::max<uint64_t>(cap, 1)));
    }
    return DEM_OK;
}
void free_list(ListBuf& L) {
    dfree(L.idB); dfree(L.cinfo); dfree(L.hist); dfree(L.force); dfree(L.cpoint); dfree(L.seg_start); dfree(L.seg_count);
    dfree(L.count); dfree(L.due);
}

ContactList as_list(const ListBuf& L) {
    ContactList c;
    c.idB = L.idB; c.cinfo = L.cinfo; c.hist = L.hist; c.seg_start = L.seg_start; c.seg_count = L.seg_count;
    c.count = L.count; c.force = L.force; c.cpoint = L.cpoint; c.due = L.due;
    return c;
}

MgDev make_mgdev(const DemCtx* c) {
    MgDev M;
    memset(&M, 0, sizeof(M));
    const MgState& g = c->mg;
    M.world = 1;
    if (!g.on) return M;
    M.rank = g.rank; M.world = g.world;
    M.has[0] = g.rank > 0; M.has[1] = g.rank < g.world - 1;
    M.my_block = g.my_block;
    for (int r = 0; r < g.world; r++) M.peer_block[r] = g.peer_block[r];
    M.cap = g.cap;
    M.epoch = g.d_ctrs;
    M.mail_ctr = g.d_ctrs + 1;
    M.block_ctr = reinterpret_cast<uint32_t*>(g.d_ctrs + 2);
    M.flag = g.d_flag;
    for (int p = 0; p < 2; p++) {
        M.active_list[p] = g.d_active_list[p];
        M.counts[p] = g.d_counts + 8 * p;
        for (int d = 0; d < 2; d++) M.send_gid[p][d] = g.d_send_gid[p][d];
    }
    M.act_sph[0] = g.d_act_sph[0]; M.act_sph[1] = g.d_act_sph[1];
    M.owner_sph = g.d_owner_sph;
    M.cut_lo = g.cut_lo; M.cut_hi = g.cut_hi;
    M.nClumpOwners = c->nClumpOwners;
    return M;
}

DevParams make_params(const DemCtx* c) {
    DevParams P;
    memset(&P, 0, sizeof(P));
    const DemSimParams& s = c->sp;
    P.nvXp2 = s.nvXp2; P.nvYp2 = s.nvYp2;
    P.l = s.l; P.voxelSize = s.voxelSize; P.inv_l = 1.0 / s.l;
    for (int k = 0; k < 3; k++) { P.LBF[k] = s.LBF[k]; P.G[k] = s.G[k]; }
    P.h = s.h;
    P.half_h = (float)(0.5 * s.h);
    P.integrator = s.integrator;
    P.nOwners = c->nOwners; P.nSpheres = c->nSpheres; P.nAnal = c->nAnal; P.nTri = c->nTri; P.nMat = c->nMat;
    P.beta = s.beta; P.approxMaxVel = s.approxMaxVel; P.expSafetyMulti = s.expSafetyMulti;
    P.expSafetyAdder = s.expSafetyAdder;
    P.maxDrift = s.cd_update_freq;
    P.state = c->d_state; P.spin = c->d_spin; P.wrench = c->d_wrench; P.acc_out = c->keep_acc ? c->d_acc : nullptr;
    P.fast_encode = (uint32_t)c->fast_encode;
    P.force_opts = (uint32_t)c->force_opts;
    P.inv_voxelSize = 1.0 / s.voxelSize;
    P.sph = c->d_sph; P.comp = c->d_comp; P.massprop = c->d_massprop; P.matpair = c->d_matpair; P.anal = c->d_anal;
    P.familyMasks = c->d_masks; P.familyExtraMargin = c->d_extra; P.presc = c->d_presc;
    P.ss = as_list(c->lists[0][c->cur]);
    P.sn = as_list(c->lists[1][c->cur]);
    P.sa = as_list(c->lists[2][c->cur]);
    P.st = as_list(c->lists[3][c->cur]);
    P.tri_n1 = c->d_tri[0]; P.tri_n2 = c->d_tri[1]; P.tri_n3 = c->d_tri[2]; P.tri_info = c->d_tri_info;
    P.flags = c->d_flags;
    if (c->mg.on) {
        const MgState& g = c->mg;
        P.active = g.d_flag;
        P.active_list = g.d_active_list[c->cur];
        P.nActivePtr = g.d_counts + 8 * c->cur + 3;
        P.halo_gid[0] = g.d_send_gid[c->cur][0]; P.halo_gid[1] = g.d_send_gid[c->cur][1];
        P.halo_counts = g.d_counts + 8 * c->cur;
    }
    P.maxvel = c->d_maxvel + c->maxvel_slot;
    P.maxvel_next = c->d_maxvel + (c->maxvel_slot ^ 1);
    P.errOutVel = s.errOutVel;
    return P;
}

// rebuild parameters: the new lists go to the buffers cur^1, the current ones are the "old" lists (history source)
CdParams make_cd(const DemCtx* c) {
    CdParams C;
    memset(&C, 0, sizeof(C));
    C.grid = c->d_grid;
    for (int k = 0; k < 3; k++) C.ext[k] = c->sp.userBoxMax[k] - c->sp.userBoxMin[k];
    // the binned region is the target box: positions are LBF relative and LBF is the target box corner
    for (int k = 0; k < 3; k++) {
        const float e = 2.f * (c->sp.userBoxMin[k] - c->sp.LBF[k]) + (c->sp.userBoxMax[k] - c->sp.userBoxMin[k]);
        if (e > C.ext[k]) C.ext[k] = e;
    }
    C.rmax = c->rmax; C.rclump = c->rclump; C.max_extra = c->max_extra; C.max_cells = c->max_cells; C.any_mask = c->any_mask;
    C.capacity = (uint32_t)c->capacity;
    C.sphF = c->d_sphF;
    C.keys[0] = c->d_keys[0]; C.keys[1] = c->d_keys[1]; C.vals[0] = c->d_vals[0]; C.vals[1] = c->d_vals[1];
    C.cellStart = c->d_cellStart; C.sortedSph = c->d_sortedSph; C.sortedAux = c->d_sortedAux; C.sortedMeta = c->d_sortedMeta;
    C.analw = c->d_analw;
    if (c->mg.on) {
        C.slab_on = 1; C.slab_lo = c->mg.cut_lo; C.slab_hi = c->mg.cut_hi;
        C.act_sph = c->mg.d_act_sph[c->cur ^ 1];
        C.act_count = c->mg.d_counts + 8 * (c->cur ^ 1) + 4;
    }
    C.oldss = as_list(c->lists[0][c->cur]);
    C.oldsn = as_list(c->lists[1][c->cur]);
    C.oldsa = as_list(c->lists[2][c->cur]);
    C.oldst = as_list(c->lists[3][c->cur]);
    C.triW1 = c->d_triW[0]; C.triW2 = c->d_triW[1]; C.triW3 = c->d_triW[2];
    C.triCellStart = c->d_triCellStart; C.triCellFill = c->d_triCellFill; C.triCellList = c->d_triCellList;
    C.tri_pair_cap = (uint32_t)c->tri_pair_cap;
    C.rs_hist = c->d_rs_hist; C.scan_tmp = c->d_scan_tmp; C.scan_desc = c->d_scan_desc;
    C.cand = c->d_cand;
    C.idA_ss = c->d_idA[0]; C.idA_sn = c->d_idA[1];
    C.status = c->d_status;
    return C;
}

void free_device(DemCtx* c) {
    dfree(c->d_state); dfree(c->d_spin); dfree(c->d_wrench); dfree(c->d_acc); dfree(c->d_sph); dfree(c->d_comp); dfree(c->d_massprop);
    dfree(c->d_matpair); dfree(c->d_anal); dfree(c->d_famblob); c->d_masks = nullptr; c->d_extra = nullptr; c->d_presc = nullptr;
    dfree(c->d_flags); dfree(c->d_maxvel); dfree(c->d_reduce); dfree(c->d_reduce_many);
    for (int kind = 0; kind < 4; kind++)
        for (int k = 0; k < 2; k++) free_list(c->lists[kind][k]);
    for (int k = 0; k < 3; k++) { dfree(c->d_tri[k]); dfree(c->d_triW[k]); }
    dfree(c->d_tri_info); dfree(c->d_triCellStart); dfree(c->d_triCellFill); dfree(c->d_triCellList);
    dfree(c->d_grid); dfree(c->d_sphF); dfree(c->d_keys[0]); dfree(c->d_keys[1]); dfree(c->d_vals[0]);
    dfree(c->d_vals[1]); dfree(c->d_cellStart); dfree(c->d_sortedSph); dfree(c->d_sortedAux); dfree(c->d_sortedMeta); dfree(c->d_analw);
    dfree(c->d_rs_hist); dfree(c->d_scan_tmp); dfree(c->d_scan_desc); dfree(c->d_idA[0]); dfree(c->d_idA[1]); dfree(c->d_cand);
    c->device_bytes = 0;
}

void free_mg(DemCtx* c) {
    MgState& g = c->mg;
    dfree(g.d_flag); dfree(g.d_active_list[0]); dfree(g.d_active_list[1]); dfree(g.d_act_sph[0]); dfree(g.d_act_sph[1]); dfree(g.d_owner_sph);
    dfree(g.d_counts); dfree(g.d_ctrs);
    for (int p = 0; p < 2; p++)
        for (int d = 0; d < 2; d++) dfree(g.d_send_gid[p][d]);
    for (int r = 0; r < (int)MG_MAX_WORLD; r++) {
        if (g.peer_block[r] && g.peer_ipc[r]) cudaIpcCloseMemHandle(g.peer_block[r]);
        g.peer_block[r] = nullptr;
        g.peer_ipc[r] = false;
    }
    if (g.my_block) cudaFree(g.my_block);
    g.my_block = nullptr;
    g.on = false;
}

constexpr size_t FAM_OFF_MASKS = 0;
constexpr size_t FAM_OFF_EXTRA = (DEM_NUM_FAMILY_MASKS + 255) / 256 * 256;
constexpr size_t FAM_OFF_PRESC = FAM_OFF_EXTRA + sizeof(float) * DEM_NUM_FAMILIES;
constexpr size_t FAM_BLOB_BYTES = FAM_OFF_PRESC + sizeof(Prescr) * DEM_NUM_FAMILIES;

// host tables -> pinned staging -> device blob, one asynchronous copy on the compute stream.  Two staging buffers
// alternate: the host waits for the upload before the previous one, which has long left its buffer.
int stage_families(DemCtx* ctx) {
    const int k = ctx->fam_slot;
    ctx->fam_slot ^= 1;
    if (!ctx->h_famblob[k]) {
        CK(cudaHostAlloc((void**)&ctx->h_famblob[k], FAM_BLOB_BYTES, cudaHostAllocDefault));
        memset(ctx->h_famblob[k], 0, FAM_BLOB_BYTES);
        CK(cudaEventCreateWithFlags(&ctx->ev_fam[k], cudaEventDisableTiming));
    } else {
        CK(cudaEventSynchronize(ctx->ev_fam[k]));  // the upload before the previous one has left this buffer
    }
    memcpy(ctx->h_famblob[k] + FAM_OFF_MASKS, ctx->h_masks.data(), DEM_NUM_FAMILY_MASKS);
    memcpy(ctx->h_famblob[k] + FAM_OFF_EXTRA, ctx->h_extra.data(), sizeof(float) * DEM_NUM_FAMILIES);
    memcpy(ctx->h_famblob[k] + FAM_OFF_PRESC, ctx->h_presc.data(), sizeof(Prescr) * DEM_NUM_FAMILIES);
    CK(cudaMemcpyAsync(ctx->d_famblob, ctx->h_famblob[k], FAM_BLOB_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev_fam[k], ctx->stream));
    return DEM_OK;
}

int alloc_lists(DemCtx* ctx, uint64_t cap) {
    const bool hist = ctx->sp.force_model == DEM_HERTZIAN;
    const bool rec = ctx->sp.record_contact_forces != 0;
    int rc;
    for (int kind = 0; kind < 4; kind++)
        for (int k = 0; k < 2; k++)
            if ((rc = alloc_list(ctx, ctx->lists[kind][k], (kind == 3 && ctx->nTri == 0) ? 1 : cap, ctx->nSpheres, hist, rec, kind == 1))) return rc;
    for (int k = 0; k < 2; k++)
        if ((rc = dalloc(ctx, &ctx->d_idA[k], cap))) return rc;
    ctx->capacity = cap;
    return DEM_OK;
}

#define NC(call)                                                                                                      \
    do {                                                                                                              \
        ncclResult_t r_ = (call);                                                                                     \
        if (r_ != ncclSuccess)                                                                                        \
            return fail(ctx, DEM_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, g_nccl.GetErrorString(r_)); \
    } while (0)

void graph_drop(DemCtx* ctx);

// Grow every contact list to `newcap` entries, keeping the contents of the CURRENT lists (the history source of the
// next rebuild).  Host-synchronous; only called with the stream idle.
int grow_lists(DemCtx* ctx, uint64_t newcap) {
    if (newcap > 0xfffffff0ull) return fail(ctx, DEM_ERR_CAPACITY, "contact list exceeds 2^32 entries");
    const uint64_t oldcap = ctx->capacity;
    const bool hist = ctx->sp.force_model == DEM_HERTZIAN, rec = ctx->sp.record_contact_forces != 0;
    for (int kind = 0; kind < (ctx->nTri ? 4 : 3); kind++) {
        int rc;
        free_list(ctx->lists[kind][ctx->cur ^ 1]);
        if ((rc = alloc_list(ctx, ctx->lists[kind][ctx->cur ^ 1], newcap, ctx->nSpheres, hist, rec, kind == 1))) return rc;
        ListBuf src = ctx->lists[kind][ctx->cur], dst;
        if ((rc = alloc_list(ctx, dst, newcap, ctx->nSpheres, hist, rec, kind == 1))) return rc;  // (due = 0: every step)
        CK(cudaMemcpy(dst.idB, src.idB, sizeof(uint32_t) * oldcap, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(dst.cinfo, src.cinfo, sizeof(uint4) * oldcap, cudaMemcpyDeviceToDevice));
        if (src.hist) CK(cudaMemcpy(dst.hist, src.hist, sizeof(float4) * oldcap, cudaMemcpyDeviceToDevice));
        if (src.force) CK(cudaMemcpy(dst.force, src.force, sizeof(float4) * oldcap, cudaMemcpyDeviceToDevice));
        if (src.cpoint) CK(cudaMemcpy(dst.cpoint, src.cpoint, sizeof(float4) * oldcap, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(dst.seg_start, src.seg_start, sizeof(uint32_t) * ((size_t)ctx->nSpheres + 1), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(dst.seg_count, src.seg_count, sizeof(uint32_t) * ((size_t)ctx->nSpheres + 1), cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(dst.count, src.count, sizeof(uint32_t) * 4, cudaMemcpyDeviceToDevice));
        free_list(src);
        ctx->lists[kind][ctx->cur] = dst;
    }
    for (int k = 0; k < 2; k++) {
        dfree(ctx->d_idA[k]);
        int rc = dalloc(ctx, &ctx->d_idA[k], newcap);
        if (rc) return rc;
    }
    ctx->capacity = newcap;
    graph_drop(ctx);
    return DEM_OK;
}

// ---- the rebuild, host side ---------------------------------------------------------------------------------------
// All kernels of ONE contact-list rebuild into the buffers cur^1, on ctx->stream; nothing here waits for the device.
// Returns the number of kernels launched.  sev (optional, 8 events) brackets the stages for dem_profile_rebuild.
int launch_rebuild_kernels(DemCtx* ctx, cudaEvent_t* sev) {
    DevParams P = make_params(ctx);
    const CdParams C = make_cd(ctx);
    const int par = ctx->cur ^ 1;
    P.ss = as_list(ctx->lists[0][par]);
    P.sn = as_list(ctx->lists[1][par]);
    P.sa = as_list(ctx->lists[2][par]);
    P.st = as_list(ctx->lists[3][par]);
    cudaStream_t s = ctx->stream;
    const MgDev M = make_mgdev(ctx);
    const MgDev* Mp = ctx->mg.on ? &M : nullptr;
    int launches = launch_zero_u32(ctx->d_flags, 2, ctx->d_flags, ctx->num_sms, s);  // capacity bits + triangle demand
    if (sev) cudaEventRecord(sev[0], s);
    launches += launch_cd_prepare(P, C, Mp, ctx->need_maxvel, 0, ctx->num_sms, s);
    launches += launch_cd_prepare(P, C, Mp, ctx->need_maxvel, 1, ctx->num_sms, s);
    if (sev) cudaEventRecord(sev[8], s);
    if (ctx->mg.on) launches += launch_mg_redistribute(P, M, ctx->d_grid, par, ctx->num_sms, s);
    if (sev) cudaEventRecord(sev[9], s);
    launches += launch_cd_prepare(P, C, Mp, ctx->need_maxvel, 2, ctx->num_sms, s);
    launches += launch_cd_triangles(P, C, 0, ctx->num_sms, s);  // triangle -> cell registration
    launches += launch_cd_triangles(P, C, 1, ctx->num_sms, s);  // per-sphere triangle candidates (+ history)
    if (sev) cudaEventRecord(sev[1], s);
    int sorted_buf = -1;  // -1: counting sort inside the sweep stage
    if (ctx->sort_mode == 0) { launches += launch_cd_sort(P, C, ctx->key_bits, s, &sorted_buf); ctx->last_sorted_buf = sorted_buf; }
    if (sev) cudaEventRecord(sev[2], s);
    launches += launch_cd_sweep(P, C, Mp, par, sorted_buf, ctx->num_sms, s, sev ? sev + 3 : nullptr);  // records sev[3..6]
    if (sev) cudaEventRecord(sev[7], s);
    return launches;
}

// adaptive update frequency: time stamp at the head of the rebuild about to be enqueued (sequence number seq_host + 1)
void tuner_mark(DemCtx* ctx) {
    DemCtx::FreqTuner& t = ctx->tuner;
    if (!t.on || ctx->mg.on) return;
    const uint32_t k = (ctx->seq_host + 1u) & 7u;
    if (!t.ev[k] && cudaEventCreate(&t.ev[k]) != cudaSuccess) { cudaGetLastError(); return; }
    cudaEventRecord(t.ev[k], ctx->stream);
    t.steps_at[k] = ctx->n_steps;
    t.freq_at[k] = (int)ctx->sp.cd_update_freq;
    t.valid[k] = true;
}
// rebuild `seq` has just been confirmed: the cycle before it (head seq-1 .. head seq) is complete on the device
void tuner_update(DemCtx* ctx, uint32_t seq) {
    DemCtx::FreqTuner& t = ctx->tuner;
    if (!t.on || ctx->mg.on) return;
    const uint32_t k1 = seq & 7u, k0 = (seq - 1u) & 7u;
    const int f = (int)ctx->sp.cd_update_freq;
    if (!t.valid[k0] || !t.valid[k1] || t.freq_at[k0] != f || t.steps_at[k1] <= t.steps_at[k0]) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.ev[k0], t.ev[k1]) != cudaSuccess) { cudaGetLastError(); return; }
    const uint64_t steps = t.steps_at[k1] - t.steps_at[k0];
    if (steps != (uint64_t)f) return;  // (a partial cycle: a host call asked for a rebuild in between)
    if (t.hold_cycles > 0) { t.hold_cycles--; return; }
    t.acc_us += (double)ms * 1000.0;
    t.acc_steps += steps;
    if (++t.acc_cycles < 3) return;
    const double us = t.acc_us / (double)t.acc_steps;
    t.acc_us = 0.0; t.acc_steps = 0; t.acc_cycles = 0;
    if (t.prev_us > 0.0 && t.prev_f != f) {
        if (us > t.prev_us * 0.995) {           // the move did not pay: go back and try the other way later
            t.dir = -t.dir;
            t.settled++;
        } else {
            t.settled = 0;
        }
    }
    if (t.settled >= 2) {                        // both neighbours are worse: stay, look again after a while
        t.settled = 0;
        t.hold_cycles = 100;
        const int best = (t.prev_us > 0.0 && t.prev_us < us) ? t.prev_f : f;
        t.prev_us = -1.0;
        if (best != f) { ctx->sp.cd_update_freq = (uint32_t)best; t.changes++; }
        return;
    }
    t.prev_us = us;
    t.prev_f = f;
    const int stepf = std::max(1, f / 6);
    const int nf = std::min(t.fmax, std::max(t.fmin, f + t.dir * stepf));
    if (nf != f) { ctx->sp.cd_update_freq = (uint32_t)nf; t.changes++; }
    else t.dir = -t.dir;
}

// host bookkeeping of a rebuild that has just been enqueued (plain launches or as the head of a cycle graph)
void note_rebuild_enqueued(DemCtx* ctx) {
    PendingRebuild& p = ctx->pending;
    p.valid = true;
    p.seq = ++ctx->seq_host;
    p.n_steps = ctx->n_steps; p.steps_since = ctx->steps_since_rebuild; p.n_rebuilds = ctx->n_rebuilds;
    p.sim_time = ctx->sim_time;
    p.cur = ctx->cur; p.maxvel_slot = ctx->maxvel_slot; p.need_maxvel = ctx->need_maxvel;
    ctx->cur ^= 1;
    ctx->n_rebuilds++;
    ctx->steps_since_rebuild = 0;
    ctx->list_freq = ctx->sp.cd_update_freq;
    ctx->need_rebuild = false;
    ctx->need_maxvel = false;
}

// Wait for the status record of the pending rebuild and act on it.  *rolled_back = true: the rebuild had overflowed, the
// arrays were grown and the host state is back where that rebuild was enqueued (need_rebuild set): the caller replays.
int confirm_rebuild(DemCtx* ctx, bool* rolled_back) {
    if (rolled_back) *rolled_back = false;
    PendingRebuild& p = ctx->pending;
    if (!p.valid) return DEM_OK;
    volatile RebuildStatus* st = ctx->h_status + (p.seq % REBUILD_STATUS_SLOTS);
    {
        // k_finish_counts writes seq last, after a system fence: spin on the pinned word (works for plain launches and
        // for graph replays alike; an event could not be waited on from inside a captured cycle)
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        while (st->seq != p.seq) {
            if ((++spins & 0x3ffu) == 0) {
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady)
                    return fail(ctx, DEM_ERR_CUDA, "the device failed during a contact-list rebuild: %s", cudaGetErrorString(q));
                if (q == cudaSuccess && st->seq != p.seq)
                    return fail(ctx, DEM_ERR_CUDA, "rebuild %u finished without leaving its status record (found %u)", p.seq, st->seq);
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
                    return fail(ctx, DEM_ERR_CUDA, "timed out waiting for contact-list rebuild %u", p.seq);
                std::this_thread::yield();
            }
        }
    }
    RebuildStatus r;
    memcpy(&r, const_cast<RebuildStatus*>(st), sizeof(r));
    p.valid = false;
    if (r.haloflags & 64u)
        return fail(ctx, DEM_ERR_CUDA, "multi-GPU: a neighbouring rank stopped answering (halo exchange timed out)");
    if (r.poison != 0u) {
        // ---- the rebuild could not hold its result: every kernel after it was a no-op.  Drain, grow, roll back. ----
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->side) CK(cudaStreamSynchronize(ctx->side));
        ctx->overflow_seen++;
        ctx->n_steps = p.n_steps; ctx->steps_since_rebuild = p.steps_since; ctx->n_rebuilds = p.n_rebuilds;
        ctx->sim_time = p.sim_time; ctx->cur = p.cur; ctx->maxvel_slot = p.maxvel_slot; ctx->need_maxvel = p.need_maxvel;
        ctx->need_rebuild = true;
        if (r.capflags & 8u)
            return fail(ctx, DEM_ERR_CAPACITY, "multi-GPU: the halo of a rank holds more owners than its buffer of %u; "
                        "use fewer ranks or a wider domain", ctx->mg.cap);
        if (r.capflags & 16u) {
            // the triangle--cell table was too small: it is scratch, so just regrow it
            dfree(ctx->d_triCellList);
            ctx->tri_pair_cap = (uint64_t)r.tri_demand + r.tri_demand / 2 + 1024;
            int rc = dalloc(ctx, &ctx->d_triCellList, ctx->tri_pair_cap);
            if (rc) return rc;
            graph_drop(ctx);
        }
        if (r.capflags & ~(8u | 16u)) {
            uint64_t need = 0;
            for (int kind = 0; kind < 4; kind++) need = std::max<uint64_t>(need, r.demand[kind]);
            // (on several GPUs the rank that overflowed may be another one: grow anyway, all ranks replay together)
            const uint64_t newcap = std::max<uint64_t>(need + need / 4 + 1024, need > ctx->capacity ? ctx->capacity * 2 : ctx->capacity);
            if (newcap > ctx->capacity) {
                int rc = grow_lists(ctx, newcap);
                if (rc) return rc;
            }
        }
        uint32_t zeros[DEM_NUM_FLAGS] = {0, 0, 0, 0, 0, 0, 0, 0};
        zeros[DEM_FLAG_SEQ] = ctx->seq_host;
        zeros[DEM_FLAG_VELOCITY] = r.velflag;
        CK(cudaMemcpy(ctx->d_flags, zeros, sizeof(zeros), cudaMemcpyHostToDevice));
        if (rolled_back) *rolled_back = true;
        return DEM_OK;
    }
    for (int kind = 0; kind < 4; kind++) ctx->n_list[kind] = r.count[kind];
    ctx->last_grid = r.grid;
    tuner_update(ctx, r.seq);
    if (ctx->mg.on) memcpy(ctx->mg.last, r.mg, sizeof(r.mg));
    if (r.velflag != 0u) {
        const uint32_t zero = 0;
        cudaMemcpyAsync(ctx->d_flags + DEM_FLAG_VELOCITY, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        return fail(ctx, DEM_ERR_VELOCITY, "an owner has a non-finite or too large velocity (max seen %.6g, limit %.6g) at t=%.9g",
                    ctx->last_grid.maxvel, ctx->sp.errOutVel, ctx->sim_time);
    }
    return DEM_OK;
}

// enqueue one rebuild with plain launches.  Confirms the previous one first (at most one rebuild is ever unconfirmed:
// its status slot and the buffers it read from stay untouched until then); when that turns into a roll-back nothing is
// enqueued and the caller's loop re-evaluates.
int enqueue_rebuild(DemCtx* ctx, float* stage_us = nullptr) {
    bool rolled = false;
    int rc = confirm_rebuild(ctx, &rolled);
    if (rc) return rc;
    if (rolled && !stage_us) return DEM_OK;
    cudaEvent_t sev[10];
    if (stage_us) for (auto& e : sev) cudaEventCreate(&e);
    tuner_mark(ctx);
    ctx->launches += launch_rebuild_kernels(ctx, stage_us ? sev : nullptr);
    note_rebuild_enqueued(ctx);
    if (stage_us) {
        CK(cudaStreamSynchronize(ctx->stream));
        // [0] margins+keys+histogram+analytical list [1] sort [2] cell-table scan [3] gather [4] sweep
        // [5] counts [6] redistribution over the ranks (inside [0]) [7] whole rebuild on the device
        for (int k = 0; k < 6; k++) cudaEventElapsedTime(&stage_us[k], sev[k], sev[k + 1]);
        cudaEventElapsedTime(&stage_us[6], sev[8], sev[9]);  // multi-GPU: ownership + halo lists (part of [0])
        cudaEventElapsedTime(&stage_us[7], sev[0], sev[7]);
        for (int k = 0; k < 8; k++) stage_us[k] *= 1000.f;
        for (auto& e : sev) cudaEventDestroy(e);
    }
    CK(cudaGetLastError());
    return DEM_OK;
}

// a rebuild NOW, confirmed before returning (dry run of DoDynamicsThenSync(0), dem_rebuild_contacts, profiling)
int rebuild_blocking(DemCtx* ctx, float* stage_us = nullptr) {
    for (int attempt = 0; attempt < 8; attempt++) {
        ctx->need_rebuild = true;
        int rc = enqueue_rebuild(ctx, stage_us);
        if (rc) return rc;
        if (!ctx->pending.valid) continue;  // the PREVIOUS rebuild had to be rolled back: go again
        bool rolled = false;
        rc = confirm_rebuild(ctx, &rolled);
        if (rc) return rc;
        if (!rolled) return DEM_OK;
    }
    return fail(ctx, DEM_ERR_CAPACITY, "contact list kept overflowing after repeated growth");
}

// decomposed runs: CTAs of the integrator, from the number of active owners the last confirmed rebuild reported (+15 %,
// rounded up to a multiple of 64 CTAs: the value only changes when the slab's population does)
int integrate_grid(const DemCtx* ctx) {
    if (!ctx->mg.on || ctx->mg.last[3] == 0) return 0;
    const uint64_t want = ((uint64_t)ctx->mg.last[3] * 115u / 100u + 255u) / 256u;
    return (int)((want + 63u) / 64u * 64u);
}

// the kernels of ONE step on ctx->stream (+ side stream); flips the max|v| slot. No bookkeeping, no rebuild.
int launch_step(DemCtx* ctx) {
    DevParams P = make_params(ctx);
    const int model = (int)ctx->sp.force_model;
    const bool rec = ctx->sp.record_contact_forces != 0;
    const bool walls = ctx->nAnal > 0 || ctx->nTri > 0;
    // The wall / mesh contact kernels are small (tens of thousands of contacts: latency, not bandwidth) and only meet
    // the sphere--sphere kernel in the wrench accumulator (atomic reductions): run them beside it on a second stream.
    cudaStream_t ws = (walls && ctx->overlap_walls && ctx->side) ? ctx->side : ctx->stream;
    if (ws != ctx->stream) {
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ws, ctx->ev_fork, 0));
    }
    if (ws != ctx->stream) {  // first, so that they are resident before the persistent kernel fills the SMs
        if (ctx->nAnal > 0) launch_force_sa(P, model, rec, ctx->sa_grid, ws);
        if (ctx->nTri > 0) launch_force_st(P, model, rec, ctx->sa_grid, ws);
        CK(cudaEventRecord(ctx->ev_join, ws));
    }
    launch_force_ss(P, model, rec, ctx->num_sms, ctx->ctas_per_sm, ctx->fast_math != 0, ctx->stream);
    if (ws == ctx->stream) {
        if (ctx->nAnal > 0) launch_force_sa(P, model, rec, ctx->sa_grid, ws);
        if (ctx->nTri > 0) launch_force_st(P, model, rec, ctx->sa_grid, ws);
    } else {
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    // integration.  On several GPUs: first the owners in the halo send lists, then -- side by side -- the exchange of
    // their records with the neighbours (k_mg_exchange: push, publish, wait, pull) and the integration of the rest
    if (ctx->mg.on) {
        const int hgrid = (int)(((uint64_t)ctx->mg.last[1] + ctx->mg.last[2]) * 115u / 100u / 256u + 64u) / 64 * 64;
        launch_integrate_halo(P, hgrid, ctx->stream);
        cudaStream_t xs = ctx->side2 ? ctx->side2 : ctx->stream;
        if (xs != ctx->stream) {
            CK(cudaEventRecord(ctx->ev_fork2, ctx->stream));
            CK(cudaStreamWaitEvent(xs, ctx->ev_fork2, 0));
        }
        ctx->launches += 1 + launch_mg_pull(P, make_mgdev(ctx), ctx->cur, ctx->num_sms, xs);
        if (xs != ctx->stream) CK(cudaEventRecord(ctx->ev_join2, xs));
        launch_integrate(P, integrate_grid(ctx), ctx->stream);
        if (xs != ctx->stream) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));
    } else {
        launch_integrate(P, 0, ctx->stream);
    }
    ctx->maxvel_slot ^= 1;  // the integrator left max |v| of the new state in the other slot
    ctx->launches += 2 + (ctx->nAnal > 0 ? 1 : 0) + (ctx->nTri > 0 ? 1 : 0);
    return DEM_OK;
}

void note_steps_enqueued(DemCtx* ctx, uint64_t n) {
    ctx->n_steps += n;
    ctx->steps_since_rebuild += n;
    for (uint64_t i = 0; i < n; i++) ctx->sim_time += (double)ctx->sp.h;
}

// ---- CUDA graph of one contact-list cycle -------------------------------------------------------------------------
// A cycle -- the rebuild and the cd_update_freq steps that use its lists -- launches the same kernels with the same
// parameters every time (all counts, the grid, exchange numbers live in device memory; only the list buffer and the
// max|v| slot alternate), so a scene whose step is shorter than the host's launch path (three to five launches, two
// event records and two stream waits: ~20 us), or a decomposed run whose ranks must not drift apart, is captured once per
// (list buffer, max|v| slot) and replayed with ONE cudaGraphLaunch; the graph stays valid for as long as the kernel
// parameters are byte-identical.
void graph_config(const DemCtx* ctx, int cfg[12]) {
    cfg[0] = (int)ctx->sp.force_model; cfg[1] = (int)ctx->sp.record_contact_forces; cfg[2] = ctx->ctas_per_sm;
    cfg[3] = ctx->fast_math; cfg[4] = ctx->overlap_walls; cfg[5] = (int)ctx->nAnal; cfg[6] = (int)ctx->nTri; cfg[7] = ctx->sa_grid;
    cfg[8] = ctx->sort_mode; cfg[9] = ctx->key_bits; cfg[10] = integrate_grid(ctx); cfg[11] = ctx->force_opts;
}
bool graph_wanted(const DemCtx* ctx) {
    if (ctx->use_graph == 0) return false;
    if (ctx->use_graph == 1 || ctx->mg.on) return true;
    return ctx->nSpheres <= 262144u;  // beyond that a step outlasts its launches
}
void graph_drop(DemCtx* ctx) {
    for (auto& row : ctx->graphs)
        for (auto& g : row) {
            if (g.exec) cudaGraphExecDestroy(g.exec);
            g.exec = nullptr;
            g.L = 0;
        }
}
// run one full cycle (the due rebuild + L = cd_update_freq steps) through the graph; sets *done on success, leaves it
// false when the caller should fall back to plain launches (or re-evaluate after a roll-back)
int run_cycle_graph(DemCtx* ctx, bool* done) {
    *done = false;
    bool rolled = false;
    int rc = confirm_rebuild(ctx, &rolled);
    if (rc) return rc;
    if (rolled || ctx->need_maxvel) return DEM_OK;
    const uint32_t L = ctx->sp.cd_update_freq;
    // (confirming the previous rebuild may have moved the update frequency: a whole cycle must still fit)
    if (L < 2 || ctx->steps_target - ctx->n_steps < L) return DEM_OK;
    DemCtx::CycleGraph& G = ctx->graphs[ctx->cur][ctx->maxvel_slot];
    const DevParams P = make_params(ctx);
    const CdParams C = make_cd(ctx);
    int cfg[12];
    graph_config(ctx, cfg);
    if (G.exec && (G.L != L || memcmp(&G.P0, &P, sizeof(P)) != 0 || memcmp(&G.C0, &C, sizeof(C)) != 0 ||
                   memcmp(G.cfg, cfg, sizeof(cfg)) != 0)) {
        cudaGraphExecDestroy(G.exec);
        G.exec = nullptr;
    }
    if (!G.exec) {
        const int slot0 = ctx->maxvel_slot, cur0 = ctx->cur;
        const uint64_t launches0 = ctx->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            ctx->use_graph = 0;
            return DEM_OK;
        }
        ctx->launches += launch_rebuild_kernels(ctx, nullptr);
        ctx->cur ^= 1;
        rc = DEM_OK;
        for (uint32_t i = 0; i < L && rc == DEM_OK; i++) rc = launch_step(ctx);
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        G.launches = ctx->launches - launches0;
        ctx->maxvel_slot = slot0;
        ctx->cur = cur0;
        ctx->launches = launches0;
        if (rc != DEM_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&G.exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            G.exec = nullptr;
            ctx->use_graph = 0;  // not worth retrying every cycle
            ctx->err.clear();
            return DEM_OK;
        }
        cudaGraphDestroy(graph);
        G.P0 = P;
        G.C0 = C;
        G.L = L;
        memcpy(G.cfg, cfg, sizeof(cfg));
    }
    tuner_mark(ctx);
    CK(cudaGraphLaunch(G.exec, ctx->stream));
    ctx->graph_launches++;
    ctx->launches += G.launches;
    note_rebuild_enqueued(ctx);
    ctx->maxvel_slot ^= (int)(L & 1u);
    note_steps_enqueued(ctx, L);
    *done = true;
    return DEM_OK;
}

// enqueue steps until n_steps has caught up with steps_target; never waits for the device except to confirm the
// rebuild before the one being enqueued
int pump(DemCtx* ctx) {
    while (ctx->n_steps < ctx->steps_target) {
        const uint32_t L = ctx->sp.cd_update_freq;
        if (ctx->need_rebuild || ctx->steps_since_rebuild >= ctx->list_freq) {
            if (graph_wanted(ctx) && !ctx->need_maxvel && L >= 2 && ctx->steps_target - ctx->n_steps >= L) {
                bool done = false;
                int rc = run_cycle_graph(ctx, &done);
                if (rc) return rc;
                if (done) continue;
            }
            const uint64_t before = ctx->seq_host;
            int rc = enqueue_rebuild(ctx);
            if (rc) return rc;
            if (ctx->seq_host == before) continue;  // rolled back instead: re-evaluate
        }
        int rc = launch_step(ctx);
        if (rc) return rc;
        note_steps_enqueued(ctx, 1);
    }
    CK(cudaGetLastError());
    return DEM_OK;
}

// wait for everything enqueued, confirm the last rebuild, replay what a failed rebuild dropped
int settle(DemCtx* ctx) {
    for (int attempt = 0; attempt < 16; attempt++) {
        CK(cudaStreamSynchronize(ctx->stream));
        bool rolled = false;
        int rc = confirm_rebuild(ctx, &rolled);
        if (rc) return rc;
        if (ctx->n_steps >= ctx->steps_target) return DEM_OK;
        rc = pump(ctx);
        if (rc) return rc;
    }
    return fail(ctx, DEM_ERR_CAPACITY, "contact list kept overflowing after repeated growth");
}

// ---- in-process GPU group behind ONE context (dem_ctx_create_group) ------------------------------------------------
std::vector<DemCtx*> group_ranks(DemCtx* ctx) {
    std::vector<DemCtx*> v{ctx};
    v.insert(v.end(), ctx->peers.begin(), ctx->peers.end());
    return v;
}
int peer_fail(DemCtx* ctx, DemCtx* peer, int rc) {
    ctx->err = "GPU " + std::to_string(peer->device) + ": " + peer->err;
    return rc;
}
// make rank 0 hold the state of all owners (after stepping, each rank only holds its slab + halo)
int ensure_merged(DemCtx* ctx) {
    if (!ctx->group_on || ctx->merged) return DEM_OK;
    std::vector<DemCtx*> v = group_ranks(ctx);
    int rc = dem_group_sync(v.data(), (int)v.size());
    if (rc) {
        for (DemCtx* c : v)
            if (c != ctx && !c->err.empty()) return peer_fail(ctx, c, rc);
        return rc;
    }
    rc = dem_group_gather(v.data(), (int)v.size());
    if (rc) return rc;
    ctx->merged = true;
    return DEM_OK;
}
// rank 0's owner arrays -> every rank, ownership re-derived from them (after the host changed owner state)
int group_broadcast_state(DemCtx* ctx) {
    if (!ctx->group_on) return DEM_OK;
    for (DemCtx* pc : ctx->peers) {
        CK(cudaMemcpyPeer(pc->d_state, pc->device, ctx->d_state, ctx->device, sizeof(OwnerState) * ctx->nOwners));
        CK(cudaMemcpyPeer(pc->d_spin, pc->device, ctx->d_spin, ctx->device, sizeof(float4) * ctx->nOwners));
    }
    for (DemCtx* c : group_ranks(ctx)) {
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(ctx, DEM_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
        const int rc = mg_reset_ownership(c);
        if (rc) return c == ctx ? rc : peer_fail(ctx, c, rc);
    }
    CK(cudaSetDevice(ctx->device));
    ctx->merged = true;
    return DEM_OK;
}
void free_mg(DemCtx* c);
// after dem_initialize of every rank: shard when the scene qualifies, otherwise the peers stay idle
int group_try_enable(DemCtx* ctx) {
    ctx->group_on = false;
    ctx->merged = true;
    if (ctx->peers.empty()) return DEM_OK;
    const size_t n = ctx->peers.size() + 1;
    if ((double)ctx->nClumpOwners < ctx->group_min_owners * (double)n) return DEM_OK;
    std::vector<DemCtx*> v = group_ranks(ctx);
    const int rc = dem_mgpu_init_local(v.data(), (int)v.size());
    if (rc == DEM_ERR_INVALID) {
        // (free-moving wall / mesh owners: the decomposition would not be exact -- stay on one GPU)
        for (DemCtx* c : v) {
            cudaSetDevice(c->device);
            free_mg(c);
            c->err.clear();
        }
        cudaSetDevice(ctx->device);
        return DEM_OK;
    }
    if (rc) return rc;
    ctx->group_on = true;
    return DEM_OK;
}
#define FORWARD_TO_PEERS(call)                                 \
    for (DemCtx* pc_ : ctx->peers) {                           \
        DemCtx* peer = pc_;                                    \
        const int rcp_ = (call);                               \
        if (rcp_) return peer_fail(ctx, peer, rcp_);           \
    }

}  // namespace

// ===============================================================================================================
extern "C" {

int dem_abi_version(void) { return DEM_B200_ABI_VERSION; }

int dem_host_figure_out_nv(const float box_min[3], const float box_max[3], uint32_t nv_p2[3], double* l,
                           double* voxel_size) {
    return dem_host_figure_out_nv_exact(box_min, box_max, -1, nv_p2, l, voxel_size);
}

int dem_host_figure_out_nv_exact(const float box_min[3], const float box_max[3], int exact_dir, uint32_t nv_p2[3], double* l,
                                 double* voxel_size) {
    // DEMSolver::figureOutNV, APIPrivate.cpp:373-487; exact_dir = -1 is m_box_dir_length_is_exact == NONE
    if (!box_min || !box_max || !nv_p2 || !l || !voxel_size || exact_dir < -1 || exact_dir > 2) return DEM_ERR_INVALID;
    float XYZ[3] = {box_max[0] - box_min[0], box_max[1] - box_min[1], box_max[2] - box_min[2]};
    int rank[3] = {0, 1, 2};
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (XYZ[i] > XYZ[j]) { std::swap(XYZ[i], XYZ[j]); std::swap(rank[i], rank[j]); }
    const float user321[3] = {XYZ[0], XYZ[1], XYZ[2]};
    int more[2] = {0, 0};
    while (XYZ[0] < XYZ[1]) {
        if (std::sqrt(2.) * XYZ[0] > XYZ[1]) break;
        more[0]++;
        XYZ[0] *= 2.;
    }
    while (XYZ[1] < XYZ[2]) {
        if (std::sqrt(2.) * XYZ[1] > XYZ[2]) break;
        more[1]++;
        XYZ[1] *= 2.;
    }
    const int total = 64 - 2 * more[0] - more[1];
    int b3 = total / 3, left = total % 3;
    int b2 = b3 + more[0], b1 = b2 + more[1];
    while (left > 0) {
        if (b3 < b2) b3++; else if (b2 < b1) b2++; else b1++;
        left--;
    }
    int bits[3] = {b3, b2, b1};
    if (exact_dir < 0) {
        const double l3 = (double)user321[0] / std::pow(2., 16) / std::pow(2., b3);
        const double l2 = (double)user321[1] / std::pow(2., 16) / std::pow(2., b2);
        const double l1 = (double)user321[2] / std::pow(2., 16) / std::pow(2., b1);
        *l = std::max(l3, std::max(l2, l1));
    } else {
        // the world spans the box EXACTLY along one axis (2^bits voxels of 2^16 l): l follows from that axis alone, and
        // the axis lends bits to the other two until they cover their lengths as well (:442-476)
        int e = 0;
        while (rank[e] != exact_dir) e++;
        const int others[2] = {e == 0 ? 1 : 0, e == 2 ? 1 : 2};
        auto unit = [&]() { return (double)user321[e] / std::pow(2., 16) / std::pow(2., bits[e]); };
        *l = unit();
        for (int k = 1; k >= 0; k--)
            while (*l * std::pow(2., 16) * std::pow(2., bits[others[k]]) < user321[others[k]]) {
                bits[e] -= 1;
                bits[others[k]] += 1;
                *l = unit();
            }
    }
    for (int p = 0; p < 3; p++) nv_p2[rank[p]] = (uint32_t)bits[p];
    *voxel_size = (double)((size_t)1 << 16) * (*l);
    return DEM_OK;
}

int dem_host_box_domain(float x, float y, float z, float user_min[3], float user_max[3], float target_min[3],
                        float target_max[3]) {
    // InstructBoxDomainDimension, APIPublic.cpp:845-872 (DEFAULT_BOX_DOMAIN_ENLARGE_RATIO = 0.2f, Defines.h:449)
    const float dims[3] = {x, y, z};
    const float ratio = 0.2f;
    for (int k = 0; k < 3; k++) {
        user_min[k] = (float)(-dims[k] / 2.);
        user_max[k] = (float)(dims[k] / 2.);
        const float enl = (float)(dims[k] * ratio / 2.);
        target_min[k] = user_min[k] - enl;
        target_max[k] = user_max[k] + enl;
    }
    return DEM_OK;
}

int dem_host_encode_positions(const DemSimParams* p, const float* xyz, uint64_t n, uint64_t* voxelID, uint16_t* locX,
                              uint16_t* locY, uint16_t* locZ) {
    if (!p || !xyz) return DEM_ERR_INVALID;
    for (uint64_t i = 0; i < n; i++) {
        double X[3];
        for (int k = 0; k < 3; k++) X[k] = (double)(xyz[3 * i + k] - p->LBF[k]);  // float subtraction, dT.cpp
        const uint64_t nx = (uint64_t)(X[0] / p->voxelSize), ny = (uint64_t)(X[1] / p->voxelSize),
                       nz = (uint64_t)(X[2] / p->voxelSize);
        locX[i] = (uint16_t)((X[0] - (double)nx * p->voxelSize) / p->l);
        locY[i] = (uint16_t)((X[1] - (double)ny * p->voxelSize) / p->l);
        locZ[i] = (uint16_t)((X[2] - (double)nz * p->voxelSize) / p->l);
        voxelID[i] = nx + (ny << p->nvXp2) + (nz << (p->nvXp2 + p->nvYp2));
    }
    return DEM_OK;
}

int dem_ctx_create(DemCtx** out, int device) {
    if (!out) return DEM_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return DEM_ERR_NO_GPU;
    if (device < 0 || device >= ndev) return DEM_ERR_INVALID;
    DemCtx* ctx = new DemCtx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return DEM_ERR_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->num_sms = prop.multiProcessorCount;
    cudaHostAlloc((void**)&ctx->h_pinned, 256 * sizeof(uint32_t), cudaHostAllocDefault);
    // status ring of the rebuilds: written by the device (k_finish_counts), polled by the host
    if (cudaHostAlloc((void**)&ctx->h_status, sizeof(RebuildStatus) * REBUILD_STATUS_SLOTS, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->d_status, ctx->h_status, 0) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return DEM_ERR_CUDA;
    }
    memset(ctx->h_status, 0, sizeof(RebuildStatus) * REBUILD_STATUS_SLOTS);
    if (const char* e = getenv("DEMB_CTAS_PER_SM")) ctx->ctas_per_sm = std::max(2, std::min(4, atoi(e)));
    if (const char* e = getenv("DEMB_FAST_MATH")) ctx->fast_math = atoi(e);
    for (int k = 0; k < 5; k++) cudaEventCreate(&ctx->ev[k]);
    cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&ctx->side2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming);
    *out = ctx;
    return DEM_OK;
}

int dem_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ndev;
}

int dem_ctx_create_group(DemCtx** out, const int* devices, int n) {
    if (!out || n < 1 || n > (int)MG_MAX_WORLD) return DEM_ERR_INVALID;
    *out = nullptr;
    DemCtx* first = nullptr;
    for (int r = 0; r < n; r++) {
        DemCtx* c = nullptr;
        const int rc = dem_ctx_create(&c, devices ? devices[r] : r);
        if (rc) {
            if (first) dem_ctx_destroy(first);
            return rc;
        }
        if (r == 0) first = c; else first->peers.push_back(c);
    }
    cudaSetDevice(first->device);
    *out = first;
    return DEM_OK;
}

int dem_ctx_destroy(DemCtx* ctx) {
    if (!ctx) return DEM_ERR_INVALID;
    if (ctx->group_on) {  // nothing may be left waiting for a peer that is about to disappear
        std::vector<DemCtx*> v = group_ranks(ctx);
        dem_group_sync(v.data(), (int)v.size());
    }
    for (DemCtx* pc : ctx->peers) dem_ctx_destroy(pc);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->mg.comm) g_nccl.CommDestroy(ctx->mg.comm);
    free_mg(ctx);
    graph_drop(ctx);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->side2) { cudaStreamSynchronize(ctx->side2); cudaStreamDestroy(ctx->side2); }
    if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    free_device(ctx);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    for (auto& e : ctx->tuner.ev)
        if (e) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) {
        if (ctx->h_famblob[k]) cudaFreeHost(ctx->h_famblob[k]);
        if (ctx->ev_fam[k]) cudaEventDestroy(ctx->ev_fam[k]);
    }
    for (int k = 0; k < 5; k++)
        if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DEM_OK;
}

const char* dem_last_error(const DemCtx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dem_set_stream(DemCtx* ctx, void* cuda_stream) {
    if (!ctx) return DEM_ERR_INVALID;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    graph_drop(ctx);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return DEM_OK;
}

int dem_set_params(DemCtx* ctx, const DemSimParams* p) {
    if (!ctx || !p) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(dem_set_params(peer, p))
    if (p->cd_update_freq < 1) return fail(ctx, DEM_ERR_INVALID, "cd_update_freq must be >= 1");
    if (!(p->l > 0) || !(p->h > 0)) return fail(ctx, DEM_ERR_INVALID, "l and h must be positive");
    if (p->force_model > DEM_HERTZIAN_FRICTIONLESS || p->integrator > DEM_EXTENDED_TAYLOR)
        return fail(ctx, DEM_ERR_INVALID, "unknown force model / integrator");
    if (ctx->initialized && (p->force_model != ctx->sp.force_model ||
                             p->record_contact_forces != ctx->sp.record_contact_forces))
        return fail(ctx, DEM_ERR_INVALID, "force model / force record cannot change after dem_initialize");
    // only what the contact margin / broad-phase grid depends on invalidates the current contact list
    const DemSimParams& o = ctx->sp;
    const bool list_stale = !ctx->params_set || p->h != o.h || p->beta != o.beta || p->approxMaxVel != o.approxMaxVel ||
                            p->expSafetyMulti != o.expSafetyMulti || p->expSafetyAdder != o.expSafetyAdder ||
                            p->cd_update_freq != o.cd_update_freq || p->l != o.l || p->nvXp2 != o.nvXp2 ||
                            p->nvYp2 != o.nvYp2 || memcmp(p->LBF, o.LBF, sizeof(o.LBF)) != 0 ||
                            memcmp(p->userBoxMin, o.userBoxMin, sizeof(o.userBoxMin)) != 0 ||
                            memcmp(p->userBoxMax, o.userBoxMax, sizeof(o.userBoxMax)) != 0;
    ctx->sp = *p;
    ctx->params_set = true;
    if (list_stale) ctx->need_rebuild = true;
    return DEM_OK;
}

int dem_upload_templates(DemCtx* ctx, uint32_t nComp, const float* radii, const float* relX, const float* relY,
                         const float* relZ, uint32_t nMassProps, const float* mass, const float* moiX,
                         const float* moiY, const float* moiZ) {
    if (ctx) { FORWARD_TO_PEERS(dem_upload_templates(peer, nComp, radii, relX, relY, relZ, nMassProps, mass, moiX, moiY, moiZ)) }
    if (!ctx || (nComp && (!radii || !relX || !relY || !relZ)) || (nMassProps && (!mass || !moiX || !moiY || !moiZ)))
        return DEM_ERR_INVALID;
    if (nComp > 65535) return fail(ctx, DEM_ERR_INVALID, "more than 65535 distinct clump components");
    ctx->h_comp.resize(nComp);
    ctx->rmax = 0.f;
    ctx->rclump = 0.f;
    for (uint32_t i = 0; i < nComp; i++) {
        ctx->h_comp[i] = make_float4(relX[i], relY[i], relZ[i], radii[i]);
        ctx->rmax = std::max(ctx->rmax, radii[i]);
        ctx->rclump = std::max(ctx->rclump, std::sqrt(relX[i] * relX[i] + relY[i] * relY[i] + relZ[i] * relZ[i]) + radii[i]);
    }
    ctx->h_massprop.resize(nMassProps);
    for (uint32_t i = 0; i < nMassProps; i++) ctx->h_massprop[i] = make_float4(mass[i], moiX[i], moiY[i], moiZ[i]);
    return DEM_OK;
}

int dem_upload_materials(DemCtx* ctx, uint32_t nMat, const float* E, const float* nu, const float* CoR,
                         const float* mu, const float* Crr) {
    if (ctx) { FORWARD_TO_PEERS(dem_upload_materials(peer, nMat, E, nu, CoR, mu, Crr)) }
    if (!ctx || !nMat || !E || !nu || !CoR || !mu || !Crr) return DEM_ERR_INVALID;
    if (nMat > 255) return fail(ctx, DEM_ERR_INVALID, "more than 255 materials");
    ctx->nMat = nMat;
    ctx->h_matpair.resize((size_t)nMat * nMat);
    for (uint32_t a = 0; a < nMat; a++)
        for (uint32_t b = 0; b < nMat; b++) {
            MatPair m;
            // matProxy2ContactParam<float>, DEMHelperKernels.cuh:433-444
            const float invE = (1.f - nu[a] * nu[a]) / E[a] + (1.f - nu[b] * nu[b]) / E[b];
            m.E_cnt = 1.f / invE;
            const float invG = 2.f * (2.f - nu[a]) * (1.f + nu[a]) / E[a] + 2.f * (2.f - nu[b]) * (1.f + nu[b]) / E[b];
            m.G_cnt = 1.f / invG;
            // FullHertzianForceModel.cu:59-60
            const float cor = CoR[a * nMat + b];
            const float loge = (float)((cor < 1e-12) ? std::log(1e-12) : (double)logf(cor));
            m.beta = (float)(loge / std::sqrt(loge * loge + 9.869604401089358));
            m.mu = mu[a * nMat + b];
            m.Crr = Crr[a * nMat + b];
            m.CoR = cor;
            m.pad0 = m.pad1 = 0.f;
            ctx->h_matpair[(size_t)a * nMat + b] = m;
        }
    return DEM_OK;
}

int dem_upload_analytical(DemCtx* ctx, uint32_t nAnal, const uint32_t* objOwner, const uint8_t* objType,
                          const uint16_t* objMaterial, const float* objNormal, const float* relPosX,
                          const float* relPosY, const float* relPosZ, const float* rotX, const float* rotY,
                          const float* rotZ, const float* size1, const float* size2, const float* size3,
                          const float* objMass) {
    if (!ctx) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(dem_upload_analytical(peer, nAnal, objOwner, objType, objMaterial, objNormal, relPosX, relPosY, relPosZ, rotX,
                                           rotY, rotZ, size1, size2, size3, objMass))
    if (nAnal > 255) return fail(ctx, DEM_ERR_INVALID, "more than 255 analytical components (objID_t is 8 bit)");
    ctx->h_anal.resize(nAnal);
    for (uint32_t i = 0; i < nAnal; i++) {
        AnalObj a;
        memset(&a, 0, sizeof(a));
        a.relx = relPosX[i]; a.rely = relPosY[i]; a.relz = relPosZ[i];
        a.rotx = rotX[i]; a.roty = rotY[i]; a.rotz = rotZ[i];
        a.size1 = size1[i]; a.size2 = size2[i]; a.size3 = size3[i];
        a.normal_sign = objNormal[i];
        a.mass = objMass[i];
        a.owner = objOwner[i]; a.type = objType[i]; a.material = objMaterial[i];
        if (a.type == DEM_ANAL_PLATE) return fail(ctx, DEM_ERR_INVALID, "plates are not supported (nor by the reference, DEMHelperKernels.cuh:491-493)");
        ctx->h_anal[i] = a;
    }
    ctx->nAnal = nAnal;
    return DEM_OK;
}

int dem_upload_families(DemCtx* ctx, const uint8_t* masks, const float* extraMargin, const DemPrescription* presc) {
    if (!ctx) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(dem_upload_families(peer, masks, extraMargin, presc))
    const std::vector<uint8_t> old_masks = ctx->h_masks;
    const std::vector<float> old_extra = ctx->h_extra;
    ctx->h_masks.assign(DEM_NUM_FAMILY_MASKS, 0);
    ctx->h_extra.assign(DEM_NUM_FAMILIES, 0.f);
    ctx->h_presc.resize(DEM_NUM_FAMILIES);
    memset(ctx->h_presc.data(), 0, sizeof(Prescr) * DEM_NUM_FAMILIES);
    if (masks) memcpy(ctx->h_masks.data(), masks, DEM_NUM_FAMILY_MASKS);
    if (extraMargin) memcpy(ctx->h_extra.data(), extraMargin, sizeof(float) * DEM_NUM_FAMILIES);
    if (presc) memcpy(ctx->h_presc.data(), presc, sizeof(Prescr) * DEM_NUM_FAMILIES);
    ctx->any_mask = 0;
    for (uint8_t m : ctx->h_masks) ctx->any_mask |= m;
    ctx->max_extra = 0.f;
    for (float e : ctx->h_extra) ctx->max_extra = std::max(ctx->max_extra, e);
    if (ctx->initialized) {
        cudaSetDevice(ctx->device);
        // stream-ordered: the tables change between the steps already enqueued and the ones enqueued after this call
        int rcf = stage_families(ctx);
        if (rcf) return rcf;
        // prescriptions act in the integrator only; the contact list depends on the masks and the extra margins
        if (old_masks != ctx->h_masks || old_extra != ctx->h_extra) ctx->need_rebuild = true;
    }
    return DEM_OK;
}

int dem_upload_owners(DemCtx* ctx, uint32_t nOwners, const uint64_t* voxelID, const uint16_t* locX,
                      const uint16_t* locY, const uint16_t* locZ, const float* oriQw, const float* oriQx,
                      const float* oriQy, const float* oriQz, const float* vX, const float* vY, const float* vZ,
                      const float* omgBarX, const float* omgBarY, const float* omgBarZ, const uint8_t* familyID,
                      const uint16_t* inertiaPropOffsets) {
    if (ctx) {
        FORWARD_TO_PEERS(dem_upload_owners(peer, nOwners, voxelID, locX, locY, locZ, oriQw, oriQx, oriQy, oriQz, vX, vY, vZ, omgBarX,
                                           omgBarY, omgBarZ, familyID, inertiaPropOffsets))
    }
    if (!ctx || (nOwners && (!voxelID || !locX || !locY || !locZ || !oriQw || !oriQx || !oriQy || !oriQz || !vX ||
                             !vY || !vZ || !omgBarX || !omgBarY || !omgBarZ || !familyID || !inertiaPropOffsets)))
        return DEM_ERR_INVALID;
    if (ctx->h_massprop.empty()) return fail(ctx, DEM_ERR_INVALID, "dem_upload_templates must precede dem_upload_owners");
    ctx->h_state.resize(nOwners);
    ctx->h_spin.resize(nOwners);
    for (uint32_t o = 0; o < nOwners; o++) {
        OwnerState s;
        s.pos.voxel = voxelID[o];
        s.pos.lx = locX[o]; s.pos.ly = locY[o]; s.pos.lz = locZ[o];
        s.pos.family = familyID[o];
        s.pos.flags = 0;
        s.quat = make_float4(oriQw[o], oriQx[o], oriQy[o], oriQz[o]);
        if (inertiaPropOffsets[o] >= ctx->h_massprop.size())
            return fail(ctx, DEM_ERR_INVALID, "owner %u refers to mass property %u (only %zu loaded)", o,
                        (unsigned)inertiaPropOffsets[o], ctx->h_massprop.size());
        s.vel = make_float4(vX[o], vY[o], vZ[o], ctx->h_massprop[inertiaPropOffsets[o]].x);
        uint32_t bits = inertiaPropOffsets[o];
        float fb;
        memcpy(&fb, &bits, 4);
        const float3 ww = host_rotate(make_float3(omgBarX[o], omgBarY[o], omgBarZ[o]), s.quat);
        s.omg = make_float4(ww.x, ww.y, ww.z, 0.f);
        ctx->h_state[o] = s;
        ctx->h_spin[o] = make_float4(omgBarX[o], omgBarY[o], omgBarZ[o], fb);
    }
    ctx->nOwners = nOwners;
    return DEM_OK;
}

int dem_upload_spheres(DemCtx* ctx, uint32_t nSpheres, const uint32_t* ownerClumpBody,
                       const uint16_t* clumpComponentOffset, const uint16_t* sphereMaterialOffset) {
    if (ctx) { FORWARD_TO_PEERS(dem_upload_spheres(peer, nSpheres, ownerClumpBody, clumpComponentOffset, sphereMaterialOffset)) }
    if (!ctx || (nSpheres && (!ownerClumpBody || !clumpComponentOffset || !sphereMaterialOffset))) return DEM_ERR_INVALID;
    ctx->h_sph.resize(nSpheres);
    uint32_t maxOwner = 0;
    for (uint32_t i = 0; i < nSpheres; i++) {
        if (clumpComponentOffset[i] >= ctx->h_comp.size())
            return fail(ctx, DEM_ERR_INVALID, "sphere %u refers to component %u (only %zu loaded)", i,
                        (unsigned)clumpComponentOffset[i], ctx->h_comp.size());
        if (ctx->nMat && sphereMaterialOffset[i] >= ctx->nMat)
            return fail(ctx, DEM_ERR_INVALID, "sphere %u refers to material %u (only %u loaded)", i,
                        (unsigned)sphereMaterialOffset[i], ctx->nMat);
        ctx->h_sph[i] = make_uint2(ownerClumpBody[i], (uint32_t)clumpComponentOffset[i] | ((uint32_t)sphereMaterialOffset[i] << 16));
        maxOwner = std::max(maxOwner, ownerClumpBody[i]);
    }
    ctx->nSpheres = nSpheres;
    ctx->nClumpOwners = nSpheres ? maxOwner + 1 : 0;
    return DEM_OK;
}

int dem_upload_triangles(DemCtx* ctx, uint32_t nTri, const uint32_t* ownerMesh, const float* node1, const float* node2,
                         const float* node3, const uint16_t* triMaterialOffset) {
    // the flattened m_mesh_facet_owner / m_mesh_facets / material arrays of dT::populateEntityArrays (dT.cpp:960-1010)
    if (ctx) { FORWARD_TO_PEERS(dem_upload_triangles(peer, nTri, ownerMesh, node1, node2, node3, triMaterialOffset)) }
    if (!ctx || (nTri && (!ownerMesh || !node1 || !node2 || !node3 || !triMaterialOffset))) return DEM_ERR_INVALID;
    ctx->h_tri1.resize(nTri); ctx->h_tri2.resize(nTri); ctx->h_tri3.resize(nTri); ctx->h_tri_info.resize(nTri);
    for (uint32_t t = 0; t < nTri; t++) {
        ctx->h_tri1[t] = make_float4(node1[3 * t], node1[3 * t + 1], node1[3 * t + 2], 0.f);
        ctx->h_tri2[t] = make_float4(node2[3 * t], node2[3 * t + 1], node2[3 * t + 2], 0.f);
        ctx->h_tri3[t] = make_float4(node3[3 * t], node3[3 * t + 1], node3[3 * t + 2], 0.f);
        ctx->h_tri_info[t] = make_uint2(ownerMesh[t], (uint32_t)triMaterialOffset[t]);
    }
    ctx->nTri = nTri;
    if (ctx->initialized) ctx->initialized = false;  // geometry changed: dem_initialize must run again
    return DEM_OK;
}

int dem_update_triangle_nodes(DemCtx* ctx, uint32_t first, uint32_t n, const float* node1, const float* node2,
                              const float* node3) {
    // SetTriNodeRelPos / UpdateTriNodeRelPos (API.h:489-491, dT.cpp:3135-3158): the facets' owner-frame node positions
    // change (a deforming mesh driven by a co-simulated FEA solver); the mesh owner's pose is untouched
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if (n == 0) return DEM_OK;
    if (!node1 || !node2 || !node3) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(dem_update_triangle_nodes(peer, first, n, node1, node2, node3))
    if ((uint64_t)first + n > ctx->nTri) return fail(ctx, DEM_ERR_INVALID, "triangle range out of bounds");
    CK(cudaSetDevice(ctx->device));
    std::vector<float4>* dst[3] = {&ctx->h_tri1, &ctx->h_tri2, &ctx->h_tri3};
    const float* src[3] = {node1, node2, node3};
    for (int k = 0; k < 3; k++) {
        for (uint32_t t = 0; t < n; t++)
            (*dst[k])[first + t] = make_float4(src[k][3 * t], src[k][3 * t + 1], src[k][3 * t + 2], 0.f);
        // stream-ordered: steps already enqueued see the old shape, steps enqueued after this call the new one
        CK(cudaMemcpyAsync(ctx->d_tri[k] + first, dst[k]->data() + first, sizeof(float4) * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));  // (the host vectors may be modified again right away)
    ctx->need_rebuild = true;                // sphere--facet candidates were found for the old shape
    return DEM_OK;
}

}  // extern "C"
namespace {
int initialize_one(DemCtx* ctx, uint64_t contact_capacity);
}
extern "C" int dem_initialize(DemCtx* ctx, uint64_t contact_capacity) {
    if (!ctx) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(initialize_one(peer, contact_capacity))
    int rc = initialize_one(ctx, contact_capacity);
    if (rc) return rc;
    return group_try_enable(ctx);
}
namespace {
int initialize_one(DemCtx* ctx, uint64_t contact_capacity) {
    if (!ctx->params_set) return fail(ctx, DEM_ERR_INVALID, "dem_set_params must precede dem_initialize");
    if (ctx->h_matpair.empty()) return fail(ctx, DEM_ERR_INVALID, "no materials uploaded");
    if (ctx->h_masks.empty()) {
        int rc = dem_upload_families(ctx, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    for (uint32_t i = 0; i < ctx->nSpheres; i++)
        if (ctx->h_sph[i].x >= ctx->nOwners) return fail(ctx, DEM_ERR_INVALID, "sphere %u refers to owner %u (only %u owners)", i, ctx->h_sph[i].x, ctx->nOwners);
    for (uint32_t i = 0; i < ctx->nAnal; i++)
        if (ctx->h_anal[i].owner >= ctx->nOwners || ctx->h_anal[i].material >= ctx->nMat)
            return fail(ctx, DEM_ERR_INVALID, "analytical component %u has a bad owner or material", i);
    for (uint32_t i = 0; i < ctx->nTri; i++)
        if (ctx->h_tri_info[i].x >= ctx->nOwners || ctx->h_tri_info[i].y >= ctx->nMat)
            return fail(ctx, DEM_ERR_INVALID, "triangle %u has a bad owner or material", i);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    free_mg(ctx);  // (a re-initialised context is a single-GPU context again)
    free_device(ctx);
    int rc;
    const uint32_t nO = ctx->nOwners, nS = ctx->nSpheres;
    if ((rc = dalloc(ctx, &ctx->d_state, nO))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_spin, nO))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_wrench, nO))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_acc, nO))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_sph, nS))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_comp, ctx->h_comp.size()))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_massprop, ctx->h_massprop.size()))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_matpair, ctx->h_matpair.size()))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_anal, ctx->h_anal.size()))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_famblob, FAM_BLOB_BYTES))) return rc;
    ctx->d_masks = reinterpret_cast<uint8_t*>(ctx->d_famblob + FAM_OFF_MASKS);
    ctx->d_extra = reinterpret_cast<float*>(ctx->d_famblob + FAM_OFF_EXTRA);
    ctx->d_presc = reinterpret_cast<Prescr*>(ctx->d_famblob + FAM_OFF_PRESC);
    if ((rc = dalloc(ctx, &ctx->d_flags, DEM_NUM_FLAGS))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_maxvel, 4))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_reduce, 4))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_reduce_many, 8))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_grid, 1))) return rc;
    CK(cudaMemcpy(ctx->d_state, ctx->h_state.data(), sizeof(OwnerState) * nO, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_spin, ctx->h_spin.data(), sizeof(float4) * nO, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctx->d_wrench, 0, sizeof(Wrench) * std::max<size_t>(nO, 1)));
    CK(cudaMemset(ctx->d_acc, 0, sizeof(Wrench) * std::max<size_t>(nO, 1)));
    CK(cudaMemcpy(ctx->d_sph, ctx->h_sph.data(), sizeof(uint2) * nS, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_comp, ctx->h_comp.data(), sizeof(float4) * ctx->h_comp.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_massprop, ctx->h_massprop.data(), sizeof(float4) * ctx->h_massprop.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_matpair, ctx->h_matpair.data(), sizeof(MatPair) * ctx->h_matpair.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_anal, ctx->h_anal.data(), sizeof(AnalObj) * ctx->h_anal.size(), cudaMemcpyHostToDevice));
    {
        int rcf = stage_families(ctx);
        if (rcf) return rcf;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaMemset(ctx->d_flags, 0, sizeof(uint32_t) * DEM_NUM_FLAGS));
    memset(ctx->h_status, 0, sizeof(RebuildStatus) * REBUILD_STATUS_SLOTS);
    ctx->seq_host = 0;
    ctx->pending = PendingRebuild();
    ctx->steps_target = ctx->n_steps;
    graph_drop(ctx);

    // broad-phase sizing: the smallest cell the device may ever pick bounds the cell table and the sort key width
    {
        const DemSimParams& s = ctx->sp;
        const float margin_min = (s.beta >= 0.f) ? s.beta : (float)((double)s.expSafetyAdder * s.h * s.cd_update_freq);
        const float cs_min = 2.f * (ctx->rmax + std::max(margin_min, 0.f)) * 1.0005f + 1e-30f;
        CdParams C = make_cd(ctx);
        double cells = 1.0;
        for (int k = 0; k < 3; k++) cells *= std::max(1.0, std::ceil((double)C.ext[k] / cs_min));
        const double cap = std::max(65536.0, std::min(67108864.0, std::max(4194304.0, 8.0 * nS)));
        ctx->max_cells = (uint32_t)std::min(cells, cap);
        ctx->key_bits = 1;
        while ((1ull << ctx->key_bits) < (unsigned long long)ctx->max_cells) ctx->key_bits++;
    }
    if ((rc = dalloc(ctx, &ctx->d_sphF, nS))) return rc;
    for (int k = 0; k < 2; k++) {
        if ((rc = dalloc(ctx, &ctx->d_keys[k], nS))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_vals[k], nS))) return rc;
    }
    if ((rc = dalloc(ctx, &ctx->d_cellStart, (size_t)ctx->max_cells + 2))) return rc;
    // (+2: the sweep stages these streams with 16-byte bulk copies that may start / end one entry outside a run)
    if ((rc = dalloc(ctx, &ctx->d_sortedSph, (size_t)nS + 2))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_sortedAux, (size_t)nS + 2))) return rc;
    CK(cudaMemset(ctx->d_sortedSph, 0, sizeof(float4) * ((size_t)nS + 2)));
    CK(cudaMemset(ctx->d_sortedAux, 0, sizeof(uint2) * ((size_t)nS + 2)));
    if ((rc = dalloc(ctx, &ctx->d_sortedMeta, nS))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_analw, ctx->h_anal.size()))) return rc;
    const size_t rs_blocks = ((size_t)nS + 4095) / 4096 + 1;
    if ((rc = dalloc(ctx, &ctx->d_rs_hist, 256 * rs_blocks))) return rc;
    const size_t scan_n = std::max<size_t>(std::max<size_t>((size_t)ctx->max_cells + 2, (size_t)nS + 2), 256 * rs_blocks);
    if ((rc = dalloc(ctx, &ctx->d_scan_tmp, scan_n / 4096 + 2))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_scan_desc, scan_n / 4096 + 8))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_cand, (size_t)nS * 12 + 12))) return rc;  // SW_REC words per sphere (kernels_sweep.cu)

    if (ctx->nTri) {
        const uint32_t nT = ctx->nTri;
        const std::vector<float4>* src[3] = {&ctx->h_tri1, &ctx->h_tri2, &ctx->h_tri3};
        for (int k = 0; k < 3; k++) {
            if ((rc = dalloc(ctx, &ctx->d_tri[k], nT))) return rc;
            if ((rc = dalloc(ctx, &ctx->d_triW[k], nT))) return rc;
            CK(cudaMemcpy(ctx->d_tri[k], src[k]->data(), sizeof(float4) * nT, cudaMemcpyHostToDevice));
        }
        if ((rc = dalloc(ctx, &ctx->d_tri_info, nT))) return rc;
        CK(cudaMemcpy(ctx->d_tri_info, ctx->h_tri_info.data(), sizeof(uint2) * nT, cudaMemcpyHostToDevice));
        if ((rc = dalloc(ctx, &ctx->d_triCellStart, (size_t)ctx->max_cells + 2))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_triCellFill, (size_t)ctx->max_cells + 2))) return rc;
        ctx->tri_pair_cap = (uint64_t)nT * 16 + 4096;
        if ((rc = dalloc(ctx, &ctx->d_triCellList, ctx->tri_pair_cap))) return rc;
    }

    uint64_t cap = contact_capacity ? contact_capacity : (uint64_t)nS * 8 + 1024;
    if ((rc = alloc_lists(ctx, cap))) return rc;
    ctx->cur = 0;

    // grid-stride force kernels: a whole number of CTAs per SM
    ctx->sa_grid = ctx->num_sms * 2;
    CK(cudaMemset(ctx->d_maxvel, 0, sizeof(float) * 4));
    ctx->maxvel_slot = 0;
    ctx->need_maxvel = true;
    ctx->initialized = true;
    ctx->need_rebuild = true;
    ctx->steps_since_rebuild = 0;
    for (auto& v : ctx->n_list) v = 0;
    return DEM_OK;
}
}  // namespace
extern "C" {

int dem_set_contacts(DemCtx* ctx, uint64_t n, const uint32_t* idA, const uint32_t* idB, const uint8_t* type,
                     const float* wildcards4) {
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if (n && (!idA || !idB || !type)) return DEM_ERR_INVALID;
    if (ctx->group_on) {  // every rank takes the whole list: a rank only ever looks up the spheres it holds
        int rcm = ensure_merged(ctx);
        if (rcm) return rcm;
        FORWARD_TO_PEERS(dem_set_contacts(peer, n, idA, idB, type, wildcards4))
    }
    CK(cudaSetDevice(ctx->device));
    { int rcs = settle(ctx); if (rcs) return rcs; }
    if (n > ctx->capacity) {
        int rc = grow_lists(ctx, n + n / 4 + 1024);
        if (rc) return rc;
    }
    // Build host-side "previous" lists grouped by sphere A so that the next rebuild carries the history over.
    const uint32_t nS = ctx->nSpheres;
    for (int which = 0; which < 3; which++) {
        std::vector<uint64_t> idx;
        for (uint64_t i = 0; i < n; i++) {
            const bool is_ss = type[i] == DEM_CNT_SPHERE_SPHERE;
            const bool is_sa = type[i] > 10;
            const bool is_st = type[i] == DEM_CNT_SPHERE_MESH;
            if ((which == 0 && is_ss) || (which == 1 && is_sa) || (which == 2 && is_st)) idx.push_back(i);
        }
        if (which == 2 && ctx->nTri == 0) break;
        std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return idA[a] < idA[b]; });
        std::vector<uint32_t> gb(idx.size());
        std::vector<uint4> cinfo(idx.size());
        std::vector<float4> hist(idx.size());
        std::vector<uint32_t> start(nS + 1, 0), count(nS + 1, 0);
        for (size_t k = 0; k < idx.size(); k++) {
            const uint64_t i = idx[k];
            if (idA[i] >= nS) return fail(ctx, DEM_ERR_INVALID, "contact %llu: bad geometry A", (unsigned long long)i);
            gb[k] = idB[i];
            float4 h = make_float4(0, 0, 0, 0);
            if (wildcards4) h = make_float4(wildcards4[4 * i], wildcards4[4 * i + 1], wildcards4[4 * i + 2], wildcards4[4 * i + 3]);
            hist[k] = h;
            const bool alive = (h.x != 0.f || h.y != 0.f || h.z != 0.f || h.w != 0.f);
            cinfo[k] = make_uint4(0, 0, 0, alive ? 0x80000000u : 0u);
            count[idA[i]]++;
        }
        for (uint32_t sp = 0, run = 0; sp < nS; sp++) { start[sp] = run; run += count[sp]; }
        ListBuf& L = (which == 0) ? ctx->lists[0][ctx->cur] : ctx->lists[which == 1 ? 2 : 3][ctx->cur];
        CK(cudaMemcpy(L.idB, gb.data(), sizeof(uint32_t) * idx.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.cinfo, cinfo.data(), sizeof(uint4) * idx.size(), cudaMemcpyHostToDevice));
        if (L.hist) CK(cudaMemcpy(L.hist, hist.data(), sizeof(float4) * idx.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.seg_start, start.data(), sizeof(uint32_t) * (nS + 1), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.seg_count, count.data(), sizeof(uint32_t) * (nS + 1), cudaMemcpyHostToDevice));
        // the old list is only a history source: its count is not used by force kernels until the rebuild swaps
        CK(cudaMemset(L.count, 0, sizeof(uint32_t) * 4));
    }
    // the sphere--sphere candidate list of the same buffer holds no restored contact: leave nothing stale behind for
    // the history search to match first
    {
        ListBuf& L = ctx->lists[1][ctx->cur];
        CK(cudaMemset(L.seg_count, 0, sizeof(uint32_t) * ((size_t)nS + 1)));
        CK(cudaMemset(L.count, 0, sizeof(uint32_t) * 4));
    }
    for (auto& v : ctx->n_list) v = 0;
    ctx->need_rebuild = true;
    return DEM_OK;
}

}  // extern "C"
namespace {
int rebuild_one(DemCtx* ctx) {
    CK(cudaSetDevice(ctx->device));
    int rc = settle(ctx);
    if (rc) return rc;
    return rebuild_blocking(ctx);
}
int step_async_one(DemCtx* ctx, uint64_t n_steps) {
    CK(cudaSetDevice(ctx->device));
    ctx->steps_target = ctx->n_steps + n_steps;
    return pump(ctx);
}
int sync_one(DemCtx* ctx) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->initialized) {
        CK(cudaStreamSynchronize(ctx->stream));
        return DEM_OK;
    }
    return settle(ctx);
}
int group_rc(DemCtx* ctx, const std::vector<DemCtx*>& v, int rc) {
    if (rc)
        for (DemCtx* c : v)
            if (c != ctx && !c->err.empty()) return peer_fail(ctx, c, rc);
    return rc;
}
}  // namespace
extern "C" {

int dem_rebuild_contacts(DemCtx* ctx) {
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if (ctx->group_on) {
        std::vector<DemCtx*> v = group_ranks(ctx);
        std::vector<int> rcs(v.size(), DEM_OK);
        std::vector<std::thread> th;
        for (size_t r = 0; r < v.size(); r++) th.emplace_back([&, r]() { rcs[r] = rebuild_one(v[r]); });
        for (auto& t : th) t.join();
        ctx->merged = false;
        for (size_t r = 0; r < v.size(); r++)
            if (rcs[r]) return r == 0 ? rcs[r] : peer_fail(ctx, v[r], rcs[r]);
        return DEM_OK;
    }
    return rebuild_one(ctx);
}

int dem_step_async(DemCtx* ctx, uint64_t n_steps) {
    if (!ctx || !ctx->initialized) return fail(ctx, DEM_ERR_INVALID, "dem_initialize has not been called");
    if (ctx->group_on) {
        std::vector<DemCtx*> v = group_ranks(ctx);
        ctx->merged = false;
        return group_rc(ctx, v, dem_group_step_async(v.data(), (int)v.size(), n_steps));
    }
    return step_async_one(ctx, n_steps);
}

int dem_sync(DemCtx* ctx) {
    if (!ctx) return DEM_ERR_INVALID;
    if (ctx->group_on) {
        std::vector<DemCtx*> v = group_ranks(ctx);
        return group_rc(ctx, v, dem_group_sync(v.data(), (int)v.size()));
    }
    return sync_one(ctx);
}

int dem_step(DemCtx* ctx, uint64_t n_steps) {
    int rc = dem_step_async(ctx, n_steps);
    if (rc) return rc;
    return dem_sync(ctx);
}

int dem_do_dynamics(DemCtx* ctx, double t) {
    if (!ctx || !ctx->initialized) return fail(ctx, DEM_ERR_INVALID, "dem_initialize has not been called");
    if (t <= 0.0) {
        // dry run: only (re)build the contact list (DoDynamicsThenSync(0), dT.cpp:2393-2398)
        return dem_rebuild_contacts(ctx);
    }
    // the reference's loop: for (double cycle = 0; cycle < t; cycle += (double)h) with the float-rounded h (dT.cpp:2401)
    uint64_t n = 0;
    const double h = (double)ctx->sp.h;
    for (double cycle = 0.0; cycle < t; cycle += h) n++;
    return dem_step(ctx, n);
}

int dem_set_sim_time(DemCtx* ctx, double t) {
    if (!ctx) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(dem_set_sim_time(peer, t))
    ctx->sim_time = t;  // host-side clock only (the kernels never see it): SetSimTime, dT.cpp:2709-2713
    return DEM_OK;
}

int dem_update_step_size(DemCtx* ctx, float h) {
    if (!ctx || !(h > 0)) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(dem_update_step_size(peer, h))
    ctx->sp.h = h;
    ctx->need_rebuild = true;  // margins depend on h
    return DEM_OK;
}

int dem_download_owner_state(DemCtx* ctx, uint32_t first, uint32_t n, uint64_t* voxelID, uint16_t* locX,
                             uint16_t* locY, uint16_t* locZ, float* oriQ, float* vel, float* omg, float* acc,
                             float* angacc, uint8_t* family) {
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if ((uint64_t)first + n > ctx->nOwners) return fail(ctx, DEM_ERR_INVALID, "owner range out of bounds");
    CK(cudaSetDevice(ctx->device));
    { int rcm = ensure_merged(ctx); if (rcm) return rcm; }
    { int rcs = settle(ctx); if (rcs) return rcs; }
    std::vector<OwnerState> st(n);
    CK(cudaMemcpy(st.data(), ctx->d_state + first, sizeof(OwnerState) * n, cudaMemcpyDeviceToHost));
    std::vector<float4> sp;
    if (omg) {
        sp.resize(n);
        CK(cudaMemcpy(sp.data(), ctx->d_spin + first, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    }
    std::vector<Wrench> ac;
    if (acc || angacc) {
        // the per-owner {a, alpha} read-out costs 32 B/owner/step, so it is only written once somebody asks: this
        // first query returns zeros, every later one the accelerations of the last step taken.
        ctx->keep_acc = true;
        ac.resize(n);
        CK(cudaMemcpy(ac.data(), ctx->d_acc + first, sizeof(Wrench) * n, cudaMemcpyDeviceToHost));
    }
    for (uint32_t i = 0; i < n; i++) {
        const OwnerState& s = st[i];
        if (voxelID) voxelID[i] = s.pos.voxel;
        if (locX) locX[i] = s.pos.lx;
        if (locY) locY[i] = s.pos.ly;
        if (locZ) locZ[i] = s.pos.lz;
        if (family) family[i] = s.pos.family;
        if (oriQ) { oriQ[4 * i] = s.quat.x; oriQ[4 * i + 1] = s.quat.y; oriQ[4 * i + 2] = s.quat.z; oriQ[4 * i + 3] = s.quat.w; }
        if (vel) { vel[3 * i] = s.vel.x; vel[3 * i + 1] = s.vel.y; vel[3 * i + 2] = s.vel.z; }
        if (omg) { omg[3 * i] = sp[i].x; omg[3 * i + 1] = sp[i].y; omg[3 * i + 2] = sp[i].z; }
        if (acc) { acc[3 * i] = ac[i].f.x; acc[3 * i + 1] = ac[i].f.y; acc[3 * i + 2] = ac[i].f.z; }
        if (angacc) { angacc[3 * i] = ac[i].t.x; angacc[3 * i + 1] = ac[i].t.y; angacc[3 * i + 2] = ac[i].t.z; }
    }
    return DEM_OK;
}

int dem_download_positions(DemCtx* ctx, uint32_t first, uint32_t n, float* xyz32, double* xyz64) {
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if ((uint64_t)first + n > ctx->nOwners) return fail(ctx, DEM_ERR_INVALID, "owner range out of bounds");
    CK(cudaSetDevice(ctx->device));
    { int rcm = ensure_merged(ctx); if (rcm) return rcm; }
    { int rcs = settle(ctx); if (rcs) return rcs; }
    std::vector<OwnerState> st(n);
    CK(cudaMemcpy(st.data(), ctx->d_state + first, sizeof(OwnerState) * n, cudaMemcpyDeviceToHost));
    const DemSimParams& p = ctx->sp;
    for (uint32_t i = 0; i < n; i++) {
        const OwnerPos& s = st[i].pos;
        const uint64_t vx = s.voxel & ((1ull << p.nvXp2) - 1ull);
        const uint64_t vy = (s.voxel >> p.nvXp2) & ((1ull << p.nvYp2) - 1ull);
        const uint64_t vz = s.voxel >> (p.nvXp2 + p.nvYp2);
        const double X[3] = {(double)vx * p.voxelSize + (double)s.lx * p.l, (double)vy * p.voxelSize + (double)s.ly * p.l,
                             (double)vz * p.voxelSize + (double)s.lz * p.l};
        for (int k = 0; k < 3; k++) {
            if (xyz64) xyz64[3 * i + k] = X[k] + (double)p.LBF[k];
            // the reference decodes in float and adds the float LBF (dT.cpp:3062-3076)
            if (xyz32) xyz32[3 * i + k] = (float)(X[k] + (double)p.LBF[k]);
        }
    }
    return DEM_OK;
}

int dem_upload_owner_state(DemCtx* ctx, uint32_t first, uint32_t n, const float* pos, const float* oriQ,
                           const float* vel, const float* omg, const uint8_t* family) {
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if ((uint64_t)first + n > ctx->nOwners) return fail(ctx, DEM_ERR_INVALID, "owner range out of bounds");
    CK(cudaSetDevice(ctx->device));
    { int rcm = ensure_merged(ctx); if (rcm) return rcm; }
    { int rcs = settle(ctx); if (rcs) return rcs; }
    std::vector<OwnerState> st(n);
    std::vector<float4> sp(n);
    CK(cudaMemcpy(st.data(), ctx->d_state + first, sizeof(OwnerState) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sp.data(), ctx->d_spin + first, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; i++) {
        OwnerState& s = st[i];
        if (pos) {
            uint64_t v;
            uint16_t lx, ly, lz;
            dem_host_encode_positions(&ctx->sp, pos + 3 * i, 1, &v, &lx, &ly, &lz);
            s.pos.voxel = v; s.pos.lx = lx; s.pos.ly = ly; s.pos.lz = lz;
        }
        if (oriQ) s.quat = make_float4(oriQ[4 * i], oriQ[4 * i + 1], oriQ[4 * i + 2], oriQ[4 * i + 3]);
        if (vel) { s.vel.x = vel[3 * i]; s.vel.y = vel[3 * i + 1]; s.vel.z = vel[3 * i + 2]; }
        if (omg) { sp[i].x = omg[3 * i]; sp[i].y = omg[3 * i + 1]; sp[i].z = omg[3 * i + 2]; }
        if (family) s.pos.family = family[i];
        const float3 ww = host_rotate(make_float3(sp[i].x, sp[i].y, sp[i].z), s.quat);
        s.omg = make_float4(ww.x, ww.y, ww.z, 0.f);
    }
    CK(cudaMemcpy(ctx->d_state + first, st.data(), sizeof(OwnerState) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_spin + first, sp.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
    ctx->need_rebuild = true;
    ctx->need_maxvel = true;
    return group_broadcast_state(ctx);  // (no-op on a single GPU)
}

int dem_add_owner_acc(DemCtx* ctx, uint32_t first, uint32_t n, const float* acc, const float* angacc_local) {
    // AddOwnerNextStepAcc / AddOwnerNextStepAngAcc (src/DEM/dT.cpp:3160-3174, DEMPrepForceKernels.cu:14-37): the reference
    // pre-loads a / alpha and skips zeroing them for one step, so the contacts of the next step add to them.  Here the
    // integrator consumes a world-frame wrench {sum F, sum T} per owner, (a, alpha) = (F / m, R^T T / I), and leaves it
    // zero: the same extra acceleration is m * a added to F and R (I * alpha) added to T, once, between two steps.
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if ((uint64_t)first + n > ctx->nOwners) return fail(ctx, DEM_ERR_INVALID, "owner range out of bounds");
    if (n == 0 || (!acc && !angacc_local)) return DEM_OK;
    CK(cudaSetDevice(ctx->device));
    { int rcm = ensure_merged(ctx); if (rcm) return rcm; }
    { int rcs = settle(ctx); if (rcs) return rcs; }
    std::vector<OwnerState> st(n);
    std::vector<float4> sp(n);
    CK(cudaMemcpy(st.data(), ctx->d_state + first, sizeof(OwnerState) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sp.data(), ctx->d_spin + first, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    std::vector<Wrench> add(n);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t mpi;
        std::memcpy(&mpi, &sp[i].w, sizeof(mpi));  // (the body-frame spin record carries the mass-property index in w)
        if (mpi >= ctx->h_massprop.size()) return fail(ctx, DEM_ERR_INVALID, "dem_add_owner_acc: owner %u has no mass properties", first + i);
        const float4 mp = ctx->h_massprop[mpi];  // mass, Ixx, Iyy, Izz
        add[i].f = make_float4(0.f, 0.f, 0.f, 0.f);
        add[i].t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (acc) add[i].f = make_float4(mp.x * acc[3 * i], mp.x * acc[3 * i + 1], mp.x * acc[3 * i + 2], 0.f);
        if (angacc_local) {
            const float3 tw = host_rotate(make_float3(mp.y * angacc_local[3 * i], mp.z * angacc_local[3 * i + 1],
                                                      mp.w * angacc_local[3 * i + 2]), st[i].quat);
            add[i].t = make_float4(tw.x, tw.y, tw.z, 0.f);
        }
    }
    // every rank of a group gets the increment: the rank that owns the body integrates it, the others consume it unused
    std::vector<Wrench> w(n);
    for (DemCtx* c : group_ranks(ctx)) {
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(ctx, DEM_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
        if (c != ctx) { int rcs = settle(c); if (rcs) return peer_fail(ctx, c, rcs); }
        if (cudaMemcpy(w.data(), c->d_wrench + first, sizeof(Wrench) * n, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(ctx, DEM_ERR_CUDA, "dem_add_owner_acc: download failed");
        for (uint32_t i = 0; i < n; i++) {
            w[i].f.x += add[i].f.x; w[i].f.y += add[i].f.y; w[i].f.z += add[i].f.z;
            w[i].t.x += add[i].t.x; w[i].t.y += add[i].t.y; w[i].t.z += add[i].t.z;
        }
        if (cudaMemcpy(c->d_wrench + first, w.data(), sizeof(Wrench) * n, cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(ctx, DEM_ERR_CUDA, "dem_add_owner_acc: upload failed");
    }
    CK(cudaSetDevice(ctx->device));
    return DEM_OK;
}

int dem_set_family_material(DemCtx* ctx, uint32_t family, uint32_t material, int meshes) {
    // SetFamilyClumpMaterial / SetFamilyMeshMaterial (src/DEM/APIPublic.cpp:1597-1604, dT.cpp:2719-2738): every sphere
    // (meshes != 0: every facet) whose owner is in `family` gets the material.  The compiled contact records carry the
    // material pair, so the contact list is rebuilt before the next step (history is kept).
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if (family > 255 || material >= ctx->nMat) return fail(ctx, DEM_ERR_INVALID, "dem_set_family_material: bad family or material");
    { int rcm = ensure_merged(ctx); if (rcm) return rcm; }
    CK(cudaSetDevice(ctx->device));
    { int rcs = settle(ctx); if (rcs) return rcs; }
    std::vector<OwnerState> st(ctx->nOwners);
    CK(cudaMemcpy(st.data(), ctx->d_state, sizeof(OwnerState) * ctx->nOwners, cudaMemcpyDeviceToHost));
    std::vector<DemCtx*> all = group_ranks(ctx);
    for (DemCtx* c : all) {
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(ctx, DEM_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
        if (c != ctx) { int rcs = settle(c); if (rcs) return peer_fail(ctx, c, rcs); }
        if (!meshes) {
            for (uint32_t i = 0; i < c->nSpheres; i++)
                if (st[c->h_sph[i].x].pos.family == family) c->h_sph[i].y = (c->h_sph[i].y & 0xffffu) | (material << 16);
            if (cudaMemcpy(c->d_sph, c->h_sph.data(), sizeof(uint2) * c->nSpheres, cudaMemcpyHostToDevice) != cudaSuccess)
                return fail(ctx, DEM_ERR_CUDA, "dem_set_family_material: upload failed");
        } else if (c->nTri) {
            for (uint32_t t = 0; t < c->nTri; t++)
                if (st[c->h_tri_info[t].x].pos.family == family) c->h_tri_info[t].y = material;
            if (cudaMemcpy(c->d_tri_info, c->h_tri_info.data(), sizeof(uint2) * c->nTri, cudaMemcpyHostToDevice) != cudaSuccess)
                return fail(ctx, DEM_ERR_CUDA, "dem_set_family_material: upload failed");
        }
        c->need_rebuild = true;
    }
    CK(cudaSetDevice(ctx->device));
    return DEM_OK;
}

int dem_download_contacts(DemCtx* ctx, uint64_t capacity, uint64_t* n_out, uint32_t* idA, uint32_t* idB, uint8_t* type,
                          float* wildcards4, float* force_xyz) {
    return dem_download_contact_records(ctx, capacity, n_out, idA, idB, type, wildcards4, force_xyz, nullptr);
}

}  // extern "C"
namespace {
struct ContactRow { uint32_t a, b; uint8_t t; float4 h; float4 f; float4 p; };
// the rows of one context's current lists, normalised to the reference's conventions (smaller sphere id = geometry A)
int collect_contact_rows(DemCtx* ctx, bool want_points, std::vector<ContactRow>& rows) {
    CK(cudaSetDevice(ctx->device));
    { int rcs = settle(ctx); if (rcs) return rcs; }
    for (int kind = 0; kind < 4; kind++) {
        const ListBuf& L = ctx->lists[kind][ctx->cur];
        const uint64_t m = ctx->n_list[kind];
        const int which = (kind == 2) ? 1 : (kind == 3 ? 2 : 0);
        // geometry A of a contact is the sphere whose segment it lies in
        std::vector<uint2> pair(m, make_uint2(0xffffffffu, 0u));
        {
            std::vector<uint32_t> gb(m), ss(ctx->nSpheres), sc(ctx->nSpheres);
            CK(cudaMemcpy(gb.data(), L.idB, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ss.data(), L.seg_start, sizeof(uint32_t) * ctx->nSpheres, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(sc.data(), L.seg_count, sizeof(uint32_t) * ctx->nSpheres, cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < m; i++) pair[i].y = gb[i];
            for (uint32_t sp = 0; sp < ctx->nSpheres; sp++)
                for (uint32_t k = 0; k < sc[sp] && (uint64_t)ss[sp] + k < m; k++) pair[(size_t)ss[sp] + k].x = sp;
        }
        std::vector<float4> hist(m, make_float4(0, 0, 0, 0)), frc(m, make_float4(0, 0, 0, 0)), cpt(m, make_float4(0, 0, 0, 0));
        if (L.hist) CK(cudaMemcpy(hist.data(), L.hist, sizeof(float4) * m, cudaMemcpyDeviceToHost));
        if (L.force) CK(cudaMemcpy(frc.data(), L.force, sizeof(float4) * m, cudaMemcpyDeviceToHost));
        if (L.cpoint && want_points) CK(cudaMemcpy(cpt.data(), L.cpoint, sizeof(float4) * m, cudaMemcpyDeviceToHost));
        std::vector<uint4> ci(m);
        CK(cudaMemcpy(ci.data(), L.cinfo, sizeof(uint4) * m, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < m; i++) {
            if (pair[i].x == 0xffffffffu) continue;  // (a slot no segment points at: decomposed lists of another cycle)
            ContactRow r;
            r.a = pair[i].x; r.b = pair[i].y;
            bool flip = false;
            if (which == 0 && r.a > r.b) { std::swap(r.a, r.b); flip = true; }
            if (which == 0) r.t = DEM_CNT_SPHERE_SPHERE;
            else if (which == 2) r.t = DEM_CNT_SPHERE_MESH;
            else r.t = (ctx->h_anal[pair[i].y].type == DEM_ANAL_PLANE) ? DEM_CNT_SPHERE_PLANE : DEM_CNT_SPHERE_CYL;
            // history words of contacts that are not alive are stale by construction: report zeros
            r.h = (ci[i].w & 0x80000000u) ? hist[i] : make_float4(0, 0, 0, 0);
            r.f = frc[i];
            r.p = cpt[i];
            if (flip) {
                // the device lists a pair with the sphere that comes first in cell order as A; the reference reports
                // the smaller sphere id as A. Swapping roles negates delta_tan and the force that "A feels".
                r.h.x = -r.h.x; r.h.y = -r.h.y; r.h.z = -r.h.z;
                r.f.x = -r.f.x; r.f.y = -r.f.y; r.f.z = -r.f.z;
            }
            rows.push_back(r);
        }
    }
    return DEM_OK;
}
}  // namespace
extern "C" {

int dem_download_contact_records(DemCtx* ctx, uint64_t capacity, uint64_t* n_out, uint32_t* idA, uint32_t* idB,
                                 uint8_t* type, float* wildcards4, float* force_xyz, float* point_xyz) {
    if (!ctx || !ctx->initialized || !n_out) return DEM_ERR_INVALID;
    std::vector<ContactRow> rows;
    if (!ctx->group_on && !idA && !idB && !type && !wildcards4 && !force_xyz && !point_xyz) {  // only the count is asked for
        CK(cudaSetDevice(ctx->device));
        { int rcs = settle(ctx); if (rcs) return rcs; }
        *n_out = ctx->n_list[0] + ctx->n_list[1] + ctx->n_list[2] + ctx->n_list[3];
        return DEM_OK;
    }
    if (ctx->group_on) {
        // every rank lists the contacts of the spheres it holds; pairs across a cut appear on both sides, with
        // identical history and force (same inputs, same arithmetic): keep one
        std::vector<DemCtx*> v = group_ranks(ctx);
        int rc = group_rc(ctx, v, dem_group_sync(v.data(), (int)v.size()));
        if (rc) return rc;
        for (DemCtx* c : v) {
            rc = collect_contact_rows(c, point_xyz != nullptr, rows);
            if (rc) return c == ctx ? rc : peer_fail(ctx, c, rc);
        }
        CK(cudaSetDevice(ctx->device));
    } else {
        int rc = collect_contact_rows(ctx, point_xyz != nullptr, rows);
        if (rc) return rc;
    }
    std::sort(rows.begin(), rows.end(), [](const ContactRow& x, const ContactRow& y) {
        if (x.t != y.t) return x.t < y.t;
        if (x.a != y.a) return x.a < y.a;
        return x.b < y.b;
    });
    if (ctx->group_on)
        rows.erase(std::unique(rows.begin(), rows.end(), [](const ContactRow& x, const ContactRow& y) {
                       return x.t == y.t && x.a == y.a && x.b == y.b;
                   }), rows.end());
    const uint64_t n = rows.size();
    *n_out = n;
    if (!idA && !idB && !type && !wildcards4 && !force_xyz && !point_xyz) return DEM_OK;
    if (capacity < n) return fail(ctx, DEM_ERR_CAPACITY, "dem_download_contacts: need room for %llu contacts", (unsigned long long)n);
    for (uint64_t i = 0; i < n; i++) {
        if (idA) idA[i] = rows[i].a;
        if (idB) idB[i] = rows[i].b;
        if (type) type[i] = rows[i].t;
        if (wildcards4) { wildcards4[4 * i] = rows[i].h.x; wildcards4[4 * i + 1] = rows[i].h.y; wildcards4[4 * i + 2] = rows[i].h.z; wildcards4[4 * i + 3] = rows[i].h.w; }
        if (force_xyz) { force_xyz[3 * i] = rows[i].f.x; force_xyz[3 * i + 1] = rows[i].f.y; force_xyz[3 * i + 2] = rows[i].f.z; }
        if (point_xyz) {  // world frame: the record is LBF-relative
            point_xyz[3 * i] = rows[i].p.x + ctx->sp.LBF[0];
            point_xyz[3 * i + 1] = rows[i].p.y + ctx->sp.LBF[1];
            point_xyz[3 * i + 2] = rows[i].p.z + ctx->sp.LBF[2];
        }
    }
    return DEM_OK;
}

int dem_get_stats(DemCtx* ctx, DemStats* out) {
    if (!ctx || !out) return DEM_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->n_steps = ctx->n_steps; out->n_rebuilds = ctx->n_rebuilds;
    out->n_contacts_ss = ctx->n_list[0] + ctx->n_list[1]; out->n_contacts_sa = ctx->n_list[2]; out->n_contacts_st = ctx->n_list[3];
    out->n_contacts_ss_touching = ctx->n_list[0];
    out->contact_capacity = ctx->capacity; out->kernel_launches = ctx->launches; out->device_bytes = ctx->device_bytes;
    out->sim_time = ctx->sim_time; out->max_margin = ctx->last_grid.max_margin; out->cell_size = ctx->last_grid.cs;
    out->n_cells[0] = ctx->last_grid.nbx; out->n_cells[1] = ctx->last_grid.nby; out->n_cells[2] = ctx->last_grid.nbz;
    out->overflow = ctx->overflow_seen;
    out->cd_update_freq = ctx->sp.cd_update_freq;
    if (ctx->group_on)  // (contacts across a cut are listed on both sides: the sums count them twice)
        for (DemCtx* pc : ctx->peers) {
            out->n_contacts_ss += pc->n_list[0] + pc->n_list[1]; out->n_contacts_sa += pc->n_list[2]; out->n_contacts_st += pc->n_list[3];
            out->n_contacts_ss_touching += pc->n_list[0];
            out->kernel_launches += pc->launches; out->device_bytes += pc->device_bytes; out->overflow += pc->overflow_seen;
        }
    return DEM_OK;
}

int dem_reduce(DemCtx* ctx, int kind, double* out) {
    if (!ctx || !ctx->initialized || !out) return DEM_ERR_INVALID;
    if (kind < DEM_REDUCE_MAX_ABSV || kind > DEM_REDUCE_SPHERE_MAX_ABSV) return fail(ctx, DEM_ERR_INVALID, "unknown reduction");
    { int rcm = ensure_merged(ctx); if (rcm) return rcm; }
    CK(cudaSetDevice(ctx->device));
    const bool per_sphere = kind >= DEM_REDUCE_SPHERE_MAX_Z;
    const double init = (kind == DEM_REDUCE_MIN_Z || kind == DEM_REDUCE_SPHERE_MIN_Z) ? 1e300
                        : ((kind == DEM_REDUCE_MAX_Z || kind == DEM_REDUCE_SPHERE_MAX_Z) ? -1e300 : 0.0);
    CK(cudaMemcpyAsync(ctx->d_reduce, &init, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DevParams P = make_params(ctx);
    P.nOwners = ctx->nClumpOwners;  // inspectors of the reference look at clumps only
    if (ctx->group_on) P.active = nullptr;  // rank 0 holds the merged state of all ranks
    ctx->launches += per_sphere ? launch_reduce_spheres(P, kind, ctx->d_reduce, ctx->stream)
                                : launch_reduce(P, kind, ctx->d_reduce, ctx->stream);
    CK(cudaMemcpyAsync(out, ctx->d_reduce, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DEM_OK;
}

int dem_reduce_many(DemCtx* ctx, uint32_t kind_mask, double out[5]) {
    if (!ctx || !ctx->initialized || !out) return DEM_ERR_INVALID;
    if (kind_mask == 0 || kind_mask >= (1u << 5)) return fail(ctx, DEM_ERR_INVALID, "unknown reduction in mask");
    { int rcm = ensure_merged(ctx); if (rcm) return rcm; }
    CK(cudaSetDevice(ctx->device));
    double* hp = reinterpret_cast<double*>(ctx->h_pinned + 112);  // pinned scratch: words 112..131
    const double init[5] = {0.0, -1e300, 1e300, 0.0, 0.0};
    memcpy(hp, init, sizeof(init));
    CK(cudaMemcpyAsync(ctx->d_reduce_many, hp, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    DevParams P = make_params(ctx);
    P.nOwners = ctx->nClumpOwners;
    if (ctx->group_on) P.active = nullptr;
    ctx->launches += launch_reduce_many(P, kind_mask, ctx->d_reduce_many, ctx->stream);
    CK(cudaMemcpyAsync(hp + 5, ctx->d_reduce_many, sizeof(init), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 5; k++)
        if (kind_mask & (1u << k)) out[k] = hp[5 + k];
    return DEM_OK;
}

int dem_mgpu_unique_id(uint8_t out[128]) {
    if (!out) return DEM_ERR_INVALID;
    if (!g_nccl.load()) return DEM_ERR_INVALID;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return DEM_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out, &id, 128);
    return DEM_OK;
}

int dem_host_slab_bounds(const DemSimParams* p, int world, int rank, float* lo, float* hi) {
    if (!p || world < 1 || rank < 0 || rank >= world || !lo || !hi) return DEM_ERR_INVALID;
    // equal-width slabs of the user's box along x, in LBF-relative coordinates; the end slabs extend to infinity
    const double x0 = (double)p->userBoxMin[0] - (double)p->LBF[0], x1 = (double)p->userBoxMax[0] - (double)p->LBF[0];
    const double w = (x1 - x0) / world;
    *lo = (rank == 0) ? -3.0e38f : (float)(x0 + w * rank);
    *hi = (rank == world - 1) ? 3.0e38f : (float)(x0 + w * (rank + 1));
    return DEM_OK;
}

int dem_host_partition_owners(const DemSimParams* p, int world, int rank, float halo, uint64_t n, const uint64_t* voxelID,
                              const uint16_t* locX, uint8_t* role, uint8_t* send) {
    if (!p || world < 1 || rank < 0 || rank >= world || (n && (!voxelID || !locX || !role))) return DEM_ERR_INVALID;
    float lo, hi;
    dem_host_slab_bounds(p, world, rank, &lo, &hi);
    const uint64_t xmask = (p->nvXp2 >= 64) ? ~0ull : ((1ull << p->nvXp2) - 1ull);
    for (uint64_t i = 0; i < n; i++) {
        // same arithmetic as the device (pos_decode, then float compare against the cuts: k_mg_classify)
        const double X = (double)(voxelID[i] & xmask) * p->voxelSize + (double)locX[i] * p->l;
        const float x = (float)X;
        const bool own = (x >= lo) && (x < hi);
        uint8_t r = 0, sd = 0;
        if (own) {
            r = 1;
            if (rank > 0 && x < lo + halo) sd |= 1;
            if (rank < world - 1 && x >= hi - halo) sd |= 2;
        } else if ((rank > 0 && x < lo && x >= lo - halo) || (rank < world - 1 && x >= hi && x < hi + halo)) {
            r = 2;
        }
        role[i] = r;
        if (send) send[i] = sd;
    }
    return DEM_OK;
}

} // extern "C"

namespace {

// (Re-)derive who owns what from the positions this rank's device arrays hold: own inside my slab, unknown elsewhere (the
// next rebuild brings the ghosts in); replicated analytical / mesh owners are owned everywhere.  The "previous cycle's
// active list" that rebuild walks is the list of these own owners; halo send lists start empty.  The per-parity sphere
// lists (act_sph, counts[.][4]) are left alone: they describe what the contact-list buffers of that parity hold.
int mg_reset_ownership(DemCtx* ctx) {
    MgState& g = ctx->mg;
    const uint32_t nO = ctx->nOwners;
    std::vector<uint8_t> flag(nO, 0);
    std::vector<uint32_t> act;
    const DemSimParams& p = ctx->sp;
    // (what the device holds now, not what was uploaded: the context may have stepped before)
    std::vector<OwnerState> st(nO);
    CK(cudaMemcpy(st.data(), ctx->d_state, sizeof(OwnerState) * nO, cudaMemcpyDeviceToHost));
    for (uint32_t o = 0; o < nO; o++) {
        bool own = true;
        if (o < ctx->nClumpOwners) {
            const OwnerPos& s = st[o].pos;
            const uint64_t vx = s.voxel & ((1ull << p.nvXp2) - 1ull);
            const float x = (float)((double)vx * p.voxelSize + (double)s.lx * p.l);
            own = (x >= g.cut_lo && x < g.cut_hi);
        }
        flag[o] = own ? 1 : 0;
        if (own) act.push_back(o);
    }
    CK(cudaMemcpy(g.d_flag, flag.data(), nO, cudaMemcpyHostToDevice));
    const int prev = ctx->cur;  // the next rebuild builds parity cur^1 and walks the list of parity cur
    CK(cudaMemcpy(g.d_active_list[prev], act.data(), sizeof(uint32_t) * act.size(), cudaMemcpyHostToDevice));
    uint32_t cnt[2][8];
    CK(cudaMemcpy(cnt, g.d_counts, sizeof(cnt), cudaMemcpyDeviceToHost));
    for (int q = 0; q < 2; q++)
        for (int k = 0; k < 4; k++) cnt[q][k] = 0;
    cnt[prev][3] = (uint32_t)act.size();
    CK(cudaMemcpy(g.d_counts, cnt, sizeof(cnt), cudaMemcpyHostToDevice));
    ctx->need_rebuild = true;
    ctx->need_maxvel = true;
    return DEM_OK;
}

// Everything of the decomposition that does not depend on how the peers' blocks get mapped: slab, buffers, the initial
// ownership.  Leaves g.my_block allocated and zeroed; the caller fills g.peer_block[] and switches g.on.
int mg_prepare(DemCtx* ctx, int rank, int world) {
    if (world > (int)MG_MAX_WORLD) return fail(ctx, DEM_ERR_INVALID, "at most %u ranks (GPUs of one box)", MG_MAX_WORLD);
    // Wall and mesh owners are replicated on every rank and feel only the contacts of that rank's slab: exact as long as
    // their motion does not depend on the forces on them, i.e. while they are fixed or fully prescribed.
    for (uint32_t o = ctx->nClumpOwners; o < ctx->nOwners; o++) {
        const Prescr& pr = ctx->h_presc[ctx->h_state[o].pos.family];
        bool ok = pr.used != 0;
        for (int k = 0; k < 3; k++) ok = ok && (pr.linVelPrescribed[k] || pr.linPosPrescribed[k]) && (pr.rotVelPrescribed[k] || pr.rotPosPrescribed);
        if (!ok)
            return fail(ctx, DEM_ERR_INVALID, "owner %u (an analytical object or a mesh, family %u) moves freely: on several "
                        "GPUs such owners must be fixed or follow a fully prescribed motion", o, (unsigned)ctx->h_state[o].pos.family);
    }
    CK(cudaSetDevice(ctx->device));
    { int rcs = settle(ctx); if (rcs) return rcs; }
    free_mg(ctx);
    MgState& g = ctx->mg;
    g.rank = rank; g.world = world;
    dem_host_slab_bounds(&ctx->sp, world, rank, &g.cut_lo, &g.cut_hi);
    const uint32_t nO = ctx->nOwners, nS = ctx->nSpheres;
    // halo buffers: a quarter of the owners per side is far more than a slab's boundary layer ever holds
    g.cap = std::max<uint32_t>(4096u, nO / (world > 2 ? 2u : 4u));
    int rc;
    if ((rc = dalloc(ctx, &g.d_flag, nO))) return rc;
    for (int p = 0; p < 2; p++) {
        if ((rc = dalloc(ctx, &g.d_active_list[p], nO))) return rc;
        for (int d = 0; d < 2; d++)
            if ((rc = dalloc(ctx, &g.d_send_gid[p][d], g.cap))) return rc;
    }
    for (int p = 0; p < 2; p++)
        if ((rc = dalloc(ctx, &g.d_act_sph[p], std::max<uint32_t>(nS, 1u)))) return rc;
    if ((rc = dalloc(ctx, &g.d_counts, 16))) return rc;
    if ((rc = dalloc(ctx, &g.d_ctrs, 4))) return rc;
    CK(cudaMemset(g.d_counts, 0, sizeof(uint32_t) * 16));
    CK(cudaMemset(g.d_ctrs, 0, sizeof(unsigned long long) * 4));
    // per owner {first sphere, number of spheres} when the spheres of every owner are contiguous (they are as the
    // reference flattens clumps, dT.cpp:638-1024): the rebuild then lists the active spheres without a pass over all
    {
        std::vector<uint2> os(nO, make_uint2(0u, 0u));
        bool contiguous = true;
        for (uint32_t i = 0; i < nS && contiguous; i++) {
            const uint32_t o = ctx->h_sph[i].x;
            if (os[o].y == 0) os[o].x = i;
            else if (os[o].x + os[o].y != i) contiguous = false;
            os[o].y++;
        }
        if (contiguous) {
            if ((rc = dalloc(ctx, &g.d_owner_sph, nO))) return rc;
            CK(cudaMemcpy(g.d_owner_sph, os.data(), sizeof(uint2) * nO, cudaMemcpyHostToDevice));
        }
    }
    if ((rc = mg_reset_ownership(ctx))) return rc;
    const size_t block_bytes = mg_block_bytes(g.cap);
    CK(cudaMalloc((void**)&g.my_block, block_bytes));
    CK(cudaMemset(g.my_block, 0, block_bytes));
    ctx->device_bytes += block_bytes;
    return DEM_OK;
}

void mg_switch_on(DemCtx* ctx) {
    ctx->sort_mode = 1;  // the radix path has no notion of inactive spheres
    ctx->mg.on = true;
    ctx->need_rebuild = true;
    ctx->need_maxvel = true;
    graph_drop(ctx);
}

}  // namespace

extern "C" {

int dem_mgpu_init(DemCtx* ctx, int rank, int world, const uint8_t unique_id[128]) {
    if (!ctx || !ctx->initialized || !unique_id || world < 1 || rank < 0 || rank >= world) return DEM_ERR_INVALID;
    if (world == 1) return DEM_OK;
    if (!g_nccl.load()) return fail(ctx, DEM_ERR_INVALID, "NCCL (libnccl.so.2) could not be loaded");
    int rc = mg_prepare(ctx, rank, world);
    if (rc) return rc;
    MgState& g = ctx->mg;
    g.local = false;
    // bootstrap: NCCL carries the cudaIpc handles of the peer blocks once; nothing after this goes through a library
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    NC(g_nccl.CommInitRank(&g.comm, world, id, rank));
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, g.my_block));
    const size_t rec = sizeof(cudaIpcMemHandle_t);
    std::vector<char> h_all(rec * world);
    char *d_me = nullptr, *d_all = nullptr;
    CK(cudaMalloc((void**)&d_me, rec));
    CK(cudaMalloc((void**)&d_all, rec * world));
    CK(cudaMemcpy(d_me, &mine, rec, cudaMemcpyHostToDevice));
    NC(g_nccl.AllGather(d_me, d_all, rec, ncclUint8, g.comm, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(h_all.data(), d_all, rec * world, cudaMemcpyDeviceToHost));
    cudaFree(d_me); cudaFree(d_all);
    for (int r = 0; r < world; r++) {
        if (r == rank) { g.peer_block[r] = g.my_block; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, h_all.data() + rec * r, sizeof(h));
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, DEM_ERR_CUDA, "rank %d cannot map the exchange block of rank %d (%s): the decomposition needs "
                        "peer access between the GPUs (NVLink / NVSwitch)", rank, r, cudaGetErrorString(e));
        }
        g.peer_block[r] = (char*)ptr;
        g.peer_ipc[r] = true;
    }
    mg_switch_on(ctx);
    return DEM_OK;
}

/* All ranks are contexts of THIS process, one per GPU: the blocks are mapped with cudaDeviceEnablePeerAccess.  Contexts
 * of a local group must be stepped together through dem_group_step / dem_group_sync (their kernels wait for each other
 * on the device). */
int dem_mgpu_init_local(DemCtx** ctxs, int world) {
    if (!ctxs || world < 1) return DEM_ERR_INVALID;
    if (world == 1) return DEM_OK;
    for (int r = 0; r < world; r++)
        if (!ctxs[r] || !ctxs[r]->initialized) return DEM_ERR_INVALID;
    for (int r = 0; r < world; r++) {
        DemCtx* ctx = ctxs[r];
        int rc = mg_prepare(ctx, r, world);
        if (rc) return rc;
        ctx->mg.local = true;
        for (int q = 0; q < world; q++) {
            if (q == r || ctxs[q]->device == ctx->device) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[q]->device));
            if (!can) return fail(ctx, DEM_ERR_CUDA, "GPU %d cannot access GPU %d's memory", ctx->device, ctxs[q]->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(ctx, DEM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctx->device, ctxs[q]->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    for (int r = 0; r < world; r++) {
        for (int q = 0; q < world; q++) ctxs[r]->mg.peer_block[q] = ctxs[q]->mg.my_block;
        mg_switch_on(ctxs[r]);
    }
    return DEM_OK;
}

/* n steps on every context of a local group.  One host thread per rank for the duration of the call: the ranks' kernels
 * wait for each other on the device, so no rank's host-side wait (a rebuild confirmation, a first-use module load, an
 * allocation while growing a list) may keep another rank's work from being enqueued. */
int dem_group_step_async(DemCtx** ctxs, int world, uint64_t n_steps) {
    if (!ctxs || world < 1) return DEM_ERR_INVALID;
    for (int r = 0; r < world; r++)
        if (!ctxs[r] || !ctxs[r]->initialized) return DEM_ERR_INVALID;
    std::vector<int> rcs(world, DEM_OK);
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++)
        th.emplace_back([&, r]() { rcs[r] = step_async_one(ctxs[r], n_steps); });
    for (auto& t : th) t.join();
    for (int r = 0; r < world; r++)
        if (rcs[r]) return rcs[r];
    return DEM_OK;
}

int dem_group_sync(DemCtx** ctxs, int world) {
    if (!ctxs || world < 1) return DEM_ERR_INVALID;
    std::vector<int> rcs(world, DEM_OK);
    std::vector<std::thread> th;
    // (a rebuild that overflowed did so on every rank alike -- the verdict is agreed on the device -- so every rank
    // rolls back and replays on its own, and they meet again in the device-side exchanges)
    for (int r = 0; r < world; r++)
        th.emplace_back([&, r]() { rcs[r] = sync_one(ctxs[r]); });
    for (auto& t : th) t.join();
    for (int r = 0; r < world; r++)
        if (rcs[r]) return rcs[r];
    return DEM_OK;
}

/* After dem_group_sync: copy into ctxs[0] the records of the clump owners the other ranks own, so that ctxs[0] holds the
 * merged state of the whole system for trackers / writers / inspectors. */
int dem_group_gather(DemCtx** ctxs, int world) {
    if (!ctxs || world < 1 || !ctxs[0]) return DEM_ERR_INVALID;
    DemCtx* ctx = ctxs[0];
    if (world == 1 || !ctx->mg.on) return DEM_OK;
    CK(cudaSetDevice(ctx->device));
    const DevParams P = make_params(ctx);
    for (int r = 1; r < world; r++)
        ctx->launches += launch_mg_gather_owned(P, ctxs[r]->d_state, ctxs[r]->d_spin, ctxs[r]->mg.d_flag, ctx->nClumpOwners, ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));
    return DEM_OK;
}

/* device-side barrier over the ranks, enqueued on the stream: what a benchmark puts right before its first timing event
 * so that the timed region starts with all GPUs level */
int dem_mgpu_barrier(DemCtx* ctx) {
    if (!ctx || !ctx->initialized) return DEM_ERR_INVALID;
    if (!ctx->mg.on) return DEM_OK;
    CK(cudaSetDevice(ctx->device));
    ctx->launches += launch_mg_barrier(make_params(ctx), make_mgdev(ctx), ctx->stream);
    return DEM_OK;
}

int dem_mgpu_info(DemCtx* ctx, uint64_t out[6]) {
    if (!ctx || !out) return DEM_ERR_INVALID;
    const MgState& g = ctx->mg;
    out[0] = g.last[0]; out[1] = g.last[3]; out[2] = g.last[1]; out[3] = g.last[2];
    out[4] = ((uint64_t)g.last[1] + g.last[2]) * 80;
    out[5] = (g.on ? (uint64_t)g.world : 1) | (g.on ? (1ull << 32) : 0ull);  // bit 32: exchange through peer memory
    return DEM_OK;
}

int dem_set_option(DemCtx* ctx, const char* name, double value) {
    if (!ctx || !name) return DEM_ERR_INVALID;
    FORWARD_TO_PEERS(dem_set_option(peer, name, value))
    const std::string n(name);
    if (n == "group_min_owners") { ctx->group_min_owners = value; return DEM_OK; }
    if (n == "ctas_per_sm") ctx->ctas_per_sm = std::max(2, std::min(4, (int)value));
    else if (n == "fast_math") ctx->fast_math = value != 0.0;
    else if (n == "use_graph") { ctx->use_graph = (int)value; if (ctx->use_graph == 0) graph_drop(ctx); }
    else if (n == "keep_acc") ctx->keep_acc = value != 0.0;
    else if (n == "fast_encode") ctx->fast_encode = value != 0.0;
    else if (n == "force_opts") ctx->force_opts = (int)value;
    else if (n == "adaptive_update_freq") { ctx->tuner.on = value != 0.0; ctx->tuner.prev_us = -1.0; ctx->tuner.acc_cycles = 0; ctx->tuner.acc_us = 0.0; ctx->tuner.acc_steps = 0; }
    else if (n == "update_freq_min") ctx->tuner.fmin = std::max(1, (int)value);
    else if (n == "update_freq_max") ctx->tuner.fmax = std::max(ctx->tuner.fmin, (int)value);
    else if (n == "sort_mode") ctx->sort_mode = (int)value;
    else if (n == "overlap_walls") ctx->overlap_walls = value != 0.0;
    else return fail(ctx, DEM_ERR_INVALID, "unknown option '%s'", name);
    return DEM_OK;
}

int dem_profile_rebuild(DemCtx* ctx, float out_us[8]) {
    if (!ctx || !ctx->initialized || !out_us) return DEM_ERR_INVALID;
    if (ctx->group_on) return fail(ctx, DEM_ERR_INVALID, "profiling hooks act on one device: create the context with dem_ctx_create");
    CK(cudaSetDevice(ctx->device));
    int rc = settle(ctx);
    if (rc) return rc;
    return rebuild_blocking(ctx, out_us);
}

int dem_profile_binning(DemCtx* ctx, uint32_t repeats, float out_us[3]) {
    if (!ctx || !ctx->initialized || !out_us || repeats == 0) return DEM_ERR_INVALID;
    if (ctx->group_on) return fail(ctx, DEM_ERR_INVALID, "profiling hooks act on one device: create the context with dem_ctx_create");
    if (ctx->mg.on) return fail(ctx, DEM_ERR_INVALID, "dem_profile_binning: single-device contexts only (shard the spheres, one context per GPU)");
    CK(cudaSetDevice(ctx->device));
    { int rcs = settle(ctx); if (rcs) return rcs; }
    cudaStream_t s = ctx->stream;
    double acc[3] = {0, 0, 0};
    for (uint32_t r = 0; r < repeats; r++) {
        DevParams P = make_params(ctx);
        CdParams C = make_cd(ctx);
        // scratch lists: the other buffers (the analytical list is emitted by the key pass)
        P.ss = as_list(ctx->lists[0][ctx->cur ^ 1]); P.sn = as_list(ctx->lists[1][ctx->cur ^ 1]);
        P.sa = as_list(ctx->lists[2][ctx->cur ^ 1]); P.st = as_list(ctx->lists[3][ctx->cur ^ 1]);
        CK(cudaEventRecord(ctx->ev[0], s));
        int launches = launch_cd_prepare(P, C, nullptr, ctx->need_maxvel, 0, ctx->num_sms, s);
        launches += launch_cd_prepare(P, C, nullptr, ctx->need_maxvel, 1, ctx->num_sms, s);
        launches += launch_cd_prepare(P, C, nullptr, ctx->need_maxvel, 2, ctx->num_sms, s);
        CK(cudaEventRecord(ctx->ev[1], s));
        int sorted_buf = -1;
        if (ctx->sort_mode == 0) { launches += launch_cd_sort(P, C, ctx->key_bits, s, &sorted_buf); ctx->last_sorted_buf = sorted_buf; }
        launches += launch_cd_sweep(P, C, nullptr, ctx->cur ^ 1, sorted_buf, ctx->num_sms, s, nullptr, /*sort_only*/ true);
        CK(cudaEventRecord(ctx->ev[2], s));
        CK(cudaEventSynchronize(ctx->ev[2]));
        ctx->launches += launches;
        ctx->need_maxvel = false;
        float ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); acc[0] += ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); acc[1] += ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2])); acc[2] += ms;
    }
    CK(cudaMemcpy(&ctx->last_grid, ctx->d_grid, sizeof(GridInfo), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) out_us[k] = (float)(acc[k] * 1000.0 / repeats);
    ctx->need_rebuild = true;  // the scratch of the real lists was reused
    return DEM_OK;
}

int dem_debug_download(DemCtx* ctx, const char* what, void* out, uint64_t n, uint64_t* n_out) {
    if (!ctx || !ctx->initialized || !what || !out || !n_out) return DEM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    { int rcs = settle(ctx); if (rcs) return rcs; }
    const std::string w(what);
    const void* src = nullptr;
    uint64_t avail = 0;
    const uint64_t nS = ctx->nSpheres;
    if (w == "sphere_keys") { src = ctx->d_keys[0]; avail = nS; }
    else if (w == "sorted_keys") { src = (ctx->sort_mode == 1) ? ctx->d_vals[0] : ctx->d_keys[ctx->last_sorted_buf]; avail = nS; }
    else if (w == "sphere_pos") { src = ctx->d_sphF; avail = nS * 4; }
    // compiled records of the sphere--sphere candidate list in use (4 words each), then the status words
    else if (w == "sn_cinfo") { src = ctx->lists[1][ctx->cur].cinfo; avail = ctx->n_list[1] * 4; }
    else if (w == "sn_due") { src = ctx->lists[1][ctx->cur].due; avail = (ctx->n_list[1] + 3) / 4; }  // one byte per candidate
    else if (w == "flags") { src = ctx->d_flags; avail = DEM_NUM_FLAGS; }
    else if (w == "sorted_ids") {
        // second word of the 16-byte sorted meta record
        const uint64_t m = std::min<uint64_t>(n, nS);
        if (m) CK(cudaMemcpy2D(out, 4, reinterpret_cast<const char*>(ctx->d_sortedMeta) + 4, 16, 4, m, cudaMemcpyDeviceToHost));
        *n_out = m;
        return DEM_OK;
    } else return fail(ctx, DEM_ERR_INVALID, "dem_debug_download: unknown array '%s'", what);
    const uint64_t m = std::min<uint64_t>(n, avail);
    if (m && src) CK(cudaMemcpy(out, src, m * 4, cudaMemcpyDeviceToHost));
    *n_out = m;
    return DEM_OK;
}

int dem_profile_steps(DemCtx* ctx, uint64_t n_steps, float out_us[8]) {
    if (!ctx || !ctx->initialized || !out_us) return DEM_ERR_INVALID;
    if (ctx->group_on) return fail(ctx, DEM_ERR_INVALID, "profiling hooks act on one device: create the context with dem_ctx_create");
    CK(cudaSetDevice(ctx->device));
    { int rcs = settle(ctx); if (rcs) return rcs; }
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaStream_t s = ctx->stream;
    const int model = (int)ctx->sp.force_model;
    const bool rec = ctx->sp.record_contact_forces != 0;
    // Events are recorded for a whole batch of steps and read back afterwards: a host synchronisation per step would
    // expose launch latencies and, on several GPUs, let the ranks drift apart between steps.
    constexpr int NE = 6, BATCH = 128;
    std::vector<cudaEvent_t> ev((size_t)NE * BATCH);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    uint64_t done = 0;
    int rc_out = DEM_OK;
    while (done < n_steps && rc_out == DEM_OK) {
        const int nb = (int)std::min<uint64_t>(BATCH, n_steps - done);
        int got = 0;
        for (int i = 0; i < nb; i++) {
            cudaEvent_t* e = &ev[(size_t)NE * i];
            CK(cudaEventRecord(e[0], s));
            if (ctx->need_rebuild || ctx->steps_since_rebuild >= ctx->list_freq) {
                const uint32_t before = ctx->seq_host;
                int rc = enqueue_rebuild(ctx);
                if (rc) { rc_out = rc; break; }
                if (ctx->seq_host == before) break;  // an earlier rebuild was rolled back: start the batch over
            }
            CK(cudaEventRecord(e[1], s));
            DevParams P = make_params(ctx);
            launch_force_ss(P, model, rec, ctx->num_sms, ctx->ctas_per_sm, ctx->fast_math != 0, s);
            CK(cudaEventRecord(e[2], s));
            if (ctx->nAnal > 0) launch_force_sa(P, model, rec, ctx->sa_grid, s);
            if (ctx->nTri > 0) launch_force_st(P, model, rec, ctx->sa_grid, s);
            CK(cudaEventRecord(e[3], s));
            // (on several GPUs the integrator also stores the halo records; [5] is the pull, including the wait)
            if (ctx->mg.on) launch_integrate_halo(P, 256, s);
            launch_integrate(P, integrate_grid(ctx), s);
            ctx->maxvel_slot ^= 1;
            CK(cudaEventRecord(e[4], s));
            if (ctx->mg.on) ctx->launches += 1 + launch_mg_pull(P, make_mgdev(ctx), ctx->cur, ctx->num_sms, s);
            CK(cudaEventRecord(e[5], s));
            ctx->launches += 2 + (ctx->nAnal > 0 ? 1 : 0) + (ctx->nTri > 0 ? 1 : 0);
            note_steps_enqueued(ctx, 1);
            got++;
        }
        if (rc_out != DEM_OK) break;
        const uint64_t steps_before = ctx->n_steps;
        ctx->steps_target = ctx->n_steps;
        { int rcs = settle(ctx); if (rcs) { rc_out = rcs; break; } }
        if (ctx->n_steps != steps_before) continue;  // rolled back inside the batch: its timings are void
        for (int i = 0; i < got; i++) {
            cudaEvent_t* e = &ev[(size_t)NE * i];
            float ms;
            CK(cudaEventElapsedTime(&ms, e[1], e[2])); acc[0] += ms;
            CK(cudaEventElapsedTime(&ms, e[2], e[3])); acc[1] += ms;
            CK(cudaEventElapsedTime(&ms, e[3], e[4])); acc[2] += ms;
            CK(cudaEventElapsedTime(&ms, e[0], e[1])); acc[3] += ms;
            CK(cudaEventElapsedTime(&ms, e[0], e[5])); acc[4] += ms;
            CK(cudaEventElapsedTime(&ms, e[4], e[5])); acc[5] += ms;
        }
        done += got;
    }
    ctx->steps_target = ctx->n_steps;
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc_out != DEM_OK) return rc_out;
    for (int k = 0; k < 8; k++) out_us[k] = n_steps ? (float)(acc[k] * 1000.0 / (double)n_steps) : 0.f;
    return DEM_OK;
}

}  // extern "C"
