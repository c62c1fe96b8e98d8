// dem_kernels.h -- host-callable launchers of the CUDA kernels (internal to libdemcore.so)
#pragma once
#include "../../include/dem_b200.h"
#include "dem_device.cuh"

namespace demb {

// broad-phase grid chosen on the device at every rebuild
struct GridInfo {
    float cs, inv_cs;
    uint32_t nbx, nby, nbz, ncells;
    float max_margin, maxvel;
    float halo;  // domain decomposition: width of the ghost layer for this rebuild
    int x0;      // domain decomposition: first cell column of this rank's table (nbx counts the columns it holds)
};

// analytical component resolved to world space for one rebuild
struct __align__(16) AnalWorld {
    float px, py, pz;  // plane point / cylinder centre (LBF-relative)
    float dx, dy, dz;  // plane normal / cylinder axis
    float margin;      // margin of the owning body
    float size1;
    float normal_sign;
    uint32_t type, family, material;
};

// scratch + in/out of one contact-list rebuild
struct CdParams {
    GridInfo* grid;
    float ext[3];        // extent of the binned region (target box size, LBF-relative)
    float rmax;          // largest template sphere radius
    float rclump;        // upper bound of the circumscribed radius of any clump about its centre of mass
    float max_extra;     // largest family extra margin
    uint32_t max_cells;  // capacity of the cell table
    uint32_t any_mask;   // != 0 when any family pair is masked
    uint32_t capacity;   // contact capacity per list
    float4* sphF;        // per sphere {x,y,z,r+margin}, LBF-relative, original order
    uint32_t* keys[2];
    uint32_t* vals[2];
    uint32_t* cellStart;  // ncells+1 (histogram, then exclusive prefix)
    float4* sortedSph;
    uint4* sortedMeta;    // {owner, sphere id, comp | material<<16, family}
    AnalWorld* analw;
    uint32_t* sortedPos;  // sphere id -> position in the cell-sorted arrays
    ContactList oldss, oldsn, oldsa, oldst;
    // triangles (sphere--triangle broad phase)
    float4* triW1;          // world-space nodes of this rebuild (LBF-relative, float); triW1.w = margin of the mesh owner
    float4* triW2;
    float4* triW3;
    uint32_t* triCellStart; // per cell: triangles registered (histogram, then exclusive prefix), max_cells+1
    uint32_t* triCellFill;  // per cell fill cursor
    uint32_t* triCellList;  // triangle ids grouped by cell
    uint32_t tri_pair_cap;
    uint32_t* rs_hist;    // radix-sort tile histograms
    uint32_t* scan_tmp;   // block sums for the scans
    // domain decomposition (0 / nullptr on a single GPU): the rebuild walks the spheres of the active owners only and
    // bins them into the cell columns of this rank's slab (+ halo)
    const uint32_t* act_sph;  // compact list of active sphere ids
    uint32_t nActSph;
    uint32_t scan_cells;      // cell-table entries to clear and scan (ncells + 1 when the host knows the grid)
    int slab_on;
    float slab_lo, slab_hi;   // my slab in LBF-relative x
};

// multi-GPU (slab decomposition) bookkeeping passed to the kernels of kernels_mgpu.cu
struct MgParams {
    uint8_t* flag;            // per global owner: 0 unknown here, 1 own, 2 ghost
    uint32_t* active_list;    // compact list of active owners (own + ghost)
    uint32_t* counts;         // device: [0] own [1] send-left [2] send-right [3] active
    uint32_t* send_gid[2];    // own owners inside the halo of the left / right cut
    uint32_t send_cap;
    uint32_t nClumpOwners;    // owners >= this index are analytical / replicated
    float cut_lo, cut_hi;     // my slab in LBF-relative x
    const GridInfo* grid;     // halo width of this rebuild is grid->halo
    int has_left, has_right;
    uint32_t* act_sph;        // compact list of the spheres of active owners; counts[4] = its length
};

int launch_mg_classify(const DevParams& P, const MgParams& M, cudaStream_t s);
int launch_mg_pack(const DevParams& P, const uint32_t* gid, uint32_t n, void* buf, cudaStream_t s);
int launch_mg_unpack(const DevParams& P, const uint32_t* gid, uint32_t n, const void* buf, uint8_t* flag, cudaStream_t s);
int launch_mg_send_map(const uint32_t* gid, uint32_t n, int32_t* slot, uint32_t nOwners, cudaStream_t s);
int launch_mg_active_spheres(const DevParams& P, const MgParams& M, cudaStream_t s);
int launch_mg_active_list(const DevParams& P, const MgParams& M, cudaStream_t s);

// per-step halo exchange through peer memory (NVLink stores into the neighbour's receive buffer + a flag)
struct MgP2P {
    const uint32_t* send_gid[2];   // my halo owners for the left / right neighbour
    const uint32_t* recv_gid[2];   // the owners the left / right neighbour sends me
    uint32_t n_send[2], n_recv[2];
    int4* peer_recv[2];            // where my records for the left / right neighbour go (in THEIR memory, this epoch's half)
    unsigned long long* peer_flag[2];  // their "data of epoch e has arrived" word for my direction
    const int4* my_recv[2];        // where the left / right neighbour's records arrive (my memory, this epoch's half)
    unsigned long long* my_flag[2];
    unsigned long long epoch;      // 1, 2, 3, ... one per exchange
    uint32_t* block_counter;       // last-block detection of the push kernel
    int has[2];                    // neighbour present on the left / right
    int publish;                   // k_mg_pull publishes my epoch first (the integrator stored the records: fused push)
};
int launch_mg_push(const DevParams& P, const MgP2P& X, cudaStream_t s);
int launch_mg_pull(const DevParams& P, const MgP2P& X, cudaStream_t s);

void launch_force_ss(const DevParams& P, int model, bool record, int num_sms, int ctas_per_sm, bool fast, cudaStream_t s);
void launch_force_sa(const DevParams& P, int model, bool record, int grid, cudaStream_t s);
void launch_force_st(const DevParams& P, int model, bool record, int grid, cudaStream_t s);
int launch_cd_triangles(const DevParams& P, const CdParams& C, int stage, cudaStream_t s);
void launch_integrate(const DevParams& P, cudaStream_t s);

// rebuild stages; each returns the number of kernels it launched
// stage 0: max |v| (when stale); stage 1: grid + margin decision; stage 2: analytical prep, clears, sphere keys + SA list
int launch_cd_prepare(const DevParams& P, const CdParams& C, bool need_maxvel, int stage, cudaStream_t s);
int launch_cd_sort(const DevParams& P, const CdParams& C, int key_bits, cudaStream_t s, int* out_buf);
int launch_cd_sweep(const DevParams& P, const CdParams& C, int sorted_buf, cudaStream_t s, cudaEvent_t* ev,
                    bool sort_only = false);
int launch_scan_exclusive(uint32_t* data, uint32_t n, uint32_t* tmp, uint32_t* total, cudaStream_t s);
int launch_reduce(const DevParams& P, int kind, double* d_out, cudaStream_t s);
int launch_reduce_many(const DevParams& P, uint32_t mask, double* d_out, cudaStream_t s);

}  // namespace demb
