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

// What a rebuild tells the host, written by its last kernel into pinned, device-mapped host memory (a ring of
// REBUILD_STATUS_SLOTS entries indexed by the rebuild's sequence number): the host never synchronises inside a rebuild,
// it reads the slot of a rebuild it knows to be complete.
constexpr uint32_t REBUILD_STATUS_SLOTS = 4;
struct RebuildStatus {
    uint32_t seq;        // sequence number of the rebuild that wrote this slot (1, 2, ...)
    uint32_t poison;     // != 0: the sequence number of the (first) rebuild that failed -- this one or an earlier one
    uint32_t capflags;   // DEM_FLAG_CAPACITY bits (all ranks' OR on several GPUs)
    uint32_t velflag;    // DEM_FLAG_VELOCITY
    uint32_t haloflags;  // DEM_FLAG_HALO
    uint32_t tri_demand; // entries the triangle--cell table needs
    uint32_t pad_[2];
    uint32_t count[4];   // contacts per list {ss touching, ss candidates, sa, st}, clamped to the capacity
    uint32_t demand[4];  // ... and unclamped
    GridInfo grid;
    uint32_t mg[8];      // multi-GPU counts of this cycle {own, send-left, send-right, active, active spheres, recv-left, recv-right, -}
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
    float4* sortedSph;    // cell-sorted stream {x,y,z,r+margin}           (+2 entries of slack: bulk copies are 16-byte units)
    uint2* sortedAux;     // ... {owner, bits(r)}                          (+2)
    uint4* sortedMeta;    // ... {owner, sphere id, comp | material<<16, family}
    AnalWorld* analw;
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
    uint32_t* scan_tmp;   // block sums for the radix sort's scans
    unsigned long long* scan_desc;  // [0] tile counter, [1..] tile descriptors of the single-pass scan
    uint32_t* cand;       // scratch: 12 words per sorted position, the candidates the count pass accepted (-> fill pass)
    uint32_t* idA_ss;     // scratch: sphere A of every contact of the new lists (sweep fill -> k_history)
    uint32_t* idA_sn;
    // domain decomposition (nullptr on a single GPU): the rebuild walks the spheres of the active owners only and
    // bins them into the cell columns of this rank's slab (+ halo)
    const uint32_t* act_sph;    // compact list of active sphere ids
    const uint32_t* act_count;  // its DEVICE-resident length
    int slab_on;
    float slab_lo, slab_hi;   // my slab in LBF-relative x
    RebuildStatus* status;    // device view of the pinned status ring
};

// multi-GPU kernels (kernels_mgpu.cu); par = parity of the cycle being built (== index of the new contact lists)
int launch_mg_redistribute(const DevParams& P, const MgDev& M, const GridInfo* grid, int par, int num_sms, cudaStream_t s);
int launch_mg_pull(const DevParams& P, const MgDev& M, int par, int num_sms, cudaStream_t s);
int launch_mg_barrier(const DevParams& P, const MgDev& M, cudaStream_t s);
int launch_mg_gather_owned(const DevParams& P, const OwnerState* peer_state, const float4* peer_spin,
                           const uint8_t* peer_flag, uint32_t nClumpOwners, cudaStream_t s);

void launch_force_ss(const DevParams& P, int model, bool record, int num_sms, int ctas_per_sm, bool fast, cudaStream_t s);
void launch_force_sa(const DevParams& P, int model, bool record, int grid, cudaStream_t s);
void launch_force_st(const DevParams& P, int model, bool record, int grid, cudaStream_t s);
int launch_cd_triangles(const DevParams& P, const CdParams& C, int stage, int num_sms, cudaStream_t s);
void launch_integrate(const DevParams& P, int grid_hint, cudaStream_t s);
void launch_integrate_halo(const DevParams& P, int grid_hint, cudaStream_t s);

// rebuild stages; each returns the number of kernels it launched
// stage 0: max |v| (when stale); stage 1: grid + margin decision (all-gathers max |v| over the ranks);
// stage 2: analytical prep, clears, sphere keys + SA list
int launch_cd_prepare(const DevParams& P, const CdParams& C, const MgDev* M, bool need_maxvel, int stage, int num_sms,
                      cudaStream_t s);
int launch_cd_sort(const DevParams& P, const CdParams& C, int key_bits, cudaStream_t s, int* out_buf);
int launch_cd_sweep(const DevParams& P, const CdParams& C, const MgDev* M, int par, int sorted_buf, int num_sms,
                    cudaStream_t s, cudaEvent_t* ev, bool sort_only = false);
void launch_sweep_count(const DevParams& P, const CdParams& C, const uint32_t* keys, int grid, cudaStream_t s);
void launch_sweep_fill(const DevParams& P, const CdParams& C, const uint32_t* keys, int grid, cudaStream_t s);
int launch_scan_exclusive(uint32_t* data, uint32_t n, uint32_t* tmp, uint32_t* total, cudaStream_t s);
// single-pass (decoupled look-back) exclusive scan of n = (n_ptr ? *n_ptr + n_add : n_add) words, in -> out (may alias);
// *total receives the sum.  desc: [0] tile counter + tile descriptors, cleared here.
int launch_scan_lookback(const uint32_t* in, uint32_t* out, const uint32_t* n_ptr, uint32_t n_add, uint32_t n_max,
                         unsigned long long* desc, uint32_t* total, const uint32_t* flags, int num_sms, cudaStream_t s,
                         const uint32_t* idx = nullptr);
int launch_zero_u32(uint32_t* p, size_t n, const uint32_t* flags, int num_sms, cudaStream_t s);
int launch_reduce(const DevParams& P, int kind, double* d_out, cudaStream_t s);
int launch_reduce_many(const DevParams& P, uint32_t mask, double* d_out, cudaStream_t s);
int launch_reduce_spheres(const DevParams& P, int kind, double* d_out, cudaStream_t s);

}  // namespace demb
