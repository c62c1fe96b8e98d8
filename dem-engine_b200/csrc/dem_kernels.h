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
    ContactList oldss, oldsn, oldsa;
    uint32_t* rs_hist;    // radix-sort tile histograms
    uint32_t* scan_tmp;   // block sums for the scans
};

void launch_force_ss(const DevParams& P, int model, bool record, int num_sms, int ctas_per_sm, cudaStream_t s);
void launch_force_sa(const DevParams& P, int model, bool record, int grid, cudaStream_t s);
void launch_integrate(const DevParams& P, cudaStream_t s);

// rebuild stages; each returns the number of kernels it launched
int launch_cd_prepare(const DevParams& P, const CdParams& C, bool need_maxvel, cudaStream_t s);
int launch_cd_sort(const DevParams& P, const CdParams& C, int key_bits, cudaStream_t s, int* out_buf);
int launch_cd_sweep(const DevParams& P, const CdParams& C, int sorted_buf, cudaStream_t s, cudaEvent_t* ev = nullptr);
int launch_scan_exclusive(uint32_t* data, uint32_t n, uint32_t* tmp, uint32_t* total, cudaStream_t s);
int launch_reduce(const DevParams& P, int kind, double* d_out, cudaStream_t s);

}  // namespace demb
