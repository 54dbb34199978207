// rbc2dx_api.h — internal interface between the C ABI (rbc2d_lib.cu) and the cluster kernels for the
// grids other than the dedicated 96 x 64 kernel (rbc2dx_lib.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>

#include "rbc2d_core.h"

namespace rbc2dx_api {

struct Plan;   // tables, per-cluster scratch and the kernel variant chosen for (grid, precision)

// type-erased EnvIO: `state` points at float or double values according to the plan's precision
struct IoRaw {
    void* state;
    const float* actions;
    float* obs;
    float* reward;
    double* nu_state;
    double* nu_obs;
    double* t;
    int* step_count;
    int* truncated;
    int* nan_flag;
    double* cell_dist;
    void* pressure;      // [B][2][NZ][NX] in the plan's precision (split mode), else nullptr
    rbc2d::VecIO vec;    // fused vector-env semantics (off by default)
    int* cfl_events = nullptr;   // [B] extra RK3 steps inserted by the CFL guard
};

// 1 when a cluster kernel is registered for this grid.  `force_cluster` (env RBC_B200_CLUSTER=1) also routes the
// 96 x 64 grid through the cluster kernel (2 CTAs per environment) for cross-validation and measurement.
int supported(int nx, int nz);
int more_supported(int nx, int nz);          // the further registered grids of rbc2dx_more.cuh
int create_more(Plan* p);                    // ... their plain-mode kernels (rbc2dx_more.cu)
int create_more_split(Plan* p);              // ... their pressure-split kernels (rbc2dx_more_split.cu)
int create(int nx, int nz, int precision, int split, int device, double lx, double lz, Plan** out);   // 0 or -1 (rbc_fail); split: pressure channels
void destroy(Plan* p);
int launch(Plan* p, const rbc2d::HostConfig& hc, const rbc2d::HostWrappers& wr, const IoRaw& io, const int* env_ids, int n,
           rbc2d::RunFlags F, cudaStream_t stream);
int grid_ctas(const Plan* p, int n_envs);    // CTAs a launch over n_envs environments uses
int cluster_size(const Plan* p);
size_t smem_bytes(const Plan* p);
const char* kernel_name(const Plan* p);

}  // namespace rbc2dx_api
