// rbc3dg_api.h — internal interface between the C ABI (rbc3d_lib.cu) and the stage-streaming 3D kernels for arbitrary grids
// (rbc3dg_lib.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "rbc3dg_core.h"

namespace rbc3dg_api {

struct Plan;   // grid, tables, per-batch scratch (predicted state, tendency slab, spectral planes, phi) of one handle

// type-erased per-environment arrays: `state` points at float or double values according to the plan's precision
struct IoRaw {
    void* state;              // [B][nstate]
    const float* actions;     // [B][heaters][heaters]
    float* obs;               // [B][4][nz][ny][nx] or nullptr
    float* reward;
    double* nusselt;
    double* t;
    int* step_count;
    int* truncated;
    int* nan_flag;
};

int supported(int nx, int ny, int nz);
int create(const rbc3dg::HostConfigG& hc, int nx, int ny, int nz, int num_envs, int precision, int device, Plan** out);   // 0 or -1 (rbc_fail)
void destroy(Plan* p);
// Rayleigh number per environment (host array [B]; nullptr restores the configuration's value for all)
int set_rayleigh(Plan* p, const double* ra_host);
// one pass over the listed environments: optional set! projection, nsub RK3 steps, epilogue (observation, Nusselt, flags, clock)
// `vec` (mode >= 0, with advance_clock): the fused vector-env semantics of rbc2d::VecIO — auto-reset from the checkpoint bank in
// next_step / same_step mode, NaN reset, episode returns, terminal observations (final_nu_a = terminal Nusselt number)
int launch(Plan* p, const IoRaw& io, const int* env_ids, int n, int nsub, int project_first, int advance_clock, cudaStream_t stream,
           int64_t* launches, const rbc2d::VecIO* vec = nullptr);
int nsub_of(const Plan* p);
size_t smem_bytes(const Plan* p);
int values_per_env(const Plan* p);

}  // namespace rbc3dg_api
