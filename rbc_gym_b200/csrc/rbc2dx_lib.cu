// rbc2dx_lib.cu — sm_100a cluster kernels for the 2D grids beyond 96 x 64 (rbc2dx_core.h) and their
// launch plumbing.  One thread-block cluster owns one environment at a time; the cluster grid is
// persistent (as many clusters as the device can keep resident) and strides over the batch.
#include <cstdlib>
#include <string>

#include "rbc2dx_kernel.cuh"

namespace rbc2dx_api {

// RBC_B200_CLUSTER=1: 96 x 64 through the 2-CTA cluster kernel; =g: through the same generic code as a single CTA
// (cluster size 1) — both for cross-validation and for measuring what the generic code costs against the dedicated kernel
static int force_cluster()
{
    const char* e = std::getenv("RBC_B200_CLUSTER");
    if (!e) return 0;
    if (e[0] == '1') return 2;
    if (e[0] == 'g') return 1;
    if (e[0] == '8') return 8;         // stress build only: 192 x 128 fp32 over 8 CTAs
    return 0;
}

int supported(int nx, int nz)
{
    if (nx == 192 && nz == 128) return 1;
    if (nx == 128 && nz == 64) return 1;
    if (nx == 96 && nz == 64 && (force_cluster() == 1 || force_cluster() == 2)) return 1;
    return more_supported(nx, nz);
}

int create(int nx, int nz, int precision, int split, int device, double lx, double lz, Plan** out)
{
    *out = nullptr;
    if (!supported(nx, nz)) return rbc_fail("rbc2dx: no kernel registered for this grid");
    Plan* p = new Plan();
    p->nx = nx; p->nz = nz; p->precision = precision; p->device = device; p->lx = lx; p->lz = lz;
    int rc = -1;
    if (more_supported(nx, nz)) {
        rc = split ? create_more_split(p) : create_more(p);
    } else if (split) {
        rc = create_split(p);
    } else if (nx == 192 && nz == 128) {
        // fp32: 4 CTAs x 32 rows, both state buffers on-chip; fp64: 8 CTAs x 16 rows, predicted state in global memory
#if defined(RBX_JITTER)
        if (precision == 32 && force_cluster() == 8) rc = create_impl<Grid<192, 128, 8, 2>, float, false, false>(p, "rbc2dx_env_kernel<192x128,cl8,f32,jitter>");
        else
#endif
        if (precision == 32) rc = create_impl<Grid<192, 128, 4, 2>, float, false, false>(p, "rbc2dx_env_kernel<192x128,cl4,f32>");
        else rc = create_impl<Grid<192, 128, 8, 2>, double, true, false>(p, "rbc2dx_env_kernel<192x128,cl8,f64>");
    } else if (nx == 128 && nz == 64) {
        // a power-of-two width: 64-point complex FFT = 4 x 16; 2 CTAs x 32 rows, 256 threads = 128 columns x 2 strips
        if (precision == 32) rc = create_impl<Grid<128, 64, 2, 2, 4>, float, false, false>(p, "rbc2dx_env_kernel<128x64,cl2,f32>");
        else rc = create_impl<Grid<128, 64, 2, 2, 4>, double, true, false>(p, "rbc2dx_env_kernel<128x64,cl2,f64>");
    } else if (nx == 96 && nz == 64 && force_cluster() == 1) {
        if (precision == 32) rc = create_impl<Grid<96, 64, 1, 4>, float, false, false>(p, "rbc2dx_env_kernel<96x64,cl1,f32>");
        else rc = rbc_fail("rbc2dx: the single-CTA generic variant is built for fp32 only");
    } else if (nx == 96 && nz == 64 && force_cluster() == 2) {
        if (precision == 32) rc = create_impl<Grid<96, 64, 2, 4>, float, false, false>(p, "rbc2dx_env_kernel<96x64,cl2,f32>");
        else rc = create_impl<Grid<96, 64, 2, 4>, double, true, false>(p, "rbc2dx_env_kernel<96x64,cl2,f64>");
    }
    if (rc) { destroy(p); return rc; }
    *out = p;
    return 0;
}

void destroy(Plan* p)
{
    if (!p) return;
    if (p->tables) cudaFree(p->tables);
    if (p->gm) cudaFree(p->gm);
    if (p->nxt) cudaFree(p->nxt);
    delete p;
}

int launch(Plan* p, const rbc2d::HostConfig& hc, const rbc2d::HostWrappers& wr, const IoRaw& io, const int* env_ids, int n,
           rbc2d::RunFlags F, cudaStream_t stream)
{
    return p->launch_fn(p, hc, wr, io, env_ids, n, F, stream);
}

int grid_ctas(const Plan* p, int n_envs) { return (n_envs < p->max_clusters ? n_envs : p->max_clusters) * p->cl; }
int cluster_size(const Plan* p) { return p->cl; }
size_t smem_bytes(const Plan* p) { return p->smem; }
const char* kernel_name(const Plan* p) { return p->name; }

}  // namespace rbc2dx_api
