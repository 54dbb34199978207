// rbc2dx_lib.cu — sm_100a cluster kernels for the 2D grids beyond 96 x 64 (rbc2dx_core.h) and their
// launch plumbing.  One thread-block cluster owns one environment at a time; the cluster grid is
// persistent (as many clusters as the device can keep resident) and strides over the batch.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "rbc2dx_api.h"
#include "rbc2dx_core.h"
#include "rbc_common.h"

using namespace rbc2dx;

template <typename G, typename Real, bool NXTG>
__global__ void __launch_bounds__(G::NT, 1)
rbc2dx_env_kernel(Consts<Real> C, EnvIO<Real> io, CtxX<Real> Xg, const int* env_ids, int n, RunFlags F)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using L = SmemLayoutX<G, Real, NXTG>;
    const int my_rank = RBX_RANK();
    const int cluster_id = blockIdx.x / G::CL, n_clusters = gridDim.x / G::CL;
    CtxX<Real> X = Xg;
    X.base = smem;
    X.gm = Xg.gm + (size_t)cluster_id * G::CL * 2 * G::NLOC;
    X.nxt_g = NXTG ? Xg.nxt_g + (size_t)cluster_id * G::CL * G::NS_SM : nullptr;
    Real* twN = reinterpret_cast<Real*>(smem + L::twN);
    Real* tw2 = reinterpret_cast<Real*>(smem + L::tw2);
    for (int q = threadIdx.x; q < 2 * G::NH; q += G::NT) { twN[q] = Xg.twN[q]; tw2[q] = Xg.tw2[q]; }
    // halo rows beyond the walls and the row paddings are never written afterwards: zero everything once
    Real* s0 = reinterpret_cast<Real*>(smem + L::s0);
    for (int q = threadIdx.x; q < G::NS_SM; q += G::NT) s0[q] = Real(0);
    if (!NXTG) {
        Real* s1 = reinterpret_cast<Real*>(smem + L::s1);
        for (int q = threadIdx.x; q < G::NS_SM; q += G::NT) s1[q] = Real(0);
    }
    if (NXTG) {
        Real* R = reinterpret_cast<Real*>(smem + L::R);
        for (int q = threadIdx.x; q < G::NR; q += G::NT) R[q] = Real(0);
    } else {                                             // this rank's Thomas pivots stay on-chip
        Real* tv = reinterpret_cast<Real*>(smem + L::tinv);
        const Real* src = Xg.tinv + (size_t)my_rank * G::NZL * G::NX;
        for (int q = threadIdx.x; q < G::NZL * G::NX; q += G::NT) tv[q] = src[q];
    }
    if (G::CL > 1 && !NXTG && threadIdx.x == 0) {
        // receive barriers of the three push channels (two per channel, alternating): one arrival = the local expect_tx
        for (int b = 0; b < 2 * NCHAN; ++b) mbar_init(smem_u32(smem + L::bars + 8 * b), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // every CTA of the cluster must be resident and initialised before the first remote store
    __syncthreads();
    RBX_SYNC_CLUSTER(G);
    SyncState S;
    for (int c = 0; c < NCHAN; ++c) S.n[c] = 0;
    for (int j = cluster_id; j < n; j += n_clusters) {
        const int env = env_ids ? env_ids[j] : j;
        env_action_step<G, Real, NXTG>(C, io, X, env, F, my_rank, S);
    }
}

namespace rbc2dx_api {

struct Plan {
    int nx = 0, nz = 0, precision = 32, device = 0, cl = 1, nt = 0, max_clusters = 0;
    size_t smem = 0, real_size = 4;
    double lx = 0, lz = 0;
    void *tables = nullptr, *gm = nullptr, *nxt = nullptr;
    const char* name = "";
    // type-erased launcher of the chosen instantiation
    int (*launch_fn)(Plan*, const rbc2d::HostConfig&, const rbc2d::HostWrappers&, const IoRaw&, const int*, int, rbc2d::RunFlags,
                     cudaStream_t) = nullptr;
};

template <typename G, typename Real, bool NXTG>
static int launch_impl(Plan* p, const rbc2d::HostConfig& hc, const rbc2d::HostWrappers& wr, const IoRaw& r, const int* env_ids, int n,
                       rbc2d::RunFlags F, cudaStream_t stream)
{
    using L = SmemLayoutX<G, Real, NXTG>;
    Consts<Real> C = make_consts<G, Real>(hc, wr);
    EnvIO<Real> io;
    io.state = (Real*)r.state; io.actions = r.actions; io.obs = r.obs; io.reward = r.reward; io.nu_state = r.nu_state;
    io.nu_obs = r.nu_obs; io.t = r.t; io.step_count = r.step_count; io.truncated = r.truncated; io.nan_flag = r.nan_flag;
    io.pressure = nullptr; io.cell_dist = r.cell_dist;
    CtxX<Real> X;
    L::fill(X);
    X.base = nullptr; X.arena_stride = 0;
    X.gm = (Real*)p->gm;
    X.nxt_g = (Real*)p->nxt;
    const Real* q = (const Real*)p->tables;
    X.tinv = q; q += (size_t)G::NZ * G::NX;
    X.spv = q; q += (size_t)G::NZ * G::NX;
    X.spw = q; q += (size_t)G::NZ * G::NX;
    X.cxl = q; q += (size_t)G::CL * 2 * G::CL * G::NX;
    X.cxr = q; q += (size_t)G::CL * 2 * G::CL * G::NX;
    X.twN = q; q += 2 * G::NH;
    X.tw2 = q;
    X.thomas_scale = (Real)((hc.lz / G::NZ) * (hc.lz / G::NZ) / G::NH);
    const int clusters = n < p->max_clusters ? n : p->max_clusters;
    if (clusters <= 0) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * G::CL);
    cfg.blockDim = dim3(G::NT);
    cfg.dynamicSmemBytes = L::total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, rbc2dx_env_kernel<G, Real, NXTG>, C, io, X, env_ids, n, F));
    return 0;
}

template <typename G, typename Real, bool NXTG>
static int create_impl(Plan* p, const char* name)
{
    using L = SmemLayoutX<G, Real, NXTG>;
    static_assert(L::total <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
    auto k = rbc2dx_env_kernel<G, Real, NXTG>;
    p->cl = G::CL; p->nt = G::NT; p->smem = L::total; p->real_size = sizeof(Real); p->name = name;
    p->launch_fn = launch_impl<G, Real, NXTG>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    // persistent grid: as many clusters as can be co-resident
    cudaLaunchConfig_t cfg = {};
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device));
    cfg.gridDim = dim3((sms / G::CL) * G::CL);
    cfg.blockDim = dim3(G::NT);
    cfg.dynamicSmemBytes = L::total;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    CK(cudaOccupancyMaxActiveClusters(&nc, k, &cfg));
    if (nc < 1) return rbc_fail("rbc2dx: the cluster kernel does not fit on this device");
    p->max_clusters = nc;
    // tables (fp64 on the host, cast to the working precision)
    std::vector<double> td(table_doubles<G>());
    HostTables T;
    double* d = td.data();
    T.tinv = d; d += (size_t)G::NZ * G::NX;
    T.spv = d; d += (size_t)G::NZ * G::NX;
    T.spw = d; d += (size_t)G::NZ * G::NX;
    T.cxl = d; d += (size_t)G::CL * 2 * G::CL * G::NX;
    T.cxr = d; d += (size_t)G::CL * 2 * G::CL * G::NX;
    T.twN = d; d += 2 * G::NH;
    T.tw2 = d;
    build_tables_host<G>(p->lx, p->lz, T);
    std::vector<Real> tr(td.begin(), td.end());
    CK(cudaMalloc(&p->tables, tr.size() * sizeof(Real)));
    CK(cudaMemcpy(p->tables, tr.data(), tr.size() * sizeof(Real), cudaMemcpyHostToDevice));
    const size_t gm_bytes = (size_t)nc * G::CL * 2 * G::NLOC * sizeof(Real);
    CK(cudaMalloc(&p->gm, gm_bytes));
    CK(cudaMemset(p->gm, 0, gm_bytes));
    if (NXTG) {
        const size_t nb = (size_t)nc * G::CL * G::NS_SM * sizeof(Real);
        CK(cudaMalloc(&p->nxt, nb));
        CK(cudaMemset(p->nxt, 0, nb));
    }
    return 0;
}

// RBC_B200_CLUSTER=1: 96 x 64 through the 2-CTA cluster kernel; =0c1: through the same generic code as a single CTA
// (cluster size 1) — both for cross-validation and for measuring what the generic code costs against the dedicated kernel
static int force_cluster()
{
    const char* e = std::getenv("RBC_B200_CLUSTER");
    if (!e) return 0;
    if (e[0] == '1') return 2;
    if (e[0] == 'g') return 1;
    return 0;
}

int supported(int nx, int nz)
{
    if (nx == 192 && nz == 128) return 1;
    if (nx == 128 && nz == 64) return 1;
    if (nx == 96 && nz == 64 && force_cluster()) return 1;
    return 0;
}

int create(int nx, int nz, int precision, int device, double lx, double lz, Plan** out)
{
    *out = nullptr;
    if (!supported(nx, nz)) return rbc_fail("rbc2dx: no kernel registered for this grid");
    Plan* p = new Plan();
    p->nx = nx; p->nz = nz; p->precision = precision; p->device = device; p->lx = lx; p->lz = lz;
    int rc = -1;
    if (nx == 192 && nz == 128) {
        // fp32: 4 CTAs x 32 rows, both state buffers on-chip; fp64: 8 CTAs x 16 rows, predicted state in global memory
        if (precision == 32) rc = create_impl<Grid<192, 128, 4, 2>, float, false>(p, "rbc2dx_env_kernel<192x128,cl4,f32>");
        else rc = create_impl<Grid<192, 128, 8, 2>, double, true>(p, "rbc2dx_env_kernel<192x128,cl8,f64>");
    } else if (nx == 128 && nz == 64) {
        // a power-of-two width: 64-point complex FFT = 4 x 16; 2 CTAs x 32 rows, 256 threads = 128 columns x 2 strips
        if (precision == 32) rc = create_impl<Grid<128, 64, 2, 2, 4>, float, false>(p, "rbc2dx_env_kernel<128x64,cl2,f32>");
        else rc = create_impl<Grid<128, 64, 2, 2, 4>, double, true>(p, "rbc2dx_env_kernel<128x64,cl2,f64>");
    } else if (nx == 96 && nz == 64 && force_cluster() == 1) {
        if (precision == 32) rc = create_impl<Grid<96, 64, 1, 4>, float, false>(p, "rbc2dx_env_kernel<96x64,cl1,f32>");
        else rc = rbc_fail("rbc2dx: the single-CTA generic variant is built for fp32 only");
    } else if (nx == 96 && nz == 64) {
        if (precision == 32) rc = create_impl<Grid<96, 64, 2, 4>, float, false>(p, "rbc2dx_env_kernel<96x64,cl2,f32>");
        else rc = create_impl<Grid<96, 64, 2, 4>, double, true>(p, "rbc2dx_env_kernel<96x64,cl2,f64>");
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
