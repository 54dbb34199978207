// rbc2dx_kernel.cuh — the cluster kernel template and its launch plumbing, shared by the translation units that
// instantiate it (rbc2dx_lib.cu: plain mode; rbc2dx_split.cu: pressure-split mode with the two pressure channels).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <vector>

#include "rbc2dx_api.h"
#include "rbc2dx_core.h"
#include "rbc_common.h"

using namespace rbc2dx;

template <typename G, typename Real, bool NXTG, bool SPLIT>
__global__ void __launch_bounds__(G::NT, 1)
rbc2dx_env_kernel(Consts<Real> C, EnvIO<Real> io, CtxX<Real> Xg, const int* env_ids, int n, RunFlags F)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using L = SmemLayoutX<G, Real, NXTG, SPLIT>;
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
    if (NXTG || SPLIT) {
        Real* R = reinterpret_cast<Real*>(smem + L::R);
        for (int q = threadIdx.x; q < G::NR; q += G::NT) R[q] = Real(0);
    } else {                                             // this rank's Thomas pivots stay on-chip (no room in split mode)
        Real* tv = reinterpret_cast<Real*>(smem + L::tinv);
        const Real* src = Xg.tinv + (size_t)my_rank * G::NZL * G::NX;
        for (int q = threadIdx.x; q < G::NZL * G::NX; q += G::NT) tv[q] = src[q];
    }
    if (G::CL > 1 && !NXTG && threadIdx.x == 0) {
        // receive barriers of the push channels (two per channel, alternating): one arrival = the local expect_tx
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
        env_action_step<G, Real, NXTG, SPLIT>(C, io, X, env, F, my_rank, S);
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

template <typename G, typename Real, bool NXTG, bool SPLIT>
static int launch_impl(Plan* p, const rbc2d::HostConfig& hc, const rbc2d::HostWrappers& wr, const IoRaw& r, const int* env_ids, int n,
                       rbc2d::RunFlags F, cudaStream_t stream)
{
    using L = SmemLayoutX<G, Real, NXTG, SPLIT>;
    Consts<Real> C = make_consts<G, Real>(hc, wr);
    EnvIO<Real> io;
    io.state = (Real*)r.state; io.actions = r.actions; io.obs = r.obs; io.reward = r.reward; io.nu_state = r.nu_state;
    io.nu_obs = r.nu_obs; io.t = r.t; io.step_count = r.step_count; io.truncated = r.truncated; io.nan_flag = r.nan_flag;
    io.pressure = SPLIT ? (Real*)r.pressure : nullptr; io.cell_dist = r.cell_dist;
    io.vec = r.vec;
    io.cfl_events = r.cfl_events;
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
    CK(cudaLaunchKernelEx(&cfg, rbc2dx_env_kernel<G, Real, NXTG, SPLIT>, C, io, X, env_ids, n, F));
    return 0;
}

template <typename G, typename Real, bool NXTG, bool SPLIT>
static int create_impl(Plan* p, const char* name)
{
    using L = SmemLayoutX<G, Real, NXTG, SPLIT>;
    static_assert(L::total <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
    auto k = rbc2dx_env_kernel<G, Real, NXTG, SPLIT>;
    p->cl = G::CL; p->nt = G::NT; p->smem = L::total; p->real_size = sizeof(Real); p->name = name;
    p->launch_fn = launch_impl<G, Real, NXTG, SPLIT>;
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

// pressure-split instantiations (rbc2dx_split.cu)
int create_split(Plan* p);

}  // namespace rbc2dx_api
