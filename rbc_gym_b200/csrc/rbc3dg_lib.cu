// rbc3dg_lib.cu — sm_100a kernels of the stage-streaming 3D action step for arbitrary grids (rbc3dg_core.h) and the plan
// object rbc3d_lib.cu drives them through.  Every kernel runs over the whole batch (blockIdx.y = environment), so a single
// 64 x 64 x 32 environment already fills the GPU with 131 072 threads per launch.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "rbc3dg_api.h"
#include "rbc3dg_core.h"
#include "rbc_common.h"

using namespace rbc3dg;

namespace {

constexpr int TB = 256;      // threads per block of the per-cell kernels

template <typename Real>
__global__ void g3_heater_kernel(Dims D, ConstsG<Real> C, const float* actions, Real* Tb, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D.ncol) return;
    Tb[(size_t)env * D.ncol + c] = (Real)heater_patch_T(C.heaters, C.heater_limit, C.b_hot, actions + (size_t)env * C.heaters * C.heaters,
                                                        c % D.nx, c / D.nx, D.nx, D.ny);
}

template <typename Real>
__global__ void __launch_bounds__(TB)
g3_tendency_kernel(Dims D, ConstsG<Real> C, const Real* nu_env, const Real* kappa_env, const Real* S_all, Real* P_all, Real* G_all,
                   const Real* Tb_all, const int* env_ids, Real dt, Real gam, Real zet, int use_prev, int store_g)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= D.nc) return;
    cell_tendency<Real>(D, C, nu_env[env], kappa_env[env], S_all + (size_t)env * D.nstate, P_all + (size_t)env * D.nstate,
                        G_all + (size_t)env * 4 * D.nc, Tb_all + (size_t)env * D.ncol, cell, dt, gam, zet, use_prev != 0, store_g != 0);
}

// one FFT stage on the device: the CTA's threads stride over the butterflies, then a barrier
struct BlockRun {
    template <typename F>
    __device__ void operator()(int n_items, F fn) const
    {
        for (int it = threadIdx.x; it < n_items; it += blockDim.x) fn(it);
        __syncthreads();
    }
};
// twiddles of both directions into shared memory, behind the plane: [nx/2] then [ny/2] complex
template <typename Real>
__device__ void stage_twiddles(const Dims& D, cx<Real>* dst, const cx<Real>* twx, const cx<Real>* twy)
{
    const int hx = D.nx >> 1, hy = D.ny >> 1;
    for (int q = threadIdx.x; q < hx + hy; q += blockDim.x) dst[q] = q < hx ? twx[q] : twy[q - hx];
}

// divergence of one level into a shared-memory plane, then the forward FFT in x and y (decimation in frequency)
template <typename Real>
__global__ void g3_div_fft_kernel(Dims D, ConstsG<Real> C, const Real* P_all, cx<Real>* Z_all, const cx<Real>* twx, const cx<Real>* twy,
                                  const int* env_ids)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<Real>* Z = reinterpret_cast<cx<Real>*>(smem_raw);
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y, k = blockIdx.x;
    const Real* P = P_all + (size_t)env * D.nstate;
    for (int c = threadIdx.x; c < D.ncol; c += blockDim.x) Z[c] = cx<Real>{cell_divergence<Real>(D, C, P, c % D.nx, c / D.nx, k), Real(0)};
    stage_twiddles<Real>(D, Z + D.ncol, twx, twy);
    __syncthreads();
    plane_fft_forward<Real>(D, Z, Z + D.ncol, Z + D.ncol + (D.nx >> 1), BlockRun{});
    cx<Real>* out = Z_all + ((size_t)env * D.nz + k) * D.ncol;
    for (int c = threadIdx.x; c < D.ncol; c += blockDim.x) out[c] = Z[c];
}

template <typename Real>
__global__ void g3_thomas_kernel(Dims D, cx<Real>* Z_all, const Real* cp, Real scale, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= D.ncol) return;
    mode_thomas<Real>(D, Z_all + (size_t)env * D.nz * D.ncol, cp, scale, m);
}

// inverse FFT of one level (decimation in time, bit-reversed order in, natural order out) -> phi
template <typename Real>
__global__ void g3_ifft_kernel(Dims D, const cx<Real>* Z_all, Real* phi_all, const cx<Real>* twx, const cx<Real>* twy, const int* env_ids)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<Real>* Z = reinterpret_cast<cx<Real>*>(smem_raw);
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y, k = blockIdx.x;
    const cx<Real>* in = Z_all + ((size_t)env * D.nz + k) * D.ncol;
    for (int c = threadIdx.x; c < D.ncol; c += blockDim.x) Z[c] = in[c];
    stage_twiddles<Real>(D, Z + D.ncol, twx, twy);
    __syncthreads();
    plane_fft_inverse<Real>(D, Z, Z + D.ncol, Z + D.ncol + (D.nx >> 1), BlockRun{});
    const Real norm = Real(1) / (Real)D.ncol;
    Real* phi = phi_all + ((size_t)env * D.nz + k) * D.ncol;
    for (int c = threadIdx.x; c < D.ncol; c += blockDim.x) phi[c] = Z[c].re * norm;
}

template <typename Real>
__global__ void __launch_bounds__(TB)
g3_correct_kernel(Dims D, ConstsG<Real> C, Real* P_all, const Real* phi_all, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= D.nc) return;
    cell_correct<Real>(D, C, P_all + (size_t)env * D.nstate, phi_all + (size_t)env * D.nc, cell);
}

// epilogue, part 1: Nusselt sum and NaN count of a chunk of cells (atomics into acc[env][2]); write-back to the environment's
// own array when the march ended in the other buffer; observation = get_state (rbc_sim3D_api.jl:106-121, rbc3D.py:229-232)
template <typename Real>
__global__ void __launch_bounds__(TB)
g3_reduce_kernel(Dims D, ConstsG<Real> C, const Real* cur_all, Real* st_all, float* obs, double* acc, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    const Real* cur = cur_all + (size_t)env * D.nstate;
    Real* st = st_all + (size_t)env * D.nstate;
    double a = 0, bad = 0;
    for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < D.nc; cell += gridDim.x * blockDim.x) {
        const int k = cell / D.ncol;
        const double b = (double)cur[D.gb + cell], u = (double)cur[D.gu + cell], v = (double)cur[D.gv + cell], w = (double)cur[D.gw + cell];
        if (b != b || u != u || v != v || w != w) bad += 1;
        // get_nusselt (rbc_sim3D_api.jl:134-159): conduction profile on a UNIT height, z = (k + 1/2)/nz
        const double zc = (k + 0.5) / D.nz, Tc = (1.0 - zc) * C.delta_b_d + C.b_top_d;
        a += (b - Tc) * w;
    }
    if (cur != st)
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < D.nstate; q += gridDim.x * blockDim.x) st[q] = cur[q];
    if (obs != nullptr) {
        float* ob = obs + (size_t)env * 4 * D.nc;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 4 * D.nc; q += gridDim.x * blockDim.x) ob[q] = (float)cur[q];
    }
    __shared__ double sa[TB], sb[TB];
    sa[threadIdx.x] = a; sb[threadIdx.x] = bad;
    __syncthreads();
    for (int s = TB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(&acc[2 * env], sa[0]); atomicAdd(&acc[2 * env + 1], sb[0]); }
}

// epilogue, part 2: Nusselt number, reward, NaN flag, clock and truncation of every listed environment
template <typename Real>
__global__ void g3_finalize_kernel(Dims D, ConstsG<Real> C, const double* kappa_env, const double* acc, rbc3dg_api::IoRaw io, const int* env_ids,
                                   int n, int advance_clock)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int env = env_ids ? env_ids[j] : j;
    const double nu = 1.0 + (acc[2 * env] / (double)D.nc) / kappa_env[env];
    io.nusselt[env] = nu;
    io.reward[env] = (float)(-nu);
    io.nan_flag[env] = acc[2 * env + 1] > 0 ? 1 : 0;
    if (advance_clock) {
        const double tn = io.t[env] + C.dt_action;
        io.t[env] = tn;
        io.step_count[env] += 1;
        io.truncated[env] = tn >= C.episode_length ? 1 : 0;
    }
}

__global__ void g3_zero_acc_kernel(double* acc, const int* env_ids, int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int env = env_ids ? env_ids[j] : j;
    acc[2 * env] = 0.0; acc[2 * env + 1] = 0.0;
}

}  // namespace

namespace rbc3dg_api {

struct Plan {
    Dims D;
    HostConfigG hc;
    int B = 0, precision = 32, device = 0;
    size_t rs = 4, smem = 0;
    int fft_threads = 256;
    void *P = nullptr, *G = nullptr, *Z = nullptr, *phi = nullptr, *Tb = nullptr, *cp = nullptr, *twx = nullptr, *twy = nullptr;
    void *nu = nullptr, *kappa = nullptr;          // per-environment diffusivities (Real)
    double* kappa_d = nullptr;                      // the same in fp64 for the Nusselt number
    double* acc = nullptr;
};

int supported(int nx, int ny, int nz) { return dims_supported(nx, ny, nz) ? 1 : 0; }
size_t smem_bytes(const Plan* p) { return p->smem; }
int values_per_env(const Plan* p) { return p->D.nstate; }

template <typename Real>
static int upload_tables(Plan* p)
{
    const Dims& D = p->D;
    std::vector<double> cp((size_t)D.nz * D.ncol), tx(D.nx), ty(D.ny);
    build_pivots_host(D, p->hc.lx, p->hc.ly, p->hc.lz, cp.data());
    build_twiddles_host(D.nx, tx.data());
    build_twiddles_host(D.ny, ty.data());
    std::vector<Real> a(cp.begin(), cp.end()), b(tx.begin(), tx.end()), c(ty.begin(), ty.end());
    CK(cudaMalloc(&p->cp, a.size() * sizeof(Real)));
    CK(cudaMalloc(&p->twx, b.size() * sizeof(Real)));
    CK(cudaMalloc(&p->twy, c.size() * sizeof(Real)));
    CK(cudaMemcpy(p->cp, a.data(), a.size() * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->twx, b.data(), b.size() * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->twy, c.data(), c.size() * sizeof(Real), cudaMemcpyHostToDevice));
    return 0;
}

template <typename Real>
static int upload_rayleigh(Plan* p, const double* ra /*[B] or nullptr*/)
{
    std::vector<Real> nu(p->B), ka(p->B);
    std::vector<double> kd(p->B);
    for (int e = 0; e < p->B; ++e) {
        const double r = ra ? ra[e] : p->hc.ra;
        if (!(r > 0)) return rbc_fail("rbc3d: Rayleigh numbers must be positive");
        nu[e] = (Real)sqrt(p->hc.pr / r); kd[e] = 1.0 / sqrt(p->hc.pr * r); ka[e] = (Real)kd[e];       // rbc_sim3D_api.jl:37-38
    }
    CK(cudaMemcpy(p->nu, nu.data(), p->B * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->kappa, ka.data(), p->B * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->kappa_d, kd.data(), p->B * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

int set_rayleigh(Plan* p, const double* ra_host)
{
    CK(cudaSetDevice(p->device));
    return p->precision == 32 ? upload_rayleigh<float>(p, ra_host) : upload_rayleigh<double>(p, ra_host);
}

int create(const HostConfigG& hc, int nx, int ny, int nz, int num_envs, int precision, int device, Plan** out)
{
    *out = nullptr;
    if (!dims_supported(nx, ny, nz)) return rbc_fail("rbc3d_create: grid must have nx, ny powers of two in 8..256 and 6 <= nz <= 256");
    Plan* p = new Plan();
    p->D = make_dims(nx, ny, nz);
    p->hc = hc; p->B = num_envs; p->precision = precision; p->device = device;
    p->rs = precision == 32 ? 4 : 8;
    p->smem = ((size_t)p->D.ncol + p->D.nx / 2 + p->D.ny / 2) * 2 * p->rs;      // one complex plane + the twiddles of both directions
    if (p->smem > 227 * 1024) { delete p; return rbc_fail("rbc3d_create: a horizontal plane of this grid does not fit the shared memory of an SM"); }
    cudaError_t e = cudaSuccess;
    if (precision == 32) {
        e = cudaFuncSetAttribute(g3_div_fft_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(g3_ifft_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    } else {
        e = cudaFuncSetAttribute(g3_div_fft_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(g3_ifft_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    }
    if (e != cudaSuccess) { delete p; return rbc_fail(std::string("rbc3d_create: ") + cudaGetErrorString(e)); }
    const size_t B = num_envs, rs = p->rs;
    const Dims& D = p->D;
    struct { void** ptr; size_t bytes; } allocs[] = {
        {&p->P, B * D.nstate * rs}, {&p->G, B * 4 * D.nc * rs}, {&p->Z, B * D.nc * 2 * rs}, {&p->phi, B * D.nc * rs},
        {&p->Tb, B * D.ncol * rs}, {&p->nu, B * rs}, {&p->kappa, B * rs}, {(void**)&p->kappa_d, B * sizeof(double)},
        {(void**)&p->acc, B * 2 * sizeof(double)}};
    for (auto& a : allocs) {
        e = cudaMalloc(a.ptr, a.bytes);
        if (e != cudaSuccess) { destroy(p); return rbc_fail(std::string("rbc3d_create: cudaMalloc: ") + cudaGetErrorString(e)); }
        cudaMemset(*a.ptr, 0, a.bytes);
    }
    int rc = precision == 32 ? upload_tables<float>(p) : upload_tables<double>(p);
    if (!rc) rc = set_rayleigh(p, nullptr);
    if (rc) { destroy(p); return rc; }
    *out = p;
    return 0;
}

void destroy(Plan* p)
{
    if (!p) return;
    void* ptrs[] = {p->P, p->G, p->Z, p->phi, p->Tb, p->cp, p->twx, p->twy, p->nu, p->kappa, p->kappa_d, p->acc};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete p;
}

template <typename Real>
static int project_t(Plan* p, const ConstsG<Real>& C, Real* buf, const int* env_ids, int n, cudaStream_t st, int64_t* launches)
{
    const Dims& D = p->D;
    const dim3 gplane(D.nz, n), gcell((D.nc + TB - 1) / TB, n), gmode((D.ncol + 127) / 128, n);
    const Real dz = (Real)(p->hc.lz / D.nz);
    g3_div_fft_kernel<Real><<<gplane, p->fft_threads, p->smem, st>>>(D, C, buf, (cx<Real>*)p->Z, (const cx<Real>*)p->twx, (const cx<Real>*)p->twy, env_ids);
    g3_thomas_kernel<Real><<<gmode, 128, 0, st>>>(D, (cx<Real>*)p->Z, (const Real*)p->cp, dz * dz, env_ids);
    g3_ifft_kernel<Real><<<gplane, p->fft_threads, p->smem, st>>>(D, (const cx<Real>*)p->Z, (Real*)p->phi, (const cx<Real>*)p->twx, (const cx<Real>*)p->twy, env_ids);
    g3_correct_kernel<Real><<<gcell, TB, 0, st>>>(D, C, buf, (const Real*)p->phi, env_ids);
    CK(cudaGetLastError());
    *launches += 4;
    return 0;
}

template <typename Real>
static int launch_t(Plan* p, const IoRaw& io, const int* env_ids, int n, int nsub, int project_first, int advance_clock, cudaStream_t st,
                    int64_t* launches)
{
    const Dims& D = p->D;
    const ConstsG<Real> C = make_consts<Real>(D, p->hc);
    const Real gam[3] = {Real(8.0 / 15.0), Real(5.0 / 12.0), Real(3.0 / 4.0)};
    const Real zet[3] = {Real(0), Real(-17.0 / 60.0), Real(-5.0 / 12.0)};
    Real* S = (Real*)io.state;
    Real* cur = S;
    Real* nxt = (Real*)p->P;
    const dim3 gcell((D.nc + TB - 1) / TB, n), gcol((D.ncol + TB - 1) / TB, n);
    if (nsub > 0) {
        g3_heater_kernel<Real><<<gcol, TB, 0, st>>>(D, C, io.actions, (Real*)p->Tb, env_ids);
        *launches += 1;
    }
    if (project_first) {
        int rc = project_t<Real>(p, C, cur, env_ids, n, st, launches);
        if (rc) return rc;
    }
    for (int sub = 0; sub < nsub; ++sub) {
        const Real dt = (sub == nsub - 1) ? C.dt_last : C.dt_full;
        for (int stage = 0; stage < 3; ++stage) {
            g3_tendency_kernel<Real><<<gcell, TB, 0, st>>>(D, C, (const Real*)p->nu, (const Real*)p->kappa, cur, nxt, (Real*)p->G, (const Real*)p->Tb,
                                                          env_ids, dt, gam[stage], zet[stage], stage > 0, stage < 2);
            *launches += 1;
            int rc = project_t<Real>(p, C, nxt, env_ids, n, st, launches);
            if (rc) return rc;
            Real* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
    const int rblocks = (D.nc + 4 * TB - 1) / (4 * TB) < 128 ? (D.nc + 4 * TB - 1) / (4 * TB) : 128;
    g3_zero_acc_kernel<<<(n + 127) / 128, 128, 0, st>>>(p->acc, env_ids, n);
    g3_reduce_kernel<Real><<<dim3(rblocks, n), TB, 0, st>>>(D, C, cur, S, io.obs, p->acc, env_ids);
    g3_finalize_kernel<Real><<<(n + 127) / 128, 128, 0, st>>>(D, C, p->kappa_d, p->acc, io, env_ids, n, advance_clock);
    CK(cudaGetLastError());
    *launches += 3;
    return 0;
}

int launch(Plan* p, const IoRaw& io, const int* env_ids, int n, int nsub, int project_first, int advance_clock, cudaStream_t stream,
           int64_t* launches)
{
    if (n <= 0) return 0;
    if (n > 65535) return rbc_fail("rbc3d: at most 65535 environments per launch on this path");
    return p->precision == 32 ? launch_t<float>(p, io, env_ids, n, nsub, project_first, advance_clock, stream, launches)
                              : launch_t<double>(p, io, env_ids, n, nsub, project_first, advance_clock, stream, launches);
}

int nsub_of(const Plan* p)
{
    return make_consts<float>(p->D, p->hc).nsub;
}

}  // namespace rbc3dg_api
