// rbc2d_lib.cu — sm_100a kernels + the C ABI declared in include/rbc_b200.h.
//
// One persistent CTA per SM; each CTA owns one environment at a time and keeps it in shared
// memory for the whole action step (rbc2d_core.h).  Environments are independent, so the grid
// simply strides over the batch.  No CPU fallback exists: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rbc_b200.h"
#include "rbc2d_core.h"
#include "rbc2dx_api.h"
#include "rbc_common.h"

using namespace rbc2d;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
int rbc_fail(const std::string& m) { g_err = m; return -1; }
static int fail(const std::string& m) { return rbc_fail(m); }

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
template <typename Real>
struct SmemLayout {
    // fp32: both state buffers, the Poisson scratch and the Thomas table live on-chip (the epilogue's
    //       reduction scratch aliases the dead state buffer).
    // fp64: one state buffer on-chip, the predicted state and the Thomas table in global memory.
    static constexpr bool kNxtGlobal = sizeof(Real) == 8;
    static constexpr bool kTinvShared = !kNxtGlobal;
    static constexpr size_t s0 = 0;
    static constexpr size_t s1 = s0 + sizeof(Real) * NS_SM;
    static constexpr size_t R = kNxtGlobal ? s1 : s1 + sizeof(Real) * NS_SM;
    static constexpr size_t tinv = R + sizeof(Real) * NR;
    static constexpr size_t red = tinv + (kTinvShared ? sizeof(Real) * NZ * NX : 0);
    static constexpr size_t Tb = red + (kNxtGlobal ? sizeof(double) * NRED * NT : 0);
    static constexpr size_t mid = Tb + sizeof(Real) * NX;
    static constexpr size_t tw48 = mid + sizeof(Real) * 2 * NX;
    static constexpr size_t tw96 = tw48 + sizeof(Real) * 96;
    // left fluxes of the warps' first columns (phase_edge_fluxes): fp64 mode keeps them in the reduction scratch, idle until the epilogue
    static constexpr size_t edge = kNxtGlobal ? red : tw96 + sizeof(Real) * 96;
    static constexpr size_t total = tw96 + sizeof(Real) * 96 + (kNxtGlobal ? 0 : sizeof(Real) * NE);
    static_assert(!kNxtGlobal || sizeof(Real) * NE <= sizeof(double) * NRED * NT, "edge-flux table must fit the reduction scratch");
    static_assert(total <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
    static constexpr size_t kAlign = 2 * sizeof(Real);     // one complex number per shared-memory access
    static_assert(s1 % kAlign == 0 && R % kAlign == 0 && tinv % kAlign == 0 && red % 8 == 0 && Tb % kAlign == 0, "alignment");
    static_assert((SX * sizeof(Real)) % kAlign == 0, "row stride must keep complex accesses aligned");
    static_assert(kNxtGlobal || sizeof(double) * NRED * NT <= sizeof(Real) * NS_SM, "reduction scratch must fit a state buffer");
};

template <typename Real, bool SPLIT>
__global__ void __launch_bounds__(NT, 1)
rbc2d_env_kernel(Consts<Real> C, Tables<Real> T, EnvIO<Real> io, Real* gm_all, Real* nxt_all, const int* env_ids, int n,
                 RunFlags F)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using L = SmemLayout<Real>;
    Ctx<Real> X;
    X.s0 = reinterpret_cast<Real*>(smem + L::s0);
    X.s1 = L::kNxtGlobal ? nxt_all + (size_t)blockIdx.x * NS_SM : reinterpret_cast<Real*>(smem + L::s1);
    X.R = reinterpret_cast<Real*>(smem + L::R);
    X.red = reinterpret_cast<double*>(smem + L::red);
    X.Tb = reinterpret_cast<Real*>(smem + L::Tb);
    X.mid = reinterpret_cast<Real*>(smem + L::mid);
    X.E = reinterpret_cast<Real*>(smem + L::edge);
    X.tw48 = reinterpret_cast<Real*>(smem + L::tw48);
    X.tw96 = reinterpret_cast<Real*>(smem + L::tw96);
    X.gm = gm_all + (size_t)blockIdx.x * 2 * NSTATE;
    if (L::kTinvShared) {
        Real* tv = reinterpret_cast<Real*>(smem + L::tinv);
        for (int q = threadIdx.x; q < NZ * NX; q += NT) tv[q] = T.tinv[q];
        X.tinv = tv;
    } else {
        X.tinv = T.tinv;
    }
    for (int q = threadIdx.x; q < 96; q += NT) { X.tw48[q] = T.tw48[q]; X.tw96[q] = T.tw96[q]; }
    // the padding columns of the on-chip rows are never read; zero them once so nothing is uninitialised
    for (int q = threadIdx.x; q < NS_SM; q += NT) { X.s0[q] = Real(0); if (!L::kNxtGlobal) X.s1[q] = Real(0); }
    __syncthreads();
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
        const int env = env_ids ? env_ids[j] : j;
        env_action_step<Real, SPLIT, L::kNxtGlobal>(C, T, io, X, env, F);
    }
}

// state[env_ids[j]] <- bank[ckpt_idx[j]] (fp64 -> Real); clock reset (initialize_simulation, rbc_sim2D_api.jl:66-68)
// (the small kernels below take the grid as run-time arguments: they serve every registered grid)
template <typename Real>
__global__ void rbc2d_reset_kernel(Real* state, const double* bank, const int* env_ids, const int* ckpt_idx, int n, int n_ep,
                                   double* t, int* step, int* trunc, int* nan, int NSTATE, int B)
{
    // indices are validated by the caller (the Python facade raises IndexError like the reference's BoundsError); an entry
    // that is out of range anyway is skipped rather than allowed to touch another environment or read past the bank
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        const int env = env_ids ? env_ids[j] : j;
        const int ep = ckpt_idx[j];
        if (env < 0 || env >= B || ep < 0 || ep >= n_ep) continue;
        const double* src = bank + (size_t)ep * NSTATE;
        Real* dst = state + (size_t)env * NSTATE;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < NSTATE; q += gridDim.x * blockDim.x) dst[q] = (Real)src[q];
        if (blockIdx.x == 0 && threadIdx.x == 0) { t[env] = 0.0; step[env] = 1; trunc[env] = 0; nan[env] = 0; }
    }
}

template <typename Real>
__global__ void rbc2d_set_fields_kernel(Real* state, const double* fields, const int* env_ids, int n, double* t, int* step,
                                        int* trunc, int* nan, int NSTATE, int B)
{
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        const int env = env_ids ? env_ids[j] : j;
        if (env < 0 || env >= B) continue;
        const double* src = fields + (size_t)j * NSTATE;
        Real* dst = state + (size_t)env * NSTATE;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < NSTATE; q += gridDim.x * blockDim.x) dst[q] = (Real)src[q];
        if (blockIdx.x == 0 && threadIdx.x == 0) { t[env] = 0.0; step[env] = 1; trunc[env] = 0; nan[env] = 0; }
    }
}

// get_state (rbc_sim2D_api.jl:102-118 + rbc2D.py:184-189): [B][C][NZ][NX] float32; w channel = faces 0..NZ-1
template <typename Real>
__global__ void rbc2d_get_state_kernel(const Real* state, const Real* pressure, float* out, int B, int channels, int NCELL, int NSTATE)
{
    const int GOFF_B = 0, GOFF_U = NCELL, GOFF_W = 2 * NCELL;
    const size_t per = (size_t)channels * NCELL, total = per * B;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int env = (int)(q / per), r = (int)(q % per), ch = r / NCELL, c = r % NCELL;
        float v;
        if (ch < 3) v = (float)state[(size_t)env * NSTATE + (ch == 0 ? GOFF_B : (ch == 1 ? GOFF_U : GOFF_W)) + c];
        else v = pressure ? (float)pressure[(size_t)env * 2 * NCELL + (size_t)(ch - 3) * NCELL + c] : 0.0f;
        out[q] = v;
    }
}

// render("rgb_array"): b -> turbo RGB (polynomial fit, same coefficients as rbc_gym_b200/colormap.py), flipped in z
template <typename Real>
__global__ void rbc2d_render_kernel(const Real* state, uint8_t* out, int B, int nx, int nz, int NSTATE, float vmin, float inv_range)
{
    const size_t ncell = (size_t)nx * nz, total = ncell * B;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int env = (int)(q / ncell), c = (int)(q % ncell), row = c / nx, i = c % nx;
        const float v = (float)state[(size_t)env * NSTATE + (size_t)(nz - 1 - row) * nx + i];
        const float x = fminf(fmaxf((v - vmin) * inv_range, 0.0f), 1.0f);
        const float r = 0.13572138f + x * (4.61539260f + x * (-42.66032258f + x * (132.13108234f + x * (-152.94239396f + x * 59.28637943f))));
        const float g = 0.09140261f + x * (2.19418839f + x * (4.84296658f + x * (-14.18503333f + x * (4.27729857f + x * 2.82956604f))));
        const float b = 0.10667330f + x * (12.64194608f + x * (-60.58204836f + x * (110.36276771f + x * (-89.90310912f + x * 27.34824973f))));
        out[3 * q + 0] = (uint8_t)(fminf(fmaxf(r, 0.0f), 1.0f) * 255.0f);
        out[3 * q + 1] = (uint8_t)(fminf(fmaxf(g, 0.0f), 1.0f) * 255.0f);
        out[3 * q + 2] = (uint8_t)(fminf(fmaxf(b, 0.0f), 1.0f) * 255.0f);
    }
}

template <typename Real>
__global__ void rbc2d_get_fields_kernel(const Real* state, double* out, size_t total)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x)
        out[q] = (double)state[q];
}

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct rbc2d_sim {
    rbc2d_config cfg;
    HostConfig hc;
    HostWrappers wr;
    double* cell_dist = nullptr;
    int B = 0, channels = 3, grid = 0, n_ep = 0;
    int nx = NX, nz = NZ, ncell = NCELL, nwf = NWF, nstate = NSTATE;   // grid of this handle
    rbc2dx_api::Plan* plan = nullptr;                                  // cluster kernels for grids other than 96 x 64
    size_t smem = 0, real_size = 4;
    cudaStream_t stream = nullptr;
    void *state = nullptr, *gm = nullptr, *nxt = nullptr, *pressure = nullptr;
    void *tinv = nullptr, *tw48 = nullptr, *tw96 = nullptr;
    double* bank = nullptr;
    double *t = nullptr, *nu_s = nullptr, *nu_o = nullptr;
    int *step = nullptr, *trunc = nullptr, *nan = nullptr;
    float *obs = nullptr, *reward = nullptr, *actions = nullptr;
    int64_t launches = 0;
    // CUDA events around the step kernel on the launch stream: a ring of the last kRing step launches, so that a
    // benchmark can read every launch of its timed region afterwards without synchronising inside it
    static constexpr int kRing = 64;
    cudaEvent_t ev0[kRing] = {}, ev1[kRing] = {};
    int64_t timed_launches = 0;
    HostPipe pipe;                                                     // rbc2d_step_host
    // fused vector-env state (rbc2d_set_autoreset / rbc2d_vec_step_dev)
    rbc_autoreset ar = {0, 0, 0, 0};
    int* pending = nullptr;
    long long* episode = nullptr;
    double* ep_return = nullptr;
    int* nan_count = nullptr;
    int* cfl_events = nullptr;       // [B] extra RK3 steps inserted by the CFL guard of the cluster kernels
    float* final_obs = nullptr;      // device staging of the final_* outputs of rbc2d_vec_step_host (allocated on first use)
    double* final_scalars = nullptr;
};

// the kernel-side view of the handle's vector-env state for one launch; `out` supplies the caller's final_* buffers
static VecIO make_vec(const rbc2d_sim* s, const rbc2d_vec_out* out)
{
    VecIO v;
    v.mode = s->ar.mode;
    v.nan_reset = s->ar.nan_reset;
    v.bank = s->bank;
    v.n_ep = s->n_ep;
    v.seed = (unsigned long long)s->ar.seed;
    v.id_offset = (unsigned long long)s->ar.env_id_offset;
    v.pending = s->pending;
    v.episode = s->episode;
    v.ep_return = s->ep_return;
    v.nan_count = s->nan_count;
    if (out) { v.final_obs = out->final_obs; v.final_nu_a = out->final_nu_state; v.final_nu_b = out->final_nu_obs; v.final_return = out->final_return; }
    return v;
}

template <typename Real>
static int upload_tables(rbc2d_sim* s)
{
    std::vector<double> tinv(NZ * NX), tw48(96), tw96(96);
    build_tables_host(s->hc.lx, s->hc.lz, tinv.data(), tw48.data(), tw96.data());
    std::vector<Real> a(tinv.begin(), tinv.end()), b(tw48.begin(), tw48.end()), c(tw96.begin(), tw96.end());
    CK(cudaMalloc(&s->tinv, a.size() * sizeof(Real)));
    CK(cudaMalloc(&s->tw48, b.size() * sizeof(Real)));
    CK(cudaMalloc(&s->tw96, c.size() * sizeof(Real)));
    CK(cudaMemcpy(s->tinv, a.data(), a.size() * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->tw48, b.data(), b.size() * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->tw96, c.data(), c.size() * sizeof(Real), cudaMemcpyHostToDevice));
    return 0;
}

template <typename Real, bool SPLIT>
static int prepare_kernel(rbc2d_sim* s)
{
    auto k = rbc2d_env_kernel<Real, SPLIT>;
    s->smem = SmemLayout<Real>::total;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem));
    int per_sm = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, NT, s->smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->cfg.device));
    if (per_sm < 1) return fail("step kernel does not fit on an SM");
    s->grid = per_sm * sms;
    return 0;
}

template <typename Real, bool SPLIT>
static int launch_env(rbc2d_sim* s, const float* actions, float* obs, float* reward, double* nu_s, double* nu_o, int* trunc,
                      int* nan, const int* env_ids, int n, RunFlags F, bool time_it, const VecIO& vec)
{
    Consts<Real> C = make_consts<Real>(s->hc, s->wr);
    Tables<Real> T{(const Real*)s->tinv, (const Real*)s->tw48, (const Real*)s->tw96,
                   (Real)((s->hc.lz / NZ) * (s->hc.lz / NZ) / 48.0)};
    EnvIO<Real> io;
    io.state = (Real*)s->state;
    io.actions = actions ? actions : s->actions;
    io.obs = obs ? obs : s->obs;
    io.reward = reward ? reward : s->reward;
    io.nu_state = nu_s ? nu_s : s->nu_s;
    io.nu_obs = nu_o ? nu_o : s->nu_o;
    io.t = s->t;
    io.step_count = s->step;
    io.truncated = trunc ? trunc : s->trunc;
    io.nan_flag = nan ? nan : s->nan;
    io.pressure = (Real*)s->pressure;
    io.cell_dist = s->wr.shaping ? s->cell_dist : nullptr;   // the peak scan runs only when reward shaping is on
    io.vec = vec;
    const int grid = n < s->grid ? n : s->grid;
    if (grid <= 0) return 0;
    const int slot = (int)(s->timed_launches % rbc2d_sim::kRing);
    if (time_it) CK(cudaEventRecord(s->ev0[slot], s->stream));
    rbc2d_env_kernel<Real, SPLIT><<<grid, NT, s->smem, s->stream>>>(C, T, io, (Real*)s->gm, (Real*)s->nxt, env_ids, n, F);
    CK(cudaGetLastError());
    if (time_it) { CK(cudaEventRecord(s->ev1[slot], s->stream)); s->timed_launches += 1; }
    s->launches += 1;
    return 0;
}

static int dispatch_env(rbc2d_sim* s, const float* actions, float* obs, float* reward, double* nu_s, double* nu_o, int* trunc,
                        int* nan, const int* env_ids, int n, RunFlags F, bool time_it, const VecIO& vec = VecIO())
{
    const bool f32 = s->cfg.precision == 32, split = s->cfg.pressure != 0;
    if (s->plan) {
        rbc2dx_api::IoRaw io;
        io.state = s->state;
        io.actions = actions ? actions : s->actions;
        io.obs = obs ? obs : s->obs;
        io.reward = reward ? reward : s->reward;
        io.nu_state = nu_s ? nu_s : s->nu_s;
        io.nu_obs = nu_o ? nu_o : s->nu_o;
        io.t = s->t;
        io.step_count = s->step;
        io.truncated = trunc ? trunc : s->trunc;
        io.nan_flag = nan ? nan : s->nan;
        io.cell_dist = s->wr.shaping ? s->cell_dist : nullptr;
        io.pressure = s->pressure;
        io.vec = vec;
        io.cfl_events = s->cfl_events;
        if (n <= 0) return 0;
        const int slot = (int)(s->timed_launches % rbc2d_sim::kRing);
        if (time_it) CK(cudaEventRecord(s->ev0[slot], s->stream));
        int rc = rbc2dx_api::launch(s->plan, s->hc, s->wr, io, env_ids, n, F, s->stream);
        if (rc) return rc;
        if (time_it) { CK(cudaEventRecord(s->ev1[slot], s->stream)); s->timed_launches += 1; }
        s->launches += 1;
        return 0;
    }
    if (f32) return split ? launch_env<float, true>(s, actions, obs, reward, nu_s, nu_o, trunc, nan, env_ids, n, F, time_it, vec)
                          : launch_env<float, false>(s, actions, obs, reward, nu_s, nu_o, trunc, nan, env_ids, n, F, time_it, vec);
    return split ? launch_env<double, true>(s, actions, obs, reward, nu_s, nu_o, trunc, nan, env_ids, n, F, time_it, vec)
                 : launch_env<double, false>(s, actions, obs, reward, nu_s, nu_o, trunc, nan, env_ids, n, F, time_it, vec);
}

int rbc_pipe_prepare(HostPipe* p, int B)
{
    if (p->copy && p->n == B) return 0;
    rbc_pipe_destroy(p);
    CK(cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking));
    for (int c = 0; c < HostPipe::kMaxChunks; ++c) CK(cudaEventCreateWithFlags(&p->done[c], cudaEventDisableTiming));
    std::vector<int> h(B);
    for (int i = 0; i < B; ++i) h[i] = i;
    CK(cudaMalloc((void**)&p->iota, (size_t)B * sizeof(int)));
    CK(cudaMemcpy(p->iota, h.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice));
    p->n = B;
    return 0;
}
void rbc_pipe_destroy(HostPipe* p)
{
    if (p->copy) cudaStreamDestroy(p->copy);
    for (int c = 0; c < HostPipe::kMaxChunks; ++c) if (p->done[c]) cudaEventDestroy(p->done[c]);
    if (p->iota) cudaFree(p->iota);
    *p = HostPipe();
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int rbc_abi_version(void) { return RBC_B200_ABI_VERSION; }
const char* rbc_last_error(void) { return g_err.c_str(); }

int rbc2d_create(const rbc2d_config* cfg, rbc2d_sim** out)
{
    if (!cfg || !out) return fail("rbc2d_create: null argument");
    *out = nullptr;
    const bool dedicated = (cfg->nx == NX && cfg->nz == NZ) && !rbc2dx_api::supported(cfg->nx, cfg->nz);
    if (!dedicated && !rbc2dx_api::supported(cfg->nx, cfg->nz))
        return fail("rbc2d_create: registered grids are (nx x nz) 96 x 64, 128 x 64, 192 x 128 and 64 x 32, 64 x 64, 96 x 32, 96 x 128, 128 x 32, 128 x 128, 192 x 64");
    if (cfg->num_envs < 1) return fail("rbc2d_create: num_envs must be >= 1");
    if (cfg->precision != 32 && cfg->precision != 64) return fail("rbc2d_create: precision must be 32 or 64");
    if (cfg->heaters < 1 || cfg->heaters > MAX_HEATERS) return fail("rbc2d_create: heaters must be in 1..32");
    if (cfg->obs_nx < 1 || cfg->obs_nz < 2 || cfg->nx % cfg->obs_nx || cfg->nz % cfg->obs_nz)
        return fail("rbc2d_create: sensors must divide the grid (and obs_nz >= 2)");
    if (!(cfg->ra > 0) || !(cfg->pr > 0) || !(cfg->dt_action > 0) || !(cfg->dt_solver > 0))
        return fail("rbc2d_create: ra, pr, dt_action, dt_solver must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail("rbc2d_create: no CUDA device available (this backend has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("rbc2d_create: bad device ordinal");
    CK(cudaSetDevice(cfg->device));

    rbc2d_sim* s = new rbc2d_sim();
    s->cfg = *cfg;
    s->B = cfg->num_envs;
    s->channels = cfg->pressure ? 5 : 3;
    s->real_size = cfg->precision == 32 ? 4 : 8;
    s->nx = cfg->nx; s->nz = cfg->nz; s->ncell = cfg->nx * cfg->nz; s->nwf = cfg->nx * (cfg->nz + 1); s->nstate = 2 * s->ncell + s->nwf;
    s->hc = HostConfig{cfg->ra, cfg->pr, 2.0 * 3.14159265358979323846, 2.0, 1.0, cfg->heater_limit, cfg->dt_action,
                       cfg->dt_solver, cfg->episode_length, cfg->heaters, cfg->obs_nz, cfg->obs_nx, s->channels,
                       // CFL guard: on in the fp32 throughput mode of the cluster kernels, off in the fp64 validation mode
                       (!dedicated && cfg->precision == 32) ? 1.4 : 0.0};
    int rc;
    const bool f32 = cfg->precision == 32, split = cfg->pressure != 0;
    if (dedicated) {
        rc = f32 ? upload_tables<float>(s) : upload_tables<double>(s);
        if (rc) { rbc2d_destroy(s); return rc; }
        if (f32) rc = split ? prepare_kernel<float, true>(s) : prepare_kernel<float, false>(s);
        else rc = split ? prepare_kernel<double, true>(s) : prepare_kernel<double, false>(s);
        if (rc) { rbc2d_destroy(s); return rc; }
    } else {
        rc = rbc2dx_api::create(cfg->nx, cfg->nz, cfg->precision, split ? 1 : 0, cfg->device, s->hc.lx, s->hc.lz, &s->plan);
        if (rc) { rbc2d_destroy(s); return rc; }
        s->smem = rbc2dx_api::smem_bytes(s->plan);
        s->grid = rbc2dx_api::grid_ctas(s->plan, 1 << 30);
    }

    const size_t B = s->B, rs = s->real_size;
    const size_t NSTATE = s->nstate, NCELL = s->ncell;
    const size_t nobs = (size_t)s->channels * cfg->obs_nz * cfg->obs_nx;
#define ALLOC(ptr, bytes)                                                         \
    do {                                                                          \
        cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                     \
        if (e_ != cudaSuccess) {                                                  \
            std::string m = std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_); \
            rbc2d_destroy(s);                                                     \
            return fail(m);                                                       \
        }                                                                         \
        cudaMemset((ptr), 0, (bytes));                                            \
    } while (0)
    ALLOC(s->state, B * NSTATE * rs);
    if (dedicated) {
        ALLOC(s->gm, (size_t)s->grid * 2 * NSTATE * rs);
        if (!f32) ALLOC(s->nxt, (size_t)s->grid * NS_SM * rs);
    }
    if (split) ALLOC(s->pressure, B * 2 * NCELL * rs);
    ALLOC(s->t, B * sizeof(double));
    ALLOC(s->nu_s, B * sizeof(double));
    ALLOC(s->nu_o, B * sizeof(double));
    ALLOC(s->step, B * sizeof(int));
    ALLOC(s->trunc, B * sizeof(int));
    ALLOC(s->nan, B * sizeof(int));
    ALLOC(s->obs, B * nobs * sizeof(float));
    ALLOC(s->reward, B * sizeof(float));
    ALLOC(s->actions, B * (size_t)cfg->heaters * sizeof(float));
    ALLOC(s->cell_dist, B * sizeof(double));
    ALLOC(s->pending, B * sizeof(int));
    ALLOC(s->episode, B * sizeof(long long));
    ALLOC(s->ep_return, B * sizeof(double));
    ALLOC(s->nan_count, sizeof(int));
    ALLOC(s->cfl_events, B * sizeof(int));
#undef ALLOC
    for (int i = 0; i < rbc2d_sim::kRing; ++i)
        if (cudaEventCreate(&s->ev0[i]) != cudaSuccess || cudaEventCreate(&s->ev1[i]) != cudaSuccess) {
            rbc2d_destroy(s);
            return fail("cudaEventCreate failed");
        }
    *out = s;
    return 0;
}

int rbc2d_destroy(rbc2d_sim* s)
{
    if (!s) return 0;
    cudaSetDevice(s->cfg.device);
    void* ptrs[] = {s->state, s->gm, s->nxt, s->pressure, s->tinv, s->tw48, s->tw96, s->bank, s->t, s->nu_s, s->nu_o,
                    s->step, s->trunc, s->nan, s->obs, s->reward, s->actions, s->cell_dist, s->pending, s->episode, s->ep_return,
                    s->nan_count, s->final_obs, s->final_scalars, s->cfl_events};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < rbc2d_sim::kRing; ++i) {
        if (s->ev0[i]) cudaEventDestroy(s->ev0[i]);
        if (s->ev1[i]) cudaEventDestroy(s->ev1[i]);
    }
    rbc2dx_api::destroy(s->plan);
    rbc_pipe_destroy(&s->pipe);
    delete s;
    return 0;
}

int rbc2d_set_stream(rbc2d_sim* s, void* stream)
{
    if (!s) return fail("null handle");
    s->stream = (cudaStream_t)stream;
    return 0;
}
int rbc2d_num_envs(const rbc2d_sim* s) { return s ? s->B : -1; }
int rbc2d_state_values_per_env(const rbc2d_sim* s) { return s ? s->nstate : -1; }

int rbc2d_load_checkpoints(rbc2d_sim* s, const double* b, const double* u, const double* w, int32_t n_ep)
{
    if (!s || !b || !u || !w || n_ep < 1) return fail("rbc2d_load_checkpoints: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    const size_t NSTATE = s->nstate, NCELL = s->ncell, NWF = s->nwf, GOFF_B = 0, GOFF_U = NCELL, GOFF_W = 2 * NCELL;
    if (s->bank) { cudaFree(s->bank); s->bank = nullptr; }
    CK(cudaMalloc((void**)&s->bank, (size_t)n_ep * NSTATE * sizeof(double)));
    for (int e = 0; e < n_ep; ++e) {
        double* d = s->bank + (size_t)e * NSTATE;
        CK(cudaMemcpyAsync(d + GOFF_B, b + (size_t)e * NCELL, NCELL * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(d + GOFF_U, u + (size_t)e * NCELL, NCELL * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(d + GOFF_W, w + (size_t)e * NWF, NWF * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    }
    CK(cudaStreamSynchronize(s->stream));
    s->n_ep = n_ep;
    return 0;
}

int rbc2d_reset_from_checkpoints_dev(rbc2d_sim* s, const int32_t* env_ids, const int32_t* ckpt_idx, int32_t n)
{
    if (!s || !ckpt_idx) return fail("rbc2d_reset_from_checkpoints_dev: bad argument");
    if (!s->bank) return fail("rbc2d_reset_from_checkpoints_dev: no checkpoint bank loaded");
    CK(cudaSetDevice(s->cfg.device));
    if (!env_ids) n = s->B;
    if (n <= 0) return 0;
    dim3 grid(8, n < 4096 ? n : 4096);
    if (s->cfg.precision == 32)
        rbc2d_reset_kernel<float><<<grid, 256, 0, s->stream>>>((float*)s->state, s->bank, env_ids, ckpt_idx, n, s->n_ep, s->t, s->step, s->trunc, s->nan, s->nstate, s->B);
    else
        rbc2d_reset_kernel<double><<<grid, 256, 0, s->stream>>>((double*)s->state, s->bank, env_ids, ckpt_idx, n, s->n_ep, s->t, s->step, s->trunc, s->nan, s->nstate, s->B);
    CK(cudaGetLastError());
    s->launches += 1;
    if (s->cfg.pressure) {   // set! projects and leaves pNHS; pHY' follows from b
        RunFlags F{0, 1, 0};
        return dispatch_env(s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, env_ids, n, F, false);
    }
    return 0;
}

int rbc2d_reset_from_fields_host(rbc2d_sim* s, const int32_t* env_ids_host, const double* fields, int32_t n, int32_t project)
{
    if (!s || !fields || n < 1 || n > s->B) return fail("rbc2d_reset_from_fields_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    double* dfields = nullptr;
    int* dids = nullptr;
    const size_t NSTATE = s->nstate;
    CK(cudaMalloc((void**)&dfields, (size_t)n * NSTATE * sizeof(double)));
    CK(cudaMemcpyAsync(dfields, fields, (size_t)n * NSTATE * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (env_ids_host) {
        CK(cudaMalloc((void**)&dids, n * sizeof(int)));
        CK(cudaMemcpyAsync(dids, env_ids_host, n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    }
    dim3 grid(8, n < 4096 ? n : 4096);
    if (s->cfg.precision == 32)
        rbc2d_set_fields_kernel<float><<<grid, 256, 0, s->stream>>>((float*)s->state, dfields, dids, n, s->t, s->step, s->trunc, s->nan, s->nstate, s->B);
    else
        rbc2d_set_fields_kernel<double><<<grid, 256, 0, s->stream>>>((double*)s->state, dfields, dids, n, s->t, s->step, s->trunc, s->nan, s->nstate, s->B);
    CK(cudaGetLastError());
    s->launches += 1;
    int rc = 0;
    if (project || s->cfg.pressure) {
        RunFlags F{0, 1, 0};
        rc = dispatch_env(s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dids, n, F, false);
    }
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(dfields);
    if (dids) cudaFree(dids);
    return rc;
}

int rbc2d_reset_from_fields_dev(rbc2d_sim* s, const int32_t* env_ids, const double* fields, int32_t n, int32_t project)
{
    if (!s || !fields || n < 1 || n > s->B) return fail("rbc2d_reset_from_fields_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    dim3 grid(8, n < 4096 ? n : 4096);
    if (s->cfg.precision == 32)
        rbc2d_set_fields_kernel<float><<<grid, 256, 0, s->stream>>>((float*)s->state, fields, env_ids, n, s->t, s->step, s->trunc, s->nan, s->nstate, s->B);
    else
        rbc2d_set_fields_kernel<double><<<grid, 256, 0, s->stream>>>((double*)s->state, fields, env_ids, n, s->t, s->step, s->trunc, s->nan, s->nstate, s->B);
    CK(cudaGetLastError());
    s->launches += 1;
    if (project || s->cfg.pressure) {
        RunFlags F{0, 1, 0};
        return dispatch_env(s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, env_ids, n, F, false);
    }
    return 0;
}

int rbc2d_step_dev(rbc2d_sim* s, const float* actions, float* obs, float* reward, double* nu_s, double* nu_o, int32_t* trunc,
                   int32_t* nan)
{
    if (!s || !actions) return fail("rbc2d_step_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    Consts<float> tmp = make_consts<float>(s->hc);
    RunFlags F{tmp.nsub, 0, 1};
    return dispatch_env(s, actions, obs, reward, nu_s, nu_o, trunc, nan, nullptr, s->B, F, true);
}

// host-buffer step shared by rbc2d_step_host and rbc2d_vec_step_host: `out` holds HOST pointers; with `vec` the launches carry
// the fused vector-env semantics and the final_* outputs pass through handle-owned device staging buffers
static int step_host_impl(rbc2d_sim* s, const float* actions, const rbc2d_vec_out& out, bool vec)
{
    CK(cudaSetDevice(s->cfg.device));
    const size_t B = s->B;
    const size_t nobs = (size_t)s->channels * s->cfg.obs_nz * s->cfg.obs_nx;
    CK(cudaMemcpyAsync(s->actions, actions, B * s->cfg.heaters * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    // chunks of whole waves of the persistent grid; the copy stream drains chunk c while chunk c+1 computes
    const int lanes = s->plan ? rbc2dx_api::grid_ctas(s->plan, 1 << 30) / rbc2dx_api::cluster_size(s->plan) : s->grid;
    int len[HostPipe::kMaxChunks];
    const int nch = rbc_pipe_chunks(s->B, lanes, len);
    int rc = rbc_pipe_prepare(&s->pipe, s->B);
    if (rc) return rc;
    VecIO V;
    const bool want_final = vec && (out.final_obs || out.final_nu_state || out.final_nu_obs || out.final_return);
    if (vec) {
        if (want_final && !s->final_obs) {
            CK(cudaMalloc((void**)&s->final_obs, B * nobs * sizeof(float)));
            CK(cudaMalloc((void**)&s->final_scalars, 3 * B * sizeof(double)));
        }
        rbc2d_vec_out dev = {};
        if (want_final) {
            dev.final_obs = s->final_obs; dev.final_nu_state = s->final_scalars; dev.final_nu_obs = s->final_scalars + B;
            dev.final_return = s->final_scalars + 2 * B;
        }
        V = make_vec(s, &dev);
    }
    Consts<float> tmp = make_consts<float>(s->hc);
    RunFlags F{tmp.nsub, 0, 1};
    const int slot = (int)(s->timed_launches % rbc2d_sim::kRing);
    CK(cudaEventRecord(s->ev0[slot], s->stream));             // the step's kernel time = span of its chunks
    size_t off = 0;
    for (int c = 0; c < nch; ++c) {
        const size_t n = (size_t)len[c];
        rc = dispatch_env(s, s->actions, s->obs, s->reward, s->nu_s, s->nu_o, s->trunc, s->nan, s->pipe.iota + off, (int)n, F, false, V);
        if (rc) return rc;
        if (c == nch - 1) { CK(cudaEventRecord(s->ev1[slot], s->stream)); s->timed_launches += 1; }
        CK(cudaEventRecord(s->pipe.done[c], s->stream));
        CK(cudaStreamWaitEvent(s->pipe.copy, s->pipe.done[c], 0));
        cudaStream_t cs = s->pipe.copy;
#define D2H(dst, src, count, type) if (dst) CK(cudaMemcpyAsync((dst) + off * (count), (src) + off * (count), n * (count) * sizeof(type), cudaMemcpyDeviceToHost, cs))
        D2H(out.obs, s->obs, nobs, float);
        D2H(out.reward, s->reward, 1, float);
        D2H(out.nu_state, s->nu_s, 1, double);
        D2H(out.nu_obs, s->nu_o, 1, double);
        D2H(out.truncated, s->trunc, 1, int);
        D2H(out.nan, s->nan, 1, int);
        if (vec) {
            D2H(out.t, s->t, 1, double);
            D2H(out.step, s->step, 1, int);
            D2H(out.episode_return, s->ep_return, 1, double);
            if (want_final) {
                D2H(out.final_obs, s->final_obs, nobs, float);
                D2H(out.final_nu_state, s->final_scalars, 1, double);
                D2H(out.final_nu_obs, s->final_scalars + B, 1, double);
                D2H(out.final_return, s->final_scalars + 2 * B, 1, double);
            }
        }
#undef D2H
        off += n;
    }
    CK(cudaStreamSynchronize(s->pipe.copy));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int rbc2d_step_host(rbc2d_sim* s, const float* actions, float* obs, float* reward, double* nu_s, double* nu_o, int32_t* trunc,
                    int32_t* nan)
{
    if (!s || !actions) return fail("rbc2d_step_host: bad argument");
    rbc2d_vec_out out = {};
    out.obs = obs; out.reward = reward; out.nu_state = nu_s; out.nu_obs = nu_o; out.truncated = trunc; out.nan = nan;
    return step_host_impl(s, actions, out, false);
}

int rbc2d_vec_step_host(rbc2d_sim* s, const float* actions, const rbc2d_vec_out* out)
{
    if (!s || !actions || !out) return fail("rbc2d_vec_step_host: bad argument");
    return step_host_impl(s, actions, *out, true);
}

int32_t rbc_checkpoint_draw(int64_t seed, int64_t global_env, int64_t episode, int32_t n_episodes)
{
    if (n_episodes < 1) return -1;
    return checkpoint_draw((unsigned long long)seed, (unsigned long long)global_env, (unsigned long long)episode, n_episodes);
}

int rbc2d_set_autoreset(rbc2d_sim* s, const rbc_autoreset* cfg)
{
    if (!s || !cfg) return fail("rbc2d_set_autoreset: bad argument");
    if (cfg->mode < 0 || cfg->mode > 2) return fail("rbc2d_set_autoreset: mode must be 0 (disabled), 1 (next_step) or 2 (same_step)");
    s->ar = *cfg;
    return 0;
}

// episode bookkeeping of a reset that the caller (or rbc2d_vec_reset_dev) performed on the fields
__global__ void rbc_vec_mark_kernel(const int* env_ids, int n, long long* episode, double* ep_return, int* pending, int restart)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int env = env_ids ? env_ids[j] : j;
        episode[env] = restart ? 1 : episode[env] + 1;
        ep_return[env] = 0.0;
        pending[env] = 0;
    }
}
// ckpt_idx[env] = draw(seed, offset + env, episode 0)
__global__ void rbc_vec_draw_kernel(int* idx, int n, unsigned long long seed, unsigned long long offset, int n_ep)
{
    for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < n; env += gridDim.x * blockDim.x)
        idx[env] = checkpoint_draw(seed, offset + (unsigned long long)env, 0ull, n_ep);
}

}  // extern "C"
int rbc_vec_mark(cudaStream_t stream, const int* env_ids, int n, long long* episode, double* ep_return, int* pending, int restart)
{
    if (n <= 0) return 0;
    rbc_vec_mark_kernel<<<(n + 255) / 256, 256, 0, stream>>>(env_ids, n, episode, ep_return, pending, restart);
    CK(cudaGetLastError());
    return 0;
}
int rbc_vec_draw(cudaStream_t stream, int* idx, int n, unsigned long long seed, unsigned long long offset, int n_ep)
{
    rbc_vec_draw_kernel<<<(n + 255) / 256, 256, 0, stream>>>(idx, n, seed, offset, n_ep);
    CK(cudaGetLastError());
    return 0;
}
extern "C" {

int rbc2d_vec_mark_reset_dev(rbc2d_sim* s, const int32_t* env_ids, int32_t n)
{
    if (!s) return fail("null handle");
    CK(cudaSetDevice(s->cfg.device));
    if (!env_ids) n = s->B;
    s->launches += 1;
    return rbc_vec_mark(s->stream, env_ids, n, s->episode, s->ep_return, s->pending, 0);
}

int rbc2d_vec_reset_dev(rbc2d_sim* s, const int32_t* ckpt_idx)
{
    if (!s) return fail("null handle");
    if (!s->bank) return fail("rbc2d_vec_reset_dev: no checkpoint bank loaded");
    CK(cudaSetDevice(s->cfg.device));
    if (!ckpt_idx) {
        // the draw of episode 0, into the pending array (idle until the mark kernel below clears it)
        int rcd = rbc_vec_draw(s->stream, s->pending, s->B, (unsigned long long)s->ar.seed, (unsigned long long)s->ar.env_id_offset, s->n_ep);
        if (rcd) return rcd;
        s->launches += 1;
        ckpt_idx = s->pending;
    }
    int rc = rbc2d_reset_from_checkpoints_dev(s, nullptr, ckpt_idx, s->B);
    if (rc) return rc;
    rc = rbc_vec_mark(s->stream, nullptr, s->B, s->episode, s->ep_return, s->pending, 1);
    if (rc) return rc;
    CK(cudaMemsetAsync(s->nan_count, 0, sizeof(int), s->stream));
    s->launches += 1;
    return 0;
}

int rbc2d_vec_step_dev(rbc2d_sim* s, const float* actions, const rbc2d_vec_out* out)
{
    if (!s || !actions || !out || !out->obs) return fail("rbc2d_vec_step_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    Consts<float> tmp = make_consts<float>(s->hc);
    RunFlags F{tmp.nsub, 0, 1};
    int rc = dispatch_env(s, actions, out->obs, out->reward, out->nu_state, out->nu_obs, out->truncated, out->nan, nullptr, s->B, F, true,
                          make_vec(s, out));
    if (rc) return rc;
    // info["t"], info["step"], running returns: plain device copies on the same stream, no synchronisation
    if (out->t) CK(cudaMemcpyAsync(out->t, s->t, (size_t)s->B * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    if (out->step) CK(cudaMemcpyAsync(out->step, s->step, (size_t)s->B * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
    if (out->episode_return) CK(cudaMemcpyAsync(out->episode_return, s->ep_return, (size_t)s->B * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}

int rbc2d_vec_nan_count(rbc2d_sim* s, int32_t clear, int64_t* count)
{
    if (!s || !count) return fail("rbc2d_vec_nan_count: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    int h = 0;
    CK(cudaMemcpyAsync(&h, s->nan_count, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if (clear) CK(cudaMemsetAsync(s->nan_count, 0, sizeof(int), s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *count = h;
    return 0;
}

int rbc2d_vec_nan_count_async(rbc2d_sim* s, int32_t* count_host)
{
    if (!s || !count_host) return fail("rbc2d_vec_nan_count_async: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(count_host, s->nan_count, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

int rbc2d_set_cfl_guard(rbc2d_sim* s, double limit)
{
    if (!s || !(limit >= 0)) return fail("rbc2d_set_cfl_guard: bad argument");
    if (limit > 0 && !s->plan) return fail("rbc2d_set_cfl_guard: the dedicated 96 x 64 kernel has no CFL guard (dt_solver 0.03 keeps CFL <= 1.2 there)");
    s->hc.cfl_limit = limit;
    return 0;
}

int rbc2d_get_cfl_events_host(rbc2d_sim* s, int32_t* out)
{
    if (!s || !out) return fail("rbc2d_get_cfl_events_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(out, s->cfl_events, (size_t)s->B * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int rbc2d_set_wrappers(rbc2d_sim* s, const rbc2d_wrappers* w)
{
    if (!s) return fail("null handle");
    s->wr = HostWrappers();
    if (!w) return 0;
    if (w->normalize_obs) {
        for (int c = 0; c < 4; ++c)
            if (!(w->obs_hi[c] > w->obs_lo[c])) return fail("rbc2d_set_wrappers: obs_hi must exceed obs_lo");
    }
    if (w->normalize_reward && w->reward_scale == 1.0) return fail("rbc2d_set_wrappers: reward_scale must differ from 1");
    s->wr.normalize_obs = w->normalize_obs; s->wr.obs_clip = w->obs_clip; s->wr.obs_maxval = w->obs_maxval;
    for (int c = 0; c < 4; ++c) { s->wr.obs_lo[c] = w->obs_lo[c]; s->wr.obs_hi[c] = w->obs_hi[c]; }
    s->wr.normalize_reward = w->normalize_reward; s->wr.reward_scale = w->reward_scale;
    s->wr.shaping = w->shaping; s->wr.shaping_weight = w->shaping_weight;
    return 0;
}

int rbc2d_get_cell_dist_host(rbc2d_sim* s, double* out)
{
    if (!s || !out) return fail("rbc2d_get_cell_dist_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(out, s->cell_dist, s->B * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int rbc2d_observe_dev(rbc2d_sim* s, float* obs, double* nu_s, double* nu_o)
{
    if (!s) return fail("null handle");
    CK(cudaSetDevice(s->cfg.device));
    RunFlags F{0, 0, 0};
    // truncated / nan flags of the handle are left untouched: write them to scratch copies
    return dispatch_env(s, nullptr, obs, nullptr, nu_s, nu_o, nullptr, nullptr, nullptr, s->B, F, false);
}

int rbc2d_observe_host(rbc2d_sim* s, float* obs, double* nu_s, double* nu_o)
{
    if (!s) return fail("null handle");
    int rc = rbc2d_observe_dev(s, s->obs, s->nu_s, s->nu_o);
    if (rc) return rc;
    const size_t B = s->B, nobs = (size_t)s->channels * s->cfg.obs_nz * s->cfg.obs_nx;
    if (obs) CK(cudaMemcpyAsync(obs, s->obs, B * nobs * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    if (nu_s) CK(cudaMemcpyAsync(nu_s, s->nu_s, B * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (nu_o) CK(cudaMemcpyAsync(nu_o, s->nu_o, B * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int rbc2d_get_state_dev(rbc2d_sim* s, float* out, int32_t channels)
{
    if (!s || !out) return fail("rbc2d_get_state_dev: bad argument");
    if (channels != 3 && channels != 5) return fail("rbc2d_get_state_dev: channels must be 3 or 5");
    if (channels == 5 && !s->pressure) return fail("rbc2d_get_state_dev: pressure channels need pressure=1 at creation");
    CK(cudaSetDevice(s->cfg.device));
    const int blocks = 148 * 8;
    if (s->cfg.precision == 32)
        rbc2d_get_state_kernel<float><<<blocks, 256, 0, s->stream>>>((const float*)s->state, (const float*)s->pressure, out, s->B, channels, s->ncell, s->nstate);
    else
        rbc2d_get_state_kernel<double><<<blocks, 256, 0, s->stream>>>((const double*)s->state, (const double*)s->pressure, out, s->B, channels, s->ncell, s->nstate);
    CK(cudaGetLastError());
    s->launches += 1;
    return 0;
}

int rbc2d_get_state_host(rbc2d_sim* s, float* out, int32_t channels)
{
    if (!s || !out) return fail("rbc2d_get_state_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    float* tmp = nullptr;
    const size_t bytes = (size_t)s->B * channels * s->ncell * sizeof(float);
    CK(cudaMalloc((void**)&tmp, bytes));
    int rc = rbc2d_get_state_dev(s, tmp, channels);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(out, tmp, bytes, cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) rc = fail(std::string("rbc2d_get_state_host: ") + cudaGetErrorString(e));
    }
    cudaFree(tmp);
    return rc;
}

int rbc2d_render_rgb_dev(rbc2d_sim* s, uint8_t* out)
{
    if (!s || !out) return fail("rbc2d_render_rgb_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    const float vmin = 1.0f, inv = 1.0f / (float)(1.0 + s->cfg.heater_limit);      // vmin=1, vmax=2+limit (rbc2D.py:244-249)
    if (s->cfg.precision == 32)
        rbc2d_render_kernel<float><<<148 * 8, 256, 0, s->stream>>>((const float*)s->state, out, s->B, s->nx, s->nz, s->nstate, vmin, inv);
    else
        rbc2d_render_kernel<double><<<148 * 8, 256, 0, s->stream>>>((const double*)s->state, out, s->B, s->nx, s->nz, s->nstate, vmin, inv);
    CK(cudaGetLastError());
    s->launches += 1;
    return 0;
}

int rbc2d_get_fields_host(rbc2d_sim* s, double* out)
{
    if (!s || !out) return fail("rbc2d_get_fields_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    const size_t total = (size_t)s->B * s->nstate;
    double* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, total * sizeof(double)));
    if (s->cfg.precision == 32) rbc2d_get_fields_kernel<float><<<148 * 8, 256, 0, s->stream>>>((const float*)s->state, tmp, total);
    else rbc2d_get_fields_kernel<double><<<148 * 8, 256, 0, s->stream>>>((const double*)s->state, tmp, total);
    s->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, tmp, total * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(std::string("rbc2d_get_fields_host: ") + cudaGetErrorString(e));
    return 0;
}

int rbc2d_get_info_host(rbc2d_sim* s, double* t, int32_t* step)
{
    if (!s) return fail("null handle");
    CK(cudaSetDevice(s->cfg.device));
    if (t) CK(cudaMemcpyAsync(t, s->t, s->B * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (step) CK(cudaMemcpyAsync(step, s->step, s->B * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int rbc2d_launch_count(const rbc2d_sim* s, int64_t* launches, int32_t* grid, int32_t* smem_bytes)
{
    if (!s) return fail("null handle");
    if (launches) *launches = s->launches;
    if (grid) *grid = s->plan ? rbc2dx_api::grid_ctas(s->plan, s->B) : (s->B < s->grid ? s->B : s->grid);
    if (smem_bytes) *smem_bytes = (int32_t)s->smem;
    return 0;
}

int rbc2d_last_step_kernel_ms(rbc2d_sim* s, float* ms)
{
    if (!s || !ms) return fail("rbc2d_last_step_kernel_ms: bad argument");
    return rbc2d_step_kernel_ms_history(s, ms, 1) == 1 ? 0 : -1;
}

int rbc2d_step_kernel_ms_history(rbc2d_sim* s, float* ms, int32_t n)
{
    if (!s || !ms || n < 1) return fail("rbc2d_step_kernel_ms_history: bad argument");
    if (s->timed_launches < 1) return fail("rbc2d_step_kernel_ms_history: no step has been launched");
    CK(cudaSetDevice(s->cfg.device));
    int64_t have = s->timed_launches < rbc2d_sim::kRing ? s->timed_launches : rbc2d_sim::kRing;
    if (n > have) n = (int32_t)have;
    for (int32_t i = 0; i < n; ++i) {                      // oldest of the requested launches first
        const int slot = (int)((s->timed_launches - n + i) % rbc2d_sim::kRing);
        CK(cudaEventSynchronize(s->ev1[slot]));
        CK(cudaEventElapsedTime(&ms[i], s->ev0[slot], s->ev1[slot]));
    }
    return n;
}

}  // extern "C"
