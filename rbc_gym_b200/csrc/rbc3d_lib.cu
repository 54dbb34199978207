// rbc3d_lib.cu — sm_100a kernel wrapper + C ABI of the 3D environment (rbc3d_* in include/rbc_b200.h).
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/rbc_b200.h"
#include "rbc3d_core.h"
#include "rbc3dg_api.h"
#include "rbc_common.h"

using namespace rbc3d;

template <typename Real, bool TILED>
struct Smem3 {
    static constexpr size_t R = 0;                       // the tile of the tiled tendency phase aliases the Poisson scratch
    static constexpr size_t Tb = R + sizeof(Real) * ((TILED && NTILE > NR) ? NTILE : NR);
    static constexpr size_t red = Tb + sizeof(Real) * NCOL;
    static constexpr size_t total = red + sizeof(double) * 2 * NT;
    static_assert(Tb % 16 == 0 && red % 16 == 0, "alignment");
    static_assert(total <= 232448, "exceeds the shared memory a CTA can opt into");
};

template <typename Real, bool SPLIT, bool TILED>
__global__ void __launch_bounds__(NT, 1)
rbc3d_env_kernel(Consts3<Real> C, EnvIO3<Real> io, Real* buf_all, Real* gm_all, const Real* tinv, Real thomas_scale,
                 const int* env_ids, int n, RunFlags3 F)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using L = Smem3<Real, TILED>;
    Ctx3<Real> X;
    X.bufA = buf_all + (size_t)blockIdx.x * 2 * NSTATE;
    X.bufB = X.bufA + NSTATE;
    X.gm = gm_all + (size_t)blockIdx.x * NG;
    X.R = reinterpret_cast<Real*>(smem + L::R);
    X.tile = TILED ? X.R : nullptr;
    X.Tb = reinterpret_cast<Real*>(smem + L::Tb);
    X.red = reinterpret_cast<double*>(smem + L::red);
    X.tinv = tinv;
    X.thomas_scale = thomas_scale;
    for (int q = threadIdx.x; q < NR; q += NT) X.R[q] = Real(0);
    __syncthreads();
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
        const int env = env_ids ? env_ids[j] : j;
        env_action_step3<Real, SPLIT, TILED>(C, io, X, env, F);
    }
}

template <typename Real>
__global__ void rbc3d_set_fields_kernel(Real* state, const double* fields, const int* env_ids, const int* src_idx, int n, int n_src,
                                        double* t, int* step, int* trunc, int* nan, int B, int NSTATE)
{
    // indices are validated by the caller (IndexError in the Python facade); out-of-range entries are skipped, never clamped
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        const int env = env_ids ? env_ids[j] : j;
        const int sidx = src_idx ? src_idx[j] : j;
        if (env < 0 || env >= B || sidx < 0 || sidx >= n_src) continue;
        const double* src = fields + (size_t)sidx * NSTATE;
        Real* dst = state + (size_t)env * NSTATE;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < NSTATE; q += gridDim.x * blockDim.x) dst[q] = (Real)src[q];
        if (blockIdx.x == 0 && threadIdx.x == 0) { t[env] = 0.0; step[env] = 1; trunc[env] = 0; nan[env] = 0; }
    }
}

template <typename Real>
__global__ void rbc3d_get_fields_kernel(const Real* state, double* out, size_t total)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x)
        out[q] = (double)state[q];
}

// render("rgb_array") of the 3D environment (rbc3D.py:247-318 renders the temperature with PyVista: volume rendering, turbo
// colormap, clim = temperature_difference, opacity "sigmoid_1", 800 x 608 window, default isometric camera).  The same picture
// ray-marched on the device, one thread per pixel: orthographic isometric view (camera along (1,1,1), z up), trilinear samples
// of b (flipped in y like the reference's np.flip(T, axis=1)), turbo polynomial, sigmoid opacity composited front to back over
// a white background.  Not pixel-identical to VTK's renderer (perspective, lighting, sampling distance differ).
template <typename Real>
__global__ void rbc3d_render_kernel(const Real* state, uint8_t* out, int B, int nx, int ny, int nz, int nstate, float lx, float ly, float lz,
                                    float vmin, float inv_range, int H, int W)
{
    const size_t total = (size_t)B * H * W;
    // orthonormal camera frame: view direction d (from the camera into the scene), right r, up u
    const float d0 = -0.57735027f, d1 = -0.57735027f, d2 = -0.57735027f;
    const float r0 = 0.70710678f, r1 = -0.70710678f, r2 = 0.0f;
    const float u0 = -0.40824829f, u1 = -0.40824829f, u2 = 0.81649658f;
    const float cx0 = 0.5f * lx, cx1 = 0.5f * ly, cx2 = 0.5f * lz;
    const float diag = sqrtf(lx * lx + ly * ly + lz * lz);
    const float scale = 1.05f * diag / (float)W;                       // scene units per pixel: the box diagonal fills the width
    const float dt = 0.5f * fminf(lx / nx, fminf(ly / ny, lz / nz));   // sampling distance: half a cell
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int env = (int)(q / ((size_t)H * W)), pix = (int)(q % ((size_t)H * W)), py = pix / W, px = pix % W;
        const float sx = (px + 0.5f - 0.5f * W) * scale, sy = (0.5f * H - py - 0.5f) * scale;
        // ray origin far in front of the box
        const float o0 = cx0 + sx * r0 + sy * u0 - d0 * diag, o1 = cx1 + sx * r1 + sy * u1 - d1 * diag, o2 = cx2 + sx * r2 + sy * u2 - d2 * diag;
        // slab test against [0,lx] x [0,ly] x [0,lz]
        float t0 = 0.0f, t1 = 3.0f * diag;
        const float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {lx, ly, lz}, o[3] = {o0, o1, o2}, dd[3] = {d0, d1, d2};
        for (int a = 0; a < 3; ++a) {
            const float ta = (lo[a] - o[a]) / dd[a], tb = (hi[a] - o[a]) / dd[a];
            t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
        }
        float cr = 0.f, cg = 0.f, cb = 0.f, trans = 1.0f;
        const Real* b = state + (size_t)env * nstate;
        for (float t = t0 + 0.5f * dt; t < t1 && trans > 0.02f; t += dt) {
            const float x = o0 + t * d0, y = o1 + t * d1, z = o2 + t * d2;
            // cell-centred trilinear sample, periodic in x and y, clamped in z; y flipped like the reference picture
            const float fx = x / lx * nx - 0.5f, fy = (ly - y) / ly * ny - 0.5f, fz = fminf(fmaxf(z / lz * nz - 0.5f, 0.f), (float)(nz - 1));
            const int i0 = (int)floorf(fx), j0 = (int)floorf(fy), k0 = min((int)fz, nz - 2 < 0 ? 0 : nz - 2);
            const float ax = fx - i0, ay = fy - j0, az = fz - k0;
            const int ia = ((i0 % nx) + nx) % nx, ib = (ia + 1) % nx, ja = ((j0 % ny) + ny) % ny, jb = (ja + 1) % ny, kb = min(k0 + 1, nz - 1);
            auto at = [&](int i, int j, int k) { return (float)b[((size_t)k * ny + j) * nx + i]; };
            const float v = (1 - az) * ((1 - ay) * ((1 - ax) * at(ia, ja, k0) + ax * at(ib, ja, k0)) + ay * ((1 - ax) * at(ia, jb, k0) + ax * at(ib, jb, k0))) +
                            az * ((1 - ay) * ((1 - ax) * at(ia, ja, kb) + ax * at(ib, ja, kb)) + ay * ((1 - ax) * at(ia, jb, kb) + ax * at(ib, jb, kb)));
            const float s = fminf(fmaxf((v - vmin) * inv_range, 0.0f), 1.0f);
            const float r = 0.13572138f + s * (4.61539260f + s * (-42.66032258f + s * (132.13108234f + s * (-152.94239396f + s * 59.28637943f))));
            const float g = 0.09140261f + s * (2.19418839f + s * (4.84296658f + s * (-14.18503333f + s * (4.27729857f + s * 2.82956604f))));
            const float bl = 0.10667330f + s * (12.64194608f + s * (-60.58204836f + s * (110.36276771f + s * (-89.90310912f + s * 27.34824973f))));
            // "sigmoid_1": opacity = sigmoid over [-1, 1] across the colour range; 0.35 of it per cell crossed, i.e. per two samples
            const float alpha_cell = 0.35f / (1.0f + __expf(-(2.0f * s - 1.0f)));
            const float a = 1.0f - sqrtf(1.0f - alpha_cell);
            cr += trans * a * fminf(fmaxf(r, 0.f), 1.f); cg += trans * a * fminf(fmaxf(g, 0.f), 1.f); cb += trans * a * fminf(fmaxf(bl, 0.f), 1.f);
            trans *= 1.0f - a;
        }
        out[3 * q + 0] = (uint8_t)(fminf(cr + trans, 1.0f) * 255.0f);
        out[3 * q + 1] = (uint8_t)(fminf(cg + trans, 1.0f) * 255.0f);
        out[3 * q + 2] = (uint8_t)(fminf(cb + trans, 1.0f) * 255.0f);
    }
}

struct rbc3d_sim {
    rbc3d_config cfg;
    HostConfig3 hc;
    rbc3dg_api::Plan* plan = nullptr;     // stage-streaming kernels for grids other than 32 x 32 x 16 (rbc3dg_lib.cu)
    int nstate = NSTATE, ncell = NC;
    int B = 0, grid = 0, n_ep = 0;
    size_t smem = 0, rs = 4;
    cudaStream_t stream = nullptr;
    void *state = nullptr, *buf = nullptr, *gm = nullptr, *tinv = nullptr;
    double *bank = nullptr, *t = nullptr, *nu = nullptr;
    int *step = nullptr, *trunc = nullptr, *nan = nullptr;
    float *obs = nullptr, *reward = nullptr, *actions = nullptr;
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    HostPipe pipe;                   // rbc3d_step_host
    // fused vector-env state (rbc3d_set_autoreset / rbc3d_vec_step_dev)
    rbc_autoreset ar = {0, 0, 0, 0};
    int* pending = nullptr;
    long long* episode = nullptr;
    double* ep_return = nullptr;
    int* nan_count = nullptr;
};

static rbc2d::VecIO make_vec3(const rbc3d_sim* s, const rbc3d_vec_out* out)
{
    rbc2d::VecIO v;
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
    if (out) { v.final_obs = out->final_obs; v.final_nu_a = out->final_nusselt; v.final_return = out->final_return; }
    return v;
}

// fp32 without the hydrostatic split runs the tiled tendency phase (the fp64 tile would not fit next to the scratch)
template <typename Real, bool SPLIT>
struct UseTile { static constexpr bool value = sizeof(Real) == 4 && !SPLIT; };

template <typename Real, bool SPLIT>
static int prepare3(rbc3d_sim* s)
{
    constexpr bool TILED = UseTile<Real, SPLIT>::value;
    auto k = rbc3d_env_kernel<Real, SPLIT, TILED>;
    s->smem = Smem3<Real, TILED>::total;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem));
    int per_sm = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, NT, s->smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->cfg.device));
    if (per_sm < 1) return rbc_fail("3D step kernel does not fit on an SM");
    s->grid = per_sm * sms;
    std::vector<double> tv((size_t)NZ * NCOL);
    build_tables3_host(s->hc.lx, s->hc.ly, s->hc.lz, tv.data());
    std::vector<Real> tr(tv.begin(), tv.end());
    CK(cudaMalloc(&s->tinv, tr.size() * sizeof(Real)));
    CK(cudaMemcpy(s->tinv, tr.data(), tr.size() * sizeof(Real), cudaMemcpyHostToDevice));
    return 0;
}

template <typename Real, bool SPLIT>
static int launch3(rbc3d_sim* s, const float* actions, float* obs, float* reward, double* nu, int* trunc, int* nan, const int* env_ids,
                   int n, RunFlags3 F, bool time_it, bool want_obs, const rbc2d::VecIO& vec)
{
    Consts3<Real> C = make_consts3<Real>(s->hc);
    EnvIO3<Real> io;
    io.state = (Real*)s->state;
    io.actions = actions ? actions : s->actions;
    io.obs = want_obs ? (obs ? obs : s->obs) : nullptr;
    io.reward = reward ? reward : s->reward;
    io.nusselt = nu ? nu : s->nu;
    io.t = s->t; io.step_count = s->step;
    io.truncated = trunc ? trunc : s->trunc;
    io.nan_flag = nan ? nan : s->nan;
    io.vec = vec;
    const int grid = n < s->grid ? n : s->grid;
    if (grid <= 0) return 0;
    if (time_it) CK(cudaEventRecord(s->ev0, s->stream));
    rbc3d_env_kernel<Real, SPLIT, UseTile<Real, SPLIT>::value><<<grid, NT, s->smem, s->stream>>>(C, io, (Real*)s->buf, (Real*)s->gm, (const Real*)s->tinv,
                                                                   (Real)((s->hc.lz / NZ) * (s->hc.lz / NZ) / 512.0), env_ids, n, F);
    CK(cudaGetLastError());
    if (time_it) { CK(cudaEventRecord(s->ev1, s->stream)); s->timed = true; }
    s->launches += 1;
    return 0;
}

static int dispatch3(rbc3d_sim* s, const float* actions, float* obs, float* reward, double* nu, int* trunc, int* nan, const int* env_ids,
                     int n, RunFlags3 F, bool time_it, bool want_obs = true, const rbc2d::VecIO& vec = rbc2d::VecIO())
{
    if (s->plan) {
        rbc3dg_api::IoRaw io;
        io.state = s->state;
        io.actions = actions ? actions : s->actions;
        io.obs = want_obs ? (obs ? obs : s->obs) : nullptr;
        io.reward = reward ? reward : s->reward;
        io.nusselt = nu ? nu : s->nu;
        io.t = s->t; io.step_count = s->step;
        io.truncated = trunc ? trunc : s->trunc;
        io.nan_flag = nan ? nan : s->nan;
        if (n <= 0) return 0;
        if (time_it) CK(cudaEventRecord(s->ev0, s->stream));
        int rc = rbc3dg_api::launch(s->plan, io, env_ids, n, F.nsub, F.project_first, F.advance_clock, s->stream, &s->launches, vec.mode >= 0 ? &vec : nullptr);
        if (rc) return rc;
        if (time_it) { CK(cudaEventRecord(s->ev1, s->stream)); s->timed = true; }
        return 0;
    }
    const bool f32 = s->cfg.precision == 32, split = s->cfg.split != 0;
    if (f32) return split ? launch3<float, true>(s, actions, obs, reward, nu, trunc, nan, env_ids, n, F, time_it, want_obs, vec)
                          : launch3<float, false>(s, actions, obs, reward, nu, trunc, nan, env_ids, n, F, time_it, want_obs, vec);
    return split ? launch3<double, true>(s, actions, obs, reward, nu, trunc, nan, env_ids, n, F, time_it, want_obs, vec)
                 : launch3<double, false>(s, actions, obs, reward, nu, trunc, nan, env_ids, n, F, time_it, want_obs, vec);
}

extern "C" {

int rbc3d_create(const rbc3d_config* cfg, rbc3d_sim** out)
{
    if (!cfg || !out) return rbc_fail("rbc3d_create: null argument");
    *out = nullptr;
    const char* force = getenv("RBC_B200_3D_GENERIC");
    const bool dedicated = cfg->nx == NX && cfg->ny == NY && cfg->nz == NZ && !(force && force[0] == '1');
    if (!dedicated && !rbc3dg_api::supported(cfg->nx, cfg->ny, cfg->nz))
        return rbc_fail("rbc3d_create: grid must have nx, ny powers of two in 8..256 and 6 <= nz <= 256");
    if (cfg->num_envs < 1) return rbc_fail("rbc3d_create: num_envs must be >= 1");
    if (cfg->precision != 32 && cfg->precision != 64) return rbc_fail("rbc3d_create: precision must be 32 or 64");
    if (cfg->heaters < 1 || cfg->heaters > MAX_HEATERS) return rbc_fail("rbc3d_create: heaters must be in 1..16");
    if (!(cfg->ra > 0) || !(cfg->pr > 0) || !(cfg->heater_duration > 0) || !(cfg->dt_solver > 0))
        return rbc_fail("rbc3d_create: ra, pr, heater_duration, dt_solver must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return rbc_fail("rbc3d_create: no CUDA device available (this backend has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return rbc_fail("rbc3d_create: bad device ordinal");
    CK(cudaSetDevice(cfg->device));
    rbc3d_sim* s = new rbc3d_sim();
    s->cfg = *cfg;
    s->B = cfg->num_envs;
    s->rs = cfg->precision == 32 ? 4 : 8;
    s->hc = HostConfig3{cfg->ra, cfg->pr, cfg->lx, cfg->ly, cfg->lz, cfg->b_min, cfg->b_max - cfg->b_min, cfg->heater_limit,
                        cfg->heater_duration, cfg->dt_solver, cfg->episode_length, cfg->heaters};
    const bool f32 = cfg->precision == 32, split = cfg->split != 0;
    int rc;
    if (dedicated) {
        rc = f32 ? (split ? prepare3<float, true>(s) : prepare3<float, false>(s))
                 : (split ? prepare3<double, true>(s) : prepare3<double, false>(s));
    } else {
        const rbc3dg::HostConfigG hg{cfg->ra, cfg->pr, cfg->lx, cfg->ly, cfg->lz, cfg->b_min, cfg->b_max - cfg->b_min, cfg->heater_limit,
                                     cfg->heater_duration, cfg->dt_solver, cfg->episode_length, cfg->heaters};
        rc = rbc3dg_api::create(hg, cfg->nx, cfg->ny, cfg->nz, cfg->num_envs, cfg->precision, cfg->device, &s->plan);
        if (!rc) {
            s->nstate = rbc3dg_api::values_per_env(s->plan);
            s->ncell = cfg->nx * cfg->ny * cfg->nz;
            s->smem = rbc3dg_api::smem_bytes(s->plan);
            s->grid = 1 << 30;                    // no persistent grid: every launch covers the environments it is given
        }
    }
    if (rc) { rbc3d_destroy(s); return rc; }
    const size_t B = s->B, rs = s->rs;
    const size_t NSTATE = s->nstate;              // shadows the compile-time constant of the dedicated kernel from here on
#define ALLOC3(ptr, bytes)                                                                      \
    do {                                                                                        \
        cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                                   \
        if (e_ != cudaSuccess) {                                                                \
            std::string m = std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_);      \
            rbc3d_destroy(s);                                                                   \
            return rbc_fail(m);                                                                 \
        }                                                                                       \
        cudaMemset((ptr), 0, (bytes));                                                          \
    } while (0)
    ALLOC3(s->state, B * NSTATE * rs);
    if (dedicated) {
        ALLOC3(s->buf, (size_t)s->grid * 2 * NSTATE * rs);
        ALLOC3(s->gm, (size_t)s->grid * NG * rs);
    }
    ALLOC3(s->t, B * sizeof(double));
    ALLOC3(s->nu, B * sizeof(double));
    ALLOC3(s->step, B * sizeof(int));
    ALLOC3(s->trunc, B * sizeof(int));
    ALLOC3(s->nan, B * sizeof(int));
    ALLOC3(s->reward, B * sizeof(float));
    ALLOC3(s->actions, B * (size_t)cfg->heaters * cfg->heaters * sizeof(float));
    ALLOC3(s->pending, B * sizeof(int));
    ALLOC3(s->episode, B * sizeof(long long));
    ALLOC3(s->ep_return, B * sizeof(double));
    ALLOC3(s->nan_count, sizeof(int));
#undef ALLOC3
    if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) {
        rbc3d_destroy(s);
        return rbc_fail("cudaEventCreate failed");
    }
    *out = s;
    return 0;
}

int rbc3d_destroy(rbc3d_sim* s)
{
    if (!s) return 0;
    cudaSetDevice(s->cfg.device);
    void* ptrs[] = {s->state, s->buf, s->gm, s->tinv, s->bank, s->t, s->nu, s->step, s->trunc, s->nan, s->obs, s->reward, s->actions, s->pending, s->episode,
                    s->ep_return, s->nan_count};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    rbc3dg_api::destroy(s->plan);
    rbc_pipe_destroy(&s->pipe);
    delete s;
    return 0;
}

int rbc3d_set_stream(rbc3d_sim* s, void* stream)
{
    if (!s) return rbc_fail("null handle");
    s->stream = (cudaStream_t)stream;
    return 0;
}
int rbc3d_state_values_per_env(const rbc3d_sim* s) { return s ? s->nstate : -1; }

int rbc3d_load_checkpoints(rbc3d_sim* s, const double* fields_host, int32_t n_episodes)
{
    if (!s || !fields_host || n_episodes < 1) return rbc_fail("rbc3d_load_checkpoints: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    if (s->bank) { cudaFree(s->bank); s->bank = nullptr; }
    CK(cudaMalloc((void**)&s->bank, (size_t)n_episodes * s->nstate * sizeof(double)));
    CK(cudaMemcpyAsync(s->bank, fields_host, (size_t)n_episodes * s->nstate * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->n_ep = n_episodes;
    return 0;
}

static int set_fields3(rbc3d_sim* s, const double* dfields, const int* env_ids_dev, const int* src_idx_dev, int n, int n_src)
{
    dim3 grid(16, n < 2048 ? n : 2048);
    if (s->cfg.precision == 32)
        rbc3d_set_fields_kernel<float><<<grid, 256, 0, s->stream>>>((float*)s->state, dfields, env_ids_dev, src_idx_dev, n, n_src, s->t, s->step, s->trunc, s->nan, s->B, s->nstate);
    else
        rbc3d_set_fields_kernel<double><<<grid, 256, 0, s->stream>>>((double*)s->state, dfields, env_ids_dev, src_idx_dev, n, n_src, s->t, s->step, s->trunc, s->nan, s->B, s->nstate);
    CK(cudaGetLastError());
    s->launches += 1;
    return 0;
}

int rbc3d_reset_from_checkpoints_dev(rbc3d_sim* s, const int32_t* env_ids, const int32_t* ckpt_idx, int32_t n)
{
    if (!s || !ckpt_idx) return rbc_fail("rbc3d_reset_from_checkpoints_dev: bad argument");
    if (!s->bank) return rbc_fail("rbc3d_reset_from_checkpoints_dev: no checkpoint bank loaded");
    CK(cudaSetDevice(s->cfg.device));
    if (!env_ids) n = s->B;
    if (n <= 0) return 0;
    return set_fields3(s, s->bank, env_ids, ckpt_idx, n, s->n_ep);
}

int rbc3d_reset_from_fields_host(rbc3d_sim* s, const int32_t* env_ids_host, const double* fields, int32_t n, int32_t project)
{
    if (!s || !fields || n < 1 || n > s->B) return rbc_fail("rbc3d_reset_from_fields_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    double* dfields = nullptr;
    int* dids = nullptr;
    CK(cudaMalloc((void**)&dfields, (size_t)n * s->nstate * sizeof(double)));
    CK(cudaMemcpyAsync(dfields, fields, (size_t)n * s->nstate * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (env_ids_host) {
        CK(cudaMalloc((void**)&dids, n * sizeof(int)));
        CK(cudaMemcpyAsync(dids, env_ids_host, n * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    }
    int rc = set_fields3(s, dfields, dids, nullptr, n, n);
    if (!rc && project) {
        RunFlags3 F{0, 1, 0};
        rc = dispatch3(s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dids, n, F, false, false);
    }
    cudaStreamSynchronize(s->stream);
    cudaFree(dfields);
    if (dids) cudaFree(dids);
    return rc;
}

int rbc3d_reset_from_fields_dev(rbc3d_sim* s, const int32_t* env_ids, const double* fields, int32_t n, int32_t project)
{
    if (!s || !fields || n < 1 || n > s->B) return rbc_fail("rbc3d_reset_from_fields_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    int rc = set_fields3(s, fields, env_ids, nullptr, n, n);
    if (!rc && project) {
        RunFlags3 F{0, 1, 0};
        rc = dispatch3(s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, env_ids, n, F, false, false);
    }
    return rc;
}

int rbc3d_step_dev(rbc3d_sim* s, const float* actions, float* obs, float* reward, double* nusselt, int32_t* trunc, int32_t* nan)
{
    if (!s || !actions) return rbc_fail("rbc3d_step_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    Consts3<float> tmp = make_consts3<float>(s->hc);
    RunFlags3 F{tmp.nsub, 0, 1};
    return dispatch3(s, actions, obs, reward, nusselt, trunc, nan, nullptr, s->B, F, true, obs != nullptr);
}

int rbc3d_set_rayleigh_per_env(rbc3d_sim* s, const double* ra_host)
{
    if (!s) return rbc_fail("null handle");
    if (!s->plan)
        return rbc_fail("rbc3d_set_rayleigh_per_env: the dedicated 32 x 32 x 16 kernel takes one Rayleigh number per handle; "
                        "set RBC_B200_3D_GENERIC=1 to run that grid through the stage-streaming kernels");
    return rbc3dg_api::set_rayleigh(s->plan, ra_host);
}

int rbc3d_set_autoreset(rbc3d_sim* s, const rbc_autoreset* cfg)
{
    if (!s || !cfg) return rbc_fail("rbc3d_set_autoreset: bad argument");
    if (cfg->mode < 0 || cfg->mode > 2) return rbc_fail("rbc3d_set_autoreset: mode must be 0 (disabled), 1 (next_step) or 2 (same_step)");
    s->ar = *cfg;
    return 0;
}

int rbc3d_vec_mark_reset_dev(rbc3d_sim* s, const int32_t* env_ids, int32_t n)
{
    if (!s) return rbc_fail("null handle");
    CK(cudaSetDevice(s->cfg.device));
    if (!env_ids) n = s->B;
    s->launches += 1;
    return rbc_vec_mark(s->stream, env_ids, n, s->episode, s->ep_return, s->pending, 0);
}

int rbc3d_vec_reset_dev(rbc3d_sim* s, const int32_t* ckpt_idx)
{
    if (!s) return rbc_fail("null handle");
    if (!s->bank) return rbc_fail("rbc3d_vec_reset_dev: no checkpoint bank loaded");
    CK(cudaSetDevice(s->cfg.device));
    if (!ckpt_idx) {
        int rcd = rbc_vec_draw(s->stream, s->pending, s->B, (unsigned long long)s->ar.seed, (unsigned long long)s->ar.env_id_offset, s->n_ep);
        if (rcd) return rcd;
        s->launches += 1;
        ckpt_idx = s->pending;
    }
    int rc = rbc3d_reset_from_checkpoints_dev(s, nullptr, ckpt_idx, s->B);
    if (rc) return rc;
    rc = rbc_vec_mark(s->stream, nullptr, s->B, s->episode, s->ep_return, s->pending, 1);
    if (rc) return rc;
    CK(cudaMemsetAsync(s->nan_count, 0, sizeof(int), s->stream));
    s->launches += 1;
    return 0;
}

int rbc3d_vec_step_dev(rbc3d_sim* s, const float* actions, const rbc3d_vec_out* out)
{
    if (!s || !actions || !out) return rbc_fail("rbc3d_vec_step_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    Consts3<float> tmp = make_consts3<float>(s->hc);
    RunFlags3 F{tmp.nsub, 0, 1};
    int rc = dispatch3(s, actions, out->obs, out->reward, out->nusselt, out->truncated, out->nan, nullptr, s->B, F, true, out->obs != nullptr,
                       make_vec3(s, out));
    if (rc) return rc;
    if (out->t) CK(cudaMemcpyAsync(out->t, s->t, (size_t)s->B * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    if (out->step) CK(cudaMemcpyAsync(out->step, s->step, (size_t)s->B * sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
    if (out->episode_return) CK(cudaMemcpyAsync(out->episode_return, s->ep_return, (size_t)s->B * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}

int rbc3d_vec_nan_count(rbc3d_sim* s, int32_t clear, int64_t* count)
{
    if (!s || !count) return rbc_fail("rbc3d_vec_nan_count: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    int h = 0;
    CK(cudaMemcpyAsync(&h, s->nan_count, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if (clear) CK(cudaMemsetAsync(s->nan_count, 0, sizeof(int), s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *count = h;
    return 0;
}

int rbc3d_vec_nan_count_async(rbc3d_sim* s, int32_t* count_host)
{
    if (!s || !count_host) return rbc_fail("rbc3d_vec_nan_count_async: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(count_host, s->nan_count, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

int rbc3d_step_host(rbc3d_sim* s, const float* actions, float* obs, float* reward, double* nusselt, int32_t* trunc, int32_t* nan)
{
    if (!s || !actions) return rbc_fail("rbc3d_step_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    const size_t B = s->B, na = (size_t)s->cfg.heaters * s->cfg.heaters, nobs = 4 * (size_t)s->ncell;
    if (obs && !s->obs) CK(cudaMalloc((void**)&s->obs, B * nobs * sizeof(float)));
    CK(cudaMemcpyAsync(s->actions, actions, B * na * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    // chunks of whole waves of the persistent grid; the copy stream drains chunk c (262 KB of observation per environment)
    // while chunk c+1 computes
    int len[HostPipe::kMaxChunks];
    const int nch = rbc_pipe_chunks(s->B, s->grid, len);
    int rc = rbc_pipe_prepare(&s->pipe, s->B);
    if (rc) return rc;
    Consts3<float> tmp = make_consts3<float>(s->hc);
    RunFlags3 F{tmp.nsub, 0, 1};
    CK(cudaEventRecord(s->ev0, s->stream));
    size_t off = 0;
    for (int c = 0; c < nch; ++c) {
        const size_t n = (size_t)len[c];
        rc = dispatch3(s, s->actions, obs ? s->obs : nullptr, s->reward, s->nu, s->trunc, s->nan, s->pipe.iota + off, (int)n, F, false, obs != nullptr);
        if (rc) return rc;
        if (c == nch - 1) { CK(cudaEventRecord(s->ev1, s->stream)); s->timed = true; }
        CK(cudaEventRecord(s->pipe.done[c], s->stream));
        CK(cudaStreamWaitEvent(s->pipe.copy, s->pipe.done[c], 0));
        cudaStream_t cs = s->pipe.copy;
        if (obs) CK(cudaMemcpyAsync(obs + off * nobs, s->obs + off * nobs, n * nobs * sizeof(float), cudaMemcpyDeviceToHost, cs));
        if (reward) CK(cudaMemcpyAsync(reward + off, s->reward + off, n * sizeof(float), cudaMemcpyDeviceToHost, cs));
        if (nusselt) CK(cudaMemcpyAsync(nusselt + off, s->nu + off, n * sizeof(double), cudaMemcpyDeviceToHost, cs));
        if (trunc) CK(cudaMemcpyAsync(trunc + off, s->trunc + off, n * sizeof(int), cudaMemcpyDeviceToHost, cs));
        if (nan) CK(cudaMemcpyAsync(nan + off, s->nan + off, n * sizeof(int), cudaMemcpyDeviceToHost, cs));
        off += n;
    }
    CK(cudaStreamSynchronize(s->pipe.copy));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int rbc3d_observe_dev(rbc3d_sim* s, float* obs, double* nusselt)
{
    if (!s) return rbc_fail("null handle");
    CK(cudaSetDevice(s->cfg.device));
    RunFlags3 F{0, 0, 0};
    return dispatch3(s, nullptr, obs, nullptr, nusselt, nullptr, nullptr, nullptr, s->B, F, false, obs != nullptr);
}

int rbc3d_render_rgb_dev(rbc3d_sim* s, uint8_t* out, int32_t height, int32_t width)
{
    if (!s || !out || height < 8 || width < 8) return rbc_fail("rbc3d_render_rgb_dev: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    const float vmin = (float)s->cfg.b_min, inv = 1.0f / (float)(s->cfg.b_max - s->cfg.b_min);          // clim = temperature_difference
    const int blocks = 148 * 8;
    if (s->cfg.precision == 32)
        rbc3d_render_kernel<float><<<blocks, 256, 0, s->stream>>>((const float*)s->state, out, s->B, s->cfg.nx, s->cfg.ny, s->cfg.nz, s->nstate,
                                                                   (float)s->cfg.lx, (float)s->cfg.ly, (float)s->cfg.lz, vmin, inv, height, width);
    else
        rbc3d_render_kernel<double><<<blocks, 256, 0, s->stream>>>((const double*)s->state, out, s->B, s->cfg.nx, s->cfg.ny, s->cfg.nz, s->nstate,
                                                                    (float)s->cfg.lx, (float)s->cfg.ly, (float)s->cfg.lz, vmin, inv, height, width);
    CK(cudaGetLastError());
    s->launches += 1;
    return 0;
}

int rbc3d_get_fields_host(rbc3d_sim* s, double* out)
{
    if (!s || !out) return rbc_fail("rbc3d_get_fields_host: bad argument");
    CK(cudaSetDevice(s->cfg.device));
    const size_t total = (size_t)s->B * s->nstate;
    double* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, total * sizeof(double)));
    if (s->cfg.precision == 32) rbc3d_get_fields_kernel<float><<<148 * 8, 256, 0, s->stream>>>((const float*)s->state, tmp, total);
    else rbc3d_get_fields_kernel<double><<<148 * 8, 256, 0, s->stream>>>((const double*)s->state, tmp, total);
    s->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, tmp, total * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return rbc_fail(std::string("rbc3d_get_fields_host: ") + cudaGetErrorString(e));
    return 0;
}

int rbc3d_get_info_host(rbc3d_sim* s, double* t, int32_t* step)
{
    if (!s) return rbc_fail("null handle");
    CK(cudaSetDevice(s->cfg.device));
    if (t) CK(cudaMemcpyAsync(t, s->t, s->B * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (step) CK(cudaMemcpyAsync(step, s->step, s->B * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int rbc3d_launch_count(const rbc3d_sim* s, int64_t* launches, int32_t* grid, int32_t* smem_bytes)
{
    if (!s) return rbc_fail("null handle");
    if (launches) *launches = s->launches;
    if (grid) *grid = s->B < s->grid ? s->B : s->grid;
    if (smem_bytes) *smem_bytes = (int32_t)s->smem;
    return 0;
}

int rbc3d_last_step_kernel_ms(rbc3d_sim* s, float* ms)
{
    if (!s || !ms) return rbc_fail("rbc3d_last_step_kernel_ms: bad argument");
    if (!s->timed) return rbc_fail("rbc3d_last_step_kernel_ms: no step has been launched");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaEventSynchronize(s->ev1));
    CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return 0;
}

}  // extern "C"
