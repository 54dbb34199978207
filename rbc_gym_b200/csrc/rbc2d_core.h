// rbc2d_core.h — the 2D Rayleigh-Benard action step as one CTA-per-environment program.
//
// Replaces, for the hot path, the reference's
//   step_simulation -> run!(simulation)            src/rbc_gym/sim/rbc_sim2D_api.jl:75-97
//   NonhydrostaticModel configuration               src/rbc_gym/sim/rbc_sim2D.jl:149-160
//   bottom_T / collate_actions_colin                src/rbc_gym/sim/rbc_sim2D.jl:87-138
//   get_state / get_observation / get_nusselt       src/rbc_gym/sim/rbc_sim2D_api.jl:102-163
//   step_contains_NaNs                              src/rbc_gym/sim/rbc_sim2D.jl:223-228
// (the PDE arithmetic of run! lives in Oceananigans.jl 0.92.0; scheme per SURVEY.md 8a).
//
// The file is plain C++ shared between nvcc (the sm_100a kernel in rbc2d_kernels.cu) and g++
// (tests/emu: a sequential "one CTA" emulator used to check the exact kernel logic against the
// fp64 oracle on a machine without a GPU).  A "phase" is a piece of per-thread work followed by
// a CTA-wide barrier; RBC_PHASE runs it for threadIdx.x on the device and for tid = 0..NT-1 on
// the host.
//
// Data layout (one environment): b[NZ][NX] centres, u[NZ][NX] left x-faces, w[NZ+1][NX] bottom
// z-faces — exactly the checkpoint layout, x fastest.  One CTA keeps the whole
// environment in shared memory for all RK3 stages of an action step.
//
// Thread map: NT = 384 = 96 columns x 4 strips of 16 rows.  In the tendency phase a thread
// marches up its strip holding 7-row sliding windows of its own column in registers, so z-fluxes
// are computed once and carried; x-neighbours come from shared memory (consecutive lanes ->
// consecutive words, conflict-free; 96 = 3 x 32 keeps periodic wrap on the same banks).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RBC_HD __host__ __device__ __forceinline__
#else
#define RBC_HD inline
#endif
#define RBC_RESTRICT __restrict__
// out-of-line on the device: code that runs once per environment visit and must not share the register budget of the march
#if defined(__CUDACC__)
#define RBC_HD_COLD __host__ __device__ __noinline__
#else
#define RBC_HD_COLD inline
#endif

#if defined(__CUDA_ARCH__)
#define RBC_PHASE_N(NTHREADS, ...) { const int tid = threadIdx.x; __VA_ARGS__ } __syncthreads();
#define RBC_UNROLL _Pragma("unroll")
#define RBC_NOUNROLL _Pragma("unroll 1")
#define RBC_UNROLL2 _Pragma("unroll 2")
#else
#define RBC_PHASE_N(NTHREADS, ...) for (int tid = 0; tid < (NTHREADS); ++tid) { __VA_ARGS__ }
#define RBC_UNROLL
#define RBC_NOUNROLL
#define RBC_UNROLL2
#endif
#define RBC_PHASE(...) RBC_PHASE_N(rbc2d::NT, __VA_ARGS__)

namespace rbc2d {

constexpr int NX = 96, NZ = 64;
constexpr int NSTRIP = 4, RS = NZ / NSTRIP;      // 4 strips of 16 rows
constexpr int NT = NX * NSTRIP;                  // 384 threads
constexpr int NCELL = NX * NZ;                   // 6144
constexpr int NWF = NX * (NZ + 1);               // 6240 (w incl. both wall faces)
constexpr int NSTATE = 2 * NCELL + NWF;          // 18528 values per environment (global / checkpoint layout)
// On-chip rows are padded to SX = 98 words: with lanes mapped to consecutive ROWS (FFT passes) a 64-bit
// access then walks the banks two at a time (98 = 3*32 + 2), and with lanes mapped to consecutive
// COLUMNS (stencils, Thomas) any stride is conflict-free.
constexpr int SX = 98;
constexpr int NROWS = 3 * NZ + 1;                // 193 rows: b (64), u (64), w (65)
constexpr int NS_SM = NROWS * SX;                // 18914 words per on-chip state buffer
constexpr int OFF_B = 0, OFF_U = NZ * SX, OFF_W = 2 * NZ * SX;       // on-chip (padded) field offsets
constexpr int GOFF_B = 0, GOFF_U = NCELL, GOFF_W = 2 * NCELL;         // global / checkpoint layout offsets
constexpr int RSTR = SX;                         // row stride (words) of the Poisson scratch
constexpr int NR = NZ * RSTR;                    // scratch size in words
constexpr int NH = NX / 2;                       // 48 complex points per packed row
constexpr int MAX_HEATERS = 32;
constexpr int NRED = 7;                          // partial sums per thread in the epilogue

// ------------------------------------------------------------------------------------------
// constants of one simulation configuration (host-computed in fp64, then cast)
// ------------------------------------------------------------------------------------------
template <typename Real>
struct Consts {
    Real idx, idz, idx2, idz2;   // 1/dx, 1/dz, 1/dx^2, 1/dz^2
    Real nu, kappa, b_top;
    Real dt_full, dt_last;       // run!'s substep schedule: (nsub-1) x dt_full + dt_last
    int nsub;
    int heaters;
    double heater_limit, lx, dx; // heater profile is evaluated in fp64 like the reference
    double kappa_d, dt_action, episode_length;
    int obs_nz, obs_nx;          // sensors (8, 48)
    int channels;                // 3, or 5 with pressure
    // wrappers fused into the epilogue (src/rbc_gym/wrappers/), applied in the order of example/run_wrapped.py:
    // RBCNormalizeObservation -> RBCNormalizeReward -> RBCRewardShaping
    int wrap_obs, obs_clip;      // rbc_normalize_observation.py:64-74
    float obs_lo[4], obs_hi[4], obs_maxval;
    int wrap_reward;             // rbc_normalize_reward.py:27-32: (r + s)/(s - 1)
    double reward_scale;
    int wrap_shaping;            // rbc_reward_shaping.py:53-140: (1-w) r + w (pi - cd)/pi
    double shaping_weight;
    // CFL guard of the cluster kernels (fp32 throughput mode): an RK3 step whose max(|w| dt/dz, |u| dt/dx) exceeds this limit is
    // taken as ceil(CFL) shorter steps by that environment; 0 = off (the reference's fixed dt, always in the fp64 mode)
    Real cfl_limit;
};

// tables in global memory, staged to shared memory by the CTA
template <typename Real>
struct Tables {
    const Real* tinv;   // [NZ][NX]  Thomas pivots per spectral word (see build_tables)
    const Real* tw48;   // [48][2]   cos/sin(2 pi j / 48)
    const Real* tw96;   // [48][2]   cos/sin(2 pi m / 96)
    Real thomas_scale;  // dz^2 / 48
};

// ------------------------------------------------------------------------------------------
// Vector-environment semantics fused into the step kernel: what gymnasium's / SB3's vector wrappers do around the
// reference's one-process-per-env setup (example/run_vectorized.py:11-31, experiments/run_sarl.py:130-153) — auto-reset
// of truncated environments from the device checkpoint bank, episode returns, terminal observations — happens inside
// the launch that steps the batch.  mode < 0: plain batched step (no vector semantics).
// ------------------------------------------------------------------------------------------
struct VecIO {
    int mode = -1;                 // 0 disabled (flags + returns only), 1 next_step (gymnasium 1.1.1), 2 same_step (SB3)
    int nan_reset = 0;             // 1: an environment with NaNs is re-initialised inside the step (truncated, reward 0)
    const double* bank = nullptr;  // [n_ep][NSTATE] fp64 checkpoint bank
    int n_ep = 0;
    unsigned long long seed = 0, id_offset = 0;   // checkpoint draw = hash(seed, id_offset + env, episode) mod n_ep
    int* pending = nullptr;        // [B] next_step: truncated on the previous call, this call only resets
    long long* episode = nullptr;  // [B] resets so far (the counter that keys the draw)
    double* ep_return = nullptr;   // [B] running episode return
    float* final_obs = nullptr;    // [B][obs...] terminal observation of environments reset inside this step
    double* final_nu_a = nullptr;  // [B] terminal nusselt_state (2D) / nusselt (3D)
    double* final_nu_b = nullptr;  // [B] terminal nusselt_obs (2D)
    double* final_return = nullptr;  // [B] return of the episode that ended
    int* nan_count = nullptr;      // one counter: environments that reported NaNs since it was last cleared
};

// Episode index for global environment g at episode counter e (same arithmetic as the 64-bit tensor expression the
// Python vector env used before the fusion, so sharded and unsharded runs draw alike)
RBC_HD int checkpoint_draw(unsigned long long seed, unsigned long long g, unsigned long long e, int n_ep)
{
    const unsigned long long M63 = 0x7FFFFFFFFFFFFFFFull;
    unsigned long long x = (g * 0x9E3779B97F4A7C15ull + e * 0xC2B2AE3D27D4EB4Full + seed * 0x165667B19E3779F9ull) & M63;
    x = ((x ^ (x >> 31)) * 0x7FB5D329728EA185ull) & M63;
    x = x ^ (x >> 27);
    return (int)(x % (unsigned long long)n_ep);
}

template <typename Real>
struct EnvIO {
    // per-environment global arrays, indexed by env
    Real* state;              // [B][NSTATE]
    const float* actions;     // [B][heaters]
    float* obs;               // [B][channels][obs_nz][obs_nx]
    float* reward;            // [B]
    double* nu_state;         // [B]
    double* nu_obs;           // [B]
    double* t;                // [B] simulation time of the episode
    int* step_count;          // [B]
    int* truncated;           // [B]
    int* nan_flag;            // [B]
    Real* pressure;           // [B][2][NZ][NX] (pHY', pNHS) or nullptr
    double* cell_dist;        // [B] Benard-cell distance (info["cell_dist"]) or nullptr
    VecIO vec;                // fused auto-reset (off by default)
    int* cfl_events = nullptr;   // [B] extra RK3 steps the CFL guard inserted so far (cluster kernels), or nullptr
};

// everything one CTA needs while it owns an environment
template <typename Real>
struct Ctx {
    Real* s0;        // shared: state buffer 0 (NS_SM, rows padded to SX)
    Real* s1;        // second state buffer (NS_SM): shared (fp32) or global (fp64)
    Real* R;         // shared: Poisson scratch / pHY' (NR)
    Real* Tb;        // shared: bottom wall temperature per column (NX)
    Real* mid;       // shared: 2*NX values exchanged between the two Thomas sweeps
    Real* tw48;      // shared twiddles
    Real* tw96;
    Real* gm;        // global: two slabs (2 x NSTATE) of stage tendencies of this CTA, used ping-pong
    const Real* tinv; // Thomas pivots [NZ][NX]: shared copy (fp32) or the global table (fp64)
    double* red;     // shared: NRED*NT doubles for the epilogue reductions (fp64 mode; fp32 aliases s0/s1)
    Real* E;         // shared: NE left fluxes of the warps' first columns (phase_edge_fluxes); fp64 mode: inside `red`
};

// what one kernel launch does with each environment it visits
struct RunFlags {
    int nsub;            // RK3 steps to take (Consts::nsub for a step, 0 for observe-only)
    int project_first;   // 1: project the loaded velocity first (Oceananigans set!, dtau = 1)
    int advance_clock;   // 1: t += dt_action, step += 1, truncation flag (step_simulation)
};

template <bool V>
struct BoolTag {
    static constexpr bool value = V;
};

RBC_HD int wrapx(int i) { return i < 0 ? i + NX : (i >= NX ? i - NX : i); }

// ------------------------------------------------------------------------------------------
// heater profile (collate_actions_colin, rbc_sim2D.jl:87-133), fp64 like the reference
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD double heater_T(const Consts<Real>& C, const float* action, double x)
{
    const int na = C.heaters;
    const double ampl = C.heater_limit, dx = 0.03;
    double mean = 0.0;
    for (int s = 0; s < na; ++s) mean += ampl * (double)action[s];
    mean /= na;
    double dev = 0.0;
    for (int s = 0; s < na; ++s) {
        double d = fabs(ampl * (double)action[s] - mean);
        if (d > dev) dev = d;
    }
    double K2 = dev / ampl;
    if (!(K2 > 1.0)) K2 = 1.0;
    const double seg = C.lx / na;
    int s = (int)floor(x / seg);
    if (s >= na) s = na - 1;
    const int sm = (s == 0) ? na - 1 : s - 1, sp = (s == na - 1) ? 0 : s + 1;
    const double T0 = 2 + (ampl * (double)action[sm] - mean) / K2;
    const double T1 = 2 + (ampl * (double)action[s] - mean) / K2;
    const double T2 = 2 + (ampl * (double)action[sp] - mean) / K2;
    const double xp = x - s * seg;
    if (xp < dx) return T0 + ((T0 - T1) / (4 * dx * dx * dx)) * (xp - 2 * dx) * (xp + dx) * (xp + dx);
    if (xp >= seg - dx)
        return T1 + ((T1 - T2) / (4 * dx * dx * dx)) * (xp - seg - 2 * dx) * (xp - seg + dx) * (xp - seg + dx);
    return T1;
}

// ------------------------------------------------------------------------------------------
// reconstruction on a 6-value window win[0..5] = c[j-3 .. j+2] around the face j-1|j
// (SURVEY 8a "Advection"); the right-biased stencil is the mirror image of the left one.
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD Real upwind5(Real vel, const Real* win)
{
    const bool pos = vel > Real(0);
    const Real x0 = pos ? win[0] : win[5], x1 = pos ? win[1] : win[4], x2 = pos ? win[2] : win[3];
    const Real x3 = pos ? win[3] : win[2], x4 = pos ? win[4] : win[1];
    const Real phi = Real(2.0 / 60.0) * x0 - Real(13.0 / 60.0) * x1 + Real(47.0 / 60.0) * x2 + Real(27.0 / 60.0) * x3 -
                     Real(3.0 / 60.0) * x4;
    return vel * phi;
}
template <typename Real>
RBC_HD Real upwind3(Real vel, const Real* win)
{
    const bool pos = vel > Real(0);
    const Real x1 = pos ? win[1] : win[4], x2 = pos ? win[2] : win[3], x3 = pos ? win[3] : win[2];
    return vel * (Real(5.0 / 6.0) * x2 + Real(2.0 / 6.0) * x3 - Real(1.0 / 6.0) * x1);
}
template <typename Real>
RBC_HD Real upwind1(Real vel, const Real* win)
{
    return vel * (vel > Real(0) ? win[2] : win[3]);
}
// order: 5, 3 or 1 (uniform across a warp: it depends on the row only)
template <typename Real>
RBC_HD Real upwind_ord(Real vel, const Real* win, int ord)
{
    if (ord == 5) return upwind5(vel, win);
    if (ord == 3) return upwind3(vel, win);
    return upwind1(vel, win);
}
template <typename Real>
RBC_HD Real centred4(Real a, Real b, Real c, Real d)   // values at j-2, j-1, j, j+1
{
    return Real(7.0 / 12.0) * (b + c) - Real(1.0 / 12.0) * (a + d);
}
template <typename Real>
RBC_HD Real centred_ord(Real a, Real b, Real c, Real d, int ord)
{
    return ord == 4 ? centred4(a, b, c, d) : (b + c) * Real(0.5);
}
// ------------------------------------------------------------------------------------------
// Two reconstructions at once.  The march is issue-bound, and on sm_100a two fp32 operations on a register pair
// issue as ONE instruction (FFMA2 / FMUL2 / FADD2, same FMA-pipe throughput, half the issue slots).  Fluxes come in
// natural pairs with identical coefficients (the two x-faces of a cell, two of the three z-faces), so the fp32
// device path evaluates them packed: the selects write straight into adjacent registers and the coefficients are
// broadcast immediates, i.e. no packing moves (checked in SASS).  Same operations in the same order as the scalar
// functions; every other instantiation (fp64, host emulator) simply calls the scalar functions twice.
// ------------------------------------------------------------------------------------------
template <typename Real>
struct Pair {
    Real a, b;
};
template <typename Real>
RBC_HD Pair<Real> upwind5_pair(Real v0, const Real* w0, Real v1, const Real* w1)
{
    return {upwind5(v0, w0), upwind5(v1, w1)};
}
template <typename Real>
RBC_HD Pair<Real> centred4_pair(Real a0, Real b0, Real c0, Real d0, Real a1, Real b1, Real c1, Real d1)
{
    return {centred4(a0, b0, c0, d0), centred4(a1, b1, c1, d1)};
}
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
// RBC_UPWIND_SYM: select-free forms of the upwind choice, kept as measured alternatives (default 0 = selects).  With
// S = (phiL + phiR)/2 (the 6-point centred reconstruction) and D = (phiL - phiR)/2 the upwinded flux  v > 0 ? v phiL : v phiR
// equals  v S + |v| D:   60 S = (w0 + w5) - 8 (w1 + w4) + 37 (w2 + w3),   60 D = (w0 - w5) - 5 (w1 - w4) + 10 (w2 - w3).
// That trades the 12 selects/compares of a pair (half-rate ALU pipe) for FMA-pipe work.  Measured on B200 with
// tools/microbench/tendency_tile.cu (cycles per tendency stage): selects 11.36 k; = 1 (all packed: the operand pairs need
// ~900 packing moves) 12.56 k; = 2 (scalar sums/differences, packed FMAs, no moves) 12.02 k.  The march is not ALU-bound.
#ifndef RBC_UPWIND_SYM
#define RBC_UPWIND_SYM 0
#endif
template <>
__device__ __forceinline__ Pair<float> upwind5_pair<float>(float v0, const float* w0, float v1, const float* w1)
{
#if RBC_UPWIND_SYM
    const float2 a0 = make_float2(w0[0], w1[0]), a1 = make_float2(w0[1], w1[1]), a2 = make_float2(w0[2], w1[2]);
    const float2 a3 = make_float2(w0[3], w1[3]), a4 = make_float2(w0[4], w1[4]), a5 = make_float2(w0[5], w1[5]);
#if RBC_UPWIND_SYM == 2
    // sums and differences as scalar adds: their results land in adjacent registers, so the packed FMAs need no packing moves
    const float2 s0 = make_float2(a0.x + a5.x, a0.y + a5.y), s1 = make_float2(a1.x + a4.x, a1.y + a4.y), s2 = make_float2(a2.x + a3.x, a2.y + a3.y);
    const float2 d0 = make_float2(a0.x - a5.x, a0.y - a5.y), d1 = make_float2(a1.x - a4.x, a1.y - a4.y), d2 = make_float2(a2.x - a3.x, a2.y - a3.y);
#else
    const float2 s0 = __fadd2_rn(a0, a5), s1 = __fadd2_rn(a1, a4), s2 = __fadd2_rn(a2, a3);
    const float2 d0 = __fadd2_rn(a0, make_float2(-a5.x, -a5.y)), d1 = __fadd2_rn(a1, make_float2(-a4.x, -a4.y)), d2 = __fadd2_rn(a2, make_float2(-a3.x, -a3.y));
#endif
    const float c0 = float(1.0 / 60.0), cs1 = float(-8.0 / 60.0), cs2 = float(37.0 / 60.0), cd1 = float(-5.0 / 60.0), cd2 = float(10.0 / 60.0);
    float2 S = __fmul2_rn(make_float2(c0, c0), s0);
    S = __ffma2_rn(make_float2(cs1, cs1), s1, S);
    S = __ffma2_rn(make_float2(cs2, cs2), s2, S);
    float2 D = __fmul2_rn(make_float2(c0, c0), d0);
    D = __ffma2_rn(make_float2(cd1, cd1), d1, D);
    D = __ffma2_rn(make_float2(cd2, cd2), d2, D);
    const float2 f = __ffma2_rn(make_float2(fabsf(v0), fabsf(v1)), D, __fmul2_rn(make_float2(v0, v1), S));
    return {f.x, f.y};
#else
    const bool p0 = v0 > 0.f, p1 = v1 > 0.f;
    const float2 x0 = make_float2(p0 ? w0[0] : w0[5], p1 ? w1[0] : w1[5]);
    const float2 x1 = make_float2(p0 ? w0[1] : w0[4], p1 ? w1[1] : w1[4]);
    const float2 x2 = make_float2(p0 ? w0[2] : w0[3], p1 ? w1[2] : w1[3]);
    const float2 x3 = make_float2(p0 ? w0[3] : w0[2], p1 ? w1[3] : w1[2]);
    const float2 x4 = make_float2(p0 ? w0[4] : w0[1], p1 ? w1[4] : w1[1]);
    const float k0 = float(2.0 / 60.0), k1 = float(-13.0 / 60.0), k2 = float(47.0 / 60.0), k3 = float(27.0 / 60.0), k4 = float(-3.0 / 60.0);
    float2 phi = __fmul2_rn(make_float2(k0, k0), x0);
    phi = __ffma2_rn(make_float2(k1, k1), x1, phi);
    phi = __ffma2_rn(make_float2(k2, k2), x2, phi);
    phi = __ffma2_rn(make_float2(k3, k3), x3, phi);
    phi = __ffma2_rn(make_float2(k4, k4), x4, phi);
    const float2 f = __fmul2_rn(make_float2(v0, v1), phi);
    return {f.x, f.y};
#endif
}
template <>
__device__ __forceinline__ Pair<float> centred4_pair<float>(float a0, float b0, float c0, float d0, float a1, float b1, float c1, float d1)
{
    const float2 in = __fadd2_rn(make_float2(b0, b1), make_float2(c0, c1));
    const float2 out = __fadd2_rn(make_float2(a0, a1), make_float2(d0, d1));
    const float h = float(7.0 / 12.0), q = float(-1.0 / 12.0);
    const float2 r = __ffma2_rn(make_float2(q, q), out, __fmul2_rn(make_float2(h, h), in));
    return {r.x, r.y};
}
#endif

// wall-order rules (0-based), SURVEY 8a: centres->z-face kf, and z-faces->centre kc
RBC_HD int ord_up_face(int kf) { return (kf >= 3 && kf <= NZ - 3) ? 5 : ((kf == 2 || kf == NZ - 2) ? 3 : 1); }
RBC_HD int ord_ce_face(int kf) { return (kf >= 2 && kf <= NZ - 2) ? 4 : 2; }
RBC_HD int ord_up_cen(int kc) { return (kc >= 2 && kc <= NZ - 3) ? 5 : ((kc == 1 || kc == NZ - 2) ? 3 : 1); }
RBC_HD int ord_ce_cen(int kc) { return (kc >= 1 && kc <= NZ - 2) ? 4 : 2; }

// ------------------------------------------------------------------------------------------
// phase: hydrostatic pressure anomaly (split mode only), column integral from the top
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void phase_phy(int tid, const Consts<Real>& C, const Real* cb, Real* phy)
{
    if (tid >= NX) return;
    const Real dz = Real(1) / C.idz;
    Real below = cb[(NZ - 1) * SX + tid];
    Real acc = -Real(0.5) * (below + (Real(2) * C.b_top - below)) * dz;
    phy[(NZ - 1) * RSTR + tid] = acc;
    for (int k = NZ - 2; k >= 0; --k) {
        const Real bk = cb[k * SX + tid];
        acc = acc - Real(0.5) * (bk + below) * dz;
        phy[k * RSTR + tid] = acc;
        below = bk;
    }
}

// ------------------------------------------------------------------------------------------
// The x-direction fluxes of a cell are shared with its neighbours: the flux through the right face of column i is
// the flux through the left face of column i+1.  Each thread therefore evaluates only the three LEFT fluxes of its
// cell (tracer flux at x-face i, u-momentum flux at x-centre i-1, w-momentum flux at (x-face i, z-face k)) and takes
// the right ones from lane+1 with a warp shuffle.  Lane 31 has no lane+1: its right fluxes belong to column
// 32(w+1) mod NX, and those are evaluated for all rows by an otherwise idle-lane prologue (phase_edge_fluxes) into
// a small shared-memory table E[flux][row][warp column].  One third of the upwind reconstructions of the march go away.
// ------------------------------------------------------------------------------------------
constexpr int NEDGE = NX / 32;                       // columns 0, 32, 64
constexpr int NE = 3 * NZ * NEDGE;                   // table size

template <typename Real>
RBC_HD void left_fluxes(const Real* RBC_RESTRICT c, int i, int k, bool want_bu, bool want_w, Real& fb, Real& fu, Real& fw)
{
    const Real* RBC_RESTRICT cb = c + OFF_B + k * SX;
    const Real* RBC_RESTRICT cu = c + OFF_U + k * SX;
    const Real* RBC_RESTRICT cw = c + OFF_W + k * SX;
    int col[6];
    RBC_UNROLL
    for (int j = 0; j < 6; ++j) col[j] = wrapx(i - 3 + j);
    if (want_bu) {
        Real bx[6], ux[6];
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) { bx[j] = cb[col[j]]; ux[j] = cu[col[j]]; }
        fb = upwind5(ux[3], bx);
        fu = upwind5(centred4(ux[1], ux[2], ux[3], ux[4]), ux);
    }
    if (want_w) {
        Real wx[6], uz[4];
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) wx[j] = cw[col[j]];
        RBC_UNROLL
        for (int j = 0; j < 4; ++j) {
            const int kk = k - 2 + j;
            uz[j] = (kk >= 0 && kk < NZ) ? c[OFF_U + kk * SX + i] : Real(0);
        }
        fw = upwind5(centred_ord(uz[0], uz[1], uz[2], uz[3], ord_ce_face(k)), wx);
    }
}

// prologue of the march: left fluxes of the first column of every warp, all rows.  Threads 0..191 do the tracer and u
// fluxes of (row, warp column), threads 192..383 the w flux.
template <typename Real>
RBC_HD void phase_edge_fluxes(int tid, const Real* RBC_RESTRICT c, Real* RBC_RESTRICT E)
{
    const int item = tid % (NZ * NEDGE), half = tid / (NZ * NEDGE);
    if (half > 1) return;
    const int k = item % NZ, ci = item / NZ;
    Real fb = Real(0), fu = Real(0), fw = Real(0);
    left_fluxes(c, ci * 32, k, half == 0, half == 1, fb, fu, fw);
    if (half == 0) {
        E[(0 * NZ + k) * NEDGE + ci] = fb;
        E[(1 * NZ + k) * NEDGE + ci] = fu;
    } else {
        E[(2 * NZ + k) * NEDGE + ci] = fw;
    }
}
static_assert(NT >= 2 * NZ * NEDGE, "the edge-flux prologue needs one thread per (row, warp column, half)");

// ------------------------------------------------------------------------------------------
// phase: tendencies + RK3 substep for one strip.  Reads the current state `c`, writes the
// predicted state `n` (U* = U + dt (gam G + zet G-)), stores G into the CTA's G- slab.
// ------------------------------------------------------------------------------------------
template <typename Real, bool SPLIT>
RBC_HD void phase_tendency(int tid, const Consts<Real>& C, const Real* RBC_RESTRICT c, Real* RBC_RESTRICT n,
                           const Real* RBC_RESTRICT phy, const Real* RBC_RESTRICT Tb, const Real* RBC_RESTRICT E, const Real* gm_in, Real* gm_out,
                           Real dt, Real gam, Real zet, bool use_gm)
{
    const int i = tid % NX, s = tid / NX, k0 = s * RS;
    const Real* RBC_RESTRICT cb = c + OFF_B;
    const Real* RBC_RESTRICT cu = c + OFF_U;
    const Real* RBC_RESTRICT cw = c + OFF_W;
    int col[6];
    RBC_UNROLL
    for (int j = 0; j < 6; ++j) col[j] = wrapx(i - 3 + j);
    const bool last_lane = (i & 31) == 31;
    const int eci = ((i + 1) % NX) / 32;                 // table column of lane 31's right neighbour

    // sliding windows of the own column; index j <-> row k-3+j
    Real bz[7], uz[7], wz[7], wxr[6];
    RBC_UNROLL
    for (int j = 0; j < 7; ++j) {
        const int k = k0 - 3 + j;
        const bool ok = (k >= 0 && k < NZ);
        bz[j] = ok ? cb[k * SX + i] : Real(0);
        uz[j] = ok ? cu[k * SX + i] : Real(0);
        wz[j] = (k >= 0 && k <= NZ) ? cw[k * SX + i] : Real(0);
    }
    RBC_UNROLL
    for (int j = 0; j < 6; ++j) wxr[j] = cw[k0 * SX + col[j]];   // w(i-3..i+2, face k0)

    // fluxes through the strip's lower boundary (carried afterwards)
    Real Fzb_lo = Real(0), Wu_lo = Real(0), Ww_lo = Real(0);
    if (k0 >= 1) {                                       // k0 is 16, 32 or 48: full-order stencils
        Fzb_lo = upwind5(wz[3], bz);
        Wu_lo = upwind5(centred4(wxr[1], wxr[2], wxr[3], wxr[4]), uz);
        Ww_lo = upwind5(centred4(wz[1], wz[2], wz[3], wz[4]), wz);
    }
    const Real tb = Tb[i];
    const Real kdx = C.kappa * C.idx2, kdz = C.kappa * C.idz2, ndx = C.nu * C.idx2, ndz = C.nu * C.idz2;
    const Real dtg = dt * gam, dtz = dt * zet;

    // previous-stage tendencies live in per-CTA global slabs [field][r][tid] (coalesced; one slab is read,
    // the other written, so loads never alias stores).  They are fetched one row ahead so their L2 latency
    // hides behind a whole row of arithmetic (stage 1 has zet = 0 and must not read).
    Real gnb = Real(0), gnu = Real(0), gnw = Real(0), gmb = Real(0), gmu = Real(0), gmw = Real(0);   // rows r and r+1
    if (use_gm) {
        gnb = gm_in[(0 * RS) * NT + tid]; gnu = gm_in[(1 * RS) * NT + tid]; gnw = gm_in[(2 * RS) * NT + tid];
        gmb = gm_in[(0 * RS + 1) * NT + tid]; gmu = gm_in[(1 * RS + 1) * NT + tid]; gmw = gm_in[(2 * RS + 1) * NT + tid];
    }

    // One row of the march.  EDGE = false is the interior fast path (rows 2..60: all stencils at full order,
    // no wall ghosts) with every order decision resolved at compile time; EDGE = true handles rows
    // 0, 1, 61, 62, 63 with the wall-order rules.  The choice is uniform across a warp.
    auto row = [&](auto edge_tag, const int r) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int k = k0 + r;
        const Real gb0 = gnb, gu0 = gnu, gw0 = gnw;
        gnb = gmb; gnu = gmu; gnw = gmw;
        if (r + 2 < RS && use_gm) {                          // two rows ahead
            gmb = gm_in[(0 * RS + r + 2) * NT + tid];
            gmu = gm_in[(1 * RS + r + 2) * NT + tid];
            gmw = gm_in[(2 * RS + r + 2) * NT + tid];
        }
        Real bx[6], ux[6], wxn[6];
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) {
            bx[j] = (j == 3) ? bz[3] : cb[k * SX + col[j]];
            ux[j] = (j == 3) ? uz[3] : cu[k * SX + col[j]];
            wxn[j] = (j == 3) ? wz[4] : cw[(k + 1) * SX + col[j]];      // face k+1 <= NZ always valid
        }
        const bool top = EDGE && (k == NZ - 1);
        const bool bot = EDGE && (k == 0);
        const int o_face_hi = EDGE ? ord_up_face(k + 1) : 5;
        const int o_ce_face = EDGE ? ord_ce_face(k) : 4;
        const int o_up_cen = EDGE ? ord_up_cen(k) : 5;
        const int o_ce_cen = EDGE ? ord_ce_cen(k) : 4;

        // left x-fluxes and upper z-fluxes.  Interior rows evaluate them in pairs (see upwind5_pair); wall rows keep the
        // scalar, order-selecting forms.
        Real Fx0, F0, Fw0, Fzb_hi, Wu_hi, Ww_hi;
        if (!EDGE) {
            const Pair<Real> ce = centred4_pair(ux[1], ux[2], ux[3], ux[4], uz[1], uz[2], uz[3], uz[4]);   // u~ at centre i-1; u at (x-face i, z-face k)
            const Pair<Real> fxu = upwind5_pair(ux[3], bx, ce.a, ux);
            const Pair<Real> fwz = upwind5_pair(ce.b, wxr, wz[4], bz + 1);
            const Pair<Real> wa = centred4_pair(wxn[1], wxn[2], wxn[3], wxn[4], wz[2], wz[3], wz[4], wz[5]);
            const Pair<Real> fz = upwind5_pair(wa.a, uz + 1, wa.b, wz + 1);
            Fx0 = fxu.a; F0 = fxu.b; Fw0 = fwz.a; Fzb_hi = fwz.b; Wu_hi = fz.a; Ww_hi = fz.b;
        } else {
            Fx0 = upwind5(ux[3], bx);
            F0 = upwind5(centred4(ux[1], ux[2], ux[3], ux[4]), ux);          // centre i-1
            Fw0 = upwind5(centred_ord(uz[1], uz[2], uz[3], uz[4], o_ce_face), wxr);   // x-face i, z-face k
            Fzb_hi = top ? Real(0) : upwind_ord(wz[4], bz + 1, o_face_hi);
            Wu_hi = top ? Real(0) : upwind_ord(centred4(wxn[1], wxn[2], wxn[3], wxn[4]), uz + 1, o_face_hi);
            Ww_hi = upwind_ord(centred_ord(wz[2], wz[3], wz[4], wz[5], o_ce_cen), wz + 1, o_up_cen);
        }
        // right x-fluxes = the neighbour's left ones
        Real Fx1, F1, Fw1;
#if defined(__CUDA_ARCH__)
        Fx1 = __shfl_down_sync(0xffffffffu, Fx0, 1);
        F1 = __shfl_down_sync(0xffffffffu, F0, 1);
        Fw1 = __shfl_down_sync(0xffffffffu, Fw0, 1);
        if (last_lane) {
            Fx1 = E[(0 * NZ + k) * NEDGE + eci];
            F1 = E[(1 * NZ + k) * NEDGE + eci];
            Fw1 = E[(2 * NZ + k) * NEDGE + eci];
        }
#else
        if (last_lane) {
            Fx1 = E[(0 * NZ + k) * NEDGE + eci];
            F1 = E[(1 * NZ + k) * NEDGE + eci];
            Fw1 = E[(2 * NZ + k) * NEDGE + eci];
        } else {
            left_fluxes(c, i + 1, k, true, true, Fx1, F1, Fw1);
        }
#endif

        // ---- tracer ----
        const Real bdn = bot ? (Real(2) * tb - bz[3]) : bz[2];
        const Real bup = top ? (Real(2) * C.b_top - bz[3]) : bz[4];
        const Real Gb = (Fx0 - Fx1) * C.idx + (Fzb_lo - Fzb_hi) * C.idz + (bx[4] - Real(2) * bx[3] + bx[2]) * kdx +
                        (bup - Real(2) * bz[3] + bdn) * kdz;

        // ---- u ----
        const Real udn = bot ? -uz[3] : uz[2];
        const Real uup = top ? -uz[3] : uz[4];
        Real Gu = (F0 - F1) * C.idx + (Wu_lo - Wu_hi) * C.idz + (ux[4] - Real(2) * ux[3] + ux[2]) * ndx +
                  (uup - Real(2) * uz[3] + udn) * ndz;
        if (SPLIT) Gu -= (phy[k * RSTR + i] - phy[k * RSTR + col[2]]) * C.idx;

        // ---- w (face k; face 0 is the wall) ----
        Real Gw = (Fw0 - Fw1) * C.idx + (Ww_lo - Ww_hi) * C.idz + (wxr[4] - Real(2) * wxr[3] + wxr[2]) * ndx +
                  (wz[4] - Real(2) * wz[3] + wz[2]) * ndz;
        if (!SPLIT) Gw += Real(0.5) * (bz[2] + bz[3]);
        if (bot) Gw = Real(0);

        // ---- RK3 substep ----
        gm_out[(0 * RS + r) * NT + tid] = Gb;
        gm_out[(1 * RS + r) * NT + tid] = Gu;
        gm_out[(2 * RS + r) * NT + tid] = Gw;
        n[OFF_B + k * SX + i] = bz[3] + dtg * Gb + dtz * gb0;
        n[OFF_U + k * SX + i] = uz[3] + dtg * Gu + dtz * gu0;
        n[OFF_W + k * SX + i] = bot ? Real(0) : wz[3] + dtg * Gw + dtz * gw0;

        // ---- slide ----
        Fzb_lo = Fzb_hi; Wu_lo = Wu_hi; Ww_lo = Ww_hi;
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) { bz[j] = bz[j + 1]; uz[j] = uz[j + 1]; wz[j] = wz[j + 1]; }
        const int kn = k + 4;
        bz[6] = (kn < NZ) ? cb[kn * SX + i] : Real(0);
        uz[6] = (kn < NZ) ? cu[kn * SX + i] : Real(0);
        wz[6] = (kn <= NZ) ? cw[kn * SX + i] : Real(0);
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) wxr[j] = wxn[j];
    };

    RBC_UNROLL
    for (int r = 0; r < RS; ++r) {
        if (r >= 2 && r < RS - 3) {
            row(BoolTag<false>{}, r);                         // rows that are interior in every strip
        } else {
            const bool edge = (r < 2) ? (s == 0) : (s == NSTRIP - 1);
            if (edge) row(BoolTag<true>{}, r);
            else row(BoolTag<false>{}, r);
        }
    }
    if (s == NSTRIP - 1) n[OFF_W + NZ * SX + i] = Real(0);       // top wall face
}

// ------------------------------------------------------------------------------------------
// phase: copy a state slab (used when the predicted state lives in global memory, fp64 mode)
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void phase_copy(int tid, const Real* src, Real* dst, int nvals)
{
    for (int q = tid; q < nvals; q += NT) dst[q] = src[q];
}

// ------------------------------------------------------------------------------------------
// small complex DFTs (SIGN = -1 forward, +1 inverse), natural in/out order
// ------------------------------------------------------------------------------------------
// aligned to its own size: every complex number in shared memory sits on a 2*sizeof(Real) boundary, and saying so lets
// the compiler emit one 64-/128-bit access even where it cannot prove the alignment of a run-time offset (without it
// the cluster kernel's FFT passes issued two 32-bit accesses with a 2-way bank conflict each)
template <typename Real>
struct alignas(2 * sizeof(Real)) cx {
    Real re, im;
};
template <typename Real> RBC_HD cx<Real> cadd(cx<Real> a, cx<Real> b) { return {a.re + b.re, a.im + b.im}; }
template <typename Real> RBC_HD cx<Real> csub(cx<Real> a, cx<Real> b) { return {a.re - b.re, a.im - b.im}; }
template <typename Real> RBC_HD cx<Real> cmul(cx<Real> a, Real c, Real s) { return {a.re * c - a.im * s, a.re * s + a.im * c}; }
template <int SIGN, typename Real> RBC_HD cx<Real> muli(cx<Real> a)   // multiply by SIGN*i
{
    return SIGN > 0 ? cx<Real>{-a.im, a.re} : cx<Real>{a.im, -a.re};
}
// 64-/128-bit shared-memory access of one complex number (address is always 2*sizeof(Real) aligned)
template <typename Real> RBC_HD cx<Real> ldc(const Real* p) { return *reinterpret_cast<const cx<Real>*>(p); }
template <typename Real> RBC_HD void stc(Real* p, cx<Real> v) { *reinterpret_cast<cx<Real>*>(p) = v; }

template <int SIGN, typename Real>
RBC_HD void dft4(cx<Real>& x0, cx<Real>& x1, cx<Real>& x2, cx<Real>& x3)
{
    const cx<Real> t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = muli<SIGN>(csub(x1, x3));
    x0 = cadd(t0, t2); x2 = csub(t0, t2);
    x1 = cadd(t1, t3); x3 = csub(t1, t3);
}
template <int SIGN, typename Real>
RBC_HD void dft8(cx<Real>* x)
{
    cx<Real> e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6], o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
    dft4<SIGN>(e0, e1, e2, e3);
    dft4<SIGN>(o0, o1, o2, o3);
    const Real h = Real(0.70710678118654752440);
    // W8^k O[k]: k=1: h(1 + SIGN i) ; k=2: SIGN i ; k=3: h(-1 + SIGN i)
    const cx<Real> w1 = cmul(o1, h, Real(SIGN) * h);
    const cx<Real> w2 = muli<SIGN>(o2);
    const cx<Real> w3 = cmul(o3, -h, Real(SIGN) * h);
    x[0] = cadd(e0, o0); x[4] = csub(e0, o0);
    x[1] = cadd(e1, w1); x[5] = csub(e1, w1);
    x[2] = cadd(e2, w2); x[6] = csub(e2, w2);
    x[3] = cadd(e3, w3); x[7] = csub(e3, w3);
}
template <int SIGN, typename Real>
RBC_HD void dft3(cx<Real>& x0, cx<Real>& x1, cx<Real>& x2)
{
    const Real q = Real(0.86602540378443864676);
    const cx<Real> s = cadd(x1, x2), d = muli<SIGN>(csub(x1, x2));
    const cx<Real> m = {x0.re - Real(0.5) * s.re, x0.im - Real(0.5) * s.im};
    x0 = cadd(x0, s);
    x1 = {m.re + q * d.re, m.im + q * d.im};
    x2 = {m.re - q * d.re, m.im - q * d.im};
}
template <int SIGN, typename Real>
RBC_HD void dft6(cx<Real>* x)
{
    cx<Real> e0 = x[0], e1 = x[2], e2 = x[4], o0 = x[1], o1 = x[3], o2 = x[5];
    dft3<SIGN>(e0, e1, e2);
    dft3<SIGN>(o0, o1, o2);
    const Real q = Real(0.86602540378443864676);
    const cx<Real> w1 = cmul(o1, Real(0.5), Real(SIGN) * q);
    const cx<Real> w2 = cmul(o2, Real(-0.5), Real(SIGN) * q);
    x[0] = cadd(e0, o0); x[3] = csub(e0, o0);
    x[1] = cadd(e1, w1); x[4] = csub(e1, w1);
    x[2] = cadd(e2, w2); x[5] = csub(e2, w2);
}

// ------------------------------------------------------------------------------------------
// The x-transform.  A row of the scratch R holds 48 complex numbers z[n] = (r[2n], r[2n+1]); the
// 96-point real FFT is a 48-point complex FFT (6 x 8 Cooley-Tukey, in place) plus the real-FFT
// split.  Lanes map to consecutive ROWS (item = group*64 + row) so that every 64-bit access of a
// warp walks the padded rows conflict-free.
//   pass A fwd : divergence of the predicted velocity (computed on the fly from u*, w*) ->
//                8-point DFTs over n2 (n = n1 + 6 n2) -> twiddle W48^(n1 k2) -> slot n1 + 6 k2
//   pass B fwd : 6-point DFTs over n1 for the column pair (k2, 8-k2) [or (0,4)] -> Z[8 k1 + k2],
//                then the real-FFT split in registers -> X[m] in slot (m>>3) + 6 (m&7)
//   pass B inv / pass A inv : the exact inverses (conjugate twiddles), leaving phi in natural order.
// ------------------------------------------------------------------------------------------
RBC_HD int slot48(int k) { return (k >> 3) + 6 * (k & 7); }

template <typename Real>
RBC_HD void fft_passA_fwd_div(int item, const Consts<Real>& C, const Real* RBC_RESTRICT p, Real* RBC_RESTRICT R,
                              const Real* RBC_RESTRICT tw48)
{
    const int n1 = item >> 6, row = item & 63;
    const Real* pu = p + OFF_U + row * SX;
    const Real* pw0 = p + OFF_W + row * SX;
    const Real* pw1 = pw0 + SX;
    cx<Real> a[8];
    RBC_UNROLL
    for (int n2 = 0; n2 < 8; ++n2) {
        const int x = 2 * (n1 + 6 * n2);
        const cx<Real> u01 = ldc(pu + x), w0 = ldc(pw0 + x), w1 = ldc(pw1 + x);
        const Real u2 = pu[(x + 2 == NX) ? 0 : x + 2];
        a[n2].re = (u01.im - u01.re) * C.idx + (w1.re - w0.re) * C.idz;
        a[n2].im = (u2 - u01.im) * C.idx + (w1.im - w0.im) * C.idz;
    }
    dft8<-1>(a);
    Real* z = R + row * RSTR;
    RBC_UNROLL
    for (int k2 = 0; k2 < 8; ++k2) {
        const int j = n1 * k2;                       // W48^(n1 k2), forward: e^{-i theta}
        stc(z + 2 * (n1 + 6 * k2), cmul(a[k2], tw48[2 * j], -tw48[2 * j + 1]));
    }
}
template <typename Real>
RBC_HD void fft_passA_inv(int item, Real* R)
{
    const int n1 = item >> 6, row = item & 63;
    Real* z = R + row * RSTR;
    cx<Real> a[8];
    RBC_UNROLL
    for (int k2 = 0; k2 < 8; ++k2) a[k2] = ldc(z + 2 * (n1 + 6 * k2));
    dft8<+1>(a);
    RBC_UNROLL
    for (int n2 = 0; n2 < 8; ++n2) stc(z + 2 * (n1 + 6 * n2), a[n2]);
}

// real-FFT split of the pair (m, 48-m): Z (48-point transform of the packed row) <-> X (96-point spectrum)
//   E = (Z[m] + conj Z[48-m])/2, O = -i (Z[m] - conj Z[48-m])/2, X[m] = E + W96^m O, X[48-m] = conj(E - W96^m O)
template <typename Real>
RBC_HD void untangle_pair(cx<Real>& zm, cx<Real>& zp, Real c, Real s)
{
    const cx<Real> E = {Real(0.5) * (zm.re + zp.re), Real(0.5) * (zm.im - zp.im)};
    const cx<Real> D = {Real(0.5) * (zm.re - zp.re), Real(0.5) * (zm.im + zp.im)};
    const cx<Real> WO = cmul(cx<Real>{D.im, -D.re}, c, -s);
    zm = {E.re + WO.re, E.im + WO.im};
    zp = {E.re - WO.re, -(E.im - WO.im)};
}
template <typename Real>
RBC_HD void tangle_pair(cx<Real>& xm, cx<Real>& xp, Real c, Real s)
{
    const cx<Real> E = {Real(0.5) * (xm.re + xp.re), Real(0.5) * (xm.im - xp.im)};
    const cx<Real> WO = {Real(0.5) * (xm.re - xp.re), Real(0.5) * (xm.im + xp.im)};
    const cx<Real> O = cmul(WO, c, s);               // conj(W96^m) * (W96^m O)
    xm = {E.re - O.im, E.im + O.re};                 // Z[m]    = E + i O
    xp = {E.re + O.im, -E.im + O.re};                // Z[48-m] = conj(E) + i conj(O)
}

// item = g*64 + row, g = 0..3: column pair (ka, kb) = (0,4) for g = 0, else (g, 8-g)
template <typename Real>
RBC_HD void fft_passB_fwd_untangle(int item, Real* RBC_RESTRICT R, const Real* RBC_RESTRICT tw96)
{
    const int g = item >> 6, row = item & 63;
    const int ka = g, kb = (g == 0) ? 4 : 8 - g;
    Real* za = R + row * RSTR + 12 * ka;
    Real* zb = R + row * RSTR + 12 * kb;
    cx<Real> a[6], b[6];
    RBC_UNROLL
    for (int n1 = 0; n1 < 6; ++n1) { a[n1] = ldc(za + 2 * n1); b[n1] = ldc(zb + 2 * n1); }
    dft6<-1>(a);                                     // a[k1] = Z[8 k1 + ka]
    dft6<-1>(b);                                     // b[k1] = Z[8 k1 + kb]
    if (g == 0) {
        // column 0: m = 8 k1: (0 | 48) real pair, 8<->40, 16<->32, 24 self-conjugate
        const Real x0 = a[0].re + a[0].im, x48 = a[0].re - a[0].im;
        a[0] = {x0, x48};
        untangle_pair(a[1], a[5], tw96[2 * 8], tw96[2 * 8 + 1]);
        untangle_pair(a[2], a[4], tw96[2 * 16], tw96[2 * 16 + 1]);
        a[3].im = -a[3].im;
        // column 4: m = 8 k1 + 4 <-> 8 (5-k1) + 4
        untangle_pair(b[0], b[5], tw96[2 * 4], tw96[2 * 4 + 1]);
        untangle_pair(b[1], b[4], tw96[2 * 12], tw96[2 * 12 + 1]);
        untangle_pair(b[2], b[3], tw96[2 * 20], tw96[2 * 20 + 1]);
    } else {
        // m = 8 k1 + g  <->  48 - m = 8 (5-k1) + (8-g)
        RBC_UNROLL
        for (int k1 = 0; k1 < 6; ++k1) {
            const int m = 8 * k1 + g;
            untangle_pair(a[k1], b[5 - k1], tw96[2 * m], tw96[2 * m + 1]);
        }
    }
    RBC_UNROLL
    for (int k1 = 0; k1 < 6; ++k1) { stc(za + 2 * k1, a[k1]); stc(zb + 2 * k1, b[k1]); }
}
template <typename Real>
RBC_HD void fft_passB_inv_tangle(int item, Real* RBC_RESTRICT R, const Real* RBC_RESTRICT tw48, const Real* RBC_RESTRICT tw96)
{
    const int g = item >> 6, row = item & 63;
    const int ka = g, kb = (g == 0) ? 4 : 8 - g;
    Real* za = R + row * RSTR + 12 * ka;
    Real* zb = R + row * RSTR + 12 * kb;
    cx<Real> a[6], b[6];
    RBC_UNROLL
    for (int k1 = 0; k1 < 6; ++k1) { a[k1] = ldc(za + 2 * k1); b[k1] = ldc(zb + 2 * k1); }
    if (g == 0) {
        const Real x0 = a[0].re, x48 = a[0].im;
        a[0] = {Real(0.5) * (x0 + x48), Real(0.5) * (x0 - x48)};      // Z[0] = E0 + i O0
        tangle_pair(a[1], a[5], tw96[2 * 8], tw96[2 * 8 + 1]);
        tangle_pair(a[2], a[4], tw96[2 * 16], tw96[2 * 16 + 1]);
        a[3].im = -a[3].im;
        tangle_pair(b[0], b[5], tw96[2 * 4], tw96[2 * 4 + 1]);
        tangle_pair(b[1], b[4], tw96[2 * 12], tw96[2 * 12 + 1]);
        tangle_pair(b[2], b[3], tw96[2 * 20], tw96[2 * 20 + 1]);
    } else {
        RBC_UNROLL
        for (int k1 = 0; k1 < 6; ++k1) {
            const int m = 8 * k1 + g;
            tangle_pair(a[k1], b[5 - k1], tw96[2 * m], tw96[2 * m + 1]);
        }
    }
    dft6<+1>(a);
    dft6<+1>(b);
    RBC_UNROLL
    for (int n1 = 0; n1 < 6; ++n1) {                  // conjugate twiddles W48^-(n1 k2)
        const int ja = n1 * ka, jb = n1 * kb;
        stc(za + 2 * n1, cmul(a[n1], tw48[2 * ja], tw48[2 * ja + 1]));
        stc(zb + 2 * n1, cmul(b[n1], tw48[2 * jb], tw48[2 * jb + 1]));
    }
}

// ------------------------------------------------------------------------------------------
// phase: tridiagonal solves in z, in place, one pair of threads per spectral word (96 real systems)
//   p(k-1) - (2 + lam dz^2) p(k) + p(k+1) = dz^2 r(k),  p(-1)=p(0), p(NZ)=p(NZ-1)
// Two-sided elimination ("burn at both ends"): thread t sweeps rows 0..31 upward, thread NX+t sweeps
// rows 63..32 downward, the two meet in a 2x2 system, then both back-substitute outward — half the
// serial FMA chain of a one-sided Thomas solve, on twice the threads.
//   tinv[k][t], k < 32 : 1/(diag_k - tinv[k-1][t])   (pivots of the upward sweep)
//   tinv[k][t], k >= 32: 1/(diag_k - tinv[k+1][t])   (pivots of the downward sweep)
// `scale` = dz^2/48 folds the inverse-FFT normalisation.  Rows are processed in blocks of 8 with all
// loads issued up front so only the 4-cycle FMA chain is serial.
// ------------------------------------------------------------------------------------------
constexpr int NZH = NZ / 2;
template <typename Real>
RBC_HD void phase_thomas_sweep(int tid, Real* RBC_RESTRICT R, const Real* RBC_RESTRICT tinv, Real* RBC_RESTRICT mid, Real scale)
{
    if (tid >= 2 * NX) return;
    const int t = tid % NX;
    const bool hi = tid >= NX;
    constexpr int BK = 8;
    // row k = k_first + step*n, n = 0..31: pointer + signed stride instead of per-element selects
    const int sgn = hi ? -1 : 1;
    Real* RBC_RESTRICT pr = R + (hi ? (NZ - 1) * RSTR : 0) + t;
    const Real* RBC_RESTRICT pt = tinv + (hi ? (NZ - 1) * NX : 0) + t;
    const int sr = sgn * RSTR, st = sgn * NX;
    Real d = Real(0);
    for (int kb = 0; kb < NZH; kb += BK) {
        Real iv[BK], rs[BK];
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) { iv[j] = pt[(kb + j) * st]; rs[j] = pr[(kb + j) * sr]; }
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) rs[j] = rs[j] * scale * iv[j];
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) { d = rs[j] - d * iv[j]; pr[(kb + j) * sr] = d; }
    }
    mid[tid] = d;                                     // d'(31) from the upward sweep, e'(32) from the downward one
}
template <typename Real>
RBC_HD void phase_thomas_back(int tid, Real* RBC_RESTRICT R, const Real* RBC_RESTRICT tinv, const Real* RBC_RESTRICT mid)
{
    if (tid >= 2 * NX) return;
    const int t = tid % NX;
    const bool hi = tid >= NX;
    constexpr int BK = 8;
    // p(31) = d - iv31 p(32),  p(32) = e - jv32 p(31)
    const Real dlo = mid[t], ehi = mid[NX + t], iv31 = tinv[(NZH - 1) * NX + t], jv32 = tinv[NZH * NX + t];
    const Real p31 = (dlo - iv31 * ehi) / (Real(1) - iv31 * jv32);
    const Real p32 = ehi - jv32 * p31;
    Real pv = hi ? p32 : p31;
    // row k = k_mid + step*o, o = distance from the middle row of this half (outward)
    const int sgn = hi ? 1 : -1;
    Real* RBC_RESTRICT pr = R + (hi ? NZH : NZH - 1) * RSTR + t;
    const Real* RBC_RESTRICT pt = tinv + (hi ? NZH : NZH - 1) * NX + t;
    const int sr = sgn * RSTR, st = sgn * NX;
    pr[0] = pv;
    for (int kb = 1; kb < NZH; kb += BK) {            // rows 30..0 (lo) / 33..63 (hi)
        Real iv[BK], dd[BK];
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) {
            const bool ok = kb + j < NZH;
            iv[j] = ok ? pt[(kb + j) * st] : Real(0);
            dd[j] = ok ? pr[(kb + j) * sr] : Real(0);
        }
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) {
            if (kb + j < NZH) {
                pv = dd[j] - iv[j] * pv;
                pr[(kb + j) * sr] = pv;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// phase: pressure correction  u -= d_x phi, w -= d_z phi  (phi = dtau * pNHS in R); strip march
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void phase_correct(int tid, const Consts<Real>& C, Real* RBC_RESTRICT p, const Real* RBC_RESTRICT R)
{
    const int i = tid % NX, s = tid / NX, im = wrapx(i - 1), k0 = s * RS;
    Real* RBC_RESTRICT pu = p + OFF_U;
    Real* RBC_RESTRICT pw = p + OFF_W;
    Real ph[RS + 1], pm[RS], uu[RS], ww[RS];
    ph[0] = (k0 >= 1) ? R[(k0 - 1) * RSTR + i] : Real(0);
    RBC_UNROLL
    for (int r = 0; r < RS; ++r) {                      // all loads first: nothing serialises behind a store
        const int k = k0 + r;
        ph[r + 1] = R[k * RSTR + i];
        pm[r] = R[k * RSTR + im];
        uu[r] = pu[k * SX + i];
        ww[r] = pw[k * SX + i];
    }
    RBC_UNROLL
    for (int r = 0; r < RS; ++r) {
        const int k = k0 + r;
        pu[k * SX + i] = uu[r] - (ph[r + 1] - pm[r]) * C.idx;
        if (k >= 1) pw[k * SX + i] = ww[r] - (ph[r + 1] - ph[r]) * C.idz;
    }
}

// ------------------------------------------------------------------------------------------
// the Poisson projection of the state `p` (FFT-x, tridiagonal-z, SURVEY 8a "Poisson solver")
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void project(const Consts<Real>& C, const Ctx<Real>& X, Real* p, Real scale)
{
    RBC_PHASE(fft_passA_fwd_div(tid, C, p, X.R, X.tw48);)
    RBC_PHASE(if (tid < 4 * NZ) fft_passB_fwd_untangle(tid, X.R, X.tw96);)
    RBC_PHASE(phase_thomas_sweep(tid, X.R, X.tinv, X.mid, scale);)
    RBC_PHASE(phase_thomas_back(tid, X.R, X.tinv, X.mid);)
    RBC_PHASE(if (tid < 4 * NZ) fft_passB_inv_tangle(tid, X.R, X.tw48, X.tw96);)
    RBC_PHASE(fft_passA_inv(tid, X.R);)
    RBC_PHASE(phase_correct(tid, C, p, X.R);)
}

// ------------------------------------------------------------------------------------------
// epilogue pieces
// ------------------------------------------------------------------------------------------
// per-thread partial sums for Nusselt numbers and the NaN flag (get_nusselt, rbc_sim2D_api.jl:142-163)
//   red[0] sum b*w (state)           red[1] sum b*w on the sensor grid
//   red[2],red[3] sum_x b on rows (0,1)      red[4],red[5] same for rows (NZ-2, NZ-1)
//   red[6] NaN count;  sensor-row sums: phase_reduce_obs_partials
template <typename Real>
RBC_HD void phase_reduce_partials(int tid, const Consts<Real>& C, const Real* p, double* red)
{
    const int i = tid % NX, s = tid / NX;
    const int oz = NZ / C.obs_nz, ox = NX / C.obs_nx;
    const bool xs = (i % ox) == 0;
    double q1 = 0, q1o = 0, ta = 0, tb = 0, bad = 0;
    for (int r = 0; r < RS; ++r) {
        const int k = s * RS + r;
        const double b = (double)p[OFF_B + k * SX + i], u = (double)p[OFF_U + k * SX + i], w = (double)p[OFF_W + k * SX + i];
        if (b != b || u != u || w != w) bad += 1;
        q1 += b * w;
        const bool zs = (k % oz) == 0;
        if (zs && xs) q1o += b * w;
        if (k == 0 || k == NZ - 2) ta += b;
        if (k == 1 || k == NZ - 1) tb += b;
    }
    const bool lower = (s * RS < NZ / 2);
    red[0 * NT + tid] = q1;
    red[1 * NT + tid] = q1o;
    red[2 * NT + tid] = lower ? ta : 0;   red[3 * NT + tid] = lower ? tb : 0;
    red[4 * NT + tid] = lower ? 0 : ta;   red[5 * NT + tid] = lower ? 0 : tb;
    red[6 * NT + tid] = bad;
}
template <typename Real>
RBC_HD void phase_reduce_obs_partials(int tid, const Consts<Real>& C, const Real* p, double* red)
{
    // sums over sensor columns of b on sensor rows 0, 1, nzo-2, nzo-1 -> red[0..3][tid]
    const int i = tid % NX, s = tid / NX;
    const int oz = NZ / C.obs_nz, ox = NX / C.obs_nx;
    double v[4] = {0, 0, 0, 0};
    if ((i % ox) == 0) {
        const int rows[4] = {0, oz, (C.obs_nz - 2) * oz, (C.obs_nz - 1) * oz};
        for (int q = 0; q < 4; ++q)
            if (rows[q] / RS == s) v[q] += (double)p[OFF_B + rows[q] * SX + i];
    }
    for (int q = 0; q < 4; ++q) red[q * NT + tid] = v[q];
}
RBC_HD double sum_serial(const double* a, int n)
{
    double s = 0;
    for (int q = 0; q < n; ++q) s += a[q];
    return s;
}
// global (compact rows of NX) <-> on-chip (rows of SX) state copies
template <typename Real>
RBC_HD void phase_load_state(int tid, const Real* g, Real* sm)
{
    for (int q = tid; q < NSTATE; q += NT) sm[(q / NX) * SX + (q % NX)] = g[q];
}
// checkpoint bank (fp64, compact) -> on-chip state: the gather + cast of a reset (initialize_from_checkpoint, rbc_sim2D.jl:173-186)
template <typename Real>
RBC_HD void phase_load_bank(int tid, const double* g, Real* sm)
{
    for (int q = tid; q < NSTATE; q += NT) sm[(q / NX) * SX + (q % NX)] = (Real)g[q];
}
template <typename Real>
RBC_HD void phase_store_state(int tid, const Real* sm, Real* g)
{
    for (int q = tid; q < NSTATE; q += NT) g[q] = sm[(q / NX) * SX + (q % NX)];
}

// ------------------------------------------------------------------------------------------
// RBCRewardShaping.compute_cell_distances (rbc_reward_shaping.py:86-140) on the mid-height row of w:
// scipy.signal.find_peaks(uy, height=1e-3) (strict local maxima, plateaus -> midpoint, end points never
// peaks), then for every pair of peaks the periodic distance on linspace(0, 2 pi, NX, endpoint=False), set to
// 0 when uy > 0 everywhere between the two peaks; result = max over pairs (0 with fewer than 2 peaks).
// Serial (one thread): ~NX steps, negligible next to an action step.
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD double cell_distance(const Real* uy)
{
    const double PI = 3.14159265358979323846;
    int peaks[NX / 2 + 1];
    int np = 0;
    int i = 1;
    const int imax = NX - 1;
    while (i < imax) {
        if (uy[i - 1] < uy[i]) {
            int ahead = i + 1;
            while (ahead < imax && uy[ahead] == uy[i]) ++ahead;
            if (uy[ahead] < uy[i]) {
                const int mid = (i + ahead - 1) / 2;
                if ((double)uy[mid] >= 0.001) peaks[np++] = mid;
                i = ahead;
            }
        }
        ++i;
    }
    if (np <= 1) return 0.0;
    double best = 0.0;
    for (int a = 0; a < np; ++a)
        for (int b = a + 1; b < np; ++b) {
            const double xa = peaks[a] * (2 * PI / NX), xb = peaks[b] * (2 * PI / NX);
            const double d1 = fabs(xb - xa), d2 = 2 * PI - d1;
            double d = d1 < d2 ? d1 : d2;
            bool pos = true;
            if (d1 < d2) {
                for (int q = peaks[a]; q < peaks[b]; ++q) pos = pos && (uy[q] > Real(0));
            } else {
                for (int q = peaks[b]; q < NX; ++q) pos = pos && (uy[q] > Real(0));
                for (int q = 0; q < peaks[a]; ++q) pos = pos && (uy[q] > Real(0));
            }
            if (pos) d = 0.0;
            if (d > best) best = d;
        }
    return best;
}

// ------------------------------------------------------------------------------------------
// epilogue of an action step: NaN check, observation, Nusselt numbers, reward, clock / truncation and the fused
// vector-env bookkeeping.  Runs once per environment visit, and a second time (second_pass) when the environment is
// re-initialised from the checkpoint bank inside the same launch.  Returns whether such a reset is due.
// ------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define RBC_COUNT_ONE(p) atomicAdd((p), 1)
#else
#define RBC_COUNT_ONE(p) (*(p) += 1)
#endif
template <typename Real, bool SPLIT, bool NXT_GLOBAL>
RBC_HD bool env_epilogue(const Consts<Real>& C, const EnvIO<Real>& io, const Ctx<Real>& X, int env, const Real* cur, double* red,
                         const RunFlags& F, bool state_changed, Real last_dtau, bool second_pass)
{
    const VecIO& V = io.vec;
    // re-read here rather than carried in registers across the march; nobody has written them since the launch started
    // (thread 0 updates them in the last phase of this function, behind several barriers)
    const int pend = (!second_pass && V.mode == 1 && V.bank != nullptr && F.advance_clock) ? V.pending[env] : 0;
    const double t_old = io.t[env];
    state_changed = state_changed || pend;
    Real* st = io.state + (size_t)env * NSTATE;
    const int oz = NZ / C.obs_nz, ox = NX / C.obs_nx, nobs = C.obs_nz * C.obs_nx;
    double nu_s = 0, nu_o = 0;
    RBC_PHASE(phase_reduce_partials(tid, C, cur, red);)
#if defined(__CUDA_ARCH__)
    __shared__ double fin[12];
#else
    double fin[12];
#endif
    RBC_PHASE(if (tid < NRED) fin[tid] = sum_serial(red + tid * NT, NT);)
    RBC_PHASE(phase_reduce_obs_partials(tid, C, cur, red);)
    RBC_PHASE(if (tid < 4) fin[NRED + tid] = sum_serial(red + tid * NT, NT);)
    {
        const double kap = C.kappa_d, dbH = kap * 1.0 / 2.0;    // kappa * db / H, db = 1, H = 2
        const double q1 = fin[0] / (double)NCELL, q1o = fin[1] / (double)nobs;
        const double T0 = fin[2] / NX, T1 = fin[3] / NX, Tm2 = fin[4] / NX, Tm1 = fin[5] / NX;
        const double g = (1.5 * Tm1 - 0.5 * Tm2 + 0.5 * T1 - 1.5 * T0) / NZ;
        nu_s = (q1 - kap * g) / dbH;
        const double o0 = fin[7] / C.obs_nx, o1 = fin[8] / C.obs_nx, om2 = fin[9] / C.obs_nx, om1 = fin[10] / C.obs_nx;
        const double go = (1.5 * om1 - 0.5 * om2 + 0.5 * o1 - 1.5 * o0) / C.obs_nz;
        nu_o = (q1o - kap * go) / dbH;
    }
    // decisions every thread takes alike (fin is shared; t_old / pend were read before the march)
    const bool vec = V.mode >= 0 && F.advance_clock;
    const bool bad = fin[6] > 0;
    const bool stepping = F.advance_clock && !pend && !second_pass;
    const bool trunc_now = stepping && (t_old + C.dt_action >= C.episode_length);
    const bool bad_reset = vec && stepping && V.nan_reset && bad;
    const bool do_reset = vec && stepping && V.bank != nullptr && ((V.mode == 2 && trunc_now) || bad_reset);
    float* ob = ((do_reset && V.final_obs != nullptr) ? V.final_obs : io.obs) + (size_t)env * C.channels * nobs;
    RBC_PHASE(
        // store the state back and emit the observation (strided sub-sample, channel-major)
        if (state_changed && !do_reset) phase_store_state(tid, cur, st);
        for (int q = tid; q < 3 * nobs; q += NT) {
            const int ch = q / nobs, zo = (q % nobs) / C.obs_nx, xo = q % C.obs_nx;
            const int off = (ch == 0 ? OFF_B : (ch == 1 ? OFF_U : OFF_W));
            float v = (float)cur[off + (zo * oz) * SX + xo * ox];
            if (C.wrap_obs) {
                v = C.obs_maxval * (2.0f * (v - C.obs_lo[ch]) / (C.obs_hi[ch] - C.obs_lo[ch]) - 1.0f);
                if (C.obs_clip) v = fminf(fmaxf(v, -C.obs_maxval), C.obs_maxval);
            }
            ob[q] = v;
        }
        if (tid == 0) {
            if (second_pass) {
                // the environment was re-initialised inside this step: its new observation is already out, the
                // Nusselt outputs follow the new state (the terminal ones went to final_*), the clock restarts
                io.nu_state[env] = nu_s;
                io.nu_obs[env] = nu_o;
                io.t[env] = 0.0;
                io.step_count[env] = 1;
                V.episode[env] += 1;
                V.ep_return[env] = 0.0;
                if (V.pending != nullptr) V.pending[env] = 0;
            } else {
                double rew = -nu_o;                                        // rbc2D.py:198-200
                if (C.wrap_reward) rew = (rew + C.reward_scale) / (C.reward_scale - 1.0);
                if (C.wrap_shaping || io.cell_dist != nullptr) {
                    const double PI = 3.14159265358979323846;
                    const double cd = cell_distance(cur + OFF_W + (NZ / 2 - 1) * SX);
                    if (io.cell_dist != nullptr) io.cell_dist[env] = cd;
                    if (C.wrap_shaping) rew = (1.0 - C.shaping_weight) * rew + C.shaping_weight * ((PI - cd) / PI);
                }
                io.nan_flag[env] = bad ? 1 : 0;
                if (pend) {
                    // gymnasium NEXT_STEP: this call only resets the environment (action ignored, reward 0)
                    io.nu_state[env] = nu_s;
                    io.nu_obs[env] = nu_o;
                    io.reward[env] = 0.0f;
                    io.truncated[env] = 0;
                    io.t[env] = 0.0;
                    io.step_count[env] = 1;
                    V.episode[env] += 1;
                    V.ep_return[env] = 0.0;
                    V.pending[env] = 0;
                } else {
                    if (bad_reset) rew = 0.0;
                    if (bad && vec && V.nan_count != nullptr) RBC_COUNT_ONE(V.nan_count);
                    io.reward[env] = (float)rew;
                    const double ret = vec ? V.ep_return[env] + (double)(float)rew : 0.0;
                    if (do_reset) {
                        if (V.final_nu_a != nullptr) V.final_nu_a[env] = nu_s;
                        if (V.final_nu_b != nullptr) V.final_nu_b[env] = nu_o;
                        if (V.final_return != nullptr) V.final_return[env] = ret;
                        io.truncated[env] = 1;
                    } else {
                        io.nu_state[env] = nu_s;
                        io.nu_obs[env] = nu_o;
                        if (F.advance_clock) {
                            const double tn = t_old + C.dt_action;
                            io.t[env] = tn;
                            io.step_count[env] += 1;
                            io.truncated[env] = (trunc_now || bad_reset) ? 1 : 0;
                        }
                        if (vec) {
                            V.ep_return[env] = ret;
                            if (V.pending != nullptr) V.pending[env] = (V.mode == 1 && trunc_now && V.bank != nullptr) ? 1 : 0;
                        }
                    }
                }
            }
        }
    )
    if (io.pressure != nullptr) {
        // pressure channels of get_state (rbc_sim2D_api.jl:114-115): pHY' from the final b (as the
        // last update_state! leaves it) and pNHS = phi / dtau of the last stage, zero-mean gauge.  They are refreshed
        // whenever the state changed (step, or the set! projection of a reset) and kept per environment, so that an
        // observe-only launch (what reset() returns) emits the channels from the stored fields.
        Real* pr = io.pressure + (size_t)env * 2 * NCELL;
        if (state_changed) {
            RBC_PHASE(
                double acc = 0;
                for (int q = tid; q < NCELL; q += NT) acc += (double)X.R[(q / NX) * RSTR + (q % NX)];
                red[tid] = acc;
            )
            RBC_PHASE(if (tid == 0) fin[11] = sum_serial(red, NT) / (double)NCELL;)
            RBC_PHASE(
                for (int q = tid; q < NCELL; q += NT)
                    pr[NCELL + q] = (Real)(((double)X.R[(q / NX) * RSTR + (q % NX)] - fin[11]) / (double)last_dtau);
            )
            RBC_PHASE(phase_phy(tid, C, cur + OFF_B, X.R);)
            RBC_PHASE(for (int q = tid; q < NCELL; q += NT) pr[q] = X.R[(q / NX) * RSTR + (q % NX)];)
        }
        if (C.channels == 5) {
            RBC_PHASE(
                for (int q = tid; q < 2 * nobs; q += NT) {
                    const int ch = q / nobs, zo = (q % nobs) / C.obs_nx, xo = q % C.obs_nx;
                    ob[3 * nobs + q] = (float)pr[ch * NCELL + (zo * oz) * NX + xo * ox];
                }
            )
        }
    }
    return do_reset;
}

// ------------------------------------------------------------------------------------------
// one action step of one environment (the whole of step_simulation + observation + reward)
// ------------------------------------------------------------------------------------------
template <typename Real, bool SPLIT, bool NXT_GLOBAL>
RBC_HD void env_action_step(const Consts<Real>& C, const Tables<Real>& T, const EnvIO<Real>& io, const Ctx<Real>& X, int env,
                            const RunFlags& F)
{
    const Real gam[3] = {Real(8.0 / 15.0), Real(5.0 / 12.0), Real(3.0 / 4.0)};
    const Real zet[3] = {Real(0), Real(-17.0 / 60.0), Real(-5.0 / 12.0)};
    Real* st = io.state + (size_t)env * NSTATE;
    const VecIO& V = io.vec;
    // fused vector-env semantics: an environment that truncated on the previous call (next_step mode) is only
    // re-initialised by this one — gather from the checkpoint bank instead of loading its state, no march
    const int pend = (V.mode == 1 && V.bank != nullptr && F.advance_clock) ? V.pending[env] : 0;
    const int nsub = pend ? 0 : F.nsub;
    const bool project_first = pend ? SPLIT : (F.project_first != 0);

    // load the environment, evaluate the heater profile for this action
    if (pend) {
        const double* src = V.bank + (size_t)checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * NSTATE;
        RBC_PHASE(phase_load_bank(tid, src, X.s0);)
    } else {
        RBC_PHASE(
            phase_load_state(tid, st, X.s0);
            if (tid < NX) X.Tb[tid] = (Real)heater_T(C, io.actions + (size_t)env * C.heaters, (tid + 0.5) * C.dx);
        )
    }
    Real* cur = X.s0;
    Real* nxt = X.s1;
    Real last_dtau = Real(1);
    if (project_first) project(C, X, cur, T.thomas_scale);
    for (int sub = 0; sub < nsub; ++sub) {
        const Real dt = (sub == nsub - 1) ? C.dt_last : C.dt_full;
        for (int stage = 0; stage < 3; ++stage) {
            if (SPLIT) { RBC_PHASE(phase_phy(tid, C, cur + OFF_B, X.R);) }
            const Real* gin = X.gm + ((stage & 1) ? 0 : NSTATE);      // stage s reads what stage s-1 wrote
            Real* gout = X.gm + ((stage & 1) ? NSTATE : 0);
            RBC_PHASE(phase_edge_fluxes(tid, cur, X.E);)
            RBC_PHASE((phase_tendency<Real, SPLIT>(tid, C, cur, nxt, X.R, X.Tb, X.E, gin, gout, dt, gam[stage], zet[stage], stage > 0));)
            Real* P;
            if (NXT_GLOBAL) {
                RBC_PHASE(phase_copy(tid, nxt, cur, NS_SM);)
                P = cur;
            } else {
                P = nxt; nxt = cur; cur = P;
            }
            project(C, X, P, T.thomas_scale);
            last_dtau = (gam[stage] + zet[stage]) * dt;
        }
    }

    // the reduction scratch aliases the dead second state buffer when that one is on-chip
    double* red = NXT_GLOBAL ? X.red : reinterpret_cast<double*>(nxt);
    // one inlined copy of the epilogue, run a second time when the environment is re-initialised inside this launch
    // (same_step auto-reset or a NaN reset: terminal outputs went to final_*, the next episode's start is gathered here)
    bool changed = nsub > 0 || project_first;
    RBC_NOUNROLL
    for (int pass = 0; pass < 2; ++pass) {
        if (!env_epilogue<Real, SPLIT, NXT_GLOBAL>(C, io, X, env, cur, red, F, changed, last_dtau, pass == 1)) break;
        const double* src = V.bank + (size_t)checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * NSTATE;
        RBC_PHASE(phase_load_bank(tid, src, cur);)
        if (SPLIT) project(C, X, cur, T.thomas_scale);
        changed = true; last_dtau = Real(1);
    }
}

// ------------------------------------------------------------------------------------------
// host-side table construction (fp64), shared by the library and the emulator
// ------------------------------------------------------------------------------------------
// run!'s substep schedule dt' = min(dt_solver, stop - t) from t = 0 (SURVEY 8a "run! semantics"):
// (nsub-1) full steps + one clipped step; remainders below 1e-10 are dropped.
inline int substep_schedule(double dt_action, double dt_solver, double* dt_last)
{
    int n = 0;
    double t = 0.0, last = dt_solver;
    while (dt_action - t > 1e-10) {
        const double d = (dt_solver < dt_action - t) ? dt_solver : dt_action - t;
        last = d; t += d; ++n;
    }
    *dt_last = last;
    return n;
}
struct HostConfig {
    double ra, pr, lx, lz, b_top, heater_limit, dt_action, dt_solver, episode_length;
    int heaters, obs_nz, obs_nx, channels;
    double cfl_limit;            // CFL guard (cluster kernels); 0 = off.  Last member: brace-initialisers that stop before it leave it 0
};
// optional fused wrappers (all off by default)
struct HostWrappers {
    int normalize_obs = 0, obs_clip = 0;
    float obs_lo[4] = {0, 0, 0, 0}, obs_hi[4] = {1, 1, 1, 1}, obs_maxval = 1.0f;
    int normalize_reward = 0;
    double reward_scale = 2.0;
    int shaping = 0;
    double shaping_weight = 0.0;
};
template <typename Real>
inline Consts<Real> make_consts(const HostConfig& h, const HostWrappers& w = HostWrappers())
{
    Consts<Real> C;
    const double dx = h.lx / NX, dz = h.lz / NZ;
    const double nu = sqrt(h.pr / h.ra), kappa = 1.0 / sqrt(h.pr * h.ra);   // rbc_sim2D_api.jl:40-41
    C.idx = (Real)(1.0 / dx); C.idz = (Real)(1.0 / dz);
    C.idx2 = (Real)(1.0 / (dx * dx)); C.idz2 = (Real)(1.0 / (dz * dz));
    C.nu = (Real)nu; C.kappa = (Real)kappa; C.b_top = (Real)h.b_top;
    double last;
    C.nsub = substep_schedule(h.dt_action, h.dt_solver, &last);
    C.dt_full = (Real)h.dt_solver; C.dt_last = (Real)last;
    C.heaters = h.heaters; C.heater_limit = h.heater_limit; C.lx = h.lx; C.dx = dx;
    C.kappa_d = kappa; C.dt_action = h.dt_action; C.episode_length = h.episode_length;
    C.obs_nz = h.obs_nz; C.obs_nx = h.obs_nx; C.channels = h.channels;
    C.wrap_obs = w.normalize_obs; C.obs_clip = w.obs_clip; C.obs_maxval = w.obs_maxval;
    for (int c = 0; c < 4; ++c) { C.obs_lo[c] = w.obs_lo[c]; C.obs_hi[c] = w.obs_hi[c]; }
    C.wrap_reward = w.normalize_reward; C.reward_scale = w.reward_scale;
    C.wrap_shaping = w.shaping; C.shaping_weight = w.shaping_weight;
    C.cfl_limit = Real(0);       // the dedicated 96 x 64 kernel runs the reference's fixed dt (CFL <= 1.2 over the shipped Ra range)
    return C;
}
// mode index of spectral word t of a row after fft_untangle
inline int word_mode(int t)
{
    const int p = t / 2;
    if (p == 0) return (t == 0) ? 0 : NX / 2;
    return 8 * (p % 6) + p / 6;
}
inline void build_tables_host(double lx, double lz, double* tinv /*NZ*NX*/, double* tw48 /*96*/, double* tw96 /*96*/)
{
    const double PI = 3.14159265358979323846;
    const double dx = lx / NX, dz = lz / NZ;
    for (int t = 0; t < NX; ++t) {
        const int m = word_mode(t);
        const double sx = 2.0 * sin(PI * m / NX) / dx, lam = sx * sx * dz * dz;
        auto diag = [&](int k) {
            double dg = -(2.0 + lam);
            if (k == 0 || k == NZ - 1) dg += 1.0;
            if (m == 0 && k == 0) dg -= 1.0;            // pin the null space of the mean mode
            return dg;
        };
        double prev = 0.0;
        for (int k = 0; k < NZ / 2; ++k) {              // upward sweep pivots
            const double iv = 1.0 / (diag(k) - prev);
            tinv[k * NX + t] = iv;
            prev = iv;
        }
        prev = 0.0;
        for (int k = NZ - 1; k >= NZ / 2; --k) {        // downward sweep pivots
            const double iv = 1.0 / (diag(k) - prev);
            tinv[k * NX + t] = iv;
            prev = iv;
        }
    }
    for (int j = 0; j < 48; ++j) { tw48[2 * j] = cos(2 * PI * j / 48); tw48[2 * j + 1] = sin(2 * PI * j / 48); }
    for (int m = 0; m < 48; ++m) { tw96[2 * m] = cos(2 * PI * m / 96); tw96[2 * m + 1] = sin(2 * PI * m / 96); }
}

}  // namespace rbc2d
