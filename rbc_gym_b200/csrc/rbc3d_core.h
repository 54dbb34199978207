// rbc3d_core.h — the 3D Rayleigh-Benard action step (RayleighBenardConvection3D-v0) as one CTA per environment.
//
// Replaces, for the 3D hot path, the reference's
//   step_simulation -> run!(simulation), preprocess_action, bottom_T(x,y,t)
//                                                   src/rbc_gym/sim/rbc_sim3D_api.jl:77-101, rbc_sim3D.jl:111-141
//   NonhydrostaticModel configuration               src/rbc_gym/sim/rbc_sim3D.jl:99-166
//   get_state / get_nusselt / NaN check             src/rbc_gym/sim/rbc_sim3D_api.jl:106-159
// (arithmetic of run!: Oceananigans.jl 0.92.0; scheme = SURVEY.md 8a with a second periodic direction.)
//
// Layout of one environment (global memory, x fastest): b,u,v [NZ][NY][NX], w [NZ+1][NY][NX] = 66 560 values
// (u on the x-face left of cell i, v on the y-face south of cell j, w on the z-face below cell k).  At fp32 that
// is 266 KB — more than a CTA's shared memory — so the fields stay in L2-resident global buffers (read through
// L1) and only the Poisson right-hand side (70 KB) lives in shared memory.  A thread owns two (i,j) columns and
// marches each from wall to wall with 7-level register windows of its own column: vertical fluxes are computed
// once and carried, every wall-order decision is resolved at compile time (the level is the unrolled loop index),
// and the x/y neighbours are coalesced 128-byte lines (a warp = one row of 32 columns).
//
// Projection: per-thread in-register FFTs — one thread transforms one 32-point real row in x (complex-16 +
// real-FFT split), then one thread transforms one 32-point complex column in y; the z direction is 1024
// independent tridiagonal systems of length 16 per environment (two per thread), then the inverses.
//
//
// Tiled variant (fp32 throughput mode, buoyancy in G_w): the tendency phase runs in two halves of 16 rows (y).
// For each half the CTA first copies the 22 rows it touches (16 + 3 halo rows per side, periodic wrap) of all
// levels of b,u,v,w into shared memory (183 KB, aliasing the Poisson scratch that is dead during this phase);
// the march then takes every stencil value — its own column and the x/y neighbours — from that tile, so the only
// global traffic of the march is the previous tendencies in and the new tendencies + predicted state out.
//
// Shared with nvcc and g++ (tests/emu) exactly like rbc2d_core.h.
#pragma once
#include "rbc2d_core.h"

#define RBC3_PHASE(...) RBC_PHASE_N(rbc3d::NT, __VA_ARGS__)

namespace rbc3d {

using rbc2d::BoolTag;
using rbc2d::cadd;
using rbc2d::centred4;
using rbc2d::centred_ord;
using rbc2d::cmul;
using rbc2d::csub;
using rbc2d::cx;
using rbc2d::dft4;
using rbc2d::dft8;
using rbc2d::ldc;
using rbc2d::stc;
using rbc2d::sum_serial;
using rbc2d::upwind5;
using rbc2d::upwind_ord;

constexpr int NX = 32, NY = 32, NZ = 16;
constexpr int NT = 512;                         // threads per CTA; two columns per thread
constexpr int NCOL = NX * NY;                   // 1024 columns
constexpr int NC = NCOL * NZ;                   // 16384 cells
constexpr int NW = NCOL * (NZ + 1);             // 17408 w values
constexpr int NSTATE = 3 * NC + NW;             // 66560 values per environment
constexpr int GB = 0, GU = NC, GV = 2 * NC, GW = 3 * NC;
constexpr int NG = 4 * NC;                      // tendency slab: b,u,v,w(faces 0..NZ-1)
constexpr int RX = 34;                          // padded row of the Poisson scratch (34 = 32 + 2: lanes on consecutive rows
constexpr int RP = NY * RX;                     //  walk the banks two at a time for 64-bit accesses)
constexpr int NR = NZ * RP;
constexpr int MAX_HEATERS = 16;
// shared-memory tile of the tiled tendency phase: TROWS rows of all levels of b,u,v (NZ planes each) and w (NZ+1)
constexpr int THALF = NY / 2, TROWS = THALF + 6, TPL = TROWS * NX;
constexpr int NTILE = (3 * NZ + NZ + 1) * TPL;  // 45 760 values

template <typename Real>
struct Consts3 {
    Real idx, idy, idz, idx2, idy2, idz2;
    Real nu, kappa, b_top;
    Real dt_full, dt_last;       // simulation-time substeps: dt_solver*t_ff, last one clipped (rbc_sim3D_api.jl:65)
    int nsub;
    int heaters;                 // patches per side (8)
    double heater_limit, b_hot;  // b_hot = min_b + delta_b
    double kappa_d, delta_b_d, b_top_d;
    double dt_action;            // heater_duration * t_ff, what `time` advances by (rbc_sim3D_api.jl:91)
    double episode_length;
};

template <typename Real>
struct EnvIO3 {
    Real* state;              // [B][NSTATE]
    const float* actions;     // [B][heaters][heaters], action[i][j] <-> patch i along x, j along y
    float* obs;               // [B][4][NZ][NY][NX] = get_state in the Python layout (rbc3D.py:229-232), or nullptr
    float* reward;            // [B] = -nusselt
    double* nusselt;          // [B]
    double* t;                // [B]
    int* step_count;
    int* truncated;
    int* nan_flag;
    rbc2d::VecIO vec;         // fused vector-env semantics (rbc2d_core.h); final_nu_a = terminal nusselt, final_obs [B][4][NZ][NY][NX]
};

template <typename Real>
struct Ctx3 {
    Real* bufA;               // global, per CTA: predicted-state ping-pong buffers (NSTATE each)
    Real* bufB;
    Real* gm;                 // global, per CTA: the tendency slab (NG), updated in place
    Real* R;                  // shared: Poisson scratch / pHY' [NZ][NY][RX]
    Real* tile;               // shared (tiled variant): NTILE values, aliases R; nullptr otherwise
    Real* Tb;                 // shared: bottom wall temperature per column (NCOL)
    double* red;              // shared: NT doubles x 2 for the epilogue reductions
    const Real* tinv;         // global: Thomas pivots [NZ][NY][NX]
    Real thomas_scale;        // dz^2 / (16*32)
};

struct RunFlags3 {
    int nsub, project_first, advance_clock;
};

RBC_HD int wrapn(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

// wall-order rules for NZ levels: centres->z-face kf, z-faces->centre kc
RBC_HD constexpr int o_up_face(int kf) { return (kf >= 3 && kf <= NZ - 3) ? 5 : ((kf == 2 || kf == NZ - 2) ? 3 : 1); }
RBC_HD constexpr int o_ce_face(int kf) { return (kf >= 2 && kf <= NZ - 2) ? 4 : 2; }
RBC_HD constexpr int o_up_cen(int kc) { return (kc >= 2 && kc <= NZ - 3) ? 5 : ((kc == 1 || kc == NZ - 2) ? 3 : 1); }
RBC_HD constexpr int o_ce_cen(int kc) { return (kc >= 1 && kc <= NZ - 2) ? 4 : 2; }

RBC_HD constexpr double tw16c(int m) { constexpr double t[16] = {1, 0.92387953251128674, 0.70710678118654757, 0.38268343236508984, 6.123233995736766e-17, -0.38268343236508973, -0.70710678118654746, -0.92387953251128674, -1, -0.92387953251128685, -0.70710678118654768, -0.38268343236509034, -1.8369701987210297e-16, 0.38268343236509, 0.70710678118654735, 0.92387953251128652}; return t[m]; }
RBC_HD constexpr double tw16s(int m) { constexpr double t[16] = {0, 0.38268343236508978, 0.70710678118654746, 0.92387953251128674, 1, 0.92387953251128674, 0.70710678118654757, 0.38268343236508989, 1.2246467991473532e-16, -0.38268343236508967, -0.70710678118654746, -0.92387953251128652, -1, -0.92387953251128663, -0.70710678118654768, -0.38268343236509039}; return t[m]; }
RBC_HD constexpr double tw32c(int m) { constexpr double t[32] = {1, 0.98078528040323043, 0.92387953251128674, 0.83146961230254524, 0.70710678118654757, 0.55557023301960229, 0.38268343236508984, 0.19509032201612833, 6.123233995736766e-17, -0.19509032201612819, -0.38268343236508973, -0.55557023301960196, -0.70710678118654746, -0.83146961230254535, -0.92387953251128674, -0.98078528040323043, -1, -0.98078528040323043, -0.92387953251128685, -0.83146961230254546, -0.70710678118654768, -0.55557023301960218, -0.38268343236509034, -0.19509032201612866, -1.8369701987210297e-16, 0.1950903220161283, 0.38268343236509, 0.55557023301960184, 0.70710678118654735, 0.83146961230254524, 0.92387953251128652, 0.98078528040323032}; return t[m]; }
RBC_HD constexpr double tw32s(int m) { constexpr double t[32] = {0, 0.19509032201612825, 0.38268343236508978, 0.55557023301960218, 0.70710678118654746, 0.83146961230254524, 0.92387953251128674, 0.98078528040323043, 1, 0.98078528040323043, 0.92387953251128674, 0.83146961230254546, 0.70710678118654757, 0.55557023301960218, 0.38268343236508989, 0.19509032201612861, 1.2246467991473532e-16, -0.19509032201612836, -0.38268343236508967, -0.55557023301960196, -0.70710678118654746, -0.83146961230254524, -0.92387953251128652, -0.98078528040323032, -1, -0.98078528040323043, -0.92387953251128663, -0.83146961230254546, -0.70710678118654768, -0.55557023301960218, -0.38268343236509039, -0.19509032201612872}; return t[m]; }

// ------------------------------------------------------------------------------------------
// heater patches: preprocess_action + bottom_T (rbc_sim3D.jl:111-141), fp64 like the reference
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD double heater_T3(const Consts3<Real>& C, const float* action, int i, int j)
{
    const int h = C.heaters, n = h * h;
    double mean = 0.0;
    for (int q = 0; q < n; ++q) mean += (double)action[q];
    mean /= n;
    double K = 1.0;
    for (int q = 0; q < n; ++q) {
        const double d = fabs((double)action[q] - mean);
        if (d > K) K = d;
    }
    int pi = (int)floor((i + 0.5) / NX * h), pj = (int)floor((j + 0.5) / NY * h);
    pi = pi < 0 ? 0 : (pi > h - 1 ? h - 1 : pi);
    pj = pj < 0 ? 0 : (pj > h - 1 ? h - 1 : pj);
    return C.b_hot + (((double)action[pi * h + pj] - mean) / K) * C.heater_limit;
}

// ------------------------------------------------------------------------------------------
// phase: hydrostatic pressure anomaly (split mode), one thread per column, top-down integral -> R
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void phase_phy3(int tid, const Consts3<Real>& C, const Real* cb, Real* phy)
{
    const Real dz = Real(1) / C.idz;
    for (int q = 0; q < NCOL / NT; ++q) {
        const int c = tid + q * NT, i = c % NX, j = c / NX;
        Real below = cb[(NZ - 1) * NCOL + c];
        Real acc = -Real(0.5) * (below + (Real(2) * C.b_top - below)) * dz;
        phy[(NZ - 1) * RP + j * RX + i] = acc;
        for (int k = NZ - 2; k >= 0; --k) {
            const Real bk = cb[k * NCOL + c];
            acc = acc - Real(0.5) * (bk + below) * dz;
            phy[k * RP + j * RX + i] = acc;
            below = bk;
        }
    }
}

// ------------------------------------------------------------------------------------------
// phase: tendencies + RK3 substep; reads `cur`, writes the predicted state `nxt`, stores G for the next stage
// ------------------------------------------------------------------------------------------
// The x-direction flux through the right face of column i is the flux through the left face of column i+1, and with
// NX = 32 the x-neighbours of a cell are the neighbouring lanes of its warp (periodic): on the device every thread
// evaluates only the left fluxes and takes the right ones from lane+1 with a shuffle — 4 of the 20 upwind
// reconstructions per cell, the u(i+1, .) windows and their loads go away.  The host emulator (threads run one after
// the other) evaluates both faces; the arithmetic is identical.
#if defined(__CUDA_ARCH__)
#define RBC3_SHARE_X 1
template <typename Real>
__device__ __forceinline__ Real from_next_lane(Real v) { return __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31); }
#else
#define RBC3_SHARE_X 0
template <typename Real>
inline Real from_next_lane(Real v) { return v; }
#endif
static_assert(NX == 32, "x-flux sharing maps one x-row of cells onto one warp");

template <typename Real, bool SPLIT, bool TILE>
RBC_HD void tendency3_column(const Consts3<Real>& C, const Real* RBC_RESTRICT src, Real* nxt, const Real* RBC_RESTRICT phy,
                             const Real* RBC_RESTRICT Tb, const Real* gm_in, Real* gm_out, Real dt, Real gam, Real zet, bool use_gm,
                             const int i, const int j, const int jl)
{
    // `src` is the environment's state in global memory (level stride NCOL, rows wrap periodically) or the
    // shared-memory tile of this half (level stride TPL, tile row = jl + 3 + dj, no wrap needed)
    constexpr int LS = TILE ? TPL : NCOL;
    constexpr int SB = 0, SU = NZ * LS, SV = 2 * NZ * LS, SW = 3 * NZ * LS;
    auto rowoff = [&](int dj) { return TILE ? (jl + 3 + dj) * NX : wrapn(j + dj, NY) * NX; };
    {
        const int c = j * NX + i;                          // global column (outputs, tendency slabs, Tb)
        const int cs = rowoff(0) + i;                      // the same column in `src`
        int xo[7], yo[7];
        RBC_UNROLL
        for (int d = 0; d < 7; ++d) { xo[d] = rowoff(0) + wrapn(i + d - 3, NX); yo[d] = rowoff(d - 3) + i; }
        const int c_ip = xo[4], c_jp = yo[4];          // columns (i+1, j) and (i, j+1)
        // u(i+1, j-2..j+1) and v(i-2..i+1, j+1) columns for the mixed fluxes Uv(i+1) and Vu(j+1)
        const int ip = wrapn(i + 1, NX);
        const int cu0 = rowoff(-2) + ip, cu1 = rowoff(-1) + ip, cu3 = rowoff(1) + ip;
        const int cv0 = rowoff(1) + wrapn(i - 2, NX), cv1 = rowoff(1) + wrapn(i - 1, NX), cv3 = cu3;

        // own-column windows: index jj <-> level k-3+jj
        Real bz[7], uz[7], vz[7], wz[7], u1z[4], v1z[4], wx[7], wy[7];
        RBC_UNROLL
        for (int jj = 0; jj < 7; ++jj) {
            const int k = jj - 3;
            const bool ok = k >= 0;
            bz[jj] = ok ? src[SB + k * LS + cs] : Real(0);
            uz[jj] = ok ? src[SU + k * LS + cs] : Real(0);
            vz[jj] = ok ? src[SV + k * LS + cs] : Real(0);
            wz[jj] = ok ? src[SW + k * LS + cs] : Real(0);
        }
        RBC_UNROLL
        for (int jj = 0; jj < 4; ++jj) {
            const int k = jj - 2;
            u1z[jj] = k >= 0 ? src[SU + k * LS + c_ip] : Real(0);
            v1z[jj] = k >= 0 ? src[SV + k * LS + c_jp] : Real(0);
        }
        RBC_UNROLL
        for (int d = 0; d < 7; ++d) { wx[d] = src[SW + xo[d]]; wy[d] = src[SW + yo[d]]; }     // wall face 0
        Real Fzb_lo = Real(0), Wu_lo = Real(0), Wv_lo = Real(0), Ww_lo = Real(0);
        const Real tb = Tb[c];
        Real gn[4] = {Real(0), Real(0), Real(0), Real(0)};
        if (use_gm) {
            RBC_UNROLL
            for (int f = 0; f < 4; ++f) gn[f] = gm_in[f * NC + c];
        }

        auto level = [&](auto edge_tag, const int k) {
            constexpr bool EDGE = decltype(edge_tag)::value;
            const Real g0[4] = {gn[0], gn[1], gn[2], gn[3]};
            if (k + 1 < NZ && use_gm) {
                RBC_UNROLL
                for (int f = 0; f < 4; ++f) gn[f] = gm_in[f * NC + (k + 1) * NCOL + c];
            }
            const Real* lb = src + SB + k * LS;
            const Real* lu = src + SU + k * LS;
            const Real* lv = src + SV + k * LS;
            const Real* lwn = src + SW + (k + 1) * LS;         // face k+1 <= NZ always exists
            Real bx[7], by[7], ux[7], uy[7], vx[7], vy[7], wxn[7], wyn[7];
            RBC_UNROLL
            for (int d = 0; d < 7; ++d) {
                bx[d] = (d == 3) ? bz[3] : lb[xo[d]];
                by[d] = (d == 3) ? bz[3] : lb[yo[d]];
                ux[d] = (d == 3) ? uz[3] : ((d == 4) ? u1z[2] : lu[xo[d]]);
                uy[d] = (d == 3) ? uz[3] : lu[yo[d]];
                vx[d] = (d == 3) ? vz[3] : lv[xo[d]];
                vy[d] = (d == 3) ? vz[3] : ((d == 4) ? v1z[2] : lv[yo[d]]);
                wxn[d] = (d == 3) ? wz[4] : lwn[xo[d]];
                wyn[d] = (d == 3) ? wz[4] : lwn[yo[d]];
            }
            const Real uc0 = lu[cu0], uc1 = lu[cu1], uc2 = ux[4], uc3 = lu[cu3];      // u(i+1, j-2..j+1)
            const Real vc0 = lv[cv0], vc1 = lv[cv1], vc2 = vy[4], vc3 = lv[cv3];      // v(i-2..i+1, j+1)
            const bool top = EDGE && (k == NZ - 1);
            const bool bot = EDGE && (k == 0);
            const int of_hi = EDGE ? o_up_face(k + 1) : 5;
            const int oc_face = EDGE ? o_ce_face(k) : 4;
            const int ou_cen = EDGE ? o_up_cen(k) : 5;
            const int oc_cen = EDGE ? o_ce_cen(k) : 4;

            // ---- tracer ----
            const Real Fx0 = upwind5(ux[3], bx), Fx1 = RBC3_SHARE_X ? from_next_lane(Fx0) : upwind5(ux[4], bx + 1);
            const Real Fy0 = upwind5(vy[3], by), Fy1 = upwind5(vy[4], by + 1);
            const Real Fzb_hi = top ? Real(0) : upwind_ord(wz[4], bz + 1, of_hi);
            const Real bdn = bot ? (Real(2) * tb - bz[3]) : bz[2];
            const Real bup = top ? (Real(2) * C.b_top - bz[3]) : bz[4];
            const Real Gb = -((Fx1 - Fx0) * C.idx + (Fy1 - Fy0) * C.idy + (Fzb_hi - Fzb_lo) * C.idz) +
                            C.kappa * ((bx[4] - Real(2) * bx[3] + bx[2]) * C.idx2 + (by[4] - Real(2) * by[3] + by[2]) * C.idy2 +
                                       (bup - Real(2) * bz[3] + bdn) * C.idz2);
            // ---- u at (x-face i, j, k) ----
            const Real uF0 = upwind5(centred4(ux[1], ux[2], ux[3], ux[4]), ux);
            const Real uF1 = RBC3_SHARE_X ? from_next_lane(uF0) : upwind5(centred4(ux[2], ux[3], ux[4], ux[5]), ux + 1);
            const Real uG0 = upwind5(centred4(vx[1], vx[2], vx[3], vx[4]), uy);          // (x-face i, y-face j)
            const Real uG1 = upwind5(centred4(vc0, vc1, vc2, vc3), uy + 1);              // (x-face i, y-face j+1)
            const Real Wu_hi = top ? Real(0) : upwind_ord(centred4(wxn[1], wxn[2], wxn[3], wxn[4]), uz + 1, of_hi);
            const Real udn = bot ? -uz[3] : uz[2], uup = top ? -uz[3] : uz[4];
            Real Gu = -((uF1 - uF0) * C.idx + (uG1 - uG0) * C.idy + (Wu_hi - Wu_lo) * C.idz) +
                      C.nu * ((ux[4] - Real(2) * ux[3] + ux[2]) * C.idx2 + (uy[4] - Real(2) * uy[3] + uy[2]) * C.idy2 +
                              (uup - Real(2) * uz[3] + udn) * C.idz2);
            // ---- v at (i, y-face j, k) ----
            const Real vF0 = upwind5(centred4(uy[1], uy[2], uy[3], uy[4]), vx);          // (x-face i,   y-face j)
            const Real vF1 = RBC3_SHARE_X ? from_next_lane(vF0) : upwind5(centred4(uc0, uc1, uc2, uc3), vx + 1);   // (x-face i+1, y-face j)
            const Real vG0 = upwind5(centred4(vy[1], vy[2], vy[3], vy[4]), vy);
            const Real vG1 = upwind5(centred4(vy[2], vy[3], vy[4], vy[5]), vy + 1);
            const Real Wv_hi = top ? Real(0) : upwind_ord(centred4(wyn[1], wyn[2], wyn[3], wyn[4]), vz + 1, of_hi);
            const Real vdn = bot ? -vz[3] : vz[2], vup = top ? -vz[3] : vz[4];
            Real Gv = -((vF1 - vF0) * C.idx + (vG1 - vG0) * C.idy + (Wv_hi - Wv_lo) * C.idz) +
                      C.nu * ((vx[4] - Real(2) * vx[3] + vx[2]) * C.idx2 + (vy[4] - Real(2) * vy[3] + vy[2]) * C.idy2 +
                              (vup - Real(2) * vz[3] + vdn) * C.idz2);
            if (SPLIT) {
                const Real* ph = phy + k * RP;
                const Real p0 = ph[j * RX + i];
                Gu -= (p0 - ph[j * RX + wrapn(i - 1, NX)]) * C.idx;
                Gv -= (p0 - ph[wrapn(j - 1, NY) * RX + i]) * C.idy;
            }
            // ---- w at (i, j, z-face k); face 0 is the wall ----
            const Real ut0 = centred_ord(uz[1], uz[2], uz[3], uz[4], oc_face), ut1 = centred_ord(u1z[0], u1z[1], u1z[2], u1z[3], oc_face);
            const Real vt0 = centred_ord(vz[1], vz[2], vz[3], vz[4], oc_face), vt1 = centred_ord(v1z[0], v1z[1], v1z[2], v1z[3], oc_face);
            const Real wF0 = upwind5(ut0, wx), wF1 = RBC3_SHARE_X ? from_next_lane(wF0) : upwind5(ut1, wx + 1);
            const Real wG0 = upwind5(vt0, wy), wG1 = upwind5(vt1, wy + 1);
            const Real Ww_hi = upwind_ord(centred_ord(wz[2], wz[3], wz[4], wz[5], oc_cen), wz + 1, ou_cen);
            Real Gw = -((wF1 - wF0) * C.idx + (wG1 - wG0) * C.idy + (Ww_hi - Ww_lo) * C.idz) +
                      C.nu * ((wx[4] - Real(2) * wx[3] + wx[2]) * C.idx2 + (wy[4] - Real(2) * wy[3] + wy[2]) * C.idy2 +
                              (wz[4] - Real(2) * wz[3] + wz[2]) * C.idz2);
            if (!SPLIT) Gw += Real(0.5) * (bz[2] + bz[3]);
            if (bot) Gw = Real(0);

            // ---- RK3 substep ----
            const int o = k * NCOL + c;
            if (gm_out != nullptr) { gm_out[0 * NC + o] = Gb; gm_out[1 * NC + o] = Gu; gm_out[2 * NC + o] = Gv; gm_out[3 * NC + o] = Gw; }   // last stage: never read
            nxt[GB + o] = bz[3] + dt * (gam * Gb + zet * g0[0]);
            nxt[GU + o] = uz[3] + dt * (gam * Gu + zet * g0[1]);
            nxt[GV + o] = vz[3] + dt * (gam * Gv + zet * g0[2]);
            nxt[GW + o] = bot ? Real(0) : wz[3] + dt * (gam * Gw + zet * g0[3]);

            // ---- slide up one level ----
            Fzb_lo = Fzb_hi; Wu_lo = Wu_hi; Wv_lo = Wv_hi; Ww_lo = Ww_hi;
            RBC_UNROLL
            for (int jj = 0; jj < 6; ++jj) { bz[jj] = bz[jj + 1]; uz[jj] = uz[jj + 1]; vz[jj] = vz[jj + 1]; wz[jj] = wz[jj + 1]; }
            const int kn = k + 4;
            bz[6] = (kn < NZ) ? src[SB + kn * LS + cs] : Real(0);
            uz[6] = (kn < NZ) ? src[SU + kn * LS + cs] : Real(0);
            vz[6] = (kn < NZ) ? src[SV + kn * LS + cs] : Real(0);
            wz[6] = (kn <= NZ) ? src[SW + kn * LS + cs] : Real(0);
            RBC_UNROLL
            for (int jj = 0; jj < 3; ++jj) { u1z[jj] = u1z[jj + 1]; v1z[jj] = v1z[jj + 1]; }
            u1z[3] = (k + 2 < NZ) ? src[SU + (k + 2) * LS + c_ip] : Real(0);
            v1z[3] = (k + 2 < NZ) ? src[SV + (k + 2) * LS + c_jp] : Real(0);
            RBC_UNROLL
            for (int d = 0; d < 7; ++d) { wx[d] = wxn[d]; wy[d] = wyn[d]; }
        };

        // The level loop is deliberately NOT fully unrolled: a wall body (run-time order selection) and the interior body
        // unrolled twice keep the march inside the instruction cache; the sliding windows cost ~45 register moves per level,
        // half of which the two-fold unrolling renames away (measured: rolled 59.1 k, x2 61.3 k, x4 58.1 k env-steps/s).
        RBC_NOUNROLL
        for (int k = 0; k < 2; ++k) level(BoolTag<true>{}, k);
        RBC_UNROLL2
        for (int k = 2; k <= NZ - 4; ++k) level(BoolTag<false>{}, k);
        RBC_NOUNROLL
        for (int k = NZ - 3; k < NZ; ++k) level(BoolTag<true>{}, k);
        nxt[GW + NZ * NCOL + c] = Real(0);               // top wall face
    }
}


// global variant: each thread marches its two columns straight from the state in global memory (through L1)
template <typename Real, bool SPLIT>
RBC_HD void phase_tendency3(int tid, const Consts3<Real>& C, const Real* cur, Real* nxt, const Real* RBC_RESTRICT phy,
                            const Real* RBC_RESTRICT Tb, const Real* gm_in, Real* gm_out, Real dt, Real gam, Real zet, bool use_gm)
{
    for (int q = 0; q < NCOL / NT; ++q) {
        const int c = tid + q * NT;
        tendency3_column<Real, SPLIT, false>(C, cur, nxt, phy, Tb, gm_in, gm_out, dt, gam, zet, use_gm, c % NX, c / NX, 0);
    }
}
// tiled variant: copy the rows half h touches into shared memory, then march from the tile
template <typename Real>
struct alignas(4 * sizeof(Real)) Vec4 {
    Real v[4];
};
template <typename Real>
RBC_HD void phase_load_tile3(int tid, const Real* RBC_RESTRICT cur, Real* RBC_RESTRICT tile, int h)
{
    // 16 bytes per access (rows are 32 values, 128-byte aligned in both layouts)
    constexpr int NV = NTILE / 4, VPL = TPL / 4, VROW = NX / 4;
    const Vec4<Real>* src = reinterpret_cast<const Vec4<Real>*>(cur);
    Vec4<Real>* dst = reinterpret_cast<Vec4<Real>*>(tile);
#if defined(__CUDA_ARCH__)
    if (sizeof(Real) == 4) {
        // asynchronous global->shared copies (LDGSTS): no register staging, all ~22 copies of a thread in flight at once
        const unsigned dbase = (unsigned)__cvta_generic_to_shared(dst);
        for (int q = tid; q < NV; q += NT) {
            const int p = q / VPL, rem = q % VPL, tr = rem / VROW, i4 = rem % VROW;
            const int f = p < 3 * NZ ? p / NZ : 3, lev = p - f * NZ;
            const int j = wrapn(h * THALF - 3 + tr, NY);
            const Vec4<Real>* g = src + (f * NC + lev * NCOL + j * NX) / 4 + i4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dbase + (unsigned)q * 16u), "l"(g) : "memory");
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
        return;
    }
#endif
    for (int q0 = tid; q0 < NV; q0 += 4 * NT) {
        Vec4<Real> val[4];
        RBC_UNROLL
        for (int a = 0; a < 4; ++a) {
            const int q = q0 + a * NT;
            if (q < NV) {
                const int p = q / VPL, rem = q % VPL, tr = rem / VROW, i4 = rem % VROW;
                const int f = p < 3 * NZ ? p / NZ : 3, lev = p - f * NZ;
                const int j = wrapn(h * THALF - 3 + tr, NY);
                val[a] = src[(f * NC + lev * NCOL + j * NX) / 4 + i4];
            }
        }
        RBC_UNROLL
        for (int a = 0; a < 4; ++a) {
            const int q = q0 + a * NT;
            if (q < NV) dst[q] = val[a];
        }
    }
}
template <typename Real, bool SPLIT>
RBC_HD void phase_tendency3_tile(int tid, const Consts3<Real>& C, const Real* tile, Real* nxt, const Real* RBC_RESTRICT phy,
                                 const Real* RBC_RESTRICT Tb, const Real* gm_in, Real* gm_out, Real dt, Real gam, Real zet, bool use_gm, int h)
{
    const int i = tid % NX, jl = tid / NX;                 // NT = 512 = 32 columns x 16 rows of the half
    tendency3_column<Real, SPLIT, true>(C, tile, nxt, phy, Tb, gm_in, gm_out, dt, gam, zet, use_gm, i, h * THALF + jl, jl);
}

// ------------------------------------------------------------------------------------------
// in-register FFTs (natural order in and out), SIGN = -1 forward / +1 inverse
//   N = 16: n = n1 + 4 n2, k = 4 k1 + k2;   N = 32: n = n1 + 4 n2 (n2 < 8), k = 8 k1 + k2
// ------------------------------------------------------------------------------------------
template <int SIGN, typename Real>
RBC_HD void fft16(cx<Real>* a)
{
    cx<Real> y[4][4];
    RBC_UNROLL
    for (int n1 = 0; n1 < 4; ++n1) {
        cx<Real> t0 = a[n1], t1 = a[n1 + 4], t2 = a[n1 + 8], t3 = a[n1 + 12];
        dft4<SIGN>(t0, t1, t2, t3);
        y[n1][0] = t0;
        y[n1][1] = cmul(t1, Real(tw16c(n1 * 1)), Real(SIGN * tw16s(n1 * 1)));
        y[n1][2] = cmul(t2, Real(tw16c(n1 * 2)), Real(SIGN * tw16s(n1 * 2)));
        y[n1][3] = cmul(t3, Real(tw16c(n1 * 3)), Real(SIGN * tw16s(n1 * 3)));
    }
    RBC_UNROLL
    for (int k2 = 0; k2 < 4; ++k2) {
        cx<Real> t0 = y[0][k2], t1 = y[1][k2], t2 = y[2][k2], t3 = y[3][k2];
        dft4<SIGN>(t0, t1, t2, t3);
        a[k2] = t0; a[4 + k2] = t1; a[8 + k2] = t2; a[12 + k2] = t3;
    }
}
template <int SIGN, typename Real>
RBC_HD void fft32(cx<Real>* a)
{
    cx<Real> y[4][8];
    RBC_UNROLL
    for (int n1 = 0; n1 < 4; ++n1) {
        cx<Real> t[8];
        RBC_UNROLL
        for (int n2 = 0; n2 < 8; ++n2) t[n2] = a[n1 + 4 * n2];
        dft8<SIGN>(t);
        RBC_UNROLL
        for (int k2 = 0; k2 < 8; ++k2) y[n1][k2] = cmul(t[k2], Real(tw32c(n1 * k2)), Real(SIGN * tw32s(n1 * k2)));
    }
    RBC_UNROLL
    for (int k2 = 0; k2 < 8; ++k2) {
        cx<Real> t0 = y[0][k2], t1 = y[1][k2], t2 = y[2][k2], t3 = y[3][k2];
        dft4<SIGN>(t0, t1, t2, t3);
        a[k2] = t0; a[8 + k2] = t1; a[16 + k2] = t2; a[24 + k2] = t3;
    }
}

// ------------------------------------------------------------------------------------------
// projection phases.  Scratch R[k][j][RX]: after the x pass, row (k, j) holds the half spectrum of the real row
// packed as 16 complex slots: slot 0 = (X[0], X[16]) (both real), slot m = X[m], m = 1..15.  After the y pass
// slot s >= 1 holds Y[ky = j] for kx = s; slot 0 is split into the two real-sequence spectra kx = 0 (rows 1..15,
// plus the real parts of rows 0 and 16) and kx = 16 (rows 31..17, plus the imaginary parts of rows 0 and 16).
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void phase_div3(int tid, const Consts3<Real>& C, const Real* p, Real* R)
{
    for (int q = 0; q < NCOL / NT; ++q) {
        const int c = tid + q * NT, i = c % NX, j = c / NX;
        const int cip = j * NX + wrapn(i + 1, NX), cjp = wrapn(j + 1, NY) * NX + i;
        Real wlo = p[GW + c];
        RBC_UNROLL
        for (int k = 0; k < NZ; ++k) {
            const Real whi = p[GW + (k + 1) * NCOL + c];
            R[k * RP + j * RX + i] = (p[GU + k * NCOL + cip] - p[GU + k * NCOL + c]) * C.idx +
                                     (p[GV + k * NCOL + cjp] - p[GV + k * NCOL + c]) * C.idy + (whi - wlo) * C.idz;
            wlo = whi;
        }
    }
}
template <typename Real>
RBC_HD void phase_fftx_fwd(int tid, Real* R)
{
    Real* z = R + (tid / NY) * RP + (tid % NY) * RX;              // one thread per row (k, j): NZ*NY = NT rows
    cx<Real> a[16];
    RBC_UNROLL
    for (int n = 0; n < 16; ++n) a[n] = ldc(z + 2 * n);
    fft16<-1>(a);
    // real-FFT split: X[m] = E + W32^m O, X[16-m] = conj(E - W32^m O)
    const Real x0 = a[0].re + a[0].im, x16 = a[0].re - a[0].im;
    a[0] = {x0, x16};
    RBC_UNROLL
    for (int m = 1; m < 8; ++m) rbc2d::untangle_pair(a[m], a[16 - m], Real(tw32c(m)), Real(tw32s(m)));
    a[8].im = -a[8].im;
    RBC_UNROLL
    for (int n = 0; n < 16; ++n) stc(z + 2 * n, a[n]);
}
template <typename Real>
RBC_HD void phase_fftx_inv(int tid, Real* R)
{
    Real* z = R + (tid / NY) * RP + (tid % NY) * RX;
    cx<Real> a[16];
    RBC_UNROLL
    for (int n = 0; n < 16; ++n) a[n] = ldc(z + 2 * n);
    const Real x0 = a[0].re, x16 = a[0].im;
    a[0] = {Real(0.5) * (x0 + x16), Real(0.5) * (x0 - x16)};
    RBC_UNROLL
    for (int m = 1; m < 8; ++m) rbc2d::tangle_pair(a[m], a[16 - m], Real(tw32c(m)), Real(tw32s(m)));
    a[8].im = -a[8].im;
    fft16<+1>(a);
    RBC_UNROLL
    for (int n = 0; n < 16; ++n) stc(z + 2 * n, a[n]);
}
// one thread per (level k, slot s): 32-point complex transform along y; items = NZ*16 = 256
template <typename Real>
RBC_HD void phase_ffty_fwd(int tid, Real* R)
{
    if (tid >= NZ * 16) return;
    const int k = tid / 16, s = tid % 16;
    Real* z = R + k * RP + 2 * s;
    cx<Real> a[32];
    RBC_UNROLL
    for (int j = 0; j < NY; ++j) a[j] = ldc(z + j * RX);
    fft32<-1>(a);
    if (s == 0) {
        // Z = A + iB with A, B the spectra of the real sequences X[0](j), X[16](j):
        //   A[ky] = (Z[ky] + conj Z[32-ky])/2 -> row ky;   B[ky] = -i (Z[ky] - conj Z[32-ky])/2 -> row 32-ky
        RBC_UNROLL
        for (int ky = 1; ky < 16; ++ky) {
            const cx<Real> zp = a[ky], zm = a[32 - ky];
            a[ky] = {Real(0.5) * (zp.re + zm.re), Real(0.5) * (zp.im - zm.im)};
            a[32 - ky] = {Real(0.5) * (zp.im + zm.im), Real(-0.5) * (zp.re - zm.re)};
        }
        // rows 0 and 16: (A, B) are both real and already sit in (re, im)
    }
    RBC_UNROLL
    for (int j = 0; j < NY; ++j) stc(z + j * RX, a[j]);
}
template <typename Real>
RBC_HD void phase_ffty_inv(int tid, Real* R)
{
    if (tid >= NZ * 16) return;
    const int k = tid / 16, s = tid % 16;
    Real* z = R + k * RP + 2 * s;
    cx<Real> a[32];
    RBC_UNROLL
    for (int j = 0; j < NY; ++j) a[j] = ldc(z + j * RX);
    if (s == 0) {
        // Z[ky] = A[ky] + i B[ky],  Z[32-ky] = conj A[ky] + i conj B[ky]
        RBC_UNROLL
        for (int ky = 1; ky < 16; ++ky) {
            const cx<Real> A = a[ky], B = a[32 - ky];
            a[ky] = {A.re - B.im, A.im + B.re};
            a[32 - ky] = {A.re + B.im, -A.im + B.re};
        }
    }
    fft32<+1>(a);
    RBC_UNROLL
    for (int j = 0; j < NY; ++j) stc(z + j * RX, a[j]);
}
// 1024 independent real tridiagonal systems of length NZ per environment, two per thread, in place
template <typename Real>
RBC_HD void phase_thomas3(int tid, Real* RBC_RESTRICT R, const Real* tinv, Real scale)
{
    for (int q = 0; q < NCOL / NT; ++q) {
        const int wd = tid + q * NT, j = wd / NX, c = wd % NX;
        Real* r = R + j * RX + c;
        const Real* tv = tinv + wd;
        Real iv[NZ], d[NZ];
        RBC_UNROLL
        for (int k = 0; k < NZ; ++k) { iv[k] = tv[k * NCOL]; d[k] = r[k * RP] * scale * iv[k]; }
        RBC_UNROLL
        for (int k = 1; k < NZ; ++k) d[k] = d[k] - d[k - 1] * iv[k];
        RBC_UNROLL
        for (int k = NZ - 2; k >= 0; --k) d[k] = d[k] - iv[k] * d[k + 1];
        RBC_UNROLL
        for (int k = 0; k < NZ; ++k) r[k * RP] = d[k];
    }
}
template <typename Real>
RBC_HD void phase_correct3(int tid, const Consts3<Real>& C, Real* p, const Real* RBC_RESTRICT R)
{
    for (int q = 0; q < NCOL / NT; ++q) {
        const int c = tid + q * NT, i = c % NX, j = c / NX;
        const int rim = j * RX + wrapn(i - 1, NX), rjm = wrapn(j - 1, NY) * RX + i, r0 = j * RX + i;
        Real below = Real(0);
        RBC_UNROLL
        for (int k = 0; k < NZ; ++k) {
            const Real ph = R[k * RP + r0];
            p[GU + k * NCOL + c] -= (ph - R[k * RP + rim]) * C.idx;
            p[GV + k * NCOL + c] -= (ph - R[k * RP + rjm]) * C.idy;
            if (k >= 1) p[GW + k * NCOL + c] -= (ph - below) * C.idz;
            below = ph;
        }
    }
}
template <typename Real>
RBC_HD void project3(const Consts3<Real>& C, const Ctx3<Real>& X, Real* p)
{
    RBC3_PHASE(phase_div3(tid, C, p, X.R);)
    RBC3_PHASE(phase_fftx_fwd(tid, X.R);)
    RBC3_PHASE(phase_ffty_fwd(tid, X.R);)
    RBC3_PHASE(phase_thomas3(tid, X.R, X.tinv, X.thomas_scale);)
    RBC3_PHASE(phase_ffty_inv(tid, X.R);)
    RBC3_PHASE(phase_fftx_inv(tid, X.R);)
    RBC3_PHASE(phase_correct3(tid, C, p, X.R);)
}

// ------------------------------------------------------------------------------------------
// epilogue of a 3D action step: write back, observation (= get_state), Nusselt number, NaN flag, clock and the fused
// vector-env bookkeeping (see rbc2d::env_epilogue).  Returns whether an in-launch re-initialisation is due.
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD bool env_epilogue3(const Consts3<Real>& C, const EnvIO3<Real>& io, const Ctx3<Real>& X, int env, const Real* cur, const RunFlags3& F,
                          bool second_pass)
{
    const rbc2d::VecIO& V = io.vec;
    const int pend = (!second_pass && V.mode == 1 && V.bank != nullptr && F.advance_clock) ? V.pending[env] : 0;
    const double t_old = io.t[env];
    Real* st = io.state + (size_t)env * NSTATE;
    RBC3_PHASE(
        double acc = 0, bad = 0;
        for (int q = 0; q < NCOL / NT; ++q) {
            const int c = tid + q * NT;
            for (int k = 0; k < NZ; ++k) {
                const double b = (double)cur[GB + k * NCOL + c], u = (double)cur[GU + k * NCOL + c];
                const double v = (double)cur[GV + k * NCOL + c], w = (double)cur[GW + k * NCOL + c];
                if (b != b || u != u || v != v || w != w) bad += 1;
                // get_nusselt (rbc_sim3D_api.jl:134-159): conduction profile on a UNIT height, z = (k + 1/2)/NZ
                const double zc = (k + 0.5) / NZ, Tc = (1.0 - zc) * C.delta_b_d + C.b_top_d;
                acc += (b - Tc) * w;
            }
        }
        X.red[tid] = acc;
        X.red[NT + tid] = bad;
    )
#if defined(__CUDA_ARCH__)
    __shared__ double fin3[2];
#else
    double fin3[2];
#endif
    RBC3_PHASE(if (tid < 2) fin3[tid] = sum_serial(X.red + tid * NT, NT);)
    const bool vec = V.mode >= 0 && F.advance_clock;
    const bool bad = fin3[1] > 0;
    const bool stepping = F.advance_clock && !pend && !second_pass;
    const bool trunc_now = stepping && (t_old + C.dt_action >= C.episode_length);
    const bool bad_reset = vec && stepping && V.nan_reset && bad;
    const bool do_reset = vec && stepping && V.bank != nullptr && ((V.mode == 2 && trunc_now) || bad_reset);
    float* const ob_base = (do_reset && V.final_obs != nullptr) ? V.final_obs : io.obs;
    RBC3_PHASE(
        if (cur != st && !do_reset) { for (int q = tid; q < NSTATE; q += NT) st[q] = cur[q]; }
        if (ob_base != nullptr && !(do_reset && V.final_obs == nullptr)) {
            float* ob = ob_base + (size_t)env * 4 * NC;
            for (int q = tid; q < 4 * NC; q += NT) ob[q] = (float)cur[q];        // b,u,v and w faces 0..NZ-1 are contiguous
        }
        if (tid == 0) {
            const double nu = 1.0 + (fin3[0] / (double)NC) / C.kappa_d;
            if (second_pass) {
                io.nusselt[env] = nu;
                io.t[env] = 0.0;
                io.step_count[env] = 1;
                V.episode[env] += 1;
                V.ep_return[env] = 0.0;
                if (V.pending != nullptr) V.pending[env] = 0;
            } else if (pend) {
                io.nusselt[env] = nu;
                io.reward[env] = 0.0f;
                io.nan_flag[env] = bad ? 1 : 0;
                io.truncated[env] = 0;
                io.t[env] = 0.0;
                io.step_count[env] = 1;
                V.episode[env] += 1;
                V.ep_return[env] = 0.0;
                V.pending[env] = 0;
            } else {
                double rew = -nu;
                if (bad_reset) rew = 0.0;
                if (bad && vec && V.nan_count != nullptr) RBC_COUNT_ONE(V.nan_count);
                io.reward[env] = (float)rew;
                io.nan_flag[env] = bad ? 1 : 0;
                const double ret = vec ? V.ep_return[env] + (double)(float)rew : 0.0;
                if (do_reset) {
                    if (V.final_nu_a != nullptr) V.final_nu_a[env] = nu;
                    if (V.final_return != nullptr) V.final_return[env] = ret;
                    io.truncated[env] = 1;
                } else {
                    io.nusselt[env] = nu;
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
    )
    return do_reset;
}

// ------------------------------------------------------------------------------------------
// one action step of one 3D environment
// ------------------------------------------------------------------------------------------
template <typename Real, bool SPLIT, bool TILED = false>
RBC_HD void env_action_step3(const Consts3<Real>& C, const EnvIO3<Real>& io, const Ctx3<Real>& X, int env, const RunFlags3& F)
{
    static_assert(!(TILED && SPLIT), "the tile aliases the scratch that holds pHY' in split mode");
    const Real gam[3] = {Real(8.0 / 15.0), Real(5.0 / 12.0), Real(3.0 / 4.0)};
    const Real zet[3] = {Real(0), Real(-17.0 / 60.0), Real(-5.0 / 12.0)};
    Real* st = io.state + (size_t)env * NSTATE;
    const rbc2d::VecIO& V = io.vec;
    // fused vector-env semantics: a pending environment (next_step mode) is only re-initialised from the checkpoint bank
    const int pend = (V.mode == 1 && V.bank != nullptr && F.advance_clock) ? V.pending[env] : 0;
    const int nsub = pend ? 0 : F.nsub;
    if (pend) {
        const double* src = V.bank + (size_t)rbc2d::checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * NSTATE;
        RBC3_PHASE(for (int q = tid; q < NSTATE; q += NT) st[q] = (Real)src[q];)
    } else {
        RBC3_PHASE(
            for (int q = 0; q < NCOL / NT; ++q) {
                const int c = tid + q * NT;
                X.Tb[c] = (Real)heater_T3(C, io.actions + (size_t)env * C.heaters * C.heaters, c % NX, c / NX);
            }
        )
    }
    if (F.project_first && !pend) project3(C, X, st);
    Real* cur = st;                        // stage 1 reads the environment's own array, then A <-> B ping-pong
    Real* nxt = X.bufA;
    int stage_no = 0;
    for (int sub = 0; sub < nsub; ++sub) {
        const Real dt = (sub == nsub - 1) ? C.dt_last : C.dt_full;
        for (int stage = 0; stage < 3; ++stage, ++stage_no) {
            if (SPLIT) { RBC3_PHASE(phase_phy3(tid, C, cur + GB, X.R);) }
            // one tendency slab, updated in place: every (column, level) entry is read (previous stage) and then
            // rewritten by the same thread, which halves the per-CTA footprint of the slabs in L2
            const Real* gin = X.gm;
            Real* gout = X.gm;
            if (TILED) {
                for (int h = 0; h < 2; ++h) {
                    RBC3_PHASE(phase_load_tile3(tid, cur, X.tile, h);)
                    RBC3_PHASE((phase_tendency3_tile<Real, false>(tid, C, X.tile, nxt, X.R, X.Tb, gin, stage < 2 ? gout : nullptr, dt, gam[stage], zet[stage], stage > 0, h));)
                }
            } else {
                RBC3_PHASE((phase_tendency3<Real, SPLIT>(tid, C, cur, nxt, X.R, X.Tb, gin, stage < 2 ? gout : nullptr, dt, gam[stage], zet[stage], stage > 0));)
            }
            project3(C, X, nxt);
            cur = nxt;
            nxt = (cur == X.bufA) ? X.bufB : X.bufA;
        }
    }
    // one inlined copy of the epilogue, run a second time when the environment is re-initialised inside this launch
    RBC_NOUNROLL
    for (int pass = 0; pass < 2; ++pass) {
        if (!env_epilogue3(C, io, X, env, cur, F, pass == 1)) break;
        const double* src = V.bank + (size_t)rbc2d::checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * NSTATE;
        RBC3_PHASE(for (int q = tid; q < NSTATE; q += NT) st[q] = (Real)src[q];)
        cur = st;
    }
}

// ------------------------------------------------------------------------------------------
// host-side construction (fp64)
// ------------------------------------------------------------------------------------------
struct HostConfig3 {
    double ra, pr, lx, ly, lz, b_top, delta_b, heater_limit, heater_duration, dt_solver, episode_length;
    int heaters;
};
template <typename Real>
inline Consts3<Real> make_consts3(const HostConfig3& h)
{
    Consts3<Real> C;
    const double dx = h.lx / NX, dy = h.ly / NY, dz = h.lz / NZ;
    const double nu = sqrt(h.pr / h.ra), kappa = 1.0 / sqrt(h.pr * h.ra);       // rbc_sim3D_api.jl:37-38
    const double t_ff = h.lz * h.lz;                                             // rbc_sim3D_api.jl:43
    C.idx = (Real)(1 / dx); C.idy = (Real)(1 / dy); C.idz = (Real)(1 / dz);
    C.idx2 = (Real)(1 / (dx * dx)); C.idy2 = (Real)(1 / (dy * dy)); C.idz2 = (Real)(1 / (dz * dz));
    C.nu = (Real)nu; C.kappa = (Real)kappa; C.b_top = (Real)h.b_top;
    double last;
    C.nsub = rbc2d::substep_schedule(h.heater_duration * t_ff, h.dt_solver * t_ff, &last);
    C.dt_full = (Real)(h.dt_solver * t_ff); C.dt_last = (Real)last;
    C.heaters = h.heaters; C.heater_limit = h.heater_limit; C.b_hot = h.b_top + h.delta_b;
    C.kappa_d = kappa; C.delta_b_d = h.delta_b; C.b_top_d = h.b_top;
    C.dt_action = h.heater_duration * t_ff; C.episode_length = h.episode_length;
    return C;
}
// Thomas pivots per spectral word (k, j, c) of the scratch after the y pass (see the layout note above)
inline void build_tables3_host(double lx, double ly, double lz, double* tinv /*NZ*NCOL*/)
{
    const double PI = 3.14159265358979323846;
    const double dx = lx / NX, dy = ly / NY, dz = lz / NZ;
    for (int j = 0; j < NY; ++j)
        for (int c = 0; c < NX; ++c) {
            const int s = c / 2;
            int kx, ky;
            if (s != 0) { kx = s; ky = j; }
            else if (j == 0 || j == NY / 2) { kx = (c == 0) ? 0 : NX / 2; ky = j; }
            else if (j < NY / 2) { kx = 0; ky = j; }
            else { kx = NX / 2; ky = NY - j; }
            const double sx = 2 * sin(PI * kx / NX) / dx, sy = 2 * sin(PI * ky / NY) / dy, lam = (sx * sx + sy * sy) * dz * dz;
            double prev = 0.0;
            for (int k = 0; k < NZ; ++k) {
                double dg = -(2.0 + lam);
                if (k == 0 || k == NZ - 1) dg += 1.0;
                if (kx == 0 && ky == 0 && k == 0) dg -= 1.0;     // pin the null space of the mean mode
                const double iv = 1.0 / (dg - prev);
                tinv[(size_t)k * NCOL + j * NX + c] = iv;
                prev = iv;
            }
        }
}

}  // namespace rbc3d
