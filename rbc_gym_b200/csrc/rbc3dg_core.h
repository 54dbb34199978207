// rbc3dg_core.h — the 3D Rayleigh-Benard action step for ANY grid (state_shape is a free keyword argument of the reference
// environment, src/rbc_gym/envs/rbc3D.py:43-60; its only Julia-produced 3D data, experiments/flowstats/flowstats_ra.py:27-36,
// was run at 64 x 64 x 32): the *stage-streaming* formulation of SURVEY.md 8d.
//
// Replaces the same reference functions as rbc3d_core.h (step_simulation / run!, preprocess_action, bottom_T, get_state,
// get_nusselt: src/rbc_gym/sim/rbc_sim3D_api.jl:77-159, rbc_sim3D.jl:99-166) for grids other than the registered
// 32 x 32 x 16, whose dedicated one-CTA-per-environment kernel stays the fast path.
//
// A 64 x 64 x 32 environment is 2.1 MB in fp32 — ten times the shared memory of an SM — so here the state and the
// previous-stage tendencies stream through HBM/L2 once per RK3 stage, exactly the traffic the roofline of SURVEY 8d
// counts (10 S values per RK3 step), and every kernel runs over the whole batch:
//   tendency   one thread per cell: the four tendencies from the current state (5th-order upwind / 4th-order centred
//              fluxes with the 5-3-1 / 4-2 wall rules, explicit diffusion, buoyancy in G_w), the RK3 substep
//              U* = U + dt (gamma G + zeta G-), G stored in place of G-;
//   div + FFT  one CTA per (environment, level): divergence of U* into a shared-memory plane, radix-2 FFT in x and y
//              (decimation in frequency: natural order in, bit-reversed order out, no permutation pass);
//   Thomas     one thread per horizontal mode: the tridiagonal solve in z with host-built pivots (stored in the same
//              bit-reversed order);
//   inverse    one CTA per (environment, level): decimation-in-time inverse (bit-reversed in, natural out) -> phi;
//   correct    one thread per cell: U = U* - grad phi.
// The Rayleigh number may differ per environment (nu, kappa arrays): the flowstats protocol runs its 14 Rayleigh
// numbers as one batch.
//
// Plain C++ shared between nvcc and g++ (tests/emu) like rbc2d_core.h: every piece of per-thread work is a function of a
// work-item index, which the kernels map to threads and the emulator to loops.
#pragma once
#include "rbc2d_core.h"

namespace rbc3dg {

using rbc2d::centred_ord;
using rbc2d::cx;
using rbc2d::upwind_ord;

struct Dims {
    int nx, ny, nz;
    int ncol, nc, nw, nstate;      // nx*ny, ncol*nz, ncol*(nz+1), 3*nc + nw
    int gb, gu, gv, gw;            // field offsets in the checkpoint layout (b,u,v [nz][ny][nx], w [nz+1][ny][nx])
    int lx2, ly2;                  // log2(nx), log2(ny)
};
inline Dims make_dims(int nx, int ny, int nz)
{
    Dims D;
    D.nx = nx; D.ny = ny; D.nz = nz;
    D.ncol = nx * ny; D.nc = D.ncol * nz; D.nw = D.ncol * (nz + 1); D.nstate = 3 * D.nc + D.nw;
    D.gb = 0; D.gu = D.nc; D.gv = 2 * D.nc; D.gw = 3 * D.nc;
    D.lx2 = 0; while ((1 << D.lx2) < nx) ++D.lx2;
    D.ly2 = 0; while ((1 << D.ly2) < ny) ++D.ly2;
    return D;
}
inline bool dims_supported(int nx, int ny, int nz)
{
    auto pow2 = [](int n) { return n >= 8 && n <= 256 && (n & (n - 1)) == 0; };
    return pow2(nx) && pow2(ny) && nz >= 6 && nz <= 256;
}

template <typename Real>
struct ConstsG {
    Real idx, idy, idz, idx2, idy2, idz2;
    Real b_top;
    Real nu, kappa;              // used when no per-environment arrays are given
    Real dt_full, dt_last;       // simulation-time substeps: dt_solver*t_ff, last one clipped (rbc_sim3D_api.jl:65)
    int nsub;
    int heaters;
    double heater_limit, b_hot;
    double kappa_d, delta_b_d, b_top_d;
    double dt_action, episode_length;
    int variant;                 // 0 = the scheme of SURVEY 8a.  Non-zero: experimental wall-order variants for the parity study of DESIGN.md section 2 (RBC_B200_G3_VARIANT)
};

// wall-order rules: centres -> z-face kf, z-faces -> centre kc
RBC_HD int o_up_face(int kf, int nz) { return (kf >= 3 && kf <= nz - 3) ? 5 : ((kf == 2 || kf == nz - 2) ? 3 : 1); }
RBC_HD int o_ce_face(int kf, int nz) { return (kf >= 2 && kf <= nz - 2) ? 4 : 2; }
RBC_HD int o_up_cen(int kc, int nz) { return (kc >= 2 && kc <= nz - 3) ? 5 : ((kc == 1 || kc == nz - 2) ? 3 : 1); }
RBC_HD int o_ce_cen(int kc, int nz) { return (kc >= 1 && kc <= nz - 2) ? 4 : 2; }

// ------------------------------------------------------------------------------------------
// heater patches: preprocess_action + bottom_T (rbc_sim3D.jl:111-141), fp64 like the reference
// ------------------------------------------------------------------------------------------
RBC_HD double heater_patch_T(int heaters, double heater_limit, double b_hot, const float* action, int i, int j, int nx, int ny)
{
    const int h = heaters, n = h * h;
    double mean = 0.0;
    for (int q = 0; q < n; ++q) mean += (double)action[q];
    mean /= n;
    double K = 1.0;
    for (int q = 0; q < n; ++q) {
        const double d = fabs((double)action[q] - mean);
        if (d > K) K = d;
    }
    int pi = (int)floor((i + 0.5) / nx * h), pj = (int)floor((j + 0.5) / ny * h);
    pi = pi < 0 ? 0 : (pi > h - 1 ? h - 1 : pi);
    pj = pj < 0 ? 0 : (pj > h - 1 ? h - 1 : pj);
    return b_hot + (((double)action[pi * h + pj] - mean) / K) * heater_limit;       // action[i][j]: patch i along x, j along y
}

// ------------------------------------------------------------------------------------------
// tendency + RK3 substep of one cell (buoyancy added to G_w, which after the projection is identical to
// Oceananigans' hydrostatic-pressure split — the 3D state exposes no pressure channel)
//   S: current state, P: predicted state (out), G: tendency slab [4][nc] read (previous stage) and rewritten in place
//
// Every value the 24 face fluxes of the cell need is loaded ONCE into register windows: seven values along x, y and z
// through the cell for each of b, u, v, w (periodic wrap in x and y; the z index is clamped, and a clamped slot is only ever
// read by a stencil whose reduced wall order ignores it), plus the six 4-value rows the mixed momentum fluxes take their
// advecting velocity from on the far faces.  ~100 loads and ~700 instructions per cell; the first version went through a
// wrapping accessor for every stencil point (~7 400 instructions per cell, 1.67 ms per 64-environment stage at 64 x 64 x 32).
// Window index d <-> offset d - 3 from the cell; a face "m-1 | m" along a direction uses slots m .. m + 5 of that window.
// ------------------------------------------------------------------------------------------
// Parity-study variants of the wall rule.  The scheme of record reduces the order of BOTH biased stencils of a
// face together (o_up_face / o_up_cen).  Variant bit 0 (centres -> z-face) and bit 1 (z-faces -> centre) let each biased stencil
// keep the highest order that fits between the walls on its own: the left-biased one (cells kf-3 .. kf+1) and the right-biased
// one (kf-2 .. kf+2) then switch order at different faces.
template <typename Real>
RBC_HD Real upwind_z_face(Real vel, const Real* win, int kf, int nz, int variant)
{
    if (!(variant & 1)) return upwind_ord(vel, win, o_up_face(kf, nz));
    const int oL = (kf >= 3 && kf <= nz - 2) ? 5 : ((kf >= 2) ? 3 : 1), oR = (kf >= 2 && kf <= nz - 3) ? 5 : ((kf <= nz - 2) ? 3 : 1);
    return upwind_ord(vel, win, vel > Real(0) ? oL : oR);
}
template <typename Real>
RBC_HD Real upwind_z_cen(Real vel, const Real* win, int kc, int nz, int variant)
{
    if (!(variant & 2)) return upwind_ord(vel, win, o_up_cen(kc, nz));
    const int oL = (kc >= 2 && kc <= nz - 2) ? 5 : ((kc >= 1) ? 3 : 1), oR = (kc >= 1 && kc <= nz - 3) ? 5 : ((kc <= nz - 2) ? 3 : 1);
    return upwind_ord(vel, win, vel > Real(0) ? oL : oR);
}

// the register windows of one cell (scalarised by the compiler once everything is inlined)
template <typename Real>
struct Windows {
    Real bx[7], by[7], bz[7], ux[7], uy[7], uz[7], vx[7], vy[7], vz[7], wx[7], wy[7], wz[7];
    // advecting-velocity rows of the far faces: slot q <-> offset q - 2 along the interpolation direction
    Real v_jp[4];       // v(i-2..i+1, j+1, k)    -> u-flux through the y-face j+1
    Real u_ip[4];       // u(i+1, j-2..j+1, k)    -> v-flux through the x-face i+1
    Real w_kp_x[4];     // w(i-2..i+1, j, k+1)    -> u-flux through the z-face k+1
    Real w_kp_y[4];     // w(i, j-2..j+1, k+1)    -> v-flux through the z-face k+1
    Real u_ip_z[4];     // u(i+1, j, k-2..k+1)    -> w-flux through the x-face i+1
    Real v_jp_z[4];     // v(i, j+1, k-2..k+1)    -> w-flux through the y-face j+1
};
template <typename Real>
struct Tend {
    Real b, u, v, w;
};
// the four tendencies of the cell at level k from its windows (tb = heater temperature below the column, read at k = 0 only)
// level k is "interior" when every z-stencil of the cell has its full order and no wall ghost is involved; the march of the tiled
// kernel runs those levels through a branch-free instantiation
RBC_HD bool interior_level(int k, int nz) { return k >= 3 && k <= nz - 4; }
// CARRY: a thread that marches a column upwards hands the four fluxes through the upper face of level k - 1 to level k, where
// they are the fluxes through its lower face — the same function of the same values, so nothing changes in the result, but four
// of the 24 reconstructions of a cell go away.  `zf` is that hand-over (in: lower-face fluxes of this level, out: upper-face
// fluxes); without CARRY it is ignored.  (The w-flux is only available from level 1 on: level 0 does not evaluate it.)
template <typename Real, bool INTERIOR, bool CARRY = false>
RBC_HD Tend<Real> tendency_from_windows_t(const ConstsG<Real>& C, Real nu, Real kappa, int nz, int k, const Windows<Real>& W, Real tb, Tend<Real>* zf = nullptr);
template <typename Real>
RBC_HD Tend<Real> tendency_from_windows(const ConstsG<Real>& C, Real nu, Real kappa, int nz, int k, const Windows<Real>& W, Real tb)
{
    return tendency_from_windows_t<Real, false, false>(C, nu, kappa, nz, k, W, tb, nullptr);
}
template <typename Real, bool INTERIOR, bool CARRY>
RBC_HD Tend<Real> tendency_from_windows_t(const ConstsG<Real>& C, Real nu, Real kappa, int nz, int k, const Windows<Real>& W, Real tb, Tend<Real>* zf)
{
    const Real *bx = W.bx, *by = W.by, *bz = W.bz, *ux = W.ux, *uy = W.uy, *uz = W.uz, *vx = W.vx, *vy = W.vy, *vz = W.vz, *wx = W.wx, *wy = W.wy, *wz = W.wz;
    if (INTERIOR) {
        // the same expressions as the general body below with every order decision resolved: 5th / 4th order everywhere
        const Real b0 = bx[3], u0 = ux[3], v0 = vx[3], w0 = wx[3];
        Tend<Real> g;
        {
            const Real Fx0 = upwind_ord(u0, bx, 5), Fx1 = upwind_ord(ux[4], bx + 1, 5);
            const Real Fy0 = upwind_ord(v0, by, 5), Fy1 = upwind_ord(vy[4], by + 1, 5);
            const Real Fz0 = CARRY ? zf->b : upwind_ord(w0, bz, 5), Fz1 = upwind_ord(wz[4], bz + 1, 5);
            if (CARRY) zf->b = Fz1;
            const Real lap = (bx[4] - Real(2) * b0 + bx[2]) * C.idx2 + (by[4] - Real(2) * b0 + by[2]) * C.idy2 + (bz[4] - Real(2) * b0 + bz[2]) * C.idz2;
            g.b = -((Fx1 - Fx0) * C.idx + (Fy1 - Fy0) * C.idy + (Fz1 - Fz0) * C.idz) + kappa * lap;
        }
        {
            const Real F0 = upwind_ord(centred_ord(ux[1], ux[2], ux[3], ux[4], 4), ux, 5);
            const Real F1 = upwind_ord(centred_ord(ux[2], ux[3], ux[4], ux[5], 4), ux + 1, 5);
            const Real G0 = upwind_ord(centred_ord(vx[1], vx[2], vx[3], vx[4], 4), uy, 5);
            const Real G1 = upwind_ord(centred_ord(W.v_jp[0], W.v_jp[1], W.v_jp[2], W.v_jp[3], 4), uy + 1, 5);
            const Real H0 = CARRY ? zf->u : upwind_ord(centred_ord(wx[1], wx[2], wx[3], wx[4], 4), uz, 5);
            const Real H1 = upwind_ord(centred_ord(W.w_kp_x[0], W.w_kp_x[1], W.w_kp_x[2], W.w_kp_x[3], 4), uz + 1, 5);
            if (CARRY) zf->u = H1;
            const Real lap = (ux[4] - Real(2) * u0 + ux[2]) * C.idx2 + (uy[4] - Real(2) * u0 + uy[2]) * C.idy2 + (uz[4] - Real(2) * u0 + uz[2]) * C.idz2;
            g.u = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap;
        }
        {
            const Real F0 = upwind_ord(centred_ord(uy[1], uy[2], uy[3], uy[4], 4), vx, 5);
            const Real F1 = upwind_ord(centred_ord(W.u_ip[0], W.u_ip[1], W.u_ip[2], W.u_ip[3], 4), vx + 1, 5);
            const Real G0 = upwind_ord(centred_ord(vy[1], vy[2], vy[3], vy[4], 4), vy, 5);
            const Real G1 = upwind_ord(centred_ord(vy[2], vy[3], vy[4], vy[5], 4), vy + 1, 5);
            const Real H0 = CARRY ? zf->v : upwind_ord(centred_ord(wy[1], wy[2], wy[3], wy[4], 4), vz, 5);
            const Real H1 = upwind_ord(centred_ord(W.w_kp_y[0], W.w_kp_y[1], W.w_kp_y[2], W.w_kp_y[3], 4), vz + 1, 5);
            if (CARRY) zf->v = H1;
            const Real lap = (vx[4] - Real(2) * v0 + vx[2]) * C.idx2 + (vy[4] - Real(2) * v0 + vy[2]) * C.idy2 + (vz[4] - Real(2) * v0 + vz[2]) * C.idz2;
            g.v = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap;
        }
        {
            const Real F0 = upwind_ord(centred_ord(uz[1], uz[2], uz[3], uz[4], 4), wx, 5);
            const Real F1 = upwind_ord(centred_ord(W.u_ip_z[0], W.u_ip_z[1], W.u_ip_z[2], W.u_ip_z[3], 4), wx + 1, 5);
            const Real G0 = upwind_ord(centred_ord(vz[1], vz[2], vz[3], vz[4], 4), wy, 5);
            const Real G1 = upwind_ord(centred_ord(W.v_jp_z[0], W.v_jp_z[1], W.v_jp_z[2], W.v_jp_z[3], 4), wy + 1, 5);
            const Real H0 = CARRY ? zf->w : upwind_ord(centred_ord(wz[1], wz[2], wz[3], wz[4], 4), wz, 5);      // interior: k >= 3
            const Real H1 = upwind_ord(centred_ord(wz[2], wz[3], wz[4], wz[5], 4), wz + 1, 5);
            if (CARRY) zf->w = H1;
            const Real lap = (wx[4] - Real(2) * w0 + wx[2]) * C.idx2 + (wy[4] - Real(2) * w0 + wy[2]) * C.idy2 + (wz[4] - Real(2) * w0 + wz[2]) * C.idz2;
            g.w = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap + Real(0.5) * (bz[2] + b0);
        }
        return g;
    }
    const int of0 = (k >= 1) ? o_up_face(k, nz) : 0, of1 = (k + 1 <= nz - 1) ? o_up_face(k + 1, nz) : 0;
    const Real b0 = bx[3], u0 = ux[3], v0 = vx[3], w0 = wx[3];
    Tend<Real> g;
    g.w = Real(0);
    {   // tracer: the advecting velocity is the face value itself
        const Real Fx0 = upwind_ord(u0, bx, 5), Fx1 = upwind_ord(ux[4], bx + 1, 5);
        const Real Fy0 = upwind_ord(v0, by, 5), Fy1 = upwind_ord(vy[4], by + 1, 5);
        const Real Fz0 = CARRY ? zf->b : (of0 ? upwind_z_face(w0, bz, k, nz, C.variant) : Real(0));
        const Real Fz1 = of1 ? upwind_z_face(wz[4], bz + 1, k + 1, nz, C.variant) : Real(0);
        if (CARRY) zf->b = Fz1;
        const Real bdn = (k == 0) ? Real(2) * tb - b0 : bz[2];
        const Real bup = (k == nz - 1) ? Real(2) * C.b_top - b0 : bz[4];
        const Real lap = (bx[4] - Real(2) * b0 + bx[2]) * C.idx2 + (by[4] - Real(2) * b0 + by[2]) * C.idy2 + (bup - Real(2) * b0 + bdn) * C.idz2;
        g.b = -((Fx1 - Fx0) * C.idx + (Fy1 - Fy0) * C.idy + (Fz1 - Fz0) * C.idz) + kappa * lap;
    }
    {   // u at (x-face i, j, k)
        const Real F0 = upwind_ord(centred_ord(ux[1], ux[2], ux[3], ux[4], 4), ux, 5);
        const Real F1 = upwind_ord(centred_ord(ux[2], ux[3], ux[4], ux[5], 4), ux + 1, 5);
        const Real G0 = upwind_ord(centred_ord(vx[1], vx[2], vx[3], vx[4], 4), uy, 5);
        const Real G1 = upwind_ord(centred_ord(W.v_jp[0], W.v_jp[1], W.v_jp[2], W.v_jp[3], 4), uy + 1, 5);
        const Real H0 = CARRY ? zf->u : (of0 ? upwind_z_face(centred_ord(wx[1], wx[2], wx[3], wx[4], 4), uz, k, nz, C.variant) : Real(0));
        const Real H1 = of1 ? upwind_z_face(centred_ord(W.w_kp_x[0], W.w_kp_x[1], W.w_kp_x[2], W.w_kp_x[3], 4), uz + 1, k + 1, nz, C.variant) : Real(0);
        if (CARRY) zf->u = H1;
        const Real dn = (k == 0) ? -u0 : uz[2], up = (k == nz - 1) ? -u0 : uz[4];
        const Real lap = (ux[4] - Real(2) * u0 + ux[2]) * C.idx2 + (uy[4] - Real(2) * u0 + uy[2]) * C.idy2 + (up - Real(2) * u0 + dn) * C.idz2;
        g.u = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap;
    }
    {   // v at (i, y-face j, k)
        const Real F0 = upwind_ord(centred_ord(uy[1], uy[2], uy[3], uy[4], 4), vx, 5);
        const Real F1 = upwind_ord(centred_ord(W.u_ip[0], W.u_ip[1], W.u_ip[2], W.u_ip[3], 4), vx + 1, 5);
        const Real G0 = upwind_ord(centred_ord(vy[1], vy[2], vy[3], vy[4], 4), vy, 5);
        const Real G1 = upwind_ord(centred_ord(vy[2], vy[3], vy[4], vy[5], 4), vy + 1, 5);
        const Real H0 = CARRY ? zf->v : (of0 ? upwind_z_face(centred_ord(wy[1], wy[2], wy[3], wy[4], 4), vz, k, nz, C.variant) : Real(0));
        const Real H1 = of1 ? upwind_z_face(centred_ord(W.w_kp_y[0], W.w_kp_y[1], W.w_kp_y[2], W.w_kp_y[3], 4), vz + 1, k + 1, nz, C.variant) : Real(0);
        if (CARRY) zf->v = H1;
        const Real dn = (k == 0) ? -v0 : vz[2], up = (k == nz - 1) ? -v0 : vz[4];
        const Real lap = (vx[4] - Real(2) * v0 + vx[2]) * C.idx2 + (vy[4] - Real(2) * v0 + vy[2]) * C.idy2 + (up - Real(2) * v0 + dn) * C.idz2;
        g.v = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap;
    }
    if (k >= 1) {   // w at (i, j, z-face k), interior faces only
        const int oc = o_ce_face(k, nz);
        const Real F0 = upwind_ord(centred_ord(uz[1], uz[2], uz[3], uz[4], oc), wx, 5);
        const Real F1 = upwind_ord(centred_ord(W.u_ip_z[0], W.u_ip_z[1], W.u_ip_z[2], W.u_ip_z[3], oc), wx + 1, 5);
        const Real G0 = upwind_ord(centred_ord(vz[1], vz[2], vz[3], vz[4], oc), wy, 5);
        const Real G1 = upwind_ord(centred_ord(W.v_jp_z[0], W.v_jp_z[1], W.v_jp_z[2], W.v_jp_z[3], oc), wy + 1, 5);
        const Real H0 = (CARRY && k >= 2) ? zf->w : upwind_z_cen(centred_ord(wz[1], wz[2], wz[3], wz[4], o_ce_cen(k - 1, nz)), wz, k - 1, nz, C.variant);
        const Real H1 = upwind_z_cen(centred_ord(wz[2], wz[3], wz[4], wz[5], o_ce_cen(k, nz)), wz + 1, k, nz, C.variant);
        if (CARRY) zf->w = H1;
        const Real lap = (wx[4] - Real(2) * w0 + wx[2]) * C.idx2 + (wy[4] - Real(2) * w0 + wy[2]) * C.idy2 + (wz[4] - Real(2) * w0 + wz[2]) * C.idz2;
        g.w = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap + Real(0.5) * (bz[2] + b0);
    }
    return g;
}
// U* = U + dt (gamma G + zeta G-); G- <- G.  `cell` = index inside a field.  The previous-stage tendencies are fetched by the
// caller BEFORE it gathers its windows (load_prev), so that their latency hides behind the whole flux evaluation.
// (`Gc`, `Pc`: the cell's slot in field b of the tendency slab / the predicted state; the other fields sit nc values apart — w of the
// state too, since D.gw = 3 nc — so that a caller that knows the extents at compile time addresses everything off one pointer)
template <typename Real>
RBC_HD Tend<Real> load_prev_at(const Real* Gc, int nc, bool use_prev)
{
    Tend<Real> p{Real(0), Real(0), Real(0), Real(0)};
    if (use_prev) { p.b = Gc[0]; p.u = Gc[nc]; p.v = Gc[2 * nc]; p.w = Gc[3 * nc]; }
    return p;
}
template <typename Real>
RBC_HD Tend<Real> load_prev(const Dims& D, const Real* G, int cell, bool use_prev)
{
    return load_prev_at<Real>(G + cell, D.nc, use_prev);
}
template <typename Real>
RBC_HD void rk3_substep_store_at(const Dims& D, Real* Pc, Real* Gc, int k, const Windows<Real>& W, const Tend<Real>& g, const Tend<Real>& prev, Real dt,
                                 Real gam, Real zet, bool store_g)
{
    const int nc = D.nc;
    Pc[0] = W.bx[3] + dt * (gam * g.b + zet * prev.b);
    Pc[nc] = W.ux[3] + dt * (gam * g.u + zet * prev.u);
    Pc[2 * nc] = W.vx[3] + dt * (gam * g.v + zet * prev.v);
    Pc[3 * nc] = (k >= 1) ? W.wx[3] + dt * (gam * g.w + zet * prev.w) : Real(0);
    if (k == D.nz - 1) Pc[3 * nc + D.ncol] = Real(0);              // top wall face
    if (store_g) { Gc[0] = g.b; Gc[nc] = g.u; Gc[2 * nc] = g.v; Gc[3 * nc] = g.w; }
}
template <typename Real>
RBC_HD void rk3_substep_store(const Dims& D, Real* P, Real* G, int cell, int colz, int k, const Windows<Real>& W, const Tend<Real>& g,
                              const Tend<Real>& prev, Real dt, Real gam, Real zet, bool store_g)
{
    (void)colz;
    rk3_substep_store_at<Real>(D, P + cell, G + cell, k, W, g, prev, dt, gam, zet, store_g);
}

// one thread (work item) per cell, every window gathered from global memory
template <typename Real>
RBC_HD void cell_tendency(const Dims& D, const ConstsG<Real>& C, Real nu, Real kappa, const Real* RBC_RESTRICT S, Real* P, Real* G,
                          const Real* RBC_RESTRICT Tb, int cell, Real dt, Real gam, Real zet, bool use_prev, bool store_g)
{
    const int nx = D.nx, ny = D.ny, nz = D.nz, ncol = D.ncol;
    const int i = cell & (nx - 1), j = (cell >> D.lx2) & (ny - 1), k = cell >> (D.lx2 + D.ly2);      // nx, ny are powers of two
    // element offsets inside one field: x- and y-windows on level k, z-windows through the column; 32-bit, shared by all fields
    int xo[7], yo[7], ox[7], oy[7], ozc[7], ozw[7];
    const int lev = k * ncol, colz = j * nx + i;
    RBC_UNROLL
    for (int d = 0; d < 7; ++d) {
        const int kk = k + d - 3;
        xo[d] = (i + d - 3) & (nx - 1);
        yo[d] = ((j + d - 3) & (ny - 1)) << D.lx2;
        ox[d] = lev + (j << D.lx2) + xo[d];
        oy[d] = lev + yo[d] + i;
        ozc[d] = (kk < 0 ? 0 : (kk > nz - 1 ? nz - 1 : kk)) * ncol + colz;      // cell-centred fields: nz levels
        ozw[d] = (kk < 0 ? 0 : (kk > nz ? nz : kk)) * ncol + colz;              // w: nz + 1 faces
    }
    const Real* RBC_RESTRICT pb = S + D.gb;
    const Real* RBC_RESTRICT pu = S + D.gu;
    const Real* RBC_RESTRICT pv = S + D.gv;
    const Real* RBC_RESTRICT pw = S + D.gw;
    const Tend<Real> prev = load_prev<Real>(D, G, cell, use_prev);
    Windows<Real> W;
    RBC_UNROLL
    for (int d = 0; d < 7; ++d) {
        W.bx[d] = pb[ox[d]]; W.by[d] = pb[oy[d]]; W.bz[d] = pb[ozc[d]];
        W.ux[d] = pu[ox[d]]; W.uy[d] = pu[oy[d]]; W.uz[d] = pu[ozc[d]];
        W.vx[d] = pv[ox[d]]; W.vy[d] = pv[oy[d]]; W.vz[d] = pv[ozc[d]];
        W.wx[d] = pw[ox[d]]; W.wy[d] = pw[oy[d]]; W.wz[d] = pw[ozw[d]];
    }
    const int up1 = ozw[4] - colz;                                   // level offset of face k + 1 (<= nz: always there)
    RBC_UNROLL
    for (int q = 0; q < 4; ++q) {
        W.v_jp[q] = pv[lev + yo[4] + xo[q + 1]];
        W.u_ip[q] = pu[lev + yo[q + 1] + xo[4]];
        W.w_kp_x[q] = pw[ox[q + 1] - lev + up1];
        W.w_kp_y[q] = pw[oy[q + 1] - lev + up1];
        W.u_ip_z[q] = pu[ozc[q + 1] - colz + (j << D.lx2) + xo[4]];
        W.v_jp_z[q] = pv[ozc[q + 1] - colz + yo[4] + i];
    }
    const Tend<Real> g = tendency_from_windows<Real>(C, nu, kappa, nz, k, W, (k == 0) ? Tb[colz] : Real(0));
    rk3_substep_store<Real>(D, P, G, cell, colz, k, W, g, prev, dt, gam, zet, store_g);
}

// divergence of the predicted velocity at one cell of level k (the Poisson right-hand side; the 1/dtau of the
// reference's formulation cancels against the dtau of the correction and is dropped on both sides)
template <typename Real>
RBC_HD Real cell_divergence(const Dims& D, const ConstsG<Real>& C, const Real* P, int i, int j, int k)
{
    const int nx = D.nx, ny = D.ny;
    const int ip = (i + 1 == nx) ? 0 : i + 1, jp = (j + 1 == ny) ? 0 : j + 1;
    const size_t q = ((size_t)k * ny + j) * nx + i;
    return (P[D.gu + ((size_t)k * ny + j) * nx + ip] - P[D.gu + q]) * C.idx + (P[D.gv + ((size_t)k * ny + jp) * nx + i] - P[D.gv + q]) * C.idy +
           (P[D.gw + q + D.ncol] - P[D.gw + q]) * C.idz;
}

// ------------------------------------------------------------------------------------------
// radix-2 butterflies on a plane Z[ny][nx] of complex values (shared memory on the device).  One call = one butterfly of one
// stage; a stage is n_lines * n/2 butterflies followed by a barrier.  Forward = decimation in frequency (spans n/2 .. 1),
// inverse = decimation in time (spans 1 .. n/2); together they need no bit-reversal pass.
//   along x: line = row (ny lines of nx), element stride 1;  along y: line = column (nx lines of ny), element stride nx
//   tw: [n/2] complex, tw[m] = exp(-2 pi i m / n)
// All sizes are powers of two and passed as logarithms (ln = log2 n, ls = log2 span, lnl = log2 n_lines): the index split is
// shifts and masks.  `lines_fastest` picks which index runs fastest over consecutive work items (= consecutive threads):
// along y it must be the line, so that a warp touches 32 neighbouring columns of one row (conflict-free) instead of 32 rows
// of one column (stride nx complex numbers = one bank: the first version's y passes were 32-way bank conflicts).
// ------------------------------------------------------------------------------------------
struct Fly {
    int a, b, tw;      // the two elements of the butterfly and its twiddle index
};
RBC_HD Fly fly_index(int ln, int ls, int lnl, int line_stride, int elem_stride, bool lines_fastest, int item)
{
    const int lh = ln - 1;                                   // log2 of the butterflies per line
    const int line = lines_fastest ? (item & ((1 << lnl) - 1)) : (item >> lh);
    const int q = lines_fastest ? (item >> lnl) : (item & ((1 << lh) - 1));
    const int jj = q & ((1 << ls) - 1), grp = q >> ls;
    Fly f;
    f.a = line * line_stride + ((grp << (ls + 1)) + jj) * elem_stride;
    f.b = f.a + (elem_stride << ls);
    f.tw = jj << (lh - ls);
    return f;
}
template <typename Real>
RBC_HD void butterfly_dif(cx<Real>* Z, int ln, int ls, int lnl, int line_stride, int elem_stride, bool lines_fastest, const cx<Real>* tw, int item)
{
    const Fly f = fly_index(ln, ls, lnl, line_stride, elem_stride, lines_fastest, item);
    const cx<Real> xa = Z[f.a], xb = Z[f.b], w = tw[f.tw];
    Z[f.a] = cx<Real>{xa.re + xb.re, xa.im + xb.im};
    const Real dr = xa.re - xb.re, di = xa.im - xb.im;
    Z[f.b] = cx<Real>{dr * w.re - di * w.im, dr * w.im + di * w.re};
}
template <typename Real>
RBC_HD void butterfly_dit_inv(cx<Real>* Z, int ln, int ls, int lnl, int line_stride, int elem_stride, bool lines_fastest, const cx<Real>* tw, int item)
{
    const Fly f = fly_index(ln, ls, lnl, line_stride, elem_stride, lines_fastest, item);
    const cx<Real> xa = Z[f.a], xb = Z[f.b], w = tw[f.tw];
    const Real tr = xb.re * w.re + xb.im * w.im, ti = xb.im * w.re - xb.re * w.im;      // xb * conj(w)
    Z[f.a] = cx<Real>{xa.re + tr, xa.im + ti};
    Z[f.b] = cx<Real>{xa.re - tr, xa.im - ti};
}
// Two consecutive radix-2 stages on the four elements they connect (a radix-4 step written as its two radix-2 halves: the same
// operations on the same values, so the result is bit-identical to running the stages one after the other): half the barriers
// and half the shared-memory traffic of the transform.  One work item = 4 elements; n/4 items per line.
//   forward (DIF): spans S = 2^ls, then S/2 (ls >= 1).  inverse (DIT): spans S = 2^ls, then 2S (ls <= ln - 2).
template <typename Real>
RBC_HD cx<Real> cmulw(cx<Real> d, cx<Real> w) { return cx<Real>{d.re * w.re - d.im * w.im, d.re * w.im + d.im * w.re}; }
template <typename Real>
RBC_HD cx<Real> cmulwc(cx<Real> b, cx<Real> w) { return cx<Real>{b.re * w.re + b.im * w.im, b.im * w.re - b.re * w.im}; }      // b * conj(w)
template <typename Real>
RBC_HD void butterfly2_dif(cx<Real>* Z, int ln, int ls, int lnl, int line_stride, int elem_stride, bool lines_fastest, const cx<Real>* tw, int item)
{
    const int lq = ln - 2, lh = ln - 1, S = 1 << ls, Sh = S >> 1;
    const int line = lines_fastest ? (item & ((1 << lnl) - 1)) : (item >> lq);
    const int q = lines_fastest ? (item >> lnl) : (item & ((1 << lq) - 1));
    const int j = q & (Sh - 1), grp = q >> (ls - 1);
    const int e0 = line * line_stride + ((grp << (ls + 1)) + j) * elem_stride, e1 = e0 + Sh * elem_stride, e2 = e0 + S * elem_stride,
              e3 = e2 + Sh * elem_stride;
    const cx<Real> x0 = Z[e0], x1 = Z[e1], x2 = Z[e2], x3 = Z[e3];
    const cx<Real> wa = tw[j << (lh - ls)], wb = tw[(j + Sh) << (lh - ls)], wc = tw[j << (lh - ls + 1)];
    // stage with span S: (e0, e2) and (e1, e3)
    const cx<Real> y0{x0.re + x2.re, x0.im + x2.im}, y2 = cmulw(cx<Real>{x0.re - x2.re, x0.im - x2.im}, wa);
    const cx<Real> y1{x1.re + x3.re, x1.im + x3.im}, y3 = cmulw(cx<Real>{x1.re - x3.re, x1.im - x3.im}, wb);
    // stage with span S/2: (e0, e1) and (e2, e3)
    Z[e0] = cx<Real>{y0.re + y1.re, y0.im + y1.im};
    Z[e1] = cmulw(cx<Real>{y0.re - y1.re, y0.im - y1.im}, wc);
    Z[e2] = cx<Real>{y2.re + y3.re, y2.im + y3.im};
    Z[e3] = cmulw(cx<Real>{y2.re - y3.re, y2.im - y3.im}, wc);
}
template <typename Real>
RBC_HD void butterfly2_dit_inv(cx<Real>* Z, int ln, int ls, int lnl, int line_stride, int elem_stride, bool lines_fastest, const cx<Real>* tw, int item)
{
    const int lq = ln - 2, lh = ln - 1, S = 1 << ls;
    const int line = lines_fastest ? (item & ((1 << lnl) - 1)) : (item >> lq);
    const int q = lines_fastest ? (item >> lnl) : (item & ((1 << lq) - 1));
    const int j = q & (S - 1), grp = q >> ls;
    const int e0 = line * line_stride + ((grp << (ls + 2)) + j) * elem_stride, e1 = e0 + S * elem_stride, e2 = e1 + S * elem_stride,
              e3 = e2 + S * elem_stride;
    const cx<Real> x0 = Z[e0], x1 = Z[e1], x2 = Z[e2], x3 = Z[e3];
    const cx<Real> wa = tw[j << (lh - ls)], wc = tw[j << (lh - ls - 1)], wd = tw[(j + S) << (lh - ls - 1)];
    // stage with span S: (e0, e1) and (e2, e3), both at position j of their group
    const cx<Real> t1 = cmulwc(x1, wa), t3 = cmulwc(x3, wa);
    const cx<Real> y0{x0.re + t1.re, x0.im + t1.im}, y1{x0.re - t1.re, x0.im - t1.im};
    const cx<Real> y2{x2.re + t3.re, x2.im + t3.im}, y3{x2.re - t3.re, x2.im - t3.im};
    // stage with span 2S: (e0, e2) at position j, (e1, e3) at position j + S
    const cx<Real> u2 = cmulwc(y2, wc), u3 = cmulwc(y3, wd);
    Z[e0] = cx<Real>{y0.re + u2.re, y0.im + u2.im};
    Z[e2] = cx<Real>{y0.re - u2.re, y0.im - u2.im};
    Z[e1] = cx<Real>{y1.re + u3.re, y1.im + u3.im};
    Z[e3] = cx<Real>{y1.re - u3.re, y1.im - u3.im};
}
// the passes of one plane as (direction, span) sequences; `run(n_items, fn)` executes fn(item) for every item and then
// synchronises (a strided thread loop + __syncthreads on the device, a plain loop in the emulator).  Stages are taken two at a
// time; an odd number of stages leaves one plain radix-2 stage.
template <typename Real, typename Run>
RBC_HD void line_fft_forward(const Dims& D, cx<Real>* Z, int ln, int lnl, int line_stride, int elem_stride, bool lines_fastest, const cx<Real>* tw, Run run)
{
    int ls = ln - 1;
    RBC_UNROLL
    for (; ls >= 1; ls -= 2) run(D.ncol >> 2, [&](int it) { butterfly2_dif<Real>(Z, ln, ls, lnl, line_stride, elem_stride, lines_fastest, tw, it); });
    if (ls == 0) run(D.ncol >> 1, [&](int it) { butterfly_dif<Real>(Z, ln, 0, lnl, line_stride, elem_stride, lines_fastest, tw, it); });
}
template <typename Real, typename Run>
RBC_HD void line_fft_inverse(const Dims& D, cx<Real>* Z, int ln, int lnl, int line_stride, int elem_stride, bool lines_fastest, const cx<Real>* tw, Run run)
{
    int ls = 0;
    RBC_UNROLL
    for (; ls + 1 <= ln - 1; ls += 2) run(D.ncol >> 2, [&](int it) { butterfly2_dit_inv<Real>(Z, ln, ls, lnl, line_stride, elem_stride, lines_fastest, tw, it); });
    if (ls == ln - 1) run(D.ncol >> 1, [&](int it) { butterfly_dit_inv<Real>(Z, ln, ls, lnl, line_stride, elem_stride, lines_fastest, tw, it); });
}
// `pitch`: elements between consecutive rows of the plane (0 = nx).  `x_rows_fastest`: consecutive work items of the x passes take the
// same butterfly of consecutive ROWS instead of consecutive butterflies of one row — with an odd pitch that makes every shared-memory
// access of the x passes conflict-free on the device (consecutive butterflies of a row are 4 or 16 elements apart in the later stages:
// 4-way bank conflicts), and a warp shares its twiddles.  The set of butterflies and their order within a line do not change.
template <typename Real, typename Run>
RBC_HD void plane_fft_forward(const Dims& D, cx<Real>* Z, const cx<Real>* twx, const cx<Real>* twy, Run run, int pitch = 0, bool x_rows_fastest = false)
{
    const int pt = pitch ? pitch : D.nx;
    line_fft_forward<Real>(D, Z, D.lx2, D.ly2, pt, 1, x_rows_fastest, twx, run);
    line_fft_forward<Real>(D, Z, D.ly2, D.lx2, 1, pt, true, twy, run);
}
template <typename Real, typename Run>
RBC_HD void plane_fft_inverse(const Dims& D, cx<Real>* Z, const cx<Real>* twx, const cx<Real>* twy, Run run, int pitch = 0, bool x_rows_fastest = false)
{
    const int pt = pitch ? pitch : D.nx;
    line_fft_inverse<Real>(D, Z, D.ly2, D.lx2, 1, pt, true, twy, run);
    line_fft_inverse<Real>(D, Z, D.lx2, D.ly2, pt, 1, x_rows_fastest, twx, run);
}

RBC_HD int bitrev(int v, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}

// ------------------------------------------------------------------------------------------
// Two levels per complex plane.  The divergence is real, so plane p of the spectral array carries level 2p in its real and level
// 2p + 1 in its imaginary part: half the FFT planes, half the spectral traffic.  The transform Z of a + i b is A + i B with the
// Hermitian spectra A, B of the two real planes, so the z-solve — real coefficients, one mode at a time — separates them through
// the pair of positions holding the wavenumbers m and -m:
//     A(m) = (Z(m) + conj Z(-m)) / 2,      B(m) = (Z(m) - conj Z(-m)) / (2 i),
// solves  T x = scale * F  for the complex column x_k(m), k = 0 .. nz-1 (Neumann ends; cp[k][position] are the host-built
// reciprocal pivots 1 / (diag_k - cp_{k-1}), the same for m and -m), and writes back  X(m) = x_even + i x_odd,
// X(-m) = conj(x_even) + i conj(x_odd), whose inverse transform is phi(level 2p) + i phi(level 2p + 1).
// One work item per storage position q; the item of the smaller position of a pair does the work (a position that is its own
// partner — wavenumbers 0 or n/2 in both directions — carries two real columns).  Between the sweeps the pair's two slots hold
// the forward-eliminated even and odd level.  Positions are in bit-reversed order (no permutation pass in the FFTs).
// ------------------------------------------------------------------------------------------
RBC_HD int partner_position(const Dims& D, int q)
{
    const int rx = q & (D.nx - 1), ry = q >> D.lx2;
    const int kx = bitrev(rx, D.lx2), ky = bitrev(ry, D.ly2);
    return (bitrev((D.ny - ky) & (D.ny - 1), D.ly2) << D.lx2) + bitrev((D.nx - kx) & (D.nx - 1), D.lx2);
}
template <typename Real>
RBC_HD void mode_pair_thomas(const Dims& D, cx<Real>* Zs, const Real* cp, Real scale, int q)
{
    const int nz = D.nz, ncol = D.ncol, nzp = (nz + 1) >> 1;
    const int qp = partner_position(D, q);
    if (qp < q) return;
    const Real h = Real(0.5) * scale;
    cx<Real> d{Real(0), Real(0)};                            // forward elimination, level after level
    for (int p = 0; p < nzp; ++p) {
        const cx<Real> A = Zs[(size_t)p * ncol + q], B = Zs[(size_t)p * ncol + qp];
        const bool odd = 2 * p + 1 < nz;
        const Real c0 = cp[(size_t)(2 * p) * ncol + q], c1 = odd ? cp[(size_t)(2 * p + 1) * ncol + q] : Real(0);
        const cx<Real> d0{((A.re + B.re) * h - d.re) * c0, ((A.im - B.im) * h - d.im) * c0};
        const cx<Real> d1{((A.im + B.im) * h - d0.re) * c1, ((B.re - A.re) * h - d0.im) * c1};
        d = odd ? d1 : d0;
        if (qp == q) Zs[(size_t)p * ncol + q] = cx<Real>{d0.re, d1.re};       // two real columns
        else { Zs[(size_t)p * ncol + q] = d0; Zs[(size_t)p * ncol + qp] = d1; }
    }
    cx<Real> x{Real(0), Real(0)};                            // back substitution: x_k = d_k - c_k x_{k+1}, x_{nz} = 0
    for (int p = nzp - 1; p >= 0; --p) {
        const bool odd = 2 * p + 1 < nz;
        cx<Real> d0 = Zs[(size_t)p * ncol + q], d1 = Zs[(size_t)p * ncol + qp];
        if (qp == q) { d1 = cx<Real>{d0.im, Real(0)}; d0 = cx<Real>{d0.re, Real(0)}; }
        const Real c0 = cp[(size_t)(2 * p) * ncol + q], c1 = odd ? cp[(size_t)(2 * p + 1) * ncol + q] : Real(0);
        cx<Real> x1{Real(0), Real(0)};
        if (odd) { x1 = cx<Real>{d1.re - c1 * x.re, d1.im - c1 * x.im}; x = x1; }
        const cx<Real> x0{d0.re - c0 * x.re, d0.im - c0 * x.im};
        x = x0;
        Zs[(size_t)p * ncol + q] = cx<Real>{x0.re - x1.im, x0.im + x1.re};                       // x_even + i x_odd
        if (qp != q) Zs[(size_t)p * ncol + qp] = cx<Real>{x0.re + x1.im, x1.re - x0.im};         // conj(x_even) + i conj(x_odd)
    }
}

// velocity correction of one cell: U = U* - grad phi (interior w faces only)
template <typename Real>
RBC_HD void cell_correct(const Dims& D, const ConstsG<Real>& C, Real* P, const Real* phi, int cell)
{
    const int nx = D.nx, ny = D.ny;
    const int i = cell % nx, j = (cell / nx) % ny, k = cell / D.ncol;
    const int im = (i == 0) ? nx - 1 : i - 1, jm = (j == 0) ? ny - 1 : j - 1;
    const Real ph = phi[cell];
    P[D.gu + cell] -= (ph - phi[((size_t)k * ny + j) * nx + im]) * C.idx;
    P[D.gv + cell] -= (ph - phi[((size_t)k * ny + jm) * nx + i]) * C.idy;
    if (k >= 1) P[D.gw + cell] -= (ph - phi[cell - D.ncol]) * C.idz;
}

// ------------------------------------------------------------------------------------------
// host-side construction (fp64)
// ------------------------------------------------------------------------------------------
struct HostConfigG {
    double ra, pr, lx, ly, lz, b_top, delta_b, heater_limit, heater_duration, dt_solver, episode_length;
    int heaters;
};
template <typename Real>
inline ConstsG<Real> make_consts(const Dims& D, const HostConfigG& h)
{
    ConstsG<Real> C;
    const double dx = h.lx / D.nx, dy = h.ly / D.ny, dz = h.lz / D.nz;
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
    C.variant = 0;
    return C;
}
// reciprocal Thomas pivots per (k, ry, rx): position (ry, rx) of the transformed plane holds mode (bitrev ry, bitrev rx)
inline void build_pivots_host(const Dims& D, double lx, double ly, double lz, double* cp /*nz*ncol*/)
{
    const double PI = 3.14159265358979323846;
    const double dx = lx / D.nx, dy = ly / D.ny, dz = lz / D.nz;
    for (int ry = 0; ry < D.ny; ++ry)
        for (int rx = 0; rx < D.nx; ++rx) {
            const int kx = bitrev(rx, D.lx2), ky = bitrev(ry, D.ly2);
            const double sx = 2 * sin(PI * kx / D.nx) / dx, sy = 2 * sin(PI * ky / D.ny) / dy, lam = (sx * sx + sy * sy) * dz * dz;
            double prev = 0.0;
            for (int k = 0; k < D.nz; ++k) {
                double dg = -(2.0 + lam);
                if (k == 0 || k == D.nz - 1) dg += 1.0;
                if (kx == 0 && ky == 0 && k == 0) dg -= 1.0;     // pin the null space of the mean mode
                const double iv = 1.0 / (dg - prev);
                cp[(size_t)k * D.ncol + ry * D.nx + rx] = iv;
                prev = iv;
            }
        }
}
inline void build_twiddles_host(int n, double* tw /*n/2 complex*/)
{
    const double PI = 3.14159265358979323846;
    for (int m = 0; m < n / 2; ++m) { tw[2 * m] = cos(2 * PI * m / n); tw[2 * m + 1] = -sin(2 * PI * m / n); }
}

}  // namespace rbc3dg
