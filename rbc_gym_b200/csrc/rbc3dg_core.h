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
// field access with the periodic wrap in x, y and a clamp in z (clamped values are only ever loaded into window slots
// that the reduced-order stencils next to the walls do not read)
// ------------------------------------------------------------------------------------------
template <typename Real>
struct Field {
    const Real* p;
    int nx, ny, nlev;
    RBC_HD Real at(int i, int j, int k) const
    {
        i = i < 0 ? i + nx : (i >= nx ? i - nx : i);
        j = j < 0 ? j + ny : (j >= ny ? j - ny : j);
        k = k < 0 ? 0 : (k >= nlev ? nlev - 1 : k);
        return p[((size_t)k * ny + j) * nx + i];
    }
    // value at offset m (in cells) from (i,j,k) along direction dir (0 x, 1 y, 2 z)
    RBC_HD Real off(int dir, int i, int j, int k, int m) const
    {
        return dir == 0 ? at(i + m, j, k) : (dir == 1 ? at(i, j + m, k) : at(i, j, k + m));
    }
};

// flux through the face "idx-1 | idx" along dq of the advected field Q, carried by the advecting field A interpolated to
// that face along da (centred, order oa = 4 or 2); Q reconstructed upwind-biased with order oq (5, 3 or 1)
template <typename Real>
RBC_HD Real flux(const Field<Real>& A, int da, int oa, const Field<Real>& Q, int dq, int oq, int i, int j, int k)
{
    const Real a = centred_ord(A.off(da, i, j, k, -2), A.off(da, i, j, k, -1), A.off(da, i, j, k, 0), A.off(da, i, j, k, 1), oa);
    Real win[6];
    for (int m = 0; m < 6; ++m) win[m] = Q.off(dq, i, j, k, m - 3);
    return upwind_ord(a, win, oq);
}
// tracer flux: the advecting velocity is the face value itself
template <typename Real>
RBC_HD Real flux_face(Real vel, const Field<Real>& Q, int dq, int oq, int i, int j, int k)
{
    Real win[6];
    for (int m = 0; m < 6; ++m) win[m] = Q.off(dq, i, j, k, m - 3);
    return upwind_ord(vel, win, oq);
}

// ------------------------------------------------------------------------------------------
// tendency + RK3 substep of one cell (buoyancy added to G_w, which after the projection is identical to
// Oceananigans' hydrostatic-pressure split — the 3D state exposes no pressure channel)
//   S: current state, P: predicted state (out), G: tendency slab [4][nc] read (previous stage) and rewritten in place
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void cell_tendency(const Dims& D, const ConstsG<Real>& C, Real nu, Real kappa, const Real* S, Real* P, Real* G, const Real* Tb,
                          int cell, Real dt, Real gam, Real zet, bool use_prev, bool store_g)
{
    const int nx = D.nx, ny = D.ny, nz = D.nz;
    const int i = cell % nx, j = (cell / nx) % ny, k = cell / D.ncol;
    const Field<Real> B{S + D.gb, nx, ny, nz}, U{S + D.gu, nx, ny, nz}, V{S + D.gv, nx, ny, nz}, W{S + D.gw, nx, ny, nz + 1};
    const int of0 = (k >= 1) ? o_up_face(k, nz) : 0, of1 = (k + 1 <= nz - 1) ? o_up_face(k + 1, nz) : 0;
    const Real b0 = B.at(i, j, k), u0 = U.at(i, j, k), v0 = V.at(i, j, k), w0 = W.at(i, j, k);
    Real gb, gu, gv, gw = Real(0);
    {   // tracer
        const Real Fx0 = flux_face(u0, B, 0, 5, i, j, k), Fx1 = flux_face(U.at(i + 1, j, k), B, 0, 5, i + 1, j, k);
        const Real Fy0 = flux_face(v0, B, 1, 5, i, j, k), Fy1 = flux_face(V.at(i, j + 1, k), B, 1, 5, i, j + 1, k);
        const Real Fz0 = of0 ? flux_face(w0, B, 2, of0, i, j, k) : Real(0);
        const Real Fz1 = of1 ? flux_face(W.at(i, j, k + 1), B, 2, of1, i, j, k + 1) : Real(0);
        const Real bdn = (k == 0) ? Real(2) * Tb[j * nx + i] - b0 : B.at(i, j, k - 1);
        const Real bup = (k == nz - 1) ? Real(2) * C.b_top - b0 : B.at(i, j, k + 1);
        const Real lap = (B.at(i + 1, j, k) - Real(2) * b0 + B.at(i - 1, j, k)) * C.idx2 + (B.at(i, j + 1, k) - Real(2) * b0 + B.at(i, j - 1, k)) * C.idy2 +
                         (bup - Real(2) * b0 + bdn) * C.idz2;
        gb = -((Fx1 - Fx0) * C.idx + (Fy1 - Fy0) * C.idy + (Fz1 - Fz0) * C.idz) + kappa * lap;
    }
    {   // u at (x-face i, j, k)
        const Real F0 = flux(U, 0, 4, U, 0, 5, i, j, k), F1 = flux(U, 0, 4, U, 0, 5, i + 1, j, k);
        const Real G0 = flux(V, 0, 4, U, 1, 5, i, j, k), G1 = flux(V, 0, 4, U, 1, 5, i, j + 1, k);
        const Real H0 = of0 ? flux(W, 0, 4, U, 2, of0, i, j, k) : Real(0);
        const Real H1 = of1 ? flux(W, 0, 4, U, 2, of1, i, j, k + 1) : Real(0);
        const Real dn = (k == 0) ? -u0 : U.at(i, j, k - 1), up = (k == nz - 1) ? -u0 : U.at(i, j, k + 1);
        const Real lap = (U.at(i + 1, j, k) - Real(2) * u0 + U.at(i - 1, j, k)) * C.idx2 + (U.at(i, j + 1, k) - Real(2) * u0 + U.at(i, j - 1, k)) * C.idy2 +
                         (up - Real(2) * u0 + dn) * C.idz2;
        gu = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap;
    }
    {   // v at (i, y-face j, k)
        const Real F0 = flux(U, 1, 4, V, 0, 5, i, j, k), F1 = flux(U, 1, 4, V, 0, 5, i + 1, j, k);
        const Real G0 = flux(V, 1, 4, V, 1, 5, i, j, k), G1 = flux(V, 1, 4, V, 1, 5, i, j + 1, k);
        const Real H0 = of0 ? flux(W, 1, 4, V, 2, of0, i, j, k) : Real(0);
        const Real H1 = of1 ? flux(W, 1, 4, V, 2, of1, i, j, k + 1) : Real(0);
        const Real dn = (k == 0) ? -v0 : V.at(i, j, k - 1), up = (k == nz - 1) ? -v0 : V.at(i, j, k + 1);
        const Real lap = (V.at(i + 1, j, k) - Real(2) * v0 + V.at(i - 1, j, k)) * C.idx2 + (V.at(i, j + 1, k) - Real(2) * v0 + V.at(i, j - 1, k)) * C.idy2 +
                         (up - Real(2) * v0 + dn) * C.idz2;
        gv = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap;
    }
    if (k >= 1) {   // w at (i, j, z-face k), interior faces only
        const int oc = o_ce_face(k, nz);
        const Real F0 = flux(U, 2, oc, W, 0, 5, i, j, k), F1 = flux(U, 2, oc, W, 0, 5, i + 1, j, k);
        const Real G0 = flux(V, 2, oc, W, 1, 5, i, j, k), G1 = flux(V, 2, oc, W, 1, 5, i, j + 1, k);
        const Real H0 = flux(W, 2, o_ce_cen(k - 1, nz), W, 2, o_up_cen(k - 1, nz), i, j, k);
        const Real H1 = flux(W, 2, o_ce_cen(k, nz), W, 2, o_up_cen(k, nz), i, j, k + 1);
        const Real lap = (W.at(i + 1, j, k) - Real(2) * w0 + W.at(i - 1, j, k)) * C.idx2 + (W.at(i, j + 1, k) - Real(2) * w0 + W.at(i, j - 1, k)) * C.idy2 +
                         (W.at(i, j, k + 1) - Real(2) * w0 + W.at(i, j, k - 1)) * C.idz2;
        gw = -((F1 - F0) * C.idx + (G1 - G0) * C.idy + (H1 - H0) * C.idz) + nu * lap + Real(0.5) * (B.at(i, j, k - 1) + b0);
    }
    // U* = U + dt (gamma G + zeta G-); G- <- G
    Real pb = Real(0), pu = Real(0), pv = Real(0), pw = Real(0);
    if (use_prev) { pb = G[cell]; pu = G[D.nc + cell]; pv = G[2 * D.nc + cell]; pw = G[3 * D.nc + cell]; }
    P[D.gb + cell] = b0 + dt * (gam * gb + zet * pb);
    P[D.gu + cell] = u0 + dt * (gam * gu + zet * pu);
    P[D.gv + cell] = v0 + dt * (gam * gv + zet * pv);
    P[D.gw + cell] = (k >= 1) ? w0 + dt * (gam * gw + zet * pw) : Real(0);
    if (k == nz - 1) P[D.gw + D.nc + j * nx + i] = Real(0);          // top wall face
    if (store_g) { G[cell] = gb; G[D.nc + cell] = gu; G[2 * D.nc + cell] = gv; G[3 * D.nc + cell] = gw; }
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
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void butterfly_dif(cx<Real>* Z, int n, int span, int line_stride, int elem_stride, const cx<Real>* tw, int item)
{
    const int half = n >> 1, line = item / half, q = item % half, jj = q % span, grp = q / span;
    const int a = line * line_stride + (grp * 2 * span + jj) * elem_stride, b = a + span * elem_stride;
    const cx<Real> xa = Z[a], xb = Z[b], w = tw[jj * (half / span)];
    Z[a] = cx<Real>{xa.re + xb.re, xa.im + xb.im};
    const Real dr = xa.re - xb.re, di = xa.im - xb.im;
    Z[b] = cx<Real>{dr * w.re - di * w.im, dr * w.im + di * w.re};
}
template <typename Real>
RBC_HD void butterfly_dit_inv(cx<Real>* Z, int n, int span, int line_stride, int elem_stride, const cx<Real>* tw, int item)
{
    const int half = n >> 1, line = item / half, q = item % half, jj = q % span, grp = q / span;
    const int a = line * line_stride + (grp * 2 * span + jj) * elem_stride, b = a + span * elem_stride;
    const cx<Real> xa = Z[a], xb = Z[b], w = tw[jj * (half / span)];
    const Real tr = xb.re * w.re + xb.im * w.im, ti = xb.im * w.re - xb.re * w.im;      // xb * conj(w)
    Z[a] = cx<Real>{xa.re + tr, xa.im + ti};
    Z[b] = cx<Real>{xa.re - tr, xa.im - ti};
}

// ------------------------------------------------------------------------------------------
// tridiagonal solve in z of one horizontal mode (Neumann ends), in place on the spectral array Zs[k][plane index];
// cp[k][plane index] are the host-built reciprocal pivots 1 / (diag_k - cp_{k-1}), scale = dz^2
// ------------------------------------------------------------------------------------------
template <typename Real>
RBC_HD void mode_thomas(const Dims& D, cx<Real>* Zs, const Real* cp, Real scale, int m)
{
    const int nz = D.nz, ncol = D.ncol;
    Real dr = Real(0), di = Real(0);
    for (int k = 0; k < nz; ++k) {
        const Real c = cp[(size_t)k * ncol + m];
        cx<Real> f = Zs[(size_t)k * ncol + m];
        dr = (f.re * scale - dr) * c;
        di = (f.im * scale - di) * c;
        Zs[(size_t)k * ncol + m] = cx<Real>{dr, di};
    }
    Real xr = dr, xi = di;
    for (int k = nz - 2; k >= 0; --k) {
        const Real c = cp[(size_t)k * ncol + m];
        const cx<Real> d = Zs[(size_t)k * ncol + m];
        xr = d.re - c * xr;
        xi = d.im - c * xi;
        Zs[(size_t)k * ncol + m] = cx<Real>{xr, xi};
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
    return C;
}
inline int bitrev(int v, int bits)
{
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
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
