// Host emulator of the stage-streaming 3D kernels for arbitrary grids (rbc3dg_core.h): every kernel of rbc3dg_lib.cu is a loop
// over the same per-work-item functions here.  Test infrastructure (see emu_rbc2d.cpp).
#include <vector>

#include "../../rbc_gym_b200/csrc/rbc3dg_core.h"

using namespace rbc3dg;

// one FFT stage in the emulator: the work items one after the other (the device runs them on the threads of a CTA + barrier)
struct SerialRun {
    template <typename F>
    void operator()(int n_items, F fn) const
    {
        for (int it = 0; it < n_items; ++it) fn(it);
    }
};

template <typename Real>
static void project(const Dims& D, const ConstsG<Real>& C, const HostConfigG& h, Real* P, std::vector<cx<Real>>& Z, std::vector<Real>& phi,
                    const std::vector<Real>& cp, const std::vector<cx<Real>>& twx, const std::vector<cx<Real>>& twy)
{
    const int nzp = (D.nz + 1) / 2;
    for (int pz = 0; pz < nzp; ++pz) {                                // g3_div_fft_kernel: one "CTA" per pair of levels
        cx<Real>* Zk = Z.data() + (size_t)pz * D.ncol;
        const bool odd = 2 * pz + 1 < D.nz;
        for (int c = 0; c < D.ncol; ++c)
            Zk[c] = cx<Real>{cell_divergence<Real>(D, C, P, c % D.nx, c / D.nx, 2 * pz), odd ? cell_divergence<Real>(D, C, P, c % D.nx, c / D.nx, 2 * pz + 1) : Real(0)};
        plane_fft_forward<Real>(D, Zk, twx.data(), twy.data(), SerialRun{});
    }
    const Real dz = (Real)(h.lz / D.nz);
    for (int q = 0; q < D.ncol; ++q) mode_pair_thomas<Real>(D, Z.data(), cp.data(), dz * dz, q);   // g3_thomas_kernel
    for (int pz = 0; pz < nzp; ++pz) {                                // g3_ifft_kernel
        cx<Real>* Zk = Z.data() + (size_t)pz * D.ncol;
        plane_fft_inverse<Real>(D, Zk, twx.data(), twy.data(), SerialRun{});
        for (int c = 0; c < D.ncol; ++c) {
            phi[(size_t)(2 * pz) * D.ncol + c] = Zk[c].re * (Real(1) / (Real)D.ncol);
            if (2 * pz + 1 < D.nz) phi[(size_t)(2 * pz + 1) * D.ncol + c] = Zk[c].im * (Real(1) / (Real)D.ncol);
        }
    }
    for (int cell = 0; cell < D.nc; ++cell) cell_correct<Real>(D, C, P, phi.data(), cell);       // g3_correct_kernel
}

template <typename Real>
static void run(const HostConfigG& h, int nx, int ny, int nz, int B, Real* state, const float* actions, const double* ra_env, double* nusselt,
                int* nan_flag, int project_first, int nsub_override)
{
    const Dims D = make_dims(nx, ny, nz);
    const ConstsG<Real> C = make_consts<Real>(D, h);
    std::vector<double> cpd((size_t)nz * D.ncol), txd(nx), tyd(ny);
    build_pivots_host(D, h.lx, h.ly, h.lz, cpd.data());
    build_twiddles_host(nx, txd.data());
    build_twiddles_host(ny, tyd.data());
    std::vector<Real> cp(cpd.begin(), cpd.end());
    std::vector<cx<Real>> twx(nx / 2), twy(ny / 2);
    for (int m = 0; m < nx / 2; ++m) twx[m] = cx<Real>{(Real)txd[2 * m], (Real)txd[2 * m + 1]};
    for (int m = 0; m < ny / 2; ++m) twy[m] = cx<Real>{(Real)tyd[2 * m], (Real)tyd[2 * m + 1]};
    std::vector<Real> Pbuf(D.nstate), G((size_t)4 * D.nc), phi(D.nc), Tb(D.ncol);
    std::vector<cx<Real>> Z(D.nc);
    const Real gam[3] = {Real(8.0 / 15.0), Real(5.0 / 12.0), Real(3.0 / 4.0)};
    const Real zet[3] = {Real(0), Real(-17.0 / 60.0), Real(-5.0 / 12.0)};
    const int nsub = nsub_override >= 0 ? nsub_override : C.nsub;
    for (int e = 0; e < B; ++e) {
        const double ra = ra_env ? ra_env[e] : h.ra;
        const Real nu = (Real)sqrt(h.pr / ra), kappa = (Real)(1.0 / sqrt(h.pr * ra));
        Real* S = state + (size_t)e * D.nstate;
        for (int c = 0; c < D.ncol; ++c)
            Tb[c] = (Real)heater_patch_T(C.heaters, C.heater_limit, C.b_hot, actions + (size_t)e * C.heaters * C.heaters, c % nx, c / nx, nx, ny);
        Real* cur = S;
        Real* nxt = Pbuf.data();
        if (project_first) project<Real>(D, C, h, cur, Z, phi, cp, twx, twy);
        for (int sub = 0; sub < nsub; ++sub) {
            const Real dt = (sub == nsub - 1) ? C.dt_last : C.dt_full;
            for (int stage = 0; stage < 3; ++stage) {
                for (int cell = 0; cell < D.nc; ++cell)
                    cell_tendency<Real>(D, C, nu, kappa, cur, nxt, G.data(), Tb.data(), cell, dt, gam[stage], zet[stage], stage > 0, stage < 2);
                project<Real>(D, C, h, nxt, Z, phi, cp, twx, twy);
                Real* t = cur; cur = nxt; nxt = t;
            }
        }
        double acc = 0, bad = 0;
        for (int cell = 0; cell < D.nc; ++cell) {
            const int k = cell / D.ncol;
            const double b = (double)cur[D.gb + cell], u = (double)cur[D.gu + cell], v = (double)cur[D.gv + cell], w = (double)cur[D.gw + cell];
            if (b != b || u != u || v != v || w != w) bad += 1;
            acc += (b - ((1.0 - (k + 0.5) / nz) * C.delta_b_d + C.b_top_d)) * w;
        }
        if (cur != S) for (int q = 0; q < D.nstate; ++q) S[q] = cur[q];
        nusselt[e] = 1.0 + (acc / (double)D.nc) / (1.0 / sqrt(h.pr * ra));
        nan_flag[e] = bad > 0;
    }
}

extern "C" int emu_rbc3dg_step(const HostConfigG* h, int nx, int ny, int nz, int precision, int B, void* state, const float* actions,
                               const double* ra_env, double* nusselt, int* nan_flag, int project_first, int nsub_override)
{
    if (!dims_supported(nx, ny, nz)) return -2;
    if (precision == 64) run<double>(*h, nx, ny, nz, B, (double*)state, actions, ra_env, nusselt, nan_flag, project_first, nsub_override);
    else if (precision == 32) run<float>(*h, nx, ny, nz, B, (float*)state, actions, ra_env, nusselt, nan_flag, project_first, nsub_override);
    else return -1;
    return 0;
}
