// Sequential emulator of rbc2dx_core.h (the cluster kernel for generic 2D grids): the CTAs of a cluster
// run one after the other inside every phase and "remote shared memory" is another slice of one host
// arena.  Test infrastructure: lets tests/ check the exact kernel logic (slab decomposition, halo pushes,
// SPIKE tridiagonal solve, N1 x N2 FFT, cross-CTA reductions) against the fp64 oracle without a GPU.
#include <cstring>
#include <vector>

#include "../../rbc_gym_b200/csrc/rbc2dx_core.h"

using namespace rbc2dx;

// fused vector-env semantics for the next emu_rbc2dx_step call (emu_rbc2dx_set_vec); mode < 0 = plain step
static rbc2d::VecIO g_vec;

template <typename G, typename Real, bool NXTG, bool SPLIT>
static void run(const HostConfig& h, const HostWrappers& wr, double* cell_dist, int B, Real* state, const float* actions, float* obs,
                float* reward, double* nu_state, double* nu_obs, double* t, int* step_count, int* truncated, int* nan_flag,
                int project_first, int nsub, Real* pressure)
{
    using L = SmemLayoutX<G, Real, NXTG, SPLIT>;
    Consts<Real> C = make_consts<G, Real>(h, wr);
    std::vector<double> td(table_doubles<G>());
    HostTables T;
    double* p = td.data();
    T.tinv = p; p += (size_t)G::NZ * G::NX;
    T.spv = p; p += (size_t)G::NZ * G::NX;
    T.spw = p; p += (size_t)G::NZ * G::NX;
    T.cxl = p; p += (size_t)G::CL * 2 * G::CL * G::NX;
    T.cxr = p; p += (size_t)G::CL * 2 * G::CL * G::NX;
    T.twN = p; p += 2 * G::NH;
    T.tw2 = p; p += 2 * G::NH;
    build_tables_host<G>(h.lx, h.lz, T);
    std::vector<Real> tr(td.begin(), td.end());
    std::vector<unsigned char> arena(L::total * G::CL, 0);
    std::vector<Real> gm((size_t)G::CL * 2 * G::NLOC, Real(0)), nxt(NXTG ? (size_t)G::CL * G::NS_SM : 1, Real(0));
    CtxX<Real> X;
    L::fill(X);
    X.base = arena.data();
    X.arena_stride = L::total;
    X.gm = gm.data();
    X.nxt_g = NXTG ? nxt.data() : nullptr;
    const Real* q = tr.data();
    X.tinv = q; q += (size_t)G::NZ * G::NX;
    X.spv = q; q += (size_t)G::NZ * G::NX;
    X.spw = q; q += (size_t)G::NZ * G::NX;
    X.cxl = q; q += (size_t)G::CL * 2 * G::CL * G::NX;
    X.cxr = q; q += (size_t)G::CL * 2 * G::CL * G::NX;
    X.twN = q; q += 2 * G::NH;
    X.tw2 = q;
    X.thomas_scale = (Real)((h.lz / G::NZ) * (h.lz / G::NZ) / G::NH);
    for (int r = 0; r < G::CL; ++r) {
        std::memcpy(arena.data() + r * L::total + X.o_twN, X.twN, sizeof(Real) * 2 * G::NH);
        std::memcpy(arena.data() + r * L::total + X.o_tw2, X.tw2, sizeof(Real) * 2 * G::NH);
    }
    EnvIO<Real> io{state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, SPLIT ? pressure : nullptr, cell_dist, g_vec};
    RunFlags F{nsub >= 0 ? nsub : C.nsub, project_first, 1};
    SyncState S{};
    for (int e = 0; e < B; ++e) env_action_step<G, Real, NXTG, SPLIT>(C, io, X, e, F, 0, S);
}

#define ARGS *h, wr, cell_dist, B
#define TAIL actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, project_first, nsub
#define DISPATCH2(GT, RT, NG)                                                                     \
    do {                                                                                          \
        if (split) run<GT, RT, NG, true>(ARGS, (RT*)state, TAIL, (RT*)pressure);                   \
        else run<GT, RT, NG, false>(ARGS, (RT*)state, TAIL, (RT*)nullptr);                         \
        return 0;                                                                                 \
    } while (0)
#define DISPATCH(GT)                                                                              \
    do {                                                                                          \
        if (precision == 64 && nxt_global) DISPATCH2(GT, double, true);                           \
        else if (precision == 64) DISPATCH2(GT, double, false);                                   \
        else if (precision == 32 && nxt_global) DISPATCH2(GT, float, true);                       \
        else if (precision == 32) DISPATCH2(GT, float, false);                                    \
        else return -1;                                                                           \
    } while (0)

extern "C" int emu_rbc2dx_step(const HostConfig* h, const HostWrappers* wp, double* cell_dist, int nx, int nz, int cl, int precision,
                               int nxt_global, int B, void* state, const float* actions, float* obs, float* reward, double* nu_state,
                               double* nu_obs, double* t, int* step_count, int* truncated, int* nan_flag, int project_first, int nsub, int split, void* pressure)
{
    HostWrappers wr;
    if (wp) wr = *wp;
    using G96_1 = Grid<96, 64, 1, 4>;
    using G96_2 = Grid<96, 64, 2, 4>;
    using G96_4 = Grid<96, 64, 4, 4>;
    using G192_4 = Grid<192, 128, 4, 2>;
    using G192_8 = Grid<192, 128, 8, 2>;
    using G192_1 = Grid<192, 128, 1, 2>;          // host-only: the whole grid in one "CTA" (plain two-sided Thomas, no SPIKE)
    using G192_2 = Grid<192, 128, 2, 2>;
    using G128_2 = Grid<128, 64, 2, 2, 4>;
    using G64_1 = Grid<64, 64, 1, 4, 4>;           // two of the further registered grids (rbc2dx_more.cuh): the 4 x 8 FFT split on one CTA,
    using G96x128_4 = Grid<96, 128, 4, 4, 6>;      // and four strips per slab with the 6 x 8 split over four CTAs
    if (nx == 96 && nz == 64 && cl == 1) DISPATCH(G96_1);
    if (nx == 96 && nz == 64 && cl == 2) DISPATCH(G96_2);
    if (nx == 96 && nz == 64 && cl == 4) DISPATCH(G96_4);
    if (nx == 192 && nz == 128 && cl == 4) DISPATCH(G192_4);
    if (nx == 192 && nz == 128 && cl == 8) DISPATCH(G192_8);
    if (nx == 192 && nz == 128 && cl == 1) DISPATCH(G192_1);
    if (nx == 192 && nz == 128 && cl == 2) DISPATCH(G192_2);
    if (nx == 128 && nz == 64 && cl == 2) DISPATCH(G128_2);
    if (nx == 64 && nz == 64 && cl == 1) DISPATCH(G64_1);
    if (nx == 96 && nz == 128 && cl == 4) DISPATCH(G96x128_4);
    return -2;
}

extern "C" void emu_rbc2dx_set_vec(int mode, int nan_reset, const double* bank, int n_ep, unsigned long long seed, unsigned long long id_offset,
                                   int* pending, long long* episode, double* ep_return, float* final_obs, double* final_nu_s,
                                   double* final_nu_o, double* final_return, int* nan_count)
{
    g_vec = rbc2d::VecIO();
    g_vec.mode = mode; g_vec.nan_reset = nan_reset; g_vec.bank = bank; g_vec.n_ep = n_ep; g_vec.seed = seed; g_vec.id_offset = id_offset;
    g_vec.pending = pending; g_vec.episode = episode; g_vec.ep_return = ep_return; g_vec.final_obs = final_obs;
    g_vec.final_nu_a = final_nu_s; g_vec.final_nu_b = final_nu_o; g_vec.final_return = final_return; g_vec.nan_count = nan_count;
}
