// Sequential "one CTA" emulator of rbc2d_core.h for machines without a GPU.
// Test infrastructure: lets tests/ check the exact kernel logic (phases, FFT passes, Thomas,
// epilogue reductions) against the fp64 oracle.  Not part of the product path.
#include <vector>
#include <cstring>
#include "../../rbc_gym_b200/csrc/rbc2d_core.h"

using namespace rbc2d;

template <typename Real, bool SPLIT>
static void run(const HostConfig& h, const HostWrappers& wr, double* cell_dist, int B, Real* state, const float* actions, float* obs, float* reward,
                double* nu_state, double* nu_obs, double* t, int* step_count, int* truncated, int* nan_flag,
                Real* pressure, bool nxt_global, const VecIO& vec = VecIO())
{
    Consts<Real> C = make_consts<Real>(h, wr);
    std::vector<double> tinv_d(NZ * NX), tw48_d(96), tw96_d(96);
    build_tables_host(h.lx, h.lz, tinv_d.data(), tw48_d.data(), tw96_d.data());
    std::vector<Real> tinv(tinv_d.begin(), tinv_d.end()), tw48(tw48_d.begin(), tw48_d.end()), tw96(tw96_d.begin(), tw96_d.end());
    Tables<Real> T{tinv.data(), tw48.data(), tw96.data(), (Real)((h.lz / NZ) * (h.lz / NZ) / 48.0)};
    EnvIO<Real> io{state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, pressure, cell_dist, vec};
    std::vector<Real> s0(NS_SM), s1(NS_SM), R(NR), Tb(NX), mid(2 * NX), gm(2 * NSTATE);
    std::vector<double> red(NRED * NT);
    std::vector<Real> E(NE);
    Ctx<Real> X{s0.data(), s1.data(), R.data(), Tb.data(), mid.data(), tw48.data(), tw96.data(), gm.data(), tinv.data(), red.data(), E.data()};
    RunFlags F{C.nsub, 0, 1};
    for (int e = 0; e < B; ++e) {
        if (nxt_global) env_action_step<Real, SPLIT, true>(C, T, io, X, e, F);
        else env_action_step<Real, SPLIT, false>(C, T, io, X, e, F);
    }
}

extern "C" int emu_rbc2d_step(const HostConfig* h, const HostWrappers* wp, double* cell_dist, int precision, int split, int nxt_global, int B, void* state,
                              const float* actions, float* obs, float* reward, double* nu_state, double* nu_obs, double* t,
                              int* step_count, int* truncated, int* nan_flag, void* pressure)
{
    HostWrappers wr;
    if (wp) wr = *wp;
    if (precision == 64) {
        if (split) run<double, true>(*h, wr, cell_dist, B, (double*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (double*)pressure, nxt_global);
        else run<double, false>(*h, wr, cell_dist, B, (double*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (double*)pressure, nxt_global);
    } else if (precision == 32) {
        if (split) run<float, true>(*h, wr, cell_dist, B, (float*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (float*)pressure, nxt_global);
        else run<float, false>(*h, wr, cell_dist, B, (float*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (float*)pressure, nxt_global);
    } else return -1;
    return 0;
}

// the same step with the fused vector-env semantics (auto-reset from a checkpoint bank, episode returns, terminal outputs)
extern "C" int emu_rbc2d_vec_step(const HostConfig* h, int precision, int split, int B, void* state, const float* actions, float* obs,
                                  float* reward, double* nu_state, double* nu_obs, double* t, int* step_count, int* truncated,
                                  int* nan_flag, void* pressure, int mode, int nan_reset, const double* bank, int n_ep,
                                  unsigned long long seed, unsigned long long id_offset, int* pending, long long* episode,
                                  double* ep_return, float* final_obs, double* final_nu_s, double* final_nu_o, double* final_return,
                                  int* nan_count)
{
    VecIO v;
    v.mode = mode; v.nan_reset = nan_reset; v.bank = bank; v.n_ep = n_ep; v.seed = seed; v.id_offset = id_offset;
    v.pending = pending; v.episode = episode; v.ep_return = ep_return; v.final_obs = final_obs; v.final_nu_a = final_nu_s;
    v.final_nu_b = final_nu_o; v.final_return = final_return; v.nan_count = nan_count;
    HostWrappers wr;
    if (precision == 64) {
        if (split) run<double, true>(*h, wr, nullptr, B, (double*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (double*)pressure, true, v);
        else run<double, false>(*h, wr, nullptr, B, (double*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (double*)pressure, true, v);
    } else {
        if (split) run<float, true>(*h, wr, nullptr, B, (float*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (float*)pressure, false, v);
        else run<float, false>(*h, wr, nullptr, B, (float*)state, actions, obs, reward, nu_state, nu_obs, t, step_count, truncated, nan_flag, (float*)pressure, false, v);
    }
    return 0;
}
extern "C" int emu_checkpoint_draw(unsigned long long seed, unsigned long long g, unsigned long long e, int n_ep)
{
    return checkpoint_draw(seed, g, e, n_ep);
}
