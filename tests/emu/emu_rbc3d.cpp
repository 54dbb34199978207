// Sequential one-CTA emulator of rbc3d_core.h (test infrastructure; see emu_rbc2d.cpp).
#include <vector>
#include "../../rbc_gym_b200/csrc/rbc3d_core.h"

using namespace rbc3d;

template <typename Real, bool SPLIT, bool TILED = false>
static void run3(const HostConfig3& h, int B, Real* state, const float* actions, float* obs, float* reward, double* nusselt,
                 double* t, int* step_count, int* truncated, int* nan_flag, int project_first, int nsub_override)
{
    Consts3<Real> C = make_consts3<Real>(h);
    std::vector<double> tinv_d((size_t)NZ * NCOL);
    build_tables3_host(h.lx, h.ly, h.lz, tinv_d.data());
    std::vector<Real> tinv(tinv_d.begin(), tinv_d.end());
    std::vector<Real> bufA(NSTATE), bufB(NSTATE), gm(NG), R(NTILE > NR ? NTILE : NR), Tb(NCOL);   // the tile aliases the scratch
    std::vector<double> red(2 * NT);
    EnvIO3<Real> io{state, actions, obs, reward, nusselt, t, step_count, truncated, nan_flag};
    Ctx3<Real> X{bufA.data(), bufB.data(), gm.data(), R.data(), TILED ? R.data() : nullptr, Tb.data(), red.data(), tinv.data(),
                 (Real)((h.lz / NZ) * (h.lz / NZ) / 512.0)};
    RunFlags3 F{nsub_override >= 0 ? nsub_override : C.nsub, project_first, 1};
    for (int e = 0; e < B; ++e) env_action_step3<Real, SPLIT, TILED>(C, io, X, e, F);
}

extern "C" int emu_rbc3d_step(const HostConfig3* h, int precision, int split, int B, void* state, const float* actions, float* obs,
                              float* reward, double* nusselt, double* t, int* step_count, int* truncated, int* nan_flag,
                              int project_first, int nsub_override)
{
    if (split == 2) {      // tiled tendency phase (shared-memory tile per half), buoyancy in G_w
        if (precision == 64) run3<double, false, true>(*h, B, (double*)state, actions, obs, reward, nusselt, t, step_count, truncated, nan_flag, project_first, nsub_override);
        else run3<float, false, true>(*h, B, (float*)state, actions, obs, reward, nusselt, t, step_count, truncated, nan_flag, project_first, nsub_override);
        return 0;
    }
    if (precision == 64) {
        if (split) run3<double, true>(*h, B, (double*)state, actions, obs, reward, nusselt, t, step_count, truncated, nan_flag, project_first, nsub_override);
        else run3<double, false>(*h, B, (double*)state, actions, obs, reward, nusselt, t, step_count, truncated, nan_flag, project_first, nsub_override);
    } else if (precision == 32) {
        if (split) run3<float, true>(*h, B, (float*)state, actions, obs, reward, nusselt, t, step_count, truncated, nan_flag, project_first, nsub_override);
        else run3<float, false>(*h, B, (float*)state, actions, obs, reward, nusselt, t, step_count, truncated, nan_flag, project_first, nsub_override);
    } else return -1;
    return 0;
}
