/*
 * rbc_b200.h — C ABI of the B200-native Rayleigh-Benard simulation backend.
 *
 * Drop-in boundary for RBC-Gym's simulation step.  In the reference, the Python environments
 * reach the simulation through juliacall: one Julia module of globals per env object
 * (src/rbc_gym/envs/rbc2D.py:111-115) exposing
 *     initialize_simulation   src/rbc_gym/sim/rbc_sim2D_api.jl:17-70
 *     step_simulation         src/rbc_gym/sim/rbc_sim2D_api.jl:75-97
 *     get_state               src/rbc_gym/sim/rbc_sim2D_api.jl:102-118
 *     get_observation         src/rbc_gym/sim/rbc_sim2D_api.jl:123-129
 *     get_info                src/rbc_gym/sim/rbc_sim2D_api.jl:134-137
 *     get_nusselt             src/rbc_gym/sim/rbc_sim2D_api.jl:142-163
 * This library replaces that bridge with a handle-based, *batched* interface: one handle owns
 * B independent environments on one GPU (replacing the process-per-env vector env of
 * example/run_vectorized.py:11-20 / experiments/run_sarl.py:130-153).
 *
 * Conventions
 *   - plain C types only; no torch/CUDA types in signatures (a CUDA stream is passed as void*).
 *   - pointers named *_dev are device pointers on the handle's device, *_host are host pointers
 *     (pinned memory makes the copies asynchronous; pageable memory works too).
 *   - every call returns 0 on success, <0 on error; rbc_last_error() returns the message of the
 *     last failing call on this thread.  NaNs in the fields are *reported* per environment
 *     (nan flag), never raised here: the Python facade turns them into the reference's
 *     RuntimeError (rbc2D.py:170-171).
 *   - all work is enqueued on the handle's stream (rbc2d_set_stream); *_host calls synchronise
 *     that stream before returning, *_dev calls do not.
 *   - a handle is not thread-safe.  There is no CPU fallback: creation fails without a GPU.
 *   - array layouts follow the Python side of the reference (already transposed, rbc2D.py:184-196):
 *     state (C, Nz, Nx) and observation (C, No_z, No_x), z = 0 at the bottom plate, x fastest.
 */
#ifndef RBC_B200_H
#define RBC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBC_B200_ABI_VERSION 2

/* Configuration = the keyword arguments of initialize_simulation (rbc_sim2D_api.jl:17) plus the
 * constants that the Julia API hard-codes (rbc_sim2D_api.jl:28-41), made explicit. */
typedef struct rbc2d_config {
    int32_t num_envs;        /* B: environments in the on-device batch                                 */
    int32_t nx, nz;          /* grid = state_shape[::-1]; registered: 96x64, 128x64, 192x128, 64x32, 64x64, 96x32, 96x128, 128x32, 128x128, 192x64 */
    int32_t obs_nx, obs_nz;  /* sensors = observation_shape[::-1]; must divide nx, nz (48, 8)          */
    int32_t heaters;         /* heater segments (12), <= 32                                            */
    double heater_limit;     /* 0.75                                                                   */
    double ra;               /* Rayleigh number                                                        */
    double pr;               /* Prandtl number, reference hard-codes 0.7                               */
    double dt_action;        /* heater_duration: simulated time per action step                        */
    double dt_solver;        /* RK3 step, reference hard-codes 0.03                                    */
    double episode_length;   /* truncation time (300)                                                  */
    int32_t precision;       /* 32 = throughput mode, 64 = validation mode                             */
    int32_t pressure;        /* 1: hydrostatic-pressure split scheme + pHY', pNHS channels (pressure=True) */
    int32_t device;          /* CUDA device ordinal                                                    */
} rbc2d_config;

/* Wrappers fused into the step epilogue (src/rbc_gym/wrappers/), applied in the order of
 * example/run_wrapped.py:15-19: RBCNormalizeObservation -> RBCNormalizeReward -> RBCRewardShaping. */
typedef struct rbc2d_wrappers {
    int32_t normalize_obs;   /* rbc_normalize_observation.py:64-74: maxval*(2(obs-lo)/(hi-lo)-1) on channels 0..3   */
    int32_t obs_clip;        /* clip to [-maxval, maxval]                                                        */
    float obs_lo[4];         /* [1, -u_limit, -u_limit, -u_limit]                                                */
    float obs_hi[4];         /* [2 + heater_limit, u_limit, u_limit, u_limit]                                    */
    float obs_maxval;
    int32_t normalize_reward; /* rbc_normalize_reward.py:27-32: (r + s)/(s - 1)                                   */
    double reward_scale;     /* s = 0.1 * Ra^0.4 in 2D                                                           */
    int32_t shaping;         /* rbc_reward_shaping.py:53-140: r <- (1-w) r + w (pi - cell_dist)/pi               */
    double shaping_weight;
} rbc2d_wrappers;

typedef struct rbc2d_sim rbc2d_sim;

int rbc_abi_version(void);
const char* rbc_last_error(void);

/* initialize_simulation, configuration part: allocates state, scratch and tables on the device. */
int rbc2d_create(const rbc2d_config* cfg, rbc2d_sim** out);
int rbc2d_destroy(rbc2d_sim* sim);
int rbc2d_set_stream(rbc2d_sim* sim, void* cuda_stream);
int rbc2d_num_envs(const rbc2d_sim* sim);
int rbc2d_state_values_per_env(const rbc2d_sim* sim);   /* 2*nx*nz + nx*(nz+1) */

/* Checkpoint bank (data/checkpoints/.../ckpt_ra*.h5 read on the host; initialize_from_checkpoint,
 * rbc_sim2D.jl:173-186).  Arrays are [n_episodes][nz|nz+1][nx] float64, z = 0 bottom. */
int rbc2d_load_checkpoints(rbc2d_sim* sim, const double* b_host, const double* u_host, const double* w_host,
                           int32_t n_episodes);

/* Reset environments from the bank: initialize_simulation with checkpoint_path.  env_ids_dev lists
 * the n environments to reset (NULL = all B, then n is ignored); ckpt_idx_dev gives the episode
 * index per listed environment.  t := 0, step := 1, the observation outputs are refreshed. */
int rbc2d_reset_from_checkpoints_dev(rbc2d_sim* sim, const int32_t* env_ids_dev, const int32_t* ckpt_idx_dev, int32_t n);

/* Reset environments from explicit fields (noise initialisation, rbc_sim2D.jl:163-171, or any
 * state).  fields_host is [n][2*nx*nz + nx*(nz+1)] float64 in checkpoint layout (b, u, w).
 * project != 0 applies the pressure projection Oceananigans' set! performs.  env_ids_host lists
 * the environments (NULL = the first n). */
int rbc2d_reset_from_fields_host(rbc2d_sim* sim, const int32_t* env_ids_host, const double* fields_host, int32_t n,
                                 int32_t project);
/* Same with everything already on the device (e.g. noise fields drawn by a device generator): no host round trip, no
 * synchronisation.  env_ids_dev may be NULL (= the first n environments). */
int rbc2d_reset_from_fields_dev(rbc2d_sim* sim, const int32_t* env_ids_dev, const double* fields_dev, int32_t n, int32_t project);

/* step_simulation for the whole batch + get_observation + get_nusselt + truncation, fused.
 *   actions      [B][heaters] float32 in [-1,1]
 *   obs          [B][C][obs_nz][obs_nx] float32, C = 3 (b,u,w) or 5 with pressure
 *   reward       [B] float32 = -nusselt_obs          (rbc2D.py:198-200)
 *   nu_state/obs [B] float64                         (info["nusselt_state"], info["nusselt_obs"])
 *   truncated    [B] int32: t >= episode_length      (rbc2D.py:178-180)
 *   nan          [B] int32: step_contains_NaNs       (rbc_sim2D.jl:223-228)
 * Any output pointer may be NULL except obs.
 * The _host variant copies the actions in and every requested output out inside the call and returns when they have
 * landed; the batch runs in wave-aligned chunks whose outputs are copied on a second stream while the next chunk
 * computes (page-locked host buffers make that overlap effective; pageable ones work, serialised by the driver). */
int rbc2d_step_dev(rbc2d_sim* sim, const float* actions_dev, float* obs_dev, float* reward_dev, double* nu_state_dev,
                   double* nu_obs_dev, int32_t* truncated_dev, int32_t* nan_dev);
int rbc2d_step_host(rbc2d_sim* sim, const float* actions_host, float* obs_host, float* reward_host, double* nu_state_host,
                    double* nu_obs_host, int32_t* truncated_host, int32_t* nan_host);

/* ------------------------------------------------------------------------------------------------
 * Vector-environment step: what the reference gets from gymnasium's / SB3's vector wrappers around one process per
 * env (gym.make_vec(..., vectorization_mode="async"), example/run_vectorized.py:11-31; SubprocVecEnv,
 * experiments/run_sarl.py:130-153) — auto-reset of truncated environments, episode returns, terminal observations —
 * fused into the ONE kernel launch that steps the batch.  Resets gather a uniformly drawn episode of the checkpoint
 * bank (rbc_sim2D.jl:176-177); the draw is rbc_checkpoint_draw(seed, env_id_offset + env, episode counter), so it does
 * not depend on how environments are sharded over GPUs.
 *   mode 0  disabled: flags and episode returns only
 *   mode 1  next_step (gymnasium 1.1.1 default): an environment that truncated is re-initialised by the NEXT call —
 *           that call ignores its action, does not march it, returns the reset observation, reward 0, truncated 0
 *   mode 2  same_step (SB3): re-initialised inside the truncating call; obs is the reset observation, the terminal
 *           observation / Nusselt numbers / episode return go to the final_* outputs (rows of truncated envs only)
 *   nan_reset != 0: an environment whose fields contain NaNs is re-initialised inside the call in either mode, reported
 *           truncated with reward 0 (the reference raises instead, rbc2D.py:170-171; the count of such events is kept
 *           in a device counter the caller can poll without synchronising the step)
 * Without a checkpoint bank (noise initialisation) nothing is reset in the kernel: the flags are reported and the caller
 * resets with rbc2d_reset_from_fields_dev.
 * ------------------------------------------------------------------------------------------------ */
typedef struct rbc_autoreset {
    int32_t mode;
    int32_t nan_reset;
    int64_t seed;
    int64_t env_id_offset;
} rbc_autoreset;

/* Outputs of a vector step; device pointers, any may be NULL except obs (final_* are only written in mode 2 / NaN resets). */
typedef struct rbc2d_vec_out {
    float* obs;              /* [B][C][obs_nz][obs_nx]                                   */
    float* reward;           /* [B]                                                      */
    double* nu_state;        /* [B] info["nusselt_state"]                                */
    double* nu_obs;          /* [B] info["nusselt_obs"]                                  */
    int32_t* truncated;      /* [B]                                                      */
    int32_t* nan;            /* [B] step_contains_NaNs of this call                      */
    double* t;               /* [B] info["t"]   (rbc2D.py:202-212)                       */
    int32_t* step;           /* [B] info["step"]                                         */
    double* episode_return;  /* [B] running return of the current episode                */
    float* final_obs;        /* [B][C][obs_nz][obs_nx]                                   */
    double* final_nu_state;  /* [B]                                                      */
    double* final_nu_obs;    /* [B]                                                      */
    double* final_return;    /* [B] return of the episode that just ended                */
} rbc2d_vec_out;

int32_t rbc_checkpoint_draw(int64_t seed, int64_t global_env, int64_t episode, int32_t n_episodes);
int rbc2d_set_autoreset(rbc2d_sim* sim, const rbc_autoreset* cfg);
/* reset(): every environment starts episode 0 from bank[ckpt_idx_dev[env]] (NULL: from its own draw); returns and pending
 * flags are cleared.  Needs a checkpoint bank. */
int rbc2d_vec_reset_dev(rbc2d_sim* sim, const int32_t* ckpt_idx_dev);
/* Counts one episode start for every environment without touching the fields (after the caller re-initialised them itself,
 * e.g. the noise initialisation through rbc2d_reset_from_fields_dev): episode counter += 1 for the listed environments
 * (NULL = all), return := 0, pending := 0. */
int rbc2d_vec_mark_reset_dev(rbc2d_sim* sim, const int32_t* env_ids_dev, int32_t n);
int rbc2d_vec_step_dev(rbc2d_sim* sim, const float* actions_dev, const rbc2d_vec_out* out);
/* The same step through HOST buffers (every pointer of `out` is a host pointer, any may be NULL): actions are copied in,
 * the requested outputs out, chunk by chunk on a second stream while the next chunk computes; returns when they have landed.
 * This is the call a caller without device tensors makes once per rollout step — what replaces the pickled pipe traffic of
 * the process-per-env vector env. */
int rbc2d_vec_step_host(rbc2d_sim* sim, const float* actions_host, const rbc2d_vec_out* out_host);
/* Number of environments that reported NaNs in vector steps since the last call with clear != 0 (synchronises the stream). */
int rbc2d_vec_nan_count(rbc2d_sim* sim, int32_t clear, int64_t* count);
/* Same counter copied to a caller-provided (pinned) host int32 without synchronising: the value is valid once the stream
 * has passed this point (poll it one step late for a deferred, sync-free NaN check). */
int rbc2d_vec_nan_count_async(rbc2d_sim* sim, int32_t* count_host_pinned);

/* CFL guard of the grids beyond 96 x 64 (cluster kernels).  The reference steps with a fixed dt_solver; on the 192 x 128 grid a
 * flow started from noise at Ra = 1e6 overshoots to CFL ~ 1.6 in its first plume burst, where the RK3 / upwind scheme is
 * linearly unstable: fp64 rides it out, fp32 loses about one environment in a thousand to NaNs.  With the guard an environment
 * whose max(|w| dt/dz, |u| dt/dx) exceeds `limit` at the start of an RK3 step takes that step as ceil(CFL) equal parts; all other
 * environments are untouched bit for bit.  Default: limit 1.4 in the fp32 mode, off (0) in the fp64 validation mode.
 * rbc2d_get_cfl_events_host: extra RK3 steps inserted per environment since creation, [B] int32. */
int rbc2d_set_cfl_guard(rbc2d_sim* sim, double limit);
int rbc2d_get_cfl_events_host(rbc2d_sim* sim, int32_t* out_host);

/* Enable/disable the fused wrappers (NULL = all off).  Takes effect from the next step/observe. */
int rbc2d_set_wrappers(rbc2d_sim* sim, const rbc2d_wrappers* w);
/* info["cell_dist"] of the last step (rbc_reward_shaping.py:61-66), [B] float64; computed only while
 * shaping is enabled (use shaping_weight = 0 to get the diagnostic without changing the reward). */
int rbc2d_get_cell_dist_host(rbc2d_sim* sim, double* out_host);

/* get_observation / get_nusselt without stepping (what reset() returns). */
int rbc2d_observe_dev(rbc2d_sim* sim, float* obs_dev, double* nu_state_dev, double* nu_obs_dev);
int rbc2d_observe_host(rbc2d_sim* sim, float* obs_host, double* nu_state_host, double* nu_obs_host);

/* get_state: [B][C][nz][nx] float32 in the Python layout, channels b,u,w(+pHY',pNHS). */
int rbc2d_get_state_dev(rbc2d_sim* sim, float* out_dev, int32_t channels);
int rbc2d_get_state_host(rbc2d_sim* sim, float* out_host, int32_t channels);

/* render("rgb_array") for the whole batch on the device (rbc2D.py:214-261 renders one env on the host with matplotlib's
 * turbo colormap): temperature channel -> turbo RGB, vmin = 1, vmax = 2 + heater_limit, origin at the top left like the
 * reference image.  out_dev: [B][nz][nx][3] uint8.  The colormap is the 7-term polynomial fit of turbo (|error| <= 1/255). */
int rbc2d_render_rgb_dev(rbc2d_sim* sim, uint8_t* out_dev);

/* Raw fields in checkpoint layout [B][2*nx*nz + nx*(nz+1)] float64 (for parity tests / checkpoint writing). */
int rbc2d_get_fields_host(rbc2d_sim* sim, double* fields_host);

/* get_info: simulation time and step counter per environment (rbc_sim2D_api.jl:134-137). */
int rbc2d_get_info_host(rbc2d_sim* sim, double* t_host, int32_t* step_host);

/* Introspection for benchmarks: kernels launched so far, CTAs per step launch, dynamic smem bytes. */
int rbc2d_launch_count(const rbc2d_sim* sim, int64_t* launches, int32_t* grid, int32_t* smem_bytes);
/* Device time of the most recent step kernel in milliseconds (CUDA events on the handle's stream;
 * synchronises). */
int rbc2d_last_step_kernel_ms(rbc2d_sim* sim, float* ms);
/* Device times of the last n step kernels (n <= 64), oldest first, from the same per-launch CUDA events; returns the number
 * written (<= n) or <0.  Lets a benchmark average over every launch of its timed region without synchronising inside it. */
int rbc2d_step_kernel_ms_history(rbc2d_sim* sim, float* ms_out, int32_t n);

/* ------------------------------------------------------------------------------------------------
 * 3D environment (rbc_gym/RayleighBenardConvection3D-v0): replaces the juliacall functions of
 * src/rbc_gym/sim/rbc_sim3D_api.jl — initialize_simulation :17-72, step_simulation :77-101,
 * get_state :106-121, get_info :126-129, get_nusselt :134-159, shutdown_simulation :164-171 —
 * called from src/rbc_gym/envs/rbc3D.py:145-245.  Same conventions as the 2D entry points.
 *   state / observation  [B][4][nz][ny][nx] float32 = (b,u,v,w) in the Python layout (rbc3D.py:229-232)
 *   action               [B][heaters][heaters] float32, action[i][j] <-> patch i along x, j along y
 *                        (rbc3D.py:206 passes the array un-transposed; rbc_sim3D.jl:131-141)
 *   fields (raw)         [B][3*nx*ny*nz + nx*ny*(nz+1)] float64: b,u,v [nz][ny][nx], w [nz+1][ny][nx]
 * ------------------------------------------------------------------------------------------------ */
typedef struct rbc3d_config {
    int32_t num_envs;
    int32_t nx, ny, nz;        /* grid = state_shape[::-1] (a free kwarg of the reference env, rbc3D.py:43-60): 32 x 32 x 16 runs a
                                  dedicated one-CTA-per-environment kernel, any other grid with nx, ny powers of two in 8..256 and
                                  6 <= nz <= 256 (e.g. the 64 x 64 x 32 of experiments/flowstats/flowstats_ra.py:27-36) the
                                  stage-streaming kernels                                                   */
    int32_t heaters;           /* patches per side (8)                                                  */
    double heater_limit;       /* 0.9                                                                   */
    double ra, pr;
    double lx, ly, lz;         /* L = domain[::-1] = [4 pi, 4 pi, 2]                                    */
    double b_min, b_max;       /* T_diff = [1, 2]                                                       */
    double heater_duration;    /* dt, in free-fall units (0.125)                                        */
    double dt_solver;          /* in free-fall units (0.01); simulation dt = dt_solver * lz^2           */
    double episode_length;     /* compared with the simulation time, which advances heater_duration*lz^2 */
    int32_t precision;         /* 32 | 64                                                               */
    int32_t split;             /* 1: hydrostatic-pressure split like Oceananigans, 0: buoyancy in G_w   */
    int32_t device;
} rbc3d_config;

typedef struct rbc3d_sim rbc3d_sim;

int rbc3d_create(const rbc3d_config* cfg, rbc3d_sim** out);
int rbc3d_destroy(rbc3d_sim* sim);                          /* shutdown_simulation */
int rbc3d_set_stream(rbc3d_sim* sim, void* cuda_stream);
int rbc3d_state_values_per_env(const rbc3d_sim* sim);
/* checkpoint bank [n_episodes][values_per_env] float64 (3D_ckpt_ra*.h5 read on the host; rbc_sim3D.jl:181-199) */
int rbc3d_load_checkpoints(rbc3d_sim* sim, const double* fields_host, int32_t n_episodes);
int rbc3d_reset_from_checkpoints_dev(rbc3d_sim* sim, const int32_t* env_ids_dev, const int32_t* ckpt_idx_dev, int32_t n);
int rbc3d_reset_from_fields_host(rbc3d_sim* sim, const int32_t* env_ids_host, const double* fields_host, int32_t n, int32_t project);
int rbc3d_reset_from_fields_dev(rbc3d_sim* sim, const int32_t* env_ids_dev, const double* fields_dev, int32_t n, int32_t project);
/* step_simulation + get_state + get_nusselt fused; obs may be NULL (skips the 262 KB/env observation write) */
int rbc3d_step_dev(rbc3d_sim* sim, const float* actions_dev, float* obs_dev, float* reward_dev, double* nusselt_dev,
                   int32_t* truncated_dev, int32_t* nan_dev);
int rbc3d_step_host(rbc3d_sim* sim, const float* actions_host, float* obs_host, float* reward_host, double* nusselt_host,
                    int32_t* truncated_host, int32_t* nan_host);
/* Vector-environment step of the 3D batch: same semantics as rbc2d_vec_step_dev (what SubprocVecEnv + Monitor give
 * experiments/run_sarl.py:130-153).  final_obs is [B][4][nz][ny][nx] — pass NULL to skip the terminal observations. */
typedef struct rbc3d_vec_out {
    float* obs;              /* [B][4][nz][ny][nx] or NULL                               */
    float* reward;           /* [B]                                                      */
    double* nusselt;         /* [B] info["nusselt"]                                      */
    int32_t* truncated;
    int32_t* nan;
    double* t;               /* [B] info["t"] (rbc3D.py:241-245)                         */
    int32_t* step;           /* [B] info["step"]                                         */
    double* episode_return;
    float* final_obs;
    double* final_nusselt;
    double* final_return;
} rbc3d_vec_out;
/* One Rayleigh number per environment (host array [B]; NULL restores cfg.ra): lets one batch sweep the Rayleigh number the way
 * experiments/flowstats/flowstats_ra.py:27-36 loops over 14 environments.  Stage-streaming kernels only. */
int rbc3d_set_rayleigh_per_env(rbc3d_sim* sim, const double* ra_host);
int rbc3d_set_autoreset(rbc3d_sim* sim, const rbc_autoreset* cfg);
int rbc3d_vec_reset_dev(rbc3d_sim* sim, const int32_t* ckpt_idx_dev);
int rbc3d_vec_mark_reset_dev(rbc3d_sim* sim, const int32_t* env_ids_dev, int32_t n);
int rbc3d_vec_step_dev(rbc3d_sim* sim, const float* actions_dev, const rbc3d_vec_out* out);
int rbc3d_vec_nan_count(rbc3d_sim* sim, int32_t clear, int64_t* count);
int rbc3d_vec_nan_count_async(rbc3d_sim* sim, int32_t* count_host_pinned);
int rbc3d_observe_dev(rbc3d_sim* sim, float* obs_dev, double* nusselt_dev);
/* render("rgb_array") for the whole batch on the device (rbc3D.py:247-318: PyVista volume rendering of the temperature, turbo
 * colormap, clim = temperature_difference, opacity "sigmoid_1", 800 x 608, isometric camera; used by example/run_wandb.py:25-59):
 * the same picture ray-marched with one thread per pixel.  out_dev: [B][height][width][3] uint8. */
int rbc3d_render_rgb_dev(rbc3d_sim* sim, uint8_t* out_dev, int32_t height, int32_t width);
int rbc3d_get_fields_host(rbc3d_sim* sim, double* fields_host);
int rbc3d_get_info_host(rbc3d_sim* sim, double* t_host, int32_t* step_host);
int rbc3d_launch_count(const rbc3d_sim* sim, int64_t* launches, int32_t* grid, int32_t* smem_bytes);
int rbc3d_last_step_kernel_ms(rbc3d_sim* sim, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* RBC_B200_H */
