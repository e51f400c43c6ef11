/*
 * heli3dof.h -- C ABI of libheli3dof.so: batched, B200-native (sm_100a) implementation of the
 * 3-DOF closed-loop control tick of BartHelder/msc-thesis.
 *
 * The reference has no FFI layer: its boundary for this path is the Python object API that
 * main_3dof.py calls.  Each entry point below therefore cites the reference METHOD it replaces
 * (file:line in the reference tree); msc-thesis_b200/*.py re-creates those classes on top of
 * this ABI through ctypes.
 *
 * Conventions
 *   - every per-episode array is a caller-owned DEVICE buffer in struct-of-arrays layout
 *     `[field][ld]` (element f of episode e at base[f*ld + e]); `ld >= n`, 16-byte aligned base;
 *   - mutable real-valued buffers are `void*`: their element type is given by the dtype codes of the
 *     call (H3_F64 = double, H3_F32 = float).  env_dtype covers the model state, reference, PID and RLS
 *     buffers, net_dtype the actor/critic/target-critic weights.  Time, accumulators, statistics and
 *     the read-only per-episode parameter rows (trim, theta_trim, ref_par, *_hyper) are always double;
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), allocate nothing,
 *     keep no global state, and are re-entrant for disjoint buffers;
 *   - return value: 0 = ok, negative = H3_E_* (never throws); h3_last_error() gives the text.
 *     Numerical failure is DATA, not an error: per-episode `flags` bit0 = |pitch| > 90 deg
 *     (heli_models.py:210-213), bit1 = NaN in actions (main_3dof.py:192-194); `term_step` is the
 *     tick index at which the episode stopped, -1 while it is alive.
 */
#ifndef HELI3DOF_H
#define HELI3DOF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H3_VERSION 100

#define H3_F64 0
#define H3_F32 1

#define H3_OK 0
#define H3_E_ARG (-1)       /* bad argument (null pointer, unsupported size, ...) */
#define H3_E_UNSUPPORTED (-2) /* combination not compiled in (e.g. hidden-layer width) */
#define H3_E_CUDA (-3)      /* CUDA launch / runtime error */

#define H3_FLAG_DONE 1
#define H3_FLAG_NAN 2

#define H3_TASK_SINUSOID 0        /* heli_models.py:233-237 */
#define H3_TASK_VELOCITY 1        /* heli_models.py:239-243 */
#define H3_TASK_STOP_OVER_POINT 2 /* heli_models.py:245-250 */

#define H3_REF_TABLE 0   /* sinusoid reference staged from a host-made table (shared clock) */
#define H3_REF_KERNEL 1  /* reference evaluated per episode in the kernel (own t, amplitude, period) */

#define H3_ACT_TANH 0    /* agents.py:52-53 */
#define H3_ACT_SIGMOID 1 /* agents.py:54-55 */

#define H3_N_STATE 7
#define H3_N_ACTION 2
#define H3_N_REG 9       /* RLS regressors: 7 state + 2 action increments (constant=False) */
#define H3_N_STATS 8

/* Physical constants of Helicopter3DOF.__init__ (heli_models.py:117-130). */
typedef struct H3EnvParams {
    double g, cl_alpha, blade_solidity, lock_number, cds, mass, rho, v_tip, rotor_radius, iy, mast,
        omega, area, tau;
    double dt;
} H3EnvParams;

/* Scalar hyper-parameters of one DHPAgent (agents.py:74-115); any of them can be overridden per
 * episode through H3Batch.{col,lon}_hyper. */
typedef struct H3AgentHyper {
    double gamma;      /* discount_factor */
    double lr_actor;   /* learning_rate_actor */
    double lr_critic;  /* learning_rate_critic */
    double tau;        /* tau_target_critic */
    double scale;      /* deg2rad(action_scaling), agents.py:46-47 */
    double reward_weight;
    int32_t n_hidden;  /* n_hidden_actor == n_hidden_critic */
    int32_t tracked_state;  /* index into the 7-state (1 = z or 4 = theta in the 3-DOF drivers) */
    int32_t ac_state;       /* ac_states[0] */
    int32_t channel;        /* control_channel_num: 0 col, 1 lon */
    int32_t final_activation; /* H3_ACT_* */
    int32_t pad_;
} H3AgentHyper;

/* Launch-uniform configuration of the tick (main_3dof.py:20-71). */
typedef struct H3Config {
    H3EnvParams env;
    int32_t env_dtype;       /* H3_F64 | H3_F32 */
    int32_t net_dtype;       /* H3_F64 | H3_F32 */
    int32_t task;            /* H3_TASK_* */
    int32_t ref_mode;        /* H3_REF_* (sinusoid only) */
    double env_kp, env_ki, env_kd, env_ref;      /* config_3dof.json env.tasks.<task> */
    double ref_amp, ref_period;                  /* heli_models.py:220-221: 10 deg, 10 s */
    double pid_kp, pid_ki, pid_kd, pid_href;     /* PID.py:6-14, main_3dof.py:68 */
    double rls_gamma;                            /* model.py:14 */
    int32_t rls_symmetric;   /* 1: covariance known symmetric, kernel keeps the upper triangle on chip */
    int32_t store_policy;    /* 0: default placement; 3: keep the collective agent in shared memory (no tensor memory) */
    int32_t switch_step;     /* tick at which the collective agent takes over (main_3dof.py:127-137) */
    int32_t update_lon;      /* main_3dof.py:87 (always 1 there) */
    int32_t win_lo, win_hi;  /* score window [lo, hi) on the lon reward, train.py:256 */
    H3AgentHyper col, lon;
} H3Config;

/* Per-episode hyper-parameter override rows (H3Batch.*_hyper is [H3_HYPER_ROWS][ld], f64). */
#define H3_HYPER_GAMMA 0
#define H3_HYPER_LR_ACTOR 1
#define H3_HYPER_LR_CRITIC 2
#define H3_HYPER_TAU 3
#define H3_HYPER_SCALE 4
#define H3_HYPER_REWARD_WEIGHT 5
#define H3_HYPER_ROWS 6

/* Network block of one agent, `[11*nh][ld]` in net dtype (torch (out,in) row-major inside a matrix):
 *   rows [0,2nh)    actor fc1  (nh,2)     agents.py:42
 *   rows [2nh,3nh)  actor fc2  (1,nh)     agents.py:44
 *   rows [3nh,5nh)  critic fc1 (nh,2)     agents.py:19
 *   rows [5nh,7nh)  critic fc2 (2,nh)     agents.py:21
 *   rows [7nh,9nh)  target critic fc1, rows [9nh,11nh) target critic fc2   agents.py:107 */

typedef struct H3Batch {
    int64_t n;   /* episodes */
    int64_t ld;  /* field stride (elements) */
    /* environment (Helicopter3DOF) */
    void* state;            /* [7][ld]  x z u w theta q lambda_i      heli_models.py:131 */
    double* t;              /* [ld]     env.t (always f64)             heli_models.py:206 */
    void* ref;              /* [2][ld]  current reference, slots 1 (z) and 4 (theta)  main_3dof.py:62 */
    void* corr_u;           /* [ld]     velocity-task integrator       heli_models.py:242 */
    const double* trim;       /* [2][ld]  trimmed [collective, cyclic]   heli_models.py:328 */
    const double* theta_trim; /* [ld]     trimmed_state[4]               heli_models.py:249 */
    const double* ref_par;    /* [2][ld] or NULL: per-episode (amplitude, period) for H3_REF_KERNEL */
    /* collective PID (CollectivePID3DOF) */
    void* pid_corr;         /* [ld]     hdot_corr                      PID.py:23 */
    /* incremental model (RecursiveLeastSquares) */
    void* rls_W;            /* [9*7][ld]  W[i][j] at row i*7+j         model.py:84 */
    void* rls_cov;          /* [9*9][ld]                               model.py:85 */
    void* rls_x;            /* [7][ld]                                 model.py:87 */
    void* rls_u;            /* [2][ld]                                 model.py:88 */
    int32_t* rls_skip;      /* [ld]                                    model.py:89 */
    /* agents (DHPAgent) */
    void* col_net;          /* [11*nh_col][ld] */
    void* lon_net;          /* [11*nh_lon][ld] */
    const double* col_hyper;  /* [H3_HYPER_ROWS][ld] or NULL */
    const double* lon_hyper;  /* [H3_HYPER_ROWS][ld] or NULL */
    /* bookkeeping (bit-exact integer fields) */
    int32_t* step;          /* [ld] next tick index                    main_3dof.py:122 */
    int32_t* flags;         /* [ld] H3_FLAG_* */
    int32_t* term_step;     /* [ld] */
    /* per-episode accumulators (f64): [0] sum r_col, [1] sum r_lon, [2] sum r_lon inside the window */
    double* acc;            /* [3][ld] */
    /* optional OUTPUT of h3_closed_loop (NULL to skip), kept for ABI stability: earlier builds counted here the ticks
     * an episode spent on data-dependent slow branches of tanh / sincos (a sort key for the host).  The tick has no such
     * branch any more; the kernels write 0. */
    int32_t* heat;          /* [ld] */
} H3Batch;

/* Optional per-tick trajectory output of h3_closed_loop (what main_3dof.py:164-175 logs):
 * row layout [tick][13][ld_traj]: state(7, pre-step), ref_theta, collective, cyclic, r_col, r_lon, alive.
 * Recorded for the first n_traj episodes only. */
typedef struct H3Traj {
    void* out;        /* env dtype, [n_ticks][13][ld_traj] or NULL */
    int64_t n_traj;
    int64_t ld_traj;
} H3Traj;
#define H3_TRAJ_ROWS 13

int h3_version(void);
const char* h3_last_error(void);

/* Helicopter3DOF.__init__ constants (heli_models.py:111-130). */
void h3_env_params_default(H3EnvParams* out, double dt);

/* Helicopter3DOF._trim(v) (heli_models.py:293-328), batched: v[n] -> trim[2][ld], state[7][ld]. f64 only. */
int h3_trim(const H3EnvParams* env, const double* v, double* trim, double* state, int64_t n, int64_t ld,
            void* stream);

/* Helicopter3DOF.step (heli_models.py:142-215) for n_ticks ticks, open loop:
 * action(k) = actions[2][ld] (if non-NULL) else trim[2][ld], plus excitation[k][2] for step k < exc_len.
 * Updates state, t, step, flags, term_step of H3Batch in place; other H3Batch fields may be NULL.
 * states_out (optional, env dtype, [n_ticks][7][ld_out]): post-step state of the first n_out episodes. */
int h3_env_step(const H3Config* cfg, const H3Batch* b, const void* actions, const double* excitation,
                int32_t exc_len, int32_t n_ticks, void* states_out, int64_t n_out, int64_t ld_out,
                void* stream);

/* Helicopter3DOF.get_ref (heli_models.py:217-255): writes b->ref from (t + dt, state), with the
 * velocity-task integrator side effect. */
int h3_get_ref(const H3Config* cfg, const H3Batch* b, void* stream);

/* CollectivePID3DOF.__call__ (PID.py:16-24): collective[ld] from b->state, updates b->pid_corr. */
int h3_pid_collective(const H3Config* cfg, const H3Batch* b, void* collective_out, void* stream);

/* RecursiveLeastSquares.update (model.py:27-59): state/action/next_state are [7|2|7][ld] env dtype. */
int h3_rls_update(const H3Config* cfg, const H3Batch* b, const void* state, const void* action,
                  const void* next_state, void* stream);

/* RecursiveLeastSquares.predict (model.py:61-69): out[7][ld]. */
int h3_rls_predict(const H3Config* cfg, const H3Batch* b, const void* state, const void* action, void* out,
                   void* stream);

/* DHPAgent.get_action (agents.py:117-121): which = 0 col, 1 lon; state[7][ld], ref[2][ld]; out[ld] (net dtype). */
int h3_agent_action(const H3Config* cfg, const H3Batch* b, int32_t which, const void* state, const void* ref,
                    void* action_out, void* stream);

/* DHPAgent.get_reward (agents.py:123-128): reward_out[ld], dr_ds_out[2][ld] (env / net dtype). */
int h3_agent_reward(const H3Config* cfg, const H3Batch* b, int32_t which, const void* next_state, const void* ref,
                    void* reward_out, void* dr_ds_out, void* stream);

/* DHPAgent.get_transition_matrices (agents.py:130-140) from b->rls_W: F_out[4][ld] (row-major 2x2),
 * G_out[2][ld], both in net dtype (the reference casts to torch.float at :140). */
int h3_agent_transition(const H3Config* cfg, const H3Batch* b, int32_t which, void* F_out, void* G_out,
                        void* stream);

/* DHPAgent.update_networks(state, next_state, ref, next_ref, dr_ds, F, G) (agents.py:152-196):
 * state/next_state [7][ld], ref/next_ref [2][ld] (env dtype); dr_ds [2][ld], F [4][ld], G [2][ld] (net dtype). */
int h3_agent_update(const H3Config* cfg, const H3Batch* b, int32_t which, const void* state,
                    const void* next_state, const void* ref, const void* next_ref, const void* dr_ds,
                    const void* F, const void* G, void* stream);

/* The fused closed-loop tick, main_3dof.py:122-197, n_ticks ticks per launch with every episode's
 * state resident on chip.  step0 >= 0: lockstep launch -- every live episode is at tick step0 (shared
 * clock; reference table and excitation window are staged through shared memory); step0 < 0: every
 * episode runs from its own b->step (teacher-forced / heterogeneous batches; H3_REF_TABLE not allowed).
 * ref_table: host-made theta_ref for H3_REF_TABLE, entry j is the reference
 * that becomes current AFTER local tick j (DEVICE pointer, n_ticks doubles) -- NULL for H3_REF_KERNEL
 * or state-dependent tasks.  excitation: DEVICE [exc_len][2] doubles or NULL (added while step < exc_len,
 * main_3dof.py:139-140 / train.py:173-174).  stats: DEVICE [H3_N_STATS] doubles (plus scratch, see
 * h3_stats_scratch_bytes) receiving this batch's sums {n, n_alive, n_done, n_nan, sum r_col, sum r_lon,
 * sum window r_lon, sum ticks} -- the vector that is all-reduced across GPUs; NULL to skip. */
int h3_closed_loop(const H3Config* cfg, const H3Batch* b, int32_t step0, const double* ref_table,
                   const double* excitation, int32_t exc_len, int32_t n_ticks, const H3Traj* traj,
                   double* stats, void* stats_scratch, void* stream);

/* Bytes of device scratch h3_closed_loop needs next to `stats` for a batch of n episodes. */
int64_t h3_stats_scratch_bytes(int64_t n);

/* Occupancy facts of the fused kernel chosen for cfg (for bench / DESIGN.md): threads per block,
 * dynamic shared memory per block, resident blocks per SM, registers per thread, episodes per warp. */
int h3_closed_loop_info(const H3Config* cfg, int32_t phase2, int32_t* block, int32_t* smem_bytes,
                        int32_t* blocks_per_sm, int32_t* regs, int32_t* lanes_per_warp);

/* The kernels' own fp64 elementary functions, exposed for accuracy tests: fn 0: y = tanh(x), table-driven (3-DOF
 * kernels); fn 1: y = sin(x), y2 = cos(x); fn 2: y = tanh(x), table-free (6-DOF per-tick kernels); fn 3: y = atan2(x, y2),
 * fn 4: y = x / y2 (y2 is an INPUT for these two); fn 5: y = sqrt(x).  DEVICE pointers, n elements. */
int h3_math_probe(int32_t fn, const double* x, double* y, double* y2, int64_t n, void* stream);

/* Roofline denominators measured with dependent-free FMA chains: runs `iters` x 8 independent
 * FMA chains per thread on a full grid; flops = 2 * fma count is written to *flops_out (host). The caller
 * times it with CUDA events on `stream`. dtype: H3_F64 | H3_F32. */
int h3_fma_peak(int32_t dtype, int32_t iters, void* scratch_out, double* flops_out, void* stream);

/* Episode initialisation on the device: everything main_3dof.py does before its tick loop (:58-71, :122), per
 * episode, from a seed and the episode's GLOBAL index -- Helicopter3DOF.reset(v) incl. the trim (heli_models.py:
 * 285-328), ref = env.get_ref() (main_3dof.py:62), RecursiveLeastSquares.reset (model.py:79-89), PID integrator 0
 * (PID.py:8), DHPAgent weights ~ N(0, nn_stdev) with the target critic a copy of the critic (agents.py:19-22,
 * :42-45, :107), counters and accumulators.  The reference draws its weights from torch's global generator; here
 * draw (agent a, matrix m, element i) of episode g is Philox4x32-10 with counter (g_lo, g_hi, 4 a + m, i / 2), key =
 * seed, followed by Box-Muller (exact recipe: oracle/philox_np.py), so a batch is the same set of episodes however it
 * is cut into shards.  Writes EVERY per-episode row of the batch except the read-only inputs ref_par / *_hyper (the
 * `const` rows trim and theta_trim ARE written). */
typedef struct H3Init {
    uint64_t seed;
    int64_t episode0;            /* global index of the batch's episode 0 (rank offset of a sharded job) */
    double v_trim;               /* trim speed, m/s (main_3dof.py:54), unless v_trim_ep */
    double sigma[4];             /* nn_stdev: col actor, col critic, lon actor, lon critic */
    double rls_cov0;             /* model.py:85 */
    const double* v_trim_ep;     /* DEVICE [ld] or NULL */
    const double* sigma_ep;      /* DEVICE [4][ld] or NULL: per-episode standard deviations (sweeps) */
    const int64_t* episode_id;   /* DEVICE [ld] or NULL: explicit global indices (overrides episode0 + e) */
} H3Init;
int h3_init_episodes(const H3Config* cfg, const H3Batch* b, const H3Init* init, void* stream);

/* Tensor memory (TMEM, 256 KB per SM on sm_100a) used as thread-private storage by the fused kernel: functional
 * probe.  Launches `blocks` CTAs of `warps` (4 or 8) warps; every thread fills its 128 private doubles with
 * (global thread index * 1024 + word), reads them back in the piece sizes the kernel uses (out[0..127][nthreads]),
 * runs `reps` read-modify-write (+1) passes (out[128..135] = words 0..7 afterwards, out[136] = checksum of a
 * dependent-load chain), and reports per-CTA cycle counts {rmw passes, dependent loads} in cycles[2][blocks].
 * out: DEVICE [137][blocks*warps*32] doubles. */
int h3_tmem_probe(int32_t warps, int32_t blocks, int32_t reps, double* out, int64_t* cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HELI3DOF_H */
