/*
 * heli6dof.h -- C ABI of the 6-DOF part of libheli3dof.so: batched, B200-native (sm_100a) implementation of
 * the 6-DOF closed-loop control tick of BartHelder/msc-thesis: Helicopter6DOF
 * (15 states, 4 controls, RK4, engine lag), LatPedPID / CollectivePID6DOF, RefGenerator, the 17 x 15
 * incremental RLS model and the two sigmoid-output IDHP agents, in the tick order of train.py:109-224.
 *
 * As for the 3-DOF path (heli3dof.h) the reference has no FFI layer; each entry point cites the reference
 * METHOD it replaces.  Conventions are those of heli3dof.h: caller-owned DEVICE buffers in struct-of-arrays
 * layout `[field][ld]`, asynchronous on `stream`, no allocation, no global state, 0 = ok / negative = H3_E_*
 * (h3_last_error() gives the text).  All real-valued buffers are double (fp64 only).
 *
 * Numerical failure is DATA: per-episode `flags` bit0 = env.t > t_max (heli_models.py:470-472), bit1 = NaN in
 * the actions (train.py:213-216), bit2 = pitch / roll envelope left (util.py:184-202, train.py:207-211),
 * bit3 = touched down in test_2 (train.py:218); `term_step` = tick index at which the episode stopped.
 */
#ifndef HELI6DOF_H
#define HELI6DOF_H

#include <stdint.h>

#include "heli3dof.h"

#ifdef __cplusplus
extern "C" {
#endif

#define H6_N_STATE 15   /* u v w p q r phi theta psi x y z lambda0_mr lambda0_tr omega  (heli_models.py:494) */
#define H6_N_ACTION 4   /* collective, longitudinal, lateral, pedal in [0, 1]            (heli_models.py:491) */
#define H6_N_REG 17     /* RLS regressors: 15 state + 2 action increments (train.py:189: actions[:2]) */
#define H6_N_HIDDEN 10  /* params.py: n_hidden_actor = n_hidden_critic = 10 for both agents */
#define H6_N_STATS 10
#define H6_WORK_ROWS 44
#define H6_HIST_ROWS 50   /* per tick: X (17), cov.X (17), 1/d (1; 0 = no RLS update that tick), Y (15) */

#define H6_FLAG_TMAX 1
#define H6_FLAG_NAN 2
#define H6_FLAG_ENVELOPE 4
#define H6_FLAG_LANDED 8

/* RefGenerator.task (util.py:119, :168-181) */
#define H6_TASK_NONE 0
#define H6_TASK_HOVER 1
#define H6_TASK_TRAIN_LON 2
#define H6_TASK_TRAIN_COL 3
#define H6_TASK_VELOCITY 4
#define H6_TASK_DESCENT 5

/* train.train(mode=...) (train.py:113-170): which schedule of events the fused loop applies */
#define H6_MODE_TRAIN 0
#define H6_MODE_TEST1 1
#define H6_MODE_TEST2 2
#define H6_MODE_FREE 3   /* no scheduled events: the host drives set_task / engine status between launches */

/* rows of H6Batch.gen (RefGenerator + FirstOrderLag state, util.py:110-120, :205-222) */
#define H6_GEN_ZREF 0
#define H6_GEN_TSWITCH 1
#define H6_GEN_A 2
#define H6_GEN_INT_ERROR_U 3
#define H6_GEN_FILTER_T0 4
#define H6_GEN_FILTER_X0 5
#define H6_GEN_FILTER_SETPOINT 6
#define H6_GEN_FILTER_TAU 7
#define H6_GEN_ROWS 8

/* rows of H6Batch.pid */
#define H6_PID_HDOT_CORR 0  /* CollectivePID6DOF.hdot_corr   PID.py:44 */
#define H6_PID_Y_REQ_INT 1  /* LatPedPID.y_req_int           PID.py:86 */
#define H6_PID_PHI_INT 2    /* LatPedPID.phi_int             PID.py:87 */
#define H6_PID_PSI_INT 3    /* LatPedPID.psi_int             PID.py:89 */

typedef struct H6Config {
    double dt;                 /* env_params['dt'] */
    double t_max;              /* env_params['t_max']; done when env.t > t_max before the step (heli_models.py:470) */
    double lat_gains[9];       /* LatPedPID.gains as stored (PID.py:50-52) */
    double col_kp, col_ki, col_kd, col_href;  /* CollectivePID6DOF (PID.py:27-30; train.py:84-87) */
    double gen_T, gen_dt;      /* RefGenerator(T=, dt=) (train.py:61) */
    double rls_gamma;          /* model.py:14 */
    double limit_pitch, limit_roll;           /* envelope_limits_reached (util.py:184), degrees */
    double P_engine_failure;   /* P_available set at tick 1000 of test_2 (train.py:164-165: Poeitr) */
    double amp[3], amp_t[2];   /* RefGenerator.A schedule: amp[0] while env.t < amp_t[0], amp[1] while < amp_t[1],
                                  else amp[2] (train.py:180-185: 10 / 15 / 20 at 20 s / 40 s) */
    double lr_decay;           /* factor on the lon learning rates at step_switch (train.py:120-121: 0.1) */
    int32_t run_mode;          /* H6_MODE_* */
    int32_t step_switch;       /* env_params['step_switch'] */
    int32_t exc_off_step;      /* excitation_phase ends at this tick (train.py:117-118: 1000) */
    int32_t rls_symmetric;     /* 1: covariance known symmetric; only its upper triangle is read and updated
                                  during a launch, the lower one is mirrored at launch exit */
    int32_t win_lo, win_hi;    /* score window on the lon reward (train.py:256: [5000, 6000)) */
    int32_t target_synced;     /* bit 0 (col) / bit 1 (lon): the caller asserts target critic == critic at entry; with
                                  tau == 1 the split path then never reads the target rows (agents.py:196) */
    int32_t pad_;
    H3AgentHyper col, lon;     /* gamma, tau, scale, reward_weight, tracked_state, ac_state, channel,
                                  final_activation; the learning rates live per episode in H6Batch.lr */
} H6Config;

typedef struct H6Batch {
    int64_t n, ld;
    double* state;       /* [15][ld]                                     heli_models.py:461 */
    double* t;           /* [ld]   env.t                                  heli_models.py:473 */
    double* P_out;       /* [ld]   engine lag state                       heli_models.py:460, :631 */
    double* P_avail;     /* [ld]   P_available                            heli_models.py:459, :477-487 */
    double* ref;         /* [3][ld] current reference, slots 0 (u, NaN if unset), 7 (theta), 11 (z) */
    double* pid;         /* [4][ld] H6_PID_* */
    double* gen;         /* [H6_GEN_ROWS][ld] */
    int32_t* task;       /* [ld]   H6_TASK_* */
    const double* trim_act;  /* [4][ld] trim controls                     heli_models.py:814 */
    const double* phi_trim;  /* [ld]   trim_state[6]                      train.py:79 */
    double* lr;          /* [4][ld] col actor, col critic, lon actor, lon critic (train.py:120-121 mutates) */
    double* rls_W;       /* [17*15][ld]  W[i][j] at row i*15+j            model.py:84 */
    double* rls_cov;     /* [17*17][ld]                                   model.py:85 */
    double* rls_x;       /* [15][ld] */
    double* rls_u;       /* [2][ld] */
    int32_t* rls_skip;   /* [ld] */
    double* col_net;     /* [11*10][ld] network block layout of heli3dof.h */
    double* lon_net;     /* [11*10][ld] */
    int32_t* step;       /* [ld] */
    int32_t* flags;      /* [ld] H6_FLAG_* */
    int32_t* term_step;  /* [ld] */
    int32_t* update_col; /* [ld] train.py:94, :122 */
    int32_t* excitation_phase; /* [ld] train.py:93, :118 */
    double* acc;         /* [3][ld] sum r_col, sum r_lon, sum r_lon inside the score window */
    double* work;        /* [H6_WORK_ROWS][ld] scratch of the split path (simulate / learn kernels), or NULL:
                            h6_closed_loop then runs the fused one-thread-per-episode kernel */
    double* hist;        /* [hist_len][H6_HIST_ROWS][ld] or NULL.  With a history buffer the split path updates only
                            the columns of W the agents read (their ac_state / tracked_state columns, agents.py:130-140)
                            every tick; the other columns -- independent recursions that nothing reads inside the
                            loop -- are brought up to date from the recorded (X, cov.X, 1/d, Y) every hist_len ticks
                            and at the end of the call, with W held in registers: same arithmetic, a third less
                            HBM traffic per tick */
    int64_t hist_len;
    int64_t tile;        /* 0: rls_W, rls_cov, col_net, lon_net are plain `[row][ld]`; 32: they are tiled
                            `[ceil(n/32)][rows][32]` (32 episodes' rows contiguous: sequential HBM pages for the
                            streaming kernels), see h6_tile_pack.  Only h6_closed_loop accepts tiled buffers. */
    int64_t rls_f32;     /* 0: rls_W / rls_cov hold doubles.  1 (needs tile = 32, work != NULL, hist == NULL): the two
                            TILED buffers hold IEEE binary32 elements (h6_tile_pack_f32; the pointers address float
                            arrays of ceil(n/32) * rows * 32 elements): the RLS estimator is STORED in single precision
                            between ticks -- half the bytes of the HBM-bound update -- while every product and sum of
                            model.py:41-59 and agents.py:130-140 is still formed in fp64.  Not the reference's
                            precision (numpy float64, model.py:84-85): a throughput option with a stated tolerance
                            (DESIGN.md section 9). */
} H6Batch;

/* Optional per-tick log of h6_closed_loop (what Logger.log_states records, util.py:30-66), for the first
 * n_traj episodes: [tick][H6_TRAJ_ROWS][ld_traj] = state(15, pre-step), ref(3), actions(4), r_col, r_lon,
 * P_out, alive. */
#define H6_TRAJ_ROWS 26
typedef struct H6Traj {
    double* out;
    int64_t n_traj, ld_traj;
} H6Traj;

/* Engine power levels of Helicopter6DOF (heli_models.py:397-401): out[0] = Paeo, out[1] = Poei, out[2] = Poeitr */
void h6_engine_power(double* out3);

/* Helicopter6DOF.trim(trim_speed, flight_path_angle, altitude) (heli_models.py:723-816), batched: per episode
 * v[n], fpa_deg[n], altitude[n] -> state[15][ld], controls[4][ld], P_out[ld]; iters[ld] (may be NULL) receives
 * the Newton iteration count, negative if not converged within 100. */
int h6_trim(const double* v, const double* fpa_deg, const double* altitude, double P_avail, double* state,
            double* controls, double* P_out, int32_t* iters, int64_t n, int64_t ld, void* stream);

/* Helicopter6DOF.calculate_state_derivatives(actions, state, trimming) (heli_models.py:489-694):
 * actions[4][ld], state[15][ld] -> dots[15][ld]; P_out[ld] is the engine integrator (read and written). */
int h6_env_derivatives(const double* actions, const double* state, double* P_out, const double* P_avail,
                       int32_t trimming, double dt, double* dots, int64_t n, int64_t ld, void* stream);

/* Helicopter6DOF.step (heli_models.py:464-475) for n_ticks ticks, open loop: action(k) = actions[4][ld] (or
 * b->trim_act if NULL) + excitation[k][4] for step k < exc_len, clipped to [0, 1].  Updates state, t, P_out,
 * step, flags, term_step.  states_out (optional, [n_ticks][15][ld_out]): post-step state of the first n_out. */
int h6_env_step(const H6Config* cfg, const H6Batch* b, const double* actions, const double* excitation,
                int32_t exc_len, int32_t n_ticks, double* states_out, int64_t n_out, int64_t ld_out, void* stream);

/* CollectivePID6DOF.__call__ (PID.py:31-45): collective[ld] from obs[15][ld]; updates b->pid row 0. */
int h6_pid_collective(const H6Config* cfg, const H6Batch* b, const double* obs, double* collective_out, void* stream);

/* LatPedPID.__call__ (PID.py:62-91): lat[ld], ped[ld] from obs[15][ld]; updates b->pid rows 1..3. */
int h6_pid_latped(const H6Config* cfg, const H6Batch* b, const double* obs, double* lat_out, double* ped_out,
                  void* stream);

/* RefGenerator.get_ref(obs, t) (util.py:122-166): ref_out[3][ld] (slots 0, 7, 11) from obs[15][ld] and t[ld];
 * amplitude[ld] (may be NULL) overwrites RefGenerator.A first (train.py:180-185).  Updates b->gen. */
int h6_refgen_get_ref(const H6Config* cfg, const H6Batch* b, const double* obs, const double* t,
                      const double* amplitude, double* ref_out, void* stream);

/* RefGenerator.set_task(task, t, obs, **kwargs) (util.py:168-181): t[ld], obs[15][ld]; z_start,
 * filter_target (NULL: not given) are [ld]; t_switch is the scalar keyword of task 'descent'. */
int h6_refgen_set_task(const H6Batch* b, int32_t task, const double* t, const double* obs, const double* z_start,
                       const double* filter_target, double t_switch, void* stream);

/* RecursiveLeastSquares.update / predict (model.py:27-69) for state_size 15, action_size 2:
 * state / next_state [15][ld], action [2][ld], out [15][ld]. */
int h6_rls_update(const H6Config* cfg, const H6Batch* b, const double* state, const double* action,
                  const double* next_state, void* stream);
int h6_rls_predict(const H6Batch* b, const double* state, const double* action, double* out, void* stream);

/* The fused closed-loop tick of train.py:109-224 (mode cfg->run_mode), n_ticks ticks per launch, every live
 * episode advancing from its own b->step.  excitation: DEVICE [exc_len][4] doubles or NULL (added while the
 * episode's excitation_phase flag is set, train.py:173-174).  stats: DEVICE [H6_N_STATS] doubles (plus
 * scratch, h6_stats_scratch_bytes(n)) receiving {n, n_alive, n_tmax, n_nan, n_envelope, n_landed,
 * sum r_col, sum r_lon, sum window r_lon, sum ticks}; NULL to skip. */
int h6_closed_loop(const H6Config* cfg, const H6Batch* b, const double* excitation, int32_t exc_len,
                   int32_t n_ticks, const H6Traj* traj, double* stats, void* stats_scratch, void* stream);

/* Layout conversion of one `[rows][ld]` array of n episodes to / from the tiled form `[ceil(n/32)][rows][32]`
 * (H6Batch.tile = 32); to_tiled != 0: plain -> tiled, else tiled -> plain.  The tiled buffer holds
 * ceil(n/32) * rows * 32 doubles. */
int h6_tile_pack(const double* src, double* dst, int32_t rows, int64_t n, int64_t ld, int32_t to_tiled, void* stream);

/* The same conversion between a plain `[rows][ld]` array of doubles and a tiled array of floats (H6Batch.rls_f32):
 * to_tiled != 0 rounds plain -> tiled (round to nearest even), else widens tiled -> plain. */
int h6_tile_pack_f32(double* plain, float* tiled, int32_t rows, int64_t n, int64_t ld, int32_t to_tiled, void* stream);

/* Bytes of device scratch h6_closed_loop needs next to `stats`. */
int64_t h6_stats_scratch_bytes(int64_t n);

/* Occupancy facts of the fused 6-DOF kernel: threads per block, resident blocks per SM, registers per thread,
 * local-memory (spill) bytes per thread. */
int h6_closed_loop_info(const H6Config* cfg, int32_t* block, int32_t* blocks_per_sm, int32_t* regs,
                        int32_t* local_bytes);

#ifdef __cplusplus
}
#endif
#endif /* HELI6DOF_H */
