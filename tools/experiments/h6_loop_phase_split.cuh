// The fused 6-DOF closed-loop kernel (train.py:109-224), one thread per episode, n_ticks ticks per launch.
//
// A tick has two phases with opposite needs, so they are separated by an explicit hand-over instead of one
// register allocation for both:
//   S ("simulate"): PIDs, reference generator, actor forward passes, RK4 of the 15-state model -- FP64 work on
//      ~60 words of per-episode state that are loaded from / stored to the H6Batch rows every tick (5 % of the
//      tick's traffic);
//   L ("learn"): the 17x15 RLS update and the two IDHP updates -- 1.2 K words per episode streamed from HBM,
//      coalesced `[row][episode]`, in software-pipelined groups (stream_groups: the loads of group g + DEPTH are
//      in flight while group g is consumed), which needs registers for load buffers, not for model state.
// What L needs from S (the regressor X, the target Y, the agents' augmented states and reward slopes: 44 words)
// is handed over in shared memory, `[row][thread]`.
#pragma once
#include "h3_tick.cuh"
#include "h6_model.cuh"

namespace h6 {

using h3::AgentH;
using h3::NetAcc;

constexpr int NS = H6_N_STATE, NR = H6_N_REG, NH = H6_N_HIDDEN;
constexpr int LOOP_BLOCK = 128;
#ifndef H6_MINB
#define H6_MINB 4
#endif
#ifndef H6_U
#define H6_U 2      // hidden units per chunk of the agent loops (NH: fully unrolled, no local arrays)
#endif
#ifndef H6_DEPTH
#define H6_DEPTH 1    // covariance groups (9 elements) in flight ahead of the one being consumed
#endif
#ifndef H6_WDEPTH
#define H6_WDEPTH 0   // W columns (17 elements) in flight ahead of the one being consumed
#endif
#ifndef H6_X_SMEM
#define H6_X_SMEM 1   // regressor read from the hand-over rows on demand instead of held in registers
#endif

// hand-over rows (shared memory, per thread)
enum { SR_X = 0, SR_Y = NR, SR_LON = NR + NS, SR_COL = NR + NS + 6, SR_ROWS = NR + NS + 12 };
enum { AG_S0 = 0, AG_S1 = 1, AG_N0 = 2, AG_N1 = 3, AG_DR = 4, AG_ON = 5 };

struct GVec {               // one episode's column of a `[rows][ld]` global array
    static constexpr bool kFence = false;
    double* p;
    int64_t ld;
    H3_DEV double& operator[](int i) { return p[(int64_t)i * ld]; }
};
typedef NetAcc<double, NH, GVec> Net;

struct LoopParams {
    H6Config C;
    ModelK K;
    H6Batch b;
    const double* excitation;
    int exc_len, n_ticks;
    H6Traj traj;
};

// Streaming order of the covariance: groups of 9 elements.  SYM: rows p and 16 - p of the upper triangle hold
// 17 - p and p + 1 elements, 18 together = two groups (rows 0..7 paired with 16..9; row 8 is the 17th group).
// Otherwise: every full row is two groups (9 + 8 elements).  Plain index arithmetic, so that everything folds to
// constants once the group loops are unrolled.
constexpr int COV_GS = 9;
template <bool SYM>
struct CovOrder {
    static constexpr int NG = SYM ? 17 : 2 * NR;
    H3_DEV static bool valid(int g, int e) { return SYM ? true : ((g & 1) * COV_GS + e) < NR; }
    H3_DEV static int row(int g, int e) {
        if (!SYM) return g >> 1;
        if (g == 16) return 8;
        const int p = g >> 1, q = (g & 1) * COV_GS + e;
        return q < NR - p ? p : NR - 1 - p;
    }
    H3_DEV static int col(int g, int e) {
        if (!SYM) return (g & 1) * COV_GS + e;
        if (g == 16) return 8 + e;
        const int p = g >> 1, q = (g & 1) * COV_GS + e;
        return q < NR - p ? p + q : (NR - 1 - p) + (q - (NR - p));
    }
};

// RecursiveLeastSquares.update (model.py:41-59), n = 17, m = 15, W / cov in global memory, X / Y in shared.
//   SYM: cov == cov^T is an invariant of the update (cov0 = c I): only the upper triangle is read and written and
//   cov.X == (X^T.cov)^T; otherwise both products are formed as the reference does (model.py:53-54).
template <bool SYM, int DEPTH>
H3_DEV void rls_update_stream(double* __restrict__ W, double* __restrict__ cov, int64_t ld, const double* sm, double gamma,
                              double inv_gamma) {
#if H6_X_SMEM
    auto X = [&](int i) { return sm[(SR_X + i) * LOOP_BLOCK]; };     // read on demand: no registers held
#else
    double Xr[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) Xr[i] = sm[(SR_X + i) * LOOP_BLOCK];
    auto X = [&](int i) { return Xr[i]; };
#endif
    double covX[NR], Xcov[SYM ? 1 : NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) covX[i] = 0;
    if (!SYM) {
#pragma unroll
        for (int i = 0; i < NR; ++i) Xcov[i] = 0;
    }
    typedef CovOrder<SYM> O;
    constexpr int GS = COV_GS, NG = O::NG;
    auto load_group = [&](int g, double (&c)[GS]) {
#pragma unroll
        for (int e = 0; e < GS; ++e)
            if (O::valid(g, e)) c[e] = cov[(int64_t)(O::row(g, e) * NR + O::col(g, e)) * ld];
    };
    h3::stream_groups<double, NG, GS, DEPTH, true>(load_group, [&](int g, double (&c)[GS]) {
#pragma unroll
        for (int e = 0; e < GS; ++e) {
            if (O::valid(g, e)) {
                const int i = O::row(g, e), j = O::col(g, e);
                covX[i] += c[e] * X(j);                         // model.py:53
                if (SYM) { if (i != j) covX[j] += c[e] * X(i); }
                else Xcov[j] += X(i) * c[e];                     // model.py:54
            }
        }
    });
    double d = gamma;                                           // model.py:55
#pragma unroll
    for (int i = 0; i < NR; ++i) d += (SYM ? covX[i] : Xcov[i]) * X(i);
    const double inv_d = 1.0 / d;
    double gain[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) gain[i] = covX[i] * inv_d;
    // model.py:49-50, :58 column by column: each weight is loaded and stored once
    h3::stream_groups<double, NS, NR, H6_WDEPTH, true>(
        [&](int j, double (&w)[NR]) {
#pragma unroll
            for (int i = 0; i < NR; ++i) w[i] = W[(int64_t)(i * NS + j) * ld];
        },
        [&](int j, double (&w)[NR]) {
            double ya = sm[(SR_Y + j) * LOOP_BLOCK], yb = 0;
#pragma unroll
            for (int i = 0; i < NR; i += 2) ya -= X(i) * w[i];
#pragma unroll
            for (int i = 1; i < NR; i += 2) yb += X(i) * w[i];
            const double err = ya - yb;
#pragma unroll
            for (int i = 0; i < NR; ++i) W[(int64_t)(i * NS + j) * ld] = w[i] + gain[i] * err;
        });
    // model.py:59
    h3::stream_groups<double, NG, GS, DEPTH, true>(load_group, [&](int g, double (&c)[GS]) {
#pragma unroll
        for (int e = 0; e < GS; ++e) {
            if (O::valid(g, e)) {
                const int i = O::row(g, e), j = O::col(g, e);
                cov[(int64_t)(i * NR + j) * ld] = (c[e] - gain[i] * (SYM ? covX[j] : Xcov[j])) * inv_gamma;
            }
        }
    });
}

// get_transition_matrices + update_networks (agents.py:130-196) of one agent from the hand-over rows
H3_DEV void agent_learn(const H3AgentHyper& h, Net& net, double lr_a, double lr_c, const double* __restrict__ W, int64_t ld,
                        const double* sm, bool tau1_fast) {
    const double s0 = sm[AG_S0 * LOOP_BLOCK], s1 = sm[AG_S1 * LOOP_BLOCK], n0 = sm[AG_N0 * LOOP_BLOCK],
                 n1 = sm[AG_N1 * LOOP_BLOCK], dr = sm[AG_DR * LOOP_BLOCK];
    double hh[NH], scr[NH], f;
    // the actor forward pass of (obs, ref) with the pre-update weights (== get_action's pass, agents.py:169)
    (void)h3::actor_forward<double, NH, H6_U>(net, s0, s1, h.scale, h.final_activation, hh, f);
    double F[2][2], G[2];                                               // agents.py:130-140, post-update W
    const int idx[2] = {h.ac_state, h.tracked_state};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = W[(int64_t)(idx[j] * NS + idx[i]) * ld];
        G[i] = W[(int64_t)((NS + h.channel) * NS + idx[i]) * ld];
    }
    AgentH<double> H;
    H.gamma = h.gamma; H.lr_a = lr_a; H.lr_c = lr_c; H.tau = h.tau; H.scale = h.scale;
    // tau == 1: after one update the target critic IS the critic (agents.py:196), so later ticks of the launch
    // neither read nor write its rows; they are refreshed from the critic at launch exit
    if (tau1_fast) h3::agent_update<double, NH, H6_U, true>(net, H, h.final_activation, s0, s1, n0, n1, 0.0, dr, F, G, hh, f, scr);
    else h3::agent_update<double, NH, H6_U, false>(net, H, h.final_activation, s0, s1, n0, n1, 0.0, dr, F, G, hh, f, scr);
}

H3_DEV void phase_fence() { asm volatile("" ::: "memory"); }

// Phase S of one tick.  Returns the termination flags of the tick (0: the episode goes on).
H3_DEV int tick_simulate(const LoopParams& P, int64_t e, int it, double* sm, int& update_col_out, bool& rls_on) {
    const H6Config& C = P.C;
    const H6Batch& b = P.b;
    const int64_t ld = b.ld;
    double obs[NS], ref[3], pid[4], gen[H6_GEN_ROWS], trim[4];
#pragma unroll
    for (int i = 0; i < NS; ++i) obs[i] = b.state[i * ld + e];
#pragma unroll
    for (int i = 0; i < 3; ++i) ref[i] = b.ref[i * ld + e];
#pragma unroll
    for (int i = 0; i < 4; ++i) { pid[i] = b.pid[i * ld + e]; trim[i] = b.trim_act[i * ld + e]; }
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) gen[i] = b.gen[i * ld + e];
    double t = b.t[e], P_out = b.P_out[e], P_avail = b.P_avail[e];
    int task = b.task[e], update_col = b.update_col[e], exc_phase = b.excitation_phase[e];
    const int k = b.step[e];

    const ObsTrig g = obs_trig(obs);
    double act[4];
    pid_latped(C, obs, g, b.phi_trim[e], trim[2], trim[3], pid[1], pid[2], pid[3], act[2], act[3]);     // train.py:110
    bool both = true;
    if (C.run_mode == H6_MODE_TRAIN) {                                                                    // :113-137
        if (k == 0) {
            gen_set_task(task, gen, H6_TASK_TRAIN_LON, 0.0, obs, 0.0, 0.0, 0.0);
            gen_get_ref(C, task, gen, obs, t, ref);
        } else if (k == C.exc_off_step) {
            exc_phase = 0;
            b.excitation_phase[e] = 0;
        } else if (k == C.step_switch) {
            b.lr[2 * ld + e] *= C.lr_decay;
            b.lr[3 * ld + e] *= C.lr_decay;
            update_col = 1;
            b.update_col[e] = 1;
            gen_set_task(task, gen, H6_TASK_TRAIN_COL, t, obs, obs[11], 0.0, 0.0);
        }
        both = !(k < C.step_switch);
    } else if (C.run_mode == H6_MODE_TEST1) {                                                             // :138-160
        if (k == 0) { gen_set_task(task, gen, H6_TASK_HOVER, 0.0, obs, 0.0, 0.0, 0.0); gen_get_ref(C, task, gen, obs, t, ref); }
        else if (k == 500) gen_set_task(task, gen, H6_TASK_VELOCITY, t, obs, 0.0, 25 - obs[0], 0.0);
        else if (k == 2000) gen_set_task(task, gen, H6_TASK_VELOCITY, t, obs, 0.0, 0 - obs[0], 0.0);
    } else if (C.run_mode == H6_MODE_TEST2) {                                                             // :161-170
        if (k == 0) { gen_set_task(task, gen, H6_TASK_DESCENT, 0.0, obs, 0.0, 0.0, 0.0); gen_get_ref(C, task, gen, obs, t, ref); }
        else if (k == 1000) { P_avail = C.P_engine_failure; b.P_avail[e] = P_avail; }
    } else {
        both = update_col != 0;
    }
    // actions (:126-137)
    Net ncol, nlon;
    ncol.v.p = b.col_net + e; ncol.v.ld = ld;
    nlon.v.p = b.lon_net + e; nlon.v.ld = ld;
    double hh[NH], f;
    const double lon_s0 = obs[C.lon.ac_state], lon_s1 = obs[C.lon.tracked_state] - ref_slot(ref, C.lon.tracked_state);
    const double col_s0 = obs[C.col.ac_state], col_s1 = obs[C.col.tracked_state] - ref_slot(ref, C.col.tracked_state);
    act[1] = trim[1] - 0.5 + h3::actor_forward<double, NH, H6_U>(nlon, lon_s0, lon_s1, C.lon.scale, C.lon.final_activation, hh, f);
    if (both) act[0] = trim[0] - 0.5 + h3::actor_forward<double, NH, H6_U>(ncol, col_s0, col_s1, C.col.scale, C.col.final_activation, hh, f);
    else act[0] = pid_collective(C, obs, g, trim[0], pid[0]);
    if (exc_phase && P.excitation != nullptr && k < P.exc_len) {                                          // :173-174
#pragma unroll
        for (int i = 0; i < 4; ++i) act[i] += P.excitation[4 * k + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) act[i] = clip01_nan(act[i]);                                              // :176
    // env.step (:179; heli_models.py:464-475)
    const Controls ctl = make_controls(P.K, act[0], act[1], act[2], act[3]);
    double nxt[NS];
    rk4_step(P.K, ctl, obs, P_out, P_avail, C.dt, nxt);
    const bool done = t > C.t_max;
    t += C.dt;
    gen[H6_GEN_A] = t < C.amp_t[0] ? C.amp[0] : (t < C.amp_t[1] ? C.amp[1] : C.amp[2]);                  // :180-185
    double nref[3];
    gen_get_ref(C, task, gen, obs, t, nref);                                                              // :186 (pre-step obs)
    // hand-over to phase L: regressor / target of the RLS update (:189; model.py:41-46) ...
    const int skip = b.rls_skip[e];
    rls_on = skip == 0;
    if (skip) b.rls_skip[e] = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (!skip) sm[(SR_X + i) * LOOP_BLOCK] = obs[i] - b.rls_x[i * ld + e];
        sm[(SR_Y + i) * LOOP_BLOCK] = nxt[i] - obs[i];
        b.rls_x[i * ld + e] = obs[i];
    }
    if (!skip) {
        sm[(SR_X + NS) * LOOP_BLOCK] = act[0] - b.rls_u[e];
        sm[(SR_X + NS + 1) * LOOP_BLOCK] = act[1] - b.rls_u[ld + e];
    }
    b.rls_u[e] = act[0];
    b.rls_u[ld + e] = act[1];
    // ... and the agents' inputs: augmented states, reward and its slope (:192-201; agents.py:123-128, :142-150)
    const double et_lon = nxt[C.lon.tracked_state] - ref_slot(ref, C.lon.tracked_state);
    const double et_col = nxt[C.col.tracked_state] - ref_slot(ref, C.col.tracked_state);
    const double r_lon = -(et_lon * et_lon) * C.lon.reward_weight;
    const double r_col = update_col ? -(et_col * et_col) * C.col.reward_weight : 0.0;
    double* sl = sm + SR_LON * LOOP_BLOCK;
    sl[AG_S0 * LOOP_BLOCK] = lon_s0; sl[AG_S1 * LOOP_BLOCK] = lon_s1;
    sl[AG_N0 * LOOP_BLOCK] = nxt[C.lon.ac_state];
    sl[AG_N1 * LOOP_BLOCK] = nxt[C.lon.tracked_state] - ref_slot(nref, C.lon.tracked_state);
    sl[AG_DR * LOOP_BLOCK] = -2 * et_lon * C.lon.reward_weight;
    double* sc = sm + SR_COL * LOOP_BLOCK;
    sc[AG_S0 * LOOP_BLOCK] = col_s0; sc[AG_S1 * LOOP_BLOCK] = col_s1;
    sc[AG_N0 * LOOP_BLOCK] = nxt[C.col.ac_state];
    sc[AG_N1 * LOOP_BLOCK] = nxt[C.col.tracked_state] - ref_slot(nref, C.col.tracked_state);
    sc[AG_DR * LOOP_BLOCK] = -2 * et_col * C.col.reward_weight;
    update_col_out = update_col;
    b.acc[e] += r_col;
    b.acc[ld + e] += r_lon;
    if (k >= C.win_lo && k < C.win_hi) b.acc[2 * ld + e] += r_lon;                                        // :256
    // termination (:207-218), evaluated on the pre-step observation
    const double R2D = 180.0 / 3.14159265358979323846;
    int fl = done ? H6_FLAG_TMAX : 0;
    if (fabs(obs[7] * R2D) > C.limit_pitch || fabs(obs[6] * R2D) > C.limit_roll) fl |= H6_FLAG_ENVELOPE;
    if (act[0] != act[0] || act[1] != act[1] || act[2] != act[2] || act[3] != act[3]) fl |= H6_FLAG_NAN;
    if (C.run_mode == H6_MODE_TEST2 && obs[11] > 0) fl |= H6_FLAG_LANDED;
    if (P.traj.out != nullptr && e < P.traj.n_traj) {                                                      // util.py:30-66
        double* o = P.traj.out + (int64_t)it * H6_TRAJ_ROWS * P.traj.ld_traj + e;
        const int64_t lt = P.traj.ld_traj;
#pragma unroll
        for (int i = 0; i < NS; ++i) o[i * lt] = obs[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) o[(NS + i) * lt] = ref[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[(NS + 3 + i) * lt] = act[i];
        o[22 * lt] = r_col; o[23 * lt] = r_lon; o[24 * lt] = P_out; o[25 * lt] = 1.0;
    }
    // the episode's state after the tick (:222-223)
#pragma unroll
    for (int i = 0; i < NS; ++i) b.state[i * ld + e] = nxt[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) b.ref[i * ld + e] = nref[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) b.pid[i * ld + e] = pid[i];
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) b.gen[i * ld + e] = gen[i];
    b.t[e] = t;
    b.P_out[e] = P_out;
    b.task[e] = task;
    b.step[e] = k + 1;
    if (fl) { b.flags[e] = fl; b.term_step[e] = k; }
    return fl;
}

template <bool SYM>
__global__ void __launch_bounds__(LOOP_BLOCK, H6_MINB) closed_loop_kernel(const __grid_constant__ LoopParams P) {
    extern __shared__ double tile[];                        // [SR_ROWS][LOOP_BLOCK]
    const H6Config& C = P.C;
    const H6Batch& b = P.b;
    const int64_t e = (int64_t)blockIdx.x * LOOP_BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    if (b.flags[e] != 0) return;                            // frozen: the episode has ended
    double* sm = tile + threadIdx.x;
    double* __restrict__ W = b.rls_W + e;
    double* __restrict__ cov = b.rls_cov + e;
    Net ncol, nlon;
    ncol.v.p = b.col_net + e; ncol.v.ld = ld;
    nlon.v.p = b.lon_net + e; nlon.v.ld = ld;
    const double inv_gamma = 1.0 / C.rls_gamma;
    int ticks_done = 0, col_updates = 0;
#pragma unroll 1
    for (int it = 0; it < P.n_ticks; ++it) {
        int update_col;
        bool rls_on;
        const int fl = tick_simulate(P, e, it, sm, update_col, rls_on);
        phase_fence();
        if (rls_on) rls_update_stream<SYM, H6_DEPTH>(W, cov, ld, sm, C.rls_gamma, inv_gamma);
        phase_fence();
        if (update_col) {
            agent_learn(C.col, ncol, b.lr[e], b.lr[ld + e], W, ld, sm + SR_COL * LOOP_BLOCK, C.col.tau == 1.0 && col_updates > 0);
            col_updates += 1;
        }
        agent_learn(C.lon, nlon, b.lr[2 * ld + e], b.lr[3 * ld + e], W, ld, sm + SR_LON * LOOP_BLOCK,
                    C.lon.tau == 1.0 && ticks_done > 0);
        phase_fence();
        ticks_done += 1;
        if (fl) break;
    }
    if (P.traj.out != nullptr && e < P.traj.n_traj) {       // rows of ticks that were not executed: alive = 0
        for (int it = ticks_done; it < P.n_ticks; ++it)
            P.traj.out[((int64_t)it * H6_TRAJ_ROWS + 25) * P.traj.ld_traj + e] = 0.0;
    }
    if (C.lon.tau == 1.0 && ticks_done > 1) {               // target critic := critic (rows 7nh.. := rows 3nh..)
#pragma unroll 1
        for (int r = 0; r < 4 * NH; ++r) nlon.v[7 * NH + r] = nlon.v[3 * NH + r];
    }
    if (C.col.tau == 1.0 && col_updates > 1) {
#pragma unroll 1
        for (int r = 0; r < 4 * NH; ++r) ncol.v[7 * NH + r] = ncol.v[3 * NH + r];
    }
    if (SYM && ticks_done > 0) {                            // restore the full matrix for the callers
#pragma unroll 1
        for (int i = 0; i < NR; ++i)
            for (int j = i + 1; j < NR; ++j) cov[(int64_t)(j * NR + i) * ld] = cov[(int64_t)(i * NR + j) * ld];
    }
}

constexpr size_t LOOP_SMEM = (size_t)SR_ROWS * LOOP_BLOCK * sizeof(double);

}  // namespace h6
