// Split form of the 6-DOF closed loop (train.py:109-224): per tick one launch of each of three kernels that are
// shaped for opposite bounds, instead of one register allocation for all of the tick.
//
//   agents_kernel (one thread per (episode, agent), 16 warps per SM): the IDHP updates of the previous tick.
//   sim_kernel   ("S", one thread per episode, 168 registers): PIDs, reference generator, actor forward passes and
//                the RK4 step of the 15-state model.  FP64 / latency work on ~60 words of per-episode state.
//   learn_kernel ("L", 17 warps per 32 episodes): the 17x15 RLS update.  Lanes are episodes (coalesced
//                `[row][episode]` accesses), warp r owns row r of the covariance (cov.X, its downdate) and column r
//                of W (prediction error + update): every thread has 17 independent loads in flight and there are 34
//                such warps per SM, so the kernel is bound by HBM bandwidth, not by exposed latency as the fused
//                one-thread-per-episode kernel is (profiles/r01_h6_summary.txt).
//
// What L needs from S (regressor X, target Y, the agents' augmented states and reward slopes, flags) travels
// through H6Batch.work, `[H6_WORK_ROWS][ld]` in global memory (44 of the tick's ~1100 words).  W, the covariance and
// the networks are normally tiled (H6Batch.tile = 32: compile-time row stride, accesses are base + immediate).
// lazy_w_kernel is the optional catch-up of the W columns nothing reads inside the loop (H6Batch.hist).
// Included by h6_api.cu after GVec / Net / LoopParams / NS / NR / NH.
#pragma once

namespace h6 {

enum { WR_X = 0, WR_Y = NR, WR_LON = NR + NS, WR_COL = NR + NS + 5, WR_FLAGS = NR + NS + 10 };
static_assert(WR_FLAGS + 1 <= H6_WORK_ROWS, "work rows");
enum { AG_S0 = 0, AG_S1 = 1, AG_N0 = 2, AG_N1 = 3, AG_DR = 4 };
enum { WF_ACTIVE = 1, WF_RLS = 2, WF_COL = 4 };

// get_transition_matrices + update_networks (agents.py:130-196) of one agent from its work rows
template <typename NetT, typename WT>
H3_DEV void agent_learn(const H3AgentHyper& h, NetT& net, double lr_a, double lr_c, const WT* __restrict__ W, int64_t lw,
                        int64_t ld, const double* wk, bool tau1_fast) {
    const double s0 = wk[AG_S0 * ld], s1 = wk[AG_S1 * ld], n0 = wk[AG_N0 * ld], n1 = wk[AG_N1 * ld], dr = wk[AG_DR * ld];
    double hh[NH], scr[NH], f;
    // the actor forward pass of (obs, ref) with the pre-update weights (== get_action's pass, agents.py:169)
    (void)h3::actor_forward<double, NH, H6_U, h3::MathExp>(net, s0, s1, h.scale, h.final_activation, hh, f);
    double F[2][2], G[2];                                               // agents.py:130-140, post-update W
    const int idx[2] = {h.ac_state, h.tracked_state};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = (double)W[(int64_t)(idx[j] * NS + idx[i]) * lw];
        G[i] = (double)W[(int64_t)((NS + h.channel) * NS + idx[i]) * lw];
    }
    AgentH<double> H;
    H.gamma = h.gamma; H.lr_a = lr_a; H.lr_c = lr_c; H.tau = h.tau; H.scale = h.scale;
    if (tau1_fast) h3::agent_update<double, NH, H6_U, true, h3::MathExp>(net, H, h.final_activation, s0, s1, n0, n1, 0.0, dr, F, G, hh, f, scr);
    else h3::agent_update<double, NH, H6_U, false, h3::MathExp>(net, H, h.final_activation, s0, s1, n0, n1, 0.0, dr, F, G, hh, f, scr);
}

// Phase S of one tick (PIDs, reference generator, actions, RK4, bookkeeping) for one episode; leaves what
// phase L needs in the work rows `wk` (global memory, [H6_WORK_ROWS][ld]).
#ifndef H6_SIM_SMEM
#define H6_SIM_SMEM 1
#endif
template <bool TILED>
H3_DEV void tick_simulate(const LoopParams& P, int64_t e, int it, double* wk, double* rk_sm) {
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
    // everything the end of the tick needs from memory is requested now, with the state, so that its latency is
    // covered by the tick's arithmetic instead of being exposed after the RK4 step
    const int skip = b.rls_skip[e];
    double xprev[NS], uprev[2], acc[3];
#pragma unroll
    for (int i = 0; i < NS; ++i) xprev[i] = b.rls_x[i * ld + e];
    uprev[0] = b.rls_u[e]; uprev[1] = b.rls_u[ld + e];
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] = b.acc[i * ld + e];

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
    typename NetOf<TILED>::type ncol, nlon;
    ncol.v.p = b.col_net + tile_off(b, e, 11 * NH); ncol.v.ld = b.ld;
    nlon.v.p = b.lon_net + tile_off(b, e, 11 * NH); nlon.v.ld = b.ld;
    double hh[NH], f;
    const double lon_s0 = obs[C.lon.ac_state], lon_s1 = obs[C.lon.tracked_state] - ref_slot(ref, C.lon.tracked_state);
    const double col_s0 = obs[C.col.ac_state], col_s1 = obs[C.col.tracked_state] - ref_slot(ref, C.col.tracked_state);
    act[1] = trim[1] - 0.5 + h3::actor_forward<double, NH, H6_U, h3::MathExp>(nlon, lon_s0, lon_s1, C.lon.scale, C.lon.final_activation, hh, f);
    if (both) act[0] = trim[0] - 0.5 + h3::actor_forward<double, NH, H6_U, h3::MathExp>(ncol, col_s0, col_s1, C.col.scale, C.col.final_activation, hh, f);
    else act[0] = pid_collective(C, obs, g, trim[0], pid[0]);
    if (exc_phase && P.excitation != nullptr && k < P.exc_len) {                                          // :173-174
#pragma unroll
        for (int i = 0; i < 4; ++i) act[i] += P.excitation[4 * k + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) act[i] = clip01_nan(act[i]);                                              // :176
    // regressor of the RLS update (:189; model.py:41-46): known before the model step
    if (skip) b.rls_skip[e] = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (!skip) wk[(WR_X + i) * ld] = obs[i] - xprev[i];
        b.rls_x[i * ld + e] = obs[i];
    }
    if (!skip) {
        wk[(WR_X + NS) * ld] = act[0] - uprev[0];
        wk[(WR_X + NS + 1) * ld] = act[1] - uprev[1];
    }
    b.rls_u[e] = act[0];
    b.rls_u[ld + e] = act[1];
    // env.step (:179; heli_models.py:464-475)
    const Controls ctl = make_controls(P.K, act[0], act[1], act[2], act[3]);
    double nxt[NS];
#if H6_SIM_SMEM
    rk4_step_smem(P.K, ctl, obs, P_out, P_avail, C.dt, nxt, rk_sm, 128);
#else
    rk4_step(P.K, ctl, obs, P_out, P_avail, C.dt, nxt);
#endif
    const bool done = t > C.t_max;
    t += C.dt;
    gen[H6_GEN_A] = t < C.amp_t[0] ? C.amp[0] : (t < C.amp_t[1] ? C.amp[1] : C.amp[2]);                  // :180-185
    double nref[3];
    gen_get_ref(C, task, gen, obs, t, nref);                                                              // :186 (pre-step obs)
    // hand-over to phase L: target of the RLS update ...
#pragma unroll
    for (int i = 0; i < NS; ++i) wk[(WR_Y + i) * ld] = nxt[i] - obs[i];
    // ... and the agents' inputs: augmented states, reward and its slope (:192-201; agents.py:123-128, :142-150)
    const double et_lon = nxt[C.lon.tracked_state] - ref_slot(ref, C.lon.tracked_state);
    const double et_col = nxt[C.col.tracked_state] - ref_slot(ref, C.col.tracked_state);
    const double r_lon = -(et_lon * et_lon) * C.lon.reward_weight;
    const double r_col = update_col ? -(et_col * et_col) * C.col.reward_weight : 0.0;
    double* sl = wk + WR_LON * ld;
    sl[AG_S0 * ld] = lon_s0; sl[AG_S1 * ld] = lon_s1;
    sl[AG_N0 * ld] = nxt[C.lon.ac_state];
    sl[AG_N1 * ld] = nxt[C.lon.tracked_state] - ref_slot(nref, C.lon.tracked_state);
    sl[AG_DR * ld] = -2 * et_lon * C.lon.reward_weight;
    double* sc = wk + WR_COL * ld;
    sc[AG_S0 * ld] = col_s0; sc[AG_S1 * ld] = col_s1;
    sc[AG_N0 * ld] = nxt[C.col.ac_state];
    sc[AG_N1 * ld] = nxt[C.col.tracked_state] - ref_slot(nref, C.col.tracked_state);
    sc[AG_DR * ld] = -2 * et_col * C.col.reward_weight;
    wk[WR_FLAGS * ld] = (double)(WF_ACTIVE | (skip ? 0 : WF_RLS) | (update_col ? WF_COL : 0));
    b.acc[e] = acc[0] + r_col;
    b.acc[ld + e] = acc[1] + r_lon;
    if (k >= C.win_lo && k < C.win_hi) b.acc[2 * ld + e] = acc[2] + r_lon;                                // :256
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
}


// The IDHP updates of the tick that learn_kernel just completed (they need its W): one thread per (episode,
// agent), blockIdx.y = 0 lon, 1 col, 128 registers -> 16 warps per SM to hide the latency of the weight loads.
//   fast: tau == 1 and the target critic equals the critic at call entry (H6Config.target_synced): the target rows
//   are then neither read nor written by the updates (agents.py:196 makes them equal again); they are refreshed from
//   the critic after an episode's last update of the call (`final`, or the episode has just ended).
//   ST: storage type of the tiled W / covariance (H6Batch.rls_f32: float; the arithmetic is fp64 either way).
template <bool TILED, typename ST = double>
__global__ void __launch_bounds__(128, 4) agents_kernel(const __grid_constant__ LoopParams P, int final) {
    const H6Config& C = P.C;
    const H6Batch& b = P.b;
    const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    const double* wk = b.work + e;
    const int wf = (int)wk[WR_FLAGS * ld];
    const bool col = blockIdx.y == 1;
    if (!(wf & WF_ACTIVE) || (col && !(wf & WF_COL))) return;
    const H3AgentHyper& h = col ? C.col : C.lon;
    const bool fast = h.tau == 1.0 && (C.target_synced & (col ? 1 : 2));
    const int64_t lw = TILED ? 32 : ld;
    const ST* __restrict__ W = reinterpret_cast<const ST*>(b.rls_W) + tile_off(b, e, NR * NS);
    typename NetOf<TILED>::type net;
    net.v.p = (col ? b.col_net : b.lon_net) + tile_off(b, e, 11 * NH); net.v.ld = lw;
    const int o = col ? 0 : 2;
    agent_learn(h, net, b.lr[o * ld + e], b.lr[(o + 1) * ld + e], W, lw, ld, wk + (col ? WR_COL : WR_LON) * ld, fast);
    if (fast && (final || b.flags[e] != 0)) {
#pragma unroll 4
        for (int r = 0; r < 4 * NH; ++r) net.v[7 * NH + r] = net.v[3 * NH + r];
    }
}

#ifndef H6_SIM_MINB
#define H6_SIM_MINB 3
#endif
// One tick of PIDs, reference generator, actions and the RK4 model step for every live episode.
template <bool TILED>
__global__ void __launch_bounds__(128, H6_SIM_MINB) sim_kernel(const __grid_constant__ LoopParams P, int it) {
    const H6Batch& b = P.b;
    const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    double* wk = b.work + e;
    if (b.flags[e] != 0) {                                  // frozen: the episode has ended
        wk[WR_FLAGS * ld] = 0.0;
        if (P.traj.out != nullptr && e < P.traj.n_traj) P.traj.out[((int64_t)it * H6_TRAJ_ROWS + 25) * P.traj.ld_traj + e] = 0.0;
        return;
    }
    __shared__ double rk_tile[H6_SIM_SMEM ? 30 * 128 : 1];
    tick_simulate<TILED>(P, e, it, wk, rk_tile + threadIdx.x);
}

// RecursiveLeastSquares.update (model.py:41-59), n = 17, m = 15: 32 episodes per block (lanes), 17 warps.
//   SYM: cov == cov^T is an invariant of the update (cov0 = c I): only the upper triangle is kept current, warp r
//   reads the r elements above the diagonal in column r and the 17 - r elements of row r from the diagonal on;
//   otherwise cov.X and X^T.cov are formed separately from the full matrix as the reference does (model.py:53-54).
//   With the tiled layout every access is base + immediate (row stride 32): no address registers.
constexpr int LEARN_WARPS = NR, LEARN_BLOCK = 32 * LEARN_WARPS;
#ifndef H6_LEARN_MINB
#define H6_LEARN_MINB 2
#endif
// loads are issued in chunks of LEARN_CH elements (a chunk in flight per thread, 34 warps per SM) so that the
// kernel fits 56 registers without spilling
#ifndef H6_LEARN_CH
#define H6_LEARN_CH 6
#endif
constexpr int LEARN_CH = H6_LEARN_CH;
#ifndef H6_LEARN_CH32
#define H6_LEARN_CH32 12    // float storage (H6Batch.rls_f32): the same registers hold twice the elements
#endif

template <bool SYM, bool TILED, typename ST = double>
__global__ void __launch_bounds__(LEARN_BLOCK, H6_LEARN_MINB) learn_kernel(H6Batch b, double gamma, unsigned live_mask,
                                                                           double* __restrict__ hist) {
    __shared__ double Xs[NR][32], cXs[NR][32], Xcs[SYM ? 1 : NR][32];
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + lane, ld = b.ld;
    constexpr int CH = sizeof(ST) == 4 ? H6_LEARN_CH32 : LEARN_CH;  // raw elements in flight per thread and chunk
    const bool in = e < b.n;
    const int wf = in ? (int)b.work[WR_FLAGS * ld + e] : 0;
    const bool on = (wf & (WF_ACTIVE | WF_RLS)) == (WF_ACTIVE | WF_RLS);
    const double* __restrict__ wk = b.work + e;
    const int64_t lw = TILED ? 32 : ld;
    ST* __restrict__ cov = reinterpret_cast<ST*>(b.rls_cov) + tile_off(b, e, NR * NR);
    ST* __restrict__ rowp = cov + (int64_t)(r * NR) * lw;           // element (r, j) at rowp[j * lw]
    const ST* __restrict__ colp = cov + (int64_t)r * lw;              // element (j, r) at colp[j * NR * lw]
    ST* __restrict__ wp = reinterpret_cast<ST*>(b.rls_W) + tile_off(b, e, NR * NS) + (int64_t)r * lw;     // W(i, r) at wp[i * NS * lw]
    // hist != nullptr: this tick's slot of the history, [H6_HIST_ROWS][ld]; columns of W outside live_mask are not
    // touched here but brought up to date later by lazy_w_kernel from the recorded X, cov.X, 1/d, Y
    double* __restrict__ hs = hist != nullptr ? hist + e : nullptr;
    const double xr = on ? wk[(WR_X + r) * ld] : 0.0;
    Xs[r][lane] = xr;
    if (hs != nullptr && in) {
        if (on) hs[(int64_t)r * ld] = xr;
        else if (r == 0) hs[(int64_t)(2 * NR) * ld] = 0.0;                     // 1/d = 0: no update this tick
    }
    __syncthreads();
    double cx = 0, xc = 0;
    if (on) {
        double a0 = 0, a1 = 0;
#pragma unroll
        for (int j0 = 0; j0 < NR; j0 += CH) {
            ST c[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int j = j0 + u;
                if (j < NR) c[u] = (SYM && j < r) ? colp[(int64_t)(j * NR) * lw] : rowp[(int64_t)j * lw];
            }
            h3::row_fence<true>();
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int j = j0 + u;
                if (j < NR) { if (u & 1) a1 += (double)c[u] * Xs[j][lane]; else a0 += (double)c[u] * Xs[j][lane]; }    // model.py:53
            }
        }
        cx = a0 + a1;
        if (!SYM) {
            a0 = 0; a1 = 0;
#pragma unroll
            for (int i0 = 0; i0 < NR; i0 += CH) {
                ST c[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (i0 + u < NR) c[u] = colp[(int64_t)((i0 + u) * NR) * lw];
                h3::row_fence<true>();
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (i0 + u < NR) { if (u & 1) a1 += Xs[i0 + u][lane] * (double)c[u]; else a0 += Xs[i0 + u][lane] * (double)c[u]; }   // model.py:54
            }
            xc = a0 + a1;
        }
    }
    cXs[r][lane] = cx;
    if (!SYM) Xcs[r][lane] = xc;
    if (hs != nullptr && on) hs[(int64_t)(NR + r) * ld] = cx;
    __syncthreads();
    if (!on) return;
    double d = gamma;                                                           // model.py:55
#pragma unroll
    for (int i = 0; i < NR; ++i) d += (SYM ? cXs[i][lane] : Xcs[i][lane]) * Xs[i][lane];
    const double inv_d = 1.0 / d, inv_gamma = 1.0 / gamma;
    if (hs != nullptr && r == NR - 1) hs[(int64_t)(2 * NR) * ld] = inv_d;
    // model.py:59, row r (re-read: the row was loaded a moment ago by this warp, it is in L1 / L2)
    {
        const double gr = cx * inv_d;
#pragma unroll
        for (int j0 = 0; j0 < NR; j0 += CH) {
            ST c[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u)
                if (j0 + u < NR && (!SYM || j0 + u >= r)) c[u] = rowp[(int64_t)(j0 + u) * lw];
            h3::row_fence<true>();
#pragma unroll
            for (int u = 0; u < CH; ++u)
                if (j0 + u < NR && (!SYM || j0 + u >= r))
                    rowp[(int64_t)(j0 + u) * lw] = (ST)(((double)c[u] - gr * (SYM ? cXs[j0 + u][lane] : Xcs[j0 + u][lane])) * inv_gamma);
        }
    }
    // model.py:49-50, :58, column r of W
    if (r < NS && hs != nullptr && !((live_mask >> r) & 1u)) {
        hs[(int64_t)(2 * NR + 1 + r) * ld] = wk[(WR_Y + r) * ld];
    } else if (r < NS) {
        ST w[NR];
        double ya = wk[(WR_Y + r) * ld], yb = 0;
#pragma unroll
        for (int i = 0; i < NR; ++i) w[i] = wp[(int64_t)(i * NS) * lw];
#pragma unroll
        for (int i = 0; i < NR; i += 2) ya -= Xs[i][lane] * (double)w[i];
#pragma unroll
        for (int i = 1; i < NR; i += 2) yb += Xs[i][lane] * (double)w[i];
        const double err = (ya - yb) * inv_d;
#pragma unroll
        for (int i = 0; i < NR; ++i) wp[(int64_t)(i * NS) * lw] = (ST)((double)w[i] + cXs[i][lane] * err);
    }
}

// The columns of W that nothing reads inside the loop (all but the agents' ac_state / tracked_state columns): column
// j of model.py:49-50, :58 is a recursion of its own (shared gain, own weights), so it can be advanced over n_slots
// recorded ticks in one go with its 17 weights in registers.  32 episodes per block, warp j = column j.
// The recorded X, cov.X and 1/d of the block's 32 episodes are first staged in shared memory by all 15 warps
// (independent coalesced loads: bandwidth-bound), then every lazy column runs its recursion out of shared memory.
constexpr int LAZY_ROWS = 2 * NR + 1;
constexpr int LAZY_MAX_SLOTS = 8;                               // 8 * 35 * 256 B = 72 KB of shared memory per block
template <bool TILED>
__global__ void __launch_bounds__(32 * NS, 2) lazy_w_kernel(H6Batch b, unsigned live_mask, int n_slots) {
    extern __shared__ double hsm[];                             // [n_slots][LAZY_ROWS][32]
    const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + lane, ld = b.ld;
    const bool in = e < b.n;
    const bool lazy_col = !((live_mask >> j) & 1u);
    const int64_t lw = TILED ? 32 : ld;
    double* __restrict__ wp = b.rls_W + tile_off(b, in ? e : 0, NR * NS) + (int64_t)j * lw;
    double w[NR], y[LAZY_MAX_SLOTS];
    if (in && lazy_col) {
#pragma unroll
        for (int i = 0; i < NR; ++i) w[i] = wp[(int64_t)(i * NS) * lw];
#pragma unroll
        for (int t = 0; t < LAZY_MAX_SLOTS; ++t)
            if (t < n_slots) y[t] = b.hist[((int64_t)t * H6_HIST_ROWS + 2 * NR + 1 + j) * ld + e];
    }
    for (int idx = j; idx < n_slots * LAZY_ROWS; idx += NS) {
        const int t = idx / LAZY_ROWS, row = idx - t * LAZY_ROWS;
        hsm[idx * 32 + lane] = in ? b.hist[((int64_t)t * H6_HIST_ROWS + row) * ld + e] : 0.0;
    }
    __syncthreads();
    if (!in || !lazy_col) return;
    bool touched = false;
#pragma unroll
    for (int t = 0; t < LAZY_MAX_SLOTS; ++t) {
        if (t < n_slots) {
            const double* __restrict__ hs = hsm + (t * LAZY_ROWS) * 32 + lane;
            const double inv_d = hs[(2 * NR) * 32];
            if (inv_d != 0.0) {
                double ya = y[t], yb = 0;
#pragma unroll
                for (int i = 0; i < NR; i += 2) ya -= hs[i * 32] * w[i];
#pragma unroll
                for (int i = 1; i < NR; i += 2) yb += hs[i * 32] * w[i];
                const double err = (ya - yb) * inv_d;
#pragma unroll
                for (int i = 0; i < NR; ++i) w[i] = w[i] + hs[(NR + i) * 32] * err;
                touched = true;
            }
        }
    }
    if (touched) {
#pragma unroll
        for (int i = 0; i < NR; ++i) wp[(int64_t)(i * NS) * lw] = w[i];
    }
}

// ----------------------------------------------------------------------------------------------
// Overlapped form: ONE launch holds blocks of two roles that work on the two halves of the batch --
//   role S (agents of the previous tick + simulate, one thread per episode, 128 episodes per block) on half A,
//   role L (the RLS update, 4 warps per 32 episodes, every warp a few rows with 17 loads in flight) on half B --
// so the FP64 / latency-bound work of one half and the HBM-bound work of the other share every SM (the block
// scheduler interleaves them: 1 S block per 4 L blocks in launch order).  The next launch swaps the halves; launch
// boundaries carry the dependencies (L needs S's hand-over rows, the agents need L's W).
// ----------------------------------------------------------------------------------------------
constexpr int MIX_BLOCK = 128, MIX_LWARPS = MIX_BLOCK / 32;

template <bool SYM, bool TILED>
H3_DEV void learn_role(const H6Batch& b, double gamma, int64_t blk, double* smem) {
    double (*Xs)[32] = reinterpret_cast<double (*)[32]>(smem);
    double (*cXs)[32] = reinterpret_cast<double (*)[32]>(smem + NR * 32);
    double (*Xcs)[32] = reinterpret_cast<double (*)[32]>(smem + 2 * NR * 32);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t e = blk * 32 + lane, ld = b.ld;
    const bool in = e < b.n;
    const int64_t ec = in ? e : 0;
    const int wf = in ? (int)b.work[WR_FLAGS * ld + e] : 0;
    const bool on = (wf & (WF_ACTIVE | WF_RLS)) == (WF_ACTIVE | WF_RLS);
    const double* __restrict__ wk = b.work + ec;
    const int64_t lw = TILED ? 32 : ld;
    double* __restrict__ cov = b.rls_cov + tile_off(b, ec, NR * NR);
    double* __restrict__ W = b.rls_W + tile_off(b, ec, NR * NS);
    for (int r = w; r < NR; r += MIX_LWARPS) Xs[r][lane] = on ? wk[(WR_X + r) * ld] : 0.0;
    __syncthreads();
    double cxr[(NR + MIX_LWARPS - 1) / MIX_LWARPS];
#pragma unroll
    for (int q = 0; q < (NR + MIX_LWARPS - 1) / MIX_LWARPS; ++q) {
        const int r = w + q * MIX_LWARPS;
        double cx = 0, xc = 0;
        if (r < NR && on) {
            double c[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) c[j] = (SYM && j < r) ? cov[(int64_t)(j * NR + r) * lw] : cov[(int64_t)(r * NR + j) * lw];
            double a0 = 0, a1 = 0;
#pragma unroll
            for (int j = 0; j < NR; ++j) { if (j & 1) a1 += c[j] * Xs[j][lane]; else a0 += c[j] * Xs[j][lane]; }   // model.py:53
            cx = a0 + a1;
            if (!SYM) {
#pragma unroll
                for (int i = 0; i < NR; ++i) c[i] = cov[(int64_t)(i * NR + r) * lw];
                a0 = 0; a1 = 0;
#pragma unroll
                for (int i = 0; i < NR; ++i) { if (i & 1) a1 += Xs[i][lane] * c[i]; else a0 += Xs[i][lane] * c[i]; }   // :54
                xc = a0 + a1;
            }
        }
        cxr[q] = cx;
        if (r < NR) { cXs[r][lane] = cx; if (!SYM) Xcs[r][lane] = xc; }
    }
    __syncthreads();
    if (!on) return;
    double d = gamma;                                                           // model.py:55
#pragma unroll
    for (int i = 0; i < NR; ++i) d += (SYM ? cXs[i][lane] : Xcs[i][lane]) * Xs[i][lane];
    const double inv_d = 1.0 / d, inv_gamma = 1.0 / gamma;
#pragma unroll
    for (int q = 0; q < (NR + MIX_LWARPS - 1) / MIX_LWARPS; ++q) {
        const int r = w + q * MIX_LWARPS;
        if (r >= NR) break;
        {                                                                       // model.py:59, row r
            const double gr = cxr[q] * inv_d;
            double c[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j)
                if (!SYM || j >= r) c[j] = cov[(int64_t)(r * NR + j) * lw];
#pragma unroll
            for (int j = 0; j < NR; ++j)
                if (!SYM || j >= r) cov[(int64_t)(r * NR + j) * lw] = (c[j] - gr * (SYM ? cXs[j][lane] : Xcs[j][lane])) * inv_gamma;
        }
        if (r < NS) {                                                           // model.py:49-50, :58, column r of W
            double wv[NR], ya = wk[(WR_Y + r) * ld], yb = 0;
#pragma unroll
            for (int i = 0; i < NR; ++i) wv[i] = W[(int64_t)(i * NS + r) * lw];
#pragma unroll
            for (int i = 0; i < NR; i += 2) ya -= Xs[i][lane] * wv[i];
#pragma unroll
            for (int i = 1; i < NR; i += 2) yb += Xs[i][lane] * wv[i];
            const double err = (ya - yb) * inv_d;
#pragma unroll
            for (int i = 0; i < NR; ++i) W[(int64_t)(i * NS + r) * lw] = wv[i] + cXs[i][lane] * err;
        }
    }
}

// mode_s bit 0: apply the pending agent updates of half A's previous tick; bit 1: simulate one tick; bit 2: the
// episode's updates of this call end here (refresh tau == 1 target critics).  n_s / n_l: blocks of each role.
template <bool SYM, bool TILED>
__global__ void __launch_bounds__(MIX_BLOCK, H6_SIM_MINB) mixed_kernel(const __grid_constant__ LoopParams PS, int it_s, int mode_s,
                                                                       int64_t n_s, H6Batch bl, double gamma, int64_t n_l) {
    __shared__ double smem[30 * MIX_BLOCK];                 // RK4 arrays (role S) / X, cov.X rows (role L)
    const int64_t q = blockIdx.x / 5;
    const int r5 = blockIdx.x % 5;
    if (r5 != 0) {                                          // role L: 4 of every 5 blocks
        const int64_t blk = 4 * q + (r5 - 1);
        if (blk < n_l) learn_role<SYM, TILED>(bl, gamma, blk, smem);
        return;
    }
    if (q >= n_s) return;
    const H6Config& C = PS.C;
    const H6Batch& b = PS.b;
    const int64_t e = q * MIX_BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    double* wk = b.work + e;
    if (mode_s & 1) {
        const int wf = (int)wk[WR_FLAGS * ld];
        if (wf & WF_ACTIVE) {
            const int64_t lw = TILED ? 32 : ld;
            const double* __restrict__ W = b.rls_W + tile_off(b, e, NR * NS);
            typename NetOf<TILED>::type ncol, nlon;
            ncol.v.p = b.col_net + tile_off(b, e, 11 * NH); ncol.v.ld = lw;
            nlon.v.p = b.lon_net + tile_off(b, e, 11 * NH); nlon.v.ld = lw;
            const bool fast_lon = C.lon.tau == 1.0 && (C.target_synced & 2), fast_col = C.col.tau == 1.0 && (C.target_synced & 1);
            if (wf & WF_COL) agent_learn(C.col, ncol, b.lr[e], b.lr[ld + e], W, lw, ld, wk + WR_COL * ld, fast_col);
            agent_learn(C.lon, nlon, b.lr[2 * ld + e], b.lr[3 * ld + e], W, lw, ld, wk + WR_LON * ld, fast_lon);
            if ((mode_s & 4) || b.flags[e] != 0) {
                if (fast_lon) {
#pragma unroll 4
                    for (int r = 0; r < 4 * NH; ++r) nlon.v[7 * NH + r] = nlon.v[3 * NH + r];
                }
                if (fast_col && (wf & WF_COL)) {
#pragma unroll 4
                    for (int r = 0; r < 4 * NH; ++r) ncol.v[7 * NH + r] = ncol.v[3 * NH + r];
                }
            }
        }
    }
    if (mode_s & 2) {
        if (b.flags[e] != 0) { wk[WR_FLAGS * ld] = 0.0; return; }
        tick_simulate<TILED>(PS, e, it_s, wk, smem + threadIdx.x);
    } else {
        wk[WR_FLAGS * ld] = 0.0;
    }
}

// lower triangle := upper triangle (once per call, after the last tick): thread = (episode, row)
template <bool TILED, typename ST = double>
__global__ void __launch_bounds__(LEARN_BLOCK) mirror_kernel(H6Batch b) {
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + lane, ld = b.ld;
    if (e >= b.n) return;
    const int64_t lw = TILED ? 32 : ld;
    ST* __restrict__ cov = reinterpret_cast<ST*>(b.rls_cov) + tile_off(b, e, NR * NR);
    ST c[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j)
        if (j > r) c[j] = cov[(int64_t)(r * NR + j) * lw];
#pragma unroll
    for (int j = 0; j < NR; ++j)
        if (j > r) cov[(int64_t)(j * NR + r) * lw] = c[j];
}

}  // namespace h6
