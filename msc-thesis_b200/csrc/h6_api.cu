// 6-DOF part of the C ABI (include/heli6dof.h): the fused closed-loop kernel of train.py:109-224 and the
// un-fused per-object kernels (trim, derivatives, open-loop step, PIDs, reference generator, RLS 17x15).
//
// Two paths behind h6_closed_loop: the split path (h6_split.cuh: a one-thread-per-episode simulate kernel + a
// 17-warps-per-32-episodes RLS kernel per tick, default when H6Batch.work is given) and the fused kernel below (one
// thread per episode streams its own W / covariance / networks from HBM every tick; latency-bound, kept for callers
// without a work buffer).  DESIGN.md section 9.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "h3_tick.cuh"
#include "h6_model.cuh"

namespace h3 {
void set_error(const char* fmt, ...);
}

namespace h6 {

using h3::AgentH;
using h3::NetAcc;
using h3::set_error;

constexpr int NS = H6_N_STATE, NR = H6_N_REG, NH = H6_N_HIDDEN;
constexpr int BLOCK = 128;
#ifndef H6_MINB
#define H6_MINB 4
#endif
#ifndef H6_U
#define H6_U 5
#endif
static inline unsigned grid_for(int64_t n) { return (unsigned)((n + BLOCK - 1) / BLOCK); }

static int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return H3_E_CUDA;
    }
    return H3_OK;
}

struct GVec {               // one episode's column of a `[rows][ld]` global array
    static constexpr bool kFence = false;
    double* p;
    int64_t ld;
    H3_DEV double& operator[](int i) { return p[(int64_t)i * ld]; }
};
typedef NetAcc<double, NH, GVec> Net;
struct TVec {               // one episode's column of a tiled `[tile][rows][32]` array: compile-time row stride
    static constexpr bool kFence = false;
    double* p;
    int64_t ld;             // unused (kept so both vectors are set up the same way)
    H3_DEV double& operator[](int i) { return p[i * 32]; }
};
typedef NetAcc<double, NH, TVec> NetTiled;
template <bool TILED> struct NetOf { typedef Net type; };
template <> struct NetOf<true> { typedef NetTiled type; };

// H6Batch.tile: offset of episode e's column and the row stride of a `rows`-row array (plain or tiled)
H3_DEV int64_t tile_off(const H6Batch& b, int64_t e, int rows) { return b.tile ? (e >> 5) * ((int64_t)rows * 32) + (e & 31) : e; }
H3_DEV int64_t tile_ld(const H6Batch& b) { return b.tile ? 32 : b.ld; }

// ----------------------------------------------------------------------------------------------
// RecursiveLeastSquares.update (model.py:41-59), n = 17, m = 15, on global-memory W / cov.
//   SYM: cov == cov^T is an invariant of the update in exact arithmetic (cov0 = c I); only the upper triangle
//   is read and written, cov.X and X^T.cov coincide.  Otherwise both products are formed as the reference does.
// ----------------------------------------------------------------------------------------------
template <bool SYM>
H3_DEV void rls_update(double* __restrict__ W, double* __restrict__ cov, int64_t ld, const double (&X)[NR],
                       const double (&Y)[NS], double gamma, double inv_gamma) {
    double covX[NR], Xcov[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) { covX[i] = 0; Xcov[i] = 0; }
    if (SYM) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
#pragma unroll
            for (int j = i; j < NR; ++j) {
                const double c = cov[(int64_t)(i * NR + j) * ld];
                covX[i] += c * X[j];
                if (j != i) covX[j] += c * X[i];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const double c = cov[(int64_t)(i * NR + j) * ld];
                covX[i] += c * X[j];        // model.py:53
                Xcov[j] += X[i] * c;        // model.py:54
            }
        }
    }
    double d = gamma;                        // model.py:55
#pragma unroll
    for (int i = 0; i < NR; ++i) d += (SYM ? covX[i] : Xcov[i]) * X[i];
    const double inv_d = 1.0 / d;
    // model.py:49-50, :58 column by column: each weight is loaded and stored once
#pragma unroll 1
    for (int j = 0; j < NS; ++j) {
        double w[NR], yh = 0;
#pragma unroll
        for (int i = 0; i < NR; ++i) { w[i] = W[(int64_t)(i * NS + j) * ld]; yh += X[i] * w[i]; }
        const double err = (Y[j] - yh) * inv_d;
#pragma unroll
        for (int i = 0; i < NR; ++i) W[(int64_t)(i * NS + j) * ld] = w[i] + covX[i] * err;
    }
    // model.py:59
    if (SYM) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const double gi = covX[i] * inv_d;
#pragma unroll
            for (int j = i; j < NR; ++j) {
                double* c = cov + (int64_t)(i * NR + j) * ld;
                *c = (*c - gi * covX[j]) * inv_gamma;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const double gi = covX[i] * inv_d;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                double* c = cov + (int64_t)(i * NR + j) * ld;
                *c = (*c - gi * Xcov[j]) * inv_gamma;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// fused closed loop
// ----------------------------------------------------------------------------------------------
struct LoopParams {
    H6Config C;
    ModelK K;
    H6Batch b;
    const double* excitation;
    int exc_len, n_ticks;
    H6Traj traj;
};

struct AgentIO {            // one agent's per-tick inputs, kept between get_action and update_networks
    double s0, s1, f;
    bool fwd;
};

template <typename NetT>
H3_DEV void agent_step_update(const H3AgentHyper& h, NetT& net, double lr_a, double lr_c, const double* __restrict__ W,
                              int64_t ld, const double (&obs)[NS], const double (&nxt)[NS], const double (&ref)[3],
                              const double (&nref)[3], AgentIO& io, double (&hh)[NH], double& reward, bool tau1_fast) {
    const int ac = h.ac_state, tr = h.tracked_state;
    if (!io.fwd) {          // the actor forward pass of this (obs, ref) was not needed for the action
        io.s0 = obs[ac];
        io.s1 = obs[tr] - ref_slot(ref, tr);
        (void)h3::actor_forward<double, NH, H6_U, h3::MathExp>(net, io.s0, io.s1, h.scale, h.final_activation, hh, io.f);
    }
    const double n0 = nxt[ac], n1 = nxt[tr] - ref_slot(nref, tr);
    const double et = nxt[tr] - ref_slot(ref, tr);                      // agents.py:125
    reward = -(et * et) * h.reward_weight;
    const double dr = -2 * et * h.reward_weight;
    double F[2][2], G[2];                                               // agents.py:130-140, post-update W
    const int idx[2] = {ac, tr};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = W[(int64_t)(idx[j] * NS + idx[i]) * ld];
        G[i] = W[(int64_t)((NS + h.channel) * NS + idx[i]) * ld];
    }
    AgentH<double> H;
    H.gamma = h.gamma; H.lr_a = lr_a; H.lr_c = lr_c; H.tau = h.tau; H.scale = h.scale;
    double scr[NH];
    // tau == 1: after one update the target critic IS the critic (agents.py:196), so later ticks of the launch
    // neither read nor write its rows; they are refreshed from the critic at launch exit
    if (tau1_fast)
        h3::agent_update<double, NH, H6_U, true, h3::MathExp>(net, H, h.final_activation, io.s0, io.s1, n0, n1, 0.0, dr, F, G, hh, io.f, scr);
    else
        h3::agent_update<double, NH, H6_U, false, h3::MathExp>(net, H, h.final_activation, io.s0, io.s1, n0, n1, 0.0, dr, F, G, hh, io.f, scr);
}

template <bool SYM, bool TILED>
__global__ void __launch_bounds__(BLOCK, H6_MINB) closed_loop_kernel(const __grid_constant__ LoopParams P) {
    const H6Config& C = P.C;
    const H6Batch& b = P.b;
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    int flags = b.flags[e];
    if (flags != 0) return;                                             // frozen: the episode has ended
    double obs[NS], ref[3], pid[4], gen[H6_GEN_ROWS], trim[4], lr[4], acc[3];
#pragma unroll
    for (int i = 0; i < NS; ++i) obs[i] = b.state[i * ld + e];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ref[i] = b.ref[i * ld + e]; acc[i] = b.acc[i * ld + e]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { pid[i] = b.pid[i * ld + e]; trim[i] = b.trim_act[i * ld + e]; lr[i] = b.lr[i * ld + e]; }
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) gen[i] = b.gen[i * ld + e];
    double t = b.t[e], P_out = b.P_out[e], P_avail = b.P_avail[e];
    const double phi_trim = b.phi_trim[e];
    int task = b.task[e], step = b.step[e], update_col = b.update_col[e], exc_phase = b.excitation_phase[e];
    int skip = b.rls_skip[e], term = -1;
    const int64_t lw = TILED ? 32 : b.ld;                   // row stride of W / cov / networks (a constant when tiled)
    double* __restrict__ W = b.rls_W + tile_off(b, e, NR * NS);
    double* __restrict__ cov = b.rls_cov + tile_off(b, e, NR * NR);
    typename NetOf<TILED>::type ncol, nlon;
    ncol.v.p = b.col_net + tile_off(b, e, 11 * NH); ncol.v.ld = lw;
    nlon.v.p = b.lon_net + tile_off(b, e, 11 * NH); nlon.v.ld = lw;
    const double inv_gamma = 1.0 / C.rls_gamma;
    int ticks_done = 0, col_updates = 0;

#pragma unroll 1
    for (int it = 0; it < P.n_ticks; ++it) {
        const int k = step;
        const ObsTrig g = obs_trig(obs);
        double act[4];
        pid_latped(C, obs, g, phi_trim, trim[2], trim[3], pid[1], pid[2], pid[3], act[2], act[3]);     // train.py:110
        bool both = true;
        if (C.run_mode == H6_MODE_TRAIN) {                                                                // :113-137
            if (k == 0) {
                gen_set_task(task, gen, H6_TASK_TRAIN_LON, 0.0, obs, 0.0, 0.0, 0.0);
                gen_get_ref(C, task, gen, obs, t, ref);
            } else if (k == C.exc_off_step) {
                exc_phase = 0;
            } else if (k == C.step_switch) {
                lr[2] *= C.lr_decay; lr[3] *= C.lr_decay;
                update_col = 1;
                gen_set_task(task, gen, H6_TASK_TRAIN_COL, t, obs, obs[11], 0.0, 0.0);
            }
            both = !(k < C.step_switch);
        } else if (C.run_mode == H6_MODE_TEST1) {                                                         // :138-160
            if (k == 0) { gen_set_task(task, gen, H6_TASK_HOVER, 0.0, obs, 0.0, 0.0, 0.0); gen_get_ref(C, task, gen, obs, t, ref); }
            else if (k == 500) gen_set_task(task, gen, H6_TASK_VELOCITY, t, obs, 0.0, 25 - obs[0], 0.0);
            else if (k == 2000) gen_set_task(task, gen, H6_TASK_VELOCITY, t, obs, 0.0, 0 - obs[0], 0.0);
        } else if (C.run_mode == H6_MODE_TEST2) {                                                         // :161-170
            if (k == 0) { gen_set_task(task, gen, H6_TASK_DESCENT, 0.0, obs, 0.0, 0.0, 0.0); gen_get_ref(C, task, gen, obs, t, ref); }
            else if (k == 1000) P_avail = C.P_engine_failure;
        } else {
            both = update_col != 0;
        }
        // actions (:126-137)
        AgentIO io_lon, io_col;
        double h_lon[NH], h_col[NH];
        io_lon.s0 = obs[C.lon.ac_state];
        io_lon.s1 = obs[C.lon.tracked_state] - ref_slot(ref, C.lon.tracked_state);
        io_lon.fwd = true;
        act[1] = trim[1] - 0.5 + h3::actor_forward<double, NH, H6_U, h3::MathExp>(nlon, io_lon.s0, io_lon.s1, C.lon.scale,
                                                                   C.lon.final_activation, h_lon, io_lon.f);
        io_col.fwd = both;
        if (both) {
            io_col.s0 = obs[C.col.ac_state];
            io_col.s1 = obs[C.col.tracked_state] - ref_slot(ref, C.col.tracked_state);
            act[0] = trim[0] - 0.5 + h3::actor_forward<double, NH, H6_U, h3::MathExp>(ncol, io_col.s0, io_col.s1, C.col.scale,
                                                                       C.col.final_activation, h_col, io_col.f);
        } else {
            act[0] = pid_collective(C, obs, g, trim[0], pid[0]);
        }
        if (exc_phase && P.excitation != nullptr && k < P.exc_len) {                                      // :173-174
#pragma unroll
            for (int i = 0; i < 4; ++i) act[i] += P.excitation[4 * k + i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) act[i] = clip01_nan(act[i]);                                          // :176
        // env.step (:179; heli_models.py:464-475)
        const Controls ctl = make_controls(P.K, act[0], act[1], act[2], act[3]);
        double nxt[NS];
        rk4_step(P.K, ctl, obs, P_out, P_avail, C.dt, nxt);
        const bool done = t > C.t_max;
        t += C.dt;
        gen[H6_GEN_A] = t < C.amp_t[0] ? C.amp[0] : (t < C.amp_t[1] ? C.amp[1] : C.amp[2]);              // :180-185
        double nref[3];
        gen_get_ref(C, task, gen, obs, t, nref);                                                          // :186 (pre-step obs)
        // RLS (:189; model.py:27-59)
        if (skip) {
            skip = 0;
        } else {
            double X[NR], Y[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) { X[i] = obs[i] - b.rls_x[i * ld + e]; Y[i] = nxt[i] - obs[i]; }
            X[NS] = act[0] - b.rls_u[e];
            X[NS + 1] = act[1] - b.rls_u[ld + e];
            rls_update<SYM>(W, cov, lw, X, Y, C.rls_gamma, inv_gamma);
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) b.rls_x[i * ld + e] = obs[i];
        b.rls_u[e] = act[0];
        b.rls_u[ld + e] = act[1];
        // agents (:192-201)
        double r_col = 0, r_lon = 0;
        if (update_col) {
            agent_step_update(C.col, ncol, lr[0], lr[1], W, lw, obs, nxt, ref, nref, io_col, h_col, r_col,
                              C.col.tau == 1.0 && col_updates > 0);
            col_updates += 1;
        }
        agent_step_update(C.lon, nlon, lr[2], lr[3], W, lw, obs, nxt, ref, nref, io_lon, h_lon, r_lon,
                          C.lon.tau == 1.0 && ticks_done > 0);
        acc[0] += r_col;
        acc[1] += r_lon;
        if (k >= C.win_lo && k < C.win_hi) acc[2] += r_lon;                                               // :256
        // termination (:207-218), evaluated on the pre-step observation
        const double R2D = 180.0 / 3.14159265358979323846;
        int fl = done ? H6_FLAG_TMAX : 0;
        if (fabs(obs[7] * R2D) > C.limit_pitch || fabs(obs[6] * R2D) > C.limit_roll) fl |= H6_FLAG_ENVELOPE;
        if (act[0] != act[0] || act[1] != act[1] || act[2] != act[2] || act[3] != act[3]) fl |= H6_FLAG_NAN;
        if (C.run_mode == H6_MODE_TEST2 && obs[11] > 0) fl |= H6_FLAG_LANDED;
        if (P.traj.out != nullptr && e < P.traj.n_traj) {                                                  // util.py:30-66
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
#pragma unroll
        for (int i = 0; i < NS; ++i) obs[i] = nxt[i];                                                     // :222-223
#pragma unroll
        for (int i = 0; i < 3; ++i) ref[i] = nref[i];
        step = k + 1;
        ticks_done += 1;
        if (fl) { flags = fl; term = k; break; }
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
        for (int i = 0; i < NR - 1; ++i) {                  // one row per step: its loads are issued together
            double r[NR];
#pragma unroll
            for (int j = 1; j < NR; ++j)
                if (j > i) r[j] = cov[(int64_t)(i * NR + j) * lw];
#pragma unroll
            for (int j = 1; j < NR; ++j)
                if (j > i) cov[(int64_t)(j * NR + i) * lw] = r[j];
        }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) b.state[i * ld + e] = obs[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { b.ref[i * ld + e] = ref[i]; b.acc[i * ld + e] = acc[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { b.pid[i * ld + e] = pid[i]; b.lr[i * ld + e] = lr[i]; }
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) b.gen[i * ld + e] = gen[i];
    b.t[e] = t; b.P_out[e] = P_out; b.P_avail[e] = P_avail;
    b.task[e] = task; b.step[e] = step; b.update_col[e] = update_col; b.excitation_phase[e] = exc_phase;
    b.rls_skip[e] = skip;
    if (flags) { b.flags[e] = flags; b.term_step[e] = term; }
}

}  // namespace h6

#include "h6_split.cuh"

namespace h6 {

// ----------------------------------------------------------------------------------------------
// statistics: deterministic two-stage reduction (block partials in scratch, then one block)
// ----------------------------------------------------------------------------------------------
constexpr int STAT_BLOCK = 256;

H3_DEV void stats_block_reduce(double (&v)[H6_N_STATS], double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < H6_N_STATS; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        if (lane == 0) sh[warp * H6_N_STATS + i] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < H6_N_STATS; ++i) {
            double a = lane < STAT_BLOCK / 32 ? sh[lane * H6_N_STATS + i] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            v[i] = a;
        }
    }
}

__global__ void __launch_bounds__(STAT_BLOCK) stats_partial_kernel(H6Batch b, double* partial) {
    __shared__ double sh[(STAT_BLOCK / 32) * H6_N_STATS];
    double v[H6_N_STATS];
#pragma unroll
    for (int i = 0; i < H6_N_STATS; ++i) v[i] = 0;
    for (int64_t e = (int64_t)blockIdx.x * STAT_BLOCK + threadIdx.x; e < b.n; e += (int64_t)gridDim.x * STAT_BLOCK) {
        const int fl = b.flags[e];
        v[0] += 1; v[1] += fl == 0; v[2] += (fl & H6_FLAG_TMAX) != 0; v[3] += (fl & H6_FLAG_NAN) != 0;
        v[4] += (fl & H6_FLAG_ENVELOPE) != 0; v[5] += (fl & H6_FLAG_LANDED) != 0;
        v[6] += b.acc[e]; v[7] += b.acc[b.ld + e]; v[8] += b.acc[2 * b.ld + e]; v[9] += (double)b.step[e];
    }
    stats_block_reduce(v, sh);
    if (threadIdx.x == 0)
        for (int i = 0; i < H6_N_STATS; ++i) partial[(int64_t)blockIdx.x * H6_N_STATS + i] = v[i];
}

__global__ void __launch_bounds__(STAT_BLOCK) stats_final_kernel(const double* partial, int n_partial, double* stats) {
    __shared__ double sh[(STAT_BLOCK / 32) * H6_N_STATS];
    double v[H6_N_STATS];
#pragma unroll
    for (int i = 0; i < H6_N_STATS; ++i) v[i] = 0;
    for (int p = threadIdx.x; p < n_partial; p += STAT_BLOCK)
        for (int i = 0; i < H6_N_STATS; ++i) v[i] += partial[(int64_t)p * H6_N_STATS + i];
    stats_block_reduce(v, sh);
    if (threadIdx.x == 0)
        for (int i = 0; i < H6_N_STATS; ++i) stats[i] = v[i];
}

static int stat_blocks(int64_t n) {
    int64_t g = (n + STAT_BLOCK - 1) / STAT_BLOCK;
    return (int)(g < 1 ? 1 : (g > 1024 ? 1024 : g));
}

// ----------------------------------------------------------------------------------------------
// un-fused kernels
// ----------------------------------------------------------------------------------------------
// Helicopter6DOF.trim (heli_models.py:723-816): Newton on 12 trim variables with a forward-difference
// Jacobian (delta = 1e-11) until max|f| <= 1e-8, per episode.
H3_DEV void trim_to_state(double (&st)[NS], double (&a)[4], const double (&tv)[12]) {
    st[0] = tv[0]; st[1] = tv[1]; st[2] = tv[2]; st[7] = tv[3]; st[6] = tv[4]; st[12] = tv[5]; st[13] = tv[6]; st[14] = tv[7];
    a[0] = tv[8]; a[1] = tv[9]; a[2] = tv[10]; a[3] = tv[11];
}
H3_DEV void trim_residual(const ModelK& K, const double (&st)[NS], const double (&a)[4], double vc, double vs, double P_avail,
                          double& P_out, double (&f)[12]) {
    double d[NS];
    // calculate_state_derivatives clips the controls (heli_models.py:492)
    const Controls c = make_controls(K, clip01_nan(a[0]), clip01_nan(a[1]), clip01_nan(a[2]), clip01_nan(a[3]));
    derivatives(K, c, st, P_out, P_avail, true, 0.0, d);
#pragma unroll
    for (int i = 0; i < 6; ++i) { f[i] = d[i]; f[6 + i] = d[9 + i]; }
    f[6] -= vc;
    f[8] += vs;
}

__global__ void __launch_bounds__(BLOCK) trim_kernel(ModelK K, const double* __restrict__ v, const double* __restrict__ fpa_deg,
                                                     const double* __restrict__ alt, double P_avail, double* state,
                                                     double* controls, double* P_out_o, int32_t* iters, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (e >= n) return;
    const double PI = 3.14159265358979323846;
    const double V = v[e], fpa = fpa_deg[e] * (PI / 180.0), altitude = alt[e];
    const double rho = air_density(K, -altitude);
    const double D = rho / 2 * (V * V) * (2 * K.F0_half);
    const double theta = asin(-D * cos(fpa) / K.W);
    double st[NS] = {V * cos(theta), 0, V * sin(theta), 0, 0, 0, 0, theta, 0, 0, 0, -altitude, 0.05, 0.05, 44.4};
    double tv[12] = {st[0], st[1], st[2], st[7], st[6], st[12], st[13], st[14], 0.75, 0.50, 0.5, 0.50};
    double a[4], f[12], fn[12], J[144], inc[12], P_out = 0;
    const double delta = 1e-11, vc = V * cos(fpa), vs = V * sin(fpa);
    int nn = 0;
    bool ok = false;
    while (nn < 100) {
        nn += 1;
        trim_to_state(st, a, tv);
        trim_residual(K, st, a, vc, vs, P_avail, P_out, f);
        double mx = 0;
        for (int i = 0; i < 12; ++i) mx = fmax(mx, fabs(f[i]));
        if (!(mx == mx)) break;
#pragma unroll 1
        for (int i = 0; i < 12; ++i) {
            double tp[12];
            for (int r = 0; r < 12; ++r) tp[r] = tv[r] + (r == i ? delta : 0.0);
            trim_to_state(st, a, tp);
            trim_residual(K, st, a, vc, vs, P_avail, P_out, fn);
            for (int r = 0; r < 12; ++r) J[r * 12 + i] = (fn[r] - f[r]) / delta;
        }
        for (int i = 0; i < 12; ++i) inc[i] = -f[i];
        // J inc = -f by elimination with partial pivoting (the reference applies pinv of the regular Jacobian)
#pragma unroll 1
        for (int c = 0; c < 12; ++c) {
            int p = c;
            for (int r = c + 1; r < 12; ++r)
                if (fabs(J[r * 12 + c]) > fabs(J[p * 12 + c])) p = r;
            if (p != c) {
                for (int q = 0; q < 12; ++q) { const double t = J[c * 12 + q]; J[c * 12 + q] = J[p * 12 + q]; J[p * 12 + q] = t; }
                const double t = inc[c]; inc[c] = inc[p]; inc[p] = t;
            }
            const double ip = 1.0 / J[c * 12 + c];
            for (int r = c + 1; r < 12; ++r) {
                const double m = J[r * 12 + c] * ip;
                for (int q = c; q < 12; ++q) J[r * 12 + q] -= m * J[c * 12 + q];
                inc[r] -= m * inc[c];
            }
        }
#pragma unroll 1
        for (int r = 11; r >= 0; --r) {
            double s = inc[r];
            for (int q = r + 1; q < 12; ++q) s -= J[r * 12 + q] * inc[q];
            inc[r] = s / J[r * 12 + r];
        }
        // heli_models.py:795-806: the step is added to the LAST perturbed vector (pedal + delta)
        tv[11] += delta;
        for (int i = 0; i < 12; ++i) tv[i] += inc[i];
        if (!(mx > 1e-8)) { ok = true; break; }
    }
    trim_to_state(st, a, tv);
    for (int i = 0; i < NS; ++i) state[i * ld + e] = st[i];
    for (int i = 0; i < 4; ++i) controls[i * ld + e] = a[i];
    P_out_o[e] = P_out;
    if (iters != nullptr) iters[e] = ok ? nn : -nn;
}

__global__ void __launch_bounds__(BLOCK) derivatives_kernel(ModelK K, const double* __restrict__ actions,
                                                            const double* __restrict__ state, double* P_out,
                                                            const double* __restrict__ P_avail, int trimming, double dt,
                                                            double* dots, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (e >= n) return;
    double s[NS], d[NS], po = P_out[e];
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = state[i * ld + e];
    const Controls c = make_controls(K, clip01_nan(actions[e]), clip01_nan(actions[ld + e]), clip01_nan(actions[2 * ld + e]),
                                     clip01_nan(actions[3 * ld + e]));
    derivatives(K, c, s, po, P_avail[e], trimming != 0, dt, d);
#pragma unroll
    for (int i = 0; i < NS; ++i) dots[i * ld + e] = d[i];
    P_out[e] = po;
}

__global__ void __launch_bounds__(BLOCK) env_step_kernel(H6Config C, ModelK K, H6Batch b, const double* __restrict__ actions,
                                                         const double* __restrict__ excitation, int exc_len, int n_ticks,
                                                         double* states_out, int64_t n_out, int64_t ld_out) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    int flags = b.flags != nullptr ? b.flags[e] : 0;
    if (flags) return;
    double s[NS], base[4], t = b.t[e], P_out = b.P_out[e];
    const double P_avail = b.P_avail[e];
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = b.state[i * ld + e];
#pragma unroll
    for (int i = 0; i < 4; ++i) base[i] = actions != nullptr ? actions[i * ld + e] : b.trim_act[i * ld + e];
    int step = b.step != nullptr ? b.step[e] : 0, term = -1;
#pragma unroll 1
    for (int it = 0; it < n_ticks; ++it) {
        double a[4], nw[NS];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = base[i];
            if (excitation != nullptr && step < exc_len) a[i] += excitation[4 * step + i];
            a[i] = clip01_nan(a[i]);
        }
        const Controls c = make_controls(K, a[0], a[1], a[2], a[3]);
        rk4_step(K, c, s, P_out, P_avail, C.dt, nw);
        const bool done = t > C.t_max;                                  // heli_models.py:470-473
        t += C.dt;
#pragma unroll
        for (int i = 0; i < NS; ++i) s[i] = nw[i];
        if (states_out != nullptr && e < n_out) {
#pragma unroll
            for (int i = 0; i < NS; ++i) states_out[((int64_t)it * NS + i) * ld_out + e] = s[i];
        }
        const int k = step;
        step += 1;
        if (done) { flags = H6_FLAG_TMAX; term = k; break; }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) b.state[i * ld + e] = s[i];
    b.t[e] = t;
    b.P_out[e] = P_out;
    if (b.step != nullptr) b.step[e] = step;
    if (flags && b.flags != nullptr) { b.flags[e] = flags; if (b.term_step != nullptr) b.term_step[e] = term; }
}

__global__ void __launch_bounds__(BLOCK) pid_kernel(H6Config C, H6Batch b, const double* __restrict__ obs_p, int which,
                                                    double* out0, double* out1) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    double o[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) o[i] = obs_p[i * ld + e];
    const ObsTrig g = obs_trig(o);
    if (which == 0) {
        double corr = b.pid[e];
        out0[e] = pid_collective(C, o, g, b.trim_act[e], corr);
        b.pid[e] = corr;
    } else {
        double y = b.pid[ld + e], ph = b.pid[2 * ld + e], ps = b.pid[3 * ld + e], lat, ped;
        pid_latped(C, o, g, b.phi_trim[e], b.trim_act[2 * ld + e], b.trim_act[3 * ld + e], y, ph, ps, lat, ped);
        b.pid[ld + e] = y; b.pid[2 * ld + e] = ph; b.pid[3 * ld + e] = ps;
        out0[e] = lat; out1[e] = ped;
    }
}

__global__ void __launch_bounds__(BLOCK) get_ref_kernel(H6Config C, H6Batch b, const double* __restrict__ obs_p,
                                                        const double* __restrict__ t, const double* __restrict__ amp,
                                                        double* ref_out) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    double o[NS], gen[H6_GEN_ROWS], ref[3];
#pragma unroll
    for (int i = 0; i < NS; ++i) o[i] = obs_p[i * ld + e];
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) gen[i] = b.gen[i * ld + e];
    if (amp != nullptr) gen[H6_GEN_A] = amp[e];
    gen_get_ref(C, b.task[e], gen, o, t[e], ref);
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) b.gen[i * ld + e] = gen[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ref_out[i * ld + e] = ref[i];
}

__global__ void __launch_bounds__(BLOCK) set_task_kernel(H6Batch b, int task, const double* __restrict__ t,
                                                         const double* __restrict__ obs_p, const double* __restrict__ z_start,
                                                         const double* __restrict__ target, double t_switch) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    double o[NS], gen[H6_GEN_ROWS];
#pragma unroll
    for (int i = 0; i < NS; ++i) o[i] = obs_p[i * ld + e];
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) gen[i] = b.gen[i * ld + e];
    int tk = b.task[e];
    gen_set_task(tk, gen, task, t[e], o, z_start != nullptr ? z_start[e] : 0.0, target != nullptr ? target[e] : 0.0, t_switch);
#pragma unroll
    for (int i = 0; i < H6_GEN_ROWS; ++i) b.gen[i * ld + e] = gen[i];
    b.task[e] = tk;
}

__global__ void __launch_bounds__(BLOCK) rls_update_kernel(H6Config C, H6Batch b, const double* __restrict__ state,
                                                           const double* __restrict__ action,
                                                           const double* __restrict__ next_state) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    if (!b.rls_skip[e]) {
        double X[NR], Y[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double s = state[i * ld + e];
            X[i] = s - b.rls_x[i * ld + e];
            Y[i] = next_state[i * ld + e] - s;
        }
        X[NS] = action[e] - b.rls_u[e];
        X[NS + 1] = action[ld + e] - b.rls_u[ld + e];
        rls_update<false>(b.rls_W + e, b.rls_cov + e, ld, X, Y, C.rls_gamma, 1.0 / C.rls_gamma);
    }
    b.rls_skip[e] = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) b.rls_x[i * ld + e] = state[i * ld + e];
    b.rls_u[e] = action[e];
    b.rls_u[ld + e] = action[ld + e];
}

__global__ void __launch_bounds__(BLOCK) rls_predict_kernel(H6Batch b, const double* __restrict__ state,
                                                            const double* __restrict__ action, double* out) {
    const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x, ld = b.ld;
    if (e >= b.n) return;
    double X[NR];
#pragma unroll
    for (int i = 0; i < NS; ++i) X[i] = state[i * ld + e] - b.rls_x[i * ld + e];        // model.py:65-66
    X[NS] = action[e] - b.rls_u[e];
    X[NS + 1] = action[ld + e] - b.rls_u[ld + e];
#pragma unroll 1
    for (int j = 0; j < NS; ++j) {
        double a = state[j * ld + e];
#pragma unroll
        for (int i = 0; i < NR; ++i) a += b.rls_W[(int64_t)(i * NS + j) * ld + e] * X[i];   // :67
        out[j * ld + e] = a;
    }
}

// plain `[rows][ld]` doubles <-> tiled `[tile][rows][32]` elements of type ST (double, or float: H6Batch.rls_f32)
template <typename ST>
__global__ void __launch_bounds__(256) tile_pack_kernel(double* __restrict__ plain, ST* __restrict__ tiled, int rows,
                                                        int64_t n, int64_t ld, int to_tiled) {
    const int64_t e = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    if (e >= n) return;
    const int64_t tb = (e >> 5) * ((int64_t)rows * 32) + (e & 31);
    for (int r = threadIdx.x >> 5; r < rows; r += 8) {
        if (to_tiled) tiled[tb + (int64_t)r * 32] = (ST)plain[(int64_t)r * ld + e];
        else plain[(int64_t)r * ld + e] = (double)tiled[tb + (int64_t)r * 32];
    }
}

// H6Batch of the episodes [off, off + cnt): every pointer advanced by `off` columns (off is a multiple of 32)
static H6Batch sub_batch(const H6Batch& src, int64_t off, int64_t cnt) {
    H6Batch b = src;
    const int64_t tiles = off / 32;
    auto adv = [&](double*& p, int rows) { if (p) p += b.tile ? tiles * (int64_t)rows * 32 : off; };
    double** plain[] = {&b.state, &b.t, &b.P_out, &b.P_avail, &b.ref, &b.pid, &b.gen, &b.lr, &b.rls_x, &b.rls_u, &b.acc, &b.work};
    for (double** p : plain) if (*p) *p += off;
    b.trim_act += off;
    b.phi_trim += off;
    int32_t** ints[] = {&b.task, &b.rls_skip, &b.step, &b.flags, &b.term_step, &b.update_col, &b.excitation_phase};
    for (int32_t** p : ints) if (*p) *p += off;
    adv(b.rls_W, NR * NS); adv(b.rls_cov, NR * NR); adv(b.col_net, 11 * NH); adv(b.lon_net, 11 * NH);
    b.n = cnt;
    b.hist = nullptr; b.hist_len = 0;
    return b;
}

#ifndef H6_MIX_MIN
#define H6_MIX_MIN 16384      // batches at least this large run the overlapped (mixed-role) launches
#endif

static bool check_agents(const H6Config* cfg, const char* who) {
    for (const H3AgentHyper* h : {&cfg->col, &cfg->lon}) {
        if (h->n_hidden != NH) { set_error("%s: the 6-DOF kernels are compiled for n_hidden = %d, got %d", who, NH, h->n_hidden); return false; }
        if (h->tracked_state < 0 || h->tracked_state >= NS || h->ac_state < 0 || h->ac_state >= NS || h->channel < 0 || h->channel > 1) {
            set_error("%s: agent state indices / channel out of range", who);
            return false;
        }
    }
    return true;
}

}  // namespace h6

using namespace h6;

extern "C" {

void h6_engine_power(double* out3) {
    const double Pe = 313000, etam = 0.95;
    out3[0] = 2 * Pe * etam * 0.8643;          // Paeo   heli_models.py:397
    out3[1] = 0.95 * (2 - 1) * Pe * etam;      // Poei   :399
    out3[2] = 1.10 * (2 - 1) * Pe * etam;      // Poeitr :401
}

int h6_trim(const double* v, const double* fpa_deg, const double* altitude, double P_avail, double* state, double* controls,
            double* P_out, int32_t* iters, int64_t n, int64_t ld, void* stream) {
    if (!v || !fpa_deg || !altitude || !state || !controls || !P_out || n < 0 || ld < n) { set_error("h6_trim: bad argument"); return H3_E_ARG; }
    if (n == 0) return H3_OK;
    trim_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(make_modelk(), v, fpa_deg, altitude, P_avail, state, controls,
                                                                 P_out, iters, n, ld);
    return check_launch("h6 trim_kernel");
}

int h6_env_derivatives(const double* actions, const double* state, double* P_out, const double* P_avail, int32_t trimming,
                       double dt, double* dots, int64_t n, int64_t ld, void* stream) {
    if (!actions || !state || !P_out || !P_avail || !dots || n < 0 || ld < n) { set_error("h6_env_derivatives: bad argument"); return H3_E_ARG; }
    if (n == 0) return H3_OK;
    derivatives_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(make_modelk(), actions, state, P_out, P_avail, trimming,
                                                                        dt, dots, n, ld);
    return check_launch("h6 derivatives_kernel");
}

int h6_env_step(const H6Config* cfg, const H6Batch* b, const double* actions, const double* excitation, int32_t exc_len,
                int32_t n_ticks, double* states_out, int64_t n_out, int64_t ld_out, void* stream) {
    if (!cfg || !b || !b->state || !b->t || !b->P_out || !b->P_avail || (!actions && !b->trim_act) || b->n < 0 || b->ld < b->n ||
        n_ticks < 0 || (states_out && ld_out < n_out)) {
        set_error("h6_env_step: bad argument");
        return H3_E_ARG;
    }
    if (b->n == 0 || n_ticks == 0) return H3_OK;
    env_step_kernel<<<grid_for(b->n), BLOCK, 0, (cudaStream_t)stream>>>(*cfg, make_modelk(), *b, actions, excitation, exc_len,
                                                                        n_ticks, states_out, n_out, ld_out);
    return check_launch("h6 env_step_kernel");
}

int h6_pid_collective(const H6Config* cfg, const H6Batch* b, const double* obs, double* collective_out, void* stream) {
    if (!cfg || !b || !obs || !collective_out || !b->pid || !b->trim_act || b->ld < b->n) { set_error("h6_pid_collective: bad argument"); return H3_E_ARG; }
    if (b->n == 0) return H3_OK;
    pid_kernel<<<grid_for(b->n), BLOCK, 0, (cudaStream_t)stream>>>(*cfg, *b, obs, 0, collective_out, nullptr);
    return check_launch("h6 pid_kernel");
}

int h6_pid_latped(const H6Config* cfg, const H6Batch* b, const double* obs, double* lat_out, double* ped_out, void* stream) {
    if (!cfg || !b || !obs || !lat_out || !ped_out || !b->pid || !b->trim_act || !b->phi_trim || b->ld < b->n) { set_error("h6_pid_latped: bad argument"); return H3_E_ARG; }
    if (b->n == 0) return H3_OK;
    pid_kernel<<<grid_for(b->n), BLOCK, 0, (cudaStream_t)stream>>>(*cfg, *b, obs, 1, lat_out, ped_out);
    return check_launch("h6 pid_kernel");
}

int h6_refgen_get_ref(const H6Config* cfg, const H6Batch* b, const double* obs, const double* t, const double* amplitude,
                      double* ref_out, void* stream) {
    if (!cfg || !b || !obs || !t || !ref_out || !b->gen || !b->task || b->ld < b->n) { set_error("h6_refgen_get_ref: bad argument"); return H3_E_ARG; }
    if (b->n == 0) return H3_OK;
    get_ref_kernel<<<grid_for(b->n), BLOCK, 0, (cudaStream_t)stream>>>(*cfg, *b, obs, t, amplitude, ref_out);
    return check_launch("h6 get_ref_kernel");
}

int h6_refgen_set_task(const H6Batch* b, int32_t task, const double* t, const double* obs, const double* z_start,
                       const double* filter_target, double t_switch, void* stream) {
    if (!b || !obs || !t || !b->gen || !b->task || b->ld < b->n || task < H6_TASK_HOVER || task > H6_TASK_DESCENT) {
        set_error("h6_refgen_set_task: bad argument");
        return H3_E_ARG;
    }
    if (b->n == 0) return H3_OK;
    set_task_kernel<<<grid_for(b->n), BLOCK, 0, (cudaStream_t)stream>>>(*b, task, t, obs, z_start, filter_target, t_switch);
    return check_launch("h6 set_task_kernel");
}

int h6_rls_update(const H6Config* cfg, const H6Batch* b, const double* state, const double* action, const double* next_state,
                  void* stream) {
    if (!cfg || !b || !state || !action || !next_state || !b->rls_W || !b->rls_cov || !b->rls_x || !b->rls_u || !b->rls_skip ||
        b->ld < b->n) {
        set_error("h6_rls_update: bad argument");
        return H3_E_ARG;
    }
    if (b->n == 0) return H3_OK;
    rls_update_kernel<<<grid_for(b->n), BLOCK, 0, (cudaStream_t)stream>>>(*cfg, *b, state, action, next_state);
    return check_launch("h6 rls_update_kernel");
}

int h6_rls_predict(const H6Batch* b, const double* state, const double* action, double* out, void* stream) {
    if (!b || !state || !action || !out || !b->rls_W || !b->rls_x || !b->rls_u || b->ld < b->n) { set_error("h6_rls_predict: bad argument"); return H3_E_ARG; }
    if (b->n == 0) return H3_OK;
    rls_predict_kernel<<<grid_for(b->n), BLOCK, 0, (cudaStream_t)stream>>>(*b, state, action, out);
    return check_launch("h6 rls_predict_kernel");
}

int h6_closed_loop(const H6Config* cfg, const H6Batch* b, const double* excitation, int32_t exc_len, int32_t n_ticks,
                   const H6Traj* traj, double* stats, void* stats_scratch, void* stream) {
    if (!cfg || !b) { set_error("h6_closed_loop: null config / batch"); return H3_E_ARG; }
    const void* need[] = {b->state, b->t, b->P_out, b->P_avail, b->ref, b->pid, b->gen, b->task, b->trim_act, b->phi_trim, b->lr,
                          b->rls_W, b->rls_cov, b->rls_x, b->rls_u, b->rls_skip, b->col_net, b->lon_net, b->step, b->flags,
                          b->term_step, b->update_col, b->excitation_phase, b->acc};
    for (const void* p : need)
        if (p == nullptr) { set_error("h6_closed_loop: H6Batch has a null field"); return H3_E_ARG; }
    if (b->n < 0 || b->ld < b->n || n_ticks < 0 || (excitation == nullptr && exc_len > 0) ||
        cfg->run_mode < H6_MODE_TRAIN || cfg->run_mode > H6_MODE_FREE || (b->tile != 0 && b->tile != 32) || (stats != nullptr && stats_scratch == nullptr)) {
        set_error("h6_closed_loop: bad argument");
        return H3_E_ARG;
    }
    if (!check_agents(cfg, "h6_closed_loop")) return H3_E_UNSUPPORTED;
    const bool f32 = b->rls_f32 != 0;
    if (f32 && (b->tile != 32 || b->work == nullptr || (b->hist != nullptr && b->hist_len > 0))) {
        set_error("h6_closed_loop: rls_f32 needs the split path on tiled buffers (tile = 32, work != NULL) without a history buffer");
        return H3_E_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (b->n > 0 && n_ticks > 0) {
        LoopParams P;
        P.C = *cfg; P.K = make_modelk(); P.b = *b; P.excitation = excitation; P.exc_len = exc_len; P.n_ticks = n_ticks;
        if (traj != nullptr) P.traj = *traj; else { P.traj.out = nullptr; P.traj.n_traj = 0; P.traj.ld_traj = 0; }
        if (b->work != nullptr) {
            // split path: per tick the agent updates of the previous tick, one simulate launch and one wide RLS
            // launch; a final agents launch completes the last tick.  (Running the two halves of
            // a batch on two streams to overlap the FP64-bound and the HBM-bound kernel was measured: no gain, each
            // kernel fills the SMs' registers and the other stream's blocks only run in its tail.)
            const bool sym = cfg->rls_symmetric != 0, til = b->tile != 0;
            const unsigned gs = grid_for(b->n), gl = (unsigned)((b->n + 31) / 32);
            const bool lazy_req = b->hist != nullptr && b->hist_len > 0;
            const char* mix_env = getenv("H6_MIX");                    // H6_MIX=1 opts in (not the default until measured)
            const bool mix_enabled = mix_env != nullptr && mix_env[0] == '1';
            if (mix_enabled && traj == nullptr && !lazy_req && !f32 && b->n >= H6_MIX_MIN) {
                // overlapped launches: halves A = [0, nA), B = [nA, n); every launch runs role S on one half and role L
                // on the other (mixed_kernel), alternating; 2 launches per tick
                const int64_t nA = ((b->n / 2 + 127) / 128) * 128, nB = b->n - nA;
                LoopParams PA = P, PB = P;
                PA.b = sub_batch(*b, 0, nA);
                PB.b = sub_batch(*b, nA, nB);
                auto mix = [&](const LoopParams& S, int it_s, int mode_s, bool with_s, const H6Batch& Lb, bool with_l) {
                    const int64_t n_s = with_s ? (S.b.n + MIX_BLOCK - 1) / MIX_BLOCK : 0, n_l = with_l ? (Lb.n + 31) / 32 : 0;
                    const int64_t groups = std::max<int64_t>(n_s, (n_l + 3) / 4);
                    if (groups == 0) return;
                    const unsigned g = (unsigned)(5 * groups);
                    if (sym && til) mixed_kernel<true, true><<<g, MIX_BLOCK, 0, st>>>(S, it_s, mode_s, n_s, Lb, cfg->rls_gamma, n_l);
                    else if (sym) mixed_kernel<true, false><<<g, MIX_BLOCK, 0, st>>>(S, it_s, mode_s, n_s, Lb, cfg->rls_gamma, n_l);
                    else if (til) mixed_kernel<false, true><<<g, MIX_BLOCK, 0, st>>>(S, it_s, mode_s, n_s, Lb, cfg->rls_gamma, n_l);
                    else mixed_kernel<false, false><<<g, MIX_BLOCK, 0, st>>>(S, it_s, mode_s, n_s, Lb, cfg->rls_gamma, n_l);
                };
                mix(PA, 0, 2, true, PB.b, false);                                   // S(A, 0)
                for (int it = 0; it < n_ticks; ++it) {
                    mix(PB, it, (it > 0 ? 1 : 0) | 2, true, PA.b, true);            // S(B, it) [+ agents(B, it-1)] || L(A, it)
                    const bool last = it == n_ticks - 1;
                    mix(PA, it + 1, last ? (1 | 4) : (1 | 2), true, PB.b, true);    // agents(A, it) + S(A, it+1) || L(B, it)
                }
                mix(PB, n_ticks, 1 | 4, true, PA.b, false);                         // agents(B, last)
                if (sym) { if (til) mirror_kernel<true><<<gl, LEARN_BLOCK, 0, st>>>(*b); else mirror_kernel<false><<<gl, LEARN_BLOCK, 0, st>>>(*b); }
            } else {
            const dim3 ga(gs, 2);
            // columns of W the agents read (agents.py:130-140); with a history buffer only these are updated every tick
            const unsigned live = (1u << cfg->col.ac_state) | (1u << cfg->col.tracked_state) | (1u << cfg->lon.ac_state) |
                                  (1u << cfg->lon.tracked_state);
            const bool lazy = b->hist != nullptr && b->hist_len > 0;
            const int T = lazy ? (int)(b->hist_len < LAZY_MAX_SLOTS ? b->hist_len : LAZY_MAX_SLOTS) : 1;
            if (lazy) {
                const int smem = LAZY_MAX_SLOTS * LAZY_ROWS * 32 * (int)sizeof(double);
                cudaFuncSetAttribute(lazy_w_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                cudaFuncSetAttribute(lazy_w_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            }
            for (int it = 0; it <= n_ticks; ++it) {
                if (it > 0) {
                    if (f32) agents_kernel<true, float><<<ga, 128, 0, st>>>(P, it == n_ticks);
                    else if (til) agents_kernel<true><<<ga, 128, 0, st>>>(P, it == n_ticks);
                    else agents_kernel<false><<<ga, 128, 0, st>>>(P, it == n_ticks);
                }
                if (it == n_ticks) {
                    if (sym) {
                        if (f32) mirror_kernel<true, float><<<gl, LEARN_BLOCK, 0, st>>>(*b);
                        else if (til) mirror_kernel<true><<<gl, LEARN_BLOCK, 0, st>>>(*b);
                        else mirror_kernel<false><<<gl, LEARN_BLOCK, 0, st>>>(*b);
                    }
                    break;
                }
                if (til) sim_kernel<true><<<gs, 128, 0, st>>>(P, it); else sim_kernel<false><<<gs, 128, 0, st>>>(P, it);
                const int slot = it % T;
                double* hs = lazy ? b->hist + (int64_t)slot * H6_HIST_ROWS * b->ld : nullptr;
                if (f32 && sym) learn_kernel<true, true, float><<<gl, LEARN_BLOCK, 0, st>>>(*b, cfg->rls_gamma, live, hs);
                else if (f32) learn_kernel<false, true, float><<<gl, LEARN_BLOCK, 0, st>>>(*b, cfg->rls_gamma, live, hs);
                else if (sym && til) learn_kernel<true, true><<<gl, LEARN_BLOCK, 0, st>>>(*b, cfg->rls_gamma, live, hs);
                else if (sym) learn_kernel<true, false><<<gl, LEARN_BLOCK, 0, st>>>(*b, cfg->rls_gamma, live, hs);
                else if (til) learn_kernel<false, true><<<gl, LEARN_BLOCK, 0, st>>>(*b, cfg->rls_gamma, live, hs);
                else learn_kernel<false, false><<<gl, LEARN_BLOCK, 0, st>>>(*b, cfg->rls_gamma, live, hs);
                if (lazy && (slot == T - 1 || it == n_ticks - 1)) {
                    const size_t sm = (size_t)(slot + 1) * LAZY_ROWS * 32 * sizeof(double);
                    if (til) lazy_w_kernel<true><<<gl, 32 * NS, sm, st>>>(*b, live, slot + 1);
                    else lazy_w_kernel<false><<<gl, 32 * NS, sm, st>>>(*b, live, slot + 1);
                }
            }
            }
        } else {
            const bool sym = cfg->rls_symmetric != 0, til = b->tile != 0;
            if (sym && til) closed_loop_kernel<true, true><<<grid_for(b->n), BLOCK, 0, st>>>(P);
            else if (sym) closed_loop_kernel<true, false><<<grid_for(b->n), BLOCK, 0, st>>>(P);
            else if (til) closed_loop_kernel<false, true><<<grid_for(b->n), BLOCK, 0, st>>>(P);
            else closed_loop_kernel<false, false><<<grid_for(b->n), BLOCK, 0, st>>>(P);
        }
        const int rc = check_launch("h6 closed_loop_kernel");
        if (rc != H3_OK) return rc;
    }
    if (stats != nullptr) {
        const int nb = stat_blocks(b->n);
        stats_partial_kernel<<<nb, STAT_BLOCK, 0, st>>>(*b, (double*)stats_scratch);
        stats_final_kernel<<<1, STAT_BLOCK, 0, st>>>((const double*)stats_scratch, nb, stats);
        return check_launch("h6 stats kernels");
    }
    return H3_OK;
}

int h6_tile_pack(const double* src, double* dst, int32_t rows, int64_t n, int64_t ld, int32_t to_tiled, void* stream) {
    if (!src || !dst || rows <= 0 || n < 0 || ld < n) { set_error("h6_tile_pack: bad argument"); return H3_E_ARG; }
    if (n == 0) return H3_OK;
    double* plain = const_cast<double*>(to_tiled ? src : dst);
    double* tiled = const_cast<double*>(to_tiled ? dst : src);
    tile_pack_kernel<double><<<(unsigned)((n + 31) / 32), 256, 0, (cudaStream_t)stream>>>(plain, tiled, rows, n, ld, to_tiled);
    return check_launch("h6 tile_pack_kernel");
}

int h6_tile_pack_f32(double* plain, float* tiled, int32_t rows, int64_t n, int64_t ld, int32_t to_tiled, void* stream) {
    if (!plain || !tiled || rows <= 0 || n < 0 || ld < n) { set_error("h6_tile_pack_f32: bad argument"); return H3_E_ARG; }
    if (n == 0) return H3_OK;
    tile_pack_kernel<float><<<(unsigned)((n + 31) / 32), 256, 0, (cudaStream_t)stream>>>(plain, tiled, rows, n, ld, to_tiled);
    return check_launch("h6 tile_pack_kernel");
}

int64_t h6_stats_scratch_bytes(int64_t n) { return (int64_t)stat_blocks(n) * H6_N_STATS * (int64_t)sizeof(double); }

int h6_closed_loop_info(const H6Config* cfg, int32_t* block, int32_t* blocks_per_sm, int32_t* regs, int32_t* local_bytes) {
    const bool sym = cfg != nullptr && cfg->rls_symmetric;
    const void* fn = sym ? (const void*)closed_loop_kernel<true, true> : (const void*)closed_loop_kernel<false, true>;
    cudaFuncAttributes a;
    cudaError_t err = cudaFuncGetAttributes(&a, fn);
    if (err != cudaSuccess) { set_error("cudaFuncGetAttributes: %s", cudaGetErrorString(err)); return H3_E_CUDA; }
    int nb = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, BLOCK, 0);
    if (err != cudaSuccess) { set_error("cudaOccupancyMaxActiveBlocksPerMultiprocessor: %s", cudaGetErrorString(err)); return H3_E_CUDA; }
    if (block) *block = BLOCK;
    if (blocks_per_sm) *blocks_per_sm = nb;
    if (regs) *regs = a.numRegs;
    if (local_bytes) *local_bytes = (int32_t)a.localSizeBytes;
    return H3_OK;
}

}  // extern "C"
