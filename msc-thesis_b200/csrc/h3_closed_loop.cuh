// The fused closed-loop kernel: main_3dof.py:122-197 for n_ticks ticks per launch, one thread per
// episode, all mutable episode state resident on chip between the entry load and the exit store.
//
// Data movement: every per-episode field is a `[field][ld]` SoA row, so a warp's
// load/store of one field is one fully coalesced 256 B (fp64) transaction.  HBM traffic per launch is
// 2 * S_ep per episode regardless of n_ticks; the per-tick work never touches global memory except
// for the optional trajectory rows.
//
// On-chip placement (STORE policy):
//   STORE_LOCAL : everything in per-thread arrays (registers, overflow in L1-backed local memory)
//   STORE_SMEM  : RLS covariance + weights and the collective agent's networks in a [row][BLOCK]
//                 shared-memory tile (conflict-free: consecutive threads hit consecutive words), the
//                 model state / PID / previous regressors / cyclic agent's networks in registers.
#pragma once
#include <type_traits>

#include "h3_tick.cuh"

namespace h3 {

enum { STORE_LOCAL = 0, STORE_SMEM = 1 };

constexpr int TICK_CHUNK = 32;      // ticks of reference / excitation staged through smem at a time

template <typename ET, typename NT>
struct LoopParams {
    EnvK<ET> K;
    int64_t n, ld;
    ET *state, *ref, *corr_u, *pid_corr, *rls_W, *rls_cov, *rls_x, *rls_u;
    double *t, *acc;
    const double *trim, *theta_trim, *ref_par, *col_hyper, *lon_hyper;
    int32_t *rls_skip, *step, *flags, *term_step;
    NT *col_net, *lon_net;
    const double* ref_table;
    const double* excitation;
    int32_t exc_len, n_ticks, step0, task, ref_mode, switch_step, update_lon, win_lo, win_hi;
    double ref_amp, ref_period, dt;
    H3AgentHyper col, lon;
    AgentH<NT> col_h, lon_h;    // launch-uniform hyper-parameters in network precision (HYP == false)
    ET col_wr, lon_wr;          // reward weights in env precision
    ET* traj;
    int64_t n_traj, ld_traj;
    double* stats;
    double* partials;       // [grid][H3_N_STATS]
    unsigned int* ticket;
};

template <typename NT>
H3_DEV AgentH<NT> load_hyper(const H3AgentHyper& h, const double* per_ep, int64_t ld, int64_t e, double& w_r) {
    AgentH<NT> H;
    double gamma = h.gamma, lr_a = h.lr_actor, lr_c = h.lr_critic, tau = h.tau, scale = h.scale;
    w_r = h.reward_weight;
    if (per_ep != nullptr) {
        gamma = per_ep[H3_HYPER_GAMMA * ld + e];
        lr_a = per_ep[H3_HYPER_LR_ACTOR * ld + e];
        lr_c = per_ep[H3_HYPER_LR_CRITIC * ld + e];
        tau = per_ep[H3_HYPER_TAU * ld + e];
        scale = per_ep[H3_HYPER_SCALE * ld + e];
        w_r = per_ep[H3_HYPER_REWARD_WEIGHT * ld + e];
    }
    H.gamma = (NT)gamma;
    H.lr_a = (NT)lr_a;
    H.lr_c = (NT)lr_c;
    H.tau = (NT)tau;
    H.scale = (NT)scale;
    return H;
}

// F[i][j] = W[idx_j][idx_i], G[i] = W[7+chan][idx_i], idx = (ac_state, tracked)   agents.py:137-140
template <typename ET, typename NT, typename WAcc>
H3_DEV void transition_matrices(WAcc& Wm, const H3AgentHyper& h, NT (&F)[2][2], NT (&G)[2]) {
    const int idx[2] = {h.ac_state, h.tracked_state};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = (NT)Wm.dyn(idx[j], idx[i]);
        G[i] = (NT)Wm.dyn(7 + h.channel, idx[i]);
    }
}

template <typename T, typename Vec>
struct WMatD : WMat<T, Vec> {       // + runtime-indexed read (shared memory: plain address arithmetic)
    H3_DEV T dyn(int i, int j) { return this->v[i * 7 + j]; }
};

template <typename T>
struct WMatLocal {                  // local-array flavour: runtime index goes through local memory
    T v[63];
    H3_DEV T& operator()(int i, int j) { return v[i * 7 + j]; }
    H3_DEV T dyn(int i, int j) { return v[i * 7 + j]; }
};

template <int BLOCK>
H3_DEV void block_reduce_stats(double (&v)[H3_N_STATS], double* sh /* [BLOCK/32][H3_N_STATS] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < H3_N_STATS; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
        if (lane == 0) sh[wid * H3_N_STATS + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < H3_N_STATS; ++i) {
            double a = 0.0;
            for (int w2 = 0; w2 < BLOCK / 32; ++w2) a += sh[w2 * H3_N_STATS + i];
            v[i] = a;
        }
    }
}

// Template flags
//   COL   : the launch contains ticks >= switch_step (collective agent acts and learns); false -> its
//           networks never leave HBM (phase 1 of main_3dof.py, PID collective)
//   SYM   : covariance symmetric -> upper triangle on chip (45 instead of 81 words)
//   HYP   : per-episode hyper-parameter rows present (sweeps); false -> launch-uniform values are read
//           from the constant bank and cost no registers
//   TAU1L : tau_target_critic == 1 for the cyclic agent (main_3dof.py:44) -> its target critic is not
//           kept on chip (see agent_update)
template <typename ET, typename NT, int NHL, int NHC, bool COL, bool SYM, bool HYP, bool TAU1L, int STORE,
          int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) closed_loop_kernel(const LoopParams<ET, NT> P) {
    constexpr int NCOV = SYM ? 45 : 81;
    constexpr int NNL = (TAU1L ? 7 : 11) * NHL, NNC = 11 * NHC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // ---- shared memory carve-up: staging tables first, then the per-thread tiles ----
    double* sh_ref = reinterpret_cast<double*>(smem_raw);               // [TICK_CHUNK]
    double* sh_exc = sh_ref + TICK_CHUNK;                               // [TICK_CHUNK][2]
    double* sh_red = sh_exc + 2 * TICK_CHUNK;                           // [BLOCK/32][H3_N_STATS]
    unsigned char* tile = reinterpret_cast<unsigned char*>(sh_red + (BLOCK / 32) * H3_N_STATS);

    const int tid = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * BLOCK + tid;
    const bool valid = e < P.n;
    const int64_t ld = P.ld;
    const int64_t ec = valid ? e : 0;          // clamp so that invalid lanes read in-bounds (never store)
    const EnvK<ET>& K = P.K;

    // ---- storage ----
    typedef typename std::conditional<STORE == STORE_SMEM, SmemVec<ET, NCOV, BLOCK>, LocalVec<ET, NCOV>>::type CovVec;
    typedef typename std::conditional<SYM, CovSym<ET, CovVec>, CovFull<ET, CovVec>>::type Cov;
    typedef typename std::conditional<STORE == STORE_SMEM, WMatD<ET, SmemVec<ET, 63, BLOCK>>, WMatLocal<ET>>::type WM;
    typedef typename std::conditional<STORE == STORE_SMEM, SmemVec<NT, NNC, BLOCK>, LocalVec<NT, NNC>>::type ColVec;
    Cov cov;
    WM Wm;
    NetAcc<NT, NHL, LocalVec<NT, NNL>> lon;
    NetAcc<NT, NHC, ColVec> col;
    if constexpr (STORE == STORE_SMEM) {
        ET* t_cov = reinterpret_cast<ET*>(tile);
        ET* t_w = t_cov + NCOV * BLOCK;
        NT* t_col = reinterpret_cast<NT*>(t_w + 63 * BLOCK);
        cov.v.p = t_cov + tid;
        Wm.v.p = t_w + tid;
        col.v.p = t_col + tid;      // only dereferenced when COL
    }

    // ---- entry load (coalesced SoA rows) ----
    // the z slot of the reference is -0 / +0 for every task (heli_models.py:236, :243, :250), so
    // `z - ref_z` is z bit for bit and the slot is not carried.
    ET s[7], refth, corr_u, pid_corr, xprev[7], uprev[2];
#pragma unroll
    for (int i = 0; i < 7; ++i) s[i] = P.state[i * ld + ec];
    refth = P.ref[ld + ec];
    corr_u = P.corr_u[ec];
    pid_corr = P.pid_corr[ec];
#pragma unroll
    for (int i = 0; i < 7; ++i) xprev[i] = P.rls_x[i * ld + ec];
    uprev[0] = P.rls_u[ec];
    uprev[1] = P.rls_u[ld + ec];
    int skip = P.rls_skip[ec];
    int step = P.step[ec];
    int flags = valid ? P.flags[ec] : H3_FLAG_DONE;
    int term_step = P.term_step[ec];
    double t = P.t[ec];
    double acc0 = P.acc[ec], acc1 = P.acc[ld + ec], acc2 = P.acc[2 * ld + ec];
    const ET trim0 = (ET)P.trim[ec], trim1 = (ET)P.trim[ld + ec];
    AgentH<NT> Hcol_e, Hlon_e;      // only live when HYP
    ET wr_col_e = (ET)0, wr_lon_e = (ET)0;
    if (HYP) {
        double wc, wl;
        Hcol_e = load_hyper<NT>(P.col, P.col_hyper, ld, ec, wc);
        Hlon_e = load_hyper<NT>(P.lon, P.lon_hyper, ld, ec, wl);
        wr_col_e = (ET)wc;
        wr_lon_e = (ET)wl;
    }
    const AgentH<NT>& Hcol = HYP ? Hcol_e : P.col_h;
    const AgentH<NT>& Hlon = HYP ? Hlon_e : P.lon_h;
    const ET wr_col = HYP ? wr_col_e : P.col_wr;
    const ET wr_lon = HYP ? wr_lon_e : P.lon_wr;
    if (SYM) {
        // upper triangle only
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) cov.v[CovSym<ET, CovVec>::idx(i, j)] = P.rls_cov[(i * 9 + j) * ld + ec];
    } else {
#pragma unroll
        for (int i = 0; i < 81; ++i) cov.v[i] = P.rls_cov[i * ld + ec];
    }
#pragma unroll
    for (int i = 0; i < 63; ++i) Wm(i / 7, i % 7) = P.rls_W[i * ld + ec];
    lon.v.load(P.lon_net + ec, ld);         // TAU1L: rows [0, 7 nh) only (actor + critic)
    if (COL) col.v.load(P.col_net + ec, ld);

    const bool lockstep = P.step0 >= 0;
    const bool in_kernel_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_KERNEL);
    const bool table_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_TABLE);

    // ---- tick loop, staged in chunks ----
    for (int c0 = 0; c0 < P.n_ticks; c0 += TICK_CHUNK) {
        const int cn = min(TICK_CHUNK, P.n_ticks - c0);
        if (lockstep) {
            __syncthreads();
            for (int i = tid; i < cn; i += BLOCK) {
                if (table_ref) sh_ref[i] = P.ref_table[c0 + i];
                const int kk = P.step0 + c0 + i;
                const bool on = (P.excitation != nullptr) && (kk < P.exc_len);
                sh_exc[2 * i] = on ? P.excitation[2 * kk] : 0.0;
                sh_exc[2 * i + 1] = on ? P.excitation[2 * kk + 1] : 0.0;
            }
            __syncthreads();
        }
        for (int jj = 0; jj < cn; ++jj) {
            if (flags != 0) continue;       // terminated (or padding) lane: frozen
            const int k = lockstep ? (P.step0 + c0 + jj) : step;
            const bool phase2 = COL && (k >= P.switch_step);                    // main_3dof.py:127-137
            // -------- actions --------
            ET sth, cth;
            Math<ET>::sincos_(s[4], &sth, &cth);
            NT hl[NHL], fl;
            const NT sl0 = (NT)pick7(s, P.lon.ac_state);
            const NT sl1 = (NT)(P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - refth);
            const NT a_lon = actor_forward<NT, NHL>(lon, sl0, sl1, Hlon.scale, P.lon.final_activation, hl, fl);
            ET cyc = trim1 + (ET)a_lon;
            NT hc[NHC], fc = (NT)0, sc0 = (NT)0, sc1 = (NT)0;
            ET colv;
            if (phase2) {
                sc0 = (NT)pick7(s, P.col.ac_state);
                sc1 = (NT)(P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - refth);
                const NT a_col = actor_forward<NT, NHC>(col, sc0, sc1, Hcol.scale, P.col.final_activation, hc, fc);
                colv = trim0 + (ET)a_col;
            } else {
                colv = pid_collective(K, s, sth, cth, pid_corr);                // PID.py:16-24
            }
            if (lockstep) {                                                     // main_3dof.py:139-140
                colv += (ET)sh_exc[2 * jj];
                cyc += (ET)sh_exc[2 * jj + 1];
            } else if (P.excitation != nullptr && k < P.exc_len) {
                colv += (ET)P.excitation[2 * k];
                cyc += (ET)P.excitation[2 * k + 1];
            }
            // -------- env.step (main_3dof.py:141) --------
            ET obs[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) obs[i] = s[i];
            if (P.task == H3_TASK_VELOCITY) (void)velocity_ref(K, s[2], corr_u);     // heli_models.py:155
            bool done;
            env_step(K, s, colv, cyc, sth, cth, done);
            t += P.dt;                                                              // heli_models.py:206
            // -------- next_ref = env.get_ref() (main_3dof.py:142) --------
            ET nrefth;
            if (table_ref) {
                nrefth = (ET)sh_ref[jj];
            } else if (in_kernel_ref) {
                double amp = P.ref_amp, period = P.ref_period;
                if (P.ref_par != nullptr) {         // re-read per tick (L2-resident) instead of pinning registers
                    amp = __ldcg(P.ref_par + ec);
                    period = __ldcg(P.ref_par + ld + ec);
                }
                nrefth = (ET)sinusoid_ref(t + P.dt, amp, period);
            } else if (P.task == H3_TASK_VELOCITY) {
                nrefth = velocity_ref(K, s[2], corr_u);
            } else {
                nrefth = stop_over_point_ref(K, s[0], s[2], (ET)__ldcg(P.theta_trim + ec));
            }
            // -------- RLS.update(obs, actions, next_obs) (main_3dof.py:145, model.py:27-59) --------
            if (skip) {
                skip = 0;
            } else {
                ET X[9], Y[7];
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    X[i] = obs[i] - xprev[i];
                    Y[i] = s[i] - obs[i];
                }
                X[7] = colv - uprev[0];
                X[8] = cyc - uprev[1];
                rls_update(K, Wm, cov, X, Y);
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) xprev[i] = obs[i];
            uprev[0] = colv;
            uprev[1] = cyc;
            // -------- cyclic agent (main_3dof.py:148-151) --------
            double r_lon = 0.0, r_col = 0.0;
            if (P.update_lon) {
                const ET e_tr = P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - refth;
                r_lon = (double)(-(e_tr * e_tr) * wr_lon);                          // agents.py:126
                const NT dr = (NT)((ET)-2 * e_tr * wr_lon);                         // agents.py:127
                NT F[2][2], G[2];
                transition_matrices<ET, NT>(Wm, P.lon, F, G);
                const NT n0 = (NT)pick7(s, P.lon.ac_state);
                const NT n1 = (NT)(P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - nrefth);
                agent_update<NT, NHL, TAU1L>(lon, Hlon, P.lon.final_activation, sl0, sl1, n0, n1, (NT)0, dr, F, G, hl, fl);
            }
            // -------- collective agent (main_3dof.py:156-159) --------
            if (phase2) {
                const ET e_tr = P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - refth;
                r_col = (double)(-(e_tr * e_tr) * wr_col);
                const NT dr = (NT)((ET)-2 * e_tr * wr_col);
                NT F[2][2], G[2];
                transition_matrices<ET, NT>(Wm, P.col, F, G);
                const NT n0 = (NT)pick7(s, P.col.ac_state);
                const NT n1 = (NT)(P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - nrefth);
                agent_update<NT, NHC, false>(col, Hcol, P.col.final_activation, sc0, sc1, n0, n1, (NT)0, dr, F, G, hc, fc);
            }
            // -------- log row (main_3dof.py:164-175): pre-step obs and ref, applied actions, rewards --------
            if (P.traj != nullptr && e < P.n_traj) {
                ET* row = P.traj + ((int64_t)(c0 + jj) * H3_TRAJ_ROWS) * P.ld_traj + e;
#pragma unroll
                for (int i = 0; i < 7; ++i) row[i * P.ld_traj] = obs[i];
                row[7 * P.ld_traj] = refth;
                row[8 * P.ld_traj] = colv;
                row[9 * P.ld_traj] = cyc;
                row[10 * P.ld_traj] = (ET)r_col;
                row[11 * P.ld_traj] = (ET)r_lon;
                row[12 * P.ld_traj] = (ET)1;
            }
            acc0 += r_col;
            acc1 += r_lon;
            if (k >= P.win_lo && k < P.win_hi) acc2 += r_lon;
            // -------- next step / termination (main_3dof.py:189-197) --------
            refth = nrefth;
            step = k + 1;
            const bool nan_act = (colv != colv) || (cyc != cyc);
            const int nf = (done ? H3_FLAG_DONE : 0) | (nan_act ? H3_FLAG_NAN : 0);
            if (nf) {
                flags = nf;
                term_step = k;
            }
        }
    }

    // ---- exit store ----
    if (valid) {
#pragma unroll
        for (int i = 0; i < 7; ++i) P.state[i * ld + e] = s[i];
        P.ref[ld + e] = refth;
        P.corr_u[e] = corr_u;
        P.pid_corr[e] = pid_corr;
#pragma unroll
        for (int i = 0; i < 7; ++i) P.rls_x[i * ld + e] = xprev[i];
        P.rls_u[e] = uprev[0];
        P.rls_u[ld + e] = uprev[1];
        P.rls_skip[e] = skip;
        P.step[e] = step;
        P.flags[e] = flags;
        P.term_step[e] = term_step;
        P.t[e] = t;
        P.acc[e] = acc0;
        P.acc[ld + e] = acc1;
        P.acc[2 * ld + e] = acc2;
        if (SYM) {
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = i; j < 9; ++j) {
                    const ET c = cov.v[CovSym<ET, CovVec>::idx(i, j)];
                    P.rls_cov[(i * 9 + j) * ld + e] = c;
                    if (j != i) P.rls_cov[(j * 9 + i) * ld + e] = c;
                }
        } else {
#pragma unroll
            for (int i = 0; i < 81; ++i) P.rls_cov[i * ld + e] = cov.v[i];
        }
#pragma unroll
        for (int i = 0; i < 63; ++i) P.rls_W[i * ld + e] = Wm(i / 7, i % 7);
        lon.v.store(P.lon_net + e, ld);
        if (TAU1L) {                // target critic == critic (agents.py:196 with tau = 1)
#pragma unroll
            for (int i = 0; i < 4 * NHL; ++i) P.lon_net[(int64_t)(7 * NHL + i) * ld + e] = lon.v[3 * NHL + i];
        }
        if (COL) col.v.store(P.col_net + e, ld);
    }

    // ---- statistics epilogue: block partials -> last block sums them in a fixed order ----
    if (P.stats != nullptr) {
        double v[H3_N_STATS];
        v[0] = valid ? 1.0 : 0.0;
        v[1] = (valid && flags == 0) ? 1.0 : 0.0;
        v[2] = (valid && (flags & H3_FLAG_DONE)) ? 1.0 : 0.0;
        v[3] = (valid && (flags & H3_FLAG_NAN)) ? 1.0 : 0.0;
        v[4] = valid ? acc0 : 0.0;
        v[5] = valid ? acc1 : 0.0;
        v[6] = valid ? acc2 : 0.0;
        v[7] = valid ? (double)step : 0.0;
        __syncthreads();
        block_reduce_stats<BLOCK>(v, sh_red);
        __shared__ bool is_last;
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < H3_N_STATS; ++i) P.partials[(int64_t)blockIdx.x * H3_N_STATS + i] = v[i];
            __threadfence();
            const unsigned int done_blocks = atomicAdd(P.ticket, 1u);
            is_last = (done_blocks == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            double a[H3_N_STATS];
#pragma unroll
            for (int i = 0; i < H3_N_STATS; ++i) a[i] = 0.0;
            for (int64_t b = tid; b < (int64_t)gridDim.x; b += BLOCK)
#pragma unroll
                for (int i = 0; i < H3_N_STATS; ++i) a[i] += __ldcg(&P.partials[b * H3_N_STATS + i]);
            __syncthreads();
            block_reduce_stats<BLOCK>(a, sh_red);
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < H3_N_STATS; ++i) P.stats[i] = a[i];
                *P.ticket = 0u;     // ready for the next launch
            }
        }
    }
}

template <typename ET, int NCOV>
constexpr size_t loop_smem_bytes(int block, int store, bool col, int nhc, size_t net_size) {
    size_t b = (size_t)(3 * TICK_CHUNK + (block / 32) * H3_N_STATS) * sizeof(double);
    if (store == STORE_SMEM) {
        b += (size_t)block * (NCOV + 63) * sizeof(ET);
        if (col) b += (size_t)block * 11 * nhc * net_size;
    }
    return b;
}

}  // namespace h3
