// The fused closed-loop kernel: main_3dof.py:122-197 for n_ticks ticks per launch, one thread per
// episode, all mutable episode state resident on chip between the entry load and the exit store.
//
// Data movement: every per-episode field is a `[field][ld]` SoA row, so a warp's
// load/store of one field is one fully coalesced transaction.  HBM traffic per launch is 2 * S_ep per
// episode regardless of n_ticks; the per-tick work never touches global memory except for the
// optional trajectory rows.
//
// On-chip placement: an episode's ~2.3-2.7 KB of fp64 state does not fit a thread's 255 registers, and
// 128 threads x that state is more than the 227 KB of shared memory -- registers AND shared memory
// together hold it.  Each warp owns a [row][LANES] shared-memory tile (conflict-free: consecutive lanes hit
// consecutive words); the PLACE mask says which groups live in the tile, the rest stays in registers.
// LANES <= 32 episodes per warp: trading a few idle lanes for shared-memory room per episode keeps four
// warps (one per SM sub-partition) resident without a single local-memory spill.
#pragma once
#include <type_traits>

#include "h3_tick.cuh"

namespace h3 {

// PLACE bits: group lives in the shared-memory tile (always there: RLS covariance + weights, collective nets)
enum { P_LON = 1, P_PREV = 2, P_MISC = 4, P_STASHC = 8, P_STASHL = 16 };

// (H3Batch.heat counted the ticks an episode spent on data-dependent slow branches of tanh / sincos; the tick has none
// left, the kernels write 0.)

constexpr int TICK_CHUNK = 16;      // ticks of reference / excitation staged through smem at a time

template <typename ET, typename NT>
struct LoopParams {
    EnvK<ET> K;
    int64_t n, ld;
    ET *state, *ref, *corr_u, *pid_corr, *rls_W, *rls_cov, *rls_x, *rls_u;
    double *t, *acc;
    const double *trim, *theta_trim, *ref_par, *col_hyper, *lon_hyper;
    int32_t *rls_skip, *step, *flags, *term_step, *heat;
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

template <typename NT, bool CG = false>
H3_DEV AgentH<NT> load_hyper(const H3AgentHyper& h, const double* per_ep, int64_t ld, int64_t e, double& w_r) {
    AgentH<NT> H;
    double gamma = h.gamma, lr_a = h.lr_actor, lr_c = h.lr_critic, tau = h.tau, scale = h.scale;
    w_r = h.reward_weight;
    if (per_ep != nullptr) {
        auto rd = [&](int row) { return CG ? __ldcg(per_ep + row * ld + e) : per_ep[row * ld + e]; };
        gamma = rd(H3_HYPER_GAMMA);
        lr_a = rd(H3_HYPER_LR_ACTOR);
        lr_c = rd(H3_HYPER_LR_CRITIC);
        tau = rd(H3_HYPER_TAU);
        scale = rd(H3_HYPER_SCALE);
        w_r = rd(H3_HYPER_REWARD_WEIGHT);
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


template <int NWARP>
H3_DEV void block_reduce_stats(double (&v)[H3_N_STATS], double* sh /* [NWARP][H3_N_STATS] */) {
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
            for (int w2 = 0; w2 < NWARP; ++w2) a += sh[w2 * H3_N_STATS + i];
            v[i] = a;
        }
    }
}

template <typename T, typename Vec>
struct WMatD : WMat<T, Vec> {};     // shared-memory flavour (runtime-indexed reads are plain address arithmetic)

// view of a group starting at element `off`
template <typename Vec, typename T>
struct SubVec {
    Vec& v;
    int off;
    H3_DEV T& operator[](int i) { return v[off + i]; }
};

// group storage: a strided column of the warp's tile, or registers
template <typename T, int N, int LANES, bool IN_SMEM>
struct Group : std::conditional<IN_SMEM, SmemVec<T, N, LANES>, LocalVec<T, N>>::type {};

template <typename ET, typename NT, int NHL, int NHC, bool COL, bool SYM, bool HYP, bool TAU1L, int PLACE, int LANES>
struct LoopLayout {
    static constexpr int NCOV = SYM ? 45 : 81;
    static constexpr int NNL = (TAU1L ? 7 : 11) * NHL, NNC = 11 * NHC;
    // stash: forward activations kept from get_action to the update (h[NH], f, s0, s1) + NH words of
    // scratch for the critic's hidden layer between the two passes of agent_update
    static constexpr int N_STASH = ((PLACE & P_STASHL) ? 2 * NHL + 3 : 0) + ((COL && (PLACE & P_STASHC)) ? 2 * NHC + 3 : 0);
    // bytes of one episode in the tile
    static constexpr size_t bytes_per_lane =
        sizeof(ET) * (NCOV + 63) + (COL ? sizeof(NT) * NNC : 0) + ((PLACE & P_LON) ? sizeof(NT) * NNL : 0) +
        ((PLACE & P_PREV) ? sizeof(ET) * 9 : 0) + ((PLACE & P_MISC) ? sizeof(double) * 4 + sizeof(ET) * 4 : 0) +
        sizeof(NT) * N_STASH;
    static constexpr size_t tile_bytes = ((bytes_per_lane * LANES + 15) / 16) * 16;
    __host__ __device__ static constexpr size_t head_bytes(int warps) {
        return sizeof(double) * warps * (3 * TICK_CHUNK + H3_N_STATS);
    }
    __host__ __device__ static constexpr size_t smem_bytes(int warps) {
        return head_bytes(warps) + tile_bytes * warps;
    }
};

// Template flags
//   COL   : the launch contains ticks >= switch_step (collective agent acts and learns); false -> its
//           networks never leave HBM (phase 1 of main_3dof.py, PID collective)
//   SYM   : covariance symmetric -> upper triangle on chip (45 instead of 81 words)
//   HYP   : per-episode hyper-parameter rows present (sweeps), re-read from L2 every tick; false -> launch-uniform values are read
//           from the constant bank and cost neither registers nor shared memory
//   TAU1L : tau_target_critic == 1 for the cyclic agent (main_3dof.py:44) -> its target critic is not
//           kept on chip (see agent_update)
//   PLACE : which optional groups live in shared memory (P_* bits); LANES episodes per warp; WARPS per block
//   UL/UC : hidden units per chunk in the cyclic / collective agent's loops (see h3_tick.cuh); a chunked
//           agent (U < NH) needs its networks and stash in shared memory
//   RD    : look-ahead depth (groups) of the RLS shared-memory streams
//   MINB  : minimum resident CTAs per SM the register allocation has to allow
template <typename ET, typename NT, int NHL, int NHC, bool COL, bool SYM, bool HYP, bool TAU1L, int PLACE, int LANES,
          int WARPS, int UL, int UC, int RD, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) closed_loop_kernel(const LoopParams<ET, NT> P) {
    static_assert(UL == NHL || ((PLACE & P_LON) && (PLACE & P_STASHL)), "chunked cyclic agent must live in smem");
    static_assert(UC == NHC || (PLACE & P_STASHC), "chunked collective agent needs its stash in smem");
    typedef LoopLayout<ET, NT, NHL, NHC, COL, SYM, HYP, TAU1L, PLACE, LANES> LY;
    constexpr int NCOV = LY::NCOV, NNC = LY::NNC;
    constexpr int NNL = LY::NNL;
    constexpr int BLOCK = WARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // reference / excitation slices are staged per warp (no CTA barrier in the tick loop: the warps of a CTA drift apart)
    double* sh_ref = reinterpret_cast<double*>(smem_raw) + warp * 3 * TICK_CHUNK;       // [TICK_CHUNK]
    double* sh_exc = sh_ref + TICK_CHUNK;                                               // [TICK_CHUNK][2]
    double* sh_red = reinterpret_cast<double*>(smem_raw) + WARPS * 3 * TICK_CHUNK;      // [WARPS][H3_N_STATS]

    const int sub = lane;                               // episode slot inside the warp
    const int64_t e = ((int64_t)blockIdx.x * WARPS + warp) * LANES + sub;
    const bool valid = (sub < LANES) && (e < P.n);
    const int64_t ld = P.ld;
    const int64_t ec = valid ? e : 0;          // idle lanes read episode 0 (in bounds) and never store
    const int lc = sub < LANES ? sub : 0;       // ... and alias slot 0's tile column (never written: flags != 0)
    const EnvK<ET>& K = P.K;
    Math<NT>::prepare();                        // tanh table -> (static) shared memory

    // ---- carve the warp's tile ----
    unsigned char* tp = smem_raw + LY::head_bytes(WARPS) + LY::tile_bytes * warp;
    typedef SmemVec<ET, NCOV, LANES> CovVec;
    typename std::conditional<SYM, CovSym<ET, CovVec>, CovFull<ET, CovVec>>::type cov;
    WMatD<ET, SmemVec<ET, 63, LANES>> Wm;
    NetAcc<NT, NHC, SmemVec<NT, NNC, LANES>> col;
    NetAcc<NT, NHL, Group<NT, NNL, LANES, (PLACE & P_LON) != 0>> lon;
    Group<ET, 9, LANES, (PLACE & P_PREV) != 0> prev;             // x_prev[7], u_prev[2]   model.py:45-46
    Group<double, 4, LANES, (PLACE & P_MISC) != 0> miscd;        // acc0, acc1, acc2, t
    Group<ET, 4, LANES, (PLACE & P_MISC) != 0> misce;            // corr_u, pid_corr, trim0, trim1
    Group<NT, 2 * NHL + 3, LANES, (PLACE & P_STASHL) != 0> stl; // hl[NHL], fl, sl0, sl1, scratch[NHL]
    Group<NT, 2 * NHC + 3, LANES, (PLACE & P_STASHC) != 0> stc; // hc[NHC], fc, sc0, sc1, scratch[NHC]
    {
        // widest element type first keeps every column naturally aligned
        double* pd = reinterpret_cast<double*>(tp);
        if constexpr ((PLACE & P_MISC) != 0) { miscd.p = pd + lc; pd += 4 * LANES; }
        ET* pe = reinterpret_cast<ET*>(pd);
        cov.v.p = pe + lc; pe += NCOV * LANES;
        Wm.v.p = pe + lc; pe += 63 * LANES;
        if constexpr ((PLACE & P_PREV) != 0) { prev.p = pe + lc; pe += 9 * LANES; }
        if constexpr ((PLACE & P_MISC) != 0) { misce.p = pe + lc; pe += 4 * LANES; }
        NT* pn = reinterpret_cast<NT*>(pe);
        col.v.p = pn + lc;                      // only dereferenced when COL
        if constexpr (COL) pn += NNC * LANES;
        if constexpr ((PLACE & P_LON) != 0) { lon.v.p = pn + lc; pn += NNL * LANES; }
        if constexpr ((PLACE & P_STASHL) != 0) { stl.p = pn + lc; pn += (2 * NHL + 3) * LANES; }
        if constexpr ((PLACE & P_STASHC) != 0) {
            stc.p = pn + lc;                    // only dereferenced when COL
            if constexpr (COL) pn += (2 * NHC + 3) * LANES;
        }
    }

    // ---- entry load (coalesced SoA rows) ----
    // the z slot of the reference is -0 / +0 for every task (heli_models.py:236, :243, :250), so
    // `z - ref_z` is z bit for bit and the slot is not carried.
    ET s[7], refth;
#pragma unroll
    for (int i = 0; i < 7; ++i) s[i] = P.state[i * ld + ec];
    refth = P.ref[ld + ec];
    int skip = P.rls_skip[ec];
    int step = P.step[ec];
    int flags = valid ? P.flags[ec] : H3_FLAG_DONE;
    int term_step = P.term_step[ec];
    // shared-memory columns are written by their owning lane only (idle lanes alias lane 0's column)
    if (valid || (PLACE & P_PREV) == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) prev[i] = P.rls_x[i * ld + ec];
        prev[7] = P.rls_u[ec];
        prev[8] = P.rls_u[ld + ec];
    }
    if (valid || (PLACE & P_MISC) == 0) {
        miscd[0] = P.acc[ec];
        miscd[1] = P.acc[ld + ec];
        miscd[2] = P.acc[2 * ld + ec];
        miscd[3] = P.t[ec];
        misce[0] = P.corr_u[ec];
        misce[1] = P.pid_corr[ec];
        misce[2] = (ET)P.trim[ec];
        misce[3] = (ET)P.trim[ld + ec];
    }
    if (valid) {
        if (SYM) {
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = i; j < 9; ++j) cov.v[CovSym<ET, CovVec>::idx(i, j)] = P.rls_cov[(i * 9 + j) * ld + ec];
        } else {
#pragma unroll
            for (int i = 0; i < 81; ++i) cov.v[i] = P.rls_cov[i * ld + ec];
        }
#pragma unroll
        for (int i = 0; i < 63; ++i) Wm.v[i] = P.rls_W[i * ld + ec];
        if (COL) col.v.load(P.col_net + ec, ld);
    }
    if constexpr ((PLACE & P_LON) != 0) {
        if (valid) lon.v.load(P.lon_net + ec, ld);                      // TAU1L: rows [0, 7 nh) only
    } else {
        lon.v.load(P.lon_net + ec, ld);
    }

    const bool lockstep = P.step0 >= 0;
    const bool in_kernel_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_KERNEL);
    const bool table_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_TABLE);

    // ---- tick loop, staged in chunks ----
    for (int c0 = 0; c0 < P.n_ticks; c0 += TICK_CHUNK) {
        const int cn = min(TICK_CHUNK, P.n_ticks - c0);
        if (lockstep) {
            static_assert(TICK_CHUNK <= 32, "one lane stages one tick");
            __syncwarp();
            if (lane < cn) {
                if (table_ref) sh_ref[lane] = P.ref_table[c0 + lane];
                const int kk = P.step0 + c0 + lane;
                const bool on = (P.excitation != nullptr) && (kk < P.exc_len);
                sh_exc[2 * lane] = on ? P.excitation[2 * kk] : 0.0;
                sh_exc[2 * lane + 1] = on ? P.excitation[2 * kk + 1] : 0.0;
            }
            __syncwarp();
        }
        for (int jj = 0; jj < cn; ++jj) {
            if (flags != 0) continue;       // terminated (or idle) lane: frozen
            const int k = lockstep ? (P.step0 + c0 + jj) : step;
            const bool phase2 = COL && (k >= P.switch_step);                    // main_3dof.py:127-137
            AgentH<NT> Hcol = P.col_h, Hlon = P.lon_h;
            ET wr_col = P.col_wr, wr_lon = P.lon_wr;
            if constexpr (HYP) {            // read-only rows, L2-resident: re-read per tick instead of holding tile space
                                            // (one more CTA per SM in the launches before the switch: 6.5 % faster)
                double wc, wl;
                Hcol = load_hyper<NT, true>(P.col, P.col_hyper, ld, ec, wc);
                Hlon = load_hyper<NT, true>(P.lon, P.lon_hyper, ld, ec, wl);
                wr_col = (ET)wc;
                wr_lon = (ET)wl;
            }
            // -------- actions (main_3dof.py:132-137); forward activations are kept for the update --------
            ET sth, cth;
            ET cyc, colv;
            {
                NT fl;
                const NT sl0 = (NT)pick7(s, P.lon.ac_state);
                const NT sl1 = (NT)(P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - refth);
                const NT a_lon = actor_forward<NT, NHL, UL>(lon, sl0, sl1, Hlon.scale, P.lon.final_activation, stl, fl);
                cyc = misce[3] + (ET)a_lon;
                stl[NHL] = fl;
                stl[NHL + 1] = sl0;
                stl[NHL + 2] = sl1;
            }
            if (phase2) {
                NT fc;
                const NT sc0 = (NT)pick7(s, P.col.ac_state);
                const NT sc1 = (NT)(P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - refth);
                const NT a_col = actor_forward<NT, NHC, UC>(col, sc0, sc1, Hcol.scale, P.col.final_activation, stc, fc);
                colv = misce[2] + (ET)a_col;
                stc[NHC] = fc;
                stc[NHC + 1] = sc0;
                stc[NHC + 2] = sc1;
            }
            ET exc0 = (ET)0, exc1 = (ET)0;                                      // main_3dof.py:139-140
            if (lockstep) {
                exc0 = (ET)sh_exc[2 * jj];
                exc1 = (ET)sh_exc[2 * jj + 1];
            } else if (P.excitation != nullptr && k < P.exc_len) {
                exc0 = (ET)P.excitation[2 * k];
                exc1 = (ET)P.excitation[2 * k + 1];
            }
            cyc += exc1;
            // sin/cos of theta (model + PID) and of the cyclic pitch (model) in one interleaved evaluation
            ET scy, ccy;
            Math<ET>::sincos2_(s[4], cyc, &sth, &cth, &scy, &ccy);
            if (!phase2) {
                ET pc = misce[1];
                colv = pid_collective(K, s, sth, cth, pc);                      // PID.py:16-24
                misce[1] = pc;
            }
            colv += exc0;
            // -------- env.step (main_3dof.py:141) --------
            ET X[9], Y[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                Y[i] = s[i];                        // pre-step obs
                X[i] = s[i] - prev[i];              // model.py:41
                prev[i] = s[i];                     // model.py:45
            }
            X[7] = colv - prev[7];                  // model.py:42
            X[8] = cyc - prev[8];
            prev[7] = colv;                         // model.py:46
            prev[8] = cyc;
            ET cu = misce[0];
            if (P.task == H3_TASK_VELOCITY) (void)velocity_ref(K, s[2], cu);    // heli_models.py:155
            bool done;
            env_step(K, s, colv, cyc, sth, cth, scy, ccy, done);
            const double t_new = miscd[3] + P.dt;                                // heli_models.py:206
            miscd[3] = t_new;
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
                nrefth = (ET)sinusoid_ref(t_new + P.dt, amp, period);
            } else if (P.task == H3_TASK_VELOCITY) {
                nrefth = velocity_ref(K, s[2], cu);
            } else {
                nrefth = stop_over_point_ref(K, s[0], s[2], (ET)__ldcg(P.theta_trim + ec));
            }
            if (P.task == H3_TASK_VELOCITY) misce[0] = cu;
            // -------- log row (main_3dof.py:164-175): pre-step obs and ref, applied actions; rewards below ----
            const bool rec = P.traj != nullptr && e < P.n_traj;
            ET* row = rec ? P.traj + ((int64_t)(c0 + jj) * H3_TRAJ_ROWS) * P.ld_traj + e : nullptr;
            if (rec) {
#pragma unroll
                for (int i = 0; i < 7; ++i) row[i * P.ld_traj] = Y[i];
                row[7 * P.ld_traj] = refth;
                row[8 * P.ld_traj] = colv;
                row[9 * P.ld_traj] = cyc;
                row[12 * P.ld_traj] = (ET)1;
            }
            const bool nan_act = (colv != colv) || (cyc != cyc);                // main_3dof.py:192
            // -------- RLS.update(obs, actions, next_obs) (main_3dof.py:145, model.py:27-59) --------
            if (skip) {
                skip = 0;
            } else {
#pragma unroll
                for (int i = 0; i < 7; ++i) Y[i] = s[i] - Y[i];                 // model.py:43
                rls_update<ET, RD>(K, Wm, cov, X, Y);
            }
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
                SubVec<decltype(stl), NT> scr{stl, NHL + 3};
                agent_update<NT, NHL, UL, TAU1L>(lon, Hlon, P.lon.final_activation, stl[NHL + 1], stl[NHL + 2], n0, n1,
                                                 (NT)0, dr, F, G, stl, stl[NHL], scr);
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
                SubVec<decltype(stc), NT> scr{stc, NHC + 3};
                agent_update<NT, NHC, UC, false>(col, Hcol, P.col.final_activation, stc[NHC + 1], stc[NHC + 2], n0, n1,
                                                 (NT)0, dr, F, G, stc, stc[NHC], scr);
            }
            if (rec) {
                row[10 * P.ld_traj] = (ET)r_col;
                row[11 * P.ld_traj] = (ET)r_lon;
            }
            miscd[0] += r_col;
            miscd[1] += r_lon;
            if (k >= P.win_lo && k < P.win_hi) miscd[2] += r_lon;
            // -------- next step / termination (main_3dof.py:189-197) --------
            refth = nrefth;
            step = k + 1;
            const int nf = (done ? H3_FLAG_DONE : 0) | (nan_act ? H3_FLAG_NAN : 0);
            if (nf) {
                flags = nf;
                term_step = k;
            }
        }
    }

    // ---- exit store ----
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    if (valid) {
#pragma unroll
        for (int i = 0; i < 7; ++i) P.state[i * ld + e] = s[i];
        P.ref[ld + e] = refth;
        P.corr_u[e] = misce[0];
        P.pid_corr[e] = misce[1];
#pragma unroll
        for (int i = 0; i < 7; ++i) P.rls_x[i * ld + e] = prev[i];
        P.rls_u[e] = prev[7];
        P.rls_u[ld + e] = prev[8];
        P.rls_skip[e] = skip;
        P.step[e] = step;
        P.flags[e] = flags;
        P.term_step[e] = term_step;
        if (P.heat != nullptr) P.heat[e] = 0;
        acc0 = miscd[0];
        acc1 = miscd[1];
        acc2 = miscd[2];
        P.t[e] = miscd[3];
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
        for (int i = 0; i < 63; ++i) P.rls_W[i * ld + e] = Wm.v[i];
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
        const bool cnt = valid;
        v[0] = cnt ? 1.0 : 0.0;
        v[1] = (cnt && flags == 0) ? 1.0 : 0.0;
        v[2] = (cnt && (flags & H3_FLAG_DONE)) ? 1.0 : 0.0;
        v[3] = (cnt && (flags & H3_FLAG_NAN)) ? 1.0 : 0.0;
        v[4] = acc0;
        v[5] = acc1;
        v[6] = acc2;
        v[7] = cnt ? (double)step : 0.0;
        __syncthreads();
        block_reduce_stats<WARPS>(v, sh_red);
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
            block_reduce_stats<WARPS>(a, sh_red);
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < H3_N_STATS; ++i) P.stats[i] = a[i];
                *P.ticket = 0u;     // ready for the next launch
            }
        }
    }
}

}  // namespace h3
