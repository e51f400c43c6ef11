// Any hidden width: the agents of the reference take arbitrary n_hidden_actor / n_hidden_critic (agents.py:74-78).
// The fused kernels are compiled for the widths of the 3-DOF drivers (6 / 10, and 8 / 8); every other width runs
// here -- the same tick, one thread per episode, with the hidden width a RUNTIME value, the network blocks, the RLS
// matrices and the accumulators left in global memory (SoA rows: a warp's access to a row is one coalesced
// transaction, the working set of a launch is L2-resident) and only the model state in registers.  It is the
// correctness path for unusual configurations, not a fast one: nothing is staged on chip, and the update recomputes
// the forward activations instead of keeping NH of them per thread.
#pragma once
#include "h3_closed_loop.cuh"

namespace h3 {

template <typename T>
struct GVec {               // strided column of a `[rows][ld]` global-memory block
    static constexpr bool kFence = false;
    T* p;
    int64_t ld;
    H3_DEV T& operator[](int i) { return p[(int64_t)i * ld]; }
};

template <typename NT>
struct RtNet {              // rows of the `[11*nh][ld]` network block (include/heli3dof.h), nh at run time
    NT* p;
    int64_t ld;
    int nh;
    H3_DEV NT& at(int r) { return p[(int64_t)r * ld]; }
    H3_DEV NT& a1(int k, int i) { return at(2 * k + i); }
    H3_DEV NT& a2(int k) { return at(2 * nh + k); }
    H3_DEV NT& c1(int k, int i) { return at(3 * nh + 2 * k + i); }
    H3_DEV NT& c2(int j, int k) { return at(5 * nh + j * nh + k); }
    H3_DEV NT& t1(int k, int i) { return at(7 * nh + 2 * k + i); }
    H3_DEV NT& t2(int j, int k) { return at(9 * nh + j * nh + k); }
};

template <typename NT>
H3_DEV NT tanh_rt(NT x) {
    return Math<NT>::tanh_(x);
}

// DHPActor.forward (agents.py:59-63)
template <typename NT>
H3_DEV NT actor_forward_rt(RtNet<NT>& net, NT s0, NT s1, NT scale, int act, NT& f) {
    typedef Math<NT> M;
    NT y = (NT)0;
    for (int k = 0; k < net.nh; ++k) y += net.a2(k) * tanh_rt(net.a1(k, 0) * s0 + net.a1(k, 1) * s1);
    f = (act == H3_ACT_TANH) ? M::tanh_(y) : M::sigmoid_(y);
    return scale * f;
}

// update_networks (agents.py:152-196), closed form with the operation order of h3_tick.cuh: agent_update; hidden
// activations are recomputed where that function keeps them (the weights they depend on are still the old ones: a
// unit's rows are written only after they were read).
template <typename NT>
H3_DEV void agent_update_rt(RtNet<NT>& net, const AgentH<NT>& H, int act, NT s0, NT s1, NT n0, NT n1, NT dr0, NT dr,
                            const NT (&F)[2][2], const NT (&G)[2]) {
    NT f;
    (void)actor_forward_rt(net, s0, s1, H.scale, act, f);
    const NT delta = (act == H3_ACT_TANH) ? H.scale * ((NT)1 - f * f) : H.scale * f * ((NT)1 - f);
    NT lam1[2] = {(NT)0, (NT)0}, lam2[2] = {(NT)0, (NT)0}, da_ds[2] = {(NT)0, (NT)0};
    for (int k = 0; k < net.nh; ++k) {
        const NT hc = tanh_rt(net.c1(k, 0) * s0 + net.c1(k, 1) * s1);
        const NT ht = tanh_rt(net.t1(k, 0) * n0 + net.t1(k, 1) * n1);
        lam1[0] += net.c2(0, k) * hc;
        lam1[1] += net.c2(1, k) * hc;
        lam2[0] += net.t2(0, k) * ht;
        lam2[1] += net.t2(1, k) * ht;
        const NT ha = tanh_rt(net.a1(k, 0) * s0 + net.a1(k, 1) * s1);
        const NT pk = delta * net.a2(k) * ((NT)1 - ha * ha);
        da_ds[0] += net.a1(k, 0) * pk;
        da_ds[1] += net.a1(k, 1) * pk;
    }
    const NT tg0 = dr0 + H.gamma * lam2[0];
    const NT tg1 = dr + H.gamma * lam2[1];
    NT e_c[2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
        e_c[j] = lam1[j] - (tg0 * (F[0][j] + G[0] * da_ds[j]) + tg1 * (F[1][j] + G[1] * da_ds[j]));
    const NT ka = H.lr_a * (tg0 * G[0] + tg1 * G[1]);
    const NT one_m_tau = (NT)1 - H.tau;
    const NT ka_s0 = ka * s0, ka_s1 = ka * s1, ka_d = ka * delta;       // as in agent_update (h3_tick.cuh)
    const NT lc_s0 = H.lr_c * s0, lc_s1 = H.lr_c * s1, lc_e0 = H.lr_c * e_c[0], lc_e1 = H.lr_c * e_c[1];
    for (int k = 0; k < net.nh; ++k) {
        const NT a10 = net.a1(k, 0), a11 = net.a1(k, 1), a2k = net.a2(k);
        const NT c10 = net.c1(k, 0), c11 = net.c1(k, 1), c20 = net.c2(0, k), c21 = net.c2(1, k);
        const NT ha = tanh_rt(a10 * s0 + a11 * s1);
        const NT hc = tanh_rt(c10 * s0 + c11 * s1);
        const NT pk = delta * a2k * ((NT)1 - ha * ha);
        const NT back = (c20 * e_c[0] + c21 * e_c[1]) * ((NT)1 - hc * hc);
        net.a1(k, 0) = a10 + ka_s0 * pk;
        net.a1(k, 1) = a11 + ka_s1 * pk;
        net.a2(k) = a2k + ka_d * ha;
        const NT n10 = c10 - lc_s0 * back;
        const NT n11 = c11 - lc_s1 * back;
        const NT n20 = c20 - lc_e0 * hc;
        const NT n21 = c21 - lc_e1 * hc;
        net.c1(k, 0) = n10;
        net.c1(k, 1) = n11;
        net.c2(0, k) = n20;
        net.c2(1, k) = n21;
        net.t1(k, 0) = H.tau * n10 + one_m_tau * net.t1(k, 0);
        net.t1(k, 1) = H.tau * n11 + one_m_tau * net.t1(k, 1);
        net.t2(0, k) = H.tau * n20 + one_m_tau * net.t2(0, k);
        net.t2(1, k) = H.tau * n21 + one_m_tau * net.t2(1, k);
    }
}

constexpr int RT_BLOCK = 128;

template <typename ET, typename NT>
__global__ void __launch_bounds__(RT_BLOCK) closed_loop_rt_kernel(const LoopParams<ET, NT> P, int nhl, int nhc) {
    const int tid = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * RT_BLOCK + tid;
    const bool valid = e < P.n;
    const int64_t ld = P.ld;
    const int64_t ec = valid ? e : 0;
    const EnvK<ET>& K = P.K;
    __shared__ double sh_red[(RT_BLOCK / 32) * H3_N_STATS];
    Math<NT>::prepare();

    CovFull<ET, GVec<ET>> cov;              // model.py:53-59 literally, on the full matrix
    cov.v.p = P.rls_cov + ec;
    cov.v.ld = ld;
    WMat<ET, GVec<ET>> Wm;
    Wm.v.p = P.rls_W + ec;
    Wm.v.ld = ld;
    RtNet<NT> col{P.col_net + ec, ld, nhc}, lon{P.lon_net + ec, ld, nhl};

    ET s[7], refth, prev[9];
#pragma unroll
    for (int i = 0; i < 7; ++i) s[i] = P.state[i * ld + ec];
    refth = P.ref[ld + ec];
    int skip = P.rls_skip[ec];
    int step = P.step[ec];
    int flags = valid ? P.flags[ec] : H3_FLAG_DONE;
    int term_step = P.term_step[ec];
#pragma unroll
    for (int i = 0; i < 7; ++i) prev[i] = P.rls_x[i * ld + ec];
    prev[7] = P.rls_u[ec];
    prev[8] = P.rls_u[ld + ec];
    double acc0 = P.acc[ec], acc1 = P.acc[ld + ec], acc2 = P.acc[2 * ld + ec], t_ep = P.t[ec];
    ET corr_u = P.corr_u[ec], pid_corr = P.pid_corr[ec];
    const ET trim0 = (ET)P.trim[ec], trim1 = (ET)P.trim[ld + ec];
    const bool lockstep = P.step0 >= 0;
    const bool in_kernel_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_KERNEL);
    const bool table_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_TABLE);

    for (int j = 0; j < P.n_ticks && flags == 0; ++j) {
        const int k = lockstep ? (P.step0 + j) : step;
        const bool phase2 = k >= P.switch_step;                                 // main_3dof.py:127-137
        double wc, wl;
        const AgentH<NT> Hcol = load_hyper<NT, true>(P.col, P.col_hyper, ld, ec, wc);
        const AgentH<NT> Hlon = load_hyper<NT, true>(P.lon, P.lon_hyper, ld, ec, wl);
        const ET wr_col = (ET)wc, wr_lon = (ET)wl;
        // -------- actions (main_3dof.py:132-137) --------
        const NT sl0 = (NT)pick7(s, P.lon.ac_state);
        const NT sl1 = (NT)(P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - refth);
        NT fdummy;
        ET cyc = trim1 + (ET)actor_forward_rt(lon, sl0, sl1, Hlon.scale, P.lon.final_activation, fdummy);
        const NT sc0 = (NT)pick7(s, P.col.ac_state);
        const NT sc1 = (NT)(P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - refth);
        ET colv = (ET)0;
        if (phase2) colv = trim0 + (ET)actor_forward_rt(col, sc0, sc1, Hcol.scale, P.col.final_activation, fdummy);
        ET exc0 = (ET)0, exc1 = (ET)0;                                          // main_3dof.py:139-140
        if (P.excitation != nullptr && k < P.exc_len) {
            exc0 = (ET)P.excitation[2 * k];
            exc1 = (ET)P.excitation[2 * k + 1];
        }
        cyc += exc1;
        ET sth, cth, scy, ccy;
        Math<ET>::sincos2_(s[4], cyc, &sth, &cth, &scy, &ccy);
        if (!phase2) colv = pid_collective(K, s, sth, cth, pid_corr);          // PID.py:16-24
        colv += exc0;
        // -------- env.step (main_3dof.py:141) --------
        ET X[9], Y[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            Y[i] = s[i];
            X[i] = s[i] - prev[i];              // model.py:41
            prev[i] = s[i];                     // model.py:45
        }
        X[7] = colv - prev[7];
        X[8] = cyc - prev[8];
        prev[7] = colv;
        prev[8] = cyc;
        if (P.task == H3_TASK_VELOCITY) (void)velocity_ref(K, s[2], corr_u);    // heli_models.py:155
        bool done;
        env_step(K, s, colv, cyc, sth, cth, scy, ccy, done);
        t_ep = t_ep + P.dt;                                                     // heli_models.py:206
        // -------- next_ref = env.get_ref() (main_3dof.py:142) --------
        ET nrefth;
        if (table_ref) {
            nrefth = (ET)P.ref_table[j];
        } else if (in_kernel_ref) {
            double amp = P.ref_amp, period = P.ref_period;
            if (P.ref_par != nullptr) {
                amp = P.ref_par[ec];
                period = P.ref_par[ld + ec];
            }
            nrefth = (ET)sinusoid_ref(t_ep + P.dt, amp, period);
        } else if (P.task == H3_TASK_VELOCITY) {
            nrefth = velocity_ref(K, s[2], corr_u);
        } else {
            nrefth = stop_over_point_ref(K, s[0], s[2], (ET)P.theta_trim[ec]);
        }
        const bool rec = P.traj != nullptr && e < P.n_traj;
        ET* row = rec ? P.traj + ((int64_t)j * H3_TRAJ_ROWS) * P.ld_traj + e : nullptr;
        if (rec) {
#pragma unroll
            for (int i = 0; i < 7; ++i) row[i * P.ld_traj] = Y[i];
            row[7 * P.ld_traj] = refth;
            row[8 * P.ld_traj] = colv;
            row[9 * P.ld_traj] = cyc;
            row[12 * P.ld_traj] = (ET)1;
        }
        const bool nan_act = (colv != colv) || (cyc != cyc);                    // main_3dof.py:192
        // -------- RLS.update (main_3dof.py:145, model.py:27-59) --------
        if (skip) {
            skip = 0;
        } else {
#pragma unroll
            for (int i = 0; i < 7; ++i) Y[i] = s[i] - Y[i];                     // model.py:43
            rls_update<ET, 1>(K, Wm, cov, X, Y);
        }
        // -------- agents (main_3dof.py:148-159) --------
        double r_lon = 0.0, r_col = 0.0;
        if (P.update_lon) {
            const ET e_tr = P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - refth;
            r_lon = (double)(-(e_tr * e_tr) * wr_lon);
            const NT dr = (NT)((ET)-2 * e_tr * wr_lon);
            NT F[2][2], G[2];
            transition_matrices<ET, NT>(Wm, P.lon, F, G);
            const NT n0 = (NT)pick7(s, P.lon.ac_state);
            const NT n1 = (NT)(P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - nrefth);
            agent_update_rt(lon, Hlon, P.lon.final_activation, sl0, sl1, n0, n1, (NT)0, dr, F, G);
        }
        if (phase2) {
            const ET e_tr = P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - refth;
            r_col = (double)(-(e_tr * e_tr) * wr_col);
            const NT dr = (NT)((ET)-2 * e_tr * wr_col);
            NT F[2][2], G[2];
            transition_matrices<ET, NT>(Wm, P.col, F, G);
            const NT n0 = (NT)pick7(s, P.col.ac_state);
            const NT n1 = (NT)(P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - nrefth);
            agent_update_rt(col, Hcol, P.col.final_activation, sc0, sc1, n0, n1, (NT)0, dr, F, G);
        }
        if (rec) {
            row[10 * P.ld_traj] = (ET)r_col;
            row[11 * P.ld_traj] = (ET)r_lon;
        }
        acc0 += r_col;
        acc1 += r_lon;
        if (k >= P.win_lo && k < P.win_hi) acc2 += r_lon;
        // -------- next step / termination (main_3dof.py:189-197) --------
        refth = nrefth;
        step = k + 1;
        const int nf = (done ? H3_FLAG_DONE : 0) | (nan_act ? H3_FLAG_NAN : 0);
        if (nf) {
            flags = nf;
            term_step = k;
        }
    }

    if (valid) {
#pragma unroll
        for (int i = 0; i < 7; ++i) P.state[i * ld + e] = s[i];
        P.ref[ld + e] = refth;
        P.corr_u[e] = corr_u;
        P.pid_corr[e] = pid_corr;
#pragma unroll
        for (int i = 0; i < 7; ++i) P.rls_x[i * ld + e] = prev[i];
        P.rls_u[e] = prev[7];
        P.rls_u[ld + e] = prev[8];
        P.rls_skip[e] = skip;
        P.step[e] = step;
        P.flags[e] = flags;
        P.term_step[e] = term_step;
        if (P.heat != nullptr) P.heat[e] = 0;
        P.t[e] = t_ep;
        P.acc[e] = acc0;
        P.acc[ld + e] = acc1;
        P.acc[2 * ld + e] = acc2;
    }

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
        block_reduce_stats<RT_BLOCK / 32>(v, sh_red);
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
            for (int64_t b = tid; b < (int64_t)gridDim.x; b += RT_BLOCK)
#pragma unroll
                for (int i = 0; i < H3_N_STATS; ++i) a[i] += __ldcg(&P.partials[b * H3_N_STATS + i]);
            __syncthreads();
            block_reduce_stats<RT_BLOCK / 32>(a, sh_red);
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < H3_N_STATS; ++i) P.stats[i] = a[i];
                *P.ticket = 0u;
            }
        }
    }
}

}  // namespace h3
