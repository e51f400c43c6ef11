// The fused closed-loop kernel for launches in which BOTH agents act and learn (ticks >= switch_step,
// main_3dof.py:127-137 / :156-159), one thread per episode, with the collective agent's networks parked in
// TENSOR MEMORY (h3_tmem.cuh).
//
// Why: in these launches an episode carries 2.3-2.7 KB of fp64 state.  Registers (256 KB per SM) and shared
// memory (227 KB) together hold 128 episodes per SM -- one warp per scheduler with a lane per episode (round 1 bought
// a second warp per scheduler with two lanes per episode, at the price of executing the model step, PID and
// covariance update twice: 145 warp-instructions per env-step against 97 here).  An sm_100a SM has a third on-chip store: 256 KB of tensor
// memory, 128 lanes x 512 columns x 32 bit, and its 32x32b access shape is exactly "N private columns per
// thread".  With the collective agent's 11 nh weights (+ the hidden activations kept from the forward pass) in
// TMEM, the RLS matrices in shared memory and the cyclic agent in registers, 8 warps x 32 episodes are resident
// per SM with ONE lane per episode: the same occupancy as the PID-collective launches, no duplicated work.
//
// Layout of a thread's TMEM range (words of NT; a double is two columns): hidden unit k owns 12 words at 12 k
//   [t1(k,0) t1(k,1) t2(0,k) t2(1,k) | c1(k,0) c1(k,1) c2(0,k) c2(1,k) | a1(k,0) a1(k,1) a2(k) h(k)]
// followed by [f, s0, s1, -] of the forward pass.  Every pass of the agent's update (col_update_tm) reads one
// contiguous piece per hidden unit.
//
// tcgen05.ld / st are warp-collective, so the warp stays converged: a terminated lane is not skipped, it keeps
// executing on sanitised inputs with its bookkeeping frozen.  Its state is written to HBM at the tick it
// terminates (`store_state`, one code site inside the resumable tick loop), exactly the state the reference
// holds when it breaks out of its loop (main_3dof.py:192-197).  Lockstep launches only (shared tick index, so
// the phase test is warp-uniform); heterogeneous batches use h3_closed_loop.cuh.
#pragma once
#include "h3_closed_loop.cuh"
#include "h3_tmem.cuh"

namespace h3 {

template <int NH>
struct ColTm {
    static constexpr int UNIT = 12;
    static constexpr int TAIL = UNIT * NH;
    static constexpr int WORDS = TAIL + 4;
    // target critic | critic | actor + a hidden activation (the actor's from the forward pass to pass A of the update,
    // the critic's from there to pass C)
    enum { T10 = 0, T11, T20, T21, C10, C11, C20, C21, A10, A11, A2, HA };
    // row of the `[11 nh][ld]` HBM block (include/heli3dof.h) that holds slot `w` of unit k (w != HA)
    H3_DEV static int row(int k, int w) {
        switch (w) {
            case A10: return 2 * k;
            case A11: return 2 * k + 1;
            case A2: return 2 * NH + k;
            case C10: return 3 * NH + 2 * k;
            case C11: return 3 * NH + 2 * k + 1;
            case C20: return 5 * NH + k;
            case C21: return 6 * NH + k;
            case T10: return 7 * NH + 2 * k;
            case T11: return 7 * NH + 2 * k + 1;
            case T20: return 9 * NH + k;
            default: return 10 * NH + k;
        }
    }
};

// DHPActor.forward of the collective agent (agents.py:59-63): weights from TMEM, h_a and (f, s0, s1) back to it.
//   UF: hidden units per group (their tanh chains are interleaved)
template <typename NT, int NH, int UF>
H3_DEV NT col_forward_tm(uint32_t tm, NT s0, NT s1, NT scale, int act) {
    typedef Math<NT> M;
    typedef ColTm<NH> L;
    static_assert(NH % UF == 0, "group width must divide the hidden width");
    NT y = (NT)0;
#pragma unroll 1
    for (int k0 = 0; k0 < NH; k0 += UF) {
        TmemBuf<NT, 4> b[UF];               // a1(k,0) a1(k,1) a2(k) (h)
#pragma unroll
        for (int u = 0; u < UF; ++u) b[u].issue(tm, L::UNIT * (k0 + u) + L::A10);
        tmem_wait_ld();
        NT pre[UF], hk[UF];
#pragma unroll
        for (int u = 0; u < UF; ++u) pre[u] = b[u][0] * s0 + b[u][1] * s1;
        M::template tanh_vec<UF>(pre, hk);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            y += b[u][2] * hk[u];
            const NT w[1] = {hk[u]};
            tmem_store<NT, 1>(tm, L::UNIT * (k0 + u) + L::HA, w);
        }
    }
    const NT f = (act == H3_ACT_TANH) ? M::tanh_(y) : M::sigmoid_(y);
    __syncwarp();
    const NT tl[4] = {f, s0, s1, (NT)0};
    tmem_store<NT, 4>(tm, L::TAIL, tl);
    return scale * f;
}

// get_reward + get_transition_matrices + update_networks (agents.py:123-196) of the collective agent, closed form
// as h3_tick.cuh: agent_update (same formulas, sums in the same order), arranged as three passes over the hidden
// units so that every loop runs with a RUNTIME unit index (TMEM addresses are plain register arithmetic) and the
// code is not replicated per chunk:
//   T : target critic on s' (:171)            -> lambda_t2, hence the target and the actor's step size
//   A : critic on s (:170), da/ds (:174), and the actor's SGD step (:184-186), which needs nothing from the other
//       units; the critic's hidden activation then takes the slot of the actor's (no longer needed)
//   C : critic SGD step (:187-188) and soft target update (:195-196)
//   UP: hidden units per group of passes T and A (their tanh chains are interleaved); V2: units per group of pass C.
template <typename NT, int NH, int UP, int V2>
H3_DEV void col_update_tm(uint32_t tm, const AgentH<NT>& H, int act, NT n0, NT n1, NT dr0, NT dr, const NT (&F)[2][2],
                          const NT (&G)[2]) {
    typedef Math<NT> M;
    typedef ColTm<NH> L;
    static_assert(NH % UP == 0 && NH % V2 == 0, "chunk widths must divide the hidden width");
    tmem_wait_st();                                     // the forward pass stored h_a and the tail
    NT tl[4];
    tmem_load<NT, 4>(tm, L::TAIL, tl);
    const NT f = tl[0], s0 = tl[1], s1 = tl[2];
    const NT delta = (act == H3_ACT_TANH) ? H.scale * ((NT)1 - f * f) : H.scale * f * ((NT)1 - f);
    // ---- pass T ----
    NT lam2[2] = {(NT)0, (NT)0};
#pragma unroll 1
    for (int k0 = 0; k0 < NH; k0 += UP) {
        TmemBuf<NT, 4> b[UP];               // t1(k,0) t1(k,1) t2(0,k) t2(1,k)
#pragma unroll
        for (int u = 0; u < UP; ++u) b[u].issue(tm, L::UNIT * (k0 + u) + L::T10);
        tmem_wait_ld();
        NT pre[UP], ht[UP];
#pragma unroll
        for (int u = 0; u < UP; ++u) pre[u] = b[u][0] * n0 + b[u][1] * n1;
        M::template tanh_vec<UP>(pre, ht);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            lam2[0] += b[u][2] * ht[u];
            lam2[1] += b[u][3] * ht[u];
        }
    }
    // target = dr_ds + gamma * lambda_t2 (:178); actor step size lr_a * (target . G) (:184)
    const NT tg0 = dr0 + H.gamma * lam2[0];
    const NT tg1 = dr + H.gamma * lam2[1];
    const NT ka = H.lr_a * (tg0 * G[0] + tg1 * G[1]);
    const NT ka_s0 = ka * s0, ka_s1 = ka * s1, ka_d = ka * delta;       // as in agent_update (h3_tick.cuh)
    // ---- pass A ----
    NT lam1[2] = {(NT)0, (NT)0}, da_ds[2] = {(NT)0, (NT)0};
#pragma unroll 1
    for (int k0 = 0; k0 < NH; k0 += UP) {
        NT pre[UP], hc[UP];
        {
            TmemBuf<NT, 2> c1[UP];          // c1(k,0) c1(k,1)
#pragma unroll
            for (int u = 0; u < UP; ++u) c1[u].issue(tm, L::UNIT * (k0 + u) + L::C10);
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < UP; ++u) pre[u] = c1[u][0] * s0 + c1[u][1] * s1;
        }
        M::template tanh_vec<UP>(pre, hc);
        __syncwarp();
        TmemBuf<NT, 6> b[UP];               // c2(0,k) c2(1,k) a1(k,0) a1(k,1) a2(k) h_a(k)
#pragma unroll
        for (int u = 0; u < UP; ++u) b[u].issue(tm, L::UNIT * (k0 + u) + L::C20);
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            lam1[0] += b[u][0] * hc[u];
            lam1[1] += b[u][1] * hc[u];
            const NT ha = b[u][5], a2k = b[u][4];
            const NT pk = delta * a2k * ((NT)1 - ha * ha);
            da_ds[0] += b[u][2] * pk;
            da_ds[1] += b[u][3] * pk;
            const NT o[4] = {b[u][2] + ka_s0 * pk, b[u][3] + ka_s1 * pk, a2k + ka_d * ha, hc[u]};
            tmem_store<NT, 4>(tm, L::UNIT * (k0 + u) + L::A10, o);
        }
    }
    // error_critic = lambda_t1 - target (F + G da_ds)  (:179)
    NT e_c[2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
        e_c[j] = lam1[j] - (tg0 * (F[0][j] + G[0] * da_ds[j]) + tg1 * (F[1][j] + G[1] * da_ds[j]));
    const NT one_m_tau = (NT)1 - H.tau;
    const NT lc_s0 = H.lr_c * s0, lc_s1 = H.lr_c * s1, lc_e0 = H.lr_c * e_c[0], lc_e1 = H.lr_c * e_c[1];
    tmem_wait_st();                                     // pass A stored the critic's hidden activations
    // ---- pass C ----
#pragma unroll 1
    for (int k0 = 0; k0 < NH; k0 += V2) {
        TmemBuf<NT, 8> b[V2];               // t1(k,:) t2(:,k) c1(k,:) c2(:,k)
        TmemBuf<NT, 1> h[V2];               // critic hidden activation (in the h_a slot)
#pragma unroll
        for (int u = 0; u < V2; ++u) {
            b[u].issue(tm, L::UNIT * (k0 + u) + L::T10);
            h[u].issue(tm, L::UNIT * (k0 + u) + L::HA);
        }
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < V2; ++u) {
            const NT hc = h[u][0], c20 = b[u][6], c21 = b[u][7];
            const NT back = (c20 * e_c[0] + c21 * e_c[1]) * ((NT)1 - hc * hc);
            const NT n10 = b[u][4] - lc_s0 * back;
            const NT n11 = b[u][5] - lc_s1 * back;
            const NT n20 = c20 - lc_e0 * hc;
            const NT n21 = c21 - lc_e1 * hc;
            const NT o[8] = {H.tau * n10 + one_m_tau * b[u][0], H.tau * n11 + one_m_tau * b[u][1],
                             H.tau * n20 + one_m_tau * b[u][2], H.tau * n21 + one_m_tau * b[u][3], n10, n11, n20, n21};
            tmem_store<NT, 8>(tm, L::UNIT * (k0 + u) + L::T10, o);
        }
    }
    tmem_wait_st();                                     // the next tick's forward pass loads these columns
}

// cyclic agent: actor + critic in registers, target critic (only carried when tau != 1) in the shared-memory tile
#ifndef H3_TM_TARGET_REGS
#define H3_TM_TARGET_REGS 1     // the cyclic agent's target critic (tau != 1) in registers: 32 episodes per warp at the price
                                // of ~430 B of spills (14 % faster than 25 episodes per warp with the target in the tile)
#endif
template <typename NT, int NH, int STRIDE, bool TARGET>
struct LonNetTm {
    LocalVec<NT, 7 * NH> v;
    typename std::conditional<H3_TM_TARGET_REGS, LocalVec<NT, 4 * NH>, SmemVec<NT, 4 * NH, STRIDE>>::type tv;
    H3_DEV NT& a1(int k, int i) { return v[2 * k + i]; }
    H3_DEV NT& a2(int k) { return v[2 * NH + k]; }
    H3_DEV NT& c1(int k, int i) { return v[3 * NH + 2 * k + i]; }
    H3_DEV NT& c2(int j, int k) { return v[5 * NH + j * NH + k]; }
    H3_DEV NT& t1(int k, int i) { return tv[2 * k + i]; }
    H3_DEV NT& t2(int j, int k) { return tv[2 * NH + j * NH + k]; }
};

//   COLT: the launch reaches the switch -- the collective agent lives in tensor memory.  false (launches that end before
//         the switch, per-episode-hyper variant): its networks stay in HBM and the columns hold the cyclic agent's TARGET
//         critic instead (tau != 1), which is then in registers only while agent_update runs.
template <typename ET, typename NT, int NHL, int NHC, bool HYP, bool TAU1L, int LANES, int WARPS_ = 4, bool COLT = true>
struct TmLayout {
    static constexpr int WARPS = WARPS_;           // 4 or 8: warps w and w + 4 share a lane quadrant, own column ranges
    static constexpr int STRIDE = LANES < 32 ? LANES + 1 : 32;      // + one dummy column for the idle lanes
    static constexpr size_t bytes_per_col = sizeof(ET) * (45 + 63) + ((TAU1L || H3_TM_TARGET_REGS) ? 0 : sizeof(NT) * 4 * NHL);
    static constexpr size_t tile_bytes = ((bytes_per_col * STRIDE + 15) / 16) * 16;
    static constexpr size_t head_bytes = sizeof(double) * WARPS * (3 * TICK_CHUNK + H3_N_STATS) + 16;
    static constexpr size_t smem_bytes = head_bytes + tile_bytes * WARPS;
    static constexpr bool TGT_TM = !COLT && !TAU1L;                 // cyclic agent's target critic in tensor memory
    static constexpr int tgt_col = (COLT ? ColTm<NHC>::WORDS : 0) * (int)(sizeof(NT) / 4);
    // after the networks: the three reward accumulators (doubles, + one pad word) -- 8 columns
    static constexpr int acc_col = tgt_col + (TGT_TM ? 4 * NHL : 0) * (int)(sizeof(NT) / 4);
    static constexpr int tm_words = acc_col + 8;
    static constexpr int tm_cols = tm_words <= 32 ? 32 : (tm_words <= 64 ? 64 : (tm_words <= 128 ? 128 : (tm_words <= 256 ? 256 : 512)));
    static_assert(tm_cols * (WARPS / 4) <= 512 && (WARPS == 4 || WARPS == 8), "collective agent does not fit a thread's tensor-memory columns");
};

//   HYP   : per-episode hyper-parameter rows (re-read from L2 every tick); TAU1L: tau_lon == 1 (target critic of the cyclic
//           agent not carried, see agent_update); LANES: episodes per warp (<= 32, as many as the tile allows);
//   RD    : look-ahead of the RLS shared-memory streams; UF / UP / V2: hidden units per group of the collective
//           agent's forward pass / critic passes / weight update
template <typename ET, typename NT, int NHL, int NHC, bool HYP, bool TAU1L, int LANES, int RD, int UF, int UP, int V2, int WARPS_,
          bool COLT = true>
__global__ void __launch_bounds__(WARPS_ * 32, 8 / WARPS_) closed_loop_tm_kernel(const LoopParams<ET, NT> P) {
    typedef TmLayout<ET, NT, NHL, NHC, HYP, TAU1L, LANES, WARPS_, COLT> LY;
    constexpr bool TGT_TM = LY::TGT_TM;
    static_assert(!TGT_TM || H3_TM_TARGET_REGS, "the target critic moves between tensor memory and registers");
    constexpr int WARPS = LY::WARPS, STRIDE = LY::STRIDE, BLOCK = WARPS * 32;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* sh_ref = reinterpret_cast<double*>(smem_raw) + warp * 3 * TICK_CHUNK;       // [TICK_CHUNK], per warp
    double* sh_exc = sh_ref + TICK_CHUNK;                                               // [TICK_CHUNK][2]
    double* sh_red = reinterpret_cast<double*>(smem_raw) + WARPS * 3 * TICK_CHUNK;      // [WARPS][H3_N_STATS]
    uint32_t* sh_slot = reinterpret_cast<uint32_t*>(sh_red + WARPS * H3_N_STATS);

    const uint32_t tbase = tmem_alloc(sh_slot, LY::tm_cols * (WARPS / 4));
    const uint32_t tm = tmem_addr(tbase, LY::tm_cols * (warp >> 2));

    const int64_t e = ((int64_t)blockIdx.x * WARPS + warp) * LANES + lane;
    const bool valid = (lane < LANES) && (e < P.n);
    const int64_t ld = P.ld;
    const int64_t ec = valid ? e : 0;               // idle lanes read episode 0 (in bounds) and never store
    const int lc = lane < LANES ? lane : LANES;     // ... and share the tile's dummy column
    const EnvK<ET>& K = P.K;
    const uint32_t tm_acc = tm + (uint32_t)LY::acc_col;     // sum r_col, sum r_lon, sum r_lon inside the score window
    const uint32_t tm_tgt = tm + (uint32_t)LY::tgt_col;     // cyclic agent's target critic (TGT_TM)
    Math<NT>::prepare();                                    // tanh table -> (static) shared memory

    // ---- carve the warp's tile ----
    unsigned char* tp = smem_raw + LY::head_bytes + LY::tile_bytes * warp;
    typedef SmemVec<ET, 45, STRIDE> CovVec;
    CovSym<ET, CovVec> cov;
    WMatD<ET, SmemVec<ET, 63, STRIDE>> Wm;
    LonNetTm<NT, NHL, STRIDE, !TAU1L> lon;
    {
        ET* pe = reinterpret_cast<ET*>(tp);
        cov.v.p = pe + lc; pe += 45 * STRIDE;
        Wm.v.p = pe + lc; pe += 63 * STRIDE;
        NT* pn = reinterpret_cast<NT*>(pe);
        if constexpr (!H3_TM_TARGET_REGS) lon.tv.p = pn + lc;      // only dereferenced when !TAU1L
    }

    // ---- entry load (coalesced SoA rows) ----
    if constexpr (COLT) {   // collective agent -> TMEM, one hidden unit at a time
        typedef ColTm<NHC> L;
#pragma unroll 1
        for (int k = 0; k < NHC; ++k) {
            NT w[L::UNIT];
#pragma unroll
            for (int i = 0; i < L::UNIT; ++i)
                w[i] = (i == L::HA) ? (NT)0 : P.col_net[(int64_t)L::row(k, i) * ld + ec];
            tmem_store<NT, L::UNIT>(tm, L::UNIT * k, w);
        }
    }
    ET s[7], refth, prev[9];
    double t_ep;
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
    // cold per-episode scalars stay out of the register file: the accumulators live in tensor memory (one load / store
    // pair per tick), the trimmed controls (read-only) are re-read from L2 every tick, the PID / velocity-task
    // integrators are read and written in place by the ticks that use them (PID collective: ticks < switch_step only)
    {
        const double a0[4] = {P.acc[ec], P.acc[ld + ec], P.acc[2 * ld + ec], 0.0};
        tmem_store<double, 4>(tm_acc, 0, a0);
    }
    t_ep = P.t[ec];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = i; j < 9; ++j) cov.v[CovSym<ET, CovVec>::idx(i, j)] = P.rls_cov[(i * 9 + j) * ld + ec];
#pragma unroll
    for (int i = 0; i < 63; ++i) Wm.v[i] = P.rls_W[i * ld + ec];
    lon.v.load(P.lon_net + ec, ld);                                     // actor + critic rows [0, 7 nh)
    if constexpr (!TAU1L) lon.tv.load(P.lon_net + (int64_t)7 * NHL * ld + ec, ld);
    if constexpr (TGT_TM) tmem_store<NT, 4 * NHL>(tm_tgt, 0, lon.tv.v);
    tmem_wait_st();
    __syncwarp();

    const bool in_kernel_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_KERNEL);
    const bool table_ref = (P.task == H3_TASK_SINUSOID) && (P.ref_mode == H3_REF_TABLE);
    NT stl[2 * NHL + 3];                    // hl[NHL], fl, (sl0, sl1: update only), scratch[NHL] of the cyclic agent

    // ---- resumable tick loop: leaves whenever a lane terminated, so that its state is stored at once ----
    int j = 0;
    bool just_died = false, finished = false;
    do {
        bool flush = false;
        for (; j < P.n_ticks && !flush; ++j) {
            const int jj = j & (TICK_CHUNK - 1);
            if (jj == 0) {                  // stage the next TICK_CHUNK ticks of reference / excitation (per warp)
                __syncwarp();
                if (lane < TICK_CHUNK && j + lane < P.n_ticks) {
                    if (table_ref) sh_ref[lane] = P.ref_table[j + lane];
                    const int kk = P.step0 + j + lane;
                    const bool on = (P.excitation != nullptr) && (kk < P.exc_len);
                    sh_exc[2 * lane] = on ? P.excitation[2 * kk] : 0.0;
                    sh_exc[2 * lane + 1] = on ? P.excitation[2 * kk + 1] : 0.0;
                }
                __syncwarp();
            }
            const bool dead = flags != 0;
            if (__all_sync(FULL, dead)) continue;       // every episode of this warp has terminated
            const int k = P.step0 + j;
            const bool phase2 = COLT && (k >= P.switch_step);                       // main_3dof.py:127-137
            if (dead) {                     // frozen episode: run on harmless values, commit nothing
#pragma unroll
                for (int i = 0; i < 7; ++i) s[i] = (ET)0;
                refth = (ET)0;
            }
            // trimmed controls (read-only rows, L2): issued here, consumed after the forward passes
            const ET trim1 = (ET)__ldcg(P.trim + ld + ec);
            ET trim0 = (ET)0;
            if (phase2) trim0 = (ET)__ldcg(P.trim + ec);
            AgentH<NT> Hcol = P.col_h, Hlon = P.lon_h;
            ET wr_col = P.col_wr, wr_lon = P.lon_wr;
            if constexpr (HYP) {            // read-only rows, L2-resident: re-read per tick instead of pinning 24 registers
                double wc, wl;
                Hcol = load_hyper<NT, true>(P.col, P.col_hyper, ld, ec, wc);
                Hlon = load_hyper<NT, true>(P.lon, P.lon_hyper, ld, ec, wl);
                wr_col = (ET)wc;
                wr_lon = (ET)wl;
            }
            // -------- actions (main_3dof.py:132-137); forward activations are kept for the update --------
            ET sth, cth, cyc, colv;
            {
                NT fl;
                const NT sl0 = (NT)pick7(s, P.lon.ac_state);
                const NT sl1 = (NT)(P.lon.tracked_state == 1 ? s[1] : pick7(s, P.lon.tracked_state) - refth);
                const NT a_lon = actor_forward<NT, NHL, NHL>(lon, sl0, sl1, Hlon.scale, P.lon.final_activation,
                                                                    stl, fl);
                cyc = trim1 + (ET)a_lon;
                stl[NHL] = fl;
            }
            if (phase2) {
                const NT sc0 = (NT)pick7(s, P.col.ac_state);
                const NT sc1 = (NT)(P.col.tracked_state == 1 ? s[1] : pick7(s, P.col.tracked_state) - refth);
                __syncwarp();
                const NT a_col = col_forward_tm<NT, NHC, UF>(tm, sc0, sc1, Hcol.scale, P.col.final_activation);
                colv = trim0 + (ET)a_col;
            }
            cyc += (ET)sh_exc[2 * jj + 1];                                          // main_3dof.py:139-140
            if (dead) cyc = (ET)0;
            ET scy, ccy;
            Math<ET>::sincos2_(s[4], cyc, &sth, &cth, &scy, &ccy);
            if (!phase2) {
                ET pc = P.pid_corr[ec];
                colv = pid_collective(K, s, sth, cth, pc);                          // PID.py:16-24
                if (!dead) P.pid_corr[ec] = pc;
            }
            colv += (ET)sh_exc[2 * jj];
            if (dead) colv = (ET)0;
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
            ET cu = (ET)0;
            if (P.task == H3_TASK_VELOCITY) {
                cu = P.corr_u[ec];
                (void)velocity_ref(K, s[2], cu);                                    // heli_models.py:155
            }
            bool done;
            env_step(K, s, colv, cyc, sth, cth, scy, ccy, done);
            const double t_new = t_ep + P.dt;                                        // heli_models.py:206
            // -------- next_ref = env.get_ref() (main_3dof.py:142) --------
            ET nrefth;
            if (table_ref) {
                nrefth = (ET)sh_ref[jj];
            } else if (in_kernel_ref) {
                double amp = P.ref_amp, period = P.ref_period;
                if (P.ref_par != nullptr) {
                    amp = __ldcg(P.ref_par + ec);
                    period = __ldcg(P.ref_par + ld + ec);
                }
                nrefth = (ET)sinusoid_ref(t_new + P.dt, amp, period);
            } else if (P.task == H3_TASK_VELOCITY) {
                nrefth = velocity_ref(K, s[2], cu);
            } else {
                nrefth = stop_over_point_ref(K, s[0], s[2], (ET)__ldcg(P.theta_trim + ec));
            }
            // -------- log row (main_3dof.py:164-175) --------
            const bool rec = P.traj != nullptr && e < P.n_traj && !dead;
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
                if (!dead) skip = 0;
            } else {
#pragma unroll
                for (int i = 0; i < 7; ++i) Y[i] = s[i] - Y[i];                     // model.py:43
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
                // the augmented state of (obs, ref) again: prev[] holds the pre-step observation (model.py:45)
                const NT sl0 = (NT)pick7(prev, P.lon.ac_state);
                const NT sl1 = (NT)(P.lon.tracked_state == 1 ? prev[1] : pick7(prev, P.lon.tracked_state) - refth);
                SubVec<decltype(stl), NT> scr{stl, NHL + 3};
                if constexpr (TGT_TM) {         // target critic: out of tensor memory for the update only
                    __syncwarp();
                    tmem_wait_st();
                    tmem_load<NT, 4 * NHL>(tm_tgt, 0, lon.tv.v);
                }
                agent_update<NT, NHL, NHL, TAU1L>(lon, Hlon, P.lon.final_activation, sl0, sl1, n0, n1, (NT)0, dr,
                                                         F, G, stl, stl[NHL], scr);
                if constexpr (TGT_TM) {
                    __syncwarp();
                    tmem_store<NT, 4 * NHL>(tm_tgt, 0, lon.tv.v);
                }
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
                __syncwarp();
                col_update_tm<NT, NHC, UP, V2>(tm, Hcol, P.col.final_activation, n0, n1, (NT)0, dr, F, G);
            }
            // -------- bookkeeping and termination (main_3dof.py:189-197): frozen for terminated episodes --------
            {   // accumulators: warp-collective round trip through tensor memory, the sums committed by live lanes only
                __syncwarp();
                tmem_wait_st();
                double a[4];
                tmem_load<double, 4>(tm_acc, 0, a);
                if (!dead) {
                    a[0] += r_col;
                    a[1] += r_lon;
                    if (k >= P.win_lo && k < P.win_hi) a[2] += r_lon;
                }
                __syncwarp();
                tmem_store<double, 4>(tm_acc, 0, a);
            }
            if (!dead) {
                if (rec) {
                    row[10 * P.ld_traj] = (ET)r_col;
                    row[11 * P.ld_traj] = (ET)r_lon;
                }
                t_ep = t_new;
                if (P.task == H3_TASK_VELOCITY) P.corr_u[ec] = cu;
                refth = nrefth;
                step = k + 1;
                const int nf = (done ? H3_FLAG_DONE : 0) | (nan_act ? H3_FLAG_NAN : 0);
                if (nf) {
                    flags = nf;
                    term_step = k;
                    just_died = true;
                }
            }
            flush = __any_sync(FULL, just_died);
        }
        finished = j >= P.n_ticks;

        // ---- store (coalesced SoA rows): episodes that terminated in the tick just executed, and at the end
        //      of the launch every live episode ----
        const bool st = valid && (just_died || (finished && flags == 0));
        __syncwarp();
        double accv[4];
        tmem_wait_st();
        tmem_load<double, 4>(tm_acc, 0, accv);
        if constexpr (TGT_TM) tmem_load<NT, 4 * NHL>(tm_tgt, 0, lon.tv.v);
        if constexpr (COLT) {
            typedef ColTm<NHC> L;
#pragma unroll 1
            for (int k = 0; k < NHC; ++k) {
                NT w[L::UNIT];
                tmem_load<NT, L::UNIT>(tm, L::UNIT * k, w);
                if (st) {
#pragma unroll
                    for (int i = 0; i < L::UNIT; ++i)
                        if (i != L::HA) P.col_net[(int64_t)L::row(k, i) * ld + e] = w[i];
                }
            }
        }
        if (st) {
#pragma unroll
            for (int i = 0; i < 7; ++i) P.state[i * ld + e] = s[i];
            P.ref[ld + e] = refth;
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
            P.acc[e] = accv[0];
            P.acc[ld + e] = accv[1];
            P.acc[2 * ld + e] = accv[2];
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int jc = i; jc < 9; ++jc) {
                    const ET c = cov.v[CovSym<ET, CovVec>::idx(i, jc)];
                    P.rls_cov[(i * 9 + jc) * ld + e] = c;
                    if (jc != i) P.rls_cov[(jc * 9 + i) * ld + e] = c;
                }
#pragma unroll
            for (int i = 0; i < 63; ++i) P.rls_W[i * ld + e] = Wm.v[i];
            lon.v.store(P.lon_net + e, ld);
            if constexpr (TAU1L) {          // target critic == critic (agents.py:196 with tau = 1)
#pragma unroll
                for (int i = 0; i < 4 * NHL; ++i) P.lon_net[(int64_t)(7 * NHL + i) * ld + e] = lon.v[3 * NHL + i];
            } else {
                lon.tv.store(P.lon_net + (int64_t)7 * NHL * ld + e, ld);
            }
        }
        just_died = false;
    } while (!finished);

    // ---- statistics epilogue: block partials -> last block sums them in a fixed order ----
    if (P.stats != nullptr) {
        double v[H3_N_STATS];
        v[0] = valid ? 1.0 : 0.0;
        v[1] = (valid && flags == 0) ? 1.0 : 0.0;
        v[2] = (valid && (flags & H3_FLAG_DONE)) ? 1.0 : 0.0;
        v[3] = (valid && (flags & H3_FLAG_NAN)) ? 1.0 : 0.0;
        // (a terminated episode's accumulators are frozen: every update above is guarded by !dead)
        double accv[4];
        __syncwarp();
        tmem_load<double, 4>(tm_acc, 0, accv);
        v[4] = valid ? accv[0] : 0.0;
        v[5] = valid ? accv[1] : 0.0;
        v[6] = valid ? accv[2] : 0.0;
        v[7] = valid ? (double)step : 0.0;
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
    tmem_free(tbase, LY::tm_cols * (WARPS / 4));
}

}  // namespace h3
