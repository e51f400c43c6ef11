// "Wide" form of the fused closed-loop kernel for launches in which both agents learn (main_3dof.py:127-161 with
// update_col), fp64, the main_3dof.py configuration (6 + 10 hidden units, tau_lon = 1, symmetric covariance, shared
// sinusoid reference table, lockstep).
//
// The one-thread-per-episode kernel (h3_closed_loop.cuh) keeps 8 warps per SM resident in these launches -- the
// episode state fills the shared memory -- and is bound by the latency of its dependent FP64 chains (2 warps per
// scheduler, FP64 pipe 38 %).  Here a CTA of 16 warps owns 32 episodes (lanes = episodes, state in a
// `[row][32]` shared-memory tile) and the WARPS split the tick's work: warp u owns hidden unit u of the two
// networks (cyclic 0-5, collective 6-15: forward pass, critic / target passes, weight update) and one item of the
// RLS update (weight column u for u < 7, covariance row u - 7 for u >= 7); warp 0 also runs the serial part
// (final activations, model step, reference, regressor).  Same shared-memory footprint per episode, 4x the warps
// per SM to cover the FP64 latency; five block barriers per tick.
//
// Arithmetic is the reference's, summed in a fixed order (unit 0, 1, ...), so results do not depend on the launch
// geometry; they agree with the narrow kernel to rounding (tests/test_gpu_parity.py).
//
// MEASURED RESULT (B200, 262,144 episodes, ticks 200-600): 23.7 ms per 100 ticks against 10.1-10.9 ms of the narrow
// kernel.  27 % of the stall samples are the barrier behind warp 0's serial section, 18 % the cross-warp sums; the
// split also costs ~70 % more instructions (every warp re-derives d, the sums, F and G).  Even with the serial
// section overlapped the instruction count bounds it at about the narrow kernel's speed, so this path is OFF unless
// H3_WIDE=1 is set; it documents why the one-thread-per-episode form stays the product path for the 3-DOF tick.
#pragma once
#include <cstdlib>

#include "h3_closed_loop.cuh"

namespace h3 {

void set_error(const char* fmt, ...);

namespace wide {

constexpr int NHL = 6, NHC = 10, NU = NHL + NHC, WARPS = NU, BLOCK = 32 * WARPS;
// shared-memory rows, 32 doubles each
enum {
    R_S = 0, R_REF = 7, R_W = 8, R_COV = 71, R_PREV = 116, R_LON = 125, R_COL = 167, R_ALIVE = 277, R_PERSIST = 278,
    R_ACT = 278, R_F = 280, R_IN = 282 /* n0, n1 lon; n0, n1 col */, R_DR = 286, R_RLSON = 288, R_YK = 289,
    R_X = 305, R_Y = 314, R_CX = 321, R_RED = 305 /* aliases X, Y, CX: free once the RLS update is done */,
    R_ROWS = 305 + 6 * NU
};
constexpr size_t SMEM = (size_t)R_ROWS * 32 * sizeof(double);

__host__ __device__ constexpr int cidx(int i, int j) { return i * 9 - (i * (i - 1)) / 2 + (j - i); }   // i <= j

H3_DEV double tanh1(double x) {
    double a[1] = {x}, y[1];
    Math<double>::tanh_vec<1>(a, y);
    return y[0];
}

__global__ void __launch_bounds__(BLOCK, 2) closed_loop_wide_kernel(const __grid_constant__ LoopParams<double, double> P) {
    extern __shared__ double T[];                                   // [R_ROWS][32]
    const int lane = threadIdx.x & 31, u = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + lane, ld = P.ld;
    const bool valid = e < P.n;
    const int64_t ec = valid ? e : P.n - 1;
    const EnvK<double>& K = P.K;
    double* t = T + lane;
#define ROW(r) t[(r) * 32]
    const bool is_lon = u < NHL;
    const int k = is_lon ? u : u - NHL;                             // hidden unit inside its network
    const H3AgentHyper& ag = is_lon ? P.lon : P.col;
    const AgentH<double>& H = is_lon ? P.lon_h : P.col_h;
    const int ac = ag.ac_state, tr = ag.tracked_state;
    // this warp's weight rows
    const int nb = is_lon ? R_LON : R_COL;
    const int a1r = nb + 2 * k, a2r = nb + (is_lon ? 2 * NHL : 2 * NHC) + k;
    const int c1r = nb + (is_lon ? 3 * NHL : 3 * NHC) + 2 * k;
    const int c2r = nb + (is_lon ? 5 * NHL : 5 * NHC) + k, c2s = is_lon ? NHL : NHC;     // c2(j, k) at c2r + j * c2s
    const int t1r = is_lon ? c1r : nb + 7 * NHC + 2 * k;           // cyclic agent: tau = 1, target == critic
    const int t2r = is_lon ? c2r : nb + 9 * NHC + k;

    // ---- entry load: rows spread over the warps, every access coalesced ----
    for (int r = u; r < R_PERSIST; r += WARPS) {
        double v = 0.0;
        if (r < R_REF) v = P.state[(int64_t)r * ld + ec];
        else if (r == R_REF) v = P.ref[ld + ec];
        else if (r < R_COV) v = P.rls_W[(int64_t)(r - R_W) * ld + ec];
        else if (r < R_PREV) {
            const int l = r - R_COV;
            int i = 0;
            while (cidx(i + 1, i + 1) <= l && i < 8) ++i;
            v = P.rls_cov[(int64_t)(i * 9 + i + (l - cidx(i, i))) * ld + ec];
        } else if (r < R_LON) {
            const int i = r - R_PREV;
            v = i < 7 ? P.rls_x[(int64_t)i * ld + ec] : P.rls_u[(int64_t)(i - 7) * ld + ec];
        } else if (r < R_COL) v = P.lon_net[(int64_t)(r - R_LON) * ld + ec];    // actor + critic rows (7 nh)
        else if (r < R_ALIVE) v = P.col_net[(int64_t)(r - R_COL) * ld + ec];
        else v = (valid && P.flags[ec] == 0) ? 1.0 : 0.0;
        ROW(r) = v;
    }
    // warp 0: the per-episode scalars of the serial part
    double tt = 0, acc0 = 0, acc1 = 0, acc2 = 0, trim0 = 0, trim1 = 0;
    int skip = 0, step = 0, flags = 0, term = -1;
    if (u == 0) {
        tt = P.t[ec]; acc0 = P.acc[ec]; acc1 = P.acc[ld + ec]; acc2 = P.acc[2 * ld + ec];
        trim0 = P.trim[ec]; trim1 = P.trim[ld + ec];
        skip = P.rls_skip[ec]; step = P.step[ec]; flags = valid ? P.flags[ec] : 1; term = P.term_step[ec];
    }
    __syncthreads();

    for (int jj = 0; jj < P.n_ticks; ++jj) {
        const int kk = P.step0 + jj;
        // ---- A: hidden unit u of the actor forward pass (agents.py:59-63) ----
        const double refth = ROW(R_REF);
        const double s0 = ROW(R_S + ac), s1 = (tr == 1) ? ROW(R_S + 1) : ROW(R_S + tr) - refth;
        const double ha = tanh1(ROW(a1r) * s0 + ROW(a1r + 1) * s1);
        ROW(R_YK + u) = ROW(a2r) * ha;
        __syncthreads();                                                                    // 1
        // ---- B (warp 0): final activations, actions, model step, reference, regressor, rewards ----
        if (u == 0) {
            double yl = 0, yc = 0;
#pragma unroll
            for (int q = 0; q < NHL; ++q) yl += ROW(R_YK + q);
#pragma unroll
            for (int q = 0; q < NHC; ++q) yc += ROW(R_YK + NHL + q);
            double fin[2] = {yl, yc}, f[2];
            Math<double>::tanh_vec<2>(fin, f);
            double cyc = trim1 + P.lon_h.scale * f[0];                                      // main_3dof.py:132-137
            double colv = trim0 + P.col_h.scale * f[1];
            const bool exc_on = P.excitation != nullptr && kk < P.exc_len;
            if (exc_on) { cyc += P.excitation[2 * kk + 1]; }                                // :139-140
            double s[7], pre[7], X[9], sth, cth, scy, ccy;
#pragma unroll
            for (int i = 0; i < 7; ++i) s[i] = ROW(R_S + i);
            Math<double>::sincos2_(s[4], cyc, &sth, &cth, &scy, &ccy);
            if (exc_on) colv += P.excitation[2 * kk];
#pragma unroll
            for (int i = 0; i < 7; ++i) { pre[i] = s[i]; X[i] = s[i] - ROW(R_PREV + i); }   // model.py:41
            X[7] = colv - ROW(R_PREV + 7);
            X[8] = cyc - ROW(R_PREV + 8);
            bool done;
            env_step(K, s, colv, cyc, sth, cth, scy, ccy, done);                            // main_3dof.py:141
            const double nref = P.ref_table[jj];                                            // :142
            const bool alive = flags == 0;
            const bool nan_act = (colv != colv) || (cyc != cyc);
            if (alive) {
                tt = tt + P.dt;
#pragma unroll
                for (int i = 0; i < 7; ++i) { ROW(R_PREV + i) = pre[i]; ROW(R_S + i) = s[i]; }   // model.py:45
                ROW(R_PREV + 7) = colv; ROW(R_PREV + 8) = cyc;
                ROW(R_REF) = nref;
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) ROW(R_X + i) = X[i];
#pragma unroll
            for (int i = 0; i < 7; ++i) ROW(R_Y + i) = s[i] - pre[i];                       // model.py:43
            ROW(R_F) = f[0]; ROW(R_F + 1) = f[1];
            ROW(R_RLSON) = (alive && !skip) ? 1.0 : 0.0;
            if (alive) skip = 0;
            // agents' inputs that depend on the new state (agents.py:123-128, :142-150)
            const int lac = P.lon.ac_state, ltr = P.lon.tracked_state, cac = P.col.ac_state, ctr = P.col.tracked_state;
            const double el = (ltr == 1) ? s[1] : pick7(s, ltr) - refth, ecl = (ctr == 1) ? s[1] : pick7(s, ctr) - refth;
            ROW(R_IN) = pick7(s, lac);
            ROW(R_IN + 1) = (ltr == 1) ? s[1] : pick7(s, ltr) - nref;
            ROW(R_IN + 2) = pick7(s, cac);
            ROW(R_IN + 3) = (ctr == 1) ? s[1] : pick7(s, ctr) - nref;
            ROW(R_DR) = -2 * el * P.lon_wr;
            ROW(R_DR + 1) = -2 * ecl * P.col_wr;
            if (alive) {
                const double r_lon = -(el * el) * P.lon_wr, r_col = -(ecl * ecl) * P.col_wr;
                acc0 += r_col; acc1 += r_lon;
                if (kk >= P.win_lo && kk < P.win_hi) acc2 += r_lon;
                step = kk + 1;
                const int nf = (done ? H3_FLAG_DONE : 0) | (nan_act ? H3_FLAG_NAN : 0);
                if (nf) { flags = nf; term = kk; }
            }
        }
        __syncthreads();                                                                    // 2
        const bool alive = ROW(R_ALIVE) != 0.0;          // alive at the start of this tick: its updates are applied
        const bool rls_on = ROW(R_RLSON) != 0.0;
        // ---- C: cov.X, row u (model.py:53) ----
        if (u < 9) {
            double a0 = 0, a1 = 0;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const double c = (j < u) ? ROW(R_COV + cidx(j, u)) : ROW(R_COV + cidx(u, j));
                if (j & 1) a1 += c * ROW(R_X + j); else a0 += c * ROW(R_X + j);
            }
            ROW(R_CX + u) = a0 + a1;
        }
        __syncthreads();                                                                    // 3
        // ---- D: weight column u (u < 7) or covariance row u - 7 (model.py:49-59) ----
        {
            double d0 = K.rls_gamma, d1 = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { if (i & 1) d1 += ROW(R_CX + i) * ROW(R_X + i); else d0 += ROW(R_CX + i) * ROW(R_X + i); }
            const double inv_d = 1.0 / (d0 + d1);
            if (u < 7) {
                double w[9], ea = ROW(R_Y + u), eb = 0;
#pragma unroll
                for (int i = 0; i < 9; ++i) w[i] = ROW(R_W + i * 7 + u);
#pragma unroll
                for (int i = 0; i < 9; ++i) { if (i & 1) eb += ROW(R_X + i) * w[i]; else ea -= ROW(R_X + i) * w[i]; }
                const double err = ea - eb;
                if (rls_on) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) ROW(R_W + i * 7 + u) = w[i] + (ROW(R_CX + i) * inv_d) * err;
                }
            } else {
                const int r = u - 7;
                const double gr = ROW(R_CX + r) * inv_d;
                if (rls_on) {
#pragma unroll
                    for (int j = 0; j < 9; ++j)
                        if (j >= r) ROW(R_COV + cidx(r, j)) = (ROW(R_COV + cidx(r, j)) - gr * ROW(R_CX + j)) * K.rls_inv_gamma;
                }
            }
        }
        __syncthreads();                                                                    // 4
        // ---- E: unit u of critic(s), target critic(s') and da/ds (agents.py:169-175) ----
        const int ao = is_lon ? 0 : 1;
        const double n0 = ROW(R_IN + 2 * ao), n1 = ROW(R_IN + 2 * ao + 1), dr = ROW(R_DR + ao), f = ROW(R_F + ao);
        double F[2][2], G[2];
        {
            const int idx[2] = {ac, tr};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int j = 0; j < 2; ++j) F[i][j] = ROW(R_W + idx[j] * 7 + idx[i]);       // agents.py:137-140
                G[i] = ROW(R_W + (7 + ag.channel) * 7 + idx[i]);
            }
        }
        const double delta = H.scale * (1.0 - f * f);
        double pre2[2] = {ROW(c1r) * s0 + ROW(c1r + 1) * s1, ROW(t1r) * n0 + ROW(t1r + 1) * n1}, h2[2];
        Math<double>::tanh_vec<2>(pre2, h2);
        const double hc = h2[0], ht = h2[1];
        const double pk = delta * ROW(a2r) * (1.0 - ha * ha);
        {
            double* red = t + (R_RED + 6 * u) * 32;
            red[0] = ROW(c2r) * hc; red[32] = ROW(c2r + c2s) * hc;
            red[64] = ROW(t2r) * ht; red[96] = ROW(t2r + c2s) * ht;
            red[128] = ROW(a1r) * pk; red[160] = ROW(a1r + 1) * pk;
        }
        __syncthreads();                                                                    // 5
        // ---- F: sums over the network's units in unit order, then the SGD step of unit u (agents.py:178-196) ----
        {
            double sum[6] = {0, 0, 0, 0, 0, 0};
            const int u0 = is_lon ? 0 : NHL, nu = is_lon ? NHL : NHC;
            for (int q = 0; q < nu; ++q) {
                const double* red = t + (R_RED + 6 * (u0 + q)) * 32;
#pragma unroll
                for (int v = 0; v < 6; ++v) sum[v] += red[v * 32];
            }
            const double tg0 = H.gamma * sum[2], tg1 = dr + H.gamma * sum[3];
            double e_c[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) e_c[j] = sum[j] - (tg0 * (F[0][j] + G[0] * sum[4 + j]) + tg1 * (F[1][j] + G[1] * sum[4 + j]));
            const double ka = H.lr_a * (tg0 * G[0] + tg1 * G[1]);
            const double c20 = ROW(c2r), c21 = ROW(c2r + c2s);
            const double back = (c20 * e_c[0] + c21 * e_c[1]) * (1.0 - hc * hc);
            if (alive) {
                const double a10 = ROW(a1r), a11 = ROW(a1r + 1), a2k = ROW(a2r), c10 = ROW(c1r), c11 = ROW(c1r + 1);
                ROW(a1r) = a10 + ka * (pk * s0);
                ROW(a1r + 1) = a11 + ka * (pk * s1);
                ROW(a2r) = a2k + ka * (delta * ha);
                const double m10 = c10 - H.lr_c * (back * s0), m11 = c11 - H.lr_c * (back * s1);
                const double m20 = c20 - H.lr_c * (e_c[0] * hc), m21 = c21 - H.lr_c * (e_c[1] * hc);
                if (!is_lon) {                                      // soft target update (agents.py:195-196)
                    const double omt = 1.0 - H.tau;
                    ROW(t1r) = H.tau * m10 + omt * ROW(t1r);
                    ROW(t1r + 1) = H.tau * m11 + omt * ROW(t1r + 1);
                    ROW(t2r) = H.tau * m20 + omt * ROW(t2r);
                    ROW(t2r + c2s) = H.tau * m21 + omt * ROW(t2r + c2s);
                }
                ROW(c1r) = m10; ROW(c1r + 1) = m11; ROW(c2r) = m20; ROW(c2r + c2s) = m21;
            }
        }
        if (u == 0 && flags != 0) ROW(R_ALIVE) = 0.0;    // terminated on this tick: frozen from the next one on
    }
    __syncthreads();

    // ---- exit store ----
    if (valid) {
        for (int r = u; r < R_ALIVE; r += WARPS) {
            const double v = ROW(r);
            if (r < R_REF) P.state[(int64_t)r * ld + e] = v;
            else if (r == R_REF) P.ref[ld + e] = v;
            else if (r < R_COV) P.rls_W[(int64_t)(r - R_W) * ld + e] = v;
            else if (r < R_PREV) {
                const int l = r - R_COV;
                int i = 0;
                while (cidx(i + 1, i + 1) <= l && i < 8) ++i;
                const int j = i + (l - cidx(i, i));
                P.rls_cov[(int64_t)(i * 9 + j) * ld + e] = v;
                if (j != i) P.rls_cov[(int64_t)(j * 9 + i) * ld + e] = v;
            } else if (r < R_LON) {
                const int i = r - R_PREV;
                if (i < 7) P.rls_x[(int64_t)i * ld + e] = v; else P.rls_u[(int64_t)(i - 7) * ld + e] = v;
            } else if (r < R_COL) {
                const int i = r - R_LON;
                P.lon_net[(int64_t)i * ld + e] = v;
                if (i >= 3 * NHL) P.lon_net[(int64_t)(i + 4 * NHL) * ld + e] = v;    // target critic == critic (tau = 1)
            } else P.col_net[(int64_t)(r - R_COL) * ld + e] = v;
        }
        if (u == 0) {
            P.t[e] = tt; P.acc[e] = acc0; P.acc[ld + e] = acc1; P.acc[2 * ld + e] = acc2;
            P.rls_skip[e] = skip; P.step[e] = step; P.flags[e] = flags; P.term_step[e] = term;
        }
    }
#undef ROW
}

// statistics of the batch after a launch (same vector as the narrow kernel's epilogue), deterministic two-stage sum
__global__ void __launch_bounds__(256) wide_stats_partial(const LoopParams<double, double> P, int nblocks) {
    __shared__ double sh[8 * H3_N_STATS];
    double v[H3_N_STATS];
#pragma unroll
    for (int i = 0; i < H3_N_STATS; ++i) v[i] = 0;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < P.n; e += (int64_t)nblocks * 256) {
        const int fl = P.flags[e];
        v[0] += 1; v[1] += fl == 0; v[2] += (fl & H3_FLAG_DONE) != 0; v[3] += (fl & H3_FLAG_NAN) != 0;
        v[4] += P.acc[e]; v[5] += P.acc[P.ld + e]; v[6] += P.acc[2 * P.ld + e]; v[7] += (double)P.step[e];
    }
    block_reduce_stats<8>(v, sh);
    if (threadIdx.x == 0)
        for (int i = 0; i < H3_N_STATS; ++i) P.partials[(int64_t)blockIdx.x * H3_N_STATS + i] = v[i];
}
__global__ void __launch_bounds__(256) wide_stats_final(const LoopParams<double, double> P, int nblocks) {
    __shared__ double sh[8 * H3_N_STATS];
    double v[H3_N_STATS];
#pragma unroll
    for (int i = 0; i < H3_N_STATS; ++i) v[i] = 0;
    for (int b = threadIdx.x; b < nblocks; b += 256)
        for (int i = 0; i < H3_N_STATS; ++i) v[i] += P.partials[(int64_t)b * H3_N_STATS + i];
    block_reduce_stats<8>(v, sh);
    if (threadIdx.x == 0)
        for (int i = 0; i < H3_N_STATS; ++i) P.stats[i] = v[i];
}

}  // namespace wide

// Is this launch one the wide kernel implements?  (everything else runs the narrow kernel)
inline bool wide_eligible(const H3Config& cfg, const H3Batch& b, int step0, int n_ticks, const double* ref_table,
                          const H3Traj* traj) {
    // opt-in (H3_WIDE=1): measured 2.2x SLOWER than the narrow kernel (header comment), kept as a tested experiment
    const char* sw = getenv("H3_WIDE");
    if (sw == nullptr || sw[0] != '1') return false;
    const bool per_ep = b.col_hyper != nullptr || b.lon_hyper != nullptr;
    return cfg.env_dtype == H3_F64 && cfg.net_dtype == H3_F64 && cfg.lon.n_hidden == wide::NHL && cfg.col.n_hidden == wide::NHC &&
           cfg.rls_symmetric != 0 && !per_ep && cfg.lon.tau == 1.0 && cfg.lon.final_activation == H3_ACT_TANH &&
           cfg.col.final_activation == H3_ACT_TANH && cfg.task == H3_TASK_SINUSOID && cfg.ref_mode == H3_REF_TABLE &&
           ref_table != nullptr && step0 >= cfg.switch_step && cfg.update_lon != 0 && (traj == nullptr || traj->out == nullptr) &&
           n_ticks > 0 && b.n > 0;
}

inline int launch_wide(const LoopParams<double, double>& P, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(wide::closed_loop_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide::SMEM);
    if (err != cudaSuccess) { set_error("cudaFuncSetAttribute(wide): %s", cudaGetErrorString(err)); return H3_E_CUDA; }
    const int64_t grid = (P.n + 31) / 32;
    wide::closed_loop_wide_kernel<<<(unsigned)grid, wide::BLOCK, wide::SMEM, stream>>>(P);
    if (P.stats != nullptr) {
        int nb = (int)((P.n + 255) / 256);
        if (nb > 512) nb = 512;
        wide::wide_stats_partial<<<nb, 256, 0, stream>>>(P, nb);
        wide::wide_stats_final<<<1, 256, 0, stream>>>(P, nb);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("closed_loop_wide_kernel launch: %s", cudaGetErrorString(err)); return H3_E_CUDA; }
    return H3_OK;
}

}  // namespace h3
