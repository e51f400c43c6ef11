// C ABI of libheli3dof.so (include/heli3dof.h) and the un-fused per-object kernels that back the
// step-by-step methods of the Python classes (Helicopter3DOF.step, CollectivePID3DOF.__call__,
// RecursiveLeastSquares.update, DHPAgent.get_action / update_networks).  The fused path is
// h3_closed_loop (h3_closed_loop.cuh).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "h3_launch.cuh"
#include "h3_tmem.cuh"

namespace h3 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return H3_E_CUDA;
    }
    return H3_OK;
}

constexpr int SMALL_BLOCK = 128;
static inline unsigned grid_for(int64_t n) { return (unsigned)((n + SMALL_BLOCK - 1) / SMALL_BLOCK); }

// ---------------------------------------------------------------------------------------------
// Helicopter3DOF._trim (heli_models.py:293-328)
// ---------------------------------------------------------------------------------------------
// one episode: trimmed [collective, cyclic] and the trimmed state [x z u w theta q lambda_i]
__device__ __forceinline__ void trim_one(const H3EnvParams& E, double vt, double& collective, double& cyclic,
                                         double (&st)[7]) {
    const double weight = E.mass * E.g;                                 // :302
    const double drag = 0.5 * E.rho * vt * vt * E.cds;                  // :303
    const double thrust = ::sqrt(drag * drag + weight * weight);        // :304
    const double c_t = thrust / (E.rho * E.v_tip * E.v_tip * E.area);   // :305
    const double theta_f = ::atan2(-drag, weight);                      // :308
    double sf, cf;
    ::sincos(theta_f, &sf, &cf);
    const double mu = vt / E.v_tip;                                     // :314
    // quartic 4 l^4 + b3 l^3 + b2 l^2 - c_t^2 = 0 (:315-318).  For l > 0 the polynomial is increasing and
    // convex (b3, b2 >= 0), so it has exactly one positive root and Newton from the hover value
    // sqrt(c_t / 2) -- where the polynomial is >= 0 -- descends onto it monotonically.
    const double b3 = 8.0 * mu * ::sin(drag / weight);
    const double b2 = 4.0 * mu * mu;
    const double c0 = c_t * c_t;
    double l = ::sqrt(c_t * 0.5);
    for (int it = 0; it < 64; ++it) {
        const double f = ((4.0 * l + b3) * l + b2) * l * l - c0;
        const double df = ((16.0 * l + 3.0 * b3) * l + 2.0 * b2) * l;
        const double ln = l - f / df;
        if (!(ln < l)) break;
        l = ln;
    }
    // 2x2 solve (:321-324): [[a, b], [c, d]] [cyclic, collective]^T = [r0, r1]^T
    const double a = 1.0 + 1.5 * mu * mu, b = -8.0 / 3.0 * mu, c = -mu, d = 2.0 / 3.0 + mu * mu;
    const double r0 = -2.0 * mu * mu * drag / weight - 2.0 * mu * l;
    const double r1 = 4.0 / E.blade_solidity * c_t / E.cl_alpha + mu * drag / weight + l;
    const double det = a * d - b * c;
    cyclic = (r0 * d - b * r1) / det;
    collective = (a * r1 - c * r0) / det;
    st[0] = 0.0;
    st[1] = 0.0;
    st[2] = vt * cf;
    st[3] = vt * sf;
    st[4] = theta_f;
    st[5] = 0.0;
    st[6] = l;
}

__global__ void __launch_bounds__(SMALL_BLOCK) trim_kernel(H3EnvParams E, const double* __restrict__ v, double* trim,
                                                           double* state, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    double collective, cyclic, st[7];
    trim_one(E, v[e], collective, cyclic, st);
    trim[e] = collective;
    trim[ld + e] = cyclic;
#pragma unroll
    for (int i = 0; i < 7; ++i) state[i * ld + e] = st[i];
}

// ---------------------------------------------------------------------------------------------
// h3_init_episodes: every per-episode field at loop entry of main_3dof.py, from (seed, global episode index)
// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so a draw depends on (key, counter) only.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double uniform53(uint32_t hi, uint32_t lo) {
    const uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return ((double)m + 0.5) * 1.1102230246251565e-16;          // 2^-53
}

// standard normals (2 j, 2 j + 1) of `stream` for global episode `ep` (Box-Muller)
__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t ep, uint32_t stream, uint32_t j, double& z0,
                                            double& z1) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)ep, (uint32_t)(ep >> 32), stream, j),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const double u1 = uniform53(r.x, r.y), u2 = uniform53(r.z, r.w);
    const double rad = ::sqrt(-2.0 * ::log(u1));
    double sn, cs;
    ::sincos(2.0 * 3.14159265358979323846 * u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

// rows [row0, row0 + cnt) of a network block <- sigma * N(0, 1)
template <typename NT>
__device__ __forceinline__ void init_matrix(NT* net, int64_t ld, int64_t e, int row0, int cnt, uint64_t seed, uint64_t ep,
                                            uint32_t stream, double sigma) {
    for (int j = 0; 2 * j < cnt; ++j) {
        double z0, z1;
        normal_pair(seed, ep, stream, (uint32_t)j, z0, z1);
        net[(int64_t)(row0 + 2 * j) * ld + e] = (NT)(sigma * z0);
        if (2 * j + 1 < cnt) net[(int64_t)(row0 + 2 * j + 1) * ld + e] = (NT)(sigma * z1);
    }
}

template <typename NT>
__device__ __forceinline__ void init_agent(NT* net, int64_t ld, int64_t e, int nh, uint64_t seed, uint64_t ep, int agent,
                                           double sig_a, double sig_c) {
    init_matrix(net, ld, e, 0, 2 * nh, seed, ep, 4 * agent + 0, sig_a);             // actor fc1   agents.py:42-45
    init_matrix(net, ld, e, 2 * nh, nh, seed, ep, 4 * agent + 1, sig_a);            // actor fc2
    init_matrix(net, ld, e, 3 * nh, 2 * nh, seed, ep, 4 * agent + 2, sig_c);        // critic fc1  agents.py:19-22
    init_matrix(net, ld, e, 5 * nh, 2 * nh, seed, ep, 4 * agent + 3, sig_c);        // critic fc2
    for (int r = 0; r < 4 * nh; ++r)                                                // target = copy  agents.py:107
        net[(int64_t)(7 * nh + r) * ld + e] = net[(int64_t)(3 * nh + r) * ld + e];
}

template <typename ET, typename NT>
__global__ void __launch_bounds__(SMALL_BLOCK)
    init_episodes_kernel(H3EnvParams E, EnvK<ET> K, H3Batch b, H3Init in, int task, double ref_amp, double ref_period,
                         int nh_col, int nh_lon) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= b.n) return;
    const int64_t ld = b.ld;
    const uint64_t ep = in.episode_id != nullptr ? (uint64_t)in.episode_id[e] : (uint64_t)(in.episode0 + e);
    // Helicopter3DOF.reset (heli_models.py:285-291)
    double collective, cyclic, st[7];
    trim_one(E, in.v_trim_ep != nullptr ? in.v_trim_ep[e] : in.v_trim, collective, cyclic, st);
    double* trim = const_cast<double*>(b.trim);
    trim[e] = collective;
    trim[ld + e] = cyclic;
    const_cast<double*>(b.theta_trim)[e] = st[4];
    ET* state = (ET*)b.state;
#pragma unroll
    for (int i = 0; i < 7; ++i) state[i * ld + e] = (ET)st[i];
    b.t[e] = 0.0;
    // ref = env.get_ref() before the loop (main_3dof.py:62; velocity task: integrates corr_u once)
    ET* ref = (ET*)b.ref;
    ET cu = (ET)0;
    if (task == H3_TASK_SINUSOID) {
        double amp = ref_amp, period = ref_period;
        if (b.ref_par != nullptr) {
            amp = b.ref_par[e];
            period = b.ref_par[ld + e];
        }
        ref[e] = (ET)-0.0;
        ref[ld + e] = (ET)sinusoid_ref(E.dt, amp, period);
    } else if (task == H3_TASK_VELOCITY) {
        ref[e] = (ET)0;
        ref[ld + e] = velocity_ref(K, (ET)st[2], cu);
    } else {
        ref[e] = (ET)0;
        ref[ld + e] = stop_over_point_ref(K, (ET)st[0], (ET)st[2], (ET)st[4]);
    }
    ((ET*)b.corr_u)[e] = cu;
    ((ET*)b.pid_corr)[e] = (ET)0;                                       // PID.py:8
    // RecursiveLeastSquares.reset (model.py:79-89)
    ET* W = (ET*)b.rls_W;
    for (int i = 0; i < 63; ++i) W[(int64_t)i * ld + e] = (ET)0;
    ET* cov = (ET*)b.rls_cov;
    for (int i = 0; i < 81; ++i) cov[(int64_t)i * ld + e] = (i % 10 == 0) ? (ET)in.rls_cov0 : (ET)0;
    for (int i = 0; i < 7; ++i) ((ET*)b.rls_x)[(int64_t)i * ld + e] = (ET)0;
    ((ET*)b.rls_u)[e] = (ET)0;
    ((ET*)b.rls_u)[ld + e] = (ET)0;
    b.rls_skip[e] = 1;
    // DHPAgent.__init__ (agents.py:104-107)
    double sg[4] = {in.sigma[0], in.sigma[1], in.sigma[2], in.sigma[3]};
    if (in.sigma_ep != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sg[i] = in.sigma_ep[(int64_t)i * ld + e];
    }
    init_agent((NT*)b.col_net, ld, e, nh_col, in.seed, ep, 0, sg[0], sg[1]);
    init_agent((NT*)b.lon_net, ld, e, nh_lon, in.seed, ep, 1, sg[2], sg[3]);
    // bookkeeping (main_3dof.py:122)
    b.step[e] = 0;
    b.flags[e] = 0;
    b.term_step[e] = -1;
    b.acc[e] = 0.0;
    b.acc[ld + e] = 0.0;
    b.acc[2 * ld + e] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// Helicopter3DOF.step, open loop, n_ticks per launch (BASELINE config 2)
// ---------------------------------------------------------------------------------------------
template <typename ET>
__global__ void __launch_bounds__(SMALL_BLOCK)
    env_step_kernel(EnvK<ET> K, double dt, ET* state, double* t, int32_t* step, int32_t* flags, int32_t* term_step,
                    const double* __restrict__ trim, const ET* __restrict__ actions,
                    const double* __restrict__ excitation, int exc_len, int n_ticks, ET* states_out, int64_t n_out,
                    int64_t ld_out, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    ET s[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) s[i] = state[i * ld + e];
    ET a0, a1;
    if (actions != nullptr) {
        a0 = actions[e];
        a1 = actions[ld + e];
    } else {
        a0 = (ET)trim[e];
        a1 = (ET)trim[ld + e];
    }
    int k = step[e], f = flags[e], ts = term_step[e];
    double tt = t[e];
    for (int j = 0; j < n_ticks && f == 0; ++j) {
        ET col = a0, cyc = a1;
        if (excitation != nullptr && k < exc_len) {
            col += (ET)excitation[2 * k];
            cyc += (ET)excitation[2 * k + 1];
        }
        ET sth, cth, scy, ccy;
        Math<ET>::sincos2_(s[4], cyc, &sth, &cth, &scy, &ccy);
        bool done;
        env_step(K, s, col, cyc, sth, cth, scy, ccy, done);
        tt += dt;
        if (states_out != nullptr && e < n_out) {
#pragma unroll
            for (int i = 0; i < 7; ++i) states_out[((int64_t)j * 7 + i) * ld_out + e] = s[i];
        }
        const bool nan_act = (col != col) || (cyc != cyc);
        const int nf = (done ? H3_FLAG_DONE : 0) | (nan_act ? H3_FLAG_NAN : 0);
        if (nf) {
            f = nf;
            ts = k;
        }
        k += 1;
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) state[i * ld + e] = s[i];
    t[e] = tt;
    step[e] = k;
    flags[e] = f;
    term_step[e] = ts;
}

// ---------------------------------------------------------------------------------------------
// Helicopter3DOF.get_ref
// ---------------------------------------------------------------------------------------------
template <typename ET>
__global__ void __launch_bounds__(SMALL_BLOCK)
    get_ref_kernel(EnvK<ET> K, int task, double dt, double amp0, double period0, const ET* __restrict__ state,
                   const double* __restrict__ t, ET* ref, ET* corr_u, const double* __restrict__ theta_trim,
                   const double* __restrict__ ref_par, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    if (task == H3_TASK_SINUSOID) {
        double amp = amp0, period = period0;
        if (ref_par != nullptr) {
            amp = ref_par[e];
            period = ref_par[ld + e];
        }
        ref[e] = (ET)-0.0;
        ref[ld + e] = (ET)sinusoid_ref(t[e] + dt, amp, period);
    } else if (task == H3_TASK_VELOCITY) {
        ET cu = corr_u[e];
        ref[e] = (ET)0;
        ref[ld + e] = velocity_ref(K, state[2 * ld + e], cu);
        corr_u[e] = cu;
    } else {
        ref[e] = (ET)0;
        ref[ld + e] = stop_over_point_ref(K, state[e], state[2 * ld + e], (ET)theta_trim[e]);
    }
}

// ---------------------------------------------------------------------------------------------
// CollectivePID3DOF.__call__
// ---------------------------------------------------------------------------------------------
template <typename ET>
__global__ void __launch_bounds__(SMALL_BLOCK)
    pid_kernel(EnvK<ET> K, const ET* __restrict__ state, ET* pid_corr, ET* out, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    ET s[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) s[i] = state[i * ld + e];
    ET sth, cth;
    Math<ET>::sincos_(s[4], &sth, &cth);
    ET corr = pid_corr[e];
    out[e] = pid_collective(K, s, sth, cth, corr);
    pid_corr[e] = corr;
}

// ---------------------------------------------------------------------------------------------
// RecursiveLeastSquares.update / predict -- state kept in global memory (un-fused path)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct GlobalVec {
    static constexpr bool kFence = false;
    T* p;
    int64_t ld;
    H3_DEV T& operator[](int i) { return p[(int64_t)i * ld]; }
};

template <typename ET>
__global__ void __launch_bounds__(SMALL_BLOCK)
    rls_update_kernel(EnvK<ET> K, int sym, ET* W, ET* cov, ET* xp, ET* up, int32_t* skip, const ET* __restrict__ state,
                      const ET* __restrict__ action, const ET* __restrict__ next_state, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    ET st[7], ns[7], ac[2];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        st[i] = state[i * ld + e];
        ns[i] = next_state[i * ld + e];
    }
    ac[0] = action[e];
    ac[1] = action[ld + e];
    if (skip[e]) {                                  // model.py:34-39
        skip[e] = 0;
    } else {
        ET X[9], Y[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            X[i] = st[i] - xp[i * ld + e];
            Y[i] = ns[i] - st[i];
        }
        X[7] = ac[0] - up[e];
        X[8] = ac[1] - up[ld + e];
        // literal model.py:49-59 on the HBM-resident matrices (rolled loops: this path is bandwidth-bound)
        (void)sym;
        ET err[7], covX[9], Xcov[9];
#pragma unroll
        for (int j = 0; j < 7; ++j) err[j] = Y[j];
#pragma unroll 1
        for (int i = 0; i < 9; ++i) {
            const ET xi = X[i];
#pragma unroll
            for (int j = 0; j < 7; ++j) err[j] -= xi * W[(i * 7 + j) * ld + e];           // :49-50
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) Xcov[i] = (ET)0;
#pragma unroll 1
        for (int i = 0; i < 9; ++i) {
            ET a = (ET)0;
            const ET xi = X[i];
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const ET c = cov[(i * 9 + j) * ld + e];
                a += c * X[j];                                                          // :53  cov X
                Xcov[j] += xi * c;                                                      // :54  X^T cov
            }
            covX[i] = a;
        }
        ET d = K.rls_gamma;
#pragma unroll
        for (int j = 0; j < 9; ++j) d += Xcov[j] * X[j];                                // :55
        const ET inv_d = (ET)1 / d;
#pragma unroll 1
        for (int i = 0; i < 9; ++i) {
            const ET gi = covX[i] * inv_d;
#pragma unroll
            for (int j = 0; j < 7; ++j) W[(i * 7 + j) * ld + e] += gi * err[j];           // :58
#pragma unroll
            for (int j = 0; j < 9; ++j)
                cov[(i * 9 + j) * ld + e] = (cov[(i * 9 + j) * ld + e] - gi * Xcov[j]) * K.rls_inv_gamma;   // :59
        }
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) xp[i * ld + e] = st[i];
    up[e] = ac[0];
    up[ld + e] = ac[1];
}

template <typename ET>
__global__ void __launch_bounds__(SMALL_BLOCK)
    rls_predict_kernel(const ET* __restrict__ W, const ET* __restrict__ xp, const ET* __restrict__ up,
                       const ET* __restrict__ state, const ET* __restrict__ action, ET* out, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    ET X[9];
#pragma unroll
    for (int i = 0; i < 7; ++i) X[i] = state[i * ld + e] - xp[i * ld + e];      // model.py:65-66
    X[7] = action[e] - up[e];
    X[8] = action[ld + e] - up[ld + e];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        ET a = state[j * ld + e];
#pragma unroll
        for (int i = 0; i < 9; ++i) a += W[(i * 7 + j) * ld + e] * X[i];        // :67
        out[j * ld + e] = a;
    }
}

// ---------------------------------------------------------------------------------------------
// DHPAgent.get_action / update_networks on global-memory weights
// ---------------------------------------------------------------------------------------------
template <typename ET, typename NT, int NH>
__global__ void __launch_bounds__(SMALL_BLOCK)
    agent_action_kernel(H3AgentHyper h, const double* __restrict__ hyper, NT* net, const ET* __restrict__ state,
                        const ET* __restrict__ ref, NT* out, int64_t n, int64_t ld) {
    Math<NT>::prepare();                        // tanh table -> shared memory (all threads, before the bounds check)
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    double w_r;
    const AgentH<NT> H = load_hyper<NT>(h, hyper, ld, e, w_r);
    NetAcc<NT, NH, GlobalVec<NT>> N;
    N.v.p = net + e;
    N.v.ld = ld;
    const int ri = (h.tracked_state == 1) ? 0 : 1;
    const NT s0 = (NT)state[h.ac_state * ld + e];
    const NT s1 = (NT)(state[h.tracked_state * ld + e] - ref[ri * ld + e]);
    NT hh[NH], f;
    out[e] = actor_forward<NT, NH, 2>(N, s0, s1, H.scale, h.final_activation, hh, f);
}

template <typename ET, typename NT, int NH>
__global__ void __launch_bounds__(SMALL_BLOCK)
    agent_update_kernel(H3AgentHyper h, const double* __restrict__ hyper, NT* net, const ET* __restrict__ state,
                        const ET* __restrict__ next_state, const ET* __restrict__ ref,
                        const ET* __restrict__ next_ref, const NT* __restrict__ dr_ds, const NT* __restrict__ Fm,
                        const NT* __restrict__ Gm, int64_t n, int64_t ld) {
    Math<NT>::prepare();                        // tanh table -> shared memory (all threads, before the bounds check)
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    double w_r;
    const AgentH<NT> H = load_hyper<NT>(h, hyper, ld, e, w_r);
    NetAcc<NT, NH, GlobalVec<NT>> N;
    N.v.p = net + e;
    N.v.ld = ld;
    const int ri = (h.tracked_state == 1) ? 0 : 1;
    const int ia = h.ac_state, it = h.tracked_state;
    const NT s0 = (NT)state[ia * ld + e];
    const NT s1 = (NT)(state[it * ld + e] - ref[ri * ld + e]);
    const NT n0 = (NT)next_state[ia * ld + e];
    const NT n1 = (NT)(next_state[it * ld + e] - next_ref[ri * ld + e]);
    NT F[2][2], G[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = Fm[(i * 2 + j) * ld + e];
        G[i] = Gm[i * ld + e];
    }
    NT hh[NH], scr[NH], f;
    // chunks of 2 hidden units: weights are addressed in global memory, activations in local arrays
    (void)actor_forward<NT, NH, 2>(N, s0, s1, H.scale, h.final_activation, hh, f);
    agent_update<NT, NH, 2, false>(N, H, h.final_activation, s0, s1, n0, n1, dr_ds[e], dr_ds[ld + e], F, G, hh, f,
                                   scr);
}

// get_reward (agents.py:123-128) and get_transition_matrices (agents.py:130-140)
template <typename ET, typename NT>
__global__ void __launch_bounds__(SMALL_BLOCK)
    agent_reward_kernel(H3AgentHyper h, const double* __restrict__ hyper, const ET* __restrict__ next_state,
                        const ET* __restrict__ ref, ET* reward, NT* dr_ds, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    const ET w_r = (ET)(hyper != nullptr ? hyper[H3_HYPER_REWARD_WEIGHT * ld + e] : h.reward_weight);
    const int ri = (h.tracked_state == 1) ? 0 : 1;
    const ET e_tr = next_state[h.tracked_state * ld + e] - ref[ri * ld + e];
    reward[e] = -(e_tr * e_tr) * w_r;
    dr_ds[e] = (NT)0;
    dr_ds[ld + e] = (NT)((ET)-2 * e_tr * w_r);
}

template <typename ET, typename NT>
__global__ void __launch_bounds__(SMALL_BLOCK)
    agent_transition_kernel(H3AgentHyper h, const ET* __restrict__ W, NT* Fm, NT* Gm, int64_t n, int64_t ld) {
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    const int idx[2] = {h.ac_state, h.tracked_state};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) Fm[(i * 2 + j) * ld + e] = (NT)W[(idx[j] * 7 + idx[i]) * ld + e];
        Gm[i * ld + e] = (NT)W[((7 + h.channel) * 7 + idx[i]) * ld + e];
    }
}

// any other width: runtime-width versions on global memory (h3_closed_loop_rt.cuh)
template <typename ET, typename NT>
__global__ void __launch_bounds__(SMALL_BLOCK)
    agent_action_rt_kernel(H3AgentHyper h, const double* __restrict__ hyper, NT* net, const ET* __restrict__ state,
                           const ET* __restrict__ ref, NT* out, int64_t n, int64_t ld) {
    Math<NT>::prepare();                        // tanh table -> shared memory (all threads, before the bounds check)
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    double w_r;
    const AgentH<NT> H = load_hyper<NT>(h, hyper, ld, e, w_r);
    RtNet<NT> N{net + e, ld, h.n_hidden};
    const int ri = (h.tracked_state == 1) ? 0 : 1;
    const NT s0 = (NT)state[h.ac_state * ld + e];
    const NT s1 = (NT)(state[h.tracked_state * ld + e] - ref[ri * ld + e]);
    NT f;
    out[e] = actor_forward_rt(N, s0, s1, H.scale, h.final_activation, f);
}

template <typename ET, typename NT>
__global__ void __launch_bounds__(SMALL_BLOCK)
    agent_update_rt_kernel(H3AgentHyper h, const double* __restrict__ hyper, NT* net, const ET* __restrict__ state,
                           const ET* __restrict__ next_state, const ET* __restrict__ ref,
                           const ET* __restrict__ next_ref, const NT* __restrict__ dr_ds, const NT* __restrict__ Fm,
                           const NT* __restrict__ Gm, int64_t n, int64_t ld) {
    Math<NT>::prepare();                        // tanh table -> shared memory (all threads, before the bounds check)
    const int64_t e = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (e >= n) return;
    double w_r;
    const AgentH<NT> H = load_hyper<NT>(h, hyper, ld, e, w_r);
    RtNet<NT> N{net + e, ld, h.n_hidden};
    const int ri = (h.tracked_state == 1) ? 0 : 1;
    const int ia = h.ac_state, it = h.tracked_state;
    const NT s0 = (NT)state[ia * ld + e];
    const NT s1 = (NT)(state[it * ld + e] - ref[ri * ld + e]);
    const NT n0 = (NT)next_state[ia * ld + e];
    const NT n1 = (NT)(next_state[it * ld + e] - next_ref[ri * ld + e]);
    NT F[2][2], G[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) F[i][j] = Fm[(i * 2 + j) * ld + e];
        G[i] = Gm[i * ld + e];
    }
    agent_update_rt(N, H, h.final_activation, s0, s1, n0, n1, dr_ds[e], dr_ds[ld + e], F, G);
}

struct AgentArgs {
    const void *state, *next_state, *ref, *next_ref, *dr_ds, *F, *G;
    void* out;
};

template <typename ET, typename NT>
static int dispatch_agent(bool update, const H3AgentHyper& h, const double* hyper, void* net, const AgentArgs& a,
                          int64_t n, int64_t ld, cudaStream_t st) {
#define H3_AGENT_CASE(NH)                                                                                      \
    case NH:                                                                                                   \
        if (update)                                                                                            \
            agent_update_kernel<ET, NT, NH><<<grid_for(n), SMALL_BLOCK, 0, st>>>(                               \
                h, hyper, (NT*)net, (const ET*)a.state, (const ET*)a.next_state, (const ET*)a.ref,             \
                (const ET*)a.next_ref, (const NT*)a.dr_ds, (const NT*)a.F, (const NT*)a.G, n, ld);             \
        else                                                                                                   \
            agent_action_kernel<ET, NT, NH><<<grid_for(n), SMALL_BLOCK, 0, st>>>(                               \
                h, hyper, (NT*)net, (const ET*)a.state, (const ET*)a.ref, (NT*)a.out, n, ld);                  \
        break;
    switch (h.n_hidden) {
        H3_AGENT_CASE(4)
        H3_AGENT_CASE(6)
        H3_AGENT_CASE(8)
        H3_AGENT_CASE(10)
        H3_AGENT_CASE(12)
        H3_AGENT_CASE(16)
        default:
            if (h.n_hidden < 1) {
                set_error("agent kernels: n_hidden must be positive, got %d", h.n_hidden);
                return H3_E_ARG;
            }
            if (update)
                agent_update_rt_kernel<ET, NT><<<grid_for(n), SMALL_BLOCK, 0, st>>>(
                    h, hyper, (NT*)net, (const ET*)a.state, (const ET*)a.next_state, (const ET*)a.ref,
                    (const ET*)a.next_ref, (const NT*)a.dr_ds, (const NT*)a.F, (const NT*)a.G, n, ld);
            else
                agent_action_rt_kernel<ET, NT><<<grid_for(n), SMALL_BLOCK, 0, st>>>(
                    h, hyper, (NT*)net, (const ET*)a.state, (const ET*)a.ref, (NT*)a.out, n, ld);
            break;
    }
#undef H3_AGENT_CASE
    return check_launch(update ? "agent_update_kernel" : "agent_action_kernel");
}

// ---------------------------------------------------------------------------------------------
// FMA-chain peak (roofline denominator): 8 independent chains per thread
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, T* out) {
    T a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
    const T b = (T)0.999999, c = (T)1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;
    }
    T s = (T)0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == (T)123456789) out[0] = s;      // never true; keeps the chains alive
}

// direct access to the kernel's own fp64 math (tests/test_gpu_math.py)
__global__ void __launch_bounds__(SMALL_BLOCK) math_probe_kernel(int fn, const double* __restrict__ x, double* y,
                                                                 double* y2, int64_t n) {
    Math<double>::prepare();
    const int64_t i = (int64_t)blockIdx.x * SMALL_BLOCK + threadIdx.x;
    if (i >= n) return;
    if (fn == 0) {
        y[i] = Math<double>::tanh_(x[i]);       // table-driven (3-DOF kernels)
    } else if (fn == 2) {
        y[i] = MathExp::tanh_(x[i]);            // table-free (6-DOF per-tick kernels)
    } else if (fn == 3) {
        y[i] = atan2_nr(x[i], y2[i]);           // y2 is the second INPUT here
    } else if (fn == 4) {
        y[i] = div_nr(x[i], y2[i]);
    } else if (fn == 5) {
        y[i] = sqrt_nr(x[i]);
    } else {
        double s, c;
        Math<double>::sincos_(x[i], &s, &c);
        y[i] = s;
        y2[i] = c;
    }
}


// ---------------------------------------------------------------------------------------------
// TMEM as thread-private storage (h3_tmem.cuh): functional probe + ld/st cost in cycles
// ---------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) tmem_probe_kernel(double* out, long long* cyc, int reps, int64_t nthreads) {
    __shared__ uint32_t slot;
    constexpr int NCOLS = 256 * (WARPS / 4);
    const uint32_t base = tmem_alloc(&slot, NCOLS);
    const int warp = threadIdx.x >> 5;
    const uint32_t addr = tmem_addr(base, 256 * (warp >> 2));       // 128 doubles of private storage per thread
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // fill: word w of thread gt = gt * 1024 + w
#pragma unroll 1
    for (int w0 = 0; w0 < 128; w0 += 16) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (double)(gt * 1024 + w0 + i);
        tmem_store<double, 16>(addr, w0, v);
    }
    tmem_wait_st();
    // read back with odd-sized pieces at runtime offsets: 12-word units (x16 + x8), then the tail
    {
#pragma unroll 1
        for (int k = 0; k < 10; ++k) {
            double v[12];
            tmem_load<double, 12>(addr, 12 * k, v);
#pragma unroll
            for (int i = 0; i < 12; ++i) out[(int64_t)(12 * k + i) * nthreads + gt] = v[i];
        }
        double t[8];
        tmem_load<double, 8>(addr, 120, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[(int64_t)(120 + i) * nthreads + gt] = t[i];
    }
    // read-modify-write round trip, as the agent update does: w += 1 for every word, `reps` times
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int k = 0; k < 10; ++k) {
            double v[12];
            tmem_load<double, 12>(addr, 12 * k, v);
#pragma unroll
            for (int i = 0; i < 12; ++i) v[i] += 1.0;
            tmem_store<double, 12>(addr, 12 * k, v);
        }
        tmem_wait_st();
    }
    long long t1 = clock64();
    // dependent load latency: single-double loads whose address depends on the previous value
    double acc = 0.0;
    int w = 0;
#pragma unroll 1
    for (int r = 0; r < reps * 8; ++r) {
        double v[1];
        tmem_load<double, 1>(addr, w, v);
        acc += v[0];
        w = (w + 1 + (__double2loint(v[0]) & 0)) & 127;
    }
    long long t2 = clock64();
    double fin[8];
    tmem_load<double, 8>(addr, 0, fin);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(int64_t)(128 + i) * nthreads + gt] = fin[i];
    out[(int64_t)136 * nthreads + gt] = acc;
    if (threadIdx.x == 0) {
        cyc[2 * blockIdx.x] = t1 - t0;
        cyc[2 * blockIdx.x + 1] = t2 - t1;
    }
    tmem_free(base, NCOLS);
}

static bool dtype_ok(int d) { return d == H3_F64 || d == H3_F32; }

}  // namespace h3

using namespace h3;

extern "C" {

int h3_version(void) { return H3_VERSION; }

const char* h3_last_error(void) { return g_err; }

void h3_env_params_default(H3EnvParams* out, double dt) {
    // heli_models.py:117-130
    out->g = 9.81;
    out->cl_alpha = 5.7;
    out->blade_solidity = 0.075;
    out->lock_number = 6;
    out->cds = 1.5;
    out->mass = 2200;
    out->rho = 1.225;
    out->v_tip = 200;
    out->rotor_radius = 7.32;
    out->iy = 10615;
    out->mast = 1;
    out->omega = out->v_tip / out->rotor_radius;
    out->area = 3.14159265358979323846 * out->rotor_radius * out->rotor_radius;
    out->tau = 0.1;
    out->dt = dt;
}

int h3_trim(const H3EnvParams* env, const double* v, double* trim, double* state, int64_t n, int64_t ld,
            void* stream) {
    if (!env || !v || !trim || !state || ld < n) {
        set_error("h3_trim: bad argument");
        return H3_E_ARG;
    }
    if (n <= 0) return H3_OK;
    trim_kernel<<<grid_for(n), SMALL_BLOCK, 0, (cudaStream_t)stream>>>(*env, v, trim, state, n, ld);
    return check_launch("trim_kernel");
}

int h3_env_step(const H3Config* cfg, const H3Batch* b, const void* actions, const double* excitation,
                int32_t exc_len, int32_t n_ticks, void* states_out, int64_t n_out, int64_t ld_out, void* stream) {
    if (!cfg || !b || !b->state || !b->t || !b->step || !b->flags || !b->term_step || (!actions && !b->trim) ||
        !dtype_ok(cfg->env_dtype) || b->ld < b->n) {
        set_error("h3_env_step: bad argument");
        return H3_E_ARG;
    }
    if (b->n <= 0 || n_ticks <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg->env_dtype == H3_F64)
        env_step_kernel<double><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            make_envk<double>(*cfg), cfg->env.dt, (double*)b->state, b->t, b->step, b->flags, b->term_step, b->trim,
            (const double*)actions, excitation, exc_len, n_ticks, (double*)states_out, n_out, ld_out, b->n, b->ld);
    else
        env_step_kernel<float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            make_envk<float>(*cfg), cfg->env.dt, (float*)b->state, b->t, b->step, b->flags, b->term_step, b->trim,
            (const float*)actions, excitation, exc_len, n_ticks, (float*)states_out, n_out, ld_out, b->n, b->ld);
    return check_launch("env_step_kernel");
}

int h3_get_ref(const H3Config* cfg, const H3Batch* b, void* stream) {
    if (!cfg || !b || !b->state || !b->t || !b->ref || !dtype_ok(cfg->env_dtype) ||
        (cfg->task == H3_TASK_VELOCITY && !b->corr_u) || (cfg->task == H3_TASK_STOP_OVER_POINT && !b->theta_trim) ||
        cfg->task < 0 || cfg->task > 2) {
        set_error("h3_get_ref: bad argument");
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg->env_dtype == H3_F64)
        get_ref_kernel<double><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            make_envk<double>(*cfg), cfg->task, cfg->env.dt, cfg->ref_amp, cfg->ref_period, (const double*)b->state,
            b->t, (double*)b->ref, (double*)b->corr_u, b->theta_trim, b->ref_par, b->n, b->ld);
    else
        get_ref_kernel<float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            make_envk<float>(*cfg), cfg->task, cfg->env.dt, cfg->ref_amp, cfg->ref_period, (const float*)b->state,
            b->t, (float*)b->ref, (float*)b->corr_u, b->theta_trim, b->ref_par, b->n, b->ld);
    return check_launch("get_ref_kernel");
}

int h3_pid_collective(const H3Config* cfg, const H3Batch* b, void* collective_out, void* stream) {
    if (!cfg || !b || !b->state || !b->pid_corr || !collective_out || !dtype_ok(cfg->env_dtype)) {
        set_error("h3_pid_collective: bad argument");
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg->env_dtype == H3_F64)
        pid_kernel<double><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(make_envk<double>(*cfg), (const double*)b->state,
                                                                   (double*)b->pid_corr, (double*)collective_out,
                                                                   b->n, b->ld);
    else
        pid_kernel<float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(make_envk<float>(*cfg), (const float*)b->state,
                                                                  (float*)b->pid_corr, (float*)collective_out, b->n,
                                                                  b->ld);
    return check_launch("pid_kernel");
}

int h3_rls_update(const H3Config* cfg, const H3Batch* b, const void* state, const void* action,
                  const void* next_state, void* stream) {
    if (!cfg || !b || !b->rls_W || !b->rls_cov || !b->rls_x || !b->rls_u || !b->rls_skip || !state || !action ||
        !next_state || !dtype_ok(cfg->env_dtype)) {
        set_error("h3_rls_update: bad argument");
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg->env_dtype == H3_F64)
        rls_update_kernel<double><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            make_envk<double>(*cfg), cfg->rls_symmetric, (double*)b->rls_W, (double*)b->rls_cov, (double*)b->rls_x,
            (double*)b->rls_u, b->rls_skip, (const double*)state, (const double*)action, (const double*)next_state,
            b->n, b->ld);
    else
        rls_update_kernel<float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            make_envk<float>(*cfg), cfg->rls_symmetric, (float*)b->rls_W, (float*)b->rls_cov, (float*)b->rls_x,
            (float*)b->rls_u, b->rls_skip, (const float*)state, (const float*)action, (const float*)next_state, b->n,
            b->ld);
    return check_launch("rls_update_kernel");
}

int h3_rls_predict(const H3Config* cfg, const H3Batch* b, const void* state, const void* action, void* out,
                   void* stream) {
    if (!cfg || !b || !b->rls_W || !b->rls_x || !b->rls_u || !state || !action || !out || !dtype_ok(cfg->env_dtype)) {
        set_error("h3_rls_predict: bad argument");
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg->env_dtype == H3_F64)
        rls_predict_kernel<double><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            (const double*)b->rls_W, (const double*)b->rls_x, (const double*)b->rls_u, (const double*)state,
            (const double*)action, (double*)out, b->n, b->ld);
    else
        rls_predict_kernel<float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
            (const float*)b->rls_W, (const float*)b->rls_x, (const float*)b->rls_u, (const float*)state,
            (const float*)action, (float*)out, b->n, b->ld);
    return check_launch("rls_predict_kernel");
}

static bool agent_desc_ok(const H3AgentHyper& h) {
    return h.tracked_state >= 0 && h.tracked_state <= 6 && h.ac_state >= 0 && h.ac_state <= 6 && h.channel >= 0 &&
           h.channel <= 1;
}

// precision pair -> 0: (f64,f64), 1: (f64,f32), 2: (f32,f32), -1: unsupported
static int precision_pair(const H3Config* cfg) {
    if (cfg->env_dtype == H3_F64 && cfg->net_dtype == H3_F64) return 0;
    if (cfg->env_dtype == H3_F64 && cfg->net_dtype == H3_F32) return 1;
    if (cfg->env_dtype == H3_F32 && cfg->net_dtype == H3_F32) return 2;
    set_error("unsupported precision pair (env %d, net %d)", cfg->env_dtype, cfg->net_dtype);
    return -1;
}

static int agent_call(bool update, const H3Config* cfg, const H3Batch* b, int32_t which, const AgentArgs& a,
                      void* stream) {
    const char* name = update ? "update" : "action";
    if (!cfg || !b || !a.state || !a.ref || (which != 0 && which != 1) ||
        (update && (!a.next_state || !a.next_ref || !a.dr_ds || !a.F || !a.G)) || (!update && !a.out)) {
        set_error("h3_agent_%s: bad argument", name);
        return H3_E_ARG;
    }
    const H3AgentHyper& h = which == 0 ? cfg->col : cfg->lon;
    void* net = which == 0 ? b->col_net : b->lon_net;
    const double* hyper = which == 0 ? b->col_hyper : b->lon_hyper;
    if (!net || !agent_desc_ok(h)) {
        set_error("h3_agent_%s: bad agent description", name);
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (precision_pair(cfg)) {
        case 0: return dispatch_agent<double, double>(update, h, hyper, net, a, b->n, b->ld, st);
        case 1: return dispatch_agent<double, float>(update, h, hyper, net, a, b->n, b->ld, st);
        case 2: return dispatch_agent<float, float>(update, h, hyper, net, a, b->n, b->ld, st);
    }
    return H3_E_UNSUPPORTED;
}

int h3_agent_action(const H3Config* cfg, const H3Batch* b, int32_t which, const void* state, const void* ref,
                    void* action_out, void* stream) {
    AgentArgs a = {state, nullptr, ref, nullptr, nullptr, nullptr, nullptr, action_out};
    return agent_call(false, cfg, b, which, a, stream);
}

int h3_agent_update(const H3Config* cfg, const H3Batch* b, int32_t which, const void* state,
                    const void* next_state, const void* ref, const void* next_ref, const void* dr_ds, const void* F,
                    const void* G, void* stream) {
    AgentArgs a = {state, next_state, ref, next_ref, dr_ds, F, G, nullptr};
    return agent_call(true, cfg, b, which, a, stream);
}

int h3_agent_reward(const H3Config* cfg, const H3Batch* b, int32_t which, const void* next_state, const void* ref,
                    void* reward_out, void* dr_ds_out, void* stream) {
    if (!cfg || !b || !next_state || !ref || !reward_out || !dr_ds_out || (which != 0 && which != 1)) {
        set_error("h3_agent_reward: bad argument");
        return H3_E_ARG;
    }
    const H3AgentHyper& h = which == 0 ? cfg->col : cfg->lon;
    const double* hyper = which == 0 ? b->col_hyper : b->lon_hyper;
    if (!agent_desc_ok(h)) {
        set_error("h3_agent_reward: bad agent description");
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (precision_pair(cfg)) {
        case 0:
            agent_reward_kernel<double, double><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
                h, hyper, (const double*)next_state, (const double*)ref, (double*)reward_out, (double*)dr_ds_out, b->n,
                b->ld);
            break;
        case 1:
            agent_reward_kernel<double, float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
                h, hyper, (const double*)next_state, (const double*)ref, (double*)reward_out, (float*)dr_ds_out, b->n,
                b->ld);
            break;
        case 2:
            agent_reward_kernel<float, float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
                h, hyper, (const float*)next_state, (const float*)ref, (float*)reward_out, (float*)dr_ds_out, b->n,
                b->ld);
            break;
        default: return H3_E_UNSUPPORTED;
    }
    return check_launch("agent_reward_kernel");
}

int h3_agent_transition(const H3Config* cfg, const H3Batch* b, int32_t which, void* F_out, void* G_out,
                        void* stream) {
    if (!cfg || !b || !b->rls_W || !F_out || !G_out || (which != 0 && which != 1)) {
        set_error("h3_agent_transition: bad argument");
        return H3_E_ARG;
    }
    const H3AgentHyper& h = which == 0 ? cfg->col : cfg->lon;
    if (!agent_desc_ok(h)) {
        set_error("h3_agent_transition: bad agent description");
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (precision_pair(cfg)) {
        case 0:
            agent_transition_kernel<double, double><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
                h, (const double*)b->rls_W, (double*)F_out, (double*)G_out, b->n, b->ld);
            break;
        case 1:
            agent_transition_kernel<double, float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
                h, (const double*)b->rls_W, (float*)F_out, (float*)G_out, b->n, b->ld);
            break;
        case 2:
            agent_transition_kernel<float, float><<<grid_for(b->n), SMALL_BLOCK, 0, st>>>(
                h, (const float*)b->rls_W, (float*)F_out, (float*)G_out, b->n, b->ld);
            break;
        default: return H3_E_UNSUPPORTED;
    }
    return check_launch("agent_transition_kernel");
}

static int loop_dispatch(const H3Config* cfg, const H3Batch* b, int32_t step0, const double* ref_table,
                         const double* excitation, int32_t exc_len, int32_t n_ticks, const H3Traj* traj,
                         double* stats, void* stats_scratch, void* stream, int force_phase2, LoopInfo* info) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (precision_pair(cfg)) {
        case 0:
            return launch_closed_loop<double, double>(*cfg, *b, step0, ref_table, excitation, exc_len, n_ticks, traj,
                                                      stats, stats_scratch, st, force_phase2, info);
        case 1:
            return launch_closed_loop<double, float>(*cfg, *b, step0, ref_table, excitation, exc_len, n_ticks, traj,
                                                     stats, stats_scratch, st, force_phase2, info);
        case 2:
            return launch_closed_loop<float, float>(*cfg, *b, step0, ref_table, excitation, exc_len, n_ticks, traj,
                                                    stats, stats_scratch, st, force_phase2, info);
    }
    return H3_E_UNSUPPORTED;
}

int h3_closed_loop(const H3Config* cfg, const H3Batch* b, int32_t step0, const double* ref_table,
                   const double* excitation, int32_t exc_len, int32_t n_ticks, const H3Traj* traj, double* stats,
                   void* stats_scratch, void* stream) {
    if (!cfg || !b) {
        set_error("h3_closed_loop: null config / batch");
        return H3_E_ARG;
    }
    if (!b->state || !b->t || !b->ref || !b->corr_u || !b->trim || !b->theta_trim || !b->pid_corr || !b->rls_W ||
        !b->rls_cov || !b->rls_x || !b->rls_u || !b->rls_skip || !b->col_net || !b->lon_net || !b->step ||
        !b->flags || !b->term_step || !b->acc || b->ld < b->n) {
        set_error("h3_closed_loop: H3Batch has a null field or ld < n");
        return H3_E_ARG;
    }
    if (cfg->task < 0 || cfg->task > 2 || (cfg->task == H3_TASK_SINUSOID && cfg->ref_mode == H3_REF_TABLE &&
                                           (step0 < 0 || ref_table == nullptr))) {
        set_error("h3_closed_loop: H3_REF_TABLE needs a lockstep launch (step0 >= 0) and a reference table");
        return H3_E_ARG;
    }
    if (!agent_desc_ok(cfg->col) || !agent_desc_ok(cfg->lon)) {
        set_error("h3_closed_loop: bad agent description");
        return H3_E_ARG;
    }
    if (excitation == nullptr) exc_len = 0;
    return loop_dispatch(cfg, b, step0, ref_table, excitation, exc_len, n_ticks, traj, stats, stats_scratch, stream,
                         -1, nullptr);
}

int64_t h3_stats_scratch_bytes(int64_t n) {
    const int64_t blocks = (n + 7) / 8;     // upper bound on the grid of any kernel variant (>= 8 episodes per warp)
    return 16 + blocks * H3_N_STATS * (int64_t)sizeof(double);
}

int h3_closed_loop_info(const H3Config* cfg, int32_t phase2, int32_t* block, int32_t* smem_bytes,
                        int32_t* blocks_per_sm, int32_t* regs, int32_t* lanes_per_warp) {
    if (!cfg) {
        set_error("h3_closed_loop_info: null config");
        return H3_E_ARG;
    }
    LoopInfo info;
    H3Batch dummy;
    memset(&dummy, 0, sizeof(dummy));
    int rc = loop_dispatch(cfg, &dummy, 0, nullptr, nullptr, 0, 1, nullptr, nullptr, nullptr, nullptr,
                           phase2 ? 1 : 0, &info);
    if (rc != H3_OK) return rc;
    if (block) *block = info.block;
    if (smem_bytes) *smem_bytes = info.smem_bytes;
    if (blocks_per_sm) *blocks_per_sm = info.blocks_per_sm;
    if (regs) *regs = info.regs;
    if (lanes_per_warp) *lanes_per_warp = info.lanes_per_warp;
    return H3_OK;
}

int h3_math_probe(int32_t fn, const double* x, double* y, double* y2, int64_t n, void* stream) {
    if ((fn < 0 || fn > 5) || !x || !y || ((fn == 1 || fn == 3 || fn == 4) && !y2)) {
        set_error("h3_math_probe: bad argument");
        return H3_E_ARG;
    }
    if (n <= 0) return H3_OK;
    math_probe_kernel<<<grid_for(n), SMALL_BLOCK, 0, (cudaStream_t)stream>>>(fn, x, y, y2, n);
    return check_launch("math_probe_kernel");
}

int h3_fma_peak(int32_t dtype, int32_t iters, void* scratch_out, double* flops_out, void* stream) {
    if (!dtype_ok(dtype) || iters <= 0 || !scratch_out) {
        set_error("h3_fma_peak: bad argument");
        return H3_E_ARG;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    if (dtype == H3_F64)
        fma_peak_kernel<double><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, (double*)scratch_out);
    else
        fma_peak_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, (float*)scratch_out);
    if (flops_out) *flops_out = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    return check_launch("fma_peak_kernel");
}

int h3_init_episodes(const H3Config* cfg, const H3Batch* b, const H3Init* init, void* stream) {
    if (!cfg || !b || !init) {
        set_error("h3_init_episodes: null argument");
        return H3_E_ARG;
    }
    if (!b->state || !b->t || !b->ref || !b->corr_u || !b->trim || !b->theta_trim || !b->pid_corr || !b->rls_W ||
        !b->rls_cov || !b->rls_x || !b->rls_u || !b->rls_skip || !b->col_net || !b->lon_net || !b->step ||
        !b->flags || !b->term_step || !b->acc || b->ld < b->n) {
        set_error("h3_init_episodes: H3Batch has a null field or ld < n");
        return H3_E_ARG;
    }
    if (cfg->task < 0 || cfg->task > 2 || !agent_desc_ok(cfg->col) || !agent_desc_ok(cfg->lon) ||
        !dtype_ok(cfg->env_dtype) || !dtype_ok(cfg->net_dtype)) {
        set_error("h3_init_episodes: bad configuration");
        return H3_E_ARG;
    }
    if (b->n <= 0) return H3_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = grid_for(b->n);
    switch (precision_pair(cfg)) {
        case 0:
            init_episodes_kernel<double, double><<<grid, SMALL_BLOCK, 0, st>>>(
                cfg->env, make_envk<double>(*cfg), *b, *init, cfg->task, cfg->ref_amp, cfg->ref_period, cfg->col.n_hidden,
                cfg->lon.n_hidden);
            break;
        case 1:
            init_episodes_kernel<double, float><<<grid, SMALL_BLOCK, 0, st>>>(
                cfg->env, make_envk<double>(*cfg), *b, *init, cfg->task, cfg->ref_amp, cfg->ref_period, cfg->col.n_hidden,
                cfg->lon.n_hidden);
            break;
        case 2:
            init_episodes_kernel<float, float><<<grid, SMALL_BLOCK, 0, st>>>(
                cfg->env, make_envk<float>(*cfg), *b, *init, cfg->task, cfg->ref_amp, cfg->ref_period, cfg->col.n_hidden,
                cfg->lon.n_hidden);
            break;
        default:
            set_error("h3_init_episodes: unsupported precision pair");
            return H3_E_UNSUPPORTED;
    }
    return check_launch("init_episodes_kernel");
}

int h3_tmem_probe(int32_t warps, int32_t blocks, int32_t reps, double* out, int64_t* cycles, void* stream) {
    if ((warps != 4 && warps != 8) || blocks <= 0 || reps < 0 || !out || !cycles) {
        set_error("h3_tmem_probe: bad argument");
        return H3_E_ARG;
    }
    const int64_t nthreads = (int64_t)blocks * warps * 32;
    if (warps == 4)
        tmem_probe_kernel<4><<<blocks, 128, 0, (cudaStream_t)stream>>>(out, (long long*)cycles, reps, nthreads);
    else
        tmem_probe_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(out, (long long*)cycles, reps, nthreads);
    return check_launch("tmem_probe_kernel");
}

}  // extern "C"
