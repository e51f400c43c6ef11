// Device functions of one closed-loop control tick, per episode (one thread = one episode).
// Every function cites the reference statement it restates; nothing here is translated line by
// line -- trig is shared between model, PID and reference, the incremental model uses one
// reciprocal and (optionally) the symmetry of the covariance, the IDHP update is in closed form
// (no autograd) and re-uses the actor forward pass of get_action.
#pragma once
#include "../../include/heli3dof.h"
#include "h3_math.cuh"

namespace h3 {

// ----------------------------------------------------------------------------------------------
// launch-uniform constants, precomputed on the host in fp64 and cast to the compute type
// ----------------------------------------------------------------------------------------------
template <typename T>
struct EnvK {
    T dt, g, inv_omega, inv_vtip, k_a1_q, k_bem, k_thrust, k_drag, inv_mass, k_q, inv_tau;
    T pid_kp, pid_ki, pid_kd, pid_href, pid_hi;
    T env_kp, env_ki, env_kd, env_ref;
    T rls_gamma, rls_inv_gamma;
    T rad2deg;
};

template <typename T>
__host__ inline EnvK<T> make_envk(const H3Config& c) {
    const H3EnvParams& e = c.env;
    EnvK<T> k;
    k.dt = (T)e.dt;
    k.g = (T)e.g;
    k.inv_omega = (T)(1.0 / e.omega);
    k.inv_vtip = (T)(1.0 / e.v_tip);
    k.k_a1_q = (T)(-16.0 / e.lock_number);
    k.k_bem = (T)(e.cl_alpha * e.blade_solidity / 4.0);
    k.k_thrust = (T)(e.rho * e.v_tip * e.v_tip * e.area);
    k.k_drag = (T)(e.cds * 0.5 * e.rho / e.mass);
    k.inv_mass = (T)(1.0 / e.mass);
    k.k_q = (T)(e.mast / e.iy);
    k.inv_tau = (T)(1.0 / e.tau);
    k.pid_kp = (T)c.pid_kp;
    k.pid_ki = (T)c.pid_ki;
    k.pid_kd = (T)c.pid_kd;
    k.pid_href = (T)c.pid_href;
    k.pid_hi = (T)(10.0 * (3.14159265358979323846 / 180.0));
    k.env_kp = (T)c.env_kp;
    k.env_ki = (T)c.env_ki;
    k.env_kd = (T)c.env_kd;
    k.env_ref = (T)c.env_ref;
    k.rls_gamma = (T)c.rls_gamma;
    k.rls_inv_gamma = (T)(1.0 / c.rls_gamma);
    k.rad2deg = (T)(180.0 / 3.14159265358979323846);
    return k;
}

// ----------------------------------------------------------------------------------------------
// Helicopter3DOF.step  (heli_models.py:157-198, :210-213)
//   sth/cth = sin/cos(theta) of the pre-step state (shared with the PID).
//   atan2 is eliminated: cos/sin(alpha_c) with alpha_c = cyc - atan2(w,u) are expanded with
//   cos(atan2(w,u)) = u/V, sin(atan2(w,u)) = w/V, so v_diml*cos(alpha_c) = (u cos cyc + w sin cyc)/v_tip.
// ----------------------------------------------------------------------------------------------
//   sc/cc = sin/cos(cyc) (evaluated together with theta's by the caller, four interleaved chains).
template <typename T>
H3_DEV void env_step(const EnvK<T>& K, T (&s)[7], T col, T cyc, T sth, T cth, T sc, T cc, bool& done) {
    typedef Math<T> M;
    const T u = s[2], w = s[3], q = s[5], li = s[6];
    const T q_diml = q * K.inv_omega;                       // :159
    const T V = M::sqrt_(u * u + w * w);                    // :160  (vv, :175)
    const T mu = (u * cc + w * sc) * K.inv_vtip;            // :163
    const T lc = (u * sc - w * cc) * K.inv_vtip;            // :164
    const T lsum = lc + li;
    const T a1 = M::div_(K.k_a1_q * q_diml + (T)(8.0 / 3.0) * mu * col - (T)2 * mu * lsum,
                         (T)1 - mu * mu * (T)0.5);          // :167
    const T ct_bem = K.k_bem * ((T)(2.0 / 3.0) * col * ((T)1 + (T)1.5 * mu * mu) - lsum);   // :170
    T s1, c1;
    M::sincos_(cyc - a1, &s1, &c1);
    const T gx = (u * c1 + w * s1) * K.inv_vtip;            // v_diml cos(alpha_c - a1)
    const T gy = (u * s1 - w * c1) * K.inv_vtip + li;       // v_diml sin(alpha_c - a1) + lambda_i
    const T ct_glau = (T)2 * li * M::sqrt_(gx * gx + gy * gy);   // :171
    const T thrust = ct_bem * K.k_thrust;                   // :174
    const T tm = thrust * K.inv_mass;
    const T dr = K.k_drag * V;
    const T x_dot = u * cth + w * sth;                      // :176
    const T z_dot = -(u * sth - w * cth);                   // :177
    const T u_dot = -K.g * sth - dr * u + tm * s1 - q * w;  // :178-181
    const T w_dot = K.g * cth - dr * w - tm * c1 + q * u;   // :182-185
    const T q_dot = -thrust * K.k_q * s1;                   // :187
    const T li_dot = (ct_bem - ct_glau) * K.inv_tau;        // :188
    s[0] += x_dot * K.dt;                                   // :191-198 explicit Euler
    s[1] += z_dot * K.dt;
    s[2] += u_dot * K.dt;
    s[3] += w_dot * K.dt;
    s[4] += q * K.dt;
    s[5] += q_dot * K.dt;
    s[6] += li_dot * K.dt;
    done = M::abs_(s[4] * K.rad2deg) > (T)90;               // :210-213
}

// ----------------------------------------------------------------------------------------------
// CollectivePID3DOF.__call__  (PID.py:16-24)
// ----------------------------------------------------------------------------------------------
template <typename T>
H3_DEV T pid_collective(const EnvK<T>& K, const T (&s)[7], T sth, T cth, T& corr) {
    const T hdot_ref = K.pid_kd * (K.pid_href + s[1]);      // h_ref - -obs[1]
    const T hdot = s[2] * sth - s[3] * cth;
    const T err = hdot_ref - hdot;
    const T col = ((T)5 + K.pid_kp * err + K.pid_ki * corr) * (T)(3.14159265358979323846 / 180.0);
    corr += K.dt * err;                                     // integrator AFTER use (:23)
    return clip_nan(col, (T)0, K.pid_hi);
}

// ----------------------------------------------------------------------------------------------
// Helicopter3DOF.get_ref  (heli_models.py:217-255).  Reference generation always runs in fp64
// (t is accumulated in fp64 for every precision path).
// ----------------------------------------------------------------------------------------------
H3_DEV double sinusoid_ref(double t_plus_dt, double amp, double period) {
    const double x = div_nr(3.14159265358979323846 * t_plus_dt, period);
    double sx, cx;
    sincos_f64(x, &sx, &cx);
    // sin x + sin 2x = sin x (1 + 2 cos x)
    return (div_nr(amp, 1.76) * (sx * (1.0 + 2.0 * cx))) * (3.14159265358979323846 / 180.0);
}

template <typename T>
H3_DEV T velocity_ref(const EnvK<T>& K, T u, T& corr_u) {   // :239-243 incl. the integrator side effect
    const T u_err = K.env_ref - u;
    const T th = (K.env_kp * u_err + K.env_ki * corr_u) * (T)(3.14159265358979323846 / 180.0);
    corr_u += u_err * K.dt;
    return th;
}

template <typename T>
H3_DEV T stop_over_point_ref(const EnvK<T>& K, T x, T u, T theta_trim) {    // :245-250
    const T x_err = K.env_ref - x;
    const T th = (K.env_kp * x_err + K.env_kd * u) * (T)(3.14159265358979323846 / 180.0);
    return (theta_trim > th) ? theta_trim : th;             // python max(th, trim): NaN th stays
}

// ----------------------------------------------------------------------------------------------
// RecursiveLeastSquares.update  (model.py:41-59), n = 9 regressors, m = 7 outputs.
//   Wm(i,j): weight accessor (9x7).  Cv: covariance accessor -- full (9x9) or symmetric-packed.
// ----------------------------------------------------------------------------------------------
// A compiler-level fence: memory operations are not moved across it.  The unrolled RLS loops stream the
// 63 + 45 shared-memory words in groups with an explicit look-ahead of DEPTH groups: the loads of group
// g + DEPTH are issued before the FMAs of group g, and the fence stops ptxas from hoisting ALL loads to the
// top (> 200 registers) or sinking them next to their use (latency exposed with one warp per scheduler).
template <bool ON>
H3_DEV void row_fence() {
    if (ON) asm volatile("" ::: "memory");
}

template <typename T, int NG, int GS, int DEPTH, bool FENCE, typename Load, typename Body>
H3_DEV void stream_groups(Load load, Body body) {
    T buf[DEPTH + 1][GS];
#pragma unroll
    for (int g = 0; g < DEPTH && g < NG; ++g) load(g, buf[g]);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (g + DEPTH < NG) load(g + DEPTH, buf[(g + DEPTH) % (DEPTH + 1)]);
        row_fence<FENCE>();
        body(g, buf[g % (DEPTH + 1)]);
    }
}

template <typename T, int DEPTH, typename WAcc, typename CAcc>
H3_DEV void rls_update(const EnvK<T>& K, WAcc& Wm, CAcc& Cv, const T (&X)[9], const T (&Y)[7]) {
    constexpr bool F = WAcc::kFence;
    // gain first: it needs the covariance only (model.py:53-55)
    T covX[9], Xcov[9];
    Cv.template matvec<DEPTH>(X, covX, Xcov);                           // :53-54
    const T d0 = (K.rls_gamma + Xcov[0] * X[0]) + Xcov[1] * X[1] + Xcov[2] * X[2];      // :55 (three partial sums)
    const T d1 = Xcov[3] * X[3] + Xcov[4] * X[4] + Xcov[5] * X[5];
    const T d2 = Xcov[6] * X[6] + Xcov[7] * X[7] + Xcov[8] * X[8];
    const T inv_d = Math<T>::div_((T)1, d0 + (d1 + d2));
    T g[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] = covX[i] * inv_d;
    // then ONE pass over W, column by column: the prediction error of output j (model.py:49-50) and the update of
    // column j (:58) need nothing from the other columns, so each weight is loaded once, not twice.
    stream_groups<T, 7, 9, DEPTH, F>(
        [&](int j, T (&w)[9]) {
#pragma unroll
            for (int i = 0; i < 9; ++i) w[i] = Wm.dyn(i, j);
        },
        [&](int j, T (&w)[9]) {
            // two partial sums keep the 9-long dependent chain short
            const T ea = ((Y[j] - X[0] * w[0]) - X[2] * w[2]) - X[4] * w[4] - X[6] * w[6] - X[8] * w[8];
            const T eb = ((X[1] * w[1] + X[3] * w[3]) + X[5] * w[5]) + X[7] * w[7];
            const T err = ea - eb;
#pragma unroll
            for (int i = 0; i < 9; ++i) Wm.ref(i, j) = w[i] + g[i] * err;
        });
    Cv.template downdate<DEPTH>(g, Xcov, K.rls_inv_gamma);              // :59
}

// covariance policies ------------------------------------------------------------------------
// Element storage is delegated to `Vec` (operator[](int) -> T&), so the same code runs on a
// register/local array or on a strided shared-memory column.
template <typename T, typename Vec>
struct CovFull {            // model.py:53-54 literally: cov.X and X^T.cov computed separately
    Vec v;                  // 81 elements, row-major
    template <int DEPTH>
    H3_DEV void matvec(const T (&X)[9], T (&covX)[9], T (&Xcov)[9]) {
#pragma unroll
        for (int i = 0; i < 9; ++i) { covX[i] = (T)0; Xcov[i] = (T)0; }
        stream_groups<T, 9, 9, DEPTH, Vec::kFence>(
            [&](int i, T (&c)[9]) {
#pragma unroll
                for (int j = 0; j < 9; ++j) c[j] = v[i * 9 + j];
            },
            [&](int i, T (&c)[9]) {
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    covX[i] += c[j] * X[j];
                    Xcov[j] += X[i] * c[j];
                }
            });
    }
    template <int DEPTH>
    H3_DEV void downdate(const T (&g)[9], const T (&Xcov)[9], T inv_gamma) {
        stream_groups<T, 9, 9, DEPTH, Vec::kFence>(
            [&](int i, T (&c)[9]) {
#pragma unroll
                for (int j = 0; j < 9; ++j) c[j] = v[i * 9 + j];
            },
            [&](int i, T (&c)[9]) {
#pragma unroll
                for (int j = 0; j < 9; ++j) v[i * 9 + j] = (c[j] - g[i] * Xcov[j]) * inv_gamma;
            });
    }
};

template <typename T, typename Vec>
struct CovSym {             // upper triangle only (45 elements): cov == cov^T is an invariant of
    Vec v;                  // model.py:59 when it holds initially (model.py:85: identity * c)
    __host__ __device__ static constexpr int idx(int i, int j) { return i * 9 - (i * (i - 1)) / 2 + (j - i); }   // i <= j
    // The 45 elements are visited in 5 groups of 9, diagonal by diagonal (group 0: the main diagonal; group
    // g: the 9-g elements of diagonal g and the g elements of diagonal 9-g), so that consecutive FMAs of the
    // matrix-vector product hit different accumulators instead of one 9-long dependent chain per row.
    __host__ __device__ static constexpr int gi(int g, int e) { return g == 0 ? e : (e < 9 - g ? e : e - (9 - g)); }
    __host__ __device__ static constexpr int gj(int g, int e) {
        return g == 0 ? e : (e < 9 - g ? e + g : e - (9 - g) + (9 - g));
    }
    template <int DEPTH>
    H3_DEV void matvec(const T (&X)[9], T (&covX)[9], T (&Xcov)[9]) {
#pragma unroll
        for (int i = 0; i < 9; ++i) covX[i] = (T)0;
        stream_groups<T, 5, 9, DEPTH, Vec::kFence>(
            [&](int g, T (&c)[9]) {
#pragma unroll
                for (int e = 0; e < 9; ++e) c[e] = v[idx(gi(g, e), gj(g, e))];
            },
            [&](int g, T (&c)[9]) {
#pragma unroll
                for (int e = 0; e < 9; ++e) {
                    const int i = gi(g, e), j = gj(g, e);
                    covX[i] += c[e] * X[j];
                    if (j != i) covX[j] += c[e] * X[i];
                }
            });
#pragma unroll
        for (int i = 0; i < 9; ++i) Xcov[i] = covX[i];
    }
    template <int DEPTH>
    H3_DEV void downdate(const T (&g)[9], const T (&Xcov)[9], T inv_gamma) {
        // forgetting factor 1 (main_3dof.py:49): x * 1 == x bit for bit, so the multiplication is left out (launch-
        // uniform branch; 45 of the tick's FP64 instructions)
        const bool unit = inv_gamma == (T)1;
        stream_groups<T, 5, 9, DEPTH, Vec::kFence>(
            [&](int gr, T (&c)[9]) {
#pragma unroll
                for (int e = 0; e < 9; ++e) c[e] = v[idx(gi(gr, e), gj(gr, e))];
            },
            [&](int gr, T (&c)[9]) {
#pragma unroll
                for (int e = 0; e < 9; ++e) {
                    const int i = gi(gr, e), j = gj(gr, e);
                    const T d = c[e] - g[i] * Xcov[j];
                    v[idx(i, j)] = unit ? d : d * inv_gamma;
                }
            });
    }
};

// ----------------------------------------------------------------------------------------------
// IDHP agent (agents.py).  Net: accessor with a1(k,i) a2(k) c1(k,i) c2(j,k) t1(k,i) t2(j,k).
//
// Loops over the hidden units run in chunks of U units: U == NH unrolls everything (weights may then
// live in registers), U < NH keeps an outer `#pragma unroll 1` loop with runtime unit indices (weights,
// activations and scratch must then be addressable, i.e. in shared or global memory).  The chunk width
// bounds how many independent tanh chains the compiler interleaves -- enough to cover the FP64 pipe
// latency from a single warp, few enough not to blow the register file.
// ----------------------------------------------------------------------------------------------
template <typename NT>
struct AgentH {             // hyper-parameters in network precision
    NT gamma, lr_a, lr_c, tau, scale;
};

// DHPActor.forward (agents.py:59-63); keeps h = tanh(W1 s) and f = final_activation(W2 h) for the update
//   M: math policy (Math<NT>; MathExp for kernels that do not stage the tanh table)
template <typename NT, int NH, int U, typename M = Math<NT>, typename Net, typename HVec>
H3_DEV NT actor_forward(Net& net, NT s0, NT s1, NT scale, int act, HVec& h, NT& f) {
    static_assert(NH % U == 0, "chunk width must divide the hidden width");
    NT y = (NT)0;
#pragma unroll 1
    for (int k0 = 0; k0 < NH; k0 += U) {
        NT pre[U], hk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = (U == NH) ? u : k0 + u;
            pre[u] = net.a1(k, 0) * s0 + net.a1(k, 1) * s1;
        }
        M::template tanh_vec<U>(pre, hk);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = (U == NH) ? u : k0 + u;
            h[k] = hk[u];
            y += net.a2(k) * hk[u];
        }
    }
    f = (act == H3_ACT_TANH) ? M::tanh_(y) : M::sigmoid_(y);
    return scale * f;
}

// get_reward + get_transition_matrices + update_networks (agents.py:123-196), closed form.
//   s0,s1 = augmented state of (obs, ref); n0,n1 = augmented state of (next_obs, next_ref);
//   dr_ds = [dr0, dr] (get_reward: dr0 = 0, dr = -2 e w_r, already in network precision);
//   F[i][j], G[i] from the POST-update RLS weights;
//   h_a, f = actor forward of the same (obs, ref) with the pre-update actor (== get_action's pass);
//   scr: NH words of scratch for the critic's hidden activations between the two passes.
//   TAU1: tau_target_critic == 1 for every episode of the launch -> the target critic IS the critic after
//   every update (agents.py:196: 1*p + 0*t), so its rows are neither kept on chip nor updated; lambda_t2
//   is evaluated with the critic's (pre-update) weights, which are the target's.
template <typename NT, int NH, int U, bool TAU1, typename M = Math<NT>, typename Net, typename HVec, typename SVec>
H3_DEV void agent_update(Net& net, const AgentH<NT>& H, int act, NT s0, NT s1, NT n0, NT n1, NT dr0, NT dr,
                         const NT (&F)[2][2], const NT (&G)[2], HVec& h_a, NT f, SVec& scr) {
    static_assert(NH % U == 0, "chunk width must divide the hidden width");
    // action.backward() (:174): delta = da/dy, p_k = da/d(pre-activation of hidden k)
    const NT delta = (act == H3_ACT_TANH) ? H.scale * ((NT)1 - f * f) : H.scale * f * ((NT)1 - f);
    // pass 1: critic(s), target_critic(s') (:170-171) and da/ds
    NT lam1[2] = {(NT)0, (NT)0}, lam2[2] = {(NT)0, (NT)0}, da_ds[2] = {(NT)0, (NT)0};
#pragma unroll 1
    for (int k0 = 0; k0 < NH; k0 += U) {
        NT pre[2 * U], act[2 * U];      // critic on s and target critic on s' of the chunk: 2U independent tanh
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = (U == NH) ? u : k0 + u;
            pre[u] = net.c1(k, 0) * s0 + net.c1(k, 1) * s1;
            pre[U + u] = TAU1 ? net.c1(k, 0) * n0 + net.c1(k, 1) * n1 : net.t1(k, 0) * n0 + net.t1(k, 1) * n1;
        }
        M::template tanh_vec<2 * U>(pre, act);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k = (U == NH) ? u : k0 + u;
            const NT hc = act[u], ht = act[U + u];
            scr[k] = hc;
            lam1[0] += net.c2(0, k) * hc;
            lam1[1] += net.c2(1, k) * hc;
            if (TAU1) {
                lam2[0] += net.c2(0, k) * ht;
                lam2[1] += net.c2(1, k) * ht;
            } else {
                lam2[0] += net.t2(0, k) * ht;
                lam2[1] += net.t2(1, k) * ht;
            }
            const NT ha = h_a[k];
            const NT pk = delta * net.a2(k) * ((NT)1 - ha * ha);
            da_ds[0] += net.a1(k, 0) * pk;
            da_ds[1] += net.a1(k, 1) * pk;
        }
    }
    // target = dr_ds + gamma * lambda_t2 ; error_critic = lambda_t1 - target (F + G da_ds)  (:178-179)
    const NT tg0 = dr0 + H.gamma * lam2[0];     // dr_ds[0] is 0 in get_reward (agents.py:124)
    const NT tg1 = dr + H.gamma * lam2[1];
    NT e_c[2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
        e_c[j] = lam1[j] - (tg0 * (F[0][j] + G[0] * da_ds[j]) + tg1 * (F[1][j] + G[1] * da_ds[j]));
    // pass 2: SGD (:184-188): wa += lr_a (target.G) da/dwa ; wc -= lr_c dL/dwc, then soft target update (:195-196)
    const NT ka = H.lr_a * (tg0 * G[0] + tg1 * G[1]);
    const NT one_m_tau = (NT)1 - H.tau;
    // step sizes folded into the launch-invariant factors of a unit's update (one rounding earlier than the literal
    // `w += lr * (a * b)`, 7 multiplications per hidden unit fewer; same factors in every kernel)
    const NT ka_s0 = ka * s0, ka_s1 = ka * s1, ka_d = ka * delta;
    const NT lc_s0 = H.lr_c * s0, lc_s1 = H.lr_c * s1, lc_e0 = H.lr_c * e_c[0], lc_e1 = H.lr_c * e_c[1];
    // (narrower chunks than pass 1: 13 values per unit are live here)
    constexpr int V = (U == NH) ? NH : ((NH % 2 == 0) ? 2 : 1);
#pragma unroll 1
    for (int k0 = 0; k0 < NH; k0 += V) {
        // all loads of the chunk first (one shared-memory latency per chunk, not per element) ...
        NT hc[V], ha[V], a10[V], a11[V], a2k[V], c10[V], c11[V], c20[V], c21[V], t10[V], t11[V], t20[V], t21[V];
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int k = (V == NH) ? u : k0 + u;
            hc[u] = scr[k];
            ha[u] = h_a[k];
            a10[u] = net.a1(k, 0);
            a11[u] = net.a1(k, 1);
            a2k[u] = net.a2(k);
            c10[u] = net.c1(k, 0);
            c11[u] = net.c1(k, 1);
            c20[u] = net.c2(0, k);
            c21[u] = net.c2(1, k);
            if (!TAU1) {
                t10[u] = net.t1(k, 0);
                t11[u] = net.t1(k, 1);
                t20[u] = net.t2(0, k);
                t21[u] = net.t2(1, k);
            }
        }
        // ... then the arithmetic and the stores
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int k = (V == NH) ? u : k0 + u;
            const NT pk = delta * a2k[u] * ((NT)1 - ha[u] * ha[u]);
            const NT back = (c20[u] * e_c[0] + c21[u] * e_c[1]) * ((NT)1 - hc[u] * hc[u]);
            // actor
            net.a1(k, 0) = a10[u] + ka_s0 * pk;
            net.a1(k, 1) = a11[u] + ka_s1 * pk;
            net.a2(k) = a2k[u] + ka_d * ha[u];
            // critic
            const NT n10 = c10[u] - lc_s0 * back;
            const NT n11 = c11[u] - lc_s1 * back;
            const NT n20 = c20[u] - lc_e0 * hc[u];
            const NT n21 = c21[u] - lc_e1 * hc[u];
            net.c1(k, 0) = n10;
            net.c1(k, 1) = n11;
            net.c2(0, k) = n20;
            net.c2(1, k) = n21;
            // target critic
            if (!TAU1) {
                net.t1(k, 0) = H.tau * n10 + one_m_tau * t10[u];
                net.t1(k, 1) = H.tau * n11 + one_m_tau * t11[u];
                net.t2(0, k) = H.tau * n20 + one_m_tau * t20[u];
                net.t2(1, k) = H.tau * n21 + one_m_tau * t21[u];
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// element storage policies
// ----------------------------------------------------------------------------------------------
template <typename T, int N>
struct LocalVec {           // registers / local memory (compile-time indices after unrolling)
    static constexpr bool kFence = false;
    T v[N];
    H3_DEV T& operator[](int i) { return v[i]; }
    H3_DEV void load(const T* g, int64_t ld) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = g[(int64_t)i * ld];
    }
    H3_DEV void store(T* g, int64_t ld) {
#pragma unroll
        for (int i = 0; i < N; ++i) g[(int64_t)i * ld] = v[i];
    }
};

template <typename T, int N, int STRIDE>
struct SmemVec {            // column of a [N][STRIDE] shared-memory tile; bank-conflict free per warp
    static constexpr bool kFence = true;
    T* p;                   // already offset by the thread index
    H3_DEV T& operator[](int i) { return p[i * STRIDE]; }
    H3_DEV void load(const T* g, int64_t ld) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i * STRIDE] = g[(int64_t)i * ld];
    }
    H3_DEV void store(T* g, int64_t ld) {
#pragma unroll
        for (int i = 0; i < N; ++i) g[(int64_t)i * ld] = p[i * STRIDE];
    }
};

template <typename T, typename Vec>
struct WMat {               // RLS weights (9,7) row-major
    static constexpr bool kFence = Vec::kFence;
    Vec v;
    H3_DEV T& operator()(int i, int j) { return v[i * 7 + j]; }
    H3_DEV T dyn(int i, int j) { return v[i * 7 + j]; }
    H3_DEV T& ref(int i, int j) { return v[i * 7 + j]; }
};

template <typename NT, int NH, typename Vec>
struct NetAcc {             // rows of the `[11*nh][ld]` network block (include/heli3dof.h)
    Vec v;
    H3_DEV NT& a1(int k, int i) { return v[2 * k + i]; }
    H3_DEV NT& a2(int k) { return v[2 * NH + k]; }
    H3_DEV NT& c1(int k, int i) { return v[3 * NH + 2 * k + i]; }
    H3_DEV NT& c2(int j, int k) { return v[5 * NH + j * NH + k]; }
    H3_DEV NT& t1(int k, int i) { return v[7 * NH + 2 * k + i]; }
    H3_DEV NT& t2(int j, int k) { return v[9 * NH + j * NH + k]; }
};

}  // namespace h3
