/* TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may build, load or call this file.
 *
 * Plain-C, scalar, one-episode-at-a-time restatement of the reference's 6-DOF closed-loop tick (SURVEY.md
 * section 8(f) row 4), written literally (same operator order, atan2, pow, LU solves, separate cov.X and
 * X^T.cov, per-element divisions):
 *
 *     Helicopter6DOF (model, RK4, trim, engine)   /root/reference/heli_models.py:347-826
 *     CollectivePID6DOF / LatPedPID               /root/reference/PID.py:27-91
 *     RefGenerator / FirstOrderLag / envelope     /root/reference/util.py:109-222
 *     RecursiveLeastSquares.update (n=17, m=15)   /root/reference/model.py:27-59
 *     DHPAgent (closed-form update, sigmoid out)  /root/reference/agents.py:52-63, :117-196
 *     episode loop, modes train/test_1/test_2     /root/reference/train.py:109-224; params.py
 *
 * All fp64, hidden width 10 for both agents (params.py).
 * PARITY PIN: checked against the reference-generated fixtures tests/golden/h6_*.npz (made by
 * oracle/make_golden6.py from the unmodified reference classes) in tests/test_oracle6.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -o _build/libheli6dof_ref.so heli6dof_ref.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NS 15
#define NR 17 /* regressors: 15 states + 2 actions (train.py:189 passes actions[:2]) */
#define NH 10
#define PI 3.14159265358979323846

enum { TASK_NONE = 0, TASK_HOVER = 1, TASK_TRAIN_LON = 2, TASK_TRAIN_COL = 3, TASK_VELOCITY = 4, TASK_DESCENT = 5 };
enum { MODE_TRAIN = 0, MODE_TEST1 = 1, MODE_TEST2 = 2 };
enum { FLAG_TMAX = 1, FLAG_NAN = 2, FLAG_ENVELOPE = 4, FLAG_LANDED = 8 };

typedef struct {
    double s[NS], t, P_out, P_avail;
    double ref[3];  /* u_ref (ref[0]), theta_ref (ref[7]), z_ref (ref[11]) */
    double pid[4];  /* hdot_corr, y_req_int, phi_int, psi_int */
    double gen[8];  /* z_ref, t_switch, A, int_error_u, filter t0, x0, setpoint, tau */
    double trim_act[4], phi_trim;
    double lr[4];   /* col actor, col critic, lon actor, lon critic (train.py:120-121 mutates lon's) */
    double ret[2], score_sum;
    double W[NR][NS], cov[NR][NR], xprev[NS], uprev[2];
    double col[11 * NH], lon[11 * NH];
    int32_t task, skip, step, flags, term_step, update_col, excitation_phase, pad;
} H6RefEpisode;

typedef struct {
    double gamma, tau, scale, w_r;
    int32_t tracked, ac_state, chan, act; /* act: 0 tanh, 1 sigmoid */
} H6RefAgent;

typedef struct {
    double dt, t_max;
    double lat_gains[9];                     /* LatPedPID.gains (PID.py:50-52) */
    double col_kp, col_ki, col_kd, col_href; /* CollectivePID6DOF */
    double gen_T, gen_dt;                    /* RefGenerator T, dt */
    double rls_gamma;
    double limit_pitch, limit_roll;
    double P_oeitr;                          /* test_2 engine failure: P_available after step 1000 */
    int32_t run_mode, step_switch, exc_len, score_lo, score_hi, pad;
    H6RefAgent col, lon;
} H6RefCfg;

/* ---- heli_models.py:351-462 (constants) -------------------------------------------------------------- */
static const double g0 = 9.80665, mass = 2200, Ixx = 1433, Iyy = 4973, Izz = 4099, Jxz = 660;
static const double xh = 0.08, yh = 0, zh = 1.48;
static const double omegareq = 44.4, R_mr = 4.912, c_mr = 0.27, Nb = 4, Cla_mr = 6.113, delta0 = 0.0074,
                    delta2 = 38.66, gamma_s = 0.0524, I_beta = 231.7, twist = -0.14, e_hinge = 0.746, m_bl = 27.3,
                    nu2 = 1.248, tau_l0_mr = 0.1, kind = 1.15;
static const double Pe = 313000, etam = 0.95, Kaeo = 0.8643, Koeitr = 1.10;
static const double K_fus = 0.83, CDS = 1.2, n_pd = 4.65;
static const double Cla_hs = 4, alpha0_hs = 0.0698, S_hs = 0.803, x_hs = 4.64;
static const double Cla_fin = 4, S_fin = 0.805, beta0_fin = -0.06116, x_fin = 5.30, z_fin = 0.97;
static const double R_tr = 0.95, c_tr = 0.18, gearing = 5.25, x_tr = 6.08, z_tr = 1.72, Cla_tr = 5.7, Nb_tr = 2,
                    k_1_tr = 1, tau_l0_tr = 0.3, M_tr = 0.7;

static double deg2rad(double x) { return x * (PI / 180.0); }
static double rad2deg(double x) { return x * (180.0 / PI); }
static double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); } /* NaN passes */

static double air_density(double altitude) { /* heli_models.py:818-826 */
    return 1.225 * pow(1 - 0.0065 * altitude / 288.15, -(9.80665 / (287.05 * -0.0065) + 1));
}

/* Dense solve with partial pivoting (stands in for LAPACK gesv / pinv of a regular matrix). */
static void solve_n(int n, double* A, double* b) {
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(A[r * n + c]) > fabs(A[p * n + c])) p = r;
        if (p != c) {
            for (int k = 0; k < n; ++k) { double t = A[c * n + k]; A[c * n + k] = A[p * n + k]; A[p * n + k] = t; }
            double t = b[c]; b[c] = b[p]; b[p] = t;
        }
        for (int r = c + 1; r < n; ++r) {
            const double f = A[r * n + c] / A[c * n + c];
            for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
            b[r] -= f * b[c];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        double acc = b[r];
        for (int k = r + 1; k < n; ++k) acc -= A[r * n + k] * b[k];
        b[r] = acc / A[r * n + r];
    }
}

/* Helicopter6DOF.calculate_state_derivatives (heli_models.py:489-694).  P_out: engine integrator state
 * (updated on EVERY call when not trimming, i.e. four times per RK4 step). */
static void derivs(const double* actions, const double* st, int trimming, double dt, double* P_out, double P_avail,
                   double* dots) {
    const double a1sl = deg2rad(-6), a1su = deg2rad(4), b1sl = deg2rad(10), b1su = deg2rad(-5.5), t0l = deg2rad(2),
                 t0u = deg2rad(18), t0trl = deg2rad(18), t0tru = deg2rad(-6), psipmu = deg2rad(-10);
    const double Wt = mass * g0;
    const double sigma_mr = Nb * c_mr / (PI * R_mr), sigma_tr = Nb_tr * c_tr / (PI * R_tr);
    const double f_tr = 1 - 3 * S_fin / (4 * (R_tr * R_tr) * PI);
    const double F0 = 1.3 * 0.73, Vol_fus = PI / 4 * 6 * 1.3;
    const double Keng = -Pe * etam * Kaeo / omegareq / 0.02;
    const double coll = clip(actions[0], 0, 1.0), lng = clip(actions[1], 0, 1.0), lat = clip(actions[2], 0, 1.0),
                 pedal = clip(actions[3], 0, 1.0);
    const double u = st[0], v = st[1], w = st[2], p = st[3], q = st[4], r = st[5], phi = st[6], theta = st[7],
                 psi = st[8], z = st[11], lambda0_mr = st[12], lambda0_tr = st[13], omega = st[14];

    const double a1s2 = (a1su - a1sl) * lat + a1sl;
    const double b1s2 = (b1su - b1sl) * lng + b1sl;
    const double theta_0 = (t0u - t0l) * coll + t0l;
    const double theta_1s = b1s2 * cos(psipmu) + a1s2 * sin(psipmu);
    const double theta_1c = a1s2 * cos(psipmu) + b1s2 * sin(psipmu);
    const double theta_0_tr = (t0tru - t0trl) * pedal + t0trl;

    const double rho = air_density(-z);
    const double gamma = rho * Cla_mr * c_mr * (R_mr * R_mr * R_mr * R_mr) / I_beta;
    const double alpha_cp = -atan2(w, u) + theta_1s;
    const double V = sqrt(u * u + v * v + w * w);
    const double omega_r = omega * R_mr;
    const double mu_x = V * cos(alpha_cp) / omega_r;
    const double mu_z = -V * sin(alpha_cp) / omega_r;

    const double eps = e_hinge / R_mr;
    const double den1 = gamma * (0.25 * (eps * eps) - eps / 3 + 0.125);
    const double den2 = gamma * ((mu_x * mu_x) * (0.0625 * (eps * eps) - eps * 0.125 + 0.0625));
    double A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, b[3];
    A[1] = gamma * mu_x * (2 * (eps * eps) - eps) / (8 * nu2);
    A[5] = (1 - nu2) / (den1 - den2);
    A[6] = gamma * mu_x * (1.0 / 6 - 0.25 * eps) / (-den1 - den2);
    A[7] = (1 - nu2) / (-den1 - den2);
    b[0] = (gamma / (2 * nu2)) *
           (theta_0 * ((0.25 - eps / 3) + (mu_x * mu_x) * (0.25 * (eps * eps) - 1.0 / 2 * eps + 0.25)) +
            mu_x * (1.0 / 2 * eps - 1.0 / 3) * theta_1s +
            ((mu_x * mu_x) / 6 + 0.2 - (mu_x * mu_x) * eps * 0.25 - 0.25 * eps) * twist -
            (1.0 / 3 - 0.5 * eps) * (lambda0_mr - mu_z) + mu_x * (1.0 / 6 - 0.25 * eps) * p / omega);
    b[1] = (gamma * (mu_x * (2 - 3 * eps) / 6 * theta_0 +
                     ((mu_x * mu_x) * (-3 * (eps * eps) + 6 * eps - 3) + 16.0 / 6 * eps - 2) / 16 * theta_1s +
                     mu_x / 4 * ((1 - 4.0 / 3 * eps) * twist - ((eps * eps) - 2 * eps + 1) * (lambda0_mr - mu_z)) +
                     (1 - 4.0 / 3 * eps) * p / (8 * omega)) -
            2 * q / omega) /
           (den1 - den2);
    b[2] = (gamma * ((8.0 / 3 * eps - 2 + (mu_x * mu_x) * (-(eps * eps) + 2 * eps - 1)) / 16 * theta_1c +
                     (1 - 4.0 / 3 * eps) * q / (8 * omega)) +
            2 * p / omega) /
           (-den1 - den2);
    solve_n(3, A, b);
    const double a1 = b[1], b1 = b[2];
    const double Ct = (Cla_mr * sigma_mr / 8) * (2 * (2.0 / 3 + mu_x * mu_x) * theta_0 + (2 * theta_1s + p / omega) * mu_x +
                                                 2 * (mu_z - lambda0_mr) + (1 + mu_x * mu_x) * twist);
    const double Cd = delta0 + delta2 * (Ct * Ct);
    const double T_mr = Ct * rho * (omega_r * omega_r) * PI * (R_mr * R_mr);
    const double a1t1s = a1 - theta_1s + gamma_s;
    const double b1t1c = b1 + theta_1c;
    const double Le = (omega_r * omega_r) * eps * m_bl * sin(b1t1c);
    const double Me = (omega_r * omega_r) * eps * m_bl * sin(a1t1s);
    const double Ct_bem_mr = Ct;
    const double Ct_gl_mr = 2 * lambda0_mr * sqrt(mu_x * mu_x + (lambda0_mr - mu_z) * (lambda0_mr - mu_z));

    const double omegar_tr = gearing * omega * R_tr;
    const double wq = w + k_1_tr * lambda0_mr * omega_r + q * x_tr;
    const double mu_x_tr = sqrt(u * u + wq * wq) / omegar_tr;
    const double mu_z_tr = -(v - x_tr * r + z_tr * p) / omegar_tr;
    const double Ct_bem_tr = Cla_tr * sigma_tr * (theta_0_tr * (2 + 3 * (mu_x_tr * mu_x_tr)) + 3 * mu_z_tr - 6 * lambda0_tr) / 12;
    const double Ct_gl_tr = 2 * lambda0_tr * sqrt(mu_x_tr * mu_x_tr + (mu_z_tr - lambda0_tr) * (mu_z_tr - lambda0_tr));
    const double T_tr = Ct_bem_tr * rho * (omegar_tr * omegar_tr) * PI * (R_tr * R_tr);
    const double Y_tr = T_tr * f_tr;
    const double L_tr = Y_tr * z_tr;
    const double N_tr = -Y_tr * x_tr;

    const double alpha_fus = atan2(w, u);
    const double R_fus = 0.5 * rho * (V * V) * F0;
    const double X_fus = -R_fus * cos(alpha_fus);
    const double Z_fus = -R_fus * sin(alpha_fus);
    const double M_fus = rho * (V * V) * K_fus * Vol_fus * alpha_fus;

    const double w_hs = w + q * x_hs;
    const double alpha_hs = alpha0_hs + atan2(w_hs, u);
    const double V_hs = sqrt(u * u + w_hs * w_hs);
    const double Z_hs = -rho / 2 * (V_hs * V_hs) * 0.65 * S_hs * Cla_hs * alpha_hs;
    const double M_hs = Z_hs * x_hs;

    const double v_fin = v - r * x_fin + p * z_fin;
    const double beta_fin = beta0_fin + atan2(v_fin, u);
    const double V_fin = sqrt(u * u + v_fin * v_fin);
    const double Y_fin = -rho / 2 * (V_fin * V_fin) * S_fin * Cla_fin * beta_fin;
    const double L_fin = z_fin * Y_fin;
    const double N_fin = -x_fin * Y_fin;

    const double P_par = CDS * rho * (V * V * V) / 2;
    const double P_i = fabs(kind * T_mr * lambda0_mr * omega_r);
    const double PpPd = sigma_mr * Cd * rho * (omega_r * omega_r * omega_r) * PI * (R_mr * R_mr) * (1 + n_pd * (mu_x * mu_x)) / 8;
    const double P_c = -w * Wt;
    const double P_tr = fabs(T_tr / M_tr * sqrt(fabs(T_tr / (2 * rho * PI * (R_tr * R_tr)))));
    const double P_req = P_par + P_i + PpPd + P_c + P_tr;

    const double tau = 0.3;
    double P_eng = Keng * (omega - omegareq * 1.02);
    if (trimming) {
        *P_out = P_eng;
    } else {
        *P_out += (P_eng - *P_out) / tau * dt;
        P_eng = *P_out;
    }
    P_eng = clip(P_eng, 0, P_avail);

    const double Xmr = -T_mr * sin(a1t1s) * cos(b1t1c);
    const double Ymr = T_mr * sin(b1t1c);
    const double Zmr = -T_mr * cos(a1t1s) * cos(b1t1c);
    const double Lmr = Ymr * zh - Zmr * yh + Le;
    const double Mmr = -Xmr * zh - Zmr * xh + Me;
    const double Nmr = P_eng / omega + Xmr * yh - Ymr * xh;

    const double X = Xmr + X_fus;
    const double Y = Ymr + Y_tr + Y_fin;
    const double Z = Zmr + Z_fus + Z_hs;
    const double Fx = -Wt * sin(theta) + X;
    const double Fy = Wt * cos(theta) * sin(phi) + Y;
    const double Fz = Wt * cos(theta) * cos(phi) + Z;
    const double L = Lmr + L_tr + L_fin;
    const double M = Mmr + M_fus + M_hs;
    const double N = Nmr + N_tr + N_fin;

    const double udot = Fx / mass - q * w + r * v;
    const double vdot = Fy / mass - r * u + p * w;
    const double wdot = Fz / mass - p * v + q * u;
    const double rdot = (N - (Iyy - Ixx) * p * q + Jxz * ((L - (Izz - Iyy) * q * r + Jxz * p * q) / Ixx - r * q)) /
                        (Izz - (Jxz * Jxz) / Ixx);
    const double qdot = (M - (Ixx - Izz) * r * p - Jxz * (p * p - r * r)) / Iyy;
    const double pdot = (L - (Izz - Iyy) * q * r + Jxz * (rdot + p * q)) / Ixx;
    const double psidot = (q * sin(phi) + r * cos(phi)) / cos(theta);
    const double thetadot = q * cos(phi) - r * sin(phi);
    const double phidot = p + psidot * sin(theta);
    const double Omegadot = rdot + (P_eng - P_req) / omega / (1.1 * Nb * I_beta);
    const double xdot = (u * cos(theta) + (v * sin(phi) + w * cos(phi)) * sin(theta)) * cos(psi) -
                        (v * cos(phi) - w * sin(phi)) * sin(psi);
    const double ydot = (u * cos(theta) + (v * sin(phi) + w * cos(phi)) * sin(theta)) * sin(psi) +
                        (v * cos(phi) - w * sin(phi)) * cos(psi);
    const double zdot = -u * sin(theta) + (v * sin(phi) + w * cos(phi)) * cos(theta);
    const double l0mr_dot = (Ct_bem_mr - Ct_gl_mr) / tau_l0_mr;
    const double l0tr_dot = (Ct_bem_tr - Ct_gl_tr) / tau_l0_tr;
    const double out[NS] = {udot, vdot, wdot, pdot, qdot, rdot, phidot, thetadot, psidot, xdot, ydot, zdot,
                            l0mr_dot, l0tr_dot, Omegadot};
    memcpy(dots, out, sizeof(out));
}

/* integrate_runge_kutta (heli_models.py:696-721) */
static void rk4(const double* actions, const double* old, double dt, double* P_out, double P_avail, double* nw) {
    double f[NS], k1[NS], k2[NS], k3[NS], k4[NS], st[NS];
    derivs(actions, old, 0, dt, P_out, P_avail, f);
    for (int i = 0; i < NS; ++i) { k1[i] = dt * f[i]; st[i] = old[i] + k1[i] / 2; }
    derivs(actions, st, 0, dt, P_out, P_avail, f);
    for (int i = 0; i < NS; ++i) { k2[i] = dt * f[i]; st[i] = old[i] + k2[i] / 2; }
    derivs(actions, st, 0, dt, P_out, P_avail, f);
    for (int i = 0; i < NS; ++i) { k3[i] = dt * f[i]; st[i] = old[i] + k3[i]; }
    derivs(actions, st, 0, dt, P_out, P_avail, f);
    for (int i = 0; i < NS; ++i) { k4[i] = dt * f[i]; nw[i] = old[i] + (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6; }
}

/* Helicopter6DOF.trim (heli_models.py:723-816): Newton on 12 trim variables, forward-difference Jacobian
 * with delta = 1e-11, until max|f| <= 1e-8.  Returns the iteration count (negative: not converged). */
static void trim_to_state(double* st, double* act, const double* tv) {
    st[0] = tv[0]; st[1] = tv[1]; st[2] = tv[2]; st[7] = tv[3]; st[6] = tv[4]; st[12] = tv[5]; st[13] = tv[6]; st[14] = tv[7];
    act[0] = tv[8]; act[1] = tv[9]; act[2] = tv[10]; act[3] = tv[11];
}
static void trim_residual(const double* st, const double* act, double V, double fpa, double P_avail, double* P_out, double* f) {
    double dot[NS];
    derivs(act, st, 1, 0.0, P_out, P_avail, dot);
    for (int i = 0; i < 6; ++i) f[i] = dot[i];
    for (int i = 0; i < 6; ++i) f[6 + i] = dot[9 + i];
    f[6] -= V * cos(fpa);
    f[8] += V * sin(fpa);
}
int32_t h6ref_trim(double trim_speed, double flight_path_angle_deg, double altitude, double P_avail, double* state_out,
                   double* controls_out, double* P_out_out) {
    const double fpa = deg2rad(flight_path_angle_deg), V = trim_speed;
    const double rho = air_density(altitude), F0 = 1.3 * 0.73, Wt = mass * g0;
    const double D = rho / 2 * (V * V) * F0;
    const double theta = asin(-D * cos(fpa) / Wt);
    double st[NS] = {V * cos(theta), 0, V * sin(theta), 0, 0, 0, 0, theta, 0, 0, 0, -altitude, 0.05, 0.05, omegareq};
    double tv[12] = {st[0], st[1], st[2], st[7], st[6], st[12], st[13], st[14], 0.75, 0.50, 0.5, 0.50};
    double act[4], f[12], fn[12], J[144], P_out = 0;
    const double delta = 1e-11;
    int nn = 0;
    for (;;) {
        if (nn > 200) return -nn;
        nn += 1;
        trim_to_state(st, act, tv);
        trim_residual(st, act, V, fpa, P_avail, &P_out, f);
        double mx = 0;
        for (int i = 0; i < 12; ++i) mx = fmax(mx, fabs(f[i]));
        for (int i = 0; i < 12; ++i) {
            double tp[12];
            memcpy(tp, tv, sizeof(tp));
            tp[i] = tv[i] + delta;
            trim_to_state(st, act, tp);
            trim_residual(st, act, V, fpa, P_avail, &P_out, fn);
            for (int r = 0; r < 12; ++r) J[r * 12 + i] = (fn[r] - f[r]) / delta;
        }
        double inc[12];
        for (int i = 0; i < 12; ++i) inc[i] = -f[i];
        solve_n(12, J, inc);
        /* heli_models.py:795-806: `trimvar = oldtrimvar + perturb` rebinds the name inside the loop, so the
         * increment is added to the LAST perturbed vector (pedal + delta), not to oldtrimvar. */
        tv[11] += delta;
        for (int i = 0; i < 12; ++i) tv[i] += inc[i];
        if (!(mx > 1e-8)) break; /* `while max(|f|) > 1e-8` sees this pass's f at the top of the next pass */
    }
    /* heli_models.py:808: trim_to_state(states, trimvar) with the LAST perturbed `states` array: every slot it does
     * not overwrite keeps its initial value, the overwritten ones come from the final trimvar. */
    trim_to_state(st, act, tv);
    memcpy(state_out, st, sizeof(st));
    memcpy(controls_out, act, sizeof(act));
    *P_out_out = P_out;
    return nn;
}

/* CollectivePID6DOF.__call__ (PID.py:31-45) */
static double pid_collective(const H6RefCfg* c, const H6RefEpisode* e, const double* obs, double* corr) {
    const double u = obs[0], v = obs[1], w = obs[2], phi = obs[6], theta = obs[7], z = obs[11], h = -z;
    const double hdot_ref = c->col_kd * (c->col_href - h);
    const double hdot = u * sin(theta) - (v * sin(phi) + w * cos(phi)) * cos(theta);
    const double err = hdot_ref - hdot;
    const double collective = e->trim_act[0] + c->col_kp * err + c->col_ki * *corr;
    *corr += c->dt * err;
    return clip(collective, 0, 1);
}

/* LatPedPID.__call__ (PID.py:62-91) */
static void pid_latped(const H6RefCfg* c, H6RefEpisode* e, const double* obs, double* lat_out, double* ped_out) {
    const double* K = c->lat_gains;
    const double u = obs[0], v = obs[1], w = obs[2], p = obs[3], r = obs[5], phi = obs[6], theta = obs[7], psi = obs[8],
                 y = obs[10];
    const double ydot = (u * cos(theta) + (v * sin(phi) + w * cos(phi)) * sin(theta)) * sin(psi) +
                        (v * cos(phi) - w * sin(phi)) * cos(psi);
    const double y_error = 0 - y;
    const double phi_req = e->phi_trim + K[0] * y_error + K[1] * e->pid[1] + K[2] * ydot;
    const double phi_error = phi_req - phi, psi_error = 0 - psi;
    const double lat = clip(e->trim_act[2] + (K[3] * phi_error + K[4] * e->pid[2] + K[5] * p) / 100, 0, 1);
    const double ped = clip(e->trim_act[3] + (K[6] * psi_error + K[7] * e->pid[3] + K[8] * r) / 100, 0, 1);
    if (0.01 < lat && lat < 0.99) { e->pid[1] += y_error * c->dt; e->pid[2] += phi_error * c->dt; }
    if (0.01 < ped && ped < 0.99) e->pid[3] += psi_error * c->dt;
    *lat_out = lat;
    *ped_out = ped;
}

/* FirstOrderLag.get_result (util.py:212-213) */
static double lag(const double* gen, double t) { return gen[5] + gen[6] * (1 - exp(-1 * (t - gen[4]) / gen[7])); }
static long r4(double t) { return lround(t * 10000.0); } /* round(t, 4) compared as an integer count of 1e-4 s */

/* RefGenerator.get_ref (util.py:122-166) -> ref[0] (u_ref or NaN), ref[7], ref[11] */
static void gen_get_ref(const H6RefCfg* c, H6RefEpisode* e, const double* obs, double t, double* ref) {
    double* g = e->gen;
    double theta_ref = 0, z_ref = 0, u_ref = NAN;
    if (e->task == TASK_HOVER) {
        theta_ref = 0; z_ref = g[0];
    } else if (e->task == TASK_TRAIN_LON) {
        theta_ref = deg2rad(g[2] * sin(2 * PI * t / c->gen_T)); z_ref = 0;
    } else if (e->task == TASK_TRAIN_COL) {
        theta_ref = obs[0] > 1 ? deg2rad(3) : deg2rad(2.5) * obs[0];
        z_ref = fmax(g[0] - 1 * (t - g[1]), -25);
    } else if (e->task == TASK_VELOCITY) {
        z_ref = g[0];
        u_ref = lag(g, t);
        const double u_error = u_ref - obs[0];
        const double ku = r4(t) < 200000 ? -2 : -3;
        theta_ref = deg2rad(ku * u_error + -0.05 * g[3]);
        g[3] += u_error * c->gen_dt;
    } else if (e->task == TASK_DESCENT) {
        if (r4(t) == 129300) { g[4] = t; g[5] = obs[0]; g[6] = -obs[0]; g[7] = 2.5; }
        if (r4(t) <= 129200) {
            theta_ref = deg2rad(-0.1);
            z_ref = g[0] + 1.93 * (t - g[1]);
        } else {
            u_ref = lag(g, t);
            const double u_error = u_ref - obs[0];
            theta_ref = deg2rad(-3.5 * u_error + 0.1 * g[3]);
            g[3] += u_error * c->gen_dt;
            z_ref = -5 + 0.5 * (t - 12.92);
        }
    }
    ref[0] = u_ref; ref[1] = theta_ref; ref[2] = z_ref;
}

/* RefGenerator.set_task (util.py:168-181) */
static void gen_set_task(H6RefEpisode* e, int task, double t, const double* obs, double z_start, double filter_target,
                         double t_switch) {
    e->task = task;
    e->gen[3] = 0;
    if (task == TASK_TRAIN_COL) e->gen[0] = z_start;
    else if (task == TASK_HOVER) e->gen[0] = obs[11];
    else if (task == TASK_VELOCITY) { e->gen[0] = z_start; e->gen[4] = t; e->gen[5] = obs[0]; e->gen[6] = filter_target; }
    else if (task == TASK_DESCENT) { e->gen[1] = t_switch; e->gen[0] = obs[11]; }
}

/* RecursiveLeastSquares.update (model.py:27-59), n = 17, m = 15 */
static void rls_update(const H6RefCfg* c, H6RefEpisode* e, const double* st, const double* act, const double* nxt) {
    if (e->skip) {
        for (int i = 0; i < NS; ++i) e->xprev[i] = st[i];
        e->uprev[0] = act[0]; e->uprev[1] = act[1];
        e->skip = 0;
        return;
    }
    double X[NR], Y[NS], err[NS], covX[NR], Xcov[NR];
    for (int i = 0; i < NS; ++i) X[i] = st[i] - e->xprev[i];
    X[NS] = act[0] - e->uprev[0]; X[NS + 1] = act[1] - e->uprev[1];
    for (int j = 0; j < NS; ++j) Y[j] = nxt[j] - st[j];
    for (int i = 0; i < NS; ++i) e->xprev[i] = st[i];
    e->uprev[0] = act[0]; e->uprev[1] = act[1];
    for (int j = 0; j < NS; ++j) {
        double yh = 0;
        for (int i = 0; i < NR; ++i) yh += X[i] * e->W[i][j];
        err[j] = Y[j] - yh;
    }
    for (int i = 0; i < NR; ++i) {
        double a = 0, b = 0;
        for (int k = 0; k < NR; ++k) { a += e->cov[i][k] * X[k]; b += X[k] * e->cov[k][i]; }
        covX[i] = a; Xcov[i] = b;
    }
    double d = c->rls_gamma;
    for (int i = 0; i < NR; ++i) d += Xcov[i] * X[i];
    for (int i = 0; i < NR; ++i)
        for (int j = 0; j < NS; ++j) e->W[i][j] = e->W[i][j] + (covX[i] * err[j]) / d;
    for (int i = 0; i < NR; ++i)
        for (int j = 0; j < NR; ++j) e->cov[i][j] = (e->cov[i][j] - (covX[i] * Xcov[j]) / d) / c->rls_gamma;
}

#define A1(n, k, i) (n)[2 * (k) + (i)]
#define A2(n, k) (n)[2 * NH + (k)]
#define C1(n, k, i) (n)[3 * NH + 2 * (k) + (i)]
#define C2(n, j, k) (n)[5 * NH + (j) * NH + (k)]
#define T1(n, k, i) (n)[7 * NH + 2 * (k) + (i)]
#define T2(n, j, k) (n)[9 * NH + (j) * NH + (k)]

static double final_act(int act, double y) { return act ? 1.0 / (1.0 + exp(-y)) : tanh(y); }

/* ref vector slot lookup: only slots 0, 7, 11 are ever non-NaN */
static double ref_slot(const double* ref3, int tracked) {
    return tracked == 0 ? ref3[0] : (tracked == 7 ? ref3[1] : (tracked == 11 ? ref3[2] : NAN));
}

static double actor(const double* n, const H6RefAgent* a, double s0, double s1) { /* agents.py:59-63 */
    double y = 0;
    for (int k = 0; k < NH; ++k) y += A2(n, k) * tanh(A1(n, k, 0) * s0 + A1(n, k, 1) * s1);
    return a->scale * final_act(a->act, y);
}

/* get_reward + get_transition_matrices + update_networks (agents.py:123-196), closed form */
static double agent_update(double* n, const H6RefAgent* a, double lr_a, double lr_c, const H6RefEpisode* e,
                           const double* st, const double* nxt, const double* ref3, const double* nref3) {
    const int tr = a->tracked, ac = a->ac_state;
    const double r_now = ref_slot(ref3, tr), r_next = ref_slot(nref3, tr);
    const double s0 = st[ac], s1 = st[tr] - r_now, n0 = nxt[ac], n1 = nxt[tr] - r_next;
    const double et = nxt[tr] - r_now;
    const double reward = -(et * et) * a->w_r, dr = -2 * et * a->w_r;
    const int idx[2] = {ac, tr};
    double F[2][2], Gm[2];
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j) F[i][j] = e->W[idx[j]][idx[i]];
        Gm[i] = e->W[NS + a->chan][idx[i]];
    }
    double ha[NH], hc[NH], y = 0, lam1[2] = {0, 0}, lam2[2] = {0, 0};
    for (int k = 0; k < NH; ++k) {
        ha[k] = tanh(A1(n, k, 0) * s0 + A1(n, k, 1) * s1);
        y += A2(n, k) * ha[k];
        hc[k] = tanh(C1(n, k, 0) * s0 + C1(n, k, 1) * s1);
        const double ht = tanh(T1(n, k, 0) * n0 + T1(n, k, 1) * n1);
        for (int j = 0; j < 2; ++j) { lam1[j] += C2(n, j, k) * hc[k]; lam2[j] += T2(n, j, k) * ht; }
    }
    const double fy = final_act(a->act, y);
    const double delta = a->scale * (a->act ? fy * (1 - fy) : (1 - fy * fy));
    double p[NH], da_ds[2] = {0, 0};
    for (int k = 0; k < NH; ++k) {
        p[k] = delta * A2(n, k) * (1 - ha[k] * ha[k]);
        da_ds[0] += A1(n, k, 0) * p[k];
        da_ds[1] += A1(n, k, 1) * p[k];
    }
    const double tg[2] = {a->gamma * lam2[0], dr + a->gamma * lam2[1]};
    double ec[2];
    for (int j = 0; j < 2; ++j)
        ec[j] = lam1[j] - (tg[0] * (F[0][j] + Gm[0] * da_ds[j]) + tg[1] * (F[1][j] + Gm[1] * da_ds[j]));
    const double tG = tg[0] * Gm[0] + tg[1] * Gm[1];
    for (int k = 0; k < NH; ++k) {
        const double back = (C2(n, 0, k) * ec[0] + C2(n, 1, k) * ec[1]) * (1 - hc[k] * hc[k]);
        A1(n, k, 0) -= (p[k] * s0) * (-tG) * lr_a;
        A1(n, k, 1) -= (p[k] * s1) * (-tG) * lr_a;
        A2(n, k) -= (delta * ha[k]) * (-tG) * lr_a;
        C1(n, k, 0) -= (back * s0) * lr_c;
        C1(n, k, 1) -= (back * s1) * lr_c;
        C2(n, 0, k) -= (ec[0] * hc[k]) * lr_c;
        C2(n, 1, k) -= (ec[1] * hc[k]) * lr_c;
        T1(n, k, 0) = a->tau * C1(n, k, 0) + (1.0 - a->tau) * T1(n, k, 0);
        T1(n, k, 1) = a->tau * C1(n, k, 1) + (1.0 - a->tau) * T1(n, k, 1);
        T2(n, 0, k) = a->tau * C2(n, 0, k) + (1.0 - a->tau) * T2(n, 0, k);
        T2(n, 1, k) = a->tau * C2(n, 1, k) + (1.0 - a->tau) * T2(n, 1, k);
    }
    return reward;
}

static double aug1(const double* st, const double* ref3, const H6RefAgent* a) { return st[a->tracked] - ref_slot(ref3, a->tracked); }

/* one iteration of the loop of train.py:109-224 */
static void tick(const H6RefCfg* c, H6RefEpisode* e, const double* excitation, double* actions_out, double* rewards_out) {
    const int k = e->step;
    double obs[NS], act[4], lat, ped;
    memcpy(obs, e->s, sizeof(obs));
    pid_latped(c, e, obs, &lat, &ped);                                                    /* :110 */
    int both = 1;
    if (c->run_mode == MODE_TRAIN) {                                                      /* :113-137 */
        if (k == 0) {
            gen_set_task(e, TASK_TRAIN_LON, 0, obs, 0, 0, 0);
            gen_get_ref(c, e, obs, e->t, e->ref);
        } else if (k == 1000) {
            e->excitation_phase = 0;
        } else if (k == c->step_switch) {
            e->lr[2] *= 0.1; e->lr[3] *= 0.1;
            e->update_col = 1;
            gen_set_task(e, TASK_TRAIN_COL, e->t, obs, obs[11], 0, 0);
        }
        both = !(k < c->step_switch);
    } else if (c->run_mode == MODE_TEST1) {                                               /* :138-160 */
        if (k == 0) { gen_set_task(e, TASK_HOVER, 0, obs, 0, 0, 0); gen_get_ref(c, e, obs, e->t, e->ref); }
        else if (k == 500) gen_set_task(e, TASK_VELOCITY, e->t, obs, 0, 25 - obs[0], 0);
        else if (k == 2000) gen_set_task(e, TASK_VELOCITY, e->t, obs, 0, 0 - obs[0], 0);
    } else {                                                                              /* :161-170 */
        if (k == 0) { gen_set_task(e, TASK_DESCENT, 0, obs, 0, 0, 0); gen_get_ref(c, e, obs, e->t, e->ref); }
        else if (k == 1000) e->P_avail = c->P_oeitr;
    }
    double ref3[3] = {e->ref[0], e->ref[1], e->ref[2]};
    act[1] = e->trim_act[1] - 0.5 + actor(e->lon, &c->lon, obs[c->lon.ac_state], aug1(obs, ref3, &c->lon));
    if (both) act[0] = e->trim_act[0] - 0.5 + actor(e->col, &c->col, obs[c->col.ac_state], aug1(obs, ref3, &c->col));
    else act[0] = pid_collective(c, e, obs, &e->pid[0]);
    act[2] = lat; act[3] = ped;
    if (e->excitation_phase && excitation && k < c->exc_len)                              /* :173-174 */
        for (int i = 0; i < 4; ++i) act[i] += excitation[4 * k + i];
    for (int i = 0; i < 4; ++i) act[i] = clip(act[i], 0, 1);                              /* :176 */
    double nxt[NS];
    rk4(act, obs, c->dt, &e->P_out, e->P_avail, nxt);                                     /* :179, heli_models.py:464-475 */
    int done = e->t > c->t_max;
    e->t += c->dt;
    memcpy(e->s, nxt, sizeof(nxt));
    e->gen[2] = e->t < 20 ? 10 : (e->t < 40 ? 15 : 20);                                   /* :180-185 */
    double nref3[3];
    gen_get_ref(c, e, obs, e->t, nref3);                                                  /* :186 (pre-step obs!) */
    rls_update(c, e, obs, act, nxt);                                                      /* :189 */
    double r_col = 0, r_lon;
    if (e->update_col) r_col = agent_update(e->col, &c->col, e->lr[0], e->lr[1], e, obs, nxt, ref3, nref3);
    r_lon = agent_update(e->lon, &c->lon, e->lr[2], e->lr[3], e, obs, nxt, ref3, nref3);
    e->ret[0] += r_col; e->ret[1] += r_lon;
    if (k >= c->score_lo && k < c->score_hi) e->score_sum += r_lon;                       /* :256 */
    int fl = done ? FLAG_TMAX : 0;
    if (fabs(rad2deg(obs[7])) > c->limit_pitch || fabs(rad2deg(obs[6])) > c->limit_roll) fl |= FLAG_ENVELOPE; /* :207 */
    if (act[0] != act[0] || act[1] != act[1] || act[2] != act[2] || act[3] != act[3]) fl |= FLAG_NAN;      /* :213 */
    if (c->run_mode == MODE_TEST2 && obs[11] > 0) fl |= FLAG_LANDED;                       /* :218 */
    e->ref[0] = nref3[0]; e->ref[1] = nref3[1]; e->ref[2] = nref3[2];
    e->step = k + 1;
    if (fl) { e->flags = fl; e->term_step = k; }
    if (actions_out) memcpy(actions_out, act, sizeof(act));
    if (rewards_out) { rewards_out[0] = r_col; rewards_out[1] = r_lon; }
}

int64_t h6ref_run(const H6RefCfg* cfg, H6RefEpisode* eps, int64_t n, int32_t n_ticks, const double* excitation,
                  int32_t threads, double* last_actions, double* last_rewards) {
    int64_t total = 0;
#pragma omp parallel for num_threads(threads) reduction(+ : total) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        H6RefEpisode* e = eps + i;
        for (int j = 0; j < n_ticks && e->flags == 0; ++j) {
            tick(cfg, e, excitation, last_actions ? last_actions + 4 * i : 0, last_rewards ? last_rewards + 2 * i : 0);
            total += 1;
        }
    }
    return total;
}

/* unit entry points for the fixtures */
void h6ref_derivs(const double* actions, const double* state, int32_t trimming, double dt, double* P_out, double P_avail,
                  double* dots) {
    derivs(actions, state, trimming, dt, P_out, P_avail, dots);
}
void h6ref_step(const double* actions, const double* state, double dt, double* P_out, double P_avail, double* next) {
    rk4(actions, state, dt, P_out, P_avail, next);
}
void h6ref_set_task(H6RefEpisode* e, int32_t task, double t, double z_start, double filter_target, double t_switch) {
    gen_set_task(e, task, t, e->s, z_start, filter_target, t_switch);
}
int64_t h6ref_sizeof_episode(void) { return (int64_t)sizeof(H6RefEpisode); }
int64_t h6ref_sizeof_cfg(void) { return (int64_t)sizeof(H6RefCfg); }
