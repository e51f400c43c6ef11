/* TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may build, load or call this file.
 *
 * Plain-C, scalar, one-episode-at-a-time restatement of the reference's 3-DOF closed-loop tick, written
 * literally (same operator order, atan2, separate cov.X and X^T.cov, per-element divisions):
 *
 *     Helicopter3DOF.step / get_ref     /root/reference/heli_models.py:142-255
 *     CollectivePID3DOF.__call__        /root/reference/PID.py:16-24
 *     RecursiveLeastSquares.update      /root/reference/model.py:27-59
 *     DHPAgent (closed-form update)     /root/reference/agents.py:117-196
 *     tick order / phase switch / break /root/reference/main_3dof.py:122-197
 *
 * All fp64 ("all64" mode), hidden widths fixed at 6 (cyclic) and 10 (collective) as in main_3dof.py:20-46.
 * PARITY PIN: checked against the reference-generated fixtures tests/golden/ticks_all64.npz and
 * seed2_all64.npz (tests/test_oracle_c.py) and against the NumPy oracle.  It exists because a compiled CPU
 * loop is the honest CPU baseline for the benchmark: the NumPy oracle and the reference itself are
 * dominated by interpreter / dispatcher overhead.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o _build/libheli3dof_ref.so heli3dof_ref.c -lm   (oracle/build_c.py)
 */
#include <math.h>
#include <stdint.h>

#define NHL 6
#define NHC 10
#define PI 3.14159265358979323846

typedef struct {
    double s[7], t, ref_th, corr_u, pid_corr, theta_trim;
    double trim[2], ret[2];
    double W[9][7], cov[9][9], xprev[7], uprev[2];
    double lon[11 * NHL], col[11 * NHC]; /* include/heli3dof.h network block layout */
    int32_t skip, step, flags, term_step;
} H3RefEpisode;

typedef struct {
    double gamma, lr_a, lr_c, tau, scale, w_r;
    int32_t tracked, ac_state, chan, pad;
} H3RefAgent;

typedef struct {
    double dt;
    double env_kp, env_ki, env_kd, env_ref, ref_amp, ref_period;
    double pid_kp, pid_ki, pid_kd, pid_href;
    double rls_gamma;
    int32_t task, switch_step, exc_len, pad;
    H3RefAgent col, lon;
} H3RefCfg;

/* heli_models.py:117-130 */
static const double G = 9.81, CL_ALPHA = 5.7, SIGMA = 0.075, LOCK = 6, CDS = 1.5, MASS = 2200, RHO = 1.225,
                    V_TIP = 200, R_ROTOR = 7.32, IY = 10615, MAST = 1, TAU = 0.1;

static double deg2rad(double x) { return x * (PI / 180.0); }

static double get_ref(const H3RefCfg* c, H3RefEpisode* e, double t_now) { /* heli_models.py:217-255 */
    const double t = t_now + c->dt;
    if (c->task == 0) {
        const double x = PI * t / c->ref_period;
        return deg2rad(c->ref_amp / 1.76 * (sin(x) + sin(2 * x)));
    }
    if (c->task == 1) {
        const double u_err = c->env_ref - e->s[2];
        const double th = deg2rad(c->env_kp * u_err + c->env_ki * e->corr_u);
        e->corr_u += u_err * c->dt;
        return th;
    }
    {
        const double x_err = c->env_ref - e->s[0];
        const double th = deg2rad(c->env_kp * x_err + c->env_ki * 0 + c->env_kd * e->s[2]);
        return (e->theta_trim > th) ? e->theta_trim : th;
    }
}

static int env_step(const H3RefCfg* c, double* s, double collective, double cyclic) { /* heli_models.py:157-213 */
    const double omega = V_TIP / R_ROTOR, area = PI * R_ROTOR * R_ROTOR;
    const double x = s[0], z = s[1], u = s[2], w = s[3], pitch = s[4], q = s[5], li = s[6];
    const double q_diml = q / omega;
    const double v_diml = sqrt(u * u + w * w) / V_TIP;
    const double alpha_c = cyclic - atan2(w, u);
    const double mu = v_diml * cos(alpha_c);
    const double lambda_c = v_diml * sin(alpha_c);
    const double a1 =
        (-16 * q_diml / LOCK + (8.0 / 3.0) * mu * collective - 2 * mu * (lambda_c + li)) / (1 - mu * mu / 2);
    const double ct_bem =
        (CL_ALPHA * SIGMA / 4) * ((2.0 / 3.0) * collective * (1 + (3.0 / 2.0) * mu * mu) - (lambda_c + li));
    const double gx = v_diml * cos(alpha_c - a1), gy = v_diml * sin(alpha_c - a1) + li;
    const double ct_glau = 2 * li * sqrt(gx * gx + gy * gy);
    const double thrust = ct_bem * RHO * V_TIP * V_TIP * area;
    const double vv = v_diml * V_TIP;
    const double x_dot = u * cos(pitch) + w * sin(pitch);
    const double z_dot = -(u * sin(pitch) - w * cos(pitch));
    const double u_dot = (-G * sin(pitch) - CDS * .5 * RHO * vv / MASS * u + thrust / MASS * sin(cyclic - a1) - q * w);
    const double w_dot = (G * cos(pitch) - CDS * .5 * RHO * vv / MASS * w - thrust / MASS * cos(cyclic - a1) + q * u);
    const double q_dot = -thrust * MAST / IY * sin(cyclic - a1);
    const double li_dot = (ct_bem - ct_glau) / TAU;
    s[0] = x + x_dot * c->dt;
    s[1] = z + z_dot * c->dt;
    s[2] = u + u_dot * c->dt;
    s[3] = w + w_dot * c->dt;
    s[4] = pitch + q * c->dt;
    s[5] = q + q_dot * c->dt;
    s[6] = li + li_dot * c->dt;
    return fabs(s[4] * (180.0 / PI)) > 90;
}

static double pid(const H3RefCfg* c, const double* s, double* corr) { /* PID.py:16-24 */
    const double hdot_ref = c->pid_kd * (c->pid_href - -s[1]);
    const double hdot = s[2] * sin(s[4]) - s[3] * cos(s[4]);
    const double err = hdot_ref - hdot;
    double col = deg2rad(5 + c->pid_kp * err + c->pid_ki * *corr);
    *corr += c->dt * err;
    const double hi = deg2rad(10);
    if (col < 0) col = 0; else if (col > hi) col = hi; /* NaN passes like np.clip */
    return col;
}

static void rls_update(const H3RefCfg* c, H3RefEpisode* e, const double* st, const double* act, const double* nxt) {
    /* model.py:27-59 */
    if (e->skip) {
        e->skip = 0;
    } else {
        double X[9], Y[7], err[7], covX[9], Xcov[9], d = c->rls_gamma;
        for (int i = 0; i < 7; ++i) { X[i] = st[i] - e->xprev[i]; Y[i] = nxt[i] - st[i]; }
        X[7] = act[0] - e->uprev[0];
        X[8] = act[1] - e->uprev[1];
        for (int j = 0; j < 7; ++j) {
            double yh = 0;
            for (int i = 0; i < 9; ++i) yh += X[i] * e->W[i][j];
            err[j] = Y[j] - yh;
        }
        for (int i = 0; i < 9; ++i) {
            double a = 0, b = 0;
            for (int j = 0; j < 9; ++j) { a += e->cov[i][j] * X[j]; b += X[j] * e->cov[j][i]; }
            covX[i] = a;
            Xcov[i] = b;
        }
        for (int j = 0; j < 9; ++j) d += Xcov[j] * X[j];
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 7; ++j) e->W[i][j] = e->W[i][j] + covX[i] * err[j] / d;
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) e->cov[i][j] = (e->cov[i][j] - covX[i] * Xcov[j] / d) / c->rls_gamma;
    }
    for (int i = 0; i < 7; ++i) e->xprev[i] = st[i];
    e->uprev[0] = act[0];
    e->uprev[1] = act[1];
}

/* network block accessors (rows of the [11*nh] block) */
#define A1(n, nh, k, i) (n)[2 * (k) + (i)]
#define A2(n, nh, k) (n)[2 * (nh) + (k)]
#define C1(n, nh, k, i) (n)[3 * (nh) + 2 * (k) + (i)]
#define C2(n, nh, j, k) (n)[5 * (nh) + (j) * (nh) + (k)]
#define T1(n, nh, k, i) (n)[7 * (nh) + 2 * (k) + (i)]
#define T2(n, nh, j, k) (n)[9 * (nh) + (j) * (nh) + (k)]

static double actor(const double* n, int nh, const H3RefAgent* a, double s0, double s1) { /* agents.py:59-63 */
    double y = 0;
    for (int k = 0; k < nh; ++k) y += A2(n, nh, k) * tanh(A1(n, nh, k, 0) * s0 + A1(n, nh, k, 1) * s1);
    return a->scale * tanh(y);
}

static double agent_update(double* n, int nh, const H3RefAgent* a, const H3RefEpisode* e, const double* st,
                           const double* nxt, double ref_th, double nref_th) { /* agents.py:123-196 */
    const int tr = a->tracked, ac = a->ac_state;
    const double r_now = (tr == 1) ? -0.0 : ref_th, r_next = (tr == 1) ? -0.0 : nref_th;
    const double s0 = st[ac], s1 = st[tr] - r_now, n0 = nxt[ac], n1 = nxt[tr] - r_next;
    const double et = nxt[tr] - r_now;
    const double reward = -(et * et) * a->w_r, dr = -2 * et * a->w_r;
    const int idx[2] = {ac, tr};
    double F[2][2], Gm[2];
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j) F[i][j] = e->W[idx[j]][idx[i]];
        Gm[i] = e->W[7 + a->chan][idx[i]];
    }
    double ha[NHC], hc[NHC], y = 0, lam1[2] = {0, 0}, lam2[2] = {0, 0};
    for (int k = 0; k < nh; ++k) {
        ha[k] = tanh(A1(n, nh, k, 0) * s0 + A1(n, nh, k, 1) * s1);
        y += A2(n, nh, k) * ha[k];
        hc[k] = tanh(C1(n, nh, k, 0) * s0 + C1(n, nh, k, 1) * s1);
        const double ht = tanh(T1(n, nh, k, 0) * n0 + T1(n, nh, k, 1) * n1);
        for (int j = 0; j < 2; ++j) { lam1[j] += C2(n, nh, j, k) * hc[k]; lam2[j] += T2(n, nh, j, k) * ht; }
    }
    const double ty = tanh(y), delta = a->scale * (1 - ty * ty);
    double p[NHC], da_ds[2] = {0, 0};
    for (int k = 0; k < nh; ++k) {
        p[k] = delta * A2(n, nh, k) * (1 - ha[k] * ha[k]);
        da_ds[0] += A1(n, nh, k, 0) * p[k];
        da_ds[1] += A1(n, nh, k, 1) * p[k];
    }
    const double tg[2] = {a->gamma * lam2[0], dr + a->gamma * lam2[1]};
    double ec[2];
    for (int j = 0; j < 2; ++j)
        ec[j] = lam1[j] - (tg[0] * (F[0][j] + Gm[0] * da_ds[j]) + tg[1] * (F[1][j] + Gm[1] * da_ds[j]));
    const double tG = tg[0] * Gm[0] + tg[1] * Gm[1];
    for (int k = 0; k < nh; ++k) {
        const double back = (C2(n, nh, 0, k) * ec[0] + C2(n, nh, 1, k) * ec[1]) * (1 - hc[k] * hc[k]);
        A1(n, nh, k, 0) -= (p[k] * s0) * (-tG) * a->lr_a;
        A1(n, nh, k, 1) -= (p[k] * s1) * (-tG) * a->lr_a;
        A2(n, nh, k) -= (delta * ha[k]) * (-tG) * a->lr_a;
        C1(n, nh, k, 0) -= (back * s0) * a->lr_c;
        C1(n, nh, k, 1) -= (back * s1) * a->lr_c;
        C2(n, nh, 0, k) -= (ec[0] * hc[k]) * a->lr_c;
        C2(n, nh, 1, k) -= (ec[1] * hc[k]) * a->lr_c;
        T1(n, nh, k, 0) = a->tau * C1(n, nh, k, 0) + (1.0 - a->tau) * T1(n, nh, k, 0);
        T1(n, nh, k, 1) = a->tau * C1(n, nh, k, 1) + (1.0 - a->tau) * T1(n, nh, k, 1);
        T2(n, nh, 0, k) = a->tau * C2(n, nh, 0, k) + (1.0 - a->tau) * T2(n, nh, 0, k);
        T2(n, nh, 1, k) = a->tau * C2(n, nh, 1, k) + (1.0 - a->tau) * T2(n, nh, 1, k);
    }
    return reward;
}

static void tick(const H3RefCfg* c, H3RefEpisode* e, const double* excitation, double* actions_out, double* rewards_out) {
    /* main_3dof.py:122-197 */
    const int k = e->step;
    const int phase2 = k >= c->switch_step;
    double st[7], act[2];
    for (int i = 0; i < 7; ++i) st[i] = e->s[i];
    const double ref_th = e->ref_th;
    const double s_lon1 = (c->lon.tracked == 1) ? st[1] - -0.0 : st[c->lon.tracked] - ref_th;
    act[1] = e->trim[1] + actor(e->lon, NHL, &c->lon, st[c->lon.ac_state], s_lon1);
    if (phase2) {
        const double s_col1 = (c->col.tracked == 1) ? st[1] - -0.0 : st[c->col.tracked] - ref_th;
        act[0] = e->trim[0] + actor(e->col, NHC, &c->col, st[c->col.ac_state], s_col1);
    } else {
        act[0] = pid(c, st, &e->pid_corr);
    }
    if (excitation && k < c->exc_len) { act[0] += excitation[2 * k]; act[1] += excitation[2 * k + 1]; }
    if (c->task == 1) (void)get_ref(c, e, e->t);            /* heli_models.py:155 */
    const int done = env_step(c, e->s, act[0], act[1]);
    e->t += c->dt;
    const double nref_th = get_ref(c, e, e->t);              /* main_3dof.py:142 */
    rls_update(c, e, st, act, e->s);
    const double r_lon = agent_update(e->lon, NHL, &c->lon, e, st, e->s, ref_th, nref_th);
    double r_col = 0;
    if (phase2) r_col = agent_update(e->col, NHC, &c->col, e, st, e->s, ref_th, nref_th);
    e->ret[0] += r_col;
    e->ret[1] += r_lon;
    e->ref_th = nref_th;
    e->step = k + 1;
    const int nan_act = (act[0] != act[0]) || (act[1] != act[1]);
    const int nf = (done ? 1 : 0) | (nan_act ? 2 : 0);
    if (nf) { e->flags = nf; e->term_step = k; }
    if (actions_out) { actions_out[0] = act[0]; actions_out[1] = act[1]; }
    if (rewards_out) { rewards_out[0] = r_col; rewards_out[1] = r_lon; }
}

/* Advance n episodes by n_ticks ticks (masked: terminated episodes freeze).  threads > 1 uses OpenMP over
 * episodes.  Returns the number of ticks executed in total. */
int64_t h3ref_run(const H3RefCfg* cfg, H3RefEpisode* eps, int64_t n, int32_t n_ticks, const double* excitation,
                  int32_t threads, double* last_actions, double* last_rewards) {
    int64_t total = 0;
#pragma omp parallel for num_threads(threads) reduction(+ : total) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        H3RefEpisode* e = eps + i;
        for (int j = 0; j < n_ticks && e->flags == 0; ++j) {
            tick(cfg, e, excitation, last_actions ? last_actions + 2 * i : 0, last_rewards ? last_rewards + 2 * i : 0);
            total += 1;
        }
    }
    return total;
}

int64_t h3ref_sizeof_episode(void) { return (int64_t)sizeof(H3RefEpisode); }
int64_t h3ref_sizeof_cfg(void) { return (int64_t)sizeof(H3RefCfg); }
