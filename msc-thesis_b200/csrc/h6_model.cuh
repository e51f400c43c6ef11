// Device functions of the 6-DOF closed-loop tick (one thread = one episode): Helicopter6DOF state derivatives
// and RK4 step with the engine lag, CollectivePID6DOF, LatPedPID, RefGenerator / FirstOrderLag.
// Each function cites the reference statement it restates (heli_models.py:347-826, PID.py:27-91,
// util.py:109-222); nothing is translated line by line: everything that depends only on the controls is
// evaluated once per tick instead of once per RK4 stage, the 3x3 flapping system is solved in closed form for
// the two angles that are used, the ISA density power law is a binomial series around sea level, trig of one
// angle is evaluated as one sincos.
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/heli6dof.h"
#include "h3_math.cuh"

namespace h6 {

using h3::Math;

// ----------------------------------------------------------------------------------------------
// launch-uniform constants (heli_models.py:351-462), derived on the host in fp64
// ----------------------------------------------------------------------------------------------
struct ModelK {
    double W, inv_mass;
    double a1_span, a1_lo, b1_span, b1_lo, t0_span, t0_lo, t0tr_span, t0tr_lo, cpsi, spsi;
    double Kgamma;                       // gamma = rho * Kgamma
    double R_mr, inv_Rmr, eps;
    double den1_c, den2_c;               // den1 = gamma*den1_c, den2 = gamma*mu_x^2*den2_c
    double A01_c, one_m_nu2, A20_c;      // A[0,1] = gamma*mu_x*A01_c, A[2,0] = gamma*mu_x*A20_c/(-den1-den2)
    double inv_2nu2;
    double b0_t0a, b0_t0b, b0_t1s, b0_twa, b0_twb, b0_lam, b0_p;    // coefficients of b[0]
    double b1_t0, b1_t1sa, b1_t1sb, b1_tw, b1_lam, b1_p;             // coefficients of b[1]
    double b2_t1ca, b2_t1cb, b2_q;                                   // coefficients of b[2]
    double Ct_c, twist, delta0, delta2, gamma_s;
    double disc_mr;                      // pi R_mr^2
    double ecc;                          // eps * m_bl
    double gear_Rtr, inv_gear_Rtr, k1tr, x_tr, z_tr, Cttr_c, disc_tr, f_tr, inv_2disc_tr, inv_Mtr;
    double F0_half, Mfus_c, x_hs, alpha0_hs, Zhs_c, x_fin, z_fin, beta0_fin, Yfin_c;
    double CDS_half, kind, PpPd_c, n_pd;
    double Keng, omega_set, inv_tau_eng;
    double xh, yh, zh;
    double Iyy_m_Ixx, Izz_m_Iyy, Ixx_m_Izz, Jxz, inv_Ixx, inv_Iyy, inv_rden, inv_rotor_inertia;
    double inv_tau_mr, inv_tau_tr;
    double rho_c[10];                    // rho(z) = sum rho_c[k] x^k, x = 0.0065 z / 288.15 (binomial series)
    double rho_expo;
};

inline ModelK make_modelk() {
    const double PI = 3.14159265358979323846;
    auto d2r = [&](double x) { return x * (PI / 180.0); };
    ModelK K;
    const double g = 9.80665, mass = 2200, Ixx = 1433, Iyy = 4973, Izz = 4099, Jxz = 660;
    const double omegareq = 44.4, R_mr = 4.912, c_mr = 0.27, Nb = 4, Cla_mr = 6.113, I_beta = 231.7, e = 0.746,
                 m_bl = 27.3, nu2 = 1.248;
    const double eps = e / R_mr, sigma_mr = Nb * c_mr / (PI * R_mr);
    K.W = mass * g;
    K.inv_mass = 1.0 / mass;
    const double a1sl = d2r(-6), a1su = d2r(4), b1sl = d2r(10), b1su = d2r(-5.5), t0l = d2r(2), t0u = d2r(18),
                 t0trl = d2r(18), t0tru = d2r(-6), psipmu = d2r(-10);
    K.a1_span = a1su - a1sl; K.a1_lo = a1sl; K.b1_span = b1su - b1sl; K.b1_lo = b1sl;
    K.t0_span = t0u - t0l; K.t0_lo = t0l; K.t0tr_span = t0tru - t0trl; K.t0tr_lo = t0trl;
    K.cpsi = std::cos(psipmu); K.spsi = std::sin(psipmu);
    K.Kgamma = Cla_mr * c_mr * (R_mr * R_mr * R_mr * R_mr) / I_beta;
    K.R_mr = R_mr; K.inv_Rmr = 1.0 / R_mr; K.eps = eps;
    K.den1_c = 0.25 * eps * eps - eps / 3 + 0.125;
    K.den2_c = 0.0625 * eps * eps - eps * 0.125 + 0.0625;
    K.A01_c = (2 * eps * eps - eps) / (8 * nu2);
    K.one_m_nu2 = 1 - nu2;
    K.A20_c = 1.0 / 6 - 0.25 * eps;
    K.inv_2nu2 = 1.0 / (2 * nu2);
    K.b0_t0a = 0.25 - eps / 3; K.b0_t0b = 0.25 * eps * eps - 0.5 * eps + 0.25; K.b0_t1s = 0.5 * eps - 1.0 / 3;
    K.b0_twa = 0.2 - 0.25 * eps; K.b0_twb = 1.0 / 6 - eps * 0.25; K.b0_lam = 1.0 / 3 - 0.5 * eps;
    K.b0_p = 1.0 / 6 - 0.25 * eps;
    K.b1_t0 = (2 - 3 * eps) / 6; K.b1_t1sa = (-3 * eps * eps + 6 * eps - 3) / 16; K.b1_t1sb = (16.0 / 6 * eps - 2) / 16;
    K.b1_tw = 1 - 4.0 / 3 * eps; K.b1_lam = eps * eps - 2 * eps + 1; K.b1_p = (1 - 4.0 / 3 * eps) / 8;
    K.b2_t1ca = (8.0 / 3 * eps - 2) / 16; K.b2_t1cb = (-eps * eps + 2 * eps - 1) / 16; K.b2_q = (1 - 4.0 / 3 * eps) / 8;
    K.Ct_c = Cla_mr * sigma_mr / 8; K.twist = -0.14; K.delta0 = 0.0074; K.delta2 = 38.66; K.gamma_s = 0.0524;
    K.disc_mr = PI * R_mr * R_mr;
    K.ecc = eps * m_bl;
    const double R_tr = 0.95, c_tr = 0.18, gearing = 5.25, Nb_tr = 2, Cla_tr = 5.7, S_fin = 0.805;
    const double sigma_tr = Nb_tr * c_tr / (PI * R_tr);
    K.gear_Rtr = gearing * R_tr; K.inv_gear_Rtr = 1.0 / (gearing * R_tr); K.k1tr = 1; K.x_tr = 6.08; K.z_tr = 1.72;
    K.Cttr_c = Cla_tr * sigma_tr / 12; K.disc_tr = PI * R_tr * R_tr;
    K.f_tr = 1 - 3 * S_fin / (4 * R_tr * R_tr * PI);
    K.inv_2disc_tr = 1.0 / (2 * PI * R_tr * R_tr); K.inv_Mtr = 1.0 / 0.7;
    const double F0 = 1.3 * 0.73, K_fus = 0.83, Vol_fus = PI / 4 * 6 * 1.3;
    K.F0_half = 0.5 * F0; K.Mfus_c = K_fus * Vol_fus;
    K.x_hs = 4.64; K.alpha0_hs = 0.0698; K.Zhs_c = -0.5 * 0.65 * 0.803 * 4;
    K.x_fin = 5.30; K.z_fin = 0.97; K.beta0_fin = -0.06116; K.Yfin_c = -0.5 * S_fin * 4;
    K.CDS_half = 1.2 / 2; K.kind = 1.15; K.PpPd_c = sigma_mr * PI * R_mr * R_mr / 8; K.n_pd = 4.65;
    const double Pe = 313000, etam = 0.95, Kaeo = 0.8643;
    K.Keng = -Pe * etam * Kaeo / omegareq / 0.02; K.omega_set = omegareq * 1.02; K.inv_tau_eng = 1.0 / 0.3;
    K.xh = 0.08; K.yh = 0; K.zh = 1.48;
    K.Iyy_m_Ixx = Iyy - Ixx; K.Izz_m_Iyy = Izz - Iyy; K.Ixx_m_Izz = Ixx - Izz; K.Jxz = Jxz;
    K.inv_Ixx = 1.0 / Ixx; K.inv_Iyy = 1.0 / Iyy; K.inv_rden = 1.0 / (Izz - Jxz * Jxz / Ixx);
    K.inv_rotor_inertia = 1.0 / (1.1 * Nb * I_beta);
    K.inv_tau_mr = 1.0 / 0.1; K.inv_tau_tr = 1.0 / 0.3;
    // heli_models.py:826: rho = 1.225 (1 - 0.0065 h / 288.15)^a, h = -z  ->  1.225 (1 + x)^a
    const double a = -(9.80665 / (287.05 * -0.0065) + 1);
    K.rho_expo = a;
    double c = 1.225;
    for (int k = 0; k < 10; ++k) { K.rho_c[k] = c; c *= (a - k) / (k + 1); }
    return K;
}

struct Controls {           // blade angles from the clipped stick positions (heli_models.py:499-507)
    double theta_0, theta_1s, theta_1c, theta_0_tr;
};

H3_DEV double clip01(double x) { return fmin(fmax(x, 0.0), 1.0); }     // NaN-propagating? fmin/fmax drop NaN:
H3_DEV double clip01_nan(double x) { return (x != x) ? x : clip01(x); } // np.clip keeps it (train.py:176)

H3_DEV Controls make_controls(const ModelK& K, double coll, double lng, double lat, double pedal) {
    Controls c;
    const double a1s2 = K.a1_span * lat + K.a1_lo, b1s2 = K.b1_span * lng + K.b1_lo;
    c.theta_0 = K.t0_span * coll + K.t0_lo;
    c.theta_1s = b1s2 * K.cpsi + a1s2 * K.spsi;
    c.theta_1c = a1s2 * K.cpsi + b1s2 * K.spsi;
    c.theta_0_tr = K.t0tr_span * pedal + K.t0tr_lo;
    return c;
}

H3_DEV double air_density(const ModelK& K, double z) {      // heli_models.py:818-826 with altitude = -z
    const double x = h3::div_nr(0.0065 * z, 288.15);
    if (fabs(x) > 0.04) return 1.225 * exp(K.rho_expo * log(1.0 + x));     // (pow has an out-of-line slow path)
    double r = K.rho_c[9];
#pragma unroll
    for (int k = 8; k >= 0; --k) r = r * x + K.rho_c[k];
    return r;
}

// Helicopter6DOF.calculate_state_derivatives (heli_models.py:489-694).  P_out is the engine lag state: when
// not trimming it is advanced on EVERY evaluation (four times per RK4 step, :629-632).
H3_DEV void derivatives(const ModelK& K, const Controls& c, const double (&s)[15], double& P_out, double P_avail,
                        bool trimming, double dt, double (&d)[15]) {
    const double u = s[0], v = s[1], w = s[2], p = s[3], q = s[4], r = s[5], phi = s[6], theta = s[7], psi = s[8],
                 z = s[11], l0mr = s[12], l0tr = s[13], omega = s[14];
    const double rho = air_density(K, z);
    const double gamma = rho * K.Kgamma;
    const double alpha_fus = h3::atan2_nr(w, u);                               // :572
    const double V2 = u * u + v * v + w * w, V = h3::sqrt_nr(V2);
    const double omega_r = omega * K.R_mr, inv_or = h3::div_nr(1.0, omega_r), inv_om = K.R_mr * inv_or;
    double sa, ca, sf, cf;
    h3::sincos_f64(c.theta_1s - alpha_fus, &sa, &ca);                   // alpha_cp, :513
    h3::sincos_f64(alpha_fus, &sf, &cf);
    const double mu_x = V * ca * inv_or, mu_z = -V * sa * inv_or, mu2 = mu_x * mu_x;
    // flapping (:522-547): rows of A x = b, only a1 and b1 are used
    const double den1 = gamma * K.den1_c, den2 = gamma * mu2 * K.den2_c;
    const double i12 = h3::div_nr(1.0, den1 - den2), i2 = -h3::div_nr(1.0, den1 + den2);
    const double A01 = gamma * mu_x * K.A01_c, A12 = K.one_m_nu2 * i12, A20 = gamma * mu_x * K.A20_c * i2,
                 A21 = K.one_m_nu2 * i2;
    const double lam = l0mr - mu_z, p_om = p * inv_om, q_om = q * inv_om;
    const double b0 = gamma * K.inv_2nu2 * (c.theta_0 * (K.b0_t0a + mu2 * K.b0_t0b) + mu_x * K.b0_t1s * c.theta_1s +
                                            (K.b0_twa + mu2 * K.b0_twb) * K.twist - K.b0_lam * lam + mu_x * K.b0_p * p_om);
    const double b1v = (gamma * (mu_x * K.b1_t0 * c.theta_0 + (mu2 * K.b1_t1sa + K.b1_t1sb) * c.theta_1s +
                                 mu_x * 0.25 * (K.b1_tw * K.twist - K.b1_lam * lam) + K.b1_p * p_om) - 2 * q_om) * i12;
    const double b2 = (gamma * ((K.b2_t1ca + mu2 * K.b2_t1cb) * c.theta_1c + K.b2_q * q_om) + 2 * p_om) * i2;
    const double m = A21 - A20 * A01;
    const double b1 = h3::div_nr(b2 - A20 * b0 - m * b1v, 1.0 - m * A12);
    const double a1 = b1v - A12 * b1;
    const double Ct = K.Ct_c * (2 * (2.0 / 3 + mu2) * c.theta_0 + (2 * c.theta_1s + p_om) * mu_x + 2 * (mu_z - l0mr) +
                                (1 + mu2) * K.twist);                    // :548-552
    const double Cd = K.delta0 + K.delta2 * Ct * Ct;
    const double or2 = omega_r * omega_r;
    const double T_mr = Ct * rho * or2 * K.disc_mr;
    double s1, c1, s2, c2;
    h3::sincos_f64(a1 - c.theta_1s + K.gamma_s, &s1, &c1);              // a1t1s
    h3::sincos_f64(b1 + c.theta_1c, &s2, &c2);                          // b1t1c
    const double Le = or2 * K.ecc * s2, Me = or2 * K.ecc * s1;          // :561-562
    const double Ct_gl_mr = 2 * l0mr * h3::sqrt_nr(mu2 + lam * lam);           // :566
    // tail rotor (:572-589)
    const double omegar_tr = K.gear_Rtr * omega, inv_otr = K.inv_gear_Rtr * inv_om;
    const double wq = w + K.k1tr * l0mr * omega_r + q * K.x_tr;
    const double mu_x_tr2 = (u * u + wq * wq) * inv_otr * inv_otr;
    const double mu_z_tr = -(v - K.x_tr * r + K.z_tr * p) * inv_otr;
    const double Ct_bem_tr = K.Cttr_c * (c.theta_0_tr * (2 + 3 * mu_x_tr2) + 3 * mu_z_tr - 6 * l0tr);
    const double dl = mu_z_tr - l0tr;
    const double Ct_gl_tr = 2 * l0tr * h3::sqrt_nr(mu_x_tr2 + dl * dl);
    const double T_tr = Ct_bem_tr * rho * omegar_tr * omegar_tr * K.disc_tr;
    const double Y_tr = T_tr * K.f_tr, L_tr = Y_tr * K.z_tr, N_tr = -Y_tr * K.x_tr;
    // fuselage (:595-601)
    const double R_fus = K.F0_half * rho * V2;
    const double X_fus = -R_fus * cf, Z_fus = -R_fus * sf;
    const double M_fus = rho * V2 * K.Mfus_c * alpha_fus;
    // horizontal stabilizer (:607-612), fin (:615-621)
    const double w_hs = w + q * K.x_hs;
    const double Z_hs = K.Zhs_c * rho * (u * u + w_hs * w_hs) * (K.alpha0_hs + h3::atan2_nr(w_hs, u));
    const double M_hs = Z_hs * K.x_hs;
    const double v_fin = v - r * K.x_fin + p * K.z_fin;
    const double Y_fin = K.Yfin_c * rho * (u * u + v_fin * v_fin) * (K.beta0_fin + h3::atan2_nr(v_fin, u));
    const double L_fin = K.z_fin * Y_fin, N_fin = -K.x_fin * Y_fin;
    // power (:626-643)
    const double P_par = K.CDS_half * rho * V2 * V;
    const double P_i = fabs(K.kind * T_mr * l0mr * omega_r);
    const double PpPd = K.PpPd_c * Cd * rho * or2 * omega_r * (1 + K.n_pd * mu2);
    const double P_c = -w * K.W;
    const double P_tr = fabs(T_tr * K.inv_Mtr * h3::sqrt_nr(fabs(h3::div_nr(T_tr * K.inv_2disc_tr, rho))));
    const double P_req = P_par + P_i + PpPd + P_c + P_tr;
    double P_eng = K.Keng * (omega - K.omega_set);
    if (trimming) {
        P_out = P_eng;
    } else {
        P_out += (P_eng - P_out) * K.inv_tau_eng * dt;
        P_eng = P_out;
    }
    P_eng = (P_eng != P_eng) ? P_eng : fmin(fmax(P_eng, 0.0), P_avail);
    // rotor forces and moments (:646-653)
    const double Xmr = -T_mr * s1 * c2, Ymr = T_mr * s2, Zmr = -T_mr * c1 * c2;
    const double Lmr = Ymr * K.zh - Zmr * K.yh + Le;
    const double Mmr = -Xmr * K.zh - Zmr * K.xh + Me;
    const double Nmr = P_eng * inv_om + Xmr * K.yh - Ymr * K.xh;
    double sph, cph, sth, cth, sps, cps;
    h3::sincos_f64(phi, &sph, &cph);
    h3::sincos_f64(theta, &sth, &cth);
    h3::sincos_f64(psi, &sps, &cps);
    const double Fx = -K.W * sth + (Xmr + X_fus);
    const double Fy = K.W * cth * sph + (Ymr + Y_tr + Y_fin);
    const double Fz = K.W * cth * cph + (Zmr + Z_fus + Z_hs);
    const double L = Lmr + L_tr + L_fin, M = Mmr + M_fus + M_hs, N = Nmr + N_tr + N_fin;
    d[0] = Fx * K.inv_mass - q * w + r * v;                             // :673-675
    d[1] = Fy * K.inv_mass - r * u + p * w;
    d[2] = Fz * K.inv_mass - p * v + q * u;
    const double Lq = L - K.Izz_m_Iyy * q * r;
    const double rdot = (N - K.Iyy_m_Ixx * p * q + K.Jxz * ((Lq + K.Jxz * p * q) * K.inv_Ixx - r * q)) * K.inv_rden;
    d[5] = rdot;
    d[4] = (M - K.Ixx_m_Izz * r * p - K.Jxz * (p * p - r * r)) * K.inv_Iyy;
    d[3] = (Lq + K.Jxz * (rdot + p * q)) * K.inv_Ixx;
    const double psidot = h3::div_nr(q * sph + r * cph, cth);                    // :681-683
    d[8] = psidot;
    d[7] = q * cph - r * sph;
    d[6] = p + psidot * sth;
    d[14] = rdot + (P_eng - P_req) * inv_om * K.inv_rotor_inertia;      // :686
    const double vw_s = v * sph + w * cph, vw_c = v * cph - w * sph;    // :689-693
    const double fwd = u * cth + vw_s * sth;
    d[9] = fwd * cps - vw_c * sps;
    d[10] = fwd * sps + vw_c * cps;
    d[11] = -u * sth + vw_s * cth;
    d[12] = (Ct - Ct_gl_mr) * K.inv_tau_mr;                             // :695-696
    d[13] = (Ct_bem_tr - Ct_gl_tr) * K.inv_tau_tr;
}

// Helicopter6DOF.integrate_runge_kutta (heli_models.py:696-721).  One copy of the derivative code, looped over
// the four stages: acc = k1 + 2 k2 + 2 k3 + k4 summed in that order, stage states old + k/2, old + k/2, old + k
// (the factors 0.5, 1, 2 are exact, so this is the reference's operator sequence).
H3_DEV void rk4_step(const ModelK& K, const Controls& c, const double (&old)[15], double& P_out, double P_avail,
                     double dt, double (&nw)[15]) {
    double st[15], acc[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { st[i] = old[i]; acc[i] = 0.0; }
#pragma unroll 1
    for (int stage = 0; stage < 4; ++stage) {
        double f[15];
        derivatives(K, c, st, P_out, P_avail, false, dt, f);
        const double wt = (stage == 0 || stage == 3) ? 1.0 : 2.0, half = (stage < 2) ? 0.5 : 1.0;
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            const double k = dt * f[i];
            acc[i] = (stage == 0) ? k : acc[i] + wt * k;
            st[i] = old[i] + k * half;
        }
    }
#pragma unroll
    for (int i = 0; i < 15; ++i) nw[i] = old[i] + h3::div_nr(acc[i], 6.0);
}

// The same step with the two arrays that live across the four stages (the step's initial state and the weighted
// sum of the increments) parked in shared memory, `sm[row * stride]`, rows [0, 15) = old, [15, 30) = acc: 60 fewer
// registers during the derivative evaluations, i.e. one more resident block per SM for the latency-bound kernel.
H3_DEV void rk4_step_smem(const ModelK& K, const Controls& c, const double (&old)[15], double& P_out, double P_avail,
                          double dt, double (&nw)[15], double* sm, int stride) {
    double st[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { st[i] = old[i]; sm[i * stride] = old[i]; }
#pragma unroll 1
    for (int stage = 0; stage < 4; ++stage) {
        double f[15];
        derivatives(K, c, st, P_out, P_avail, false, dt, f);
        const double wt = (stage == 0 || stage == 3) ? 1.0 : 2.0, half = (stage < 2) ? 0.5 : 1.0;
#pragma unroll
        for (int i = 0; i < 15; ++i) {
            const double k = dt * f[i];
            sm[(15 + i) * stride] = (stage == 0) ? k : sm[(15 + i) * stride] + wt * k;
            st[i] = sm[i * stride] + k * half;
        }
    }
#pragma unroll
    for (int i = 0; i < 15; ++i) nw[i] = sm[i * stride] + h3::div_nr(sm[(15 + i) * stride], 6.0);
}

// ----------------------------------------------------------------------------------------------
// PIDs.  sph..cps: sin / cos of phi, theta, psi of the observation (shared by both controllers)
// ----------------------------------------------------------------------------------------------
struct ObsTrig {
    double sph, cph, sth, cth, sps, cps;
};
H3_DEV ObsTrig obs_trig(const double (&o)[15]) {
    ObsTrig t;
    h3::sincos_f64(o[6], &t.sph, &t.cph);
    h3::sincos_f64(o[7], &t.sth, &t.cth);
    h3::sincos_f64(o[8], &t.sps, &t.cps);
    return t;
}

// CollectivePID6DOF.__call__ (PID.py:31-45)
H3_DEV double pid_collective(const H6Config& C, const double (&o)[15], const ObsTrig& g, double col_trim, double& corr) {
    const double hdot_ref = C.col_kd * (C.col_href - (-o[11]));
    const double hdot = o[0] * g.sth - (o[1] * g.sph + o[2] * g.cph) * g.cth;
    const double err = hdot_ref - hdot;
    const double collective = col_trim + C.col_kp * err + C.col_ki * corr;
    corr += C.dt * err;
    return clip01_nan(collective);
}

// LatPedPID.__call__ (PID.py:62-91); integ = {y_req_int, phi_int, psi_int}
H3_DEV void pid_latped(const H6Config& C, const double (&o)[15], const ObsTrig& g, double phi_trim, double lat_trim,
                       double ped_trim, double& y_int, double& phi_int, double& psi_int, double& lat, double& ped) {
    const double* G = C.lat_gains;
    const double ydot = (o[0] * g.cth + (o[1] * g.sph + o[2] * g.cph) * g.sth) * g.sps + (o[1] * g.cph - o[2] * g.sph) * g.cps;
    const double y_error = 0.0 - o[10];
    const double phi_req = phi_trim + G[0] * y_error + G[1] * y_int + G[2] * ydot;
    const double phi_error = phi_req - o[6], psi_error = 0.0 - o[8];
    lat = clip01_nan(lat_trim + (G[3] * phi_error + G[4] * phi_int + G[5] * o[3]) / 100);
    ped = clip01_nan(ped_trim + (G[6] * psi_error + G[7] * psi_int + G[8] * o[5]) / 100);
    if (0.01 < lat && lat < 0.99) { y_int += y_error * C.dt; phi_int += phi_error * C.dt; }
    if (0.01 < ped && ped < 0.99) psi_int += psi_error * C.dt;
}

// ----------------------------------------------------------------------------------------------
// RefGenerator (util.py:109-181).  gen: the H6_GEN_* words of one episode; ref = {u_ref, theta_ref, z_ref}
// ----------------------------------------------------------------------------------------------
H3_DEV double lag_filter(const double (&gen)[H6_GEN_ROWS], double t) {     // FirstOrderLag.get_result, util.py:212
    return gen[H6_GEN_FILTER_X0] +
           gen[H6_GEN_FILTER_SETPOINT] * (1.0 - exp(-1.0 * (t - gen[H6_GEN_FILTER_T0]) / gen[H6_GEN_FILTER_TAU]));
}
H3_DEV long long round4(double t) { return llrint(t * 10000.0); }          // round(t, 4) as a count of 1e-4 s

H3_DEV void gen_get_ref(const H6Config& C, int task, double (&gen)[H6_GEN_ROWS], const double (&obs)[15], double t,
                        double (&ref)[3]) {
    const double D2R = 3.14159265358979323846 / 180.0;
    double theta_ref = 0.0, z_ref = 0.0, u_ref = nan("");
    if (task == H6_TASK_HOVER) {
        z_ref = gen[H6_GEN_ZREF];
    } else if (task == H6_TASK_TRAIN_LON) {
        theta_ref = (gen[H6_GEN_A] * sin(2 * 3.14159265358979323846 * t / C.gen_T)) * D2R;
    } else if (task == H6_TASK_TRAIN_COL) {
        theta_ref = obs[0] > 1 ? 3 * D2R : (2.5 * D2R) * obs[0];
        z_ref = fmax(gen[H6_GEN_ZREF] - 1 * (t - gen[H6_GEN_TSWITCH]), -25.0);
    } else if (task == H6_TASK_VELOCITY) {
        z_ref = gen[H6_GEN_ZREF];
        u_ref = lag_filter(gen, t);
        const double u_error = u_ref - obs[0];
        const double ku = round4(t) < 200000 ? -2.0 : -3.0;
        theta_ref = (ku * u_error + -0.05 * gen[H6_GEN_INT_ERROR_U]) * D2R;
        gen[H6_GEN_INT_ERROR_U] += u_error * C.gen_dt;
    } else if (task == H6_TASK_DESCENT) {
        const long long t4 = round4(t);
        if (t4 == 129300) {
            gen[H6_GEN_FILTER_T0] = t; gen[H6_GEN_FILTER_X0] = obs[0]; gen[H6_GEN_FILTER_SETPOINT] = -obs[0];
            gen[H6_GEN_FILTER_TAU] = 2.5;
        }
        if (t4 <= 129200) {
            theta_ref = -0.1 * D2R;
            z_ref = gen[H6_GEN_ZREF] + 1.93 * (t - gen[H6_GEN_TSWITCH]);
        } else {
            u_ref = lag_filter(gen, t);
            const double u_error = u_ref - obs[0];
            theta_ref = (-3.5 * u_error + 0.1 * gen[H6_GEN_INT_ERROR_U]) * D2R;
            gen[H6_GEN_INT_ERROR_U] += u_error * C.gen_dt;
            z_ref = -5 + 0.5 * (t - 12.92);
        }
    }
    ref[0] = u_ref; ref[1] = theta_ref; ref[2] = z_ref;
}

H3_DEV void gen_set_task(int& task_io, double (&gen)[H6_GEN_ROWS], int task, double t, const double (&obs)[15],
                         double z_start, double filter_target, double t_switch) {   // util.py:168-181
    task_io = task;
    gen[H6_GEN_INT_ERROR_U] = 0;
    if (task == H6_TASK_TRAIN_COL) gen[H6_GEN_ZREF] = z_start;
    else if (task == H6_TASK_HOVER) gen[H6_GEN_ZREF] = obs[11];
    else if (task == H6_TASK_VELOCITY) {
        gen[H6_GEN_ZREF] = z_start;
        gen[H6_GEN_FILTER_T0] = t; gen[H6_GEN_FILTER_X0] = obs[0]; gen[H6_GEN_FILTER_SETPOINT] = filter_target;
    } else if (task == H6_TASK_DESCENT) { gen[H6_GEN_TSWITCH] = t_switch; gen[H6_GEN_ZREF] = obs[11]; }
}

// reference slot of a tracked state: only slots 0, 7, 11 are ever set (util.py:123, :164-165)
H3_DEV double ref_slot(const double (&ref)[3], int tracked) {
    return tracked == 0 ? ref[0] : (tracked == 7 ? ref[1] : (tracked == 11 ? ref[2] : nan("")));
}

}  // namespace h6
