// Precision-dispatched device math for the 3-DOF tick kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "h3_coef.cuh"

namespace h3 {

#define H3_DEV __device__ __forceinline__

// fp64 tanh / sincos written for this kernel: every polynomial coefficient is a constant-bank operand of
// its DFMA (tools/fit_math_coef.py derives them; <= 1.6 ulp end to end), where the inlined CUDA library
// routine spends two UMOV issue slots per coefficient.  tanh has two straight-line branches:
//   |x| <  0.55 : x + x z P(z), z = x^2                      (hidden pre-activations of the cyclic agent)
//   otherwise   : E / (E + 2), E = expm1(2x)                 (collective agent: inputs in m and m/s)
// and the branch is taken once per CHUNK of hidden units (tanh_vec), so each side is straight-line code in
// which the compiler interleaves the independent polynomial chains.
// sin/cos use one polynomial pair on |x| <= pi/4 (flight angles) and the library, out of line, beyond.
static __device__ __noinline__ void sincos_libm(double x, double* s, double* c) { ::sincos(x, s, c); }

// ~20-bit reciprocal seed from the high word: one MUFU.RCP64H.  (__frcp_rn on the float-converted value is
// IEEE-rounded and drags a slow-path branch into every element, which also stops the interleaving.)
H3_DEV double rcp_seed(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

H3_DEV double tanh_small(double x) {
    const double z = x * x;
    double p = H3_TANH_C[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, z, H3_TANH_C[i]);
    return fma(x * z, p, x);
}

// tanh(x) = E / (E + 2), E = expm1(2x) = 2^n expm1(r) + (2^n - 1), 2x = n ln2 + r: one straight-line formula
// without cancellation for every x (2.4 ulp), used whenever a chunk holds an argument >= 0.55.
H3_DEV double tanh_any(double x) {
    const double xc = x > 20.0 ? 20.0 : (x < -20.0 ? -20.0 : x);    // |tanh| == 1 to the last bit beyond 19.1; NaN passes
    const double y = xc + xc;
    // n = rint(y / ln2) by the 1.5 * 2^52 trick: the integer lands in the low word of t
    const double t = fma(y, H3_LN2[2], 6755399441055744.0);
    const double nf = t - 6755399441055744.0;
    double r = fma(nf, -H3_LN2[0], y);
    r = fma(nf, -H3_LN2[1], r);                             // |r| <= ln2 / 2
    double q = H3_EXP_C[10];
#pragma unroll
    for (int i = 9; i >= 0; --i) q = fma(q, r, H3_EXP_C[i]);
    const double p = fma(r * r, q, r);                      // expm1(r)
    const int n = __double2loint(t);                        // in [-58, 58]; NaN: E is NaN whatever n is
    const double sc = __hiloint2double((n + 1023) << 20, 0);
    const double E = fma(sc, p, sc - 1.0);
    const double den = E + 2.0;                             // in [1, 2^58]: no special cases in the division
    double rc = rcp_seed(den);
    rc = fma(rc, fma(-den, rc, 1.0), rc);
    rc = fma(rc, fma(-den, rc, 1.0), rc);
    const double q0 = E * rc;
    return fma(fma(-q0, den, E), rc, q0);
}

H3_DEV double tanh_f64(double x) {
    if (::fabs(x) < 0.55) return tanh_small(x);             // NaN compares false -> tanh_any -> NaN
    return tanh_any(x);
}

// sin / cos polynomials on the reduced argument |r| <= pi/4
H3_DEV void sincos_poly(double r, double* s, double* c) {
    const double z = r * r;
    double ps = H3_SIN_C[7], pc = H3_COS_C[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) {
        ps = fma(ps, z, H3_SIN_C[i]);
        pc = fma(pc, z, H3_COS_C[i]);
    }
    *s = fma(r * z, ps, r);
    *c = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// |x| <= 1e5: Cody-Waite reduction by multiples of pi/2 (three FMA pieces), polynomial pair, quadrant swap --
// straight-line code.  Pitch angles between 45 and 90 degrees (episodes on their way to termination) stay on
// this path instead of dragging their whole warp through the library's out-of-line routine.
H3_DEV void sincos_reduced(double x, double* s, double* c) {
    const double t = fma(x, H3_PIO2[3], 6755399441055744.0);      // q = rint(x * 2/pi) in the low word of t
    const double q = t - 6755399441055744.0;
    double r = fma(-q, H3_PIO2[0], x);
    r = fma(-q, H3_PIO2[1], r);
    r = fma(-q, H3_PIO2[2], r);
    double sr, cr;
    sincos_poly(r, &sr, &cr);
    const int n = __double2loint(t);
    const double a = (n & 1) ? cr : sr, b = (n & 1) ? sr : cr;    // sin x, cos x up to sign
    *s = (n & 2) ? -a : a;
    *c = ((n + 1) & 2) ? -b : b;
}

H3_DEV void sincos_f64(double x, double* s, double* c) {
    const double ax = ::fabs(x);
    if (ax <= 0.78539816339744828) {
        sincos_poly(x, s, c);
    } else if (ax <= 1.0e5) {
        sincos_reduced(x, s, c);
    } else {
        sincos_libm(x, s, c);                                     // huge, inf, NaN
    }
}

#ifndef H3_TANH_G
#define H3_TANH_G 5
#endif

// tanh_any on elements [G0, min(G0 + H3_TANH_G, U)) of x, interleaved, then the next group
template <int G0, int U>
H3_DEV void tanh_any_group(const double (&x)[U], double (&y)[U]) {
    constexpr int G = (U - G0) < H3_TANH_G ? (U - G0) : H3_TANH_G;
    if constexpr (G > 0) {
        double r[G], q[G], t[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const double xu = x[G0 + u];
            const double xc = xu > 20.0 ? 20.0 : (xu < -20.0 ? -20.0 : xu);
            const double yy = xc + xc;
            t[u] = fma(yy, H3_LN2[2], 6755399441055744.0);
            const double nf = t[u] - 6755399441055744.0;
            r[u] = fma(nf, -H3_LN2[1], fma(nf, -H3_LN2[0], yy));
            q[u] = H3_EXP_C[10];
        }
#pragma unroll
        for (int i = 9; i >= 0; --i)
#pragma unroll
            for (int u = 0; u < G; ++u) q[u] = fma(q[u], r[u], H3_EXP_C[i]);
        double E[G], den[G], rc[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const double p = fma(r[u] * r[u], q[u], r[u]);
            const double sc = __hiloint2double((__double2loint(t[u]) + 1023) << 20, 0);
            E[u] = fma(sc, p, sc - 1.0);
            den[u] = E[u] + 2.0;
            rc[u] = rcp_seed(den[u]);
        }
#pragma unroll
        for (int u = 0; u < G; ++u) rc[u] = fma(rc[u], fma(-den[u], rc[u], 1.0), rc[u]);
#pragma unroll
        for (int u = 0; u < G; ++u) rc[u] = fma(rc[u], fma(-den[u], rc[u], 1.0), rc[u]);
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const double q0 = E[u] * rc[u];
            y[G0 + u] = fma(fma(-q0, den[u], E[u]), rc[u], q0);
        }
        tanh_any_group<G0 + G, U>(x, y);
    }
}

template <typename T>
struct Math;

template <>
struct Math<double> {
    static H3_DEV double tanh_(double x) { return tanh_f64(x); }
    static H3_DEV double tanh_wide_(double x) { return tanh_any(x); }
    // U independent arguments at once; the polynomial loops run coefficient-major so that consecutive DFMAs
    // belong to different elements (a single warp per scheduler then covers the DFMA latency by itself).
    // Returns whether the short (polynomial) branch was taken: the kernels keep that per episode as a sort key, so
    // that episodes taking the same branches can be grouped into the same warps (ClosedLoop3DOF.compact).
    // wide: every element takes the full-range formula (no test, no divergence): for the collective agent, whose
    // inputs are a speed in m/s and an altitude error in m -- its pre-activations leave the polynomial's range as soon
    // as the episode is under way, and a warp in which only SOME lanes have left it pays for both branches.
    template <int U>
    static H3_DEV bool tanh_vec(const double (&x)[U], double (&y)[U], bool wide = false) {
        if (wide) {
            tanh_any_group<0, U>(x, y);
            return false;
        }
        bool all_small = true;
#pragma unroll
        for (int u = 0; u < U; ++u) all_small = all_small && (::fabs(x[u]) < 0.55);
        if (all_small) {
            double z[U], p[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                z[u] = x[u] * x[u];
                p[u] = H3_TANH_C[11];
            }
#pragma unroll
            for (int i = 10; i >= 0; --i)
#pragma unroll
                for (int u = 0; u < U; ++u) p[u] = fma(p[u], z[u], H3_TANH_C[i]);
#pragma unroll
            for (int u = 0; u < U; ++u) y[u] = fma(x[u] * z[u], p[u], x[u]);
        } else {
            // at most H3_TANH_G elements in flight: each carries ~6 live doubles through the pipeline below
            tanh_any_group<0, U>(x, y);
        }
        return all_small;
    }
    // sin/cos of two independent angles (theta and the cyclic pitch): four interleaved polynomial chains
    static H3_DEV bool sincos2_(double a, double b, double* sa, double* ca, double* sb, double* cb) {
        const bool poly = ::fabs(a) <= 0.78539816339744828 && ::fabs(b) <= 0.78539816339744828;
        if (poly) {
            const double za = a * a, zb = b * b;
            double psa = H3_SIN_C[7], pca = H3_COS_C[7], psb = H3_SIN_C[7], pcb = H3_COS_C[7];
#pragma unroll
            for (int i = 6; i >= 0; --i) {
                psa = fma(psa, za, H3_SIN_C[i]);
                pca = fma(pca, za, H3_COS_C[i]);
                psb = fma(psb, zb, H3_SIN_C[i]);
                pcb = fma(pcb, zb, H3_COS_C[i]);
            }
            *sa = fma(a * za, psa, a);
            *ca = fma(za * za, pca, fma(-0.5, za, 1.0));
            *sb = fma(b * zb, psb, b);
            *cb = fma(zb * zb, pcb, fma(-0.5, zb, 1.0));
        } else if (::fabs(a) <= 1.0e5 && ::fabs(b) <= 1.0e5) {
            sincos_reduced(a, sa, ca);
            sincos_reduced(b, sb, cb);
        } else {
            sincos_f64(a, sa, ca);
            sincos_f64(b, sb, cb);
        }
        return poly;
    }
    static H3_DEV double exp_(double x) { return ::exp(x); }
    static H3_DEV double sqrt_(double x) { return ::sqrt(x); }
    static H3_DEV double abs_(double x) { return ::fabs(x); }
    static H3_DEV void sincos_(double x, double* s, double* c) { sincos_f64(x, s, c); }
    static H3_DEV double sigmoid_(double x) { return 1.0 / (1.0 + ::exp(-x)); }
};

template <>
struct Math<float> {
    static H3_DEV float tanh_(float x) { return ::tanhf(x); }
    static H3_DEV float tanh_wide_(float x) { return ::tanhf(x); }
    template <int U>
    static H3_DEV bool tanh_vec(const float (&x)[U], float (&y)[U], bool = false) {
#pragma unroll
        for (int u = 0; u < U; ++u) y[u] = ::tanhf(x[u]);
        return true;
    }
    static H3_DEV bool sincos2_(float a, float b, float* sa, float* ca, float* sb, float* cb) {
        ::sincosf(a, sa, ca);
        ::sincosf(b, sb, cb);
        return true;
    }
    static H3_DEV float exp_(float x) { return ::expf(x); }
    static H3_DEV float sqrt_(float x) { return ::sqrtf(x); }
    static H3_DEV float abs_(float x) { return ::fabsf(x); }
    static H3_DEV void sincos_(float x, float* s, float* c) { ::sincosf(x, s, c); }
    static H3_DEV float sigmoid_(float x) { return 1.0f / (1.0f + ::expf(-x)); }
};

// np.clip semantics (PID.py:24): NaN passes through (fmin/fmax would drop it).
template <typename T>
H3_DEV T clip_nan(T x, T lo, T hi) {
    return x < lo ? lo : (x > hi ? hi : x);
}

// element `idx` of a 7-vector held in registers (runtime but warp-uniform index)
template <typename T, int N>
H3_DEV T pick7(const T (&s)[N], int idx) {
    static_assert(N >= 7, "a state vector (or an array that starts with one)");
    T r = s[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) r = (idx == i) ? s[i] : r;
    return r;
}

}  // namespace h3
