// Precision-dispatched device math for the 3-DOF tick kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "h3_coef.cuh"

namespace h3 {

#define H3_DEV __device__ __forceinline__

// fp64 tanh / sincos written for this kernel: every polynomial coefficient is a constant-bank operand of
// its DFMA (tools/fit_math_coef.py derives them), where the inlined CUDA library routine spends two UMOV issue
// slots per coefficient.  tanh is ONE straight-line, table-driven formula for every argument (tanh_tab below): no
// data-dependent branch, so episodes whose pre-activations are small and episodes that saturate share a warp for free.
// sin/cos: Cody-Waite reduction by multiples of pi/2 and one polynomial pair, also one formula for every argument.

// ~20-bit reciprocal seed from the high word: one MUFU.RCP64H.  (__frcp_rn on the float-converted value is
// IEEE-rounded and drags a slow-path branch into every element, which also stops the interleaving.)
H3_DEV double rcp_seed(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

// Division and square root without the compiler's slow-path call sites (subnormal / huge operands): reciprocal seed + one
// third-order step (relative error < 2^-58), quotient with one residual correction -- within 1 ulp of a / b for normal
// operands; b == 0 gives NaN instead of +-inf (such an episode has diverged and ends on the NaN check).
H3_DEV double div_nr(double a, double b) {
    double rc = rcp_seed(b);
    const double e = fma(-b, rc, 1.0);
    rc = fma(rc, fma(e, e, e), rc);
    const double q = a * rc;
    return fma(fma(-q, b, a), rc, q);
}

// sqrt: MUFU.RSQ64H seed, two coupled Newton steps on (g ~ sqrt a, h ~ 1 / (2 sqrt a)), one residual correction.
// Arguments below 1e-290 (incl. 0 and subnormals, which the seed flushes) give 0; negative arguments NaN.
H3_DEV double sqrt_nr(double a) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double g = a * r, h = 0.5 * r;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    g = fma(fma(-g, g, a), h, g);
    return (a < 1.0e-290 && a >= 0.0) ? 0.0 : g;
}

// atan2 with ONE division and no call site: the quotient min / max of the magnitudes is reduced to |t| <= tan(pi/8)
// through atan(q) = pi/4 + atan((q - 1) / (q + 1)) applied to numerator and denominator BEFORE dividing
// ((mn - mx) / (mn + mx)), then t + t z A(z); octant and sign fixed up by selects.  <= 3 ulp (tests/test_gpu_math.py);
// atan2(0, 0) = 0; NaN in, NaN out; infinite arguments give NaN (no live state is infinite).
H3_DEV double atan2_nr(double y, double x) {
    const double ax = ::fabs(x), ay = ::fabs(y);
    const bool swap = ay > ax;
    const double mx = swap ? ay : ax, mn = swap ? ax : ay;
    const bool hi = mn > 0.41421356237309503 * mx;
    double t = div_nr(hi ? mn - mx : mn, hi ? mn + mx : mx);
    t = (mx == 0.0) ? 0.0 : t;
    const double z = t * t;
    double p = H3_ATAN_C[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, z, H3_ATAN_C[i]);
    double a = fma(t * z, p, t);
    a = hi ? a + 0.78539816339744831 : a;
    a = swap ? 1.5707963267948966 - a : a;
    a = (x < 0.0) ? 3.1415926535897931 - a : a;
    return ::copysign(a, y);
}

// Full-range tanh from a 313-entry table in shared memory (2.5 KB per CTA, staged by tanh_table_stage()):
//   x0 = floor(16 |x|) / 16, d = |x| - x0 in [0, 1/16]      (exact: DFMA.RM with the 1.5 * 2^52 constant, one DFMA)
//   tanh(x0 + d) = t0 + (1 - t0^2) p / (1 + t0 p),  t0 = tanh(x0) (table),  p = tanh(d) = d + d z Q(z), z = d^2
// Every term is non-negative, so nothing cancels: 1.4 ulp measured (tests/test_gpu_math.py); 17 FP64-pipe instructions.
// (Round 2 started with two branches -- an 11th-degree polynomial below 0.55, 14 instructions, and the expm1-based
// E / (E + 2) beyond, 29 instructions, chosen per chunk of hidden units: warps whose lanes disagreed paid for both.)
// The division is one MUFU.RCP64H seed
// (relative error e0 < 2^-19) and one third-order step rc (1 + e0 + e0^2): the quotient only scales the correction
// term.  Beyond 19.5 the result is 1 to the last bit (t0 = 1, 1 - t0^2 = 0); NaN comes out as NaN (the table index is
// clamped, d is NaN).  32 lanes with unrelated arguments gather 32 different words: a few bank conflicts per call on
// the LSU, which these kernels leave mostly idle.
#define H3_TANHT_MAGIC 6755399441055744.0
__shared__ double h3_tanht[H3_TANHT_N];

// every thread of the CTA; contains __syncthreads()
H3_DEV void tanh_table_stage() {
    for (int i = threadIdx.x; i < H3_TANHT_N; i += blockDim.x) h3_tanht[i] = H3_TANHT_G[i];
    __syncthreads();
}

H3_DEV double tanh_tab(double x) {
    const double ax = ::fabs(x);
    const double xc = ax > H3_TANHT_XMAX ? H3_TANHT_XMAX : ax;          // NaN passes
    const double t = __fma_rd(xc, 16.0, H3_TANHT_MAGIC);                // floor(16 xc) in the low word of t
    const double d = fma(t - H3_TANHT_MAGIC, -0.0625, xc);
    const unsigned k = min((unsigned)__double2loint(t), (unsigned)(H3_TANHT_N - 1));
    const double t0 = h3_tanht[k];
    const double z = d * d;
    const double q = fma(fma(fma(H3_TANHT_C[3], z, H3_TANHT_C[2]), z, H3_TANHT_C[1]), z, H3_TANHT_C[0]);
    const double p = fma(d * z, q, d);
    const double S = fma(-t0, t0, 1.0);
    const double den = fma(t0, p, 1.0);                                 // in [1, 1.0625]
    double rc = rcp_seed(den);
    const double e0 = fma(-den, rc, 1.0);
    rc = fma(rc, fma(e0, e0, e0), rc);
    return ::copysign(fma(p * S, rc, t0), x);
}

// Table-free full-range tanh: E / (E + 2), E = expm1(2x) = 2^n expm1(r) + (2^n - 1), 2x = n ln2 + r -- one straight-line
// formula without cancellation for every x (2.4 ulp, 29 FP64-pipe instructions).  For kernels that are not bound by the
// FP64 pipe and are too short to stage the table: the per-tick kernels of the 6-DOF loop (MathExp below).
H3_DEV double tanh_exp(double x) {
    const double xc = x > 20.0 ? 20.0 : (x < -20.0 ? -20.0 : x);    // |tanh| == 1 to the last bit beyond 19.1; NaN passes
    const double y = xc + xc;
    const double t = fma(y, H3_LN2[2], H3_TANHT_MAGIC);             // n = rint(y / ln2) in the low word of t
    const double nf = t - H3_TANHT_MAGIC;
    double r = fma(nf, -H3_LN2[0], y);
    r = fma(nf, -H3_LN2[1], r);                                     // |r| <= ln2 / 2
    double q = H3_EXP_C[10];
#pragma unroll
    for (int i = 9; i >= 0; --i) q = fma(q, r, H3_EXP_C[i]);
    const double p = fma(r * r, q, r);                              // expm1(r)
    const int n = __double2loint(t);                                // in [-58, 58]; NaN: E is NaN whatever n is
    const double sc = __hiloint2double((n + 1023) << 20, 0);
    const double E = fma(sc, p, sc - 1.0);
    const double den = E + 2.0;                                     // in [1, 2^58]: no special cases in the division
    double rc = rcp_seed(den);
    rc = fma(rc, fma(-den, rc, 1.0), rc);
    rc = fma(rc, fma(-den, rc, 1.0), rc);
    const double q0 = E * rc;
    return fma(fma(-q0, den, E), rc, q0);
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

// Cody-Waite reduction by multiples of pi/2 (three FMA pieces of 53 bits each), polynomial pair, quadrant swap --
// straight-line code for EVERY argument: no polynomial-only branch for small angles (a warp whose lanes disagree pays for
// both sides; 5 instructions saved when they agree) and no out-of-line library fallback (a call site costs registers
// around it even when it is never taken).  q = rint(x * 2/pi) lands in the low word of x * 2/pi + 1.5 * 2^52 and is exact
// as a double below 2^51, so the pair is valid for |x| < 3.5e15 (2 ulp in the ranges the loops visit, tests/
// test_gpu_math.py); inf and NaN give NaN.  Angles of the closed loops are bounded by the physics: the pitch-angle
// termination (heli_models.py:210-213) ends an episode long before.
H3_DEV void sincos_f64(double x, double* s, double* c) {
    const double t = fma(x, H3_PIO2[3], 6755399441055744.0);
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

#ifndef H3_TANH_G
#define H3_TANH_G 10
#endif

// tanh_tab on elements [G0, min(G0 + H3_TANH_G, U)) of x, stage by stage so that the chains interleave, then the
// next group
template <int G0, int U>
H3_DEV void tanh_tab_group(const double (&x)[U], double (&y)[U]) {
    constexpr int G = (U - G0) < H3_TANH_G ? (U - G0) : H3_TANH_G;
    if constexpr (G > 0) {
        double d[G], t0[G], q[G], z[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const double ax = ::fabs(x[G0 + u]);
            const double xc = ax > H3_TANHT_XMAX ? H3_TANHT_XMAX : ax;
            const double t = __fma_rd(xc, 16.0, H3_TANHT_MAGIC);
            d[u] = fma(t - H3_TANHT_MAGIC, -0.0625, xc);
            t0[u] = h3_tanht[min((unsigned)__double2loint(t), (unsigned)(H3_TANHT_N - 1))];
            z[u] = d[u] * d[u];
            q[u] = H3_TANHT_C[3];
        }
#pragma unroll
        for (int i = 2; i >= 0; --i)
#pragma unroll
            for (int u = 0; u < G; ++u) q[u] = fma(q[u], z[u], H3_TANHT_C[i]);
        double p[G], den[G], rc[G], e0[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
            p[u] = fma(d[u] * z[u], q[u], d[u]);
            den[u] = fma(t0[u], p[u], 1.0);
            rc[u] = rcp_seed(den[u]);
        }
#pragma unroll
        for (int u = 0; u < G; ++u) e0[u] = fma(-den[u], rc[u], 1.0);
#pragma unroll
        for (int u = 0; u < G; ++u) rc[u] = fma(rc[u], fma(e0[u], e0[u], e0[u]), rc[u]);
#pragma unroll
        for (int u = 0; u < G; ++u)
            y[G0 + u] = ::copysign(fma(p[u] * fma(-t0[u], t0[u], 1.0), rc[u], t0[u]), x[G0 + u]);
        tanh_tab_group<G0 + G, U>(x, y);
    }
}

template <typename T>
struct Math;

template <>
struct Math<double> {
    // once per kernel, by every thread, before the first tanh: stages the table of tanh_tab (contains __syncthreads())
    static H3_DEV void prepare() { tanh_table_stage(); }
    static H3_DEV double tanh_(double x) { return tanh_tab(x); }
    // U independent arguments at once, stage by stage, so that consecutive DFMAs belong to different elements (a single
    // warp per scheduler then covers the DFMA latency by itself)
    template <int U>
    static H3_DEV void tanh_vec(const double (&x)[U], double (&y)[U]) {
        tanh_tab_group<0, U>(x, y);
    }
    // sin/cos of two independent angles (theta and the cyclic pitch): straight-line, the four polynomial chains interleave
    static H3_DEV void sincos2_(double a, double b, double* sa, double* ca, double* sb, double* cb) {
        sincos_f64(a, sa, ca);
        sincos_f64(b, sb, cb);
    }
    static H3_DEV double exp_(double x) { return ::exp(x); }
    static H3_DEV double sqrt_(double x) { return sqrt_nr(x); }
    static H3_DEV double div_(double a, double b) { return div_nr(a, b); }
    static H3_DEV double abs_(double x) { return ::fabs(x); }
    static H3_DEV void sincos_(double x, double* s, double* c) { sincos_f64(x, s, c); }
    // (IEEE division on purpose: exp overflows to inf -> 0; and with an inlined division the compiler predicates the
    // tanh / sigmoid choice of the final activation, i.e. evaluates both for every lane: -2 % in the 3-DOF kernels)
    static H3_DEV double sigmoid_(double x) { return 1.0 / (1.0 + ::exp(-x)); }
};

// fp64 math without the shared-memory table (tanh_exp): nothing to stage
struct MathExp : Math<double> {
    static H3_DEV void prepare() {}
    static H3_DEV double tanh_(double x) { return tanh_exp(x); }
    template <int U>
    static H3_DEV void tanh_vec(const double (&x)[U], double (&y)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) y[u] = tanh_exp(x[u]);
    }
};

template <>
struct Math<float> {
    static H3_DEV void prepare() {}
    static H3_DEV float tanh_(float x) { return ::tanhf(x); }
    template <int U>
    static H3_DEV void tanh_vec(const float (&x)[U], float (&y)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) y[u] = ::tanhf(x[u]);
    }
    static H3_DEV void sincos2_(float a, float b, float* sa, float* ca, float* sb, float* cb) {
        ::sincosf(a, sa, ca);
        ::sincosf(b, sb, cb);
    }
    static H3_DEV float exp_(float x) { return ::expf(x); }
    static H3_DEV float sqrt_(float x) { return ::sqrtf(x); }
    static H3_DEV float div_(float a, float b) { return a / b; }
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
