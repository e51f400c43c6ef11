"""Generates msc-thesis_b200/csrc/h3_coef.cuh: the constants of the kernels' own fp64 tanh / sin / cos.

sin / cos: polynomial pair on |x| <= pi/4 (flight angles; beyond, Cody-Waite reduction by pi/2 in three pieces).
tanh: table of tanh(k / 16), k = 0 .. 312, and the polynomial of tanh on [0, 1/16] (h3_math.cuh: tanh_tab); exp on
|r| <= ln2 / 2 and ln2 in two pieces for the table-free formula the 6-DOF kernels use (tanh_exp).

Everything is computed here from first principles with mpmath (Chebyshev interpolation of the reduced function at
60 digits, converted to the monomial basis and rounded to double; table entries correctly rounded), not taken from
any library.  Polynomial coefficients go to __constant__ memory so that every DFMA reads its coefficient as a
constant-bank operand; the table is a __device__ array that the kernels stage into shared memory.

    python tools/fit_math_coef.py          # rewrites the header and prints the measured accuracy
"""
import os
import struct

import mpmath as mp
import numpy as np

mp.mp.dps = 60
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "msc-thesis_b200", "csrc",
                   "h3_coef.cuh")


def cheb_fit(f, a, b, deg):
    """Monomial coefficients (lowest first) of the degree-`deg` Chebyshev interpolant of f on [a, b]."""
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    ys = [f(x) for x in xs]
    # solve the Vandermonde system in high precision (n <= 14: perfectly fine at 60 digits)
    A = mp.matrix(n, n)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
    c = mp.lu_solve(A, mp.matrix(ys))
    return [c[i] for i in range(n)]


def f_tanh(z):      # tanh(x) = x + x z P(z), z = x^2
    if z == 0:
        return mp.mpf(-1) / 3
    x = mp.sqrt(z)
    return (mp.tanh(x) / x - 1) / z


def f_sin(z):       # sin(x) = x + x z S(z)
    if z == 0:
        return mp.mpf(-1) / 6
    x = mp.sqrt(z)
    return (mp.sin(x) / x - 1) / z


def f_cos(z):       # cos(x) = 1 - z/2 + z^2 C(z)
    if z == 0:
        return mp.mpf(1) / 24
    x = mp.sqrt(z)
    return (mp.cos(x) - 1 + z / 2) / (z * z)


def f_exp(r):       # exp(r) = 1 + r + r^2 Q(r)
    if abs(r) < mp.mpf("1e-12"):        # the middle Chebyshev node is ~1e-61: use the series there
        return mp.mpf(1) / 2 + r / 6 + r * r / 24
    return (mp.exp(r) - 1 - r) / (r * r)


def f_atan(z):      # atan(t) = t + t z A(z), z = t^2
    if z == 0:
        return mp.mpf(-1) / 3
    t = mp.sqrt(z)
    return (mp.atan(t) / t - 1) / z


def horner(coef, z):
    acc = np.float64(coef[-1])
    for c in coef[-2::-1]:
        acc = acc * z + np.float64(c)
    return acc


def ulp_err(approx, exact_fn, xs):
    worst = 0.0
    for x, y in zip(xs, approx):
        e = exact_fn(mp.mpf(float(x)))
        if e == 0:
            continue
        ulp = mp.mpf(2) ** (mp.floor(mp.log(abs(e), 2)) - 52)
        worst = max(worst, float(abs(mp.mpf(float(y)) - e) / ulp))
    return worst


def main():
    rng = np.random.default_rng(0)
    spec = []
    # sin / cos, |x| <= pi/4
    zmax = (mp.pi / 4) ** 2
    cs = [float(c) for c in cheb_fit(f_sin, mp.mpf(0), zmax, 7)]
    cc = [float(c) for c in cheb_fit(f_cos, mp.mpf(0), zmax, 7)]
    xs = np.concatenate([rng.uniform(-np.pi / 4, np.pi / 4, 4000), rng.uniform(-1e-3, 1e-3, 1000)])
    z = xs * xs
    ys = xs + xs * (z * horner(cs, z))
    yc = (1.0 - 0.5 * z) + (z * z) * horner(cc, z)
    spec.append(("H3_SIN_C", cs, ulp_err(ys, mp.sin, xs)))
    spec.append(("H3_COS_C", cc, ulp_err(yc, mp.cos, xs)))
    # exp(r), |r| <= ln2/2: the table-free tanh of the 6-DOF kernels (h3_math.cuh: tanh_exp)
    half = mp.log(2) / 2
    ce = [float(c) for c in cheb_fit(f_exp, -half, half, 10)]
    xs = rng.uniform(-float(half), float(half), 5000)
    ye = 1.0 + (xs + (xs * xs) * horner(ce, xs))
    spec.append(("H3_EXP_C", ce, ulp_err(ye, mp.exp, xs)))
    ln2 = mp.log(2)
    ln2_hi = float(ln2)
    ln2_lo = float(ln2 - mp.mpf(ln2_hi))
    spec.append(("H3_LN2", [ln2_hi, ln2_lo, float(1 / ln2)], 0.0))
    # atan on |t| <= tan(pi/8) (h3_math.cuh: atan2_nr reduces every quotient to this range with ONE division)
    tmax = mp.tan(mp.pi / 8)
    ca = [float(c) for c in cheb_fit(f_atan, mp.mpf(0), tmax ** 2, 11)]
    xs = rng.uniform(-float(tmax), float(tmax), 5000)
    z = xs * xs
    spec.append(("H3_ATAN_C", ca, ulp_err(xs + xs * (z * horner(ca, z)), mp.atan, xs)))
    # pi/2 in three double pieces (Cody-Waite reduction with FMAs) and 2/pi
    pio2 = mp.pi / 2
    p1 = float(pio2)
    p2 = float(pio2 - mp.mpf(p1))
    p3 = float(pio2 - mp.mpf(p1) - mp.mpf(p2))
    spec.append(("H3_PIO2", [p1, p2, p3, float(2 / mp.pi)], 0.0))
    # table-driven full-range tanh (h3_math.cuh: tanh_tab): x0 = floor(16 |x|) / 16, d = |x| - x0 in [0, 1/16],
    # tanh(x0 + d) = t0 + (1 - t0^2) p / (1 + t0 p), t0 = tanh(x0) from a table, p = tanh(d) = d + d z Q(z)
    tab_h, tab_xmax = mp.mpf(1) / 16, mp.mpf("19.5")
    cq = [float(c) for c in cheb_fit(f_tanh, mp.mpf(0), tab_h ** 2, 3)]
    xs = rng.uniform(0.0, float(tab_h), 5000)
    z = xs * xs
    spec.append(("H3_TANHT_C", cq, ulp_err(xs + xs * (z * horner(cq, z)), mp.tanh, xs)))
    tab = [float(mp.tanh(k * tab_h)) for k in range(int(tab_xmax / tab_h) + 1)]
    assert tab[-1] == 1.0 and tab[0] == 0.0
    lines = ["// GENERATED by tools/fit_math_coef.py -- do not edit.  Coefficients (lowest order first) of the reduced",
             "// functions  sin(x) = x + x z S(z),  cos(x) = 1 - z/2 + z^2 C(z),  z = x^2, |x| <= pi/4, and of",
             "// tanh(d) = d + d z Q(z), z = d^2, 0 <= d <= 1/16, with the table of tanh(k / 16) (h3_math.cuh: tanh_tab).",
             "// Hex = IEEE-754 bits of the double.",
             "#pragma once", ""]
    for name, coef, err in spec:
        lines.append("// max error measured against mpmath on 5000 points (plain Horner in double): %.2f ulp" % err)
        lines.append("static __constant__ double %s[%d] = {" % (name, len(coef)))
        for c in coef:
            bits = struct.unpack("<Q", struct.pack("<d", c))[0]
            lines.append("    %s,  // 0x%016x" % (repr(c), bits))
        lines.append("};")
        lines.append("")
        print(name, "degree", len(coef) - 1, "max err %.2f ulp" % err)
    lines.append("// tanh(k / 16), k = 0 .. %d, correctly rounded (mpmath, 60 digits); staged into shared memory by the kernels" % (len(tab) - 1))
    lines.append("#define H3_TANHT_N %d" % len(tab))
    lines.append("#define H3_TANHT_XMAX %s" % repr(float(tab_xmax)))
    lines.append("static __device__ const double H3_TANHT_G[H3_TANHT_N] = {")
    for i in range(0, len(tab), 4):
        lines.append("    " + " ".join("%s," % float(t).hex() for t in tab[i:i + 4]))
    lines.append("};")
    lines.append("")
    with open(OUT, "w") as f:
        f.write("\n".join(lines))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
