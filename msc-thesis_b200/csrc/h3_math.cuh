// Precision-dispatched device math for the 3-DOF tick kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace h3 {

#define H3_DEV __device__ __forceinline__

template <typename T>
struct Math;

template <>
struct Math<double> {
    static H3_DEV double tanh_(double x) { return ::tanh(x); }
    static H3_DEV double exp_(double x) { return ::exp(x); }
    static H3_DEV double sqrt_(double x) { return ::sqrt(x); }
    static H3_DEV double abs_(double x) { return ::fabs(x); }
    static H3_DEV void sincos_(double x, double* s, double* c) { ::sincos(x, s, c); }
    static H3_DEV double sigmoid_(double x) { return 1.0 / (1.0 + ::exp(-x)); }
};

template <>
struct Math<float> {
    static H3_DEV float tanh_(float x) { return ::tanhf(x); }
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
template <typename T>
H3_DEV T pick7(const T (&s)[7], int idx) {
    T r = s[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) r = (idx == i) ? s[i] : r;
    return r;
}

}  // namespace h3
