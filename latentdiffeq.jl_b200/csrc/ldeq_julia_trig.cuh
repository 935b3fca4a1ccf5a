// ldeq_julia_trig.cuh -- Base.sin / Base.cos for Float32 as Julia 1.8 computes them (base/special/trig.jl and
// rem_pio2.jl, a port of FreeBSD msun k_sinf.c / k_cosf.c / e_rem_pio2f.c) [3P, restated from the published algorithm].
//
// Used by the forward-dual pullback (ldeq_fwdsens.cuh), which restates the reference's ForwardDiffSensitivity solves
// literally: in a Float32 adaptive solve the last bit of sin(x) moves the scaled error estimate by ~1e-3 relative
// (sum_j btilde_j k_j cancels heavily), i.e. it decides accept/reject on borderline steps.  libdevice's sinf, glibc's
// sinf and Julia's sin(::Float32) disagree in the last bit for about 1 % of the arguments; the dual solves follow
// Julia: widen to Float64, reduce by multiples of pi/2 (exact subtraction of the Float64 constant up to 9pi/4, two-term
// Cody-Waite beyond), evaluate the msun kernels in Float64 with NO fused multiply-adds, round to Float32 once.  The CPU
// oracle (oracle/ldeq_oracle.cpp::jl_sinf) is the same arithmetic; tests assert bit equality on 2^22 arguments.
#pragma once

#include "ldeq_common.cuh"

namespace ldeq {

__device__ __forceinline__ double jl_sin_kernel(double x) {
    const double S1 = -0.16666666641626524, S2 = 0.008333329385889463, S3 = -0.00019839334836096632, S4 = 2.718311493989822e-06;
    const double z = __dmul_rn(x, x), w = __dmul_rn(z, z), r = __dadd_rn(S3, __dmul_rn(z, S4)), s = __dmul_rn(z, x);
    return __dadd_rn(__dadd_rn(x, __dmul_rn(s, __dadd_rn(S1, __dmul_rn(z, S2)))), __dmul_rn(__dmul_rn(s, w), r));
}
__device__ __forceinline__ double jl_cos_kernel(double x) {
    const double C0 = -0.499999997251031, C1 = 0.04166662332373906, C2 = -0.001388676377460993, C3 = 2.439044879627741e-05;
    const double z = __dmul_rn(x, x), w = __dmul_rn(z, z), r = __dadd_rn(C2, __dmul_rn(z, C3));
    return __dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(z, C0)), __dmul_rn(w, C1)), __dmul_rn(__dmul_rn(w, z), r));
}
// rem_pio2_kernel(x::Float32): quadrant (only n mod 4 is used) and the reduced argument; |x| < Float32(pi)/2 * 2^28
__device__ __forceinline__ int jl_rem_pio2f(float x, double* y) {
    const double PI = 3.141592653589793, pio2_1 = 1.57079631090164184570e+00, pio2_1t = 1.58932547735281966916e-08,
                 inv_pio2 = 6.36619772367581382433e-01;
    const double xd = (double)x, ax = fabs(xd);
    const bool pos = x > 0.0f;
    if (ax <= PI * 5 / 4) {
        if (ax <= PI * 3 / 4) { *y = pos ? __dadd_rn(xd, -(PI / 2)) : __dadd_rn(xd, PI / 2); return pos ? 1 : -1; }
        *y = pos ? __dadd_rn(xd, -PI) : __dadd_rn(xd, PI);
        return pos ? 2 : -2;
    }
    if (ax <= PI * 9 / 4) {
        if (ax <= PI * 7 / 4) { *y = pos ? __dadd_rn(xd, -(PI * 3 / 2)) : __dadd_rn(xd, PI * 3 / 2); return pos ? 3 : -3; }
        *y = pos ? __dadd_rn(xd, -(PI * 4 / 2)) : __dadd_rn(xd, PI * 4 / 2);
        return pos ? 4 : -4;
    }
    const double fn = rint(__dmul_rn(xd, inv_pio2));
    const double r = __dadd_rn(xd, -__dmul_rn(fn, pio2_1)), w = __dmul_rn(fn, pio2_1t);
    *y = __dadd_rn(r, -w);
    return (int)(long long)fn;
}
// sin and cos of one Float32 argument (ForwardDiff's sin(::Dual) evaluates both)
__device__ __forceinline__ void jl_sincosf(float x, float* s, float* c) {
    const float ax = fabsf(x);
    if (ax < 0.7853982f) {  // Float32(pi)/4: no reduction
        const double xd = (double)x;
        *s = ax < 0.00034526698f ? x : (float)jl_sin_kernel(xd);      // sqrt(eps(Float32))
        *c = ax < 0.00024414062f ? 1.0f : (float)jl_cos_kernel(xd);   // sqrt(eps(Float32)/2)
        return;
    }
    if (!(ax < 8.4e8f)) {  // Payne-Hanek territory in Julia (and NaN/Inf): Float64 libdevice, rounded once
        double sd, cd;
        sincos((double)x, &sd, &cd);
        *s = (float)sd;
        *c = (float)cd;
        return;
    }
    double y;
    const int n = jl_rem_pio2f(x, &y) & 3;
    const float sk = (float)jl_sin_kernel(y), ck = (float)jl_cos_kernel(y);
    *s = n == 0 ? sk : n == 1 ? ck : n == 2 ? -sk : -ck;
    *c = n == 0 ? ck : n == 1 ? -sk : n == 2 ? -ck : sk;
}
template <class S> __device__ __forceinline__ void s_sincos_julia(S x, S* s, S* c);
template <> __device__ __forceinline__ void s_sincos_julia<float>(float x, float* s, float* c) { jl_sincosf(x, s, c); }
template <> __device__ __forceinline__ void s_sincos_julia<double>(double x, double* s, double* c) { sincos(x, s, c); }

}  // namespace ldeq
