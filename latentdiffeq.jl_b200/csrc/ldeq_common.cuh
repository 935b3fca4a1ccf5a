// ldeq_common.cuh -- shared device code for the Tsit5 integrator kernels (sm_100a).
//
// Algorithm notes (what is computed, not how): OrdinaryDiffEq's Tsit5 with FSAL, RMS error norm,
// PI step-size controller with DiffEqBase.fastpow, Hairer initial step and saveat through the
// dense interpolant -- the semantics `solve(ens_prob, Tsit5(), EnsembleThreads(); saveat = t, ...)`
// has at reference src/models/GOKU.jl:121.
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
// NVRTC (user-defined right-hand sides are compiled at run time): no host headers
typedef unsigned int uint32_t;
#endif

namespace ldeq {

// ---- Tsit5 tableau, rounded to the state type S at compile time --------------------------------
template <class S> struct Tab {
    static constexpr double c2 = 0.161, c3 = 0.327, c4 = 0.9, c5 = 0.9800255409045097;
    static constexpr S a21 = (S)0.161;
    static constexpr S a31 = (S)-0.008480655492356989, a32 = (S)0.335480655492357;
    static constexpr S a41 = (S)2.8971530571054935, a42 = (S)-6.359448489975075, a43 = (S)4.3622954328695815;
    static constexpr S a51 = (S)5.325864828439257, a52 = (S)-11.748883564062828, a53 = (S)7.4955393428898365,
                       a54 = (S)-0.09249506636175525;
    static constexpr S a61 = (S)5.86145544294642, a62 = (S)-12.92096931784711, a63 = (S)8.159367898576159,
                       a64 = (S)-0.071584973281401, a65 = (S)-0.028269050394068383;
    static constexpr S a71 = (S)0.09646076681806523, a72 = (S)0.01, a73 = (S)0.4798896504144996,
                       a74 = (S)1.379008574103742, a75 = (S)-3.290069515436081, a76 = (S)2.324710524099774;
    static constexpr S bt1 = (S)-0.00178001105222577714, bt2 = (S)-0.0008164344596567469,
                       bt3 = (S)0.007880878010261995, bt4 = (S)-0.1447110071732629, bt5 = (S)0.5823571654525552,
                       bt6 = (S)-0.45808210592918697, bt7 = (S)0.015151515151515152;
    // dense output b_j(Theta)
    static constexpr S r11 = (S)1.0, r12 = (S)-2.763706197274826, r13 = (S)2.9132554618219126,
                       r14 = (S)-1.0530884977290216;
    static constexpr S r22 = (S)0.13169999999999998, r23 = (S)-0.2234, r24 = (S)0.1017;
    static constexpr S r32 = (S)3.9302962368947516, r33 = (S)-5.941033872131505, r34 = (S)2.490627285651253;
    static constexpr S r42 = (S)-12.411077166933676, r43 = (S)30.33818863028232, r44 = (S)-16.548102889244902;
    static constexpr S r52 = (S)37.50931341651104, r53 = (S)-88.1789048947664, r54 = (S)47.37952196281928;
    static constexpr S r62 = (S)-27.896526289197286, r63 = (S)65.09189467479366, r64 = (S)-34.87065786149661;
    static constexpr S r72 = (S)1.5, r73 = (S)-4.0, r74 = (S)2.5;
};

// dense-output weights b_1..b_7 at Theta
template <class S> __device__ __forceinline__ void interp_weights(S th, S* bw) {
    using Tb = Tab<S>;
    const S th2 = th * th;
    bw[0] = th * (Tb::r11 + th * (Tb::r12 + th * (Tb::r13 + th * Tb::r14)));
    bw[1] = th2 * (Tb::r22 + th * (Tb::r23 + th * Tb::r24));
    bw[2] = th2 * (Tb::r32 + th * (Tb::r33 + th * Tb::r34));
    bw[3] = th2 * (Tb::r42 + th * (Tb::r43 + th * Tb::r44));
    bw[4] = th2 * (Tb::r52 + th * (Tb::r53 + th * Tb::r54));
    bw[5] = th2 * (Tb::r62 + th * (Tb::r63 + th * Tb::r64));
    bw[6] = th2 * (Tb::r72 + th * (Tb::r73 + th * Tb::r74));
}

// ---- solver options as the kernels see them ----------------------------------------------------
struct KOpts {
    double abstol, reltol;
    double dt, dtmax, dtmin;
    double gamma, qmin, qmax, beta1, beta2, qoldinit, qsteady_min, qsteady_max;
    long long maxiters;
    int adaptive;
    int controller_pow;
};

// the save grid as the kernels see it: t0 + k h when the host has verified that bit for bit (uniform), and the row
// stride of the (z,B,T) arrays (a kernel may work on a column slab of a wider batch)
struct GridInfo {
    double t0, h;
    int uniform;
    int ld;  // row stride (in trajectories) of the (z,B,T) arrays: the kernel may work on a column slab of a wider batch
    int sort;  // reverse pass: re-deal the CTA's trajectories to lanes by their accepted-step count (0: batch order)
};

enum { RET_SUCCESS = 0, RET_MAXITERS = 1, RET_DTLESSTHANMIN = 2, RET_UNSTABLE = 3 };

// ---- DiffEqBase.fastpow: Float32 rational log2 on the significand, Float32 exp2 ------------------
__device__ __forceinline__ float fastlog2_dev(float x) {
    const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
    const uint32_t ux = __float_as_uint(x);
    const int ex = (int)((ux & 0x7F800000u) >> 23);
    const bool greater = (ux & 0x00400000u) != 0u;
    float signif = __uint_as_float((ux & 0x007FFFFFu) | (greater ? 0x3f000000u : 0x3f800000u));
    const float fexp = (float)(ex - (greater ? 126 : 127));
    signif = signif - 1.0f;
    return fexp + __fdiv_rn(__fmul_rn(signif, __fadd_rn(__fmul_rn(a, signif), b)), __fadd_rn(signif, c));
}
// DiffEqBase.fastpow(x, y) for Float64 arguments: demote to Float32, exp2(y * fastlog2(x))
__device__ __forceinline__ float ctrl_powf(double x, float y, int exact) {
    if (exact) return (float)pow(x, (double)y);
    if (x == 0.0) return 0.0f;
    return exp2f(__fmul_rn(y, fastlog2_dev(fabsf((float)x))));
}

// PI controller (OrdinaryDiffEq stepsize_controller!/step_accept_controller!/step_reject_controller!).
// Returns accept; updates the controller memory and the next step-size proposal.  The two fastpow factors are
// Float32 by construction; their quotient and the final 1/q are taken in Float32 as well (a relative 1e-7 on the
// proposed dt, far below anything the error control can see) so that no Float64 division is issued.
// The memory is kept as qold_pow = fastpow(qold, beta2): after an accepted step qold = max(EEst, qoldinit), so the
// next step's denominator is exp2(beta2 * fastlog2(EEst)) from the SAME logarithm as this step's numerator
// (bit-identical to evaluating fastpow(qold, beta2) then, one logarithm per step instead of two).
struct PiState {
    float qold_pow;  // fastpow(qold, beta2)
    // loop invariants, converted to Float32 once per trajectory (they were re-derived from the Float64 options on every
    // step: 10 F2F + 2 MUFU per step on the quarter-rate XU pipe)
    float inv_gamma, qlo, qhi, beta1, beta2, qsteady_min, qsteady_max, qoldinit, lg_qoldinit, q22_init;
    int exact_pow;
};
__device__ __forceinline__ PiState pi_init(const KOpts& o) {
    PiState st;
    st.inv_gamma = (float)(1.0 / o.gamma);
    st.qlo = (float)(1.0 / o.qmax);
    st.qhi = (float)(1.0 / o.qmin);
    st.beta1 = (float)o.beta1;
    st.beta2 = (float)o.beta2;
    st.qsteady_min = (float)o.qsteady_min;
    st.qsteady_max = (float)o.qsteady_max;
    st.qoldinit = (float)o.qoldinit;
    st.lg_qoldinit = fastlog2_dev(st.qoldinit);
    st.exact_pow = o.controller_pow;
    st.q22_init = ctrl_powf(o.qoldinit, (float)o.beta2, o.controller_pow);
    st.qold_pow = st.q22_init;
    return st;
}
__device__ __forceinline__ bool pi_controller(const KOpts& o, double EEst, double dts, double dtmax, PiState& st,
                                              double& dt_next) {
    float q, q11 = 1.0f, q22_next = st.qold_pow;
    if (EEst == 0.0) {
        q = st.qlo;
        q22_next = st.q22_init;
    } else {
        if (st.exact_pow) {
            q11 = (float)pow(EEst, o.beta1);
            q22_next = (float)pow(fmax(EEst, o.qoldinit), o.beta2);
        } else {
            const float ef = fabsf((float)EEst);
            const float lg = fastlog2_dev(ef);
            q11 = exp2f(__fmul_rn(st.beta1, lg));
            // fastpow(max(EEst, qoldinit), beta2): the Float32 demotion commutes with max
            q22_next = exp2f(__fmul_rn(st.beta2, ef >= st.qoldinit ? lg : st.lg_qoldinit));
        }
        q = __fdiv_rn(q11, st.qold_pow);
        q = fmaxf(st.qlo, fminf(st.qhi, q * st.inv_gamma));
    }
    const bool accept = EEst <= 1.0;
    if (accept) {
        if (st.qsteady_min <= q && q <= st.qsteady_max) q = 1.0f;
        st.qold_pow = q22_next;
        dt_next = fmin(dtmax, dts * (double)__frcp_rn(q));
    } else {
        dt_next = dts * (double)__frcp_rn(fminf(st.qhi, q11 * st.inv_gamma));
    }
    return accept;
}

// ulp(max(|a|,|b|)) for the tstop snap (fixed_t_for_floatingpoint_error!)
__device__ __forceinline__ double ulp_of(double x) {
    x = fabs(x);
    return __longlong_as_double(__double_as_longlong(x) + 1) - x;
}

// ---- small typed helpers -------------------------------------------------------------------------
template <class S> __device__ __forceinline__ S s_fma(S a, S b, S c);
template <> __device__ __forceinline__ float s_fma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double s_fma<double>(double a, double b, double c) { return fma(a, b, c); }
template <class S> __device__ __forceinline__ S s_abs(S a);
template <> __device__ __forceinline__ float s_abs<float>(float a) { return fabsf(a); }
template <> __device__ __forceinline__ double s_abs<double>(double a) { return fabs(a); }
template <class S> __device__ __forceinline__ S s_sqrt(S a);
template <> __device__ __forceinline__ float s_sqrt<float>(float a) { return sqrtf(a); }
template <> __device__ __forceinline__ double s_sqrt<double>(double a) { return sqrt(a); }
template <class S> __device__ __forceinline__ S s_max(S a, S b);
template <> __device__ __forceinline__ float s_max<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double s_max<double>(double a, double b) { return fmax(a, b); }
template <class S> __device__ __forceinline__ void s_sincos(S x, S* s, S* c);
template <> __device__ __forceinline__ void s_sincos<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <> __device__ __forceinline__ void s_sincos<double>(double x, double* s, double* c) { sincos(x, s, c); }
template <class S> __device__ __forceinline__ S s_sin(S x);
template <> __device__ __forceinline__ float s_sin<float>(float x) { return sinf(x); }
template <> __device__ __forceinline__ double s_sin<double>(double x) { return sin(x); }
template <class S> __device__ __forceinline__ S s_nan();
template <> __device__ __forceinline__ float s_nan<float>() { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double s_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <class S> __device__ __forceinline__ bool s_finite(S a);
template <> __device__ __forceinline__ bool s_finite<float>(float a) { return isfinite(a); }
template <> __device__ __forceinline__ bool s_finite<double>(double a) { return isfinite(a); }

// ---- Float32 sine / cosine of the pendulum right-hand sides ------------------------------------------
// sinf / sincosf of libdevice cost ~25 / ~40 instructions (quadrant selection, a Payne-Hanek slow path behind a branch);
// the integrator kernels are instruction-issue bound and evaluate 6-7 of them per step.  These versions reduce by
// multiples of pi with a three-term Cody-Waite split (k pi_hi exact in the FMA), then ONE odd minimax polynomial of
// degree 9 on [-pi/2, pi/2] (approximation error 6e-9 relative) and, for the cosine, an even one of degree 10
// (2.4e-10 absolute): 13 / 19 instructions.  Measured against Float64 (tests/test_goku_gpu.py::test_fast_sincos_accuracy):
// sine <= 1.8 ulp everywhere on |x| <= 1e4 (mean 0.33 ulp, 99.1 % of the arguments within 1 ulp), cosine <= 2e-7 absolute.
// Arguments beyond 1e4 (never reached by a pendulum angle in radians, but legal) take libdevice's sinf / sincosf.
#define LDEQ_SINCOS_FAST_MAX 1.0e4f
__device__ __forceinline__ float fast_sin_reduce(float x, unsigned& sign) {
    const float kf = fmaf(x, 0.318309886183790672f, 12582912.0f);  // 1.5 * 2^23: the integer lands in the low mantissa bits
    sign = __float_as_uint(kf) << 31;
    const float k = kf - 12582912.0f;
    float r = fmaf(k, -3.14159274101257324f, x);
    r = fmaf(k, 8.74227765734758577e-08f, r);
    return fmaf(k, 3.43024902734575342e-15f, r);
}
__device__ __forceinline__ float fast_sin_poly(float r, float r2) {
    float p = fmaf(r2, 2.6057520614159343e-06f, -1.9809590781843947e-04f);
    p = fmaf(r2, p, 8.3330660931473012e-03f);
    p = fmaf(r2, p, -1.6666659545045038e-01f);
    return fmaf(r * r2, p, r);
}
__device__ __forceinline__ float fast_cos_poly(float r2) {
    float q = fmaf(r2, -2.6076832445861094e-07f, 2.4761871604948579e-05f);
    q = fmaf(r2, q, -1.3888403241539062e-03f);
    q = fmaf(r2, q, 4.1666640709316493e-02f);
    q = fmaf(r2, q, -4.9999999549160074e-01f);
    return fmaf(r2, q, 1.0f);
}
// out of line: six or seven inlined copies of libdevice's slow path would triple the kernels' code size
static __device__ __noinline__ float slow_sinf(float x) { return sinf(x); }
static __device__ __noinline__ void slow_sincosf(float x, float* s, float* c) { sincosf(x, s, c); }
// The integrators call the unchecked forms: the range test is made ONCE per step on the step's end points (a step whose
// u_n or u_{n+1} leaves the fast range is redone with libdevice's functions, see tsit5_stages' SAFE flag); the reduction
// itself stays valid far beyond the tested range (the magic-number rounding up to |x| < 2^22 pi).
__device__ __forceinline__ float fast_sinf_unchecked(float x) {
    unsigned sg;
    const float r = fast_sin_reduce(x, sg);
    return __uint_as_float(__float_as_uint(fast_sin_poly(r, r * r)) ^ sg);
}
__device__ __forceinline__ void fast_sincosf_unchecked(float x, float* s, float* c) {
    unsigned sg;
    const float r = fast_sin_reduce(x, sg);
    const float r2 = r * r;
    *s = __uint_as_float(__float_as_uint(fast_sin_poly(r, r2)) ^ sg);
    *c = __uint_as_float(__float_as_uint(fast_cos_poly(r2)) ^ sg);
}
__device__ __forceinline__ float fast_sinf(float x) {
    if (!(fabsf(x) <= LDEQ_SINCOS_FAST_MAX)) return slow_sinf(x);
    unsigned sg;
    const float r = fast_sin_reduce(x, sg);
    return __uint_as_float(__float_as_uint(fast_sin_poly(r, r * r)) ^ sg);
}
__device__ __forceinline__ void fast_sincosf(float x, float* s, float* c) {
    if (!(fabsf(x) <= LDEQ_SINCOS_FAST_MAX)) { slow_sincosf(x, s, c); return; }
    unsigned sg;
    const float r = fast_sin_reduce(x, sg);
    const float r2 = r * r;
    *s = __uint_as_float(__float_as_uint(fast_sin_poly(r, r2)) ^ sg);
    *c = __uint_as_float(__float_as_uint(fast_cos_poly(r2)) ^ sg);
}
template <class S> __device__ __forceinline__ S s_sin_fast(S x);
template <> __device__ __forceinline__ float s_sin_fast<float>(float x) { return fast_sinf(x); }
template <> __device__ __forceinline__ double s_sin_fast<double>(double x) { return sin(x); }
template <class S> __device__ __forceinline__ void s_sincos_fast(S x, S* s, S* c);
template <> __device__ __forceinline__ void s_sincos_fast<float>(float x, float* s, float* c) { fast_sincosf(x, s, c); }
template <> __device__ __forceinline__ void s_sincos_fast<double>(double x, double* s, double* c) { sincos(x, s, c); }
template <class S> __device__ __forceinline__ S s_sin_unchecked(S x);
template <> __device__ __forceinline__ float s_sin_unchecked<float>(float x) { return fast_sinf_unchecked(x); }
template <> __device__ __forceinline__ double s_sin_unchecked<double>(double x) { return sin(x); }
template <class S> __device__ __forceinline__ void s_sincos_unchecked(S x, S* s, S* c);
template <> __device__ __forceinline__ void s_sincos_unchecked<float>(float x, float* s, float* c) { fast_sincosf_unchecked(x, s, c); }
template <> __device__ __forceinline__ void s_sincos_unchecked<double>(double x, double* s, double* c) { sincos(x, s, c); }
// quotient for the scaled error estimate: MUFU.RCP + one multiply in Float32 (2 ulp; the estimate itself carries the
// cancellation noise of sum_j btilde_j k_j, ~1e-3 relative), IEEE in Float64
template <class S> __device__ __forceinline__ S s_div_fast(S a, S b);
template <> __device__ __forceinline__ float s_div_fast<float>(float a, float b) { return __fdividef(a, b); }
template <> __device__ __forceinline__ double s_div_fast<double>(double a, double b) { return a / b; }

// vector load/store of one trajectory's ZD-vector (coalesced 8/16-byte accesses for ZD = 2)
template <class S, int ZD> __device__ __forceinline__ void load_vec(const S* __restrict__ p, S* v) {
#pragma unroll
    for (int i = 0; i < ZD; ++i) v[i] = p[i];
}
template <> __device__ __forceinline__ void load_vec<float, 2>(const float* __restrict__ p, float* v) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void load_vec<double, 2>(const double* __restrict__ p, double* v) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x; v[1] = t.y;
}
template <class S, int ZD> __device__ __forceinline__ void store_vec(S* __restrict__ p, const S* v) {
#pragma unroll
    for (int i = 0; i < ZD; ++i) p[i] = v[i];
}
template <> __device__ __forceinline__ void store_vec<float, 2>(float* __restrict__ p, const float* v) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}
template <> __device__ __forceinline__ void store_vec<double, 2>(double* __restrict__ p, const double* v) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

// ---- ZD-vector arithmetic with a scalar coefficient -------------------------------------------------
// Every stage combination, dense-output coefficient, error estimate and adjoint update of the integrators is
// "scalar x ZD-vector (+ ZD-vector)".  For the two-dimensional Float32 state of the GOKU pendulums that is exactly one
// packed fma.rn.f32x2 (FFMA2 in SASS, sm_100): half the issue slots of two FFMAs, and the scalar rides along as an
// immediate / 32-bit register broadcast.  Measured at 2^20 x 200 (same box, A/B builds): the adjoint kernel gains 8 %
// (1.185 -> 1.097 ms: its reverse sweep is 21 + 21 vector updates per step with no scalar work in between), the forward
// kernel LOSES 3 % (1.240 -> 1.281 ms: pairing the operands costs register moves around the scalar sine chains), so the
// kernels choose per call site (PACK).  NVRTC (user right-hand sides) takes the generic loops.
template <class S, int N, bool PACK = true> struct VecOps {
    // out = a * x
    __device__ __forceinline__ static void scale(S* out, S a, const S* x) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = a * x[i];
    }
    // y += a * x
    __device__ __forceinline__ static void axpy(S* y, S a, const S* x) {
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = s_fma<S>(a, x[i], y[i]);
    }
    // out = a * x + y
    __device__ __forceinline__ static void fma(S* out, S a, const S* x, const S* y) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = s_fma<S>(a, x[i], y[i]);
    }
    // y += x
    __device__ __forceinline__ static void add(S* y, const S* x) {
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] += x[i];
    }
};
#if !defined(__CUDACC_RTC__) && !defined(LDEQ_NO_F32X2)
template <> struct VecOps<float, 2, true> {
    __device__ __forceinline__ static void scale(float* out, float a, const float* x) {
        const float2 r = __fmul2_rn(make_float2(a, a), make_float2(x[0], x[1]));
        out[0] = r.x; out[1] = r.y;
    }
    __device__ __forceinline__ static void axpy(float* y, float a, const float* x) {
        const float2 r = __ffma2_rn(make_float2(a, a), make_float2(x[0], x[1]), make_float2(y[0], y[1]));
        y[0] = r.x; y[1] = r.y;
    }
    __device__ __forceinline__ static void fma(float* out, float a, const float* x, const float* y) {
        const float2 r = __ffma2_rn(make_float2(a, a), make_float2(x[0], x[1]), make_float2(y[0], y[1]));
        out[0] = r.x; out[1] = r.y;
    }
    __device__ __forceinline__ static void add(float* y, const float* x) {
        const float2 r = __fadd2_rn(make_float2(y[0], y[1]), make_float2(x[0], x[1]));
        y[0] = r.x; y[1] = r.y;
    }
};
#endif

}  // namespace ldeq
