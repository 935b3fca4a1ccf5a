// ldeq_dual.cuh -- forward-mode dual numbers and the adapter that turns a user-defined right-hand side
//   template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t);
// into the RHS interface of the integrator kernels (ldeq_rhs.cuh).  The reference lets the user write any
// `f!(du, u, p, t)` in the diffeq struct (examples/pendulum_friction-less/pendulum.jl:19-26); here the same
// function is written once in CUDA C and instantiated with float, double and Dual<S, ZD+PD>: for the 2-3
// dimensional systems of this model family the full Jacobian by forward mode (one dual evaluation) is the
// cheapest VJP.
#pragma once

#include "ldeq_common.cuh"
#include "ldeq_julia_trig.cuh"

namespace ldeq {

template <class S, int N> struct Dual {
    typedef S value_type;
    S v;
    S d[N];
    __device__ __forceinline__ Dual() {}
    __device__ __forceinline__ Dual(S x) : v(x) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = (S)0;
    }
    __device__ __forceinline__ Dual(int x) : v((S)x) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = (S)0;
    }
    template <class U, class = decltype((S)U())>
    __device__ __forceinline__ Dual(U x) : v((S)x) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = (S)0;
    }
};

#define LDEQ_DUAL_T template <class S, int N> __device__ __forceinline__
LDEQ_DUAL_T Dual<S, N> operator+(Dual<S, N> a, const Dual<S, N>& b) { a.v += b.v; _Pragma("unroll") for (int i = 0; i < N; ++i) a.d[i] += b.d[i]; return a; }
LDEQ_DUAL_T Dual<S, N> operator-(Dual<S, N> a, const Dual<S, N>& b) { a.v -= b.v; _Pragma("unroll") for (int i = 0; i < N; ++i) a.d[i] -= b.d[i]; return a; }
LDEQ_DUAL_T Dual<S, N> operator-(Dual<S, N> a) { a.v = -a.v; _Pragma("unroll") for (int i = 0; i < N; ++i) a.d[i] = -a.d[i]; return a; }
LDEQ_DUAL_T Dual<S, N> operator+(Dual<S, N> a) { return a; }
LDEQ_DUAL_T Dual<S, N> operator*(const Dual<S, N>& a, const Dual<S, N>& b) {
    Dual<S, N> r; r.v = a.v * b.v;
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
LDEQ_DUAL_T Dual<S, N> operator/(const Dual<S, N>& a, const Dual<S, N>& b) {
    // ForwardDiff's quotient rule with true divisions (the forward-dual pullback restates the reference literally)
    Dual<S, N> r; r.v = a.v / b.v;
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
    return r;
}
// mixed with plain scalars (anything convertible to S)
#define LDEQ_DUAL_MIXED(OP)                                                                                          \
    template <class S, int N, class U, class = decltype((S)U())>                                                    \
    __device__ __forceinline__ Dual<S, N> operator OP(const Dual<S, N>& a, U b) { return a OP Dual<S, N>((S)b); }    \
    template <class S, int N, class U, class = decltype((S)U())>                                                    \
    __device__ __forceinline__ Dual<S, N> operator OP(U a, const Dual<S, N>& b) { return Dual<S, N>((S)a) OP b; }
LDEQ_DUAL_MIXED(+)
LDEQ_DUAL_MIXED(-)
LDEQ_DUAL_MIXED(*)
LDEQ_DUAL_MIXED(/)
#undef LDEQ_DUAL_MIXED
LDEQ_DUAL_T Dual<S, N>& operator+=(Dual<S, N>& a, const Dual<S, N>& b) { a = a + b; return a; }
LDEQ_DUAL_T Dual<S, N>& operator-=(Dual<S, N>& a, const Dual<S, N>& b) { a = a - b; return a; }
LDEQ_DUAL_T Dual<S, N>& operator*=(Dual<S, N>& a, const Dual<S, N>& b) { a = a * b; return a; }
LDEQ_DUAL_T Dual<S, N>& operator/=(Dual<S, N>& a, const Dual<S, N>& b) { a = a / b; return a; }
LDEQ_DUAL_T bool operator<(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v < b.v; }
LDEQ_DUAL_T bool operator>(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v > b.v; }
LDEQ_DUAL_T bool operator<=(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v <= b.v; }
LDEQ_DUAL_T bool operator>=(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v >= b.v; }

// chain rule helper: f(a) with derivative fp
LDEQ_DUAL_T Dual<S, N> dual_chain(const Dual<S, N>& a, S f, S fp) {
    Dual<S, N> r; r.v = f;
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = fp * a.d[i];
    return r;
}
#undef LDEQ_DUAL_T

}  // namespace ldeq

// elementary functions, found by argument-dependent lookup from user code written as sin(x), exp(x), ...
// sin / cos of a Float32 dual follow Julia's Base.sin / Base.cos bit for bit (ldeq_julia_trig.cuh): the forward-dual
// pullback restates the reference's dual-number solves literally
namespace ldeq {
template <class S, int N> __device__ __forceinline__ Dual<S, N> sin(const Dual<S, N>& a) { S s, c; s_sincos_julia<S>(a.v, &s, &c); return dual_chain(a, s, c); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> cos(const Dual<S, N>& a) { S s, c; s_sincos_julia<S>(a.v, &s, &c); return dual_chain(a, c, -s); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> tan(const Dual<S, N>& a) { const S t = ::tan(a.v); return dual_chain(a, t, (S)1 + t * t); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> exp(const Dual<S, N>& a) { const S e = ::exp(a.v); return dual_chain(a, e, e); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> log(const Dual<S, N>& a) { return dual_chain(a, (S)::log(a.v), (S)1 / a.v); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> sqrt(const Dual<S, N>& a) { const S r = ::sqrt(a.v); return dual_chain(a, r, (S)0.5 / r); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> tanh(const Dual<S, N>& a) { const S t = ::tanh(a.v); return dual_chain(a, t, (S)1 - t * t); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> sinh(const Dual<S, N>& a) { return dual_chain(a, (S)::sinh(a.v), (S)::cosh(a.v)); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> cosh(const Dual<S, N>& a) { return dual_chain(a, (S)::cosh(a.v), (S)::sinh(a.v)); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> atan(const Dual<S, N>& a) { return dual_chain(a, (S)::atan(a.v), (S)1 / ((S)1 + a.v * a.v)); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> fabs(const Dual<S, N>& a) { return dual_chain(a, (S)::fabs(a.v), a.v < (S)0 ? (S)-1 : (S)1); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> abs(const Dual<S, N>& a) { return fabs(a); }
template <class S, int N, class U> __device__ __forceinline__ Dual<S, N> pow(const Dual<S, N>& a, U e) {
    const S ev = (S)e;
    return dual_chain(a, (S)::pow(a.v, ev), ev * (S)::pow(a.v, ev - (S)1));
}
template <class S, int N> __device__ __forceinline__ Dual<S, N> pow(const Dual<S, N>& a, const Dual<S, N>& e) { return exp(e * log(a)); }

// ---- the adapter ---------------------------------------------------------------------------------------------
template <class S, int ZD_, int PD_> struct UserRHS;

}  // namespace ldeq

// the user's function lives in the global namespace
template <class S> __device__ void ldeq_user_rhs(S* du, const S* u, const S* p, S t);

namespace ldeq {

template <class S, int ZD_, int PD_> struct UserRHS {
    static constexpr int ZD = ZD_;
    static constexpr int PD = PD_;
    struct Aux {};
    __device__ __forceinline__ static bool fast_ok(const S*) { return true; }  // no reduced-range fast path to guard
    template <bool SAFE = true>
    __device__ __forceinline__ static void f(S* du, const S* u, const S* p, double t) { ::ldeq_user_rhs<S>(du, u, p, (S)t); }
    template <bool SAFE = true>
    __device__ __forceinline__ static void f(S* du, const S* u, const S* p, double t, Aux&) { ::ldeq_user_rhs<S>(du, u, p, (S)t); }
    // ubar += (df/du)^T kbar, pbar += (df/dp)^T kbar through one forward-mode evaluation with ZD+PD directions
    __device__ __forceinline__ static void vjp(S* ubar, S* pbar, const S* u, const S* p, double t, const S* kbar,
                                               const Aux&) {
        using D = Dual<S, ZD + PD>;
        D ud[ZD], pd[PD > 0 ? PD : 1], dud[ZD];
#pragma unroll
        for (int i = 0; i < ZD; ++i) { ud[i] = D(u[i]); ud[i].d[i] = (S)1; }
#pragma unroll
        for (int i = 0; i < PD; ++i) { pd[i] = D(p[i]); pd[i].d[ZD + i] = (S)1; }
        ::ldeq_user_rhs<D>(dud, ud, pd, D((S)t));
#pragma unroll
        for (int r = 0; r < ZD; ++r) {
#pragma unroll
            for (int i = 0; i < ZD; ++i) ubar[i] = s_fma<S>(dud[r].d[i], kbar[r], ubar[i]);
#pragma unroll
            for (int i = 0; i < PD; ++i) pbar[i] = s_fma<S>(dud[r].d[ZD + i], kbar[r], pbar[i]);
        }
    }
};

}  // namespace ldeq
