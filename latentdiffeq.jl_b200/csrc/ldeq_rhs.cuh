// ldeq_rhs.cuh -- right-hand sides the integrator kernels are instantiated with.
//
// An RHS is a struct with
//   static constexpr int ZD, PD;                       state / parameter dimension
//   struct Aux;                                        whatever f wants to keep for the VJP
//   static void f(S* du, const S* u, const S* p, double t, Aux& aux);
//   static void vjp(S* ubar, S* pbar, const S* u, const S* p, double t, const S* kbar, const Aux& aux);
//       ubar += (df/du)^T kbar ; pbar += (df/dp)^T kbar
//
// The two built-ins are the reference's example diffeq structs
// (examples/pendulum_friction-less/pendulum.jl:19-26 and :65-74).
#pragma once

#include "ldeq_common.cuh"

namespace ldeq {

template <class S, bool FRICTION> struct PendulumRHS {
    static constexpr int ZD = 2;
    static constexpr int PD = 1;
    struct Aux {
        S s, c;  // sin x, cos x of the stage input
    };
    // G = 10f0, b = 0.7f0, m = 1f0 are Float32 literals in the reference
    static constexpr S G = (S)10.0f;
    static constexpr S BM = (S)0.7f / (S)1.0f;

    // SAFE = false: the 13/19-instruction sine / cosine without their range test (the integrator tests the step's end
    // points once, fast_ok, and redoes the step with SAFE = true -- libdevice -- if one of them is out of range)
    template <bool SAFE> __device__ __forceinline__ static void sc(S x, S* s, S* c) {
        if constexpr (SAFE) s_sincos_fast<S>(x, s, c); else s_sincos_unchecked<S>(x, s, c);
    }
    template <bool SAFE> __device__ __forceinline__ static S sn(S x) {
        if constexpr (SAFE) return s_sin_fast<S>(x); else return s_sin_unchecked<S>(x);
    }
    __device__ __forceinline__ static bool fast_ok(const S* u) { return sizeof(S) == 8 || s_abs<S>(u[0]) <= (S)LDEQ_SINCOS_FAST_MAX; }

    template <bool SAFE = true>
    __device__ __forceinline__ static void f(S* du, const S* u, const S* p, double, Aux& aux) {
        sc<SAFE>(u[0], &aux.s, &aux.c);
        const S w = -G / p[0];
        du[0] = u[1];
        du[1] = FRICTION ? s_fma<S>(w, aux.s, -BM * u[1]) : w * aux.s;
    }
    // forward-only variant (no cosine needed)
    template <bool SAFE = true>
    __device__ __forceinline__ static void f(S* du, const S* u, const S* p, double) {
        const S w = -G / p[0];
        du[0] = u[1];
        const S sx = sn<SAFE>(u[0]);
        du[1] = FRICTION ? s_fma<S>(w, sx, -BM * u[1]) : w * sx;
    }
    __device__ __forceinline__ static void vjp(S* ubar, S* pbar, const S*, const S* p, double, const S* kbar,
                                               const Aux& aux) {
        const S w = -G / p[0];
        // J = [[0, 1], [w cos x, -b/m]];  df2/dL = (G/L^2) sin x
        ubar[0] = s_fma<S>(w * aux.c, kbar[1], ubar[0]);
        ubar[1] += FRICTION ? kbar[0] - BM * kbar[1] : kbar[0];
        pbar[0] = s_fma<S>((G / (p[0] * p[0])) * aux.s, kbar[1], pbar[0]);
    }
};

}  // namespace ldeq
