// ldeq_erk.cuh -- the other values of the diffeq struct's `solver` field: OrdinaryDiffEq's DP5, BS3
// and RK4 as table-driven FSAL explicit Runge-Kutta methods for the integrator kernels of ldeq_tsit5.cuh.
//
// The reference's diffeq structs carry `solver = Tsit5()` (examples/pendulum_friction-less/pendulum.jl:11,58) and hand
// it to `solve` (src/models/GOKU.jl:121); the field is the user's to set.  In OrdinaryDiffEq all three methods end a
// step with fsallast = f(u_{n+1}), which becomes k1 of the next step, so one stage loop serves them: NS stages
// (the FSAL stage included), row NS-1 of `a` is b.  A method supplies
//   a(j,l), c(j)     the tableau (exact rationals rounded once, like OrdinaryDiffEq's constant caches)
//   bt(j)            btilde = b - bhat of the embedded pair (error estimate; all zero for RK4 = fixed step only)
//   r(j,m)           dense output  u(t_n + Theta dt) = u_n + dt Theta sum_{m=0..3} Theta^m sum_j r(j,m) k_j
//                    -- the Horner form the kernels evaluate (one contraction per step, four FMAs per save point)
//   ORDER            the exponent of the initial-step formula and the source of the controller defaults
// Dense outputs as OrdinaryDiffEq defines them, expanded into r(j,m):
//   DP5     Hairer's dopri5 `contd5`:  P = K1 + (1-Th)(K2 + Th(K3 + (1-Th) K4)),  K1 = sum b_j k_j, K2 = k1 - K1,
//           K3 = K1 - k7 - K2, K4 = sum d_j k_j   =>  c0 = k1, c1 = 3K1 - 2k1 - k7 + K4, c2 = -2K1 + k1 + k7 - 2K4, c3 = K4
//   BS3/RK4 cubic Hermite on (u_n, k1, u_{n+1}, k_last): the same with K4 = 0.
// The compile-time loops below are fully unrolled; coefficient look-ups fold to immediates and zero entries vanish.
#pragma once

#include "ldeq_common.cuh"

namespace ldeq {

struct TabDP5 {
    static constexpr int NS = 7, ORDER = 5;
    __host__ __device__ static constexpr double a(int j, int l) {
        switch (j * 8 + l) {
            case 1 * 8 + 0: return 1.0 / 5;
            case 2 * 8 + 0: return 3.0 / 40;
            case 2 * 8 + 1: return 9.0 / 40;
            case 3 * 8 + 0: return 44.0 / 45;
            case 3 * 8 + 1: return -56.0 / 15;
            case 3 * 8 + 2: return 32.0 / 9;
            case 4 * 8 + 0: return 19372.0 / 6561;
            case 4 * 8 + 1: return -25360.0 / 2187;
            case 4 * 8 + 2: return 64448.0 / 6561;
            case 4 * 8 + 3: return -212.0 / 729;
            case 5 * 8 + 0: return 9017.0 / 3168;
            case 5 * 8 + 1: return -355.0 / 33;
            case 5 * 8 + 2: return 46732.0 / 5247;
            case 5 * 8 + 3: return 49.0 / 176;
            case 5 * 8 + 4: return -5103.0 / 18656;
            case 6 * 8 + 0: return 35.0 / 384;
            case 6 * 8 + 2: return 500.0 / 1113;
            case 6 * 8 + 3: return 125.0 / 192;
            case 6 * 8 + 4: return -2187.0 / 6784;
            case 6 * 8 + 5: return 11.0 / 84;
            default: return 0.0;
        }
    }
    __host__ __device__ static constexpr double c(int j) {
        return j == 1 ? 1.0 / 5 : j == 2 ? 3.0 / 10 : j == 3 ? 4.0 / 5 : j == 4 ? 8.0 / 9 : j >= 5 ? 1.0 : 0.0;
    }
    __host__ __device__ static constexpr double bt(int j) {
        return j == 0 ? 71.0 / 57600 : j == 2 ? -71.0 / 16695 : j == 3 ? 71.0 / 1920 : j == 4 ? -17253.0 / 339200
             : j == 5 ? 22.0 / 525 : j == 6 ? -1.0 / 40 : 0.0;
    }
    __host__ __device__ static constexpr double d(int j) {
        return j == 0 ? -12715105075.0 / 11282082432.0 : j == 2 ? 87487479700.0 / 32700410799.0
             : j == 3 ? -10690763975.0 / 1880347072.0 : j == 4 ? 701980252875.0 / 199316789632.0
             : j == 5 ? -1453857185.0 / 822651844.0 : j == 6 ? 69997945.0 / 29380423.0 : 0.0;
    }
};

struct TabBS3 {
    static constexpr int NS = 4, ORDER = 3;
    __host__ __device__ static constexpr double a(int j, int l) {
        switch (j * 8 + l) {
            case 1 * 8 + 0: return 1.0 / 2;
            case 2 * 8 + 1: return 3.0 / 4;
            case 3 * 8 + 0: return 2.0 / 9;
            case 3 * 8 + 1: return 1.0 / 3;
            case 3 * 8 + 2: return 4.0 / 9;
            default: return 0.0;
        }
    }
    __host__ __device__ static constexpr double c(int j) { return j == 1 ? 1.0 / 2 : j == 2 ? 3.0 / 4 : j == 3 ? 1.0 : 0.0; }
    __host__ __device__ static constexpr double bt(int j) {
        return j == 0 ? -5.0 / 72 : j == 1 ? 1.0 / 12 : j == 2 ? 1.0 / 9 : j == 3 ? -1.0 / 8 : 0.0;
    }
    __host__ __device__ static constexpr double d(int) { return 0.0; }
};

// the classical method as OrdinaryDiffEq runs it: four stages + fsallast = f(u_{n+1}).  No embedded pair: its adaptive
// mode in OrdinaryDiffEq is a defect-control estimate that is not restated, so the entry points accept RK4 with
// adaptive = 0 only (LDEQ_ERR_UNSUPPORTED otherwise).
struct TabRK4 {
    static constexpr int NS = 5, ORDER = 4;
    __host__ __device__ static constexpr double a(int j, int l) {
        switch (j * 8 + l) {
            case 1 * 8 + 0: return 1.0 / 2;
            case 2 * 8 + 1: return 1.0 / 2;
            case 3 * 8 + 2: return 1.0;
            case 4 * 8 + 0: return 1.0 / 6;
            case 4 * 8 + 1: return 1.0 / 3;
            case 4 * 8 + 2: return 1.0 / 3;
            case 4 * 8 + 3: return 1.0 / 6;
            default: return 0.0;
        }
    }
    __host__ __device__ static constexpr double c(int j) { return j == 1 || j == 2 ? 1.0 / 2 : j >= 3 ? 1.0 : 0.0; }
    __host__ __device__ static constexpr double bt(int) { return 0.0; }
    __host__ __device__ static constexpr double d(int) { return 0.0; }
};

// r(j,m) of the DP5 / Hermite dense outputs (header comment); b_j = a(NS-1, j), the FSAL stage has b = 0
template <class TB> __host__ __device__ constexpr double erk_r(int j, int m) {
    const double b = j < TB::NS - 1 ? TB::a(TB::NS - 1, j) : 0.0;
    const double first = j == 0 ? 1.0 : 0.0, last = j == TB::NS - 1 ? 1.0 : 0.0;
    return m == 0 ? first
         : m == 1 ? 3.0 * b - 2.0 * first - last + TB::d(j)
         : m == 2 ? -2.0 * b + first + last - 2.0 * TB::d(j)
                  : TB::d(j);
}

// The method interface of the integrator bodies (ldeq_tsit5.cuh: Tsit5M is the hand-unrolled Tsit5), generic version.
template <class TB> struct ErkM {
    static constexpr int NS = TB::NS, ORDER = TB::ORDER;
    __host__ __device__ static constexpr double a(int j, int l) { return TB::a(j, l); }
    __host__ __device__ static constexpr double c(int j) { return TB::c(j); }
    __host__ __device__ static constexpr double bt(int j) { return TB::bt(j); }
    __host__ __device__ static constexpr double r(int j, int m) { return erk_r<TB>(j, m); }

    // k[0] holds f(u) on entry (FSAL); fills k[1..NS-1] and un.  KEEP (reverse pass): also the stage inputs g[0..NS-2]
    // and the right-hand side's auxiliaries; k[NS-1] = f(u_{n+1}) is then left to the caller (see tsit5_stages).
    template <class RHS, class S, bool KEEP, bool SAFE, bool PACK>
    __device__ __forceinline__ static void stages(const S* u, const S* p, double t, double dts, S (*k)[RHS::ZD], S* un,
                                                  S (*g)[RHS::ZD], typename RHS::Aux* aux) {
        constexpr int ZD = RHS::ZD;
        using V = VecOps<S, ZD, PACK>;
        const S h = (S)dts;
        S gi[ZD], acc[ZD];
        if constexpr (KEEP) {
#pragma unroll
            for (int i = 0; i < ZD; ++i) g[0][i] = u[i];
            RHS::template f<SAFE>(k[0], u, p, t, aux[0]);
        }
#pragma unroll
        for (int j = 1; j < NS; ++j) {
#pragma unroll
            for (int i = 0; i < ZD; ++i) acc[i] = (S)0;
#pragma unroll
            for (int l = 0; l < j; ++l)
                if (TB::a(j, l) != 0.0) V::axpy(acc, (S)TB::a(j, l), k[l]);
            if (j < NS - 1) {
                V::fma(gi, h, acc, u);
                if constexpr (KEEP) {
#pragma unroll
                    for (int i = 0; i < ZD; ++i) g[j][i] = gi[i];
                    RHS::template f<SAFE>(k[j], gi, p, t + TB::c(j) * dts, aux[j]);
                } else {
                    RHS::template f<SAFE>(k[j], gi, p, t + TB::c(j) * dts);
                }
            } else {
                V::fma(un, h, acc, u);
                if constexpr (!KEEP) RHS::template f<SAFE>(k[j], un, p, t + dts);
            }
        }
    }

    // scaled RMS error estimate of the embedded pair
    template <class S, int ZD>
    __device__ __forceinline__ static double eest(const S* u, const S* un, S (*k)[ZD], double dts, S abstol, S reltol) {
        using V = VecOps<S, ZD, false>;
        const S h = (S)dts;
        S s[ZD];
#pragma unroll
        for (int i = 0; i < ZD; ++i) s[i] = (S)0;
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (TB::bt(j) != 0.0) V::axpy(s, (S)TB::bt(j), k[j]);
        S e2 = (S)0;
#pragma unroll
        for (int i = 0; i < ZD; ++i) {
            const S utilde = h * s[i];
            const S sk = s_fma<S>(s_max<S>(s_abs<S>(u[i]), s_abs<S>(un[i])), reltol, abstol);
            const S q = s_div_fast<S>(utilde, sk);
            e2 = s_fma<S>(q, q, e2);
        }
        return (double)s_sqrt<S>(e2 / (S)ZD);
    }

    // Horner coefficients of the dense output, c_m = sum_j r(j,m) k_j, and the adjoint of that contraction
    template <class S, int ZD> __device__ __forceinline__ static void interp_coeffs(S (*k)[ZD], S (*c)[ZD]) {
        using V = VecOps<S, ZD, false>;
#pragma unroll
        for (int i = 0; i < ZD; ++i) c[0][i] = k[0][i];
#pragma unroll
        for (int m = 1; m < 4; ++m) {
#pragma unroll
            for (int i = 0; i < ZD; ++i) c[m][i] = (S)0;
#pragma unroll
            for (int j = 0; j < NS; ++j)
                if (erk_r<TB>(j, m) != 0.0) V::axpy(c[m], (S)erk_r<TB>(j, m), k[j]);
        }
    }
    template <class S, int ZD> __device__ __forceinline__ static void interp_coeffs_adj(S (*cb)[ZD], S (*kbar)[ZD]) {
        using V = VecOps<S, ZD>;
        V::add(kbar[0], cb[0]);
#pragma unroll
        for (int m = 1; m < 4; ++m)
#pragma unroll
            for (int j = 0; j < NS; ++j)
                if (erk_r<TB>(j, m) != 0.0) V::axpy(kbar[j], (S)erk_r<TB>(j, m), cb[m]);
    }

    // Reverse sweep through the stages of one taped step (discrete adjoint): on entry kbar holds the cotangents of the
    // slopes from the dense output, ubn the cotangent of u_{n+1}, ub the direct cotangent of u_n from the save points;
    // on exit ub = cotangent of u_n (complete), pbar accumulated.
    template <class RHS, class S>
    __device__ __forceinline__ static void reverse_sweep(const S* p, double tn, double dtn, S (*g)[RHS::ZD], const S* un,
                                                         typename RHS::Aux* aux, const typename RHS::Aux& aux_next,
                                                         S (*kbar)[RHS::ZD], S* ubn, S* ub, S* pbar) {
        constexpr int ZD = RHS::ZD;
        using V = VecOps<S, ZD>;
        const S h = (S)dtn;
        S v[ZD], gb[ZD];
        // k_NS = f(u_{n+1}) (it is also the next step's k1, whose adjoint was already folded into ubn)
        RHS::vjp(ubn, pbar, un, p, tn + dtn, kbar[NS - 1], aux_next);
        V::scale(v, h, ubn);
        V::add(ub, ubn);
#pragma unroll
        for (int l = 0; l < NS - 1; ++l)
            if (TB::a(NS - 1, l) != 0.0) V::axpy(kbar[l], (S)TB::a(NS - 1, l), v);
#pragma unroll
        for (int j = NS - 2; j >= 1; --j) {
#pragma unroll
            for (int i = 0; i < ZD; ++i) gb[i] = (S)0;
            RHS::vjp(gb, pbar, g[j], p, tn + TB::c(j) * dtn, kbar[j], aux[j]);
            V::scale(v, h, gb);
            V::add(ub, gb);
#pragma unroll
            for (int l = 0; l < j; ++l)
                if (TB::a(j, l) != 0.0) V::axpy(kbar[l], (S)TB::a(j, l), v);
        }
        RHS::vjp(ub, pbar, g[0], p, tn, kbar[0], aux[0]);
    }
};

using DP5M = ErkM<TabDP5>;
using BS3M = ErkM<TabBS3>;
using RK4M = ErkM<TabRK4>;

}  // namespace ldeq
