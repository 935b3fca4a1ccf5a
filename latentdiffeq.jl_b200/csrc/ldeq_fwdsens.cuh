// ldeq_fwdsens.cuh -- the reference's OWN gradient algorithm for the GOKU solve: sensealg = ForwardDiffSensitivity()
// (examples/pendulum_friction-less/pendulum.jl:11,58; SciMLSensitivity 7.10 `_concrete_solve_adjoint`).
//
// Pullback per trajectory: seed theta with Dual partials and re-solve, dtheta = sum_k (du_k/dtheta)^T Delta_k; seed u0
// and re-solve, dz0 likewise.  The error norm of a Dual state includes the partials (sse(value) + sum sse(partials),
// partials counted as entries), so each dual solve takes its OWN accepted-step sequence -- this is what makes the
// reference's gradient differ from the discrete adjoint of the primal steps (tsit5_bwd_kernel) by up to the solver
// tolerance.  The body restates that algorithm literally, including Julia's arithmetic promotion (a Float32 state
// meets the Float64 dt: `uprev + dt*(...)` is evaluated in Float64 and rounded on the store; the dense-output
// polynomials are evaluated in Float64), so that its step sequences are the oracle's (oracle/ldeq_oracle.cpp::solve_one
// with NP > 0).  One thread per trajectory, two launches (theta-seeded NP = p_dim, u0-seeded NP = z_dim); the
// cotangent is consumed on the fly, nothing is stored.  RHS: ZD, PD and  template <class D> f(D* du, const D* u,
// const D* p, double t)  on duals -- the built-in pendulums (ldeq_fwdsens.cu) or a user's NVRTC function.
#pragma once

#include "ldeq_dual.cuh"

namespace ldeq {

#define LDEQ_DL _Pragma("unroll") for (int i_ = 0; i_ < N; ++i_)
template <class S, int N> __device__ __forceinline__ Dual<S, N> dual_fma(S s, Dual<S, N> k, Dual<S, N> acc) {
    acc.v = s_fma<S>(s, k.v, acc.v);
    LDEQ_DL acc.d[i_] = s_fma<S>(s, k.d[i_], acc.d[i_]);
    return acc;
}
template <class S, int N> __device__ __forceinline__ Dual<S, N> dual_scale(S a, Dual<S, N> b) {
    b.v *= a;
    LDEQ_DL b.d[i_] *= a;
    return b;
}
// uprev + dt*sum with dt in Float64: promoted, fused, rounded back to S on the store
template <class S, int N> __device__ __forceinline__ Dual<S, N> dual_axpy_time(Dual<S, N> uprev, double dt, Dual<S, N> sum) {
    Dual<S, N> r;
    r.v = (S)fma(dt, (double)sum.v, (double)uprev.v);
    LDEQ_DL r.d[i_] = (S)fma(dt, (double)sum.d[i_], (double)uprev.d[i_]);
    return r;
}
template <class S, int N> __device__ __forceinline__ Dual<S, N> dual_scale_time(double dt, Dual<S, N> sum) {
    Dual<S, N> r;
    r.v = (S)(dt * (double)sum.v);
    LDEQ_DL r.d[i_] = (S)(dt * (double)sum.d[i_]);
    return r;
}
template <class S, int N> __device__ __forceinline__ S dual_sse(Dual<S, N> a, bool wp) {
    S s = a.v * a.v;
    if (wp) { LDEQ_DL s += a.d[i_] * a.d[i_]; }
    return s;
}
template <class S, int N> __device__ __forceinline__ S dual_absnorm(Dual<S, N> a, bool wp) { return wp ? s_sqrt<S>(dual_sse(a, true)) : s_abs<S>(a.v); }
template <class S, int N> __device__ __forceinline__ Dual<S, N> dual_div_s(Dual<S, N> a, S s) {
    a.v /= s;
    LDEQ_DL a.d[i_] /= s;
    return a;
}
template <class S, int N> __device__ __forceinline__ S dual_rms(const Dual<S, N>* a, int n, bool wp) {
    S s = (S)0;
    for (int i = 0; i < n; ++i) s += dual_sse(a[i], wp);
    const int len = wp ? n * (1 + N) : n;
    return s_sqrt<S>(s / (S)len);
}
#undef LDEQ_DL

// SEED_P: partials seeded on theta (NP = PD) -> dout = dtheta (p,B);  else on u0 (NP = ZD) -> dout = dz0 (z,B)
template <class RHS, class S, int NP, bool SEED_P>
__device__ __forceinline__ void
tsit5_fwdsens_body(const S* __restrict__ z0, const S* __restrict__ theta, const double* __restrict__ tg, int B, int T,
                   KOpts o, int norm_partials, const S* __restrict__ dtraj, const int* __restrict__ primal_ret,
                   S* __restrict__ dout) {
    constexpr int Z = RHS::ZD, PD = RHS::PD;
    using D = Dual<S, NP>;
    using Tb = Tab<S>;
    using Td = Tab<double>;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const bool wp = norm_partials != 0;
    D u[Z], k[7][Z], unew[Z], tmp[Z], sum, L[PD];
    for (int i = 0; i < Z; ++i) {
        u[i] = D(z0[(size_t)b * Z + i]);
        if (!SEED_P) u[i].d[i] = (S)1;
    }
    for (int i = 0; i < PD; ++i) {
        L[i] = D(theta[(size_t)b * PD + i]);
        if (SEED_P) L[i].d[i] = (S)1;
    }

    const double t0 = tg[0], tend = tg[T - 1];
    const double dtmax = o.dtmax > 0.0 ? o.dtmax : (tend - t0);
    const double dtmin = o.dtmin > 0.0 ? o.dtmin : fmax(2.220446049250313e-16, ulp_of(t0));
    const S abstol = (S)o.abstol, reltol = (S)o.reltol;
    double acc[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) acc[q] = 0.0;
    // save point 0 is u0 itself
    for (int i = 0; i < Z; ++i)
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] += (double)u[i].d[q] * (double)dtraj[(size_t)b * Z + i];

    RHS::f(k[0], u, L, t0);  // fsalfirst
    double t = t0, dt;
    if (o.adaptive && !(o.dt > 0.0)) {
        // Hairer initial step on the dual state
        S sk[Z];
        for (int i = 0; i < Z; ++i) {
            sk[i] = abstol + dual_absnorm(u[i], wp) * reltol;
            tmp[i] = dual_div_s(u[i], sk[i]);
        }
        const double d0 = (double)dual_rms(tmp, Z, wp);
        for (int i = 0; i < Z; ++i) tmp[i] = dual_div_s(k[0][i], sk[i]);
        const double d1 = (double)dual_rms(tmp, Z, wp);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = fmin(dt0, dtmax);
        if (dt0 < 10.0 * 2.220446049250313e-16) {
            dt = fmax(1e-6, dtmin);
        } else {
            D u1[Z], f1[Z];
            for (int i = 0; i < Z; ++i) u1[i] = dual_axpy_time(u[i], dt0, k[0][i]);
            RHS::f(f1, u1, L, t0 + dt0);
            for (int i = 0; i < Z; ++i) tmp[i] = dual_div_s(f1[i] - k[0][i], sk[i]);
            const double d2 = (double)dual_rms(tmp, Z, wp) / dt0;
            const double m = fmax(d1, d2);
            const double dt1 = (m <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : ::pow(10.0, -(2.0 + log10(m)) / 5.0);
            dt = fmax(dtmin, fmin(100.0 * dt0, fmin(dt1, dtmax)));
        }
    } else {
        dt = o.dt;
    }
    PiState pst = pi_init(o);
    int ks = 1, ret = RET_SUCCESS;
    long long iters = 0;
    if (!(dt > 0.0) || !isfinite(dt)) ret = RET_DTLESSTHANMIN;
    while (ks < T && ret == RET_SUCCESS) {
        if (iters >= o.maxiters) { ret = RET_MAXITERS; break; }
        ++iters;
        const double dts = fmin(dt, tend - t);
        double tnew = t + dts;
        if (::fabs(tnew - tend) < 100.0 * ulp_of(fmax(::fabs(t), ::fabs(tend)))) tnew = tend;
        // ---- one Tsit5 step on duals ----
        for (int i = 0; i < Z; ++i) { sum = dual_scale(Tb::a21, k[0][i]); tmp[i] = dual_axpy_time(u[i], dts, sum); }
        RHS::f(k[1], tmp, L, t + Tb::c2 * dts);
        for (int i = 0; i < Z; ++i) {
            sum = dual_scale(Tb::a31, k[0][i]); sum = dual_fma(Tb::a32, k[1][i], sum);
            tmp[i] = dual_axpy_time(u[i], dts, sum);
        }
        RHS::f(k[2], tmp, L, t + Tb::c3 * dts);
        for (int i = 0; i < Z; ++i) {
            sum = dual_scale(Tb::a41, k[0][i]); sum = dual_fma(Tb::a42, k[1][i], sum); sum = dual_fma(Tb::a43, k[2][i], sum);
            tmp[i] = dual_axpy_time(u[i], dts, sum);
        }
        RHS::f(k[3], tmp, L, t + Tb::c4 * dts);
        for (int i = 0; i < Z; ++i) {
            sum = dual_scale(Tb::a51, k[0][i]); sum = dual_fma(Tb::a52, k[1][i], sum); sum = dual_fma(Tb::a53, k[2][i], sum);
            sum = dual_fma(Tb::a54, k[3][i], sum);
            tmp[i] = dual_axpy_time(u[i], dts, sum);
        }
        RHS::f(k[4], tmp, L, t + Tb::c5 * dts);
        for (int i = 0; i < Z; ++i) {
            sum = dual_scale(Tb::a61, k[0][i]); sum = dual_fma(Tb::a62, k[1][i], sum); sum = dual_fma(Tb::a63, k[2][i], sum);
            sum = dual_fma(Tb::a64, k[3][i], sum); sum = dual_fma(Tb::a65, k[4][i], sum);
            tmp[i] = dual_axpy_time(u[i], dts, sum);
        }
        RHS::f(k[5], tmp, L, t + dts);
        for (int i = 0; i < Z; ++i) {
            sum = dual_scale(Tb::a71, k[0][i]); sum = dual_fma(Tb::a72, k[1][i], sum); sum = dual_fma(Tb::a73, k[2][i], sum);
            sum = dual_fma(Tb::a74, k[3][i], sum); sum = dual_fma(Tb::a75, k[4][i], sum); sum = dual_fma(Tb::a76, k[5][i], sum);
            unew[i] = dual_axpy_time(u[i], dts, sum);
        }
        RHS::f(k[6], unew, L, t + dts);
        double EEst = 0.0;
        if (o.adaptive) {
            for (int i = 0; i < Z; ++i) {
                sum = dual_scale(Tb::bt1, k[0][i]); sum = dual_fma(Tb::bt2, k[1][i], sum); sum = dual_fma(Tb::bt3, k[2][i], sum);
                sum = dual_fma(Tb::bt4, k[3][i], sum); sum = dual_fma(Tb::bt5, k[4][i], sum); sum = dual_fma(Tb::bt6, k[5][i], sum);
                sum = dual_fma(Tb::bt7, k[6][i], sum);
                const D ut = dual_scale_time(dts, sum);
                const S a0 = dual_absnorm(u[i], wp), a1 = dual_absnorm(unew[i], wp);
                const S sk = abstol + (a0 > a1 ? a0 : a1) * reltol;
                tmp[i] = dual_div_s(ut, sk);
            }
            EEst = (double)dual_rms(tmp, Z, wp);
        }
        bool finite = true;
        for (int i = 0; i < Z; ++i) finite = finite && s_finite<S>(unew[i].v);
        if (!finite || EEst != EEst) { ret = RET_UNSTABLE; break; }
        bool accept = true;
        double dt_next = dt;
        if (o.adaptive) accept = pi_controller(o, EEst, dts, dtmax, pst, dt_next);
        if (accept) {
            // saveat: every pending time <= tnew, dense output evaluated in Float64 (Theta is Float64)
            while (ks < T && tg[ks] <= tnew) {
                const double tsv = tg[ks];
                D out[Z];
                if (tsv == tnew) {
                    for (int i = 0; i < Z; ++i) out[i] = unew[i];
                } else {
                    const double T1 = (tsv - t) / dts, T2 = T1 * T1;
                    double bb[7];
                    bb[0] = T1 * (Td::r11 + T1 * (Td::r12 + T1 * (Td::r13 + T1 * Td::r14)));
                    bb[1] = T2 * (Td::r22 + T1 * (Td::r23 + T1 * Td::r24));
                    bb[2] = T2 * (Td::r32 + T1 * (Td::r33 + T1 * Td::r34));
                    bb[3] = T2 * (Td::r42 + T1 * (Td::r43 + T1 * Td::r44));
                    bb[4] = T2 * (Td::r52 + T1 * (Td::r53 + T1 * Td::r54));
                    bb[5] = T2 * (Td::r62 + T1 * (Td::r63 + T1 * Td::r64));
                    bb[6] = T2 * (Td::r72 + T1 * (Td::r73 + T1 * Td::r74));
                    for (int i = 0; i < Z; ++i) {
                        double sv = 0.0, sd[NP];
#pragma unroll
                        for (int q = 0; q < NP; ++q) sd[q] = 0.0;
#pragma unroll
                        for (int j = 0; j < 7; ++j) {
                            sv = fma(bb[j], (double)k[j][i].v, sv);
#pragma unroll
                            for (int q = 0; q < NP; ++q) sd[q] = fma(bb[j], (double)k[j][i].d[q], sd[q]);
                        }
                        out[i].v = (S)fma(dts, sv, (double)u[i].v);
#pragma unroll
                        for (int q = 0; q < NP; ++q) out[i].d[q] = (S)fma(dts, sd[q], (double)u[i].d[q]);
                    }
                }
                for (int i = 0; i < Z; ++i) {
                    const double dv = (double)dtraj[((size_t)ks * B + b) * Z + i];
#pragma unroll
                    for (int q = 0; q < NP; ++q) acc[q] += (double)out[i].d[q] * dv;
                }
                ++ks;
            }
            t = tnew;
            for (int i = 0; i < Z; ++i) { u[i] = unew[i]; k[0][i] = k[6][i]; }
        }
        if (o.adaptive) dt = dt_next;
        if (ks < T && o.adaptive && (!(::fabs(dt) > dtmin) || !isfinite(dt))) { ret = RET_DTLESSTHANMIN; break; }
    }
    // a failed solve (dual or primal) contributes nothing: its NaN block is a constant of the differentiation (GOKU.jl:114)
    const bool ok = ret == RET_SUCCESS && primal_ret[b] == RET_SUCCESS;
#pragma unroll
    for (int q = 0; q < NP; ++q) dout[(size_t)b * NP + q] = ok ? (S)acc[q] : (S)0;
}

}  // namespace ldeq
