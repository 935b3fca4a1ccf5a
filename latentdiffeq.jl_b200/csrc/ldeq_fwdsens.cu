// ldeq_fwdsens.cu -- the reference's OWN gradient algorithm for the GOKU solve: sensealg = ForwardDiffSensitivity()
// (examples/pendulum_friction-less/pendulum.jl:11,58; SciMLSensitivity 7.10 `_concrete_solve_adjoint`).
//
// Pullback per trajectory: seed theta with Dual partials and re-solve, dtheta = sum_k (du_k/dtheta)^T Delta_k; seed u0
// and re-solve, dz0 likewise.  The error norm of a Dual state includes the partials (sse(value) + sum sse(partials),
// partials counted as entries), so each dual solve takes its OWN accepted-step sequence -- this is what makes the
// reference's gradient differ from the discrete adjoint of the primal steps (tsit5_bwd_kernel) by up to the solver
// tolerance.  This kernel restates that algorithm literally, including Julia's arithmetic promotion (a Float32 state
// meets the Float64 dt: `uprev + dt*(...)` is evaluated in Float64 and rounded on the store; the dense-output
// polynomials are evaluated in Float64), so that its step sequences are the oracle's (oracle/ldeq_oracle.cpp::solve_one
// with NP > 0).  One thread per trajectory, two launches (theta-seeded NP = p_dim, u0-seeded NP = z_dim); the
// cotangent is consumed on the fly, nothing is stored.  LDEQ_SENSE_FORWARD_DUAL selects it; the default reverse pass
// stays the discrete adjoint (5x cheaper, equal to this within the solver tolerance).
#include "ldeq_internal.h"

namespace ldeq {

template <class S, int NP> struct FD {
    S v;
    S d[NP];
};
#define FD_LOOP for (int i_ = 0; i_ < NP; ++i_)
template <class S, int NP> __device__ __forceinline__ FD<S, NP> fd_make(S v) {
    FD<S, NP> r;
    r.v = v;
#pragma unroll
    FD_LOOP r.d[i_] = (S)0;
    return r;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> operator-(FD<S, NP> a, FD<S, NP> b) {
    a.v -= b.v;
#pragma unroll
    FD_LOOP a.d[i_] -= b.d[i_];
    return a;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> operator-(FD<S, NP> a) {
    a.v = -a.v;
#pragma unroll
    FD_LOOP a.d[i_] = -a.d[i_];
    return a;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> operator*(FD<S, NP> a, FD<S, NP> b) {
    FD<S, NP> r;
    r.v = a.v * b.v;
#pragma unroll
    FD_LOOP r.d[i_] = a.d[i_] * b.v + a.v * b.d[i_];
    return r;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> operator*(S a, FD<S, NP> b) {
    b.v *= a;
#pragma unroll
    FD_LOOP b.d[i_] *= a;
    return b;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> operator/(FD<S, NP> a, FD<S, NP> b) {
    FD<S, NP> r;
    r.v = a.v / b.v;
#pragma unroll
    FD_LOOP r.d[i_] = (a.d[i_] - r.v * b.d[i_]) / b.v;
    return r;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> fd_sin(FD<S, NP> a) {
    FD<S, NP> r;
    S sn, cs;
    s_sincos<S>(a.v, &sn, &cs);
    r.v = sn;
#pragma unroll
    FD_LOOP r.d[i_] = cs * a.d[i_];
    return r;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> fd_fma(S s, FD<S, NP> k, FD<S, NP> acc) {
    acc.v = s_fma<S>(s, k.v, acc.v);
#pragma unroll
    FD_LOOP acc.d[i_] = s_fma<S>(s, k.d[i_], acc.d[i_]);
    return acc;
}
// uprev + dt*sum with dt in Float64: promoted, fused, rounded back to S on the store
template <class S, int NP> __device__ __forceinline__ FD<S, NP> fd_axpy_time(FD<S, NP> uprev, double dt, FD<S, NP> sum) {
    FD<S, NP> r;
    r.v = (S)fma(dt, (double)sum.v, (double)uprev.v);
#pragma unroll
    FD_LOOP r.d[i_] = (S)fma(dt, (double)sum.d[i_], (double)uprev.d[i_]);
    return r;
}
template <class S, int NP> __device__ __forceinline__ FD<S, NP> fd_scale_time(double dt, FD<S, NP> sum) {
    FD<S, NP> r;
    r.v = (S)(dt * (double)sum.v);
#pragma unroll
    FD_LOOP r.d[i_] = (S)(dt * (double)sum.d[i_]);
    return r;
}
template <class S, int NP> __device__ __forceinline__ S fd_sse(FD<S, NP> a, bool wp) {
    S s = a.v * a.v;
    if (wp) {
#pragma unroll
        FD_LOOP s += a.d[i_] * a.d[i_];
    }
    return s;
}
template <class S, int NP> __device__ __forceinline__ S fd_absnorm(FD<S, NP> a, bool wp) { return wp ? s_sqrt<S>(fd_sse(a, true)) : s_abs<S>(a.v); }
template <class S, int NP> __device__ __forceinline__ FD<S, NP> fd_div_s(FD<S, NP> a, S s) {
    a.v /= s;
#pragma unroll
    FD_LOOP a.d[i_] /= s;
    return a;
}
template <class S, int NP> __device__ __forceinline__ S fd_rms(const FD<S, NP>* a, int n, bool wp) {
    S s = (S)0;
    for (int i = 0; i < n; ++i) s += fd_sse(a[i], wp);
    const int len = wp ? n * (1 + NP) : n;
    return s_sqrt<S>(s / (S)len);
}

// pendulum.jl:19-26 / :65-74 on duals (G, b, m are Float32 literals of the reference, rounded to S)
// mgl = -G/L is a constant of the trajectory: evaluated once by the caller (same value as evaluating it per call)
template <class S, int NP, bool FRICTION>
__device__ __forceinline__ void fd_rhs(FD<S, NP>* du, const FD<S, NP>* u, FD<S, NP> mgl) {
    du[0] = u[1];
    const FD<S, NP> a = mgl * fd_sin(u[0]);
    if (FRICTION) {
        const S bm = (S)0.7f / (S)1.0f;
        du[1] = a - bm * u[1];
    } else {
        du[1] = a;
    }
}

// SEED_P: partials seeded on theta (NP = 1) -> dout = dtheta (B);  else on u0 (NP = 2) -> dout = dz0 (2,B)
template <class S, int NP, bool FRICTION, bool SEED_P>
__global__ void __launch_bounds__(128)
tsit5_fwdsens_kernel(const S* __restrict__ z0, const S* __restrict__ theta, const double* __restrict__ tg, int B, int T,
                     KOpts o, int norm_partials, const S* __restrict__ dtraj, const int* __restrict__ primal_ret,
                     S* __restrict__ dout) {
    constexpr int Z = 2;
    using D = FD<S, NP>;
    using Tb = Tab<S>;
    using Td = Tab<double>;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const bool wp = norm_partials != 0;
    D u[Z], k[7][Z], unew[Z], tmp[Z], sum;
    for (int i = 0; i < Z; ++i) {
        u[i] = fd_make<S, NP>(z0[(size_t)b * Z + i]);
        if (!SEED_P) u[i].d[i] = (S)1;
    }
    D L = fd_make<S, NP>(theta[b]);
    if (SEED_P) L.d[0] = (S)1;
    L = (-fd_make<S, NP>((S)10.0f)) / L;  // from here on L holds -G/L (pendulum.jl:24, :72)

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

    fd_rhs<S, NP, FRICTION>(k[0], u, L);  // fsalfirst
    double t = t0, dt;
    if (o.adaptive && !(o.dt > 0.0)) {
        // Hairer initial step on the dual state
        S sk[Z];
        for (int i = 0; i < Z; ++i) {
            sk[i] = abstol + fd_absnorm(u[i], wp) * reltol;
            tmp[i] = fd_div_s(u[i], sk[i]);
        }
        const double d0 = (double)fd_rms(tmp, Z, wp);
        for (int i = 0; i < Z; ++i) tmp[i] = fd_div_s(k[0][i], sk[i]);
        const double d1 = (double)fd_rms(tmp, Z, wp);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = fmin(dt0, dtmax);
        if (dt0 < 10.0 * 2.220446049250313e-16) {
            dt = fmax(1e-6, dtmin);
        } else {
            D u1[Z], f1[Z];
            for (int i = 0; i < Z; ++i) u1[i] = fd_axpy_time(u[i], dt0, k[0][i]);
            fd_rhs<S, NP, FRICTION>(f1, u1, L);
            for (int i = 0; i < Z; ++i) tmp[i] = fd_div_s(f1[i] - k[0][i], sk[i]);
            const double d2 = (double)fd_rms(tmp, Z, wp) / dt0;
            const double m = fmax(d1, d2);
            const double dt1 = (m <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(m)) / 5.0);
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
        if (fabs(tnew - tend) < 100.0 * ulp_of(fmax(fabs(t), fabs(tend)))) tnew = tend;
        // ---- one Tsit5 step on duals ----
        for (int i = 0; i < Z; ++i) { sum = Tb::a21 * k[0][i]; tmp[i] = fd_axpy_time(u[i], dts, sum); }
        fd_rhs<S, NP, FRICTION>(k[1], tmp, L);
        for (int i = 0; i < Z; ++i) {
            sum = Tb::a31 * k[0][i]; sum = fd_fma(Tb::a32, k[1][i], sum);
            tmp[i] = fd_axpy_time(u[i], dts, sum);
        }
        fd_rhs<S, NP, FRICTION>(k[2], tmp, L);
        for (int i = 0; i < Z; ++i) {
            sum = Tb::a41 * k[0][i]; sum = fd_fma(Tb::a42, k[1][i], sum); sum = fd_fma(Tb::a43, k[2][i], sum);
            tmp[i] = fd_axpy_time(u[i], dts, sum);
        }
        fd_rhs<S, NP, FRICTION>(k[3], tmp, L);
        for (int i = 0; i < Z; ++i) {
            sum = Tb::a51 * k[0][i]; sum = fd_fma(Tb::a52, k[1][i], sum); sum = fd_fma(Tb::a53, k[2][i], sum);
            sum = fd_fma(Tb::a54, k[3][i], sum);
            tmp[i] = fd_axpy_time(u[i], dts, sum);
        }
        fd_rhs<S, NP, FRICTION>(k[4], tmp, L);
        for (int i = 0; i < Z; ++i) {
            sum = Tb::a61 * k[0][i]; sum = fd_fma(Tb::a62, k[1][i], sum); sum = fd_fma(Tb::a63, k[2][i], sum);
            sum = fd_fma(Tb::a64, k[3][i], sum); sum = fd_fma(Tb::a65, k[4][i], sum);
            tmp[i] = fd_axpy_time(u[i], dts, sum);
        }
        fd_rhs<S, NP, FRICTION>(k[5], tmp, L);
        for (int i = 0; i < Z; ++i) {
            sum = Tb::a71 * k[0][i]; sum = fd_fma(Tb::a72, k[1][i], sum); sum = fd_fma(Tb::a73, k[2][i], sum);
            sum = fd_fma(Tb::a74, k[3][i], sum); sum = fd_fma(Tb::a75, k[4][i], sum); sum = fd_fma(Tb::a76, k[5][i], sum);
            unew[i] = fd_axpy_time(u[i], dts, sum);
        }
        fd_rhs<S, NP, FRICTION>(k[6], unew, L);
        double EEst = 0.0;
        if (o.adaptive) {
            for (int i = 0; i < Z; ++i) {
                sum = Tb::bt1 * k[0][i]; sum = fd_fma(Tb::bt2, k[1][i], sum); sum = fd_fma(Tb::bt3, k[2][i], sum);
                sum = fd_fma(Tb::bt4, k[3][i], sum); sum = fd_fma(Tb::bt5, k[4][i], sum); sum = fd_fma(Tb::bt6, k[5][i], sum);
                sum = fd_fma(Tb::bt7, k[6][i], sum);
                const D ut = fd_scale_time(dts, sum);
                const S a0 = fd_absnorm(u[i], wp), a1 = fd_absnorm(unew[i], wp);
                const S sk = abstol + (a0 > a1 ? a0 : a1) * reltol;
                tmp[i] = fd_div_s(ut, sk);
            }
            EEst = (double)fd_rms(tmp, Z, wp);
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
        if (ks < T && o.adaptive && (!(fabs(dt) > dtmin) || !isfinite(dt))) { ret = RET_DTLESSTHANMIN; break; }
    }
    // a failed solve (dual or primal) contributes nothing: its NaN block is a constant of the differentiation (GOKU.jl:114)
    const bool ok = ret == RET_SUCCESS && primal_ret[b] == RET_SUCCESS;
#pragma unroll
    for (int q = 0; q < NP; ++q) dout[(size_t)b * NP + q] = ok ? (S)acc[q] : (S)0;
}

template <class S, bool FR>
static cudaError_t launch_fwdsens_t(const ldeq_tape* tp, const void* dtraj, void* dz0, void* dtheta, cudaStream_t s) {
    const int B = tp->B, grid = (B + 127) / 128;
    tsit5_fwdsens_kernel<S, 1, FR, true><<<grid, 128, 0, s>>>((const S*)tp->u, (const S*)tp->theta, tp->tgrid, B, tp->T, tp->kopts, 1,
                                                             (const S*)dtraj, tp->retcode, (S*)dtheta);
    tsit5_fwdsens_kernel<S, 2, FR, false><<<grid, 128, 0, s>>>((const S*)tp->u, (const S*)tp->theta, tp->tgrid, B, tp->T, tp->kopts, 1,
                                                              (const S*)dtraj, tp->retcode, (S*)dz0);
    return cudaGetLastError();
}

cudaError_t launch_fwdsens(const ldeq_tape* tp, const void* dtraj, void* dz0, void* dtheta, cudaStream_t s) {
    const bool fr = tp->rhs_kind == LDEQ_RHS_PENDULUM_FRICTION;
    if (tp->dtype == LDEQ_F32)
        return fr ? launch_fwdsens_t<float, true>(tp, dtraj, dz0, dtheta, s) : launch_fwdsens_t<float, false>(tp, dtraj, dz0, dtheta, s);
    return fr ? launch_fwdsens_t<double, true>(tp, dtraj, dz0, dtheta, s) : launch_fwdsens_t<double, false>(tp, dtraj, dz0, dtheta, s);
}

}  // namespace ldeq
