// ldeq_fwdsens.cu -- LDEQ_SENSE_FORWARD_DUAL for the built-in right-hand sides: instantiations of ldeq_fwdsens.cuh (the
// reference's ForwardDiffSensitivity pullback restated literally).  Compiled with -fmad=false (build.py) like the oracle's
// -ffp-contract=off: only the explicit fma() calls fuse.  The default reverse pass stays the discrete adjoint
// (~7x cheaper, equal to this within the solver tolerance).
#include "ldeq_internal.h"
#include "ldeq_fwdsens.cuh"
#include <cstdlib>

namespace ldeq {

// pendulum.jl:19-26 / :65-74 on duals (G, b, m are Float32 literals of the reference, rounded to S)
template <bool FRICTION> struct PendulumDualRHS {
    static constexpr int ZD = 2, PD = 1;
    template <class D> __device__ __forceinline__ static void f(D* du, const D* u, const D* p, double) {
        typedef typename D::value_type S;
        const D G((S)10.0f);
        du[0] = u[1];
        const D a = (-G / p[0]) * sin(u[0]);
        if (FRICTION) {
            const S bm = (S)0.7f / (S)1.0f;
            du[1] = a - bm * u[1];
        } else {
            du[1] = a;
        }
    }
};

template <class S, int NP, bool FRICTION, bool SEED_P>
__global__ void __launch_bounds__(LDEQ_FWDSENS_THREADS, 512 / LDEQ_FWDSENS_THREADS)
tsit5_fwdsens_kernel(const S* __restrict__ z0, const S* __restrict__ theta, const double* __restrict__ tg, int B, GridInfo gi, int T,
                     KOpts o, int norm_partials, const int* __restrict__ sort_key, const S* __restrict__ dtraj, const int* __restrict__ primal_ret,
                     S* __restrict__ dout) {
    tsit5_fwdsens_body<PendulumDualRHS<FRICTION>, S, NP, SEED_P>(z0, theta, tg, B, gi, T, o, norm_partials, sort_key, dtraj, primal_ret, dout);
}

template <class S, bool FR>
static cudaError_t launch_fwdsens_t(const ldeq_tape* tp, const void* dtraj, int ld, void* dz0, void* dtheta, cudaStream_t s) {
    const int B = tp->B, grid = (B + LDEQ_FWDSENS_THREADS - 1) / LDEQ_FWDSENS_THREADS;
    const GridInfo gi{tp->grid_t0, tp->grid_h, tp->grid_uniform, ld};
    // LDEQ_FWDSENS_SORT=0 keeps the trajectory -> lane assignment of the batch order (A/B switch; the results are identical)
    static const bool sort = [] { const char* e = getenv("LDEQ_FWDSENS_SORT"); return !(e && e[0] == '0'); }();
    const int* key = sort ? tp->naccept : nullptr;
    tsit5_fwdsens_kernel<S, 1, FR, true><<<grid, LDEQ_FWDSENS_THREADS, 0, s>>>((const S*)tp->u, (const S*)tp->theta, tp->tgrid, B, gi, tp->T, tp->kopts, 1, key,
                                                             (const S*)dtraj, tp->retcode, (S*)dtheta);
    tsit5_fwdsens_kernel<S, 2, FR, false><<<grid, LDEQ_FWDSENS_THREADS, 0, s>>>((const S*)tp->u, (const S*)tp->theta, tp->tgrid, B, gi, tp->T, tp->kopts, 1, key,
                                                              (const S*)dtraj, tp->retcode, (S*)dz0);
    return cudaGetLastError();
}

cudaError_t launch_fwdsens(const ldeq_tape* tp, const void* dtraj, int ld, void* dz0, void* dtheta, cudaStream_t s) {
    const bool fr = tp->rhs_kind == LDEQ_RHS_PENDULUM_FRICTION;
    if (tp->dtype == LDEQ_F32)
        return fr ? launch_fwdsens_t<float, true>(tp, dtraj, ld, dz0, dtheta, s) : launch_fwdsens_t<float, false>(tp, dtraj, ld, dz0, dtheta, s);
    return fr ? launch_fwdsens_t<double, true>(tp, dtraj, ld, dz0, dtheta, s) : launch_fwdsens_t<double, false>(tp, dtraj, ld, dz0, dtheta, s);
}

}  // namespace ldeq
