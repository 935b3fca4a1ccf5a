// ldeq_mlp_cadj.cuh -- LDEQ_SENSE_INTERPOLATING_ADJOINT for the LatentODE path: the reverse pass the REFERENCE runs.
//
// DiffEqFlux's NeuralODE differentiates its solve (src/models/LatentODE.jl:70-72 under Zygote) with SciMLSensitivity's
// InterpolatingAdjoint(autojacvec = ZygoteVJP()) [3P; restated from the published algorithm, the same
// restatement as oracle/mlp.py::interpolating_adjoint]: ONE backward ODE on the augmented state
//     y = [lambda (D x B); mu (n_params)],   lambda' = -(df/du)^T lambda,   mu' = -(df/dp)^T lambda,
// integrated from t_end to t_0 by adaptive Tsit5 at the solve's abstol / reltol -- RMS error norm over the whole
// augmented vector, Hairer initial step, PI controller -- with u(t) read from the forward solution's dense output and a
// PresetTimeCallback at every save time (the step is clipped to it, lambda += dtraj[k] there, f is re-evaluated).
//
// Included by ldeq_mlp.cu (uses its dense_bwd_input / mlp_fwd building blocks).  One cooperative launch:
//   * a CTA owns a tile of TB trajectories for the whole sweep: lambda, the seven stage derivatives and the MLP
//     activations live in its shared memory; the stages of a step need no exchange (mu does not feed back into f);
//   * the forward dense output is rebuilt first: the tape holds (t_n, dt_n, u_n) of every accepted step, the CTA
//     recomputes k_1..k_7 for its rows and keeps them in a global scratch (7 D floats per row and step);
//   * mu enters a step only through  mu_new = mu - dt sum_j b_j k_j^mu  and the error estimate  dt sum_j btilde_j k_j^mu,
//     both linear in the per-row weight gradients: every stage's VJP adds its weight gradient, scaled by b_j and by
//     btilde_j, into the CTA's two private accumulators (RED atomics, one add per address and stage); after the seventh
//     stage the grid meets, every CTA sums its slice of the parameters over all CTAs in a fixed order, forms
//     mu_new / the scaled error of its slice and of its lambda rows, and one deterministic grid sum yields the error
//     norm: two grid barriers per attempted step, every CTA takes the same accept / reject decision.
#pragma once

namespace ldeq {

// gW{a,b}[k*N + n] (+)= w{a,b} * sum_b dy[n][b] x[k][b];  g{a,b}b[n] (+)= w{a,b} * sum_b dy[n][b].
// store: the first stage of an attempt overwrites (no separate zeroing pass); later stages add with reduction atomics
// (the accumulators are private to the CTA and every address gets one add per stage, so the sums do not depend on timing).
template <class S, int TB>
__device__ void dense_bwd_params2(S* __restrict__ gWa, S* __restrict__ gba, S* __restrict__ gWb, S* __restrict__ gbb, S wa, S wb,
                                  bool store, const S* __restrict__ dy, const S* __restrict__ x, int K, int N) {
    int SL = 1;
    while (SL * 2 * N <= MLP_THREADS && SL < 16 && SL * 2 <= K) SL *= 2;
    const int kper = (K + SL - 1) / SL;
    for (int w = threadIdx.x; w < N * SL; w += MLP_THREADS) {
        const int s = w / N, n = w - s * N;
        const int k0 = s * kper, k1 = min(K, k0 + kper);
        S d[TB];
#pragma unroll
        for (int b = 0; b < TB; ++b) d[b] = dy[n * TB + b];
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
            S a = (S)0;
#pragma unroll
            for (int b = 0; b < TB; ++b) a = s_fma<S>(d[b], x[k * TB + b], a);
            const size_t idx = (size_t)k * N + n;
            if (store) {
                gWa[idx] = wa * a;
                gWb[idx] = wb * a;
            } else {
                if (wa != (S)0) atomicAdd(gWa + idx, wa * a);
                atomicAdd(gWb + idx, wb * a);
            }
        }
        if (s == 0) {
            S a = (S)0;
#pragma unroll
            for (int b = 0; b < TB; ++b) a += d[b];
            if (store) {
                gba[n] = wa * a;
                gbb[n] = wb * a;
            } else {
                if (wa != (S)0) atomicAdd(gba + n, wa * a);
                atomicAdd(gbb + n, wb * a);
            }
        }
    }
    __syncthreads();
}

// mlp_vjp (ldeq_mlp.cu) with the two weighted accumulators.  (The VJP helpers of the general variants are __noinline__: the
// kernel evaluates them at four call sites, and nine kernel variants with every dense product inlined four times took the
// translation unit to five minutes of compile time.  The resident variant -- the default at the reference's network --
// keeps its helpers inlined: out of line it lost 87 -> 116 ms, the network descriptor then lives behind a pointer.)
template <class S, int TB>
__device__ __noinline__ void mlp_vjp2(const MlpNet& net, const S* __restrict__ P, const S* __restrict__ Pt, S* __restrict__ gA, S* __restrict__ gB,
                         S wa, S wb, bool store, const S* x, const S* kbar, S* gbar, S* acts, S* dbuf0, S* dbuf1, S* ytmp,
                         S* red, int HW) {
    const S* in = x;
    for (int l = 0; l < net.n_layers; ++l) {
        const bool last = l + 1 == net.n_layers;
        S* out = last ? ytmp : acts + (size_t)l * HW * TB;
        dense_fwd<S, TB>(P + net.w_off[l], P + net.b_off[l], in, out, red, net.dims[l], net.dims[l + 1], !last);
        in = out;
    }
    const S* dy = kbar;
    for (int l = net.n_layers - 1; l >= 0; --l) {
        const int K = net.dims[l], N = net.dims[l + 1];
        const S* xin = l == 0 ? x : acts + (size_t)(l - 1) * HW * TB;
        dense_bwd_params2<S, TB>(gA + net.w_off[l], gA + net.b_off[l], gB + net.w_off[l], gB + net.b_off[l], wa, wb, store, dy, xin, K, N);
        S* dx = l == 0 ? gbar : ((l & 1) ? dbuf1 : dbuf0);
        dense_bwd_input<S, TB>(Pt + net.w_off[l], dy, dx, red, K, N);
        if (l > 0) {
            for (int i = threadIdx.x; i < K * TB; i += MLP_THREADS) dx[i] = xin[i] > (S)0 ? dx[i] : (S)0;
            __syncthreads();
        }
        dy = dx;
    }
}

template <class S> __device__ __forceinline__ S tab_b(int i) { return tab_a<S>(6, i); }  // b_j = a_7j, b_7 = 0
template <class S> __device__ __forceinline__ S tab_c(int j) {
    using Tb = Tab<S>;
    switch (j) {
        case 1: return Tb::c2; case 2: return Tb::c3; case 3: return Tb::c4; case 4: return Tb::c5; case 5: case 6: return (S)1;
    }
    return (S)0;
}

// ---- BATCH: the weight gradients of an attempt in ONE pass ---------------------------------------------------------------
// The seven stages of an attempted step add  b_j (df/dp)^T lambda_j  and  btilde_j (df/dp)^T lambda_j  into the CTA's two
// accumulators.  Doing that per stage costs two reduction atomics per parameter and stage (at C2: 655 k per CTA and
// attempt, 84 M per attempt over the grid -- the L2's atomic throughput, not the arithmetic, set the time of an attempt).
// Instead every stage leaves its layer inputs x_l and pre-activation cotangents dy_l in a shared-memory record, and one
// pass after the last stage forms  gA = sum_j b_j dy_j x_j^T,  gB = sum_j btilde_j dy_j x_j^T  with plain stores:
// 2 stores per parameter and attempt instead of 14 atomics.
template <class S, int TB> struct CadjRec {
    // one stage's record: x (D) | kbar (D) | acts [(L-1) HW] | dys [(L-1) HW], each x TB
    __host__ __device__ static size_t floats(int D, int HW, int L) { return (size_t)(2 * D + 2 * (L - 1) * HW) * TB; }
};

// forward + input-VJP through the MLP, leaving the record of this stage in `rec` (no parameter accumulation)
template <class S, int TB>
__device__ __noinline__ void mlp_vjp_rec(const MlpNet& net, const S* __restrict__ P, const S* __restrict__ Pt, S* rec, S* gbar, S* ytmp, S* red, int HW) {
    // (general kernels: MLP_THREADS threads, weights through L1 / L2)
    const int D = net.dims[0], L = net.n_layers;
    S* xs = rec;
    S* kb = rec + D * TB;
    S* acts = kb + D * TB;
    S* dys = acts + (size_t)(L - 1) * HW * TB;
    const S* in = xs;
    for (int l = 0; l < L; ++l) {
        const bool last = l + 1 == L;
        S* out = last ? ytmp : acts + (size_t)l * HW * TB;
        dense_fwd<S, TB>(P + net.w_off[l], P + net.b_off[l], in, out, red, net.dims[l], net.dims[l + 1], !last);
        in = out;
    }
    const S* dy = kb;
    for (int l = L - 1; l >= 0; --l) {
        const int K = net.dims[l], N = net.dims[l + 1];
        const S* xin = l == 0 ? xs : acts + (size_t)(l - 1) * HW * TB;
        S* dx = l == 0 ? gbar : dys + (size_t)(l - 1) * HW * TB;
        dense_bwd_input<S, TB>(Pt + net.w_off[l], dy, dx, red, K, N);
        if (l > 0) {
            for (int i = threadIdx.x; i < K * TB; i += MLP_THREADS) dx[i] = xin[i] > (S)0 ? dx[i] : (S)0;
            __syncthreads();
        }
        dy = dx;
    }
}

// the same from the shared-memory-resident weight image (ldeq_mlp_common.cuh: dense_res, RES_THREADS threads, Float32):
// forward and transposed products read the one image, the relu mask is applied in the transposed product's epilogue
template <int TB>
__device__ void mlp_vjp_rec_res(const MlpNet& net, const float* img, float* rec, float* gbar, float* ytmp, int HW) {
    const int D = net.dims[0], L = net.n_layers;
    float* xs = rec;
    float* kb = rec + D * TB;
    float* acts = kb + D * TB;
    float* dys = acts + (size_t)(L - 1) * HW * TB;
    const float* in = xs;
    for (int l = 0; l < L; ++l) {
        const bool last = l + 1 == L;
        float* out = last ? ytmp : acts + (size_t)l * HW * TB;
        dense_res<TB>(net, img, l, 0, in, out, !last, nullptr);
        in = out;
    }
    const float* dy = kb;
    for (int l = L - 1; l >= 0; --l) {
        const float* xin = l == 0 ? xs : acts + (size_t)(l - 1) * HW * TB;
        float* dx = l == 0 ? gbar : dys + (size_t)(l - 1) * HW * TB;
        dense_res<TB>(net, img, l, 1, dy, dx, false, l > 0 ? xin : nullptr);
        dy = dx;
    }
}

// gA = sum_j wa[j] dy_j x_j^T, gB = sum_j wb[j] dy_j x_j^T over the stages in `jmask`, for every layer; plain stores
template <class S, int TB, int NT>
__device__ void cadj_param_pass(const MlpNet& net, const S* recs, size_t rec_stride, unsigned jmask, const S* wa, const S* wb,
                                S* __restrict__ gA, S* __restrict__ gB, int HW) {
    const int D = net.dims[0], L = net.n_layers;
    for (int l = 0; l < L; ++l) {
        const int K = net.dims[l], N = net.dims[l + 1];
        // offsets of this layer's input and cotangent inside a record
        const size_t o_x = l == 0 ? 0 : (size_t)(2 * D + (size_t)(l - 1) * HW) * TB;
        const size_t o_dy = l == L - 1 ? (size_t)D * TB : (size_t)(2 * D + (size_t)(L - 1) * HW + (size_t)l * HW) * TB;
        S* gWa = gA + net.w_off[l];
        S* gWb = gB + net.w_off[l];
        int SL = 1;
        while (SL * 2 * N <= NT && SL < 16 && SL * 2 <= K) SL *= 2;
        const int kper = (K + SL - 1) / SL;
        for (int w = threadIdx.x; w < N * SL; w += NT) {
            const int sl = w / N, n = w - sl * N;
            const int k0 = sl * kper, k1 = min(K, k0 + kper);
            S da[7][TB], db[7][TB];
            S sa = (S)0, sb = (S)0;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
#pragma unroll
                for (int b = 0; b < TB; ++b) { da[j][b] = (S)0; db[j][b] = (S)0; }
                if (!((jmask >> j) & 1u)) continue;
                const S* dy = recs + j * rec_stride + o_dy;
#pragma unroll
                for (int b = 0; b < TB; ++b) {
                    const S d = dy[n * TB + b];
                    da[j][b] = wa[j] * d;
                    db[j][b] = wb[j] * d;
                    sa += da[j][b];
                    sb += db[j][b];
                }
            }
#pragma unroll 2
            for (int k = k0; k < k1; ++k) {
                S a = (S)0, bsum = (S)0;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    if (!((jmask >> j) & 1u)) continue;
                    const S* x = recs + j * rec_stride + o_x + (size_t)k * TB;
#pragma unroll
                    for (int b = 0; b < TB; ++b) {
                        const S xv = x[b];
                        a = s_fma<S>(da[j][b], xv, a);
                        bsum = s_fma<S>(db[j][b], xv, bsum);
                    }
                }
                gWa[(size_t)k * N + n] = a;
                gWb[(size_t)k * N + n] = bsum;
            }
            if (sl == 0) {
                gA[net.b_off[l] + n] = sa;
                gB[net.b_off[l] + n] = sb;
            }
        }
    }
    __syncthreads();
}

struct CadjShared {
    double tc, dt, dts, tnew, tq, th, dtn, sum;
    int n, ks, accept, ret;
};

// status[0..3] = {accepted, rejected, retcode, -} of the backward solve
// BATCH: stage records + one weight-gradient pass per attempt (above).  RES (Float32, implies BATCH): the weights are staged
// once into shared memory (the padded image of ldeq_mlp_res.cu) and every dense product reads them from there with
// RES_THREADS threads; the image leaves room for ONE working record, so the seven records of an attempt live in `grec`
// (global scratch, L2-resident: 47 kB per CTA written and read once per attempt).
template <class S, int TB, bool BATCH, bool RES>
__global__ void __launch_bounds__(RES ? RES_THREADS : MLP_THREADS)
mlp_cadj_kernel(MlpNet net, const S* __restrict__ P, const S* __restrict__ Pt, const double* __restrict__ tg, int B, int T, KOpts o,
                const S* __restrict__ dtraj, MlpTapeView<S> tape, const int* __restrict__ retcode, const int* __restrict__ naccept,
                S* __restrict__ dense, S* __restrict__ gscr, S* __restrict__ mubuf, double* __restrict__ partials,
                S* __restrict__ dz0, S* __restrict__ dparams, int* __restrict__ status, double* __restrict__ trace, int trace_cap,
                S* __restrict__ grec) {
    constexpr int NT = RES ? RES_THREADS : MLP_THREADS;
    static_assert(!RES || (BATCH && sizeof(S) == 4), "the resident variant is Float32 and keeps stage records");
    cg::grid_group grid = cg::this_grid();
    const int D = net.dims[0], HW = net.max_width, DT = D * TB, NP = net.n_params;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S* Wimg = reinterpret_cast<S*>(smem_raw);                      // RES: the padded weight image
    S* Y = Wimg + (RES ? ((net.n_img + 3) & ~3) : 0);  // lambda of this tile, feature-major [D][TB]
    S* KL = Y + DT;                          // [7][D][TB] stage derivatives of lambda
    S* G = KL + 7 * DT;                      // stage input / candidate
    S* X = G + DT;                           // u(t) from the dense output
    S* GB = X + DT;                          // J^T lambda
    S* ytmp = GB + DT;
    S* acts = ytmp + DT;
    S* dbuf0 = acts + (size_t)(net.n_layers - 1) * HW * TB;
    S* dbuf1 = dbuf0 + HW * TB;
    S* red = dbuf1 + HW * TB;                // [NT][TB] (general kernels only)
    S* recs = red + (RES ? 0 : NT * TB);     // BATCH: [7] stage records (CadjRec); RES: one working record
    const size_t rec_stride = CadjRec<S, TB>::floats(D, HW, net.n_layers);
    S* grecs = RES ? grec + (size_t)blockIdx.x * 7 * rec_stride : recs;   // where the attempt's seven records are read from
    if constexpr (RES) stage_weight_image<NT>(net, reinterpret_cast<const float*>(P), reinterpret_cast<float*>(Wimg));
    __shared__ CadjShared sh;
    __shared__ S s_wa[7], s_wb[7];
    __shared__ double s_red[NT / 32];

    const int b0 = blockIdx.x * TB;
    const int ncta = gridDim.x;
    const int na_f = naccept[0];             // one integrator for the batch: the same count on every row
    const bool fwd_ok = retcode[0] == RET_SUCCESS && na_f >= 1 && na_f <= tape.cap;
    S* gA = gscr + (size_t)blockIdx.x * 2 * NP;  // sum_j b_j (df/dp)^T lambda_j of this tile
    S* gB = gA + NP;                             // sum_j btilde_j ...
    S* mu = mubuf, *mu_new = mubuf + NP, *mu_f0 = mubuf + 2 * (size_t)NP;
    // this CTA's slice of the parameter vector in the grid reductions
    const int chunk = (NP + ncta - 1) / ncta;
    const int p_lo = min(NP, blockIdx.x * chunk), p_hi = min(NP, p_lo + chunk);
    const S abstol = (S)o.abstol, reltol = (S)o.reltol;
    const double nall = (double)D * (double)B + (double)NP;
    int gs_parity = 0;

    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        double r = 0.0;
        if (threadIdx.x == 0) {
            for (int i = 0; i < NT / 32; ++i) r += s_red[i];
            sh.sum = r;
        }
        __syncthreads();
        r = sh.sum;
        __syncthreads();
        return r;
    };
    auto finish = [&](bool nan_out) {
        for (int i = threadIdx.x; i < DT; i += NT) {
            const int d = i / TB, b = i - d * TB;
            if (b0 + b < B) dz0[(size_t)(b0 + b) * D + d] = nan_out ? s_nan<S>() : Y[i];
        }
        for (int p = p_lo + threadIdx.x; p < p_hi; p += NT) dparams[p] = nan_out ? s_nan<S>() : mu[p];
    };

    if (!fwd_ok) {  // a failed forward solve contributes no gradient (GOKU.jl:114 convention, uniform over the batch)
        for (int i = threadIdx.x; i < DT; i += NT) Y[i] = (S)0;
        for (int p = p_lo + threadIdx.x; p < p_hi; p += NT) mu[p] = (S)0;
        __syncthreads();
        finish(false);
        if (blockIdx.x == 0 && threadIdx.x == 0) { status[0] = 0; status[1] = 0; status[2] = RET_SUCCESS; }
        return;
    }

    // ---- the forward dense output of this tile: k_1..k_7 of every accepted step -------------------------------------
    const size_t Bld = (size_t)ncta * TB;
    for (int n = 0; n < na_f; ++n) {
        const double dtn = tape.dt[(size_t)n * B];
        for (int i = threadIdx.x; i < DT; i += NT) {
            const int d = i / TB, b = i - d * TB;
            Y[i] = b0 + b < B ? tape.u[((size_t)n * B + b0 + b) * D + d] : (S)0;
        }
        __syncthreads();
        for (int j = 0; j < 7; ++j) {
            if (j == 0) {
                for (int i = threadIdx.x; i < DT; i += NT) G[i] = Y[i];
            } else {
                for (int i = threadIdx.x; i < DT; i += NT) {
                    S acc = tab_a<S>(j, 0) * KL[i];
                    for (int q = 1; q < j; ++q) acc = s_fma<S>(tab_a<S>(j, q), KL[q * DT + i], acc);
                    G[i] = s_fma<S>((S)dtn, acc, Y[i]);
                }
            }
            __syncthreads();
            if constexpr (RES) mlp_fwd_res<TB>(net, reinterpret_cast<const float*>(Wimg), reinterpret_cast<const float*>(G), reinterpret_cast<float*>(KL + j * DT),
                                               reinterpret_cast<float*>(dbuf0), reinterpret_cast<float*>(dbuf1), nullptr);
            else mlp_fwd<S, TB>(net, P, G, KL + j * DT, dbuf0, dbuf1, red);
            __syncthreads();
        }
        for (int i = threadIdx.x; i < 7 * DT; i += NT) {
            const int j = i / DT, r = i - j * DT, d = r / TB, b = r - d * TB;
            dense[(((size_t)n * 7 + j) * D + d) * Bld + b0 + b] = KL[i];
        }
        __syncthreads();
    }

    // rhs of the lambda part at time tq with stage input `lam`: KL[j] = -(df/du)^T lam; weight gradients into gA / gB
    auto eval = [&](double tq, const S* lam, int j, S wa, S wb, bool store) {
        if (threadIdx.x == 0) {
            int n = sh.n;
            while (n > 0 && tape.t[(size_t)n * B] > tq) --n;
            while (n + 1 < na_f && tape.t[(size_t)(n + 1) * B] <= tq) ++n;
            sh.n = n;
            sh.dtn = tape.dt[(size_t)n * B];
            sh.th = (tq - tape.t[(size_t)n * B]) / sh.dtn;
        }
        __syncthreads();
        {
            const int n = sh.n;
            const S dtn = (S)sh.dtn;
            S bw[7];
            interp_weights<S>((S)sh.th, bw);
            for (int i = threadIdx.x; i < DT; i += NT) {
                const int d = i / TB, b = i - d * TB;
                S v = (S)0;
                if (b0 + b < B) {
                    S acc = (S)0;
#pragma unroll
                    for (int q = 0; q < 7; ++q) acc = s_fma<S>(bw[q], dense[(((size_t)n * 7 + q) * D + d) * Bld + b0 + b], acc);
                    v = s_fma<S>(dtn, acc, tape.u[((size_t)n * B + b0 + b) * D + d]);
                }
                if (BATCH) {
                    S* rec = recs + (RES ? 0 : (size_t)j * rec_stride);
                    rec[i] = v;               // x of this stage
                    rec[DT + i] = lam[i];     // the stage's lambda (cotangent of f's output)
                } else {
                    X[i] = v;
                }
            }
        }
        __syncthreads();
        if constexpr (RES) {
            if (threadIdx.x == 0) { s_wa[j] = wa; s_wb[j] = wb; }
            mlp_vjp_rec_res<TB>(net, reinterpret_cast<const float*>(Wimg), reinterpret_cast<float*>(recs), reinterpret_cast<float*>(GB),
                                reinterpret_cast<float*>(ytmp), HW);
            S* dst = grecs + (size_t)j * rec_stride;     // park the record (read back by this CTA only, after its barriers)
            for (int i = threadIdx.x; i < (int)rec_stride; i += NT) dst[i] = recs[i];
        } else if (BATCH) {
            if (threadIdx.x == 0) { s_wa[j] = wa; s_wb[j] = wb; }
            mlp_vjp_rec<S, TB>(net, P, Pt, recs + (size_t)j * rec_stride, GB, ytmp, red, HW);
        } else {
            mlp_vjp2<S, TB>(net, P, Pt, gA, gB, wa, wb, store, X, lam, GB, acts, dbuf0, dbuf1, ytmp, red, HW);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < DT; i += NT) KL[j * DT + i] = -GB[i];
        __syncthreads();
    };
    // sum over all CTAs of accumulator `which` (0: gA, 1: gB) at parameter p, in CTA order
    auto acc_sum = [&](int which, int p) -> S {
        const S* base = gscr + (size_t)which * NP + p;
        S s0 = (S)0, s1 = (S)0;
        int c = 0;
        for (; c + 1 < ncta; c += 2) {
            s0 += __ldcg(base + (size_t)c * 2 * NP);
            s1 += __ldcg(base + (size_t)(c + 1) * 2 * NP);
        }
        if (c < ncta) s0 += __ldcg(base + (size_t)c * 2 * NP);
        return s0 + s1;
    };

    // ---- initial state: the callback at t_end has fired ---------------------------------------------------------------
    const double t0 = tg[0], tend = tg[T - 1];
    const double dtmax = o.dtmax > 0.0 ? o.dtmax : (tend - t0);
    const double dtmin = o.dtmin > 0.0 ? o.dtmin : fmax(2.220446049250313e-16, ulp_of(tend));
    for (int i = threadIdx.x; i < DT; i += NT) {
        const int d = i / TB, b = i - d * TB;
        Y[i] = b0 + b < B ? dtraj[((size_t)(T - 1) * B + b0 + b) * D + d] : (S)0;
    }
    for (int p = p_lo + threadIdx.x; p < p_hi; p += NT) mu[p] = (S)0;
    if (threadIdx.x == 0) { sh.n = na_f - 1; sh.tc = tend; sh.ks = T - 2; sh.ret = RET_SUCCESS; sh.dt = o.dt; }
    __syncthreads();

    if (o.adaptive && !(o.dt > 0.0)) {
        // ode_determine_initdt on the augmented state, time running backwards
        eval(tend, Y, 0, (S)1, (S)0, true);
        if (BATCH) cadj_param_pass<S, TB, NT>(net, grecs, rec_stride, 1u, s_wa, s_wb, gA, gB, HW);
        grid.sync();
        double a0 = 0.0, a1 = 0.0;
        for (int p = p_lo + threadIdx.x; p < p_hi; p += NT) {
            const S f0 = -acc_sum(0, p);  // mu' = -(df/dp)^T lambda; mu = 0: sk = abstol
            mu_f0[p] = f0;
            const S r = f0 / abstol;
            a1 += (double)(r * r);
        }
        for (int i = threadIdx.x; i < DT; i += NT) {
            if (b0 + (i % TB) < B) {
                const S sk = s_fma<S>(s_abs<S>(Y[i]), reltol, abstol);
                const S r0 = Y[i] / sk, r1 = KL[i] / sk;
                a0 += (double)(r0 * r0);
                a1 += (double)(r1 * r1);
            }
        }
        a0 = block_sum(a0);
        a1 = block_sum(a1);
        const double d0 = sqrt(grid_sum(a0, partials, grid, gs_parity) / nall);
        const double d1 = sqrt(grid_sum(a1, partials, grid, gs_parity) / nall);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = fmin(dt0, dtmax);
        for (int i = threadIdx.x; i < DT; i += NT) G[i] = s_fma<S>(-(S)dt0, KL[i], Y[i]);
        __syncthreads();
        eval(tend - dt0, G, 1, (S)1, (S)0, true);
        if (BATCH) cadj_param_pass<S, TB, NT>(net, grecs, rec_stride, 2u, s_wa, s_wb, gA, gB, HW);
        grid.sync();
        double a2 = 0.0;
        for (int p = p_lo + threadIdx.x; p < p_hi; p += NT) {
            const S r = (-acc_sum(0, p) - mu_f0[p]) / abstol;
            a2 += (double)(r * r);
        }
        for (int i = threadIdx.x; i < DT; i += NT) {
            if (b0 + (i % TB) < B) {
                const S sk = s_fma<S>(s_abs<S>(Y[i]), reltol, abstol);
                const S r = (KL[DT + i] - KL[i]) / sk;
                a2 += (double)(r * r);
            }
        }
        a2 = block_sum(a2);
        const double d2 = sqrt(grid_sum(a2, partials, grid, gs_parity) / nall) / dt0;
        const double m = fmax(d1, d2);
        const double dt1 = (m <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(m)) / 5.0);
        if (threadIdx.x == 0) sh.dt = fmax(dtmin, fmin(100.0 * dt0, fmin(dt1, dtmax)));
        __syncthreads();
    }

    // ---- the backward sweep ---------------------------------------------------------------------------------------------
    PiState pst = pi_init(o);
    int na = 0, nr = 0;
    long long iters = 0;
    int ks = T - 2;
    double tc = tend, dt = sh.dt;
    int ret = (!(dt > 0.0) || !isfinite(dt)) ? RET_DTLESSTHANMIN : RET_SUCCESS;
    while (ks >= 0 && ret == RET_SUCCESS) {
        if (iters >= o.maxiters) { ret = RET_MAXITERS; break; }
        ++iters;
        const double tstop = tg[ks];
        const double dts = fmin(dt, tc - tstop);
        double tnew = tc - dts;
        if (fabs(tnew - tstop) < 100.0 * ulp_of(fmax(fabs(tc), fabs(tstop)))) tnew = tstop;
        // seven stages; f is re-evaluated at the start of every step (after a callback it has to be, otherwise it equals k_7)
        eval(tc, Y, 0, tab_b<S>(0), tab_bt<S>(0), true);
        for (int j = 1; j < 7; ++j) {
            for (int i = threadIdx.x; i < DT; i += NT) {
                S acc = tab_a<S>(j, 0) * KL[i];
                for (int q = 1; q < j; ++q) acc = s_fma<S>(tab_a<S>(j, q), KL[q * DT + i], acc);
                G[i] = s_fma<S>(-(S)dts, acc, Y[i]);
            }
            __syncthreads();
            eval(tc - (double)tab_c<double>(j) * dts, G, j, j < 6 ? tab_b<S>(j) : (S)0, tab_bt<S>(j), false);
        }
        // G = lambda candidate (a_7j = b_j).  Every tile's weight gradients must be complete before the parameter slices are summed.
        if (BATCH) cadj_param_pass<S, TB, NT>(net, grecs, rec_stride, 0x7fu, s_wa, s_wb, gA, gB, HW);
        grid.sync();
        double e2 = 0.0;
        for (int p = p_lo + threadIdx.x; p < p_hi; p += NT) {
            // k_j^mu = -sum_rows (df/dp)^T lambda_j:  mu_new = mu - dts sum b_j k_j^mu = mu + dts * sum_cta gA
            const S m_old = mu[p];
            const S m_new = s_fma<S>((S)dts, acc_sum(0, p), m_old);
            mu_new[p] = m_new;
            if (o.adaptive) {
                const S em = (S)dts * acc_sum(1, p);
                const S r = em / s_fma<S>(s_max<S>(s_abs<S>(m_old), s_abs<S>(m_new)), reltol, abstol);
                e2 += (double)(r * r);
            }
        }
        if (o.adaptive) {
            for (int i = threadIdx.x; i < DT; i += NT) {
                if (b0 + (i % TB) < B) {
                    S acc = tab_bt<S>(0) * KL[i];
                    for (int q = 1; q < 7; ++q) acc = s_fma<S>(tab_bt<S>(q), KL[q * DT + i], acc);
                    const S el = (S)dts * acc;
                    const S r = el / s_fma<S>(s_max<S>(s_abs<S>(Y[i]), s_abs<S>(G[i])), reltol, abstol);
                    e2 += (double)(r * r);
                }
            }
        }
        bool accept = true;
        double dt_next = dt;
        if (o.adaptive) {
            e2 = block_sum(e2);
            const double EEst = sqrt(grid_sum(e2, partials, grid, gs_parity) / nall);
            if (!(EEst == EEst) || isinf(EEst)) { ret = RET_UNSTABLE; break; }
            accept = pi_controller(o, EEst, dts, dtmax, pst, dt_next);
            if (trace && blockIdx.x == 0 && threadIdx.x == 0 && iters <= trace_cap) {
                double* tr = trace + 4 * (iters - 1);
                tr[0] = tc; tr[1] = dts; tr[2] = EEst; tr[3] = accept ? 1.0 : 0.0;
            }
        } else {
            grid.sync();  // mu_new of the neighbours' slices is not read, but gA / gB may be overwritten only after everyone summed
        }
        if (accept) {
            ++na;
            tc = tnew;
            for (int p = p_lo + threadIdx.x; p < p_hi; p += NT) mu[p] = mu_new[p];
            if (tc == tstop) {
                // the callback: cotangent of save point ks
                for (int i = threadIdx.x; i < DT; i += NT) {
                    const int d = i / TB, b = i - d * TB;
                    Y[i] = G[i] + (b0 + b < B ? dtraj[((size_t)ks * B + b0 + b) * D + d] : (S)0);
                }
                --ks;
            } else {
                for (int i = threadIdx.x; i < DT; i += NT) Y[i] = G[i];
            }
            __syncthreads();
        } else {
            ++nr;
        }
        if (o.adaptive) dt = dt_next;
        if (ks >= 0 && o.adaptive && (!(fabs(dt) > dtmin) || !isfinite(dt))) { ret = RET_DTLESSTHANMIN; break; }
    }
    __syncthreads();
    finish(ret != RET_SUCCESS);
    if (blockIdx.x == 0 && threadIdx.x == 0) { status[0] = na; status[1] = nr; status[2] = ret; }
}

template <class S, int TB> static size_t cadj_smem(const MlpNet& net, bool batch, bool res = false) {
    const size_t D = net.dims[0], HW = net.max_width;
    size_t n = (12 * D + (net.n_layers - 1) * HW + 2 * HW + (res ? 0 : MLP_THREADS)) * TB;
    if (res) n += ((size_t)net.n_img + 3) & ~(size_t)3;
    if (batch) n += (res ? 1 : 7) * CadjRec<S, TB>::floats(net.dims[0], net.max_width, net.n_layers);
    return (n * sizeof(S) + 15) & ~(size_t)15;
}

}  // namespace ldeq
