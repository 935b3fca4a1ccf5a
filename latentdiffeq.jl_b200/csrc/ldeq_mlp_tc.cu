// ldeq_mlp_tc.cu -- LatentODE forward solve with the MLP right-hand side on the 5th-generation tensor cores
// (tcgen05 + TMEM), `LDEQ_MLP_MATH_BF16X3`.
//
// Same semantics as mlp_fwd_kernel (ldeq_mlp.cu; reference src/models/LatentODE.jl:61-78 with the NODE struct's
// Chain(Dense(D,H,relu), Dense(H,H,relu), Dense(H,D)), examples/pendulum_friction-less/nODE.jl:14-16), for
// batches where batch x hidden is a genuine dense contraction.
//
// One CTA integrates a tile of 128 trajectories for the whole time span; thread r owns trajectory (row) r:
// its state, step size and controller memory stay in registers, exactly as in the GOKU kernel.  One RHS
// evaluation is three dependent GEMMs with M = 128 (the tile), issued by one thread as tcgen05.mma:
//   * B operand = weights, staged ONCE per CTA in shared memory as bf16 "hi" and "lo" images in the UMMA
//     canonical K-major (no swizzle) layout (195 kB for 16-200-200-16, widths padded to multiples of 16);
//   * A operand = activations, read by the MMA straight from TMEM: the epilogue of layer l (tcgen05.ld the
//     fp32 accumulators, + bias, relu, split into bf16 hi/lo, tcgen05.st) writes the A operand of layer l+1,
//     so activations never touch shared memory;
//   * fp32 parity: x*w ~ x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (three MMAs per K-step, fp32 accumulate in TMEM),
//     relative error 2^-16 per product instead of bf16's 2^-8;
//   * the stage slopes k1..k6 live in the 96 TMEM columns left over (512 = 208 accumulator + 2*104 operand + 96).
// TMEM column map: see the TC_COL_* constants below.
#include "ldeq_gridsum.cuh"
#include <cuda_bf16.h>

#include "ldeq_internal.h"


namespace ldeq {

#define TC_THREADS 256       // 8 warps: warps w and w+4 share TMEM lane quadrant w%4 and split the columns
#define TC_ROWS 128          // trajectories per CTA (M of the MMA)
#define TC_MAXW 208          // largest padded layer width this kernel is built for (N of one MMA, K = 13 steps)
// TMEM column map (512 columns): two 208-column regions used in ping-pong + the stage slopes.
//   layer 1: A = g in R1[0:16)        -> D = R0        epilogue 1 rewrites R0 IN PLACE as the bf16 hi/lo A operand
//   layer 2: A = R0                   -> D = R1        epilogue 2 rewrites R1 in place
//   layer 3: A = R1                   -> D = R0[0:16)
// In-place layout of an A operand: the 16 fp32 accumulator columns of K-step kk become 8 columns of packed bf16
// "hi" pairs followed by 8 columns of "lo" pairs.  Because an epilogue only touches the columns it has just read,
// it can run on one column half while the MMAs of the other half are still in flight.
#define TC_COL_R0 0
#define TC_COL_R1 208
#define TC_COL_K 416         // stage slopes k1..k6: 6 x 16 columns [416, 512)
#define TC_TMEM_COLS 512

struct TcNet {
    int d;        // state dimension (<= 16, padded to 16)
    int n1, n2;   // hidden widths padded to multiples of 16 (<= TC_MAXW)
    int img_off[6];   // byte offsets of the weight images in shared memory: W1h, W1l, W2h, W2l, W3h, W3l
    int bias_off;     // byte offset of the padded fp32 biases (n1 + n2 + 16 floats)
    int smem_bytes;   // total image size
};

__device__ __constant__ float c_a[7][6] = {
    {0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.161f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {-0.008480655492356989f, 0.335480655492357f, 0.f, 0.f, 0.f, 0.f},
    {2.8971530571054935f, -6.359448489975075f, 4.3622954328695815f, 0.f, 0.f, 0.f},
    {5.325864828439257f, -11.748883564062828f, 7.4955393428898365f, -0.09249506636175525f, 0.f, 0.f},
    {5.86145544294642f, -12.92096931784711f, 8.159367898576159f, -0.071584973281401f, -0.028269050394068383f, 0.f},
    {0.09646076681806523f, 0.01f, 0.4798896504144996f, 1.379008574103742f, -3.290069515436081f, 2.324710524099774f}};
__device__ __constant__ float c_bt[7] = {-0.00178001105222577714f, -0.0008164344596567469f, 0.007880878010261995f,
                                         -0.1447110071732629f,      0.5823571654525552f,     -0.45808210592918697f,
                                         0.015151515151515152f};
__device__ __constant__ float c_r[7][4] = {{1.0f, -2.763706197274826f, 2.9132554618219126f, -1.0530884977290216f},
                                           {0.f, 0.13169999999999998f, -0.2234f, 0.1017f},
                                           {0.f, 3.9302962368947516f, -5.941033872131505f, 2.490627285651253f},
                                           {0.f, -12.411077166933676f, 30.33818863028232f, -16.548102889244902f},
                                           {0.f, 37.50931341651104f, -88.1789048947664f, 47.37952196281928f},
                                           {0.f, -27.896526289197286f, 65.09189467479366f, -34.87065786149661f},
                                           {0.f, 1.5f, -4.0f, 2.5f}};

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LDEQ_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LDEQ_DONE_%=;\n\t"
        "bra LDEQ_WAIT_%=;\n\t"
        "LDEQ_DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]; kind::f16 (bf16 inputs, fp32 accumulate), one CTA
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// 16 fp32 values -> bf16 hi / lo halves packed two per 32-bit TMEM column (even element in the low half).
// cvt.rn.bf16x2.f32 converts a pair in one instruction; bf16 -> fp32 is a shift.
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // first source -> upper half
    return r;
}
__device__ __forceinline__ void split_pack16(const float* x, uint32_t* hi, uint32_t* lo) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t h = pack_bf16x2(x[2 * j], x[2 * j + 1]);
        const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xFFFF0000u);
        hi[j] = h;
        lo[j] = pack_bf16x2(x[2 * j] - h0, x[2 * j + 1] - h1);
    }
}

// shared-memory matrix descriptor: K-major, no swizzle, canonical core matrices of 8 rows x 16 bytes;
// SBO (between 8-row groups) = 128 B, LBO (between 8-element K groups) = n_rows/8 * 128 B  (version 1 = sm_100)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = BF16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---- weight images -----------------------------------------------------------------------------------------
// flat Flux parameters -> bf16 hi/lo images in the canonical layout + padded biases, in global memory
__global__ void mlp_tc_prep_kernel(TcNet net, const float* __restrict__ P, int d_in, int h1, int h2, unsigned char* __restrict__ img) {
    // layer l: W (N x K) column-major at P[w_off + k*N + n]
    const int Ks[3] = {16, net.n1, net.n2};
    const int Ns[3] = {net.n1, net.n2, 16};
    const int Kr[3] = {d_in, h1, h2};
    const int Nr[3] = {h1, h2, d_in};
    int w_off = 0;
    for (int l = 0; l < 3; ++l) {
        const int Kp = Ks[l], Np = Ns[l];
        __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(img + net.img_off[2 * l]);
        __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(img + net.img_off[2 * l + 1]);
        const int lbo_elems = (Np / 8) * 64;  // 128 bytes = 64 bf16 per core matrix
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Kp * Np; i += gridDim.x * blockDim.x) {
            const int n = i % Np, k = i / Np;
            const float w = (n < Nr[l] && k < Kr[l]) ? P[w_off + (size_t)k * Nr[l] + n] : 0.f;
            const __nv_bfloat16 h = __float2bfloat16_rn(w);
            const __nv_bfloat16 lw = __float2bfloat16_rn(w - __bfloat162float(h));
            const int off = (k / 8) * lbo_elems + (n / 8) * 64 + (n % 8) * 8 + (k % 8);
            hi[off] = h;
            lo[off] = lw;
        }
        float* bias = reinterpret_cast<float*>(img + net.bias_off) + (l == 0 ? 0 : l == 1 ? net.n1 : net.n1 + net.n2);
        const int b_off = w_off + Kr[l] * Nr[l];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += gridDim.x * blockDim.x)
            bias[i] = i < Nr[l] ? P[b_off + i] : 0.f;
        w_off = b_off + Nr[l];
    }
}

template <class S> struct MlpTapeViewTc {
    double* t;
    double* dt;
    S* u;
    int cap;
};

// ---- forward ---------------------------------------------------------------------------------------------------
// Phases of the single-call-site state machine (the RHS evaluation is inlined exactly once).
enum { PH_F0 = 0, PH_INITDT = 1, PH_STAGE = 2 };

// epilogue of a hidden layer on the chunks [c_lo, c_hi) of a region: accumulators -> relu(acc + bias) -> bf16 hi/lo
// written back IN PLACE (8 hi columns + 8 lo columns per 16-column chunk) as the A operand of the next layer
__device__ __forceinline__ void tc_epilogue_inplace(uint32_t region_addr, const float* __restrict__ bias, int c_lo, int c_hi) {
    for (int c = c_lo; c < c_hi; c += 16) {
        float v[16];
        uint32_t hi[8], lo[8];
        tc_ld16(region_addr + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + bias[c + i], 0.f);
        split_pack16(v, hi, lo);
        tc_st8(region_addr + c, hi);
        tc_st8(region_addr + c + 8, lo);
    }
}

// MMAs of one layer for the output columns [n_lo, n_lo + n) : D[:, n_lo:n_lo+n] = A (K = 16*ksteps) x W[n_lo:n_lo+n, :]^T,
// bf16x3: hi*hi + hi*lo + lo*hi.  A operand in the in-place layout (hi at a_addr + 16*kk, lo at a_addr + 16*kk + 8).
__device__ __forceinline__ void tc_issue_layer(uint32_t d_addr, uint32_t a_addr, uint32_t w_hi, uint32_t w_lo, uint32_t lbo, int n_lo,
                                               int n, int ksteps) {
    if (n <= 0) return;  // a layer narrower than 32 has no second column half
    const uint32_t idesc = make_idesc(n);
    const uint32_t row_off = (uint32_t)(n_lo / 8) * 128;  // SBO = 128 bytes per 8-row group
    for (int kk = 0; kk < ksteps; ++kk) {
        const uint64_t bh = make_b_desc(w_hi + row_off + kk * 2 * lbo, lbo), bl = make_b_desc(w_lo + row_off + kk * 2 * lbo, lbo);
        tc_mma_ts(d_addr + n_lo, a_addr + 16 * kk, bh, idesc, kk > 0);
        tc_mma_ts(d_addr + n_lo, a_addr + 16 * kk, bl, idesc, 1);
        tc_mma_ts(d_addr + n_lo, a_addr + 16 * kk + 8, bh, idesc, 1);
    }
}

template <bool GLOBAL>
__global__ void __launch_bounds__(TC_THREADS, 1)
mlp_tc_fwd_kernel(TcNet net, const unsigned char* __restrict__ img_global, const float* __restrict__ z0,
                  const double* __restrict__ tg, int B, int T, KOpts o, float* __restrict__ traj, int* __restrict__ retcode,
                  int* __restrict__ naccept, int* __restrict__ nreject, MlpTapeViewTc<float> tape, double* __restrict__ partials) {
    cg::grid_group grid = cg::this_grid();
    int gs_parity = 0;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t mbar[3];  // [0] first column half, [1] second column half, [2] output layer
    __shared__ uint32_t tmem_base_s;
    __shared__ double red_s[TC_ROWS / 32];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int quad = warp & 3, hf = warp >> 2;     // TMEM lane quadrant / column half of this warp
    const int row = quad * 32 + (tid & 31);           // trajectory (row of the tile) this thread works on
    const int D = net.d;
    // stage the weight images (and biases) once
    {
        const uint4* src = reinterpret_cast<const uint4*>(img_global);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < net.smem_bytes / 16; i += TC_THREADS) dst[i] = src[i];
    }
    const float* bias1 = reinterpret_cast<const float*>(smem + net.bias_off);
    const float* bias2 = bias1 + net.n1;
    const float* bias3 = bias2 + net.n2;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // generic-proxy writes of the weight images must be visible to the tensor-core (async) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);  // this warp's 32-lane quadrant
    const uint32_t R0 = lane_addr + TC_COL_R0, R1 = lane_addr + TC_COL_R1, KS = lane_addr + TC_COL_K;
    uint32_t par_half = 0, par_out = 0;  // phase parity of the barrier this warp waits on for its half / the output layer

    const uint32_t lbo1 = (net.n1 / 8) * 128, lbo2 = (net.n2 / 8) * 128, lbo3 = (16 / 8) * 128;
    const uint32_t sbase = smem_u32(smem);
    // column halves of the two hidden layers (multiples of 16)
    const int n1a = ((net.n1 / 16 + 1) / 2) * 16, n2a = ((net.n2 / 16 + 1) / 2) * 16;

    const double t0 = tg[0], tend = tg[T - 1];
    const double dtmax = o.dtmax > 0.0 ? o.dtmax : (tend - t0);
    const double dtmin = o.dtmin > 0.0 ? o.dtmin : fmax(2.220446049250313e-16, ulp_of(t0));
    const float abstol = (float)o.abstol, reltol = (float)o.reltol;
    const int ntiles = (B + TC_ROWS - 1) / TC_ROWS;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile * TC_ROWS + row;
        const bool live = b < B;
        const bool writer = hf == 0;  // both column halves carry the row's state redundantly (identical arithmetic, hence
        // identical values: TMEM stores of state are issued by both); only one half writes to global memory
        float u[16], g[16], kout[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = (live && i < D) ? z0[(size_t)b * D + i] : 0.f;
        if (live && writer) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < D) traj[(size_t)b * D + i] = u[i];  // save point 0 is u0 itself
        }
        PiState pist = pi_init(o);
        double t = t0, dt = o.dt, dts = 0.0, tnew = t0, dt0 = 0.0, d1n = 0.0;
        int na = 0, nr = 0, ks = 1, ret = RET_SUCCESS, stage = 0;
        long long iters = 0;
        bool active = live && T > 1;
        int phase = PH_F0;

        for (;;) {
            // ---- input of this RHS evaluation -------------------------------------------------------------------
            if (phase == PH_F0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) g[i] = u[i];
            } else if (phase == PH_INITDT) {
                float k1[16];
                tc_ld16(KS, k1);  // f0 = k1
#pragma unroll
                for (int i = 0; i < 16; ++i) g[i] = fmaf((float)dt0, k1[i], u[i]);
            } else {
                if (stage == 1 && active) {
                    if (iters >= o.maxiters) { ret = RET_MAXITERS; active = false; }
                    else {
                        ++iters;
                        dts = fmin(dt, tend - t);
                        tnew = t + dts;
                        if (fabs(tnew - tend) < 100.0 * ulp_of(fmax(fabs(t), fabs(tend)))) tnew = tend;
                    }
                }
                float acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.f;
                for (int q = 0; q < stage; ++q) {
                    float kq[16];
                    tc_ld16(KS + 16 * q, kq);
                    const float a = c_a[stage][q];
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] = fmaf(a, kq[i], acc[i]);
                }
                const float h = (float)dts;
#pragma unroll
                for (int i = 0; i < 16; ++i) g[i] = fmaf(h, acc[i], u[i]);
            }
            // ---- RHS: three GEMMs on the tensor cores, activations TMEM -> TMEM ------------------------------------
            {
                {  // g -> R1[0:16) as bf16 hi (8 columns) + lo (8 columns)
                    uint32_t hi[8], lo[8];
                    split_pack16(g, hi, lo);
                    tc_st8(R1, hi);
                    tc_st8(R1 + 8, lo);
                    tc_wait_st();
                }
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t wh = sbase + net.img_off[0], wl = sbase + net.img_off[1];
                    tc_issue_layer(tmem_base + TC_COL_R0, tmem_base + TC_COL_R1, wh, wl, lbo1, 0, n1a, 1);
                    tc_commit(&mbar[0]);
                    tc_issue_layer(tmem_base + TC_COL_R0, tmem_base + TC_COL_R1, wh, wl, lbo1, n1a, net.n1 - n1a, 1);
                    tc_commit(&mbar[1]);
                }
                mbar_wait(&mbar[hf], par_half);
                par_half ^= 1;
                tc_fence_after();
                // epilogue 1 (this warp's column half), in place in R0 -> A operand of layer 2
                tc_epilogue_inplace(R0, bias1, hf ? n1a : 0, hf ? net.n1 : n1a);
                tc_wait_st();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t wh = sbase + net.img_off[2], wl = sbase + net.img_off[3];
                    tc_issue_layer(tmem_base + TC_COL_R1, tmem_base + TC_COL_R0, wh, wl, lbo2, 0, n2a, net.n1 / 16);
                    tc_commit(&mbar[0]);
                    tc_issue_layer(tmem_base + TC_COL_R1, tmem_base + TC_COL_R0, wh, wl, lbo2, n2a, net.n2 - n2a, net.n1 / 16);
                    tc_commit(&mbar[1]);
                }
                mbar_wait(&mbar[hf], par_half);
                par_half ^= 1;
                tc_fence_after();
                // epilogue 2, in place in R1 -> A operand of layer 3
                tc_epilogue_inplace(R1, bias2, hf ? n2a : 0, hf ? net.n2 : n2a);
                tc_wait_st();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    tc_issue_layer(tmem_base + TC_COL_R0, tmem_base + TC_COL_R1, sbase + net.img_off[4], sbase + net.img_off[5], lbo3, 0, 16,
                                   net.n2 / 16);
                    tc_commit(&mbar[2]);
                }
                mbar_wait(&mbar[2], par_out);
                par_out ^= 1;
                tc_fence_after();
                tc_ld16(R0, kout);
#pragma unroll
                for (int i = 0; i < 16; ++i) kout[i] = i < D ? kout[i] + bias3[i] : 0.f;
            }
            // ---- what the evaluation was for ---------------------------------------------------------------------
            bool all_done = false;
            if (phase == PH_F0) {
                tc_st16(KS, kout);  // k1 = f(u0)  (fsalfirst)
                tc_wait_st();
                if (o.adaptive && !(o.dt > 0.0)) {
                    // Hairer initial step, first half
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (i < D) {
                            const float sk = fmaf(fabsf(u[i]), reltol, abstol);
                            const float a = u[i] / sk, c = kout[i] / sk;
                            s0 = fmaf(a, a, s0);
                            s1 = fmaf(c, c, s1);
                        }
                    double e0 = s0, e1 = s1, n = D;
                    if (GLOBAL) {
                        // one dt for the batch: RMS over all D*B entries (summed by the `writer` half only)
                        double v0 = (live && writer) ? e0 : 0.0, v1 = (live && writer) ? e1 : 0.0;
                        for (int off = 16; off > 0; off >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, off); v1 += __shfl_xor_sync(0xffffffffu, v1, off); }
                        if (writer && (tid & 31) == 0) red_s[quad] = v0;
                        __syncthreads();
                        double b0 = 0.0;
                        for (int w = 0; w < TC_ROWS / 32; ++w) b0 += red_s[w];
                        __syncthreads();
                        if (writer && (tid & 31) == 0) red_s[quad] = v1;
                        __syncthreads();
                        double b1 = 0.0;
                        for (int w = 0; w < TC_ROWS / 32; ++w) b1 += red_s[w];
                        __syncthreads();
                        e0 = grid_sum(b0, partials, grid, gs_parity);
                        e1 = grid_sum(b1, partials, grid, gs_parity);
                        n = (double)D * (double)B;
                    }
                    const double d0 = (double)sqrtf((float)(e0 / n));
                    d1n = (double)sqrtf((float)(e1 / n));
                    dt0 = (d0 < 1e-5 || d1n < 1e-5) ? 1e-6 : 0.01 * (d0 / d1n);
                    dt0 = fmin(dt0, dtmax);
                    phase = PH_INITDT;
                } else {
                    phase = PH_STAGE;
                    stage = 1;
                    if (active && (!(dt > 0.0) || !isfinite(dt))) { ret = RET_DTLESSTHANMIN; active = false; }
                }
            } else if (phase == PH_INITDT) {
                float k1[16];
                tc_ld16(KS, k1);
                float s2 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (i < D) {
                        const float sk = fmaf(fabsf(u[i]), reltol, abstol);
                        const float a = (kout[i] - k1[i]) / sk;
                        s2 = fmaf(a, a, s2);
                    }
                double e2 = s2, n = D;
                if (GLOBAL) {
                    double v = (live && writer) ? e2 : 0.0;
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if (writer && (tid & 31) == 0) red_s[quad] = v;
                    __syncthreads();
                    double bsum = 0.0;
                    for (int w = 0; w < TC_ROWS / 32; ++w) bsum += red_s[w];
                    __syncthreads();
                    e2 = grid_sum(bsum, partials, grid, gs_parity);
                    n = (double)D * (double)B;
                }
                const double d2 = (double)sqrtf((float)(e2 / n)) / dt0;
                if (dt0 < 10.0 * 2.220446049250313e-16) {
                    dt = fmax(1e-6, dtmin);
                } else {
                    const double m = fmax(d1n, d2);
                    const double dt1 = (m <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(m)) / 5.0);
                    dt = fmax(dtmin, fmin(100.0 * dt0, fmin(dt1, dtmax)));
                }
                phase = PH_STAGE;
                stage = 1;
                if (active && (!(dt > 0.0) || !isfinite(dt))) { ret = RET_DTLESSTHANMIN; active = false; }
            } else if (stage < 6) {
                tc_st16(KS + 16 * stage, kout);  // k_{stage+1}
                tc_wait_st();
                ++stage;
            } else {
                // ---- end of an attempted step: g = u_{n+1}, kout = k7 -------------------------------------------------
                // dense-output coefficients and the error estimate from k1..k7 in one pass over the TMEM slots
                float c2[16], c3[16], c4[16], et[16], k1v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    c2[i] = c_r[6][1] * kout[i]; c3[i] = c_r[6][2] * kout[i]; c4[i] = c_r[6][3] * kout[i];
                    et[i] = c_bt[6] * kout[i];
                }
                for (int q = 0; q < 6; ++q) {
                    float kq[16];
                    tc_ld16(KS + 16 * q, kq);
                    const float r2 = c_r[q][1], r3 = c_r[q][2], r4 = c_r[q][3], bt = c_bt[q];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        c2[i] = fmaf(r2, kq[i], c2[i]); c3[i] = fmaf(r3, kq[i], c3[i]); c4[i] = fmaf(r4, kq[i], c4[i]);
                        et[i] = fmaf(bt, kq[i], et[i]);
                        if (q == 0) k1v[i] = kq[i];
                    }
                }
                const float h = (float)dts;
                bool finite = true;
                float e2f = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (i < D) {
                        finite = finite && isfinite(g[i]);
                        const float sk = fmaf(fmaxf(fabsf(u[i]), fabsf(g[i])), reltol, abstol);
                        const float a = h * et[i] / sk;
                        e2f = fmaf(a, a, e2f);
                    }
                bool accept = true;
                double dt_next = dt;
                if (o.adaptive) {
                    double e2 = e2f, n = D;
                    if (GLOBAL) {
                        double v = (active && writer) ? e2 : 0.0;
                        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                        if (writer && (tid & 31) == 0) red_s[quad] = v;
                        __syncthreads();
                        double bsum = 0.0;
                        for (int w = 0; w < TC_ROWS / 32; ++w) bsum += red_s[w];
                        __syncthreads();
                        e2 = grid_sum(bsum, partials, grid, gs_parity);
                        n = (double)D * (double)B;
                    }
                    const double EEst = (double)sqrtf((float)(e2 / n));
                    if (EEst != EEst) finite = false;
                    accept = pi_controller(o, EEst, dts, dtmax, pist, dt_next);
                }
                // every thread of the warp has read k1..k6 by now; the FSAL store below may overwrite slot 0 only
                // after the other half has read it as well
                tc_fence_before();
                __syncthreads();
                tc_fence_after();
                if (active) {
                    if (!finite) {
                        ret = RET_UNSTABLE;
                        active = false;
                    } else {
                        if (accept) {
                            if (writer && tape.cap > 0 && na < tape.cap) {
                                tape.t[(size_t)na * B + b] = t;
                                tape.dt[(size_t)na * B + b] = dts;
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (i < D) tape.u[((size_t)na * B + b) * D + i] = u[i];
                            }
                            ++na;
                            // saveat through the dense interpolant (Horner form)
                            const double inv = 1.0 / dts;
                            while (ks < T && tg[ks] <= tnew) {
                                if (writer) {
                                    const double tsv = tg[ks];
                                    float* dst = traj + ((size_t)ks * B + b) * D;
                                    if (tsv == tnew) {
#pragma unroll
                                        for (int i = 0; i < 16; ++i)
                                            if (i < D) dst[i] = g[i];
                                    } else {
                                        const float th = (float)((tsv - t) * inv);
                                        const float hth = h * th;
#pragma unroll
                                        for (int i = 0; i < 16; ++i)
                                            if (i < D) dst[i] = fmaf(hth, fmaf(th, fmaf(th, fmaf(th, c4[i], c3[i]), c2[i]), k1v[i]), u[i]);
                                    }
                                }
                                ++ks;
                            }
                            t = tnew;
#pragma unroll
                            for (int i = 0; i < 16; ++i) u[i] = g[i];
                            if (ks >= T) active = false;
                        } else {
                            ++nr;
                        }
                        dt = dt_next;
                        if (active && o.adaptive && (!(fabs(dt) > dtmin) || !isfinite(dt))) { ret = RET_DTLESSTHANMIN; active = false; }
                    }
                }
                // FSAL: k7 becomes k1 of the next step for the rows that accepted, the others keep their k1
                // (tcgen05.st is warp-aligned: every thread executes it, the DATA is selected per row)
                {
                    const bool take = accept && finite;
#pragma unroll
                    for (int i = 0; i < 16; ++i) k1v[i] = take ? kout[i] : k1v[i];
                    tc_st16(KS, k1v);
                    tc_wait_st();
                }
                stage = 1;
                tc_fence_before();
                all_done = __syncthreads_or(active ? 1 : 0) == 0;
                tc_fence_after();
            }
            if (all_done) break;
            if (phase != PH_STAGE && T <= 1) break;
        }
        if (live && writer) {
            if (ret != RET_SUCCESS) {
                for (int k = 0; k < T; ++k)
                    for (int i = 0; i < D; ++i) traj[((size_t)k * B + b) * D + i] = __int_as_float(0x7fc00000);
            }
            if (retcode) retcode[b] = ret;
            if (naccept) naccept[b] = na;
            if (nreject) nreject[b] = nr;
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS));
}

}  // namespace ldeq

using namespace ldeq;

// Host side: called from ldeq_mlp_solve_fwd when opts->mlp_math == LDEQ_MLP_MATH_BF16X3.
// Returns LDEQ_ERR_UNSUPPORTED for shapes this kernel is not built for (the caller reports it; no silent fallback).
int ldeq_mlp_tc_forward(ldeq_handle* h, const int32_t* dims, int n_layers, const float* params, const float* z0,
                        const double* d_tgrid, int B, int T, const KOpts& ko, int norm_mode, float* traj, int32_t* ret,
                        int32_t* na, int32_t* nr, double* tape_t, double* tape_dt, float* tape_u, int tape_cap,
                        cudaStream_t s) {
    if (n_layers != 3) return set_err(h, LDEQ_ERR_UNSUPPORTED, "bf16x3 tensor-core path: exactly 3 dense layers");
    const int D = dims[0], H1 = dims[1], H2 = dims[2];
    auto pad16 = [](int x) { return (x + 15) / 16 * 16; };
    if (D > 16 || pad16(H1) > TC_MAXW || pad16(H2) > TC_MAXW)
        return set_err(h, LDEQ_ERR_UNSUPPORTED, "bf16x3 tensor-core path: state dim <= 16, hidden widths <= 208");
    TcNet net;
    net.d = D; net.n1 = pad16(H1); net.n2 = pad16(H2);
    int off = 0;
    const int sizes[3] = {16 * net.n1 * 2, net.n1 * net.n2 * 2, net.n2 * 16 * 2};
    for (int l = 0; l < 3; ++l) {
        net.img_off[2 * l] = off; off += (sizes[l] + 127) / 128 * 128;
        net.img_off[2 * l + 1] = off; off += (sizes[l] + 127) / 128 * 128;
    }
    net.bias_off = off; off += (net.n1 + net.n2 + 16) * 4;
    net.smem_bytes = (off + 15) / 16 * 16;
    if (net.smem_bytes > 220 * 1024) return set_err(h, LDEQ_ERR_UNSUPPORTED, "bf16x3 tensor-core path: weights do not fit in shared memory");
    int rc = ensure_scratch(h, 2, (size_t)net.smem_bytes);
    if (rc) return rc;
    rc = ensure_scratch(h, 0, sizeof(double) * 4096);
    if (rc) return rc;
    unsigned char* img = (unsigned char*)h->scratch[2];
    mlp_tc_prep_kernel<<<64, 256, 0, s>>>(net, params, D, H1, H2, img);
    LDEQ_CUDA(cudaGetLastError());
    const int tiles = (B + TC_ROWS - 1) / TC_ROWS;
    MlpTapeViewTc<float> tv{tape_t, tape_dt, tape_u, tape_cap};
    double* partials = (double*)h->scratch[0];
    KOpts kov = ko;
    const unsigned char* imgc = img;
    void* args[] = {&net, &imgc, &z0, &d_tgrid, &B, &T, &kov, &traj, &ret, &na, &nr, &tv, &partials};
    // fixed-step mode computes no error norm: no grid barrier, no co-residency limit
    if (norm_mode == LDEQ_NORM_GLOBAL && ko.adaptive) {
        if (tiles > h->sm_count)
            return set_err(h, LDEQ_ERR_UNSUPPORTED, "bf16x3 tensor-core path, global norm: B <= 128 * SM count (all tiles co-resident)");
        LDEQ_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, net.smem_bytes));
        LDEQ_CUDA(cudaLaunchCooperativeKernel((void*)mlp_tc_fwd_kernel<true>, dim3(tiles), dim3(TC_THREADS), args, net.smem_bytes, s));
    } else {
        const int grid = tiles < h->sm_count ? tiles : h->sm_count;
        LDEQ_CUDA(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, net.smem_bytes));
        LDEQ_CUDA(cudaLaunchKernel((void*)mlp_tc_fwd_kernel<false>, dim3(grid), dim3(TC_THREADS), args, net.smem_bytes, s));
    }
    h->launches += 2;
    return LDEQ_OK;
}
