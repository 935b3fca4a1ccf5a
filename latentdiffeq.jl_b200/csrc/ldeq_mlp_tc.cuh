// ldeq_mlp_tc.cuh -- shared device code of the tensor-core (tcgen05 + TMEM) LatentODE kernels: PTX wrappers, the
// bf16 hi/lo operand split, shared-memory matrix descriptors, the weight-image builder and the layer issue / epilogue
// helpers.  Used by ldeq_mlp_tc.cu (forward solve) and ldeq_mlp_tc_bwd.cu (discrete adjoint + weight gradients).
#pragma once

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

static __device__ __constant__ float c_a[7][6] = {
    {0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {0.161f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {-0.008480655492356989f, 0.335480655492357f, 0.f, 0.f, 0.f, 0.f},
    {2.8971530571054935f, -6.359448489975075f, 4.3622954328695815f, 0.f, 0.f, 0.f},
    {5.325864828439257f, -11.748883564062828f, 7.4955393428898365f, -0.09249506636175525f, 0.f, 0.f},
    {5.86145544294642f, -12.92096931784711f, 8.159367898576159f, -0.071584973281401f, -0.028269050394068383f, 0.f},
    {0.09646076681806523f, 0.01f, 0.4798896504144996f, 1.379008574103742f, -3.290069515436081f, 2.324710524099774f}};
static __device__ __constant__ float c_bt[7] = {-0.00178001105222577714f, -0.0008164344596567469f, 0.007880878010261995f,
                                         -0.1447110071732629f,      0.5823571654525552f,     -0.45808210592918697f,
                                         0.015151515151515152f};
static __device__ __constant__ float c_r[7][4] = {{1.0f, -2.763706197274826f, 2.9132554618219126f, -1.0530884977290216f},
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
// general form: leading-dimension / stride-dimension byte offsets given explicitly (an MN-major operand swaps their roles:
// LBO = stride between 8-row K groups, SBO = stride between 8-element MN chunks; verified on the device, scripts/probe/)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = BF16 (bits 7, 10), K-major both, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---- weight images -----------------------------------------------------------------------------------------
// flat Flux parameters -> bf16 hi/lo images in the canonical layout + padded biases, in global memory
static __global__ void mlp_tc_prep_kernel(TcNet net, const float* __restrict__ P, int d_in, int h1, int h2, unsigned char* __restrict__ img) {
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
        // the first padding column of a hidden layer carries bias 1 (its weights are zero): the activation there is the
        // constant 1, which turns the bias gradients into one more column of the weight-gradient GEMMs (ldeq_mlp_tc_bwd.cu);
        // the next layer's image has zero rows for the padding, so the forward pass is unaffected
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += gridDim.x * blockDim.x)
            bias[i] = i < Nr[l] ? P[b_off + i] : (l < 2 && i == Nr[l]) ? 1.f : 0.f;
        w_off = b_off + Nr[l];
    }
}

template <class S> struct MlpTapeViewTc {
    double* t;
    double* dt;
    S* u;
    int cap;
};

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

// instruction descriptor with explicit majors (bit 15: A is MN-major, bit 16: B is MN-major)
__device__ __forceinline__ uint32_t make_idesc_mn(int n, int a_mn, int b_mn) {
    return make_idesc(n) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// The same weight image read TRANSPOSED: D[:, n_lo:n_lo+n] = A (K' = 16*ksteps) x W^T, where the image was built for the
// forward product (N = out, K = in, K-major, lbo_fwd = (N/8)*128).  As the B operand of the transposed product the roles
// swap -- N' = in, K' = out -- and the very same bytes are the MN-major canonical layout with LBO' = 128 (between 8-row K'
// groups) and SBO' = lbo_fwd (between 8-element N' chunks): no second set of images is needed for the reverse pass.
__device__ __forceinline__ void tc_issue_layer_T(uint32_t d_addr, uint32_t a_addr, uint32_t w_hi, uint32_t w_lo, uint32_t lbo_fwd,
                                                 int n_lo, int n, int ksteps) {
    if (n <= 0) return;
    const uint32_t idesc = make_idesc_mn(n, 0, 1);
    const uint32_t col_off = (uint32_t)(n_lo / 8) * lbo_fwd;
    for (int kk = 0; kk < ksteps; ++kk) {
        const uint64_t bh = make_smem_desc(w_hi + col_off + kk * 256, 128, lbo_fwd), bl = make_smem_desc(w_lo + col_off + kk * 256, 128, lbo_fwd);
        tc_mma_ts(d_addr + n_lo, a_addr + 16 * kk, bh, idesc, kk > 0);
        tc_mma_ts(d_addr + n_lo, a_addr + 16 * kk, bl, idesc, 1);
        tc_mma_ts(d_addr + n_lo, a_addr + 16 * kk + 8, bh, idesc, 1);
    }
}

// ---- TMA (bulk async copy) staging of the weight images: global -> shared, completion on an mbarrier ------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one thread: the whole image in pieces of at most 32 KiB (sizes are multiples of 16 bytes)
__device__ __forceinline__ void stage_image_tma(unsigned char* smem_dst, const unsigned char* gsrc, uint32_t bytes, uint64_t* bar) {
    mbar_expect_tx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
        bulk_g2s(smem_dst + off, gsrc + off, n, bar);
    }
}

}  // namespace ldeq

// ---- host: geometry of the padded network and its weight images (shared by the forward and the reverse pass) -------------
// Hidden widths are padded to a multiple of 16 with at least ONE spare column (the constant-1 column of the bias gradients).
static inline int ldeq_tc_make_net(ldeq_handle* h, const int32_t* dims, int n_layers, ldeq::TcNet* net) {
    using namespace ldeq;
    if (n_layers != 3) return set_err(h, LDEQ_ERR_UNSUPPORTED, "bf16x3 tensor-core path: exactly 3 dense layers");
    const int D = dims[0], H1 = dims[1], H2 = dims[2];
    auto pad16 = [](int x) { return (x + 15) / 16 * 16; };
    if (D > 16 || pad16(H1 + 1) > TC_MAXW || pad16(H2 + 1) > TC_MAXW)
        return set_err(h, LDEQ_ERR_UNSUPPORTED, "bf16x3 tensor-core path: state dim <= 16, hidden widths <= 207");
    net->d = D; net->n1 = pad16(H1 + 1); net->n2 = pad16(H2 + 1);
    int off = 0;
    const int sizes[3] = {16 * net->n1 * 2, net->n1 * net->n2 * 2, net->n2 * 16 * 2};
    for (int l = 0; l < 3; ++l) {
        net->img_off[2 * l] = off; off += (sizes[l] + 127) / 128 * 128;
        net->img_off[2 * l + 1] = off; off += (sizes[l] + 127) / 128 * 128;
    }
    net->bias_off = off; off += (net->n1 + net->n2 + 16) * 4;
    net->smem_bytes = (off + 15) / 16 * 16;
    if (net->smem_bytes > 220 * 1024) return set_err(h, LDEQ_ERR_UNSUPPORTED, "bf16x3 tensor-core path: weights do not fit in shared memory");
    return LDEQ_OK;
}
