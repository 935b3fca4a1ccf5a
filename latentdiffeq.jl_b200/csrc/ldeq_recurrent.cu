// ldeq_recurrent.cu -- the recurrent pattern extractor as persistent kernels.
//
// Reference: apply_pattern_extractor (src/models/GOKU.jl:30-49, src/models/LatentODE.jl:20-34) runs three two-layer
// recurrent stacks over the 50 frames of every sequence -- Chain(RNN(F,H,relu), RNN(H,H,relu)) on the reversed
// sequence, Chain(LSTM(F,H), LSTM(H,H)) forwards and a second one on the reversed sequence (GOKU.jl:224-234; F = 32,
// H = 16 by default) -- keeps only the final hidden state of each and resets the states (`Flux.reset!`).  In Flux that
// is T dependent cell applications per stack, each a handful of tiny BLAS calls [3P Flux 0.13.6 recurrent.jl: RNNCell
// h' = s(Wi x + Wh h + b); LSTMCell gates = Wi x + Wh h + b, input/forget/cell/output = sigm, sigm, tanh, sigm,
// c' = f c + i g, h' = o tanh(c')].
//
// Here one kernel launch integrates a whole stack (both layers, all T steps):
//   * 4 lanes per sequence, each owning 4 of the 16 hidden units of both layers (their gate rows, their c), the
//     sequence's hidden vectors are re-assembled with warp shuffles; 32 sequences per CTA;
//   * the stack's weights are staged ONCE into shared memory in the order the lanes consume them (input-major, the
//     16 rows of a lane contiguous, lane groups 20 floats apart so that one 128-bit load per lane group is conflict-free);
//   * the forward pass keeps h (and c) of every step on a tape; the reverse pass (back-propagation through time)
//     recomputes the gates from the taped states, pulls the cotangents through the transposed weights (reduce-scatter
//     over the 4 lanes), writes d x, and accumulates the weight gradients in REGISTERS across all T steps: after every
//     step the CTA parks its 32 x (deltas, inputs) in shared memory and each thread adds its 4 x 6 tile of the
//     (rows x inputs) outer product; every CTA stores its partial gradient and a second kernel sums the CTAs in a
//     fixed order (deterministic: no floating-point atomics anywhere);
//   * the three stacks run in ONE launch per pass (blockIdx.y), each with its own d x buffer, added in a fixed order.
// Parameters travel as one flat Float32 vector per stack in `Flux.destructure` order: per layer Wi (rows x in,
// column-major), Wh (rows x H), b (rows), state0 (H) [LSTM: h0 (H), c0 (H)]; rows = H (RNN) or 4H (LSTM, gate-major).
#include <cstring>

#include "ldeq_internal.h"

namespace ldeq {

constexpr int PE_H = 16;        // hidden units per layer (rnn_output_dim default, GOKU.jl:201)
constexpr int PE_LANES = 4;     // lanes per sequence
constexpr int PE_SPB = 32;      // sequences per CTA
constexpr int PE_THREADS = PE_LANES * PE_SPB;

template <int G> struct PeDims {
    static constexpr int RL = 4 * G;       // gate rows a lane owns per layer (4 units x G gates)
    static constexpr int RG = RL + 4;      // stride between lane groups in the image (bank-conflict padding)
    static constexpr int RS = 4 * RG + 4;  // image stride per input (+4: columns j, j+1, j+2, j+3 fall into different banks)
    static constexpr int R = 16 * G;       // rows of a layer
    static constexpr int NS = G == 4 ? 64 : 32;  // taped floats per sequence and step: h1 [c1] h2 [c2]
};

__host__ __device__ inline int pe_layer_params(int G, int in) { return 16 * G * in + 16 * G * PE_H + 16 * G + (G == 4 ? 2 : 1) * PE_H; }
__host__ __device__ inline int pe_stack_params(int G, int F) { return pe_layer_params(G, F) + pe_layer_params(G, PE_H); }

// image row of (lane group g, local row r) <-> row of the Flux matrices: gate * H + 4 g + k, r = gate * 4 + k
template <int G> __device__ __forceinline__ int pe_flux_row(int g, int r) { return (r >> 2) * PE_H + 4 * g + (r & 3); }

// Stage one layer: img[j * RS + g * RG + r] = [Wi | Wh](row, j); bias[g * RG + r]
template <int G>
__device__ void pe_stage_layer(const float* __restrict__ p, int in, float* img, float* bias) {
    using D = PeDims<G>;
    const int IN = in + PE_H;
    const float* Wi = p;
    const float* Wh = p + D::R * in;
    const float* b = Wh + D::R * PE_H;
    for (int e = threadIdx.x; e < IN * 4 * D::RL; e += blockDim.x) {
        const int j = e / (4 * D::RL), q = e - j * 4 * D::RL, g = q / D::RL, r = q - g * D::RL;
        const int row = pe_flux_row<G>(g, r);
        img[j * D::RS + g * D::RG + r] = j < in ? Wi[j * D::R + row] : Wh[(j - in) * D::R + row];
    }
    for (int e = threadIdx.x; e < 4 * D::RL; e += blockDim.x) {
        const int g = e / D::RL, r = e - g * D::RL;
        bias[g * D::RG + r] = b[pe_flux_row<G>(g, r)];
    }
}

__device__ __forceinline__ float pe_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// acc[r] = bias + sum_j W[row r][j] in[j] for the lane's RL rows; `a` has NA entries, `hb` the 16 recurrent ones
template <int G, int NA>
__device__ __forceinline__ void pe_gates(const float* __restrict__ img, const float* __restrict__ bias, int g, const float* a, const float* hb,
                                         float* acc) {
    using D = PeDims<G>;
    const float* base = img + g * D::RG;
#pragma unroll
    for (int q = 0; q < G; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + g * D::RG + 4 * q);
        acc[4 * q] = b4.x; acc[4 * q + 1] = b4.y; acc[4 * q + 2] = b4.z; acc[4 * q + 3] = b4.w;
    }
#pragma unroll
    for (int j = 0; j < NA + PE_H; ++j) {
        const float v = j < NA ? a[j] : hb[j - NA];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(base + j * D::RS + 4 * q);
            acc[4 * q] = fmaf(w.x, v, acc[4 * q]);
            acc[4 * q + 1] = fmaf(w.y, v, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w.z, v, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(w.w, v, acc[4 * q + 3]);
        }
    }
}

// every lane of a sequence gets all 16 values from the 4 owned by each lane (unit 4 q + k lives in lane q)
__device__ __forceinline__ void pe_gather(const float* own, float* full) {
    const int lane0 = (threadIdx.x & 31) & ~3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) full[4 * q + k] = __shfl_sync(0xffffffffu, own[k], lane0 | q);
}

// one cell update from the lane's gate pre-activations
template <int G>
__device__ __forceinline__ void pe_cell(const float* acc, float* h, float* c) {
    if constexpr (G == 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float i = pe_sigmoid(acc[k]), f = pe_sigmoid(acc[4 + k]), gg = tanhf(acc[8 + k]), o = pe_sigmoid(acc[12 + k]);
            c[k] = fmaf(f, c[k], i * gg);
            h[k] = o * tanhf(c[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = fmaxf(acc[k], 0.f);
    }
}

template <int G, int F> struct PeSmem {
    using D = PeDims<G>;
    static constexpr int IMG1 = (F + PE_H) * D::RS, IMG2 = (2 * PE_H) * D::RS, BIAS = 4 * D::RG;
    static constexpr int WEIGHTS = IMG1 + IMG2 + 2 * BIAS;
    // reverse pass staging, per sequence: delta1 (R) | in1 (F + H) | delta2 (R) | in2 (2 H)
    static constexpr int STG = 2 * D::R + F + 3 * PE_H + 4;   // + 4: the rows of the 8 sequences of a warp start in different banks
    static constexpr size_t fwd_bytes = (size_t)WEIGHTS * 4;
    static constexpr size_t bwd_bytes = (size_t)(WEIGHTS + 2 * PE_SPB * STG) * 4;
};

// initial states of a layer in the flat vector: [h0] or [h0, c0]
template <int G> __device__ __forceinline__ const float* pe_state0(const float* layer_params, int in) {
    return layer_params + 16 * G * in + 16 * G * PE_H + 16 * G;
}

// ---- forward ------------------------------------------------------------------------------------------------------
// x (F,B,T) = [T][B][F]; reverse: the stack reads frame T-1-s at step s (GOKU.jl:39).  out: final h of layer 2 into
// out[b * ostride + ooff + unit].  tape (may be null): [T][B][NS] = h1 [c1] h2 [c2] after every step.
template <int G, int F>
__device__ __forceinline__ void
pe_fwd_body(const float* __restrict__ x, int B, int T, int reverse, const float* __restrict__ params, float* __restrict__ out, int ostride,
            int ooff, float* __restrict__ tape) {
    using D = PeDims<G>;
    using SM = PeSmem<G, F>;
    extern __shared__ __align__(16) float pe_smem[];
    float* img1 = pe_smem;
    float* img2 = img1 + SM::IMG1;
    float* bias1 = img2 + SM::IMG2;
    float* bias2 = bias1 + SM::BIAS;
    const float* p1 = params;
    const float* p2 = params + pe_layer_params(G, F);
    pe_stage_layer<G>(p1, F, img1, bias1);
    pe_stage_layer<G>(p2, PE_H, img2, bias2);
    __syncthreads();

    const int g = threadIdx.x & 3;
    const int b = blockIdx.x * PE_SPB + (threadIdx.x >> 2);
    const bool live = b < B;
    const int bb = live ? b : B - 1;
    float h1[4], c1[4], h2[4], c2[4];
    {
        const float* s1 = pe_state0<G>(p1, F);
        const float* s2 = pe_state0<G>(p2, PE_H);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            h1[k] = s1[4 * g + k];
            h2[k] = s2[4 * g + k];
            c1[k] = G == 4 ? s1[PE_H + 4 * g + k] : 0.f;
            c2[k] = G == 4 ? s2[PE_H + 4 * g + k] : 0.f;
        }
    }
    for (int s = 0; s < T; ++s) {
        const int frame = reverse ? T - 1 - s : s;
        float xin[F], hf[PE_H], acc[D::RL];
        const float4* xp = reinterpret_cast<const float4*>(x + ((size_t)frame * B + bb) * F);
#pragma unroll
        for (int i = 0; i < F / 4; ++i) {
            const float4 v = __ldg(xp + i);
            xin[4 * i] = v.x; xin[4 * i + 1] = v.y; xin[4 * i + 2] = v.z; xin[4 * i + 3] = v.w;
        }
        pe_gather(h1, hf);
        pe_gates<G, F>(img1, bias1, g, xin, hf, acc);
        pe_cell<G>(acc, h1, c1);
        float h1f[PE_H];
        pe_gather(h1, h1f);
        pe_gather(h2, hf);
        pe_gates<G, PE_H>(img2, bias2, g, h1f, hf, acc);
        pe_cell<G>(acc, h2, c2);
        if (tape && live) {
            float* tp = tape + ((size_t)s * B + b) * D::NS + 4 * g;
            *reinterpret_cast<float4*>(tp) = make_float4(h1[0], h1[1], h1[2], h1[3]);
            if constexpr (G == 4) {
                *reinterpret_cast<float4*>(tp + 16) = make_float4(c1[0], c1[1], c1[2], c1[3]);
                *reinterpret_cast<float4*>(tp + 32) = make_float4(h2[0], h2[1], h2[2], h2[3]);
                *reinterpret_cast<float4*>(tp + 48) = make_float4(c2[0], c2[1], c2[2], c2[3]);
            } else {
                *reinterpret_cast<float4*>(tp + 16) = make_float4(h2[0], h2[1], h2[2], h2[3]);
            }
        }
    }
    if (live) {
#pragma unroll
        for (int k = 0; k < 4; ++k) out[(size_t)b * ostride + ooff + 4 * g + k] = h2[k];
    }
}

// ---- reverse pass (back-propagation through time) ---------------------------------------------------------------------
// The reverse pass works through a per-sequence row of shared memory instead of registers and shuffles (a fully
// unrolled register formulation of the four dense products of a step spills): lane g owns the units u = 4 k + g
// (interleaved, so that the columns j = 4 i + g of the transposed products it evaluates are its own units), the
// sequence's inputs and pre-activation cotangents are parked in the row
//     [ delta1 (R, image order) | in1 = x_t, h1_{t-1} (F + H) | delta2 (R) | in2 = h1_t, h2_{t-1} (2 H) ]
// where all four lanes -- and, after the step's barrier, the weight-gradient tiles of the whole CTA -- read them.
template <int G> __device__ __forceinline__ int pe_flux_row_il(int g, int r) { return (r >> 2) * PE_H + 4 * (r & 3) + g; }

template <int G>
__device__ void pe_stage_layer_il(const float* __restrict__ p, int in, float* img, float* bias) {
    using D = PeDims<G>;
    const int IN = in + PE_H;
    const float* Wi = p;
    const float* Wh = p + D::R * in;
    const float* b = Wh + D::R * PE_H;
    for (int e = threadIdx.x; e < IN * 4 * D::RL; e += blockDim.x) {
        const int j = e / (4 * D::RL), q = e - j * 4 * D::RL, g = q / D::RL, r = q - g * D::RL;
        const int row = pe_flux_row_il<G>(g, r);
        img[j * D::RS + g * D::RG + r] = j < in ? Wi[j * D::R + row] : Wh[(j - in) * D::R + row];
    }
    for (int e = threadIdx.x; e < 4 * D::RL; e += blockDim.x) {
        const int g = e / D::RL, r = e - g * D::RL;
        bias[g * D::RG + r] = b[pe_flux_row_il<G>(g, r)];
    }
}

// acc[r] = bias + sum_j W[own row r][j] in[j], inputs read from the sequence's shared-memory row (IN a multiple of 4)
template <int G>
__device__ __forceinline__ void pe_gates_s(const float* __restrict__ img, const float* __restrict__ bias, int g, const float* __restrict__ in, int IN,
                                           float* acc) {
    using D = PeDims<G>;
    const float* base = img + g * D::RG;
#pragma unroll
    for (int q = 0; q < G; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + g * D::RG + 4 * q);
        acc[4 * q] = b4.x; acc[4 * q + 1] = b4.y; acc[4 * q + 2] = b4.z; acc[4 * q + 3] = b4.w;
    }
#pragma unroll 2
    for (int j4 = 0; j4 < IN / 4; ++j4) {
        const float4 v4 = *reinterpret_cast<const float4*>(in + 4 * j4);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
            for (int q = 0; q < G; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(base + (4 * j4 + jj) * D::RS + 4 * q);
                acc[4 * q] = fmaf(w.x, v[jj], acc[4 * q]);
                acc[4 * q + 1] = fmaf(w.y, v[jj], acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(w.z, v[jj], acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(w.w, v[jj], acc[4 * q + 3]);
            }
        }
    }
}

// (W^T delta)[j] over ALL rows of the layer for one input column j; dall = the sequence's delta in image order
template <int G>
__device__ __forceinline__ float pe_col_dot(const float* __restrict__ img, int j, const float* dall) {
    using D = PeDims<G>;
    const float* col = img + j * D::RS;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const float4 w = *reinterpret_cast<const float4*>(col + gq * D::RG + 4 * q);
            const float* d = dall + gq * D::RL + 4 * q;
            s0 = fmaf(w.x, d[0], s0);
            s1 = fmaf(w.y, d[1], s1);
            s0 = fmaf(w.z, d[2], s0);
            s1 = fmaf(w.w, d[3], s1);
        }
    }
    return s0 + s1;
}

// cotangent of one cell at one step: from (dh, dc) of its outputs and the recomputed gates to the pre-activation
// cotangents delta (lane's RL rows); dc is replaced by the cotangent of c_{t-1}
template <int G>
__device__ __forceinline__ void pe_cell_bwd(const float* acc, const float* h_t, const float* c_t, const float* c_prev, const float* dh, float* dc,
                                            float* delta) {
    if constexpr (G == 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float i = pe_sigmoid(acc[k]), f = pe_sigmoid(acc[4 + k]), gg = tanhf(acc[8 + k]), o = pe_sigmoid(acc[12 + k]);
            const float tc = tanhf(c_t[k]);
            const float dct = fmaf(dh[k] * o, 1.f - tc * tc, dc[k]);
            delta[k] = dct * gg * i * (1.f - i);
            delta[4 + k] = dct * c_prev[k] * f * (1.f - f);
            delta[8 + k] = dct * i * (1.f - gg * gg);
            delta[12 + k] = dh[k] * tc * o * (1.f - o);
            dc[k] = dct * f;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) delta[k] = h_t[k] > 0.f ? dh[k] : 0.f;
    }
}

// store a thread's (G rows x TI inputs) weight-gradient tile into the CTA's partial of one layer (Wi | Wh in Flux order)
template <int G, int TI>
__device__ __forceinline__ void pe_flush_tile(const float (&tile)[G][TI], float* gl, int in, int rg, int ig) {
    using D = PeDims<G>;
    float* gWi = gl;
    float* gWh = gl + D::R * in;
#pragma unroll
    for (int r = 0; r < G; ++r) {
        const int ri = rg * G + r;                       // image row = lane group * RL + local row
        const int row = pe_flux_row_il<G>(ri / D::RL, ri % D::RL);
#pragma unroll
        for (int i = 0; i < TI; ++i) {
            const int j = ig * TI + i;
            if (j < in) gWi[(size_t)j * D::R + row] = tile[r][i];      // every (row, input) pair has one owner thread in the CTA
            else gWh[(size_t)(j - in) * D::R + row] = tile[r][i];
        }
    }
}

// dx: this stack's own cotangent buffer of x (the three stacks of the pattern extractor read the same frames; their
// buffers are added in a fixed order afterwards).  gpart: [n_cta][n_params] partial gradients, row blockIdx.x written here.
// dout: cotangent of the final h of layer 2.
template <int G, int F>
__device__ __forceinline__ void
pe_bwd_body(const float* __restrict__ x, int B, int T, int reverse, const float* __restrict__ params, const float* __restrict__ tape,
            const float* __restrict__ dout, int ostride, int ooff, float* __restrict__ dx, float* __restrict__ gpart) {
    using D = PeDims<G>;
    using SM = PeSmem<G, F>;
    constexpr int IN1 = F + PE_H, IN2 = 2 * PE_H;
    constexpr int TI1 = IN1 / 8, TI2 = IN2 / 8;   // inputs per thread tile; rows per tile = G
    constexpr int O_D1 = 0, O_IN1 = D::R, O_D2 = D::R + IN1, O_IN2 = 2 * D::R + IN1;
    extern __shared__ __align__(16) float pe_smem[];
    float* img1 = pe_smem;
    float* img2 = img1 + SM::IMG1;
    float* bias1 = img2 + SM::IMG2;
    float* bias2 = bias1 + SM::BIAS;
    float* stage = bias2 + SM::BIAS;   // [2][PE_SPB][STG]
    const float* p1 = params;
    const float* p2 = params + pe_layer_params(G, F);
    pe_stage_layer_il<G>(p1, F, img1, bias1);
    pe_stage_layer_il<G>(p2, PE_H, img2, bias2);
    __syncthreads();

    const int g = threadIdx.x & 3, sl = threadIdx.x >> 2;
    const int b = blockIdx.x * PE_SPB + sl;
    const bool live = b < B;
    const int bb = live ? b : B - 1;
    const float* s1 = pe_state0<G>(p1, F);
    const float* s2 = pe_state0<G>(p2, PE_H);

    // own units: u = 4 k + g
    float dh1[4], dc1[4], dh2[4], dc2[4], db1[D::RL], db2[D::RL];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dh1[k] = dc1[k] = dc2[k] = 0.f;
        dh2[k] = live ? dout[(size_t)b * ostride + ooff + 4 * k + g] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < D::RL; ++r) db1[r] = db2[r] = 0.f;
    // weight-gradient tiles of this thread: rows rg G .. rg G + G - 1 (image order), inputs ig TI .. ig TI + TI - 1
    const int rg = threadIdx.x >> 3, ig = threadIdx.x & 7;
    float gw1[G][TI1], gw2[G][TI2];
#pragma unroll
    for (int r = 0; r < G; ++r) {
#pragma unroll
        for (int i = 0; i < TI1; ++i) gw1[r][i] = 0.f;
#pragma unroll
        for (int i = 0; i < TI2; ++i) gw2[r][i] = 0.f;
    }

    // taped states of the own units after step s (s < 0: the trainable initial states)
    auto load_state = [&](int s, float* h1o, float* c1o, float* h2o, float* c2o) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int u = 4 * k + g;
            if (s < 0) {
                h1o[k] = s1[u];
                h2o[k] = s2[u];
                c1o[k] = G == 4 ? s1[PE_H + u] : 0.f;
                c2o[k] = G == 4 ? s2[PE_H + u] : 0.f;
            } else {
                const float* tp = tape + ((size_t)s * B + bb) * D::NS + u;
                h1o[k] = tp[0];
                if constexpr (G == 4) { c1o[k] = tp[16]; h2o[k] = tp[32]; c2o[k] = tp[48]; }
                else { h2o[k] = tp[16]; c1o[k] = 0.f; c2o[k] = 0.f; }
            }
        }
    };

    float h1t[4], c1t[4], h2t[4], c2t[4];
    load_state(T - 1, h1t, c1t, h2t, c2t);
    for (int s = T - 1; s >= 0; --s) {
        const int frame = reverse ? T - 1 - s : s;
        float h1p[4], c1p[4], h2p[4], c2p[4];
        load_state(s - 1, h1p, c1p, h2p, c2p);
        float* st = stage + ((size_t)(s & 1) * PE_SPB + sl) * SM::STG;   // this sequence's row
        // ---- park the step's inputs: in1 = [x_t | h1_{t-1}], in2 = [h1_t | h2_{t-1}]
        {
            const float4* xq = reinterpret_cast<const float4*>(x + ((size_t)frame * B + bb) * F) + g * (F / 16);
#pragma unroll
            for (int i = 0; i < F / 16; ++i) *reinterpret_cast<float4*>(st + O_IN1 + g * (F / 4) + 4 * i) = __ldg(xq + i);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                st[O_IN1 + F + 4 * k + g] = h1p[k];
                st[O_IN2 + 4 * k + g] = h1t[k];
                st[O_IN2 + PE_H + 4 * k + g] = h2p[k];
            }
        }
        __syncwarp();
        float acc[D::RL], delta[D::RL], dall[D::R];
        // ---- layer 2
        if constexpr (G == 4) pe_gates_s<G>(img2, bias2, g, st + O_IN2, IN2, acc);
        pe_cell_bwd<G>(acc, h2t, c2t, c2p, dh2, dc2, delta);
#pragma unroll
        for (int q = 0; q < G; ++q) *reinterpret_cast<float4*>(st + O_D2 + g * D::RL + 4 * q) = make_float4(delta[4 * q], delta[4 * q + 1], delta[4 * q + 2], delta[4 * q + 3]);
#pragma unroll
        for (int r = 0; r < D::RL; ++r) db2[r] += delta[r];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < D::R / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(st + O_D2 + 4 * q);
            dall[4 * q] = v.x; dall[4 * q + 1] = v.y; dall[4 * q + 2] = v.z; dall[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dh1[k] += pe_col_dot<G>(img2, 4 * k + g, dall);          // W_i2^T delta2: cotangent of h1_t, own units
            dh2[k] = pe_col_dot<G>(img2, PE_H + 4 * k + g, dall);    // W_h2^T delta2: cotangent of h2_{t-1}
        }
        // ---- layer 1
        if constexpr (G == 4) pe_gates_s<G>(img1, bias1, g, st + O_IN1, IN1, acc);
        pe_cell_bwd<G>(acc, h1t, c1t, c1p, dh1, dc1, delta);
#pragma unroll
        for (int q = 0; q < G; ++q) *reinterpret_cast<float4*>(st + O_D1 + g * D::RL + 4 * q) = make_float4(delta[4 * q], delta[4 * q + 1], delta[4 * q + 2], delta[4 * q + 3]);
#pragma unroll
        for (int r = 0; r < D::RL; ++r) db1[r] += delta[r];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < D::R / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(st + O_D1 + 4 * q);
            dall[4 * q] = v.x; dall[4 * q + 1] = v.y; dall[4 * q + 2] = v.z; dall[4 * q + 3] = v.w;
        }
        {
            float* dp = dx + ((size_t)frame * B + bb) * F;
#pragma unroll 2
            for (int i = 0; i < F / 4; ++i) {                          // W_i1^T delta1: cotangent of x, columns 4 i + g
                const float v = pe_col_dot<G>(img1, 4 * i + g, dall);
                if (live) dp[4 * i + g] = v;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dh1[k] = pe_col_dot<G>(img1, F + 4 * k + g, dall);   // W_h1^T delta1: cotangent of h1_{t-1}
#pragma unroll
        for (int k = 0; k < 4; ++k) { h1t[k] = h1p[k]; c1t[k] = c1p[k]; h2t[k] = h2p[k]; c2t[k] = c2p[k]; }
        __syncthreads();
        // ---- weight gradients: this thread's tiles of delta^T [inputs] over the CTA's 32 sequences
        {
            const float* sb = stage + (size_t)(s & 1) * PE_SPB * SM::STG;
#pragma unroll 4
            for (int q = 0; q < PE_SPB; ++q) {
                const float* row = sb + q * SM::STG;
                float d1[G], d2[G], a1[TI1], a2[TI2];
#pragma unroll
                for (int r = 0; r < G; ++r) { d1[r] = row[O_D1 + rg * G + r]; d2[r] = row[O_D2 + rg * G + r]; }
#pragma unroll
                for (int i = 0; i < TI1; ++i) a1[i] = row[O_IN1 + ig * TI1 + i];
#pragma unroll
                for (int i = 0; i < TI2; ++i) a2[i] = row[O_IN2 + ig * TI2 + i];
#pragma unroll
                for (int r = 0; r < G; ++r) {
#pragma unroll
                    for (int i = 0; i < TI1; ++i) gw1[r][i] = fmaf(d1[r], a1[i], gw1[r][i]);
#pragma unroll
                    for (int i = 0; i < TI2; ++i) gw2[r][i] = fmaf(d2[r], a2[i], gw2[r][i]);
                }
            }
        }
        // the row of parity (s & 1) is written again at step s - 2: every warp passes the barrier of step s - 1 only after it
        // has finished this loop
    }

    // ---- flush: this CTA's partial gradient (plain stores: deterministic); pe_reduce_kernel sums the CTAs in order
    float* g1 = gpart + (size_t)blockIdx.x * pe_stack_params(G, F);
    float* g2 = g1 + pe_layer_params(G, F);
    pe_flush_tile<G, TI1>(gw1, g1, F, rg, ig);
    pe_flush_tile<G, TI2>(gw2, g2, PE_H, rg, ig);
    // biases and initial states: sum over the 8 sequences of the warp (lanes with the same g), park the 4 warps' sums in
    // shared memory and add them in a fixed order
    auto warp_sum_same_g = [](float v) {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        return v;
    };
    constexpr int NRED = 2 * D::R + 4 * PE_H;   // db1 | db2 | dh1_0 | dh2_0 | dc1_0 | dc2_0 in Flux row / unit order
    __syncthreads();                            // the staging rows are free now
    float* red = stage;                         // [4 warps][NRED]
    const int warp = threadIdx.x >> 5;
    const bool writer = (threadIdx.x & 31) < 4;
#pragma unroll
    for (int r = 0; r < D::RL; ++r) {
        const float v1 = warp_sum_same_g(db1[r]), v2 = warp_sum_same_g(db2[r]);
        if (writer) {
            red[warp * NRED + pe_flux_row_il<G>(g, r)] = v1;
            red[warp * NRED + D::R + pe_flux_row_il<G>(g, r)] = v2;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = warp_sum_same_g(live ? dh1[k] : 0.f), c = warp_sum_same_g(live ? dh2[k] : 0.f);
        const float e = warp_sum_same_g(live ? dc1[k] : 0.f), f = warp_sum_same_g(live ? dc2[k] : 0.f);
        if (writer) {
            float* rw = red + warp * NRED + 2 * D::R;
            rw[4 * k + g] = a;
            rw[PE_H + 4 * k + g] = c;
            rw[2 * PE_H + 4 * k + g] = e;
            rw[3 * PE_H + 4 * k + g] = f;
        }
    }
    __syncthreads();
    float* gb1 = g1 + D::R * F + D::R * PE_H;
    float* gb2 = g2 + D::R * PE_H + D::R * PE_H;
    for (int i = threadIdx.x; i < NRED; i += PE_THREADS) {
        const float v = (red[i] + red[NRED + i]) + (red[2 * NRED + i] + red[3 * NRED + i]);
        if (i < D::R) gb1[i] = v;
        else if (i < 2 * D::R) gb2[i - D::R] = v;
        else {
            const int q = (i - 2 * D::R) / PE_H, u = (i - 2 * D::R) % PE_H;   // q: dh1_0, dh2_0, dc1_0, dc2_0
            if (q == 0) gb1[D::R + u] = v;
            else if (q == 1) gb2[D::R + u] = v;
            else if (G == 4) (q == 2 ? gb1 : gb2)[D::R + PE_H + u] = v;
        }
    }
}

// One launch per pass: blockIdx.y selects the stack (0: relu-RNN on the reversed sequence, 1: LSTM forwards,
// 2: LSTM on the reversed sequence; GOKU.jl:39-41).  The three stacks are independent, so at one GPU's share of a
// training batch (B = 8192: 256 CTAs per stack) they fill the machine together instead of one after the other.
struct PeStackArgs {
    const float* params[3];
    float* tape[3];
    float* out[3];      // forward: final states; reverse pass: cotangents of the final states (read)
    int ostride[3], ooff[3];
    float* dx[3];       // reverse pass: one cotangent buffer per stack (summed in a fixed order afterwards)
    float* gpart[3];    // reverse pass: [n_cta][n_params] partial gradients
};

template <int F>
__global__ void __launch_bounds__(PE_THREADS)
pe_fwd_kernel(const float* __restrict__ x, int B, int T, PeStackArgs a) {
    const int y = blockIdx.y;
    if (y == 0) pe_fwd_body<1, F>(x, B, T, 1, a.params[0], a.out[0], a.ostride[0], a.ooff[0], a.tape[0]);
    else pe_fwd_body<4, F>(x, B, T, y == 2, a.params[y], a.out[y], a.ostride[y], a.ooff[y], a.tape[y]);
}

template <int F>
__global__ void __launch_bounds__(PE_THREADS, 2)
pe_bwd_kernel(const float* __restrict__ x, int B, int T, PeStackArgs a) {
    const int y = blockIdx.y;
    if (y == 0) pe_bwd_body<1, F>(x, B, T, 1, a.params[0], a.tape[0], a.out[0], a.ostride[0], a.ooff[0], a.dx[0], a.gpart[0]);
    else pe_bwd_body<4, F>(x, B, T, y == 2, a.params[y], a.tape[y], a.out[y], a.ostride[y], a.ooff[y], a.dx[y], a.gpart[y]);
}

// dparams[p] = sum over the CTAs' partials in CTA order (deterministic); blockIdx.y = stack
struct PeReduceArgs {
    const float* gpart[3];
    float* dparams[3];
    int n[3];
};
__global__ void pe_reduce_kernel(PeReduceArgs a, int n_cta) {
    const int y = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n[y]) return;
    const float* src = a.gpart[y] + p;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 3 < n_cta; c += 4) {
        s0 += src[(size_t)c * a.n[y]];
        s1 += src[(size_t)(c + 1) * a.n[y]];
        s2 += src[(size_t)(c + 2) * a.n[y]];
        s3 += src[(size_t)(c + 3) * a.n[y]];
    }
    for (; c < n_cta; ++c) s0 += src[(size_t)c * a.n[y]];
    a.dparams[y][p] = (s0 + s1) + (s2 + s3);
}

// dx = dx0 + dx1 + dx2 in that order
__global__ void pe_sum_dx_kernel(float4* __restrict__ dx, const float4* __restrict__ d1, const float4* __restrict__ d2, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = dx[i];
        const float4 b = __ldcs(d1 + i), c = __ldcs(d2 + i);
        a.x = (a.x + b.x) + c.x; a.y = (a.y + b.y) + c.y; a.z = (a.z + b.z) + c.z; a.w = (a.w + b.w) + c.w;
        dx[i] = a;
    }
}

}  // namespace ldeq

using namespace ldeq;

struct ldeq_pe_tape {
    int B = 0, T = 0, F = 0, H = 0;
    bool has_lstm = false;
    void* base = nullptr;          // one stream-ordered allocation
    float *rnn = nullptr, *lf = nullptr, *lb = nullptr;
};

namespace {

template <int F> int pe_fwd_launch(ldeq_handle* h, const float* x, int B, int T, const float* rnn, const float* lf, const float* lb, float* z0o,
                                   float* tho, ldeq_pe_tape* tape, cudaStream_t s) {
    const int grid = (B + PE_SPB - 1) / PE_SPB, ny = lf ? 3 : 1;
    const size_t smem = lf ? PeSmem<4, F>::fwd_bytes : PeSmem<1, F>::fwd_bytes;
    // per device, a few microseconds: set on every call rather than cached per process
    cudaFuncSetAttribute(pe_fwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PeSmem<4, F>::fwd_bytes);
    PeStackArgs a{};
    a.params[0] = rnn; a.params[1] = lf; a.params[2] = lb;
    a.tape[0] = tape ? tape->rnn : nullptr; a.tape[1] = tape ? tape->lf : nullptr; a.tape[2] = tape ? tape->lb : nullptr;
    a.out[0] = z0o; a.out[1] = tho; a.out[2] = tho;
    a.ostride[0] = PE_H; a.ostride[1] = a.ostride[2] = 2 * PE_H;
    a.ooff[0] = 0; a.ooff[1] = 0; a.ooff[2] = PE_H;
    pe_fwd_kernel<F><<<dim3(grid, ny), PE_THREADS, smem, s>>>(x, B, T, a);
    h->launches += 1;
    LDEQ_CUDA(cudaGetLastError());
    return LDEQ_OK;
}

template <int F> int pe_bwd_launch(ldeq_handle* h, const ldeq_pe_tape* tape, const float* x, const float* rnn, const float* lf, const float* lb,
                                   const float* dz0o, const float* dtho, float* dx, float* drnn, float* dlf, float* dlb, cudaStream_t s) {
    const int B = tape->B, T = tape->T, grid = (B + PE_SPB - 1) / PE_SPB, ny = tape->has_lstm ? 3 : 1;
    const size_t smem = tape->has_lstm ? PeSmem<4, F>::bwd_bytes : PeSmem<1, F>::bwd_bytes;
    cudaFuncSetAttribute(pe_bwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PeSmem<4, F>::bwd_bytes);
    const size_t nr = pe_stack_params(1, F), nl = pe_stack_params(4, F), nx = (size_t)T * B * F;
    // scratch (stream-ordered): per-CTA partial gradients of every stack, and the two extra cotangent buffers
    const size_t n_part = (size_t)grid * (nr + (ny == 3 ? 2 * nl : 0)), n_dx = ny == 3 ? 2 * nx : 0;
    float* scratch = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&scratch, (n_part + n_dx) * sizeof(float), s);
    if (e != cudaSuccess) return set_err(h, LDEQ_ERR_NOMEM, "cudaMallocAsync(pattern extractor reverse-pass scratch)", e);
    PeStackArgs a{};
    a.params[0] = rnn; a.params[1] = lf; a.params[2] = lb;
    a.tape[0] = tape->rnn; a.tape[1] = tape->lf; a.tape[2] = tape->lb;
    a.out[0] = const_cast<float*>(dz0o); a.out[1] = a.out[2] = const_cast<float*>(dtho);
    a.ostride[0] = PE_H; a.ostride[1] = a.ostride[2] = 2 * PE_H;
    a.ooff[0] = 0; a.ooff[1] = 0; a.ooff[2] = PE_H;
    a.gpart[0] = scratch; a.gpart[1] = scratch + (size_t)grid * nr; a.gpart[2] = a.gpart[1] + (size_t)grid * nl;
    a.dx[0] = dx; a.dx[1] = scratch + n_part; a.dx[2] = a.dx[1] + nx;
    pe_bwd_kernel<F><<<dim3(grid, ny), PE_THREADS, smem, s>>>(x, B, T, a);
    PeReduceArgs r{};
    r.gpart[0] = a.gpart[0]; r.gpart[1] = a.gpart[1]; r.gpart[2] = a.gpart[2];
    r.dparams[0] = drnn; r.dparams[1] = dlf; r.dparams[2] = dlb;
    r.n[0] = (int)nr; r.n[1] = r.n[2] = (int)nl;
    pe_reduce_kernel<<<dim3((unsigned)((nl + 127) / 128), ny), 128, 0, s>>>(r, grid);
    h->launches += 2;
    if (ny == 3) {
        pe_sum_dx_kernel<<<h->sm_count * 8, 256, 0, s>>>((float4*)dx, (const float4*)a.dx[1], (const float4*)a.dx[2], nx / 4);
        h->launches += 1;
    }
    e = cudaGetLastError();
    cudaFreeAsync(scratch, s);
    if (e != cudaSuccess) return set_err(h, LDEQ_ERR_CUDA, "pattern extractor reverse pass launch", e);
    return LDEQ_OK;
}

bool pe_aligned(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

int ldeq_pattern_extractor_param_count(int cell, int F, int H) {
    if ((cell != 0 && cell != 1) || H != PE_H || F < 1) return LDEQ_ERR_INVALID;
    return pe_stack_params(cell == 1 ? 4 : 1, F);
}

int ldeq_pattern_extractor_fwd(ldeq_handle* h, const float* x, int B, int T, int F, int H, const float* rnn_params, const float* lstm_f_params,
                               const float* lstm_b_params, float* z0_out, float* theta_out, ldeq_pe_tape** tape_out, ldeq_stream stream) {
    if (!h) return LDEQ_ERR_INVALID;
    if (tape_out) *tape_out = nullptr;
    if (!x || !rnn_params || !z0_out) return set_err(h, LDEQ_ERR_INVALID, "null argument");
    if ((lstm_f_params == nullptr) != (lstm_b_params == nullptr) || (lstm_f_params && !theta_out))
        return set_err(h, LDEQ_ERR_INVALID, "lstm_f_params, lstm_b_params and theta_out come together (GOKU) or not at all (LatentODE)");
    if (B < 1 || T < 1) return set_err(h, LDEQ_ERR_INVALID, "B and T must be >= 1");
    if (H != PE_H || (F != 16 && F != 32 && F != 64))
        return set_err(h, LDEQ_ERR_UNSUPPORTED, "pattern extractor kernels: rnn_output_dim = 16 and rnn_input_dim in {16, 32, 64} (GOKU.jl:200-201 defaults: 32, 16)");
    if (!pe_aligned(x)) return set_err(h, LDEQ_ERR_INVALID, "x must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaSetDevice(h->device));
    ldeq_pe_tape* tape = nullptr;
    if (tape_out) {
        tape = new ldeq_pe_tape();
        tape->B = B; tape->T = T; tape->F = F; tape->H = H; tape->has_lstm = lstm_f_params != nullptr;
        const size_t nr = (size_t)T * B * PeDims<1>::NS, nl = tape->has_lstm ? (size_t)T * B * PeDims<4>::NS : 0;
        cudaError_t e = cudaMallocAsync(&tape->base, (nr + 2 * nl) * sizeof(float), s);
        if (e != cudaSuccess) { delete tape; return set_err(h, LDEQ_ERR_NOMEM, "cudaMallocAsync(pattern extractor tape)", e); }
        tape->rnn = (float*)tape->base;
        tape->lf = tape->rnn + nr;
        tape->lb = tape->lf + nl;
    }
    int rc = F == 16 ? pe_fwd_launch<16>(h, x, B, T, rnn_params, lstm_f_params, lstm_b_params, z0_out, theta_out, tape, s)
           : F == 32 ? pe_fwd_launch<32>(h, x, B, T, rnn_params, lstm_f_params, lstm_b_params, z0_out, theta_out, tape, s)
                     : pe_fwd_launch<64>(h, x, B, T, rnn_params, lstm_f_params, lstm_b_params, z0_out, theta_out, tape, s);
    if (rc) { if (tape) { cudaFreeAsync(tape->base, s); delete tape; } return rc; }
    if (tape_out) *tape_out = tape;
    return LDEQ_OK;
}

int ldeq_pattern_extractor_bwd(ldeq_handle* h, ldeq_pe_tape* tape, const float* x, const float* rnn_params, const float* lstm_f_params,
                               const float* lstm_b_params, const float* dz0_out, const float* dtheta_out, float* dx, float* d_rnn_params,
                               float* d_lstm_f_params, float* d_lstm_b_params, ldeq_stream stream) {
    if (!h) return LDEQ_ERR_INVALID;
    if (!tape || !x || !rnn_params || !dz0_out || !dx || !d_rnn_params) return set_err(h, LDEQ_ERR_INVALID, "null argument");
    if (tape->has_lstm && (!lstm_f_params || !lstm_b_params || !dtheta_out || !d_lstm_f_params || !d_lstm_b_params))
        return set_err(h, LDEQ_ERR_INVALID, "the tape was recorded with the LSTM stacks: their parameters, cotangent and gradient buffers are needed");
    if (!pe_aligned(x) || !pe_aligned(dx)) return set_err(h, LDEQ_ERR_INVALID, "x and dx must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaSetDevice(h->device));
    const int F = tape->F;
    return F == 16 ? pe_bwd_launch<16>(h, tape, x, rnn_params, lstm_f_params, lstm_b_params, dz0_out, dtheta_out, dx, d_rnn_params, d_lstm_f_params, d_lstm_b_params, s)
         : F == 32 ? pe_bwd_launch<32>(h, tape, x, rnn_params, lstm_f_params, lstm_b_params, dz0_out, dtheta_out, dx, d_rnn_params, d_lstm_f_params, d_lstm_b_params, s)
                   : pe_bwd_launch<64>(h, tape, x, rnn_params, lstm_f_params, lstm_b_params, dz0_out, dtheta_out, dx, d_rnn_params, d_lstm_f_params, d_lstm_b_params, s);
}

void ldeq_pe_tape_free(ldeq_handle* h, ldeq_pe_tape* tape, ldeq_stream stream) {
    if (!tape) return;
    if (h) cudaSetDevice(h->device);
    if (tape->base) cudaFreeAsync(tape->base, (cudaStream_t)stream);
    delete tape;
}

}  // extern "C"
