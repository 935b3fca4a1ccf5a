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

struct ldeq_pe_tape {
    int B = 0, T = 0, F = 0, H = 0;
    bool has_lstm = false;
    void* base = nullptr;          // one stream-ordered allocation
    float *rnn = nullptr, *lf = nullptr, *lb = nullptr;
};

#define PE_NAMESPACE pe16
#define PE_HIDDEN 16
#include "ldeq_recurrent_kernels.inc"
#undef PE_NAMESPACE
#undef PE_HIDDEN
// LatentODE's default pattern extractor is Chain(RNN(32,32,relu), RNN(32,32,relu)) (LatentODE.jl:100-124): 8 lanes per
// sequence; the LSTM bodies are not instantiated at this size (their reverse pass keeps all 4 H pre-activation cotangents
// of a sequence in registers)
#define PE_NAMESPACE pe32
#define PE_HIDDEN 32
#define PE_RNN_ONLY 1
#include "ldeq_recurrent_kernels.inc"
#undef PE_NAMESPACE
#undef PE_HIDDEN
#undef PE_RNN_ONLY

using namespace ldeq;

namespace {
bool pe_aligned(const void* p) { return ((uintptr_t)p & 15) == 0; }
}  // namespace

extern "C" {

int ldeq_pattern_extractor_param_count(int cell, int F, int H) {
    if ((cell != 0 && cell != 1) || (H != 16 && H != 32) || F < 1) return LDEQ_ERR_INVALID;
    return H == 16 ? pe16::pe_stack_params(cell == 1 ? 4 : 1, F) : pe32::pe_stack_params(cell == 1 ? 4 : 1, F);
}

int ldeq_pattern_extractor_fwd(ldeq_handle* h, const float* x, int B, int T, int F, int H, const float* rnn_params, const float* lstm_f_params,
                               const float* lstm_b_params, float* z0_out, float* theta_out, ldeq_pe_tape** tape_out, ldeq_stream stream) {
    if (!h) return LDEQ_ERR_INVALID;
    if (tape_out) *tape_out = nullptr;
    if (!x || !rnn_params || !z0_out) return set_err(h, LDEQ_ERR_INVALID, "null argument");
    if ((lstm_f_params == nullptr) != (lstm_b_params == nullptr) || (lstm_f_params && !theta_out))
        return set_err(h, LDEQ_ERR_INVALID, "lstm_f_params, lstm_b_params and theta_out come together (GOKU) or not at all (LatentODE)");
    if (B < 1 || T < 1) return set_err(h, LDEQ_ERR_INVALID, "B and T must be >= 1");
    const bool ok16 = H == 16 && (F == 16 || F == 32 || F == 64);
    const bool ok32 = H == 32 && (F == 32 || F == 64) && !lstm_f_params;
    if (!ok16 && !ok32)
        return set_err(h, LDEQ_ERR_UNSUPPORTED, "pattern extractor kernels: rnn_output_dim = 16 with rnn_input_dim in {16, 32, 64} (GOKU.jl:200-201 "
                                                "defaults: 32, 16), or the relu-RNN stack alone with rnn_output_dim = 32 and rnn_input_dim in {32, 64} "
                                                "(LatentODE.jl:101-102 defaults: 32, 32)");
    if (!pe_aligned(x)) return set_err(h, LDEQ_ERR_INVALID, "x must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    LDEQ_CUDA(cudaSetDevice(h->device));
    ldeq_pe_tape* tape = nullptr;
    if (tape_out) {
        tape = new ldeq_pe_tape();
        tape->B = B; tape->T = T; tape->F = F; tape->H = H; tape->has_lstm = lstm_f_params != nullptr;
        const size_t nr = (size_t)T * B * 2 * H, nl = tape->has_lstm ? (size_t)T * B * 4 * H : 0;   // h1 [c1] h2 [c2] per step
        cudaError_t e = cudaMallocAsync(&tape->base, (nr + 2 * nl) * sizeof(float), s);
        if (e != cudaSuccess) { delete tape; return set_err(h, LDEQ_ERR_NOMEM, "cudaMallocAsync(pattern extractor tape)", e); }
        tape->rnn = (float*)tape->base;
        tape->lf = tape->rnn + nr;
        tape->lb = tape->lf + nl;
    }
    int rc;
#define LDEQ_PE_FWD(NS, FV) NS::pe_fwd_launch<FV>(h, x, B, T, rnn_params, lstm_f_params, lstm_b_params, z0_out, theta_out, tape, s)
    if (H == 16) rc = F == 16 ? LDEQ_PE_FWD(pe16, 16) : F == 32 ? LDEQ_PE_FWD(pe16, 32) : LDEQ_PE_FWD(pe16, 64);
    else rc = F == 32 ? LDEQ_PE_FWD(pe32, 32) : LDEQ_PE_FWD(pe32, 64);
#undef LDEQ_PE_FWD
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
#define LDEQ_PE_BWD(NS, FV) NS::pe_bwd_launch<FV>(h, tape, x, rnn_params, lstm_f_params, lstm_b_params, dz0_out, dtheta_out, dx, d_rnn_params, d_lstm_f_params, d_lstm_b_params, s)
    if (tape->H == 16) return F == 16 ? LDEQ_PE_BWD(pe16, 16) : F == 32 ? LDEQ_PE_BWD(pe16, 32) : LDEQ_PE_BWD(pe16, 64);
    return F == 32 ? LDEQ_PE_BWD(pe32, 32) : LDEQ_PE_BWD(pe32, 64);
#undef LDEQ_PE_BWD
}

void ldeq_pe_tape_free(ldeq_handle* h, ldeq_pe_tape* tape, ldeq_stream stream) {
    if (!tape) return;
    if (h) cudaSetDevice(h->device);
    if (tape->base) cudaFreeAsync(tape->base, (cudaStream_t)stream);
    delete tape;
}

}  // extern "C"
