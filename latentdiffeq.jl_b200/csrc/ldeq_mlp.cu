// ldeq_mlp.cu -- LatentODE path (placeholder until the MLP integrator lands in this round).
#include "ldeq_internal.h"
using namespace ldeq;
extern "C" {
int ldeq_mlp_solve_fwd(ldeq_handle* h, int, const void*, const void*, const int32_t*, int, const double*, int, int,
                       const ldeq_opts*, void*, int32_t*, int32_t*, int32_t*, ldeq_mlp_tape**, ldeq_stream) {
    return set_err(h, LDEQ_ERR_UNSUPPORTED, "ldeq_mlp_solve_fwd: not built yet");
}
int ldeq_mlp_solve_bwd(ldeq_handle* h, ldeq_mlp_tape*, const void*, void*, void*, ldeq_stream) {
    return set_err(h, LDEQ_ERR_UNSUPPORTED, "ldeq_mlp_solve_bwd: not built yet");
}
void ldeq_mlp_tape_free(ldeq_handle*, ldeq_mlp_tape*, ldeq_stream) {}
}
