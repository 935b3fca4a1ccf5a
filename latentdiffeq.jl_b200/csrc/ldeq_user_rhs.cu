// ldeq_user_rhs.cu -- user-defined right-hand sides compiled with NVRTC (placeholder).
#include "ldeq_internal.h"
using namespace ldeq;
extern "C" {
int ldeq_rhs_from_source(ldeq_handle* h, const char*, int, int, ldeq_rhs** out) {
    if (out) *out = nullptr;
    return set_err(h, LDEQ_ERR_UNSUPPORTED, "ldeq_rhs_from_source: not built yet");
}
void ldeq_rhs_free(ldeq_handle*, ldeq_rhs* rhs) { delete rhs; }
}
