// kexp_krylov.cu -- shift-and-invert Arnoldi (krylov_exp.py:50-109) on the device.
#include "kexp_internal.h"

namespace kexp {
struct KrylovState {};
void krylov_free(kexp_ctx* ctx) { delete ctx->kry; ctx->kry = nullptr; }
}  // namespace kexp

extern "C" {
int kexp_set_krylov_operator(kexp_ctx* ctx, double, const int32_t*, const int32_t*, const double*,
                             const int32_t*, const int32_t*, const double*, const int32_t*, const int32_t*, const double*,
                             const int32_t*, const int32_t*, const double*, double, int) {
    return kexp::fail(ctx, "krylov path not built yet");
}
int kexp_krylov_apply(kexp_ctx* ctx, const double*, double*, double, int, int, double*, double*, double*) {
    return kexp::fail(ctx, "krylov path not built yet");
}
int kexp_dense_expm(kexp_ctx* ctx, const double*, int, double*) { return kexp::fail(ctx, "krylov path not built yet"); }
}
