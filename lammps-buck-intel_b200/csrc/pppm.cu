// placeholder translation unit replaced below
#include "internal.h"
struct PppmState { int dummy; };
void b2_pppm_free(b200md_ctx *ctx) { delete ctx->pppm; ctx->pppm = nullptr; }
int b2_pppm_compute(b200md_ctx *ctx, int, int, double *, double *) { return b2_fail(ctx, B200MD_EINVAL, "pppm not set up"); }
extern "C" {
int b200md_pppm_setup(b200md_ctx *ctx, const b200md_pppm_params *) { return b2_fail(ctx, B200MD_EINVAL, "todo"); }
int b200md_pppm_compute(b200md_ctx *ctx, int, int, double *, double *) { return b2_fail(ctx, B200MD_EINVAL, "todo"); }
int b200md_pppm_compute_host(b200md_ctx *ctx, int, int, int, const double *, const double *, double *, double *, double *) { return b2_fail(ctx, B200MD_EINVAL, "todo"); }
int b200md_pppm_download(b200md_ctx *ctx, double *, double *, double *, double *, double *, double *) { return b2_fail(ctx, B200MD_EINVAL, "todo"); }
int b200md_fft3d_host(b200md_ctx *ctx, double *, int, int, int, int) { return b2_fail(ctx, B200MD_EINVAL, "todo"); }
}
