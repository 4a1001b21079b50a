// comm.cu — multi-GPU plumbing (filled in with the slab decomposition).
#include "internal.h"
struct CommState { int rank = 0, nranks = 1; };
void b2_comm_free(b200md_ctx *ctx) { delete ctx->comm; ctx->comm = nullptr; }
extern "C" {
int b200md_comm_unique_id(void *id128) { (void)id128; return B200MD_ECOMM; }
int b200md_comm_init(b200md_ctx *ctx, int rank, int nranks, const void *id128) {
  (void)rank; (void)id128;
  if (nranks == 1) return 0;
  return b2_fail(ctx, B200MD_ECOMM, "multi-GPU decomposition not built into this library yet");
}
int b200md_comm_finalize(b200md_ctx *ctx) { (void)ctx; return 0; }
}
