// comm.cu — NCCL plumbing of the multi-GPU path: one process per GPU, z-slab decomposition.
//
// Replaces what the reference reaches through upstream objects: Comm::exchange/borders/forward_comm (atom
// migration and ghost-atom halo; consumed at pair_buck_intel.cpp:86,290), GridComm (pppm_intel.cpp:185,219-220),
// Remap + FFT3d transposes (:664,:835,:903,930,958) and the two MPI_Allreduce calls (:260,:273).
// Only primitives live here: neighbour exchange (grouped ncclSend/ncclRecv over NVLink), all-to-all (grouped
// send/recv), small all-reduces.  The decomposition logic is in neigh.cu (atoms) and pppm.cu (grid).
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"

struct CommState {
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  int lower = 0, upper = 0;
  DevBuf<int> cnt;  // [4] device counters for count exchanges
};

#define NCCL_OK(ctx, call)                                                                                      \
  do {                                                                                                          \
    ncclResult_t r__ = (call);                                                                                  \
    if (r__ != ncclSuccess)                                                                                     \
      return b2_fail(ctx, B200MD_ECOMM, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

void b2_comm_free(b200md_ctx *ctx) {
  if (!ctx->comm) return;
  if (ctx->comm->comm) ncclCommDestroy(ctx->comm->comm);
  ctx->comm->cnt.free_();
  delete ctx->comm;
  ctx->comm = nullptr;
}

int b2_comm_nranks(const b200md_ctx *ctx) { return ctx->comm ? ctx->comm->nranks : 1; }
int b2_comm_rank(const b200md_ctx *ctx) { return ctx->comm ? ctx->comm->rank : 0; }

// exchange with the two z neighbours: send s_lo to the lower rank and s_hi to the upper rank, receive what the
// upper rank sent down (r_from_hi) and what the lower rank sent up (r_from_lo).  Sizes in bytes.  With two ranks
// both neighbours are the same peer; NCCL matches the two messages in issue order, which this ordering respects.
int b2_comm_exchange(b200md_ctx *ctx, const void *s_lo, size_t n_lo, const void *s_hi, size_t n_hi, void *r_from_hi,
                     size_t n_from_hi, void *r_from_lo, size_t n_from_lo) {
  CommState *cs = ctx->comm;
  NCCL_OK(ctx, ncclGroupStart());
  if (n_lo) NCCL_OK(ctx, ncclSend(s_lo, n_lo, ncclChar, cs->lower, cs->comm, ctx->stream));
  if (n_hi) NCCL_OK(ctx, ncclSend(s_hi, n_hi, ncclChar, cs->upper, cs->comm, ctx->stream));
  if (n_from_hi) NCCL_OK(ctx, ncclRecv(r_from_hi, n_from_hi, ncclChar, cs->upper, cs->comm, ctx->stream));
  if (n_from_lo) NCCL_OK(ctx, ncclRecv(r_from_lo, n_from_lo, ncclChar, cs->lower, cs->comm, ctx->stream));
  NCCL_OK(ctx, ncclGroupEnd());
  return 0;
}

// counts first: every rank always sends/receives one int per direction, so the pattern is static
int b2_comm_exchange_counts(b200md_ctx *ctx, int n_lo, int n_hi, int *n_from_hi, int *n_from_lo) {
  CommState *cs = ctx->comm;
  RESERVE(ctx, cs->cnt, 8);
  int h[2] = {n_lo, n_hi};
  CUDA_OK(ctx, cudaMemcpyAsync(cs->cnt.p, h, 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  TRY(b2_comm_exchange(ctx, cs->cnt.p, sizeof(int), cs->cnt.p + 1, sizeof(int), cs->cnt.p + 2, sizeof(int),
                       cs->cnt.p + 3, sizeof(int)));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, cs->cnt.p + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  *n_from_hi = ((int *)ctx->h_pinned)[0];
  *n_from_lo = ((int *)ctx->h_pinned)[1];
  return 0;
}

int b2_comm_allreduce_sum(b200md_ctx *ctx, double *dev, int n) {
  if (b2_comm_nranks(ctx) == 1) return 0;
  NCCL_OK(ctx, ncclAllReduce(dev, dev, n, ncclDouble, ncclSum, ctx->comm->comm, ctx->stream));
  return 0;
}

int b2_comm_allreduce_max_int(b200md_ctx *ctx, int *dev, int n) {
  if (b2_comm_nranks(ctx) == 1) return 0;
  NCCL_OK(ctx, ncclAllReduce(dev, dev, n, ncclInt, ncclMax, ctx->comm->comm, ctx->stream));
  return 0;
}

int b2_comm_allgather_int(b200md_ctx *ctx, int value, int *host_out) {
  CommState *cs = ctx->comm;
  if (!cs) { host_out[0] = value; return 0; }
  DevBuf<int> buf;
  if (buf.reserve((size_t)cs->nranks + 1)) return b2_fail(ctx, B200MD_ENOMEM, "allgather: out of memory");
  CUDA_OK(ctx, cudaMemcpyAsync(buf.p + cs->nranks, &value, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_OK(ctx, ncclAllGather(buf.p + cs->nranks, buf.p, 1, ncclInt, cs->comm, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(host_out, buf.p, cs->nranks * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  buf.free_();
  return 0;
}

// every rank contributes `bytes[rank]` bytes at displacement disp[rank] of recv (device memory, may be the buffer the
// contribution already sits in): one broadcast per contributing rank inside a group
int b2_comm_allgatherv(b200md_ctx *ctx, const void *send, void *recv, const size_t *bytes, const size_t *disp) {
  CommState *cs = ctx->comm;
  if (!cs) return 0;
  NCCL_OK(ctx, ncclGroupStart());
  for (int r = 0; r < cs->nranks; r++)
    if (bytes[r])
      NCCL_OK(ctx, ncclBroadcast(r == cs->rank ? send : (const char *)recv + disp[r], (char *)recv + disp[r], bytes[r],
                                 ncclChar, r, cs->comm, ctx->stream));
  NCCL_OK(ctx, ncclGroupEnd());
  return 0;
}

// all-to-all with per-peer element counts (bytes) and displacements: the FFT transposes
int b2_comm_alltoallv(b200md_ctx *ctx, const void *sbuf, const size_t *scount, const size_t *sdisp, void *rbuf,
                      const size_t *rcount, const size_t *rdisp) {
  CommState *cs = ctx->comm;
  NCCL_OK(ctx, ncclGroupStart());
  for (int p = 0; p < cs->nranks; p++) {
    if (scount[p]) NCCL_OK(ctx, ncclSend((const char *)sbuf + sdisp[p], scount[p], ncclChar, p, cs->comm, ctx->stream));
    if (rcount[p]) NCCL_OK(ctx, ncclRecv((char *)rbuf + rdisp[p], rcount[p], ncclChar, p, cs->comm, ctx->stream));
  }
  NCCL_OK(ctx, ncclGroupEnd());
  return 0;
}

int b2_comm_barrier(b200md_ctx *ctx) {
  CommState *cs = ctx->comm;
  if (!cs) return 0;
  RESERVE(ctx, cs->cnt, 8);
  // an all-reduce cannot complete on any rank before every rank has launched it, i.e. has finished the stream work
  // ahead of it: stores into peer memory issued before the barrier are complete when kernels after it run
  NCCL_OK(ctx, ncclAllReduce(cs->cnt.p + 4, cs->cnt.p + 5, 1, ncclInt, ncclSum, cs->comm, ctx->stream));
  return 0;
}

int b2_comm_peer_alloc(b200md_ctx *ctx, PeerBuf &pb, size_t bytes, int *ok) {
  CommState *cs = ctx->comm;
  *ok = 0;
  if (!cs || cs->nranks < 2) return 0;
  const int P = cs->nranks, me = cs->rank;
  b2_comm_peer_free(ctx, pb);
  int good = 1;
  void *loc = nullptr;
  cudaIpcMemHandle_t h;
  std::memset(&h, 0, sizeof(h));
  if (cudaMalloc(&loc, bytes) != cudaSuccess) { good = 0; loc = nullptr; cudaGetLastError(); }
  if (good && cudaIpcGetMemHandle(&h, loc) != cudaSuccess) { good = 0; cudaGetLastError(); }
  // all-gather the handles (64 B each) and every rank's status
  DevBuf<unsigned char> hb;
  const size_t rec = sizeof(cudaIpcMemHandle_t) + sizeof(int);
  if (hb.reserve(rec * (size_t)(P + 1))) { if (loc) cudaFree(loc); return b2_fail(ctx, B200MD_ENOMEM, "peer alloc: out of memory"); }
  std::vector<unsigned char> host(rec * (size_t)P);
  std::memcpy(host.data(), &h, sizeof(h));
  std::memcpy(host.data() + sizeof(h), &good, sizeof(int));
  CUDA_OK(ctx, cudaMemcpyAsync(hb.p + rec * P, host.data(), rec, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_OK(ctx, ncclAllGather(hb.p + rec * P, hb.p, rec, ncclChar, cs->comm, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(host.data(), hb.p, rec * P, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < P; q++) {
    int g;
    std::memcpy(&g, host.data() + rec * q + sizeof(h), sizeof(int));
    if (!g) good = 0;
  }
  void *peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (good) {
    for (int q = 0; q < P; q++) {
      if (q == me) { peer[q] = loc; continue; }
      cudaIpcMemHandle_t hq;
      std::memcpy(&hq, host.data() + rec * q, sizeof(hq));
      if (cudaIpcOpenMemHandle(&peer[q], hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        good = 0;
        peer[q] = nullptr;
        cudaGetLastError();
      }
    }
  }
  // the decision is collective: one rank without peer access sends everybody to the NCCL path
  int *flag = (int *)hb.p;
  CUDA_OK(ctx, cudaMemcpyAsync(flag, &good, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_OK(ctx, ncclAllReduce(flag, flag, 1, ncclInt, ncclMin, cs->comm, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(&good, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  hb.free_();
  if (!good) {
    for (int q = 0; q < P; q++)
      if (q != me && peer[q]) cudaIpcCloseMemHandle(peer[q]);
    if (loc) cudaFree(loc);
    return 0;
  }
  pb.local = loc;
  pb.bytes = bytes;
  for (int q = 0; q < P; q++) pb.peer[q] = peer[q];
  *ok = 1;
  return 0;
}

void b2_comm_peer_free(b200md_ctx *ctx, PeerBuf &pb) {
  if (!pb.local) return;
  const int me = b2_comm_rank(ctx);
  cudaStreamSynchronize(ctx->stream);
  for (int q = 0; q < 8; q++) {
    if (pb.peer[q] && q != me && pb.peer[q] != pb.local) cudaIpcCloseMemHandle(pb.peer[q]);
    pb.peer[q] = nullptr;
  }
  cudaFree(pb.local);
  pb.local = nullptr;
  pb.bytes = 0;
}

CommGroup::CommGroup(b200md_ctx *c) : ctx(c) {
  if (ctx->comm && ncclGroupStart() == ncclSuccess) open = true;
}
CommGroup::~CommGroup() {
  if (open) ncclGroupEnd();
}
int CommGroup::send(const void *buf, size_t bytes, int peer) {
  if (!open) return b2_fail(ctx, B200MD_ECOMM, "NCCL group is not open");
  if (bytes) NCCL_OK(ctx, ncclSend(buf, bytes, ncclChar, peer, ctx->comm->comm, ctx->stream));
  return 0;
}
int CommGroup::recv(void *buf, size_t bytes, int peer) {
  if (!open) return b2_fail(ctx, B200MD_ECOMM, "NCCL group is not open");
  if (bytes) NCCL_OK(ctx, ncclRecv(buf, bytes, ncclChar, peer, ctx->comm->comm, ctx->stream));
  return 0;
}
int CommGroup::end() {
  if (!open) return 0;
  open = false;
  NCCL_OK(ctx, ncclGroupEnd());
  return 0;
}

extern "C" {

int b200md_comm_unique_id(void *id128) {
  if (!id128) return B200MD_EINVAL;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return B200MD_ECOMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

int b200md_comm_init(b200md_ctx *ctx, int rank, int nranks, const void *id128) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return b2_fail(ctx, B200MD_EINVAL, "b200md_comm_init: bad rank");
  if (nranks > 8) return b2_fail(ctx, B200MD_EINVAL, "b200md_comm_init: at most 8 ranks (one NVSwitch node)");
  cudaSetDevice(ctx->device);
  b2_comm_free(ctx);
  // a k-space state keeps the rank count and the slab plan it was set up with: drop it, b200md_pppm_setup must follow
  b2_pppm_free(ctx);
  ctx->neigh.ready = false;
  if (nranks == 1) return 0;
  if (!id128) return b2_fail(ctx, B200MD_EINVAL, "b200md_comm_init: missing ncclUniqueId");
  CommState *cs = new CommState();
  ctx->comm = cs;
  cs->rank = rank;
  cs->nranks = nranks;
  cs->lower = (rank + nranks - 1) % nranks;
  cs->upper = (rank + 1) % nranks;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  NCCL_OK(ctx, ncclCommInitRank(&cs->comm, nranks, id, rank));
  ctx->neigh.ready = false;
  // several GPUs: the k-space solve runs on its own stream underneath the pair kernel by default — part of its time
  // is NCCL wait in the FFT transposes, which the pair kernel fills (4 GPUs: 20.3 -> 19.5 ms per step,
  // profiles/r01_overlap.txt).  B200MD_OVERLAP=0 turns it off.
  const char *ov = getenv("B200MD_OVERLAP");
  ctx->overlap = !(ov && ov[0] == '0');
  return 0;
}

int b200md_comm_finalize(b200md_ctx *ctx) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  b2_comm_free(ctx);
  return 0;
}

}  // extern "C"
