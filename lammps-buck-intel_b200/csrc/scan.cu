// scan.cu — exclusive prefix sums used by cell binning, ghost compaction and the CSR neighbour
// offsets (int32 counts -> int32 / int64 offsets).  Three-level block scan: integer adds only, so the
// result is order-independent and bitwise reproducible.
#include "internal.h"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *total, T *smem /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < (blockDim.x >> 5) ? smem[lane] : (T)0;
    T wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  T res = smem[warp] + incl - v;
  __syncthreads();
  return res;
}

// pass 1: per-tile local exclusive scan, tile totals to sums[]
template <class TIN, class T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const TIN *__restrict__ in, T *__restrict__ out,
                                                            T *__restrict__ sums, size_t n) {
  __shared__ T smem[32];
  __shared__ T total;
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  T vals[SCAN_ITEMS];
  T tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    vals[k] = (base + k < n) ? (T)in[base + k] : (T)0;
    tsum += vals[k];
  }
  T off = block_exclusive_scan<T>(tsum, &total, smem);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = off;
    off += vals[k];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) add_offsets(T *__restrict__ out, const T *__restrict__ sums_scanned,
                                                             size_t n, T *__restrict__ total_slot) {
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  const T add = sums_scanned[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) out[base + k] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0 && total_slot) *total_slot = sums_scanned[gridDim.x];
}

template <class T>
__global__ void set_total_single(T *out, const T *sums, size_t n) {
  out[n] = sums[0];
}

template <class TIN, class T>
int scan_rec(b200md_ctx *ctx, const TIN *in, T *out, size_t n, T *ws, T *total_slot) {
  // ws layout: [sums(nt+1)] [scanned(nt+1)] [recursive ws ...]
  const size_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
  T *sums = ws;
  T *scanned = ws + (nt + 1);
  T *next = scanned + (nt + 1);
  scan_tiles<TIN, T><<<(unsigned)nt, SCAN_THREADS, 0, ctx->stream>>>(in, out, sums, n);
  KERNEL_OK(ctx, "scan_tiles");
  if (nt == 1) {
    if (total_slot) {
      set_total_single<T><<<1, 1, 0, ctx->stream>>>(total_slot, sums, 0);
      KERNEL_OK(ctx, "set_total_single");
    }
    return 0;
  }
  TRY((scan_rec<T, T>(ctx, sums, scanned, nt, next, scanned + nt)));
  add_offsets<T><<<(unsigned)nt, SCAN_THREADS, 0, ctx->stream>>>(out, scanned, n, total_slot);
  KERNEL_OK(ctx, "add_offsets");
  return 0;
}

}  // namespace

size_t b2_scan_ws_bytes(size_t n) {
  size_t bytes = 0;
  while (true) {
    const size_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    bytes += 2 * (nt + 1) * sizeof(long long);
    if (nt <= 1) break;
    n = nt;
  }
  return bytes + 64;
}

int b2_exclusive_scan_i32(b200md_ctx *ctx, const int *in, int *out, size_t n, void *ws) {
  if (n == 0) {
    CUDA_OK(ctx, cudaMemsetAsync(out, 0, sizeof(int), ctx->stream));
    return 0;
  }
  return scan_rec<int, int>(ctx, in, out, n, (int *)ws, out + n);
}

int b2_exclusive_scan_i32_i64(b200md_ctx *ctx, const int *in, long long *out, size_t n, void *ws) {
  if (n == 0) {
    CUDA_OK(ctx, cudaMemsetAsync(out, 0, sizeof(long long), ctx->stream));
    return 0;
  }
  return scan_rec<int, long long>(ctx, in, out, n, (long long *)ws, out + n);
}
