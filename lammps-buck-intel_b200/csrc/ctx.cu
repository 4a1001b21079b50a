// ctx.cu — context, units, box, atom upload/download (replaces FixIntel + IntelBuffers ownership,
// intel_buffers.h:272-312, and IntelBuffers::thr_pack, intel_buffers.h:185-203).
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "internal.h"

static std::string g_create_error;
static void harvest_timers(b200md_ctx *ctx) {
  if (ctx->t_pending.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto &r : ctx->t_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ctx->t_ms[r.id] += ms;
      ctx->t_calls[r.id]++;
    }
    ctx->t_pool.push_back(r.a);
    ctx->t_pool.push_back(r.b);
  }
  ctx->t_pending.clear();
}


int b2_fail(b200md_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_create_error = buf;
  return code;
}

static const char *kTimerNames[T_COUNT] = {"neigh",       "comm",        "pair",         "make_rho",   "fft",
                                           "poisson",     "fieldforce",  "nve",          "other",      "k_nb_mask",
                                           "k_nb_fill",   "k_rho_tiles", "k_rho_fold",   "k_fft_x_fwd", "k_fft_y_fwd",
                                           "k_fft_z_poisson", "k_fft_y_inv", "k_fft_x_inv", "k_nve_initial",
                                           "k_nve_final"};

namespace {

__global__ void pack_atoms(int n, const double *__restrict__ x, const double *__restrict__ v,
                           const double *__restrict__ q, const int *__restrict__ type,
                           double4 *__restrict__ xq, double4 *__restrict__ vv, int *__restrict__ tag,
                           int *__restrict__ type_out, int first_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xq[i] = make_double4(x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2], q ? q[i] : 0.0);
  vv[i] = v ? make_double4(v[3 * (size_t)i], v[3 * (size_t)i + 1], v[3 * (size_t)i + 2], 0.0)
            : make_double4(0.0, 0.0, 0.0, 0.0);
  tag[i] = first_id + i;
  type_out[i] = type[i];
}

__global__ void set_x_by_tag(int n, const double *__restrict__ x, const int *__restrict__ tag,
                             double4 *__restrict__ xq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tag[i];
  double4 p = xq[i];
  p.x = x[3 * (size_t)t];
  p.y = x[3 * (size_t)t + 1];
  p.z = x[3 * (size_t)t + 2];
  xq[i] = p;
}

// sorted -> host order; out3 is [n][3], outw optional scalar (w component)
__global__ void unpack_by_tag(int n, const double4 *__restrict__ src, const int *__restrict__ tag,
                              double *__restrict__ out3, double *__restrict__ outw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tag[i];
  const double4 s = src[i];
  if (out3) {
    out3[3 * (size_t)t] = s.x;
    out3[3 * (size_t)t + 1] = s.y;
    out3[3 * (size_t)t + 2] = s.z;
  }
  if (outw) outw[t] = s.w;
}

__global__ void to_float_copy(int first, int count, const double4 *__restrict__ xq, float4 *__restrict__ xqf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const double4 p = xq[first + i];
  xqf[first + i] = make_float4((float)p.x, (float)p.y, (float)p.z, (float)p.w);
}

// measurement helpers (bench.py roofline denominators): register-resident FMA chains, and a plain copy
template <class T>
__global__ void __launch_bounds__(256) k_fma_peak(int iters, T seed, T *__restrict__ out) {
  T a0 = seed + threadIdx.x, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3, a4 = a0 + (T)4, a5 = a0 + (T)5,
    a6 = a0 + (T)6, a7 = a0 + (T)7;
  const T m = (T)0.999999, c = (T)1e-6;
  for (int i = 0; i < iters; i++) {
    a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
    a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
  }
  const T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == (T)-12345.678) out[0] = s;
}
__global__ void k_copy16(size_t n, const double2 *__restrict__ in, double2 *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace

int b2_refresh_float_copy(b200md_ctx *ctx, int first, int count) {
  if (ctx->prec != B200MD_PREC_MIXED || count <= 0) return 0;
  to_float_copy<<<cdiv(count, 256), 256, 0, ctx->stream>>>(first, count, ctx->xq.p, ctx->xqf.p);
  KERNEL_OK(ctx, "to_float_copy");
  return 0;
}

// sorted -> host order into a staging area of 3*nlocal doubles (on the main stream)
int b2_unpack_to_stage(b200md_ctx *ctx, const double4 *src, double *stage3) {
  if (ctx->nlocal == 0) return 0;
  unpack_by_tag<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(ctx->nlocal, src, ctx->tag.p, stage3, nullptr);
  KERNEL_OK(ctx, "unpack_by_tag");
  return 0;
}

extern "C" {

int b200md_version(void) { return 100; }

const char *b200md_last_error(const b200md_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int b200md_ctx_create(int device, int precision, b200md_ctx **out) {
  if (!out) return b2_fail(nullptr, B200MD_EINVAL, "b200md_ctx_create: out is NULL");
  *out = nullptr;
  if (precision != B200MD_PREC_DOUBLE && precision != B200MD_PREC_MIXED)
    return b2_fail(nullptr, B200MD_EINVAL,
                   "precision mode %d not provided on the device (double and mixed only)", precision);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return b2_fail(nullptr, B200MD_ENODEV, "no CUDA device: %s (there is no CPU fallback)",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev)
    return b2_fail(nullptr, B200MD_ENODEV, "device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return b2_fail(nullptr, B200MD_ENODEV, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return b2_fail(nullptr, B200MD_ENODEV,
                   "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                   prop.minor);
  if (cudaSetDevice(device) != cudaSuccess) return b2_fail(nullptr, B200MD_ECUDA, "cudaSetDevice failed");
  b200md_ctx *ctx = new b200md_ctx();
  ctx->device = device;
  ctx->prec = precision;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return b2_fail(nullptr, B200MD_ECUDA, "cudaStreamCreate failed");
  }
  cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = greatest priority (numerically lowest)
    cudaStreamCreateWithPriority(&ctx->kstream, cudaStreamNonBlocking, hi);
    cudaEventCreateWithFlags(&ctx->ev_pre, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_k, cudaEventDisableTiming);
    ctx->main_stream = ctx->stream;
    // measured on one B200 (profiles/r01_overlap.txt): hiding PPPM under the pair kernel slows that kernel by more
    // than PPPM costs (10.96 -> 17.4 ms: they contend for L1/shared memory), so on one GPU the overlap is opt-in
    // (B200MD_OVERLAP=1); b200md_comm_init turns it on for several GPUs
    {
      const char *ov = getenv("B200MD_OVERLAP");
      ctx->overlap = ov && ov[0] != '0';
    }
  }
  cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
  cudaEventCreate(&ctx->ev_a);
  cudaEventCreate(&ctx->ev_b);
  ctx->h_pinned_bytes = 1 << 16;
  if (cudaMallocHost((void **)&ctx->h_pinned, ctx->h_pinned_bytes) != cudaSuccess) {
    delete ctx;
    return b2_fail(nullptr, B200MD_ENOMEM, "cudaMallocHost failed");
  }
  *out = ctx;
  return 0;
}

void b200md_ctx_destroy(b200md_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  b2_pppm_free(ctx);
  b2_comm_free(ctx);
  ctx->xq.free_(); ctx->v.free_(); ctx->f.free_(); ctx->xqf.free_();
  ctx->nve_group.free_(); ctx->nve_rmass.free_();
  ctx->type.free_(); ctx->tag.free_(); ctx->inv_tag.free_(); ctx->stage.free_();
  ctx->ev_partial.free_(); ctx->ev_out.free_();
  PairState &ps = ctx->pair;
  ps.coeff_d.free_(); ps.coeff_f.free_(); ps.cutneighsq.free_();
  ps.ctab_d.free_(); ps.ctab_f.free_(); ps.dtab_d.free_(); ps.dtab_f.free_(); ps.exptab.free_();
  NeighState &ns = ctx->neigh;
  ns.bin_of.free_(); ns.bin_sorted.free_(); ns.goff.free_(); ns.gsrc_tmp.free_(); ns.gshift_tmp.free_();
  ns.gbin.free_(); ns.gperm.free_(); ns.bin_count.free_(); ns.bin_start.free_(); ns.bin_end.free_(); ns.bin_cursor.free_();
  ns.perm.free_(); ns.ghost_src.free_(); ns.ghost_shift.free_(); ns.ghost_cnt.free_();
  ns.numneigh.free_(); ns.offsets.free_(); ns.entries.free_(); ns.mask_words.free_(); ns.mask_off.free_(); ns.maskbuf.free_();
  ns.halo_idx.free_(); ns.halo_sbuf.free_(); ns.halo_rbuf.free_(); ns.halo_stype.free_(); ns.halo_rtype.free_(); ns.halo_stag.free_(); ns.halo_rtag.free_();
  ns.mig_flag.free_(); ns.mig_off.free_(); ns.mig_send.free_(); ns.mig_recv.free_(); ns.xhold.free_(); ns.flags.free_();
  ns.scan_ws.free_(); ns.tmp4a.free_(); ns.tmp4b.free_(); ns.tmpi_a.free_(); ns.tmpi_b.free_();
  harvest_timers(ctx);
  for (cudaEvent_t e : ctx->t_pool) cudaEventDestroy(e);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_pre) cudaEventDestroy(ctx->ev_pre);
  if (ctx->ev_k) cudaEventDestroy(ctx->ev_k);
  if (ctx->kstream) cudaStreamDestroy(ctx->kstream);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int b200md_set_units(b200md_ctx *ctx, double qqrd2e, double ftm2v) {
  if (!ctx) return B200MD_EINVAL;
  ctx->qqrd2e = qqrd2e;
  ctx->ftm2v = ftm2v;
  return 0;
}

int b200md_set_box(b200md_ctx *ctx, const double boxlo[3], const double boxhi[3], const int periodic[3]) {
  if (!ctx) return B200MD_EINVAL;
  for (int d = 0; d < 3; d++) {
    if (!std::isfinite(boxlo[d]) || !std::isfinite(boxhi[d]))
      return b2_fail(ctx, B200MD_ENONFINITE, "Non-numeric box dimensions - simulation unstable");
    if (!(boxhi[d] > boxlo[d])) return b2_fail(ctx, B200MD_EINVAL, "box dimension %d is empty", d);
    ctx->boxlo[d] = boxlo[d];
    ctx->boxhi[d] = boxhi[d];
    ctx->prd[d] = boxhi[d] - boxlo[d];
    ctx->periodic[d] = periodic ? periodic[d] : 1;
  }
  ctx->box_set = true;
  ctx->neigh.ready = false;
  ctx->triclinic = false;
  ctx->tilt[0] = ctx->tilt[1] = ctx->tilt[2] = 0.0;
  return 0;
}

int b200md_set_box_triclinic(b200md_ctx *ctx, const double boxlo[3], const double boxhi[3], double xy, double xz,
                             double yz) {
  if (!ctx) return B200MD_EINVAL;
  if (!std::isfinite(xy) || !std::isfinite(xz) || !std::isfinite(yz))
    return b2_fail(ctx, B200MD_ENONFINITE, "Non-numeric box dimensions - simulation unstable");
  const int periodic[3] = {1, 1, 1};
  TRY(b200md_set_box(ctx, boxlo, boxhi, periodic));
  // Domain::set_initial_box [UPSTREAM]: a tilt beyond half a box length is refused
  if (std::fabs(xy / ctx->prd[0]) > 0.5 || std::fabs(xz / ctx->prd[0]) > 0.5 || std::fabs(yz / ctx->prd[1]) > 0.5)
    return b2_fail(ctx, B200MD_EINVAL, "Triclinic box skew is too large");
  ctx->triclinic = xy != 0.0 || xz != 0.0 || yz != 0.0;
  ctx->tilt[0] = xy; ctx->tilt[1] = xz; ctx->tilt[2] = yz;
  return 0;
}

int b200md_atoms_upload(b200md_ctx *ctx, int nlocal, int ntypes, const double *x, const double *v,
                        const double *q, const int *type, const double *mass) {
  if (!ctx || nlocal < 0 || !x || !type || !mass || ntypes < 1)
    return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_upload: bad arguments");
  if (ntypes > B2_MAXTYPES)
    return b2_fail(ctx, B200MD_EINVAL, "ntypes %d > %d supported on the device", ntypes, B2_MAXTYPES);
  for (int i = 0; i < nlocal; i++)
    if (type[i] < 1 || type[i] > ntypes)
      return b2_fail(ctx, B200MD_EINVAL, "atom %d has type %d outside 1..%d", i, type[i], ntypes);
  cudaSetDevice(ctx->device);
  ctx->ev_pre_valid = false;
  ctx->nlocal = nlocal;
  ctx->nghost = 0;
  ctx->ntypes = ntypes;
  ctx->has_q = q != nullptr;
  ctx->mass.assign(mass, mass + ntypes + 1);
  const size_t n = (size_t)nlocal;
  RESERVE(ctx, ctx->xq, n);
  RESERVE(ctx, ctx->v, n);
  RESERVE(ctx, ctx->f, n);
  RESERVE(ctx, ctx->type, n);
  RESERVE(ctx, ctx->tag, n);
  if (ctx->prec == B200MD_PREC_MIXED) RESERVE(ctx, ctx->xqf, n);
  // staging: x(3n) v(3n) q(n) doubles + type(n) ints
  RESERVE(ctx, ctx->stage, 8 * n + 16);
  double *sx = ctx->stage.p, *sv = sx + 3 * n, *sq = sv + 3 * n;
  int *st = (int *)(sq + n);
  CUDA_OK(ctx, cudaMemcpyAsync(sx, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (v) CUDA_OK(ctx, cudaMemcpyAsync(sv, v, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (q) CUDA_OK(ctx, cudaMemcpyAsync(sq, q, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(st, type, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ctx->first_id = 0;
  if (b2_comm_nranks(ctx) > 1) {   // global ids: rank r's atoms follow those of ranks < r
    std::vector<int> counts(b2_comm_nranks(ctx));
    TRY(b2_comm_allgather_int(ctx, nlocal, counts.data()));
    for (int r = 0; r < b2_comm_rank(ctx); r++) ctx->first_id += counts[r];
  }
  if (nlocal) {
    pack_atoms<<<cdiv(nlocal, 256), 256, 0, ctx->stream>>>(nlocal, sx, v ? sv : nullptr, q ? sq : nullptr, st,
                                                           ctx->xq.p, ctx->v.p, ctx->tag.p, ctx->type.p, ctx->first_id);
    KERNEL_OK(ctx, "pack_atoms");
    CUDA_OK(ctx, cudaMemsetAsync(ctx->f.p, 0, n * sizeof(double4), ctx->stream));
  }
  TRY(b2_refresh_float_copy(ctx, 0, nlocal));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->neigh.ready = false;
  ctx->neigh.last_build = -1;
  ctx->nve_ready = false;
  ctx->nve_grouped = ctx->nve_has_rmass = false;   // a new set of atoms: group / rmass have to be given again
  return 0;
}

int b200md_atoms_set_x(b200md_ctx *ctx, const double *x) {
  if (!ctx || !x) return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_set_x: bad arguments");
  if (b2_comm_nranks(ctx) > 1)
    return b2_fail(ctx, B200MD_EINVAL, "host-order access is single-GPU only (atoms migrate between ranks)");
  ctx->ev_pre_valid = false;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->nlocal;
  if (!n) return 0;
  RESERVE(ctx, ctx->stage, 8 * n + 16);
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->stage.p, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  set_x_by_tag<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(ctx->nlocal, ctx->stage.p, ctx->tag.p, ctx->xq.p);
  KERNEL_OK(ctx, "set_x_by_tag");
  TRY(b2_refresh_float_copy(ctx, 0, ctx->nlocal));
  return 0;
}

int b200md_atoms_download_ids(b200md_ctx *ctx, int capacity, int *n_out, int *ids, double *x, double *v, double *f) {
  if (!ctx || !n_out) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  const int n = ctx->nlocal;
  *n_out = n;
  if (n > capacity) return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_download_ids: capacity %d < %d owned atoms", capacity, n);
  if (!n) return 0;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<double4> h((size_t)n);
  auto fetch = [&](const double4 *src, double *dst) -> int {
    CUDA_OK(ctx, cudaMemcpy(h.data(), src, (size_t)n * sizeof(double4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) { dst[3 * (size_t)i] = h[i].x; dst[3 * (size_t)i + 1] = h[i].y; dst[3 * (size_t)i + 2] = h[i].z; }
    return 0;
  };
  if (ids) CUDA_OK(ctx, cudaMemcpy(ids, ctx->tag.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  if (x) TRY(fetch(ctx->xq.p, x));
  if (v) TRY(fetch(ctx->v.p, v));
  if (f) TRY(fetch(ctx->f.p, f));
  return 0;
}

int b200md_atoms_download(b200md_ctx *ctx, double *x, double *v, double *f, double *eatom) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  if (b2_comm_nranks(ctx) > 1)
    return b2_fail(ctx, B200MD_EINVAL, "host-order download is single-GPU only: use b200md_atoms_download_ids");
  const size_t n = (size_t)ctx->nlocal;
  if (!n) return 0;
  RESERVE(ctx, ctx->stage, 8 * n + 16);
  double *s3 = ctx->stage.p, *sw = s3 + 3 * n;
  const int nb = cdiv(ctx->nlocal, 256);
  if (x) {
    unpack_by_tag<<<nb, 256, 0, ctx->stream>>>(ctx->nlocal, ctx->xq.p, ctx->tag.p, s3, nullptr);
    KERNEL_OK(ctx, "unpack_by_tag");
    CUDA_OK(ctx, cudaMemcpyAsync(x, s3, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (v) {
    unpack_by_tag<<<nb, 256, 0, ctx->stream>>>(ctx->nlocal, ctx->v.p, ctx->tag.p, s3, nullptr);
    KERNEL_OK(ctx, "unpack_by_tag");
    CUDA_OK(ctx, cudaMemcpyAsync(v, s3, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (f || eatom) {
    unpack_by_tag<<<nb, 256, 0, ctx->stream>>>(ctx->nlocal, ctx->f.p, ctx->tag.p, f ? s3 : nullptr,
                                               eatom ? sw : nullptr);
    KERNEL_OK(ctx, "unpack_by_tag");
    if (f) CUDA_OK(ctx, cudaMemcpyAsync(f, s3, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (eatom) CUDA_OK(ctx, cudaMemcpyAsync(eatom, sw, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int b200md_atoms_count(const b200md_ctx *ctx, int *nlocal, int *nghost) {
  if (!ctx) return B200MD_EINVAL;
  if (nlocal) *nlocal = ctx->nlocal;
  if (nghost) *nghost = ctx->nghost;
  return 0;
}

int b200md_timers_enable(b200md_ctx *ctx, int on) {
  if (!ctx) return B200MD_EINVAL;
  ctx->timers_on = on != 0;
  return 0;
}
int b200md_timers_get(b200md_ctx *ctx, double *ms, long *calls, int n) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  harvest_timers(ctx);
  for (int i = 0; i < n && i < T_COUNT; i++) {
    if (ms) ms[i] = ctx->t_ms[i];
    if (calls) calls[i] = ctx->t_calls[i];
  }
  return 0;
}
int b200md_timers_reset(b200md_ctx *ctx) {
  if (!ctx) return B200MD_EINVAL;
  harvest_timers(ctx);
  for (int i = 0; i < T_COUNT; i++) { ctx->t_ms[i] = 0; ctx->t_calls[i] = 0; }
  return 0;
}
// kind 0: FP64 FMA TFLOP/s, 1: FP32 FMA TFLOP/s, 2: HBM copy GB/s (read + write bytes).  Best of 5.
int b200md_microbench(b200md_ctx *ctx, int kind, double *value) {
  if (!ctx || !value) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0.0;
  DevBuf<double2> buf;
  const size_t ncopy = (size_t)1 << 26;  // 2 x 1 GiB
  if (kind == 2 && buf.reserve(2 * ncopy)) return b2_fail(ctx, B200MD_ENOMEM, "microbench: out of memory");
  RESERVE(ctx, ctx->ev_out, 32);
  for (int rep = 0; rep < 6; rep++) {
    const int iters = 1 << 14, blocks = ctx->sm_count * 8;
    cudaEventRecord(a, ctx->stream);
    if (kind == 0) k_fma_peak<double><<<blocks, 256, 0, ctx->stream>>>(iters, 1.0, ctx->ev_out.p);
    else if (kind == 1) k_fma_peak<float><<<blocks, 256, 0, ctx->stream>>>(iters, 1.0f, (float *)ctx->ev_out.p);
    else k_copy16<<<ctx->sm_count * 16, 512, 0, ctx->stream>>>(ncopy, buf.p, buf.p + ncopy);
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    ctx->launches++;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep == 0) continue;  // warm-up
    double v;
    if (kind == 2) v = 2.0 * ncopy * sizeof(double2) / (ms * 1e-3) / 1e9;
    else v = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
    best = v > best ? v : best;
  }
  buf.free_();
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return b2_fail(ctx, B200MD_ECUDA, "microbench failed: %s", cudaGetErrorString(e));
  *value = best;
  return 0;
}

int b200md_timer_count(void) { return T_COUNT; }
const char *b200md_timer_name(int i) { return i >= 0 && i < T_COUNT ? kTimerNames[i] : ""; }
long b200md_launch_count(const b200md_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
