// nve.cu — fix nve/intel on the device, plus the resident Verlet loop.
//
// Replaces FixNVEIntel::initial_integrate / final_integrate / reset_dt (fix_nve_intel.cpp:60-99,
// 103-127, 129-194; group all, per-type mass) and the stock Verlet::run order the reference plugs into
// (SURVEY.md §3.1, App. A.6).  The arithmetic is un-fused (mul then add), like the reference's AVX build,
// so x and v are bit-identical to the CPU oracle step for step given identical forces.
// dtf/mass rides in v.w; the mixed-mode float copy of the positions (IntelBuffers::thr_pack,
// intel_buffers.h:185-203) is written by the same kernel instead of a separate pack pass.
#include <vector>

#include "internal.h"

namespace {

// {x,y,z,w}[n] -> [n][3] in the same (device) order
__global__ void k_pack3(int n, const double4 *__restrict__ src, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 v = src[i];
  out[3 * (size_t)i] = v.x;
  out[3 * (size_t)i + 1] = v.y;
  out[3 * (size_t)i + 2] = v.z;
}

// _dtfm of FixNVEIntel::reset_dt (fix_nve_intel.cpp:147-190): dtf / mass[type] or dtf / rmass[i]; 0 outside the group
__global__ void k_nve_set_dtfm(int n, const int *__restrict__ type, const double *__restrict__ mass, double dtf,
                               double4 *__restrict__ v, const int *__restrict__ tag, int first_id,
                               const int *__restrict__ grp, const double *__restrict__ rmass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 vi = v[i];
  const int t = tag[i] - first_id;
  const double m = rmass ? rmass[t] : mass[type[i]];
  vi.w = (grp && !grp[t]) ? 0.0 : dtf / m;
  v[i] = vi;
}

__global__ void k_nve_initial(int n, double dtv, double4 *__restrict__ xq, double4 *__restrict__ v,
                              const double4 *__restrict__ f, float4 *__restrict__ xqf, int grouped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 vi = v[i];
  if (grouped && vi.w == 0.0) return;   // `if (_dtfm[i] != 0.0)`, fix_nve_intel.cpp:92: neither v nor x moves
  const double4 fi = f[i];
  double4 xi = xq[i];
  vi.x = __dadd_rn(vi.x, __dmul_rn(vi.w, fi.x));
  vi.y = __dadd_rn(vi.y, __dmul_rn(vi.w, fi.y));
  vi.z = __dadd_rn(vi.z, __dmul_rn(vi.w, fi.z));
  xi.x = __dadd_rn(xi.x, __dmul_rn(dtv, vi.x));
  xi.y = __dadd_rn(xi.y, __dmul_rn(dtv, vi.y));
  xi.z = __dadd_rn(xi.z, __dmul_rn(dtv, vi.z));
  v[i] = vi;
  xq[i] = xi;
  if (xqf) xqf[i] = make_float4((float)xi.x, (float)xi.y, (float)xi.z, (float)xi.w);
}

__global__ void k_nve_final(int n, double4 *__restrict__ v, const double4 *__restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 vi = v[i];
  const double4 fi = f[i];
  vi.x = __dadd_rn(vi.x, __dmul_rn(vi.w, fi.x));
  vi.y = __dadd_rn(vi.y, __dmul_rn(vi.w, fi.y));
  vi.z = __dadd_rn(vi.z, __dmul_rn(vi.w, fi.z));
  v[i] = vi;
}

// sum 1/2 m v^2: per-block partials in a fixed order, reduced by one block
__global__ void __launch_bounds__(256) k_ke_partial(int n, const double4 *__restrict__ v, const int *__restrict__ type,
                                                    const double *__restrict__ mass, double *__restrict__ partial,
                                                    const int *__restrict__ tag, int first_id,
                                                    const double *__restrict__ rmass) {
  __shared__ double s[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double a = 0.0;
  if (i < n) {
    const double4 vi = v[i];
    const double m = rmass ? rmass[tag[i] - first_id] : mass[type[i]];
    a = 0.5 * m * (vi.x * vi.x + vi.y * vi.y + vi.z * vi.z);
  }
  s[threadIdx.x] = a;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
__global__ void __launch_bounds__(256) k_sum1(int n, const double *__restrict__ partial, double *__restrict__ out) {
  __shared__ double s[256];
  double a = 0.0;
  for (int r = threadIdx.x; r < n; r += 256) a += partial[r];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

}  // namespace

static int upload_mass(b200md_ctx *ctx, double **dmass) {
  // mass table sits behind ev_out (ev_out[16..32))
  RESERVE(ctx, ctx->ev_out, 32);
  if ((int)ctx->mass.size() > 16) return b2_fail(ctx, B200MD_EINVAL, "too many atom types");
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->ev_out.p + 16, ctx->mass.data(), ctx->mass.size() * sizeof(double),
                               cudaMemcpyHostToDevice, ctx->stream));
  *dmass = ctx->ev_out.p + 16;
  return 0;
}

int b2_nve_initial(b200md_ctx *ctx) {
  ctx->ev_pre_valid = false;   // positions change: a pending k-space overlap marker is stale
  if (!ctx->nve_ready) return b2_fail(ctx, B200MD_EINVAL, "nve integrate before b200md_nve_setup");
  ScopedTimer tm(ctx, T_NVE);
  if (ctx->nlocal == 0) return 0;
  ScopedTimer tk(ctx, K_NVE_INITIAL);
  k_nve_initial<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(
      ctx->nlocal, ctx->dtv, ctx->xq.p, ctx->v.p, ctx->f.p, ctx->prec == B200MD_PREC_MIXED ? ctx->xqf.p : nullptr,
      ctx->nve_grouped ? 1 : 0);
  KERNEL_OK(ctx, "k_nve_initial");
  return 0;
}

int b2_nve_final(b200md_ctx *ctx) {
  if (!ctx->nve_ready) return b2_fail(ctx, B200MD_EINVAL, "nve integrate before b200md_nve_setup");
  ScopedTimer tm(ctx, T_NVE);
  if (ctx->nlocal == 0) return 0;
  ScopedTimer tk(ctx, K_NVE_FINAL);
  k_nve_final<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(ctx->nlocal, ctx->v.p, ctx->f.p);
  KERNEL_OK(ctx, "k_nve_final");
  return 0;
}

int b2_kinetic_energy(b200md_ctx *ctx, double *ke) {
  *ke = 0.0;
  if (ctx->nlocal == 0 && b2_comm_nranks(ctx) == 1) return 0;
  double *dmass;
  TRY(upload_mass(ctx, &dmass));
  const int nb = cdiv(ctx->nlocal, 256);
  RESERVE(ctx, ctx->ev_partial, (size_t)nb + 1);
  if (nb > 0) {
    k_ke_partial<<<nb, 256, 0, ctx->stream>>>(ctx->nlocal, ctx->v.p, ctx->type.p, dmass, ctx->ev_partial.p, ctx->tag.p,
                                              0 /* global ids */, ctx->nve_has_rmass ? ctx->nve_rmass.p : nullptr);
    KERNEL_OK(ctx, "k_ke_partial");
  }
  k_sum1<<<1, 256, 0, ctx->stream>>>(nb, ctx->ev_partial.p, ctx->ev_out.p + 8);
  KERNEL_OK(ctx, "k_sum1");
  TRY(b2_comm_allreduce_sum(ctx, ctx->ev_out.p + 8, 1));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->ev_out.p + 8, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  *ke = ctx->h_pinned[0];
  return 0;
}

static int forces(b200md_ctx *ctx, int eflag, int vflag, double *thermo) {
  double ev[8] = {0};
  TRY(b2_pair_compute(ctx, eflag, vflag, (eflag || vflag) ? ev : nullptr));
  double ek = 0.0, vk[6] = {0};
  if (ctx->pppm || ctx->pppm6) TRY(b2_pppm_compute(ctx, eflag, vflag, &ek, vk));
  if (thermo) {
    for (int k = 0; k < 8; k++) thermo[k] = ev[k];
    thermo[8] = ek;
    for (int k = 0; k < 6; k++) thermo[9 + k] = vk[k];
  }
  return 0;
}

extern "C" {

int b200md_nve_setup(b200md_ctx *ctx, double dt) {
  if (!ctx || !(dt > 0)) return b2_fail(ctx, B200MD_EINVAL, "b200md_nve_setup: bad timestep");
  cudaSetDevice(ctx->device);
  ctx->dt = dt;
  ctx->dtv = dt;                     // FixNVEIntel::reset_dt, fix_nve_intel.cpp:130-131
  ctx->dtf = 0.5 * dt * ctx->ftm2v;
  if (ctx->nlocal > 0) {
    double *dmass;
    TRY(upload_mass(ctx, &dmass));
    k_nve_set_dtfm<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(
        ctx->nlocal, ctx->type.p, dmass, ctx->dtf, ctx->v.p, ctx->tag.p, 0 /* tables are indexed by global id */,
        ctx->nve_grouped ? ctx->nve_group.p : nullptr, ctx->nve_has_rmass ? ctx->nve_rmass.p : nullptr);
    KERNEL_OK(ctx, "k_nve_set_dtfm");
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->nve_ready = true;
  return 0;
}

int b200md_nve_set_group(b200md_ctx *ctx, const int *ingroup, const double *rmass) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->nlocal;
  const bool multi = b2_comm_nranks(ctx) > 1;
  // several GPUs: collective, every rank passes the same kind of arrays (a rank without atoms may pass any non-NULL
  // pointer); the tables are gathered over the ranks and indexed by global id, so they follow migrating atoms
  ctx->nve_grouped = ingroup != nullptr && (n > 0 || multi);
  ctx->nve_has_rmass = rmass != nullptr && (n > 0 || multi);
  ctx->nve_ready = false;   // _dtfm has to be rebuilt (reset_dt)
  if (ctx->nve_grouped) {
    std::vector<int> g(n + 1);
    for (size_t i = 0; i < n; i++) g[i] = ingroup[i] ? 1 : 0;
    TRY(b2_atoms_global_table<int>(ctx, g.data(), 1, ctx->nve_group, nullptr));
  }
  if (ctx->nve_has_rmass) {
    int bad = 0;
    for (size_t i = 0; i < n; i++)
      if (!(rmass[i] > 0.0)) bad = 1;
    if (multi) {   // agree on the verdict before the collective gather
      std::vector<int> all(b2_comm_nranks(ctx));
      TRY(b2_comm_allgather_int(ctx, bad, all.data()));
      for (int v : all) bad |= v;
    }
    if (bad) return b2_fail(ctx, B200MD_EINVAL, "per-atom mass must be positive");
    TRY(b2_atoms_global_table<double>(ctx, rmass, 1, ctx->nve_rmass, nullptr));
  }
  return 0;
}

int b200md_atoms_set_special(b200md_ctx *ctx, int maxspecial, const int *nspecial, const int *special) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  ctx->neigh.ready = false;   // the next list carries (or drops) the bits
  if (!nspecial || !special || maxspecial <= 0) {
    ctx->sp_max = 0;
    return 0;
  }
  const bool multi = b2_comm_nranks(ctx) > 1;
  const size_t n = (size_t)ctx->nlocal;
  // several GPUs: collective; partner ids are GLOBAL ids (b200md_atoms_download_ids), every rank passes the rows of its
  // own upload and the same maxspecial
  long nglobal = (long)n;
  int bad = maxspecial > 32 ? 1 : 0;
  if (multi) {
    std::vector<int> all(b2_comm_nranks(ctx));
    TRY(b2_comm_allgather_int(ctx, ctx->nlocal, all.data()));
    nglobal = 0;
    for (int v : all) nglobal += v;
    TRY(b2_comm_allgather_int(ctx, maxspecial, all.data()));
    for (int v : all)
      if (v != maxspecial) bad = 2;
  }
  size_t bad_atom = 0;
  if (!bad)
    for (size_t i = 0; i < n && !bad; i++) {
      const int a = nspecial[3 * i], b = nspecial[3 * i + 1], c = nspecial[3 * i + 2];
      if (a < 0 || b < a || c < b || c > maxspecial) { bad = 3; bad_atom = i; break; }
      for (int k = 0; k < c; k++)
        if (special[i * maxspecial + k] < 0 || (long)special[i * maxspecial + k] >= nglobal) { bad = 4; bad_atom = i; break; }
    }
  if (multi) {
    std::vector<int> all(b2_comm_nranks(ctx));
    const int mine = bad;
    TRY(b2_comm_allgather_int(ctx, mine, all.data()));
    for (int v : all)
      if (v && !bad) bad = 5;
  }
  if (bad == 1) return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_set_special: maxspecial %d > 32", maxspecial);
  if (bad == 2) return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_set_special: maxspecial differs between the ranks");
  if (bad == 3) return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_set_special: bad cumulative counts for atom %zu", bad_atom);
  if (bad == 4) return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_set_special: a partner of atom %zu is not an atom", bad_atom);
  if (bad) return b2_fail(ctx, B200MD_EINVAL, "b200md_atoms_set_special: rejected on another rank");
  TRY(b2_atoms_global_table<int>(ctx, nspecial, 3, ctx->sp_count, nullptr));
  TRY(b2_atoms_global_table<int>(ctx, special, (size_t)maxspecial, ctx->sp_list, nullptr));
  ctx->sp_max = maxspecial;
  return 0;
}

int b200md_nve_initial_integrate(b200md_ctx *ctx) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  return b2_nve_initial(ctx);
}

int b200md_nve_final_integrate(b200md_ctx *ctx) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  return b2_nve_final(ctx);
}

int b200md_setup_forces(b200md_ctx *ctx, int eflag, int vflag, double *thermo) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  TRY(b2_neigh_build(ctx));
  TRY(forces(ctx, eflag, vflag, thermo));
  if (thermo) TRY(b2_kinetic_energy(ctx, &thermo[15]));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int b200md_run_timed(b200md_ctx *ctx, long nsteps, double *thermo, double *elapsed_ms) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  TRY(b200md_run(ctx, nsteps, thermo));
  CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  CUDA_OK(ctx, cudaEventSynchronize(ctx->ev_b));
  float ms = 0;
  CUDA_OK(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

int b200md_step_host(b200md_ctx *ctx, const double *x_in, double *x_out, double *f_out) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  if (b2_comm_nranks(ctx) > 1)
    return b2_fail(ctx, B200MD_EINVAL, "b200md_step_host is single-GPU only (atoms migrate between ranks)");
  if (!ctx->neigh.ready) return b2_fail(ctx, B200MD_EINVAL, "b200md_step_host before b200md_setup_forces");
  if (x_in) TRY(b200md_atoms_set_x(ctx, x_in));
  const size_t n = (size_t)ctx->nlocal;
  RESERVE(ctx, ctx->stage, 8 * n + 16);
  ctx->ntimestep++;
  TRY(b2_nve_initial(ctx));
  int rebuilt = 0;
  TRY(b200md_neigh_decide(ctx, ctx->ntimestep, &rebuilt));
  // positions are final for this step: their download runs on the copy stream underneath the force kernels
  if (x_out && n) {
    TRY(b2_unpack_to_stage(ctx, ctx->xq.p, ctx->stage.p));
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));
    CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    CUDA_OK(ctx, cudaMemcpyAsync(x_out, ctx->stage.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
  }
  TRY(forces(ctx, 0, 0, nullptr));
  TRY(b2_nve_final(ctx));
  if (f_out && n) {
    TRY(b2_unpack_to_stage(ctx, ctx->f.p, ctx->stage.p + 3 * n));
    CUDA_OK(ctx, cudaMemcpyAsync(f_out, ctx->stage.p + 3 * n, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return 0;
}

// One timestep with the results handed to the host in DEVICE order together with the atoms' global ids — the form of
// b200md_step_host that also works on several GPUs, where atoms migrate between the ranks at every rebuild and the
// host therefore owns no fixed slice: positions stay resident (no upload), every step downloads this rank's ids, x and
// f.  *n_out = atoms this rank owns after the step's migration; fails if capacity is smaller.
int b200md_step_host_ids(b200md_ctx *ctx, int capacity, int *n_out, int *ids, double *x_out, double *f_out) {
  if (!ctx || !n_out) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  if (!ctx->neigh.ready) return b2_fail(ctx, B200MD_EINVAL, "b200md_step_host_ids before b200md_setup_forces");
  ctx->ntimestep++;
  TRY(b2_nve_initial(ctx));
  int rebuilt = 0;
  TRY(b200md_neigh_decide(ctx, ctx->ntimestep, &rebuilt));
  const size_t n = (size_t)ctx->nlocal;
  *n_out = ctx->nlocal;
  if (ctx->nlocal > capacity)
    return b2_fail(ctx, B200MD_EINVAL, "b200md_step_host_ids: capacity %d < %d owned atoms", capacity, ctx->nlocal);
  RESERVE(ctx, ctx->stage, 8 * n + 16);
  const int nb = cdiv(ctx->nlocal, 256);
  if (n) {   // positions and ids are final for this step: their download runs underneath the force kernels
    if (x_out) {
      k_pack3<<<nb, 256, 0, ctx->stream>>>(ctx->nlocal, ctx->xq.p, ctx->stage.p);
      KERNEL_OK(ctx, "k_pack3");
    }
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));
    CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    if (x_out) CUDA_OK(ctx, cudaMemcpyAsync(x_out, ctx->stage.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (ids) CUDA_OK(ctx, cudaMemcpyAsync(ids, ctx->tag.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
  }
  TRY(forces(ctx, 0, 0, nullptr));
  TRY(b2_nve_final(ctx));
  if (f_out && n) {
    k_pack3<<<nb, 256, 0, ctx->stream>>>(ctx->nlocal, ctx->f.p, ctx->stage.p + 3 * n);
    KERNEL_OK(ctx, "k_pack3");
    CUDA_OK(ctx, cudaMemcpyAsync(f_out, ctx->stage.p + 3 * n, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return 0;
}

int b200md_run(b200md_ctx *ctx, long nsteps, double *thermo) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  if (!ctx->neigh.ready) return b2_fail(ctx, B200MD_EINVAL, "b200md_run before b200md_setup_forces");
  for (long s = 0; s < nsteps; s++) {
    const bool last = (s == nsteps - 1) && thermo;
    ctx->ntimestep++;
    TRY(b2_nve_initial(ctx));
    int rebuilt = 0;
    TRY(b200md_neigh_decide(ctx, ctx->ntimestep, &rebuilt));
    TRY(forces(ctx, last ? 1 : 0, last ? 1 : 0, last ? thermo : nullptr));
    TRY(b2_nve_final(ctx));
    if (last) TRY(b2_kinetic_energy(ctx, &thermo[15]));
  }
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
