// pair.cu — the four Buckingham pair kernels (sm_100a).
//
// Replaces eval<EVFLAG,EFLAG,NEWTON_PAIR> of
//   PairBuckIntel               pair_buck_intel.cpp:127-365        (inner jj loop :241-317)
//   PairBuckCoulCutIntel        pair_buck_coul_cut_intel.cpp:134-402 (:259-353)
//   PairBuckCoulLongIntel       pair_buck_coul_long_intel.cpp:134-453 (:275-405; erfc :296-307; table :317-340)
//   PairBuckLongCoulLongIntel   pair_buck_long_coul_long_intel.cpp:215-539 (Coulomb :350-409, dispersion :410-473)
// and pack_force_const of each (…:391-443, :431-492, :481-566, :573-646).
//
// Design (B200): FULL neighbour list, newton off — each owned atom accumulates only its own force, so
// there are no force atomics and no thread-private force arrays to reduce (the reference's
// IP_PRE_fdotr_acc_force, pair_buck_intel.cpp:332-334, disappears).  TPA lanes cooperate on one atom:
// consecutive lanes read consecutive CSR entries (coalesced), gather {x,y,z,q} of j as one 32 B sector
// (16 B in mixed mode), accumulate in double, and combine with a fixed xor-shuffle tree => bitwise
// reproducible forces.  Energy/virial: per-block partial sums in a fixed order, then one single-block
// reduction — no FP atomics.  Per-type-pair constants live in shared memory.
// Cut-off test is rsq < cutsq.  Every list entry carries ev_pre = 1/2 (i is owned; the
// mirrored entry supplies the other half), which reproduces the NEWTON_PAIR=0 tallies of :296-313.
#include "pair_kernel.cuh"

using namespace pairk;

int b2_launch_pair_double(b200md_ctx *ctx, const PairView &v, long long total_entries, int evflag, double *ev_dev,
                          int has_special) {
  return launch_pair<double>(ctx, v, total_entries, evflag, ev_dev, has_special);
}

namespace {

int finish_ev(b200md_ctx *ctx, int eflag, int vflag, double *ev) {
  // copy back ev_out[0..8) and mask by the flags actually requested (ev_setup semantics)
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->ev_out.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 8; k++) ev[k] = 0.0;
  if (eflag & 1) { ev[0] = ctx->h_pinned[0]; ev[1] = ctx->h_pinned[1]; }
  if (vflag & 3) for (int k = 2; k < 8; k++) ev[k] = ctx->h_pinned[k];
  return 0;
}

__global__ void k_pack_host_atoms(int n, const double *__restrict__ x, const double *__restrict__ q,
                                  double4 *__restrict__ xq, float4 *__restrict__ xqf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = make_double4(x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2], q ? q[i] : 0.0);
  xq[i] = p;
  if (xqf) xqf[i] = make_float4((float)p.x, (float)p.y, (float)p.z, (float)p.w);
}

}  // namespace

int b2_pair_compute(b200md_ctx *ctx, int eflag, int vflag, double *ev) {
  if (!ctx->pair.ready) return b2_fail(ctx, B200MD_EINVAL, "pair compute before b200md_pair_setup");
  if (!ctx->neigh.ready) return b2_fail(ctx, B200MD_EINVAL, "pair compute before a neighbour build");
  const int evflag = ((eflag & 3) || (vflag & 3)) ? 1 : 0;
  if (evflag && !ev) return b2_fail(ctx, B200MD_EINVAL, "ev is NULL but energy/virial requested");
  // everything the k-space solver reads (positions, types) is final here: mark it, so that PPPM may start on its own
  // stream underneath this kernel
  ctx->ev_pre_valid = false;
  if (ctx->overlap && (ctx->pppm || ctx->pppm6)) {
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_pre, ctx->stream));
    ctx->ev_pre_valid = true;
  }
  ScopedTimer tm(ctx, T_PAIR);
  RESERVE(ctx, ctx->ev_out, 32);
  PairView v;
  v.nlocal = ctx->nlocal;
  v.x = ctx->prec == B200MD_PREC_MIXED ? (const void *)ctx->xqf.p : (const void *)ctx->xq.p;
  v.type = ctx->type.p;
  v.numneigh = ctx->neigh.numneigh.p;
  v.offsets = ctx->neigh.offsets.p;
  v.entries = ctx->neigh.entries.p;
  v.f = ctx->f.p;
  v.packed_type = ctx->neigh.packed_type ? 1 : 0;
  // lists built on the device carry special-bond bits only for molecular systems (b200md_atoms_set_special)
  const int has_special = ctx->sp_max > 0 ? 1 : 0;
  if (ctx->prec == B200MD_PREC_MIXED)
    TRY(b2_launch_pair_float(ctx, v, ctx->neigh.total_entries, evflag, ctx->ev_out.p, has_special));
  else TRY(b2_launch_pair_double(ctx, v, ctx->neigh.total_entries, evflag, ctx->ev_out.p, has_special));
  if (evflag) {
    TRY(b2_comm_allreduce_sum(ctx, ctx->ev_out.p, 8));   // global tallies over the ranks (no-op on one GPU)
    TRY(finish_ev(ctx, eflag, vflag, ev));
  }
  return 0;
}

extern "C" {

int b200md_pair_setup(b200md_ctx *ctx, const b200md_pair_params *p) {
  if (!ctx || !p) return b2_fail(ctx, B200MD_EINVAL, "b200md_pair_setup: NULL argument");
  cudaSetDevice(ctx->device);
  if (p->style < B200MD_PAIR_BUCK || p->style > B200MD_PAIR_LJ_LONG_COUL_LONG)
    return b2_fail(ctx, B200MD_EINVAL, "unknown pair style %d", p->style);
  if (p->ntypes < 1 || p->ntypes > B2_MAXTYPES)
    return b2_fail(ctx, B200MD_EINVAL, "ntypes %d outside 1..%d", p->ntypes, B2_MAXTYPES);
  if (!p->cutsq || !p->cut_ljsq || !p->buck1 || !p->buck2 || !p->rhoinv || !p->a || !p->c || !p->offset)
    return b2_fail(ctx, B200MD_EINVAL, "b200md_pair_setup: missing per-type-pair arrays");
  const bool coul = p->style != B200MD_PAIR_BUCK;
  if (p->style == B200MD_PAIR_BUCK_COUL_CUT && !p->cut_coulsq)
    return b2_fail(ctx, B200MD_EINVAL, "buck/coul/cut needs cut_coulsq");
  PairState &ps = ctx->pair;
  ps.p = *p;
  const int tp1 = p->ntypes + 1;
  ps.tp1 = tp1;
  const int n = tp1 * tp1;
  std::vector<double> hd((size_t)n * C_N, 0.0);
  std::vector<float> hf((size_t)n * C_N, 0.0f);
  ps.h_cutsq.assign(n, 0.0);
  double cutmax = 0.0;
  bool same_cut = true;
  for (int i = 1; i < tp1; i++)
    for (int j = 1; j < tp1; j++) {
      const int ij = i * tp1 + j;
      double *d = &hd[(size_t)ij * C_N];
      d[C_CUTSQ] = p->cutsq[ij];
      d[C_CUT_LJSQ] = p->cut_ljsq[ij];
      d[C_CUT_COULSQ] = (coul && p->cut_coulsq) ? p->cut_coulsq[ij] : 0.0;
      d[C_BUCK1] = p->buck1[ij];
      d[C_BUCK2] = p->buck2[ij];
      d[C_RHOINV] = p->rhoinv[ij];
      d[C_A] = p->a[ij];
      d[C_C] = p->c[ij];
      d[C_OFFSET] = p->offset[ij];
      for (int k = 0; k < C_N; k++) hf[(size_t)ij * C_N + k] = (float)d[k];
      ps.h_cutsq[ij] = p->cutsq[ij];
      if (p->cut_ljsq[ij] != p->cutsq[ij]) same_cut = false;
      cutmax = std::max(cutmax, std::sqrt(p->cutsq[ij]));
    }
  ps.cutmax = cutmax;
  ps.same_cut = same_cut;
  RESERVE(ctx, ps.coeff_d, hd.size());
  RESERVE(ctx, ps.coeff_f, hf.size());
  CUDA_OK(ctx, cudaMemcpy(ps.coeff_d.p, hd.data(), hd.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(ctx, cudaMemcpy(ps.coeff_f.p, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice));
  {
    // 2^(k/64), correctly rounded: fast_exp's table (pair_kernel.cuh)
    double et[B2_EXP_TAB];
    for (int k = 0; k < B2_EXP_TAB; k++) et[k] = (double)exp2l((long double)k / B2_EXP_TAB);
    RESERVE(ctx, ps.exptab, B2_EXP_TAB);
    CUDA_OK(ctx, cudaMemcpy(ps.exptab.p, et, sizeof(et), cudaMemcpyHostToDevice));
    // fast_exp adds n>>6 to the exponent field without an underflow path: arguments stay above -700
    // (and never positive): every exponential of the loop is checked here — exp(-r/rho), exp(-(g_ewald r)^2) of the
    // real-space Ewald term, exp(-g_ewald_6^2 r^2) of the buck/long dispersion term
    for (int i = 1; i < tp1; i++)
      for (int j = 1; j < tp1; j++) {
        if (p->style == B200MD_PAIR_LJ_LONG_COUL_LONG) continue;   // no exp(-r/rho) in the Lennard-Jones style
        if (!(p->rhoinv[i * tp1 + j] > 0.0))
          return b2_fail(ctx, B200MD_EINVAL, "buck rho for types %d-%d must be positive", i, j);
        if (std::sqrt(p->cut_ljsq[i * tp1 + j]) * p->rhoinv[i * tp1 + j] > 700.0)
          return b2_fail(ctx, B200MD_EINVAL, "buck rho for types %d-%d is too small for its cutoff: exp(-r/rho) underflows", i, j);
      }
    const bool lcl = p->style == B200MD_PAIR_BUCK_LONG_COUL_LONG || p->style == B200MD_PAIR_LJ_LONG_COUL_LONG;
    const bool ewald1 = p->style == B200MD_PAIR_BUCK_COUL_LONG || (lcl && ((p->ewald_order >> 1) & 1));
    const bool ewald6 = lcl && ((p->ewald_order >> 6) & 1);
    if (ewald1 && p->g_ewald * cutmax * p->g_ewald * cutmax > 700.0)
      return b2_fail(ctx, B200MD_EINVAL, "g_ewald %g is too large for the cutoff %g: exp(-(g r)^2) underflows", p->g_ewald, cutmax);
    if (ewald6 && p->g_ewald_6 * cutmax * p->g_ewald_6 * cutmax > 700.0)
      return b2_fail(ctx, B200MD_EINVAL, "g_ewald_6 %g is too large for the cutoff %g: exp(-(g r)^2) underflows", p->g_ewald_6, cutmax);
  }
  // tables
  if (p->ncoultablebits) {
    if (!p->rtable || !p->drtable || !p->ftable || !p->dftable || !p->etable || !p->detable || !p->ctable ||
        !p->dctable)
      return b2_fail(ctx, B200MD_EINVAL, "ncoultablebits set but a Coulomb table is NULL");
    const int nt = 1 << p->ncoultablebits;
    std::vector<double> td((size_t)nt * 8);
    std::vector<float> tf((size_t)nt * 8);
    const double *src[8] = {p->rtable, p->drtable, p->ftable, p->dftable, p->etable, p->detable, p->ctable, p->dctable};
    for (int t = 0; t < nt; t++)
      for (int k = 0; k < 8; k++) {
        td[(size_t)t * 8 + k] = src[k][t];
        tf[(size_t)t * 8 + k] = (float)src[k][t];
      }
    RESERVE(ctx, ps.ctab_d, td.size());
    RESERVE(ctx, ps.ctab_f, tf.size());
    CUDA_OK(ctx, cudaMemcpy(ps.ctab_d.p, td.data(), td.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(ctx, cudaMemcpy(ps.ctab_f.p, tf.data(), tf.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (p->ndisptablebits) {
    if (!p->rdisptable || !p->drdisptable || !p->fdisptable || !p->dfdisptable || !p->edisptable || !p->dedisptable)
      return b2_fail(ctx, B200MD_EINVAL, "ndisptablebits set but a dispersion table is NULL");
    const int nt = 1 << p->ndisptablebits;
    std::vector<double> td((size_t)nt * 6);
    std::vector<float> tf((size_t)nt * 6);
    const double *src[6] = {p->rdisptable, p->drdisptable, p->fdisptable, p->dfdisptable, p->edisptable, p->dedisptable};
    for (int t = 0; t < nt; t++)
      for (int k = 0; k < 6; k++) {
        td[(size_t)t * 6 + k] = src[k][t];
        tf[(size_t)t * 6 + k] = (float)src[k][t];
      }
    RESERVE(ctx, ps.dtab_d, td.size());
    RESERVE(ctx, ps.dtab_f, tf.size());
    CUDA_OK(ctx, cudaMemcpy(ps.dtab_d.p, td.data(), td.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(ctx, cudaMemcpy(ps.dtab_f.p, tf.data(), tf.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  // pointers in the kept copy are host pointers of the caller: drop them
  ps.p.cutsq = ps.p.cut_ljsq = ps.p.cut_coulsq = ps.p.buck1 = ps.p.buck2 = ps.p.rhoinv = nullptr;
  ps.p.a = ps.p.c = ps.p.offset = nullptr;
  ps.p.rtable = ps.p.drtable = ps.p.ftable = ps.p.dftable = ps.p.etable = ps.p.detable = nullptr;
  ps.p.ctable = ps.p.dctable = nullptr;
  ps.p.rdisptable = ps.p.drdisptable = ps.p.fdisptable = ps.p.dfdisptable = ps.p.edisptable = ps.p.dedisptable = nullptr;
  ps.ready = true;
  ctx->neigh.ready = false;
  return 0;
}

int b200md_pair_compute(b200md_ctx *ctx, int eflag, int vflag, double ev[8]) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  return b2_pair_compute(ctx, eflag, vflag, ev);
}

int b200md_pair_eval_host(b200md_ctx *ctx, int eflag, int vflag, int nlocal, int nall, const double *x,
                          const int *type, const double *q, const int *numneigh, const long *cnumneigh,
                          const int *firstneigh, double *f, double ev[8]) {
  if (!ctx || !x || !type || !numneigh || !cnumneigh || !firstneigh || !f || nlocal < 0 || nall < nlocal)
    return b2_fail(ctx, B200MD_EINVAL, "b200md_pair_eval_host: bad arguments");
  if (!ctx->pair.ready) return b2_fail(ctx, B200MD_EINVAL, "pair eval before b200md_pair_setup");
  cudaSetDevice(ctx->device);
  const int evflag = ((eflag & 3) || (vflag & 3)) ? 1 : 0;
  if (evflag && !ev) return b2_fail(ctx, B200MD_EINVAL, "ev is NULL but energy/virial requested");
  // the caller's arrays index device memory (per-type rows of the constant table, gathered atoms): check them here, this
  // entry point is the boundary for host-built lists
  const int ntypes = ctx->pair.tp1 - 1;
  for (int i = 0; i < nall; i++)
    if (type[i] < 1 || type[i] > ntypes)
      return b2_fail(ctx, B200MD_EINVAL, "b200md_pair_eval_host: type[%d] = %d outside 1..%d", i, type[i], ntypes);
  long long total = 0;
  int has_special = 0;
  for (int i = 0; i < nlocal; i++) {
    if (numneigh[i] < 0 || cnumneigh[i] < 0)
      return b2_fail(ctx, B200MD_EINVAL, "b200md_pair_eval_host: negative count or offset for atom %d", i);
    total = std::max(total, (long long)cnumneigh[i] + numneigh[i]);
    for (int k = 0; k < numneigh[i]; k++) {
      const int e = firstneigh[cnumneigh[i] + k];
      if ((e & B2_NEIGHMASK) >= nall)
        return b2_fail(ctx, B200MD_EINVAL, "b200md_pair_eval_host: list entry %d of atom %d points past nall = %d",
                       e & B2_NEIGHMASK, i, nall);
      if ((unsigned)e >> B2_SBBITS) has_special = 1;
    }
  }
  DevBuf<double> dx, dq;
  DevBuf<double4> dxq, df;
  DevBuf<float4> dxqf;
  DevBuf<int> dtype, dnum, dent;
  DevBuf<long long> doff;
  int rc = 0;
  auto cleanup = [&]() {
    dx.free_(); dq.free_(); dxq.free_(); df.free_(); dxqf.free_(); dtype.free_(); dnum.free_(); dent.free_(); doff.free_();
  };
  const bool mixed = ctx->prec == B200MD_PREC_MIXED;
  if (dx.reserve(3 * (size_t)nall + 1) || dq.reserve((size_t)nall + 1) || dxq.reserve((size_t)nall + 1) ||
      df.reserve((size_t)nlocal + 1) || (mixed && dxqf.reserve((size_t)nall + 1)) || dtype.reserve((size_t)nall + 1) ||
      dnum.reserve((size_t)nlocal + 1) || dent.reserve((size_t)total + 1) || doff.reserve((size_t)nlocal + 1) ||
      ctx->ev_out.reserve(32)) {
    cleanup();
    return b2_fail(ctx, B200MD_ENOMEM, "out of device memory in b200md_pair_eval_host");
  }
  cudaStream_t s = ctx->stream;
  cudaMemcpyAsync(dx.p, x, 3 * (size_t)nall * sizeof(double), cudaMemcpyHostToDevice, s);
  if (q) cudaMemcpyAsync(dq.p, q, (size_t)nall * sizeof(double), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dtype.p, type, (size_t)nall * sizeof(int), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dnum.p, numneigh, (size_t)nlocal * sizeof(int), cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(doff.p, cnumneigh, (size_t)nlocal * sizeof(long long), cudaMemcpyHostToDevice, s);
  if (total) cudaMemcpyAsync(dent.p, firstneigh, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, s);
  if (nall) {
    k_pack_host_atoms<<<cdiv(nall, 256), 256, 0, s>>>(nall, dx.p, q ? dq.p : nullptr, dxq.p, mixed ? dxqf.p : nullptr);
    ctx->launches++;
  }
  PairView v;
  v.nlocal = nlocal;
  v.x = mixed ? (const void *)dxqf.p : (const void *)dxq.p;
  v.type = dtype.p; v.numneigh = dnum.p; v.offsets = doff.p; v.entries = dent.p; v.f = df.p; v.packed_type = 0;
  rc = mixed ? b2_launch_pair_float(ctx, v, total, evflag, ctx->ev_out.p, has_special)
             : b2_launch_pair_double(ctx, v, total, evflag, ctx->ev_out.p, has_special);
  if (!rc) {
    cudaMemcpyAsync(f, df.p, (size_t)nlocal * sizeof(double4), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = b2_fail(ctx, B200MD_ECUDA, "pair eval failed: %s", cudaGetErrorString(e));
  }
  if (!rc && evflag) rc = finish_ev(ctx, eflag, vflag, ev);
  cleanup();
  return rc;
}

}  // extern "C"
