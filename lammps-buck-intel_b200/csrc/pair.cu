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
#include <algorithm>
#include <cmath>
#include <cstring>

#include "internal.h"

namespace {

enum { C_CUTSQ = 0, C_CUT_LJSQ, C_CUT_COULSQ, C_BUCK1, C_BUCK2, C_RHOINV, C_A, C_C, C_OFFSET, C_N };

template <class flt_t>
struct PairConsts {
  int tp1;
  flt_t qqrd2e, g_ewald, tabinnersq, tabinnerdispsq, g2, g6, g8;
  flt_t special_lj[4], special_coul[4];
  int ncoulmask, ncoulshiftbits, ndispmask, ndispshiftbits;
  int order1, order6, coultable, disptable;
};

template <class flt_t> struct V4;
template <> struct V4<double> { typedef double4 type; };
template <> struct V4<float> { typedef float4 type; };

__device__ __forceinline__ double m_exp(double x) { return exp(x); }
__device__ __forceinline__ float m_exp(float x) { return expf(x); }
__device__ __forceinline__ double m_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float m_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double m_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float m_rcp(float x) { return 1.0f / x; }

struct PairView {  // device pointers of one evaluation
  int nlocal;
  const void *x;         // double4* or float4*
  const int *type;
  const int *numneigh;
  const long long *offsets;
  const int *entries;
  double4 *f;
};

template <int STYLE, class flt_t, int EVFLAG, int TPA>
__global__ void __launch_bounds__(256)
k_pair(const int nlocal, const typename V4<flt_t>::type *__restrict__ x, const int *__restrict__ type,
       const int *__restrict__ numneigh, const long long *__restrict__ offsets,
       const int *__restrict__ entries, const PairConsts<flt_t> pc, const flt_t *__restrict__ coeff,
       const flt_t *__restrict__ ctab, const flt_t *__restrict__ dtab, double4 *__restrict__ f,
       double *__restrict__ ev_partial) {
  typedef typename V4<flt_t>::type vec4;
  __shared__ flt_t s_coeff[(B2_MAXTYPES + 1) * (B2_MAXTYPES + 1) * C_N];
  __shared__ double s_ev[8][8];
  for (int k = threadIdx.x; k < pc.tp1 * pc.tp1 * C_N; k += blockDim.x) s_coeff[k] = coeff[k];
  __syncthreads();

  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int sub = threadIdx.x & (TPA - 1);
  const int i = gtid / TPA;
  const bool active = i < nlocal;

  double fx = 0.0, fy = 0.0, fz = 0.0;
  double sevdwl = 0.0, secoul = 0.0, sv0 = 0.0, sv1 = 0.0, sv2 = 0.0, sv3 = 0.0, sv4 = 0.0, sv5 = 0.0;

  if (active) {
    const vec4 xi = x[i];
    const flt_t qtmp = xi.w;
    const flt_t *ci = s_coeff + type[i] * pc.tp1 * C_N;
    const int jnum = numneigh[i];
    const int *jlist = entries + offsets[i];

    for (int jj = sub; jj < jnum; jj += TPA) {
      const int e = jlist[jj];
      const int sbindex = (e >> B2_SBBITS) & 3;
      const int j = e & B2_NEIGHMASK;
      const vec4 xj = x[j];
      const flt_t *cij = ci + type[j] * C_N;
      const flt_t delx = xi.x - xj.x;
      const flt_t dely = xi.y - xj.y;
      const flt_t delz = xi.z - xj.z;
      const flt_t rsq = delx * delx + dely * dely + delz * delz;
      if (rsq < cij[C_CUTSQ]) {
        const flt_t rinv = m_rsqrt(rsq);
        const flt_t r = rsq * rinv;
        const flt_t r2inv = rinv * rinv;
        flt_t forcecoul = (flt_t)0, forcebuck = (flt_t)0, evdwl = (flt_t)0, ecoul = (flt_t)0;

        if (STYLE == B200MD_PAIR_BUCK_COUL_CUT) {
          if (rsq < cij[C_CUT_COULSQ]) {
            forcecoul = pc.qqrd2e * qtmp * xj.w * rinv;
            if (sbindex) forcecoul *= pc.special_coul[sbindex];
            if (EVFLAG) ecoul = forcecoul;
          }
        }
        if (STYLE == B200MD_PAIR_BUCK_COUL_LONG || (STYLE == B200MD_PAIR_BUCK_LONG_COUL_LONG && pc.order1)) {
          if (!pc.coultable || rsq <= pc.tabinnersq) {
            const flt_t A1 = (flt_t)0.254829592, A2 = (flt_t)-0.284496736, A3 = (flt_t)1.421413741;
            const flt_t A4 = (flt_t)-1.453152027, A5 = (flt_t)1.061405429;
            const flt_t EWALD_F = (flt_t)1.12837917, EWALD_P = (flt_t)0.3275911;
            const flt_t grij = pc.g_ewald * r;
            const flt_t expm2 = m_exp(-grij * grij);
            const flt_t t = m_rcp((flt_t)1.0 + EWALD_P * grij);
            const flt_t erfc = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * expm2;
            const flt_t prefactor = pc.qqrd2e * qtmp * xj.w * rinv;
            forcecoul = prefactor * (erfc + EWALD_F * grij * expm2);
            if (EVFLAG) ecoul = prefactor * erfc;
            if (sbindex) {
              const flt_t adjust = ((flt_t)1.0 - pc.special_coul[sbindex]) * prefactor;
              forcecoul -= adjust;
              if (EVFLAG) ecoul -= adjust;
            }
          } else {
            const float rsq_lookup = (float)rsq;
            const int itable = (__float_as_int(rsq_lookup) & pc.ncoulmask) >> pc.ncoulshiftbits;
            const flt_t *tb = ctab + 8 * itable;  // {r,dr,f,df,e,de,c,dc}
            const flt_t fraction = ((flt_t)rsq_lookup - tb[0]) * tb[1];
            const flt_t qiqj = qtmp * xj.w;
            forcecoul = qiqj * (tb[2] + fraction * tb[3]);
            if (EVFLAG) ecoul = qiqj * (tb[4] + fraction * tb[5]);
            if (sbindex) {
              const flt_t prefactor = qiqj * (tb[6] + fraction * tb[7]);
              const flt_t adjust = ((flt_t)1.0 - pc.special_coul[sbindex]) * prefactor;
              forcecoul -= adjust;
              if (EVFLAG) ecoul -= adjust;
            }
          }
        }

        if (rsq < cij[C_CUT_LJSQ]) {
          const flt_t r6inv = r2inv * r2inv * r2inv;
          const flt_t rexp = m_exp(-r * cij[C_RHOINV]);
          if (STYLE == B200MD_PAIR_BUCK_LONG_COUL_LONG && pc.order6) {
            if (!pc.disptable || rsq <= pc.tabinnerdispsq) {
              const flt_t grij2 = pc.g2 * rsq;
              const flt_t a2 = m_rcp(grij2);
              const flt_t x2 = a2 * m_exp(-grij2) * cij[C_C];
              forcebuck = r * rexp * cij[C_BUCK1] -
                          pc.g8 * x2 * rsq * ((((flt_t)6.0 * a2 + (flt_t)6.0) * a2 + (flt_t)3.0) * a2 + (flt_t)1.0);
              if (EVFLAG) evdwl = rexp * cij[C_A] - pc.g6 * x2 * ((a2 + (flt_t)1.0) * a2 + (flt_t)0.5);
            } else {
              const float rsq_lookup = (float)rsq;
              const int itable = (__float_as_int(rsq_lookup) & pc.ndispmask) >> pc.ndispshiftbits;
              const flt_t *tb = dtab + 6 * itable;  // {r,dr,f,df,e,de}
              const flt_t fd = (rsq - tb[0]) * tb[1];
              forcebuck = r * rexp * cij[C_BUCK1] - (tb[2] + fd * tb[3]) * cij[C_C];
              if (EVFLAG) evdwl = rexp * cij[C_A] - (tb[4] + fd * tb[5]) * cij[C_C];
            }
            if (sbindex) {
              const flt_t t = pc.special_lj[sbindex] - (flt_t)1.0;
              forcebuck += t * r * rexp * cij[C_BUCK1] - t * r6inv * cij[C_BUCK2];
              if (EVFLAG) evdwl += t * rexp * cij[C_A] - t * r6inv * cij[C_C];
            }
          } else {
            forcebuck = r * rexp * cij[C_BUCK1] - r6inv * cij[C_BUCK2];
            if (EVFLAG) evdwl = rexp * cij[C_A] - r6inv * cij[C_C] - cij[C_OFFSET];
            if (sbindex) {
              const flt_t factor_lj = pc.special_lj[sbindex];
              forcebuck *= factor_lj;
              if (EVFLAG) evdwl *= factor_lj;
            }
          }
        }

        const flt_t fpair = (forcecoul + forcebuck) * r2inv;
        const double dfx = (double)(delx * fpair), dfy = (double)(dely * fpair), dfz = (double)(delz * fpair);
        fx += dfx;
        fy += dfy;
        fz += dfz;
        if (EVFLAG) {
          sevdwl += 0.5 * (double)evdwl;
          secoul += 0.5 * (double)ecoul;
          const flt_t hf = (flt_t)0.5 * fpair;  // IP_PRE_ev_tally_nbor with ev_pre = 1/2
          sv0 += (double)(hf * delx * delx);
          sv1 += (double)(hf * dely * dely);
          sv2 += (double)(hf * delz * delz);
          sv3 += (double)(hf * delx * dely);
          sv4 += (double)(hf * delx * delz);
          sv5 += (double)(hf * dely * delz);
        }
      }
    }
  }

  // fixed-shape xor tree over the TPA lanes of an atom
#pragma unroll
  for (int d = TPA >> 1; d > 0; d >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, d);
    fy += __shfl_xor_sync(0xffffffffu, fy, d);
    fz += __shfl_xor_sync(0xffffffffu, fz, d);
    if (EVFLAG) {
      sevdwl += __shfl_xor_sync(0xffffffffu, sevdwl, d);
      secoul += __shfl_xor_sync(0xffffffffu, secoul, d);
    }
  }
  if (active && sub == 0) f[i] = make_double4(fx, fy, fz, EVFLAG ? sevdwl + secoul : 0.0);

  if (EVFLAG) {
    // block tally: each atom's lane 0 carries its energies; virial terms are still per lane
    double vals[8] = {(active && sub == 0) ? sevdwl : 0.0, (active && sub == 0) ? secoul : 0.0,
                      sv0, sv1, sv2, sv3, sv4, sv5};
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], d);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < 8; k++) s_ev[warp][k] = vals[k];
    __syncthreads();
    if (threadIdx.x < 8) {
      double s = 0.0;
      const int nw = blockDim.x >> 5;
      for (int w = 0; w < nw; w++) s += s_ev[w][threadIdx.x];
      ev_partial[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
    }
  }
}

// single block, fixed order: thread t sums partial rows t, t+256, ... then a shared-memory tree
__global__ void __launch_bounds__(256) k_ev_reduce(int nrows, const double *__restrict__ partial, double *__restrict__ out) {
  __shared__ double s[256];
  for (int k = 0; k < 8; k++) {
    double a = 0.0;
    for (int r = threadIdx.x; r < nrows; r += 256) a += partial[(size_t)r * 8 + k];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
      if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = s[0];
    __syncthreads();
  }
}

template <class flt_t>
PairConsts<flt_t> make_consts(const PairState &ps) {
  PairConsts<flt_t> pc;
  const b200md_pair_params &p = ps.p;
  pc.tp1 = ps.tp1;
  pc.qqrd2e = (flt_t)0;  // filled by caller
  pc.g_ewald = (flt_t)p.g_ewald;
  pc.tabinnersq = (flt_t)p.tabinnersq;
  pc.tabinnerdispsq = (flt_t)p.tabinnerdispsq;
  const flt_t g2 = (flt_t)(p.g_ewald_6 * p.g_ewald_6);
  pc.g2 = g2;
  pc.g6 = g2 * g2 * g2;
  pc.g8 = pc.g6 * g2;
  for (int k = 0; k < 4; k++) {
    pc.special_lj[k] = (flt_t)p.special_lj[k];
    pc.special_coul[k] = (flt_t)p.special_coul[k];
  }
  pc.special_lj[0] = pc.special_coul[0] = (flt_t)1.0;  // pair_buck_intel.cpp:414-417
  pc.ncoulmask = p.ncoulmask;
  pc.ncoulshiftbits = p.ncoulshiftbits;
  pc.ndispmask = p.ndispmask;
  pc.ndispshiftbits = p.ndispshiftbits;
  pc.order1 = (p.ewald_order >> 1) & 1;
  pc.order6 = (p.ewald_order >> 6) & 1;
  pc.coultable = p.ncoultablebits != 0;
  pc.disptable = p.ndisptablebits != 0;
  return pc;
}

int pick_tpa(const b200md_ctx *ctx, int nlocal, long long total_entries) {
  // enough lanes per atom to coalesce the CSR row reads, fewer when the rows are short
  const double avg = nlocal > 0 ? (double)total_entries / nlocal : 0.0;
  if (avg >= 256.0) return 8;
  if (avg >= 48.0) return 8;
  if (avg >= 16.0) return 4;
  (void)ctx;
  return 4;
}

template <int STYLE, class flt_t, int EVFLAG>
int launch_tpa(b200md_ctx *ctx, const PairView &v, int tpa, const PairConsts<flt_t> &pc, const flt_t *coeff,
               const flt_t *ctab, const flt_t *dtab, double *ev_partial, int nblocks) {
  typedef typename V4<flt_t>::type vec4;
#define LAUNCH(T)                                                                                          \
  k_pair<STYLE, flt_t, EVFLAG, T><<<nblocks, 256, 0, ctx->stream>>>(                                        \
      v.nlocal, (const vec4 *)v.x, v.type, v.numneigh, v.offsets, v.entries, pc, coeff, ctab, dtab, v.f, \
      ev_partial)
  switch (tpa) {
    case 4: LAUNCH(4); break;
    case 8: LAUNCH(8); break;
    case 16: LAUNCH(16); break;
    default: LAUNCH(32); break;
  }
#undef LAUNCH
  KERNEL_OK(ctx, "k_pair");
  return 0;
}

template <class flt_t>
int launch_pair(b200md_ctx *ctx, const PairView &v, long long total_entries, int evflag, double *ev_dev) {
  PairState &ps = ctx->pair;
  PairConsts<flt_t> pc = make_consts<flt_t>(ps);
  pc.qqrd2e = (flt_t)ctx->qqrd2e;
  const flt_t *coeff, *ctab, *dtab;
  if (sizeof(flt_t) == 8) {
    coeff = (const flt_t *)ps.coeff_d.p; ctab = (const flt_t *)ps.ctab_d.p; dtab = (const flt_t *)ps.dtab_d.p;
  } else {
    coeff = (const flt_t *)ps.coeff_f.p; ctab = (const flt_t *)ps.ctab_f.p; dtab = (const flt_t *)ps.dtab_f.p;
  }
  const int tpa = pick_tpa(ctx, v.nlocal, total_entries);
  const int nblocks = cdiv((long)v.nlocal * tpa, 256);
  if (nblocks == 0) {
    if (evflag) CUDA_OK(ctx, cudaMemsetAsync(ev_dev, 0, 8 * sizeof(double), ctx->stream));
    return 0;
  }
  if (evflag) RESERVE(ctx, ctx->ev_partial, (size_t)nblocks * 8);
  double *evp = ctx->ev_partial.p;
#define STYLE_CASE(S)                                                                              \
  case S:                                                                                          \
    if (evflag) TRY((launch_tpa<S, flt_t, 1>(ctx, v, tpa, pc, coeff, ctab, dtab, evp, nblocks)));  \
    else TRY((launch_tpa<S, flt_t, 0>(ctx, v, tpa, pc, coeff, ctab, dtab, evp, nblocks)));         \
    break;
  switch (ps.p.style) {
    STYLE_CASE(B200MD_PAIR_BUCK)
    STYLE_CASE(B200MD_PAIR_BUCK_COUL_CUT)
    STYLE_CASE(B200MD_PAIR_BUCK_COUL_LONG)
    STYLE_CASE(B200MD_PAIR_BUCK_LONG_COUL_LONG)
    default: return b2_fail(ctx, B200MD_EINVAL, "unknown pair style %d", ps.p.style);
  }
#undef STYLE_CASE
  if (evflag) {
    k_ev_reduce<<<1, 256, 0, ctx->stream>>>(nblocks, evp, ev_dev);
    KERNEL_OK(ctx, "k_ev_reduce");
  }
  return 0;
}

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
  if (ctx->prec == B200MD_PREC_MIXED) TRY(launch_pair<float>(ctx, v, ctx->neigh.total_entries, evflag, ctx->ev_out.p));
  else TRY(launch_pair<double>(ctx, v, ctx->neigh.total_entries, evflag, ctx->ev_out.p));
  if (evflag) TRY(finish_ev(ctx, eflag, vflag, ev));
  return 0;
}

extern "C" {

int b200md_pair_setup(b200md_ctx *ctx, const b200md_pair_params *p) {
  if (!ctx || !p) return b2_fail(ctx, B200MD_EINVAL, "b200md_pair_setup: NULL argument");
  cudaSetDevice(ctx->device);
  if (p->style < B200MD_PAIR_BUCK || p->style > B200MD_PAIR_BUCK_LONG_COUL_LONG)
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
      cutmax = std::max(cutmax, std::sqrt(p->cutsq[ij]));
    }
  ps.cutmax = cutmax;
  RESERVE(ctx, ps.coeff_d, hd.size());
  RESERVE(ctx, ps.coeff_f, hf.size());
  CUDA_OK(ctx, cudaMemcpy(ps.coeff_d.p, hd.data(), hd.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(ctx, cudaMemcpy(ps.coeff_f.p, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice));
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
  long long total = 0;
  for (int i = 0; i < nlocal; i++) total = std::max(total, (long long)cnumneigh[i] + numneigh[i]);
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
  v.type = dtype.p; v.numneigh = dnum.p; v.offsets = doff.p; v.entries = dent.p; v.f = df.p;
  rc = mixed ? launch_pair<float>(ctx, v, total, evflag, ctx->ev_out.p)
             : launch_pair<double>(ctx, v, total, evflag, ctx->ev_out.p);
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
