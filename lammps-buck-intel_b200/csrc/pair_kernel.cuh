// pair_kernel.cuh — templates of the Buckingham pair kernels; instantiated for <double> in pair.cu and for
// <float> (mixed mode) in pair_mixed.cu.  pair_mixed.cu is compiled with -fmad=false: in float the nearest-
// neighbour terms of a pair (r*rexp*buck1 - r6inv*buck2 + Coulomb) cancel to ~10% of their size and the pair
// forces of a tetrahedron cancel again, so a fused-vs-unfused ulp shows up at 1e-5 of the net force; without
// contraction the float pipeline performs exactly the IEEE operations of the reference's AVX build.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "internal.h"

namespace pairk {

enum { C_CUTSQ = 0, C_CUT_LJSQ, C_CUT_COULSQ, C_BUCK1, C_BUCK2, C_RHOINV, C_A, C_C, C_OFFSET, C_N };

enum { MC_L = 0, MC_MAGIC, MC_LN2, MC_E7, MC_E6, MC_E5, MC_E4, MC_E3, MC_HALF, MC_ONE, MC_R375, MC_EWP, MC_A1, MC_A2,
       MC_A3, MC_A4, MC_A5, MC_EWF, MC_N };

template <class flt_t>
struct PairConsts {
  int tp1;
  flt_t qqrd2e, g_ewald, tabinnersq, tabinnerdispsq, g2, g6, g8;
  flt_t special_lj[4], special_coul[4];
  int ncoulmask, ncoulshiftbits, ndispmask, ndispshiftbits;
  int order1, order6, coultable, disptable;
  int same_cut;     // cut_ljsq == cutsq for every type pair: the second cut-off test is skipped
  flt_t mc[MC_N];   // numeric constants of the double-precision math kernels (see fill_math_consts)
};

template <class flt_t> struct V4;
template <> struct V4<double> { typedef double4 type; };
template <> struct V4<float> { typedef float4 type; };

// ---- double-precision math of the hot loop ---------------------------------------------------------------------
// The loop is bound by the FP64 pipe and by issue slots (ncu: profiles/r01_*), so the libdevice routines (generic
// range checks, ~17 DFMA per exp, a correctly-rounded divide) are replaced by kernels sized for this loop's argument
// ranges and error budget: exp is accurate to 4e-14 relative, rsqrt / rcp to a few ulp, against a parity bar of 1e-9 on
// forces (measured: max force error vs the oracle 6e-13, __graft_entry__.smoke).
//   exp:   x = (n/64) ln2 + r, |r| <= ln2/128 (one FMA: n * fl(ln2/64) is exact inside it);
//          exp(x) = 2^(n>>6) * T[n&63] * (1 + expm1(r)), T = 2^(k/64) in shared memory, expm1 by a degree-4 Taylor
//          polynomial (remainder r^5/120 < 4e-14).  8 FP64 ops, no branches.
//   rsqrt: MUFU.RSQ64H seed (~2^-20) + one third-order step (error e^3).  rcp likewise.
#ifndef B2_EXP_TAB
#define B2_EXP_TAB 64
#endif
#if B2_EXP_TAB == 64
#define B2_EXP_SHIFT 6
#elif B2_EXP_TAB == 32
#define B2_EXP_SHIFT 5
#else
#define B2_EXP_SHIFT 4
#endif
// The numeric constants of these kernels travel in the kernel-parameter block (PairConsts::mc, constant bank 0), not
// as literals: a double literal costs two UMOV/IMAD.MOV issue slots every time it is used (the compiler re-
// materialises it inside the loop — 45 of the 223 instructions per pair in the first profile), a parameter is loaded
// once into a (uniform) register ahead of the loop.

static inline void fill_math_consts(double *mc) {
#if B2_EXP_TAB == 64
  mc[MC_L] = 92.332482616893656877;        // 64 / ln 2
  mc[MC_LN2] = 1.0830424696249145e-02;     // ln2/64 to full double precision (the reduction is a single FMA)
#elif B2_EXP_TAB == 32
  mc[MC_L] = 46.16624130844683;
  mc[MC_LN2] = 0.02166084939249829;        // ln2/32
#else
  mc[MC_L] = 23.083120654223414;
  mc[MC_LN2] = 0.04332169878499658;        // ln2/16
#endif
  mc[MC_MAGIC] = 6755399441055744.0;       // 1.5 * 2^52: the low word of (t + MAGIC) is rint(t)
  mc[MC_E7] = 1.0 / 5040.0; mc[MC_E6] = 1.0 / 720.0;
  mc[MC_E5] = 1.0 / 120.0; mc[MC_E4] = 1.0 / 24.0; mc[MC_E3] = 1.0 / 6.0;
  mc[MC_HALF] = 0.5; mc[MC_ONE] = 1.0; mc[MC_R375] = 0.375;
  mc[MC_EWP] = 0.3275911;
  mc[MC_A1] = 0.254829592; mc[MC_A2] = -0.284496736; mc[MC_A3] = 1.421413741; mc[MC_A4] = -1.453152027;
  mc[MC_A5] = 1.061405429;
  mc[MC_EWF] = 1.12837917;
}
static inline void fill_math_consts(float *) {}

__device__ __forceinline__ double fast_exp(const double x, const double *__restrict__ s_tab, const double *mc) {
  const double t = fma(x, mc[MC_L], mc[MC_MAGIC]);
  const int n = __double2loint(t);
  const double nf = t - mc[MC_MAGIC];
  // one FMA: nf * (ln2/64) is exact inside it, and |nf| * |fl(ln2/64) - ln2/64| < 4e-15 for |x| < 700
  const double r = fma(nf, -mc[MC_LN2], x);
#if B2_EXP_TAB == 64
  // |r| <= ln2/128: the degree-4 remainder r^5/120 is < 4e-14 relative — four orders inside the 1e-9 parity bar
  double p = mc[MC_E4];
#elif B2_EXP_TAB == 32
  double p = fma(r, mc[MC_E6], mc[MC_E5]);
  p = fma(p, r, mc[MC_E4]);
#else
  double p = fma(r, mc[MC_E7], mc[MC_E6]);
  p = fma(p, r, mc[MC_E5]);
  p = fma(p, r, mc[MC_E4]);
#endif
  p = fma(p, r, mc[MC_E3]);
  p = fma(p, r, mc[MC_HALF]);
  p = fma(p, r * r, r);
  const double T = s_tab[n & (B2_EXP_TAB - 1)];
  const double v = fma(T, p, T);
  return __hiloint2double(__double2hiint(v) + ((n >> B2_EXP_SHIFT) << 20), __double2loint(v));
}
__device__ __forceinline__ double fast_rsqrt(const double x, const double *mc) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x, -(y * y), mc[MC_ONE]);
  const double c = fma(e, mc[MC_R375], mc[MC_HALF]);
  return fma(c, y * e, y);
}
__device__ __forceinline__ double fast_rcp(const double x, const double *mc) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, mc[MC_ONE]);
  return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double m_exp(double x, const double *s_tab, const double *mc) { return fast_exp(x, s_tab, mc); }
__device__ __forceinline__ float m_exp(float x, const float *, const float *) { return expf(x); }
__device__ __forceinline__ double m_rcp(double x, const double *mc) { return fast_rcp(x, mc); }
__device__ __forceinline__ float m_rcp(float x, const float *) { return 1.0f / x; }
// rsq with the reference's un-fused rounding ((dx*dx + dy*dy) + dz*dz, AVX build: no FMA) so that the
// cut-off decisions and, in mixed mode, r itself are bit-identical to the CPU path
__device__ __forceinline__ double m_rsq(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ float m_rsq(float dx, float dy, float dz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// fused flavour for the production double kernel (mixed mode keeps the reference's float expression): the list criterion (neigh.cu) keeps the un-fused expression and
// with it the bit-exact pair set; here a 1-ulp difference can only matter for a pair within 1 ulp of the force cut-off
__device__ __forceinline__ double m_rsq_fused(double dx, double dy, double dz) { return fma(dz, dz, fma(dy, dy, dx * dx)); }
__device__ __forceinline__ float m_rsq_fused(float dx, float dy, float dz) { return m_rsq(dx, dy, dz); }
// r, 1/r, 1/r^2.  double: one rsqrt (error ~1 ulp, far inside 1e-9).  float: exp(-r/rho) amplifies an ulp of r
// by r/rho ~ 50, so mixed mode replays the reference's own correctly-rounded sequence: COUL_LONG takes
// r2inv = 1/rsq, r = 1/sqrt(r2inv) (pair_buck_coul_long_intel.cpp:287-288), the others r = sqrt(rsq)
// (pair_buck_intel.cpp:254-255).
template <int COUL_LONG>
__device__ __forceinline__ void m_r(double rsq, double &r, double &rinv, double &r2inv, const double *mc) {
  rinv = fast_rsqrt(rsq, mc);
  r = rsq * rinv;
  r2inv = rinv * rinv;
}
template <int COUL_LONG>
__device__ __forceinline__ void m_r(float rsq, float &r, float &rinv, float &r2inv, const float *) {
  r2inv = __fdiv_rn(1.0f, rsq);
  r = COUL_LONG ? __fdiv_rn(1.0f, __fsqrt_rn(r2inv)) : __fsqrt_rn(rsq);
  rinv = __fdiv_rn(1.0f, r);
}

// real-space Ewald kernel, pair_buck_coul_long_intel.cpp:296-308.  double: fused arithmetic, one true divide.
// float: the reference's own operation order with correctly-rounded ops — the alternating-charge sum cancels to
// ~1% of its terms, so an ulp per pair is already 1e-5 of the net force (the reference's own mixed mode sits
// ~1.7e-5 from its double mode on data.aC); replaying the order keeps the two float paths together.
template <class flt_t>
__device__ __forceinline__ void ewald_real(flt_t g_ewald, flt_t qqrd2e, flt_t qi, flt_t qj, flt_t r, flt_t rinv,
                                           flt_t &grij, flt_t &expm2, flt_t &erfc, flt_t &prefactor,
                                           const flt_t *s_tab, const flt_t *mc);
template <>
__device__ __forceinline__ void ewald_real<double>(double g_ewald, double qqrd2e, double qi, double qj, double r,
                                                   double rinv, double &grij, double &expm2, double &erfc,
                                                   double &prefactor, const double *s_tab, const double *mc) {
  grij = g_ewald * r;
  expm2 = fast_exp(-grij * grij, s_tab, mc);
  const double t = fast_rcp(fma(mc[MC_EWP], grij, mc[MC_ONE]), mc);
  erfc = t * (mc[MC_A1] + t * (mc[MC_A2] + t * (mc[MC_A3] + t * (mc[MC_A4] + t * mc[MC_A5])))) * expm2;
  prefactor = qqrd2e * qi * qj * rinv;
}
template <>
__device__ __forceinline__ void ewald_real<float>(float g_ewald, float qqrd2e, float qi, float qj, float r,
                                                  float rinv, float &grij, float &expm2, float &erfc,
                                                  float &prefactor, const float *, const float *) {
  const float A1 = 0.254829592f, A2 = -0.284496736f, A3 = 1.421413741f, A4 = -1.453152027f, A5 = 1.061405429f;
  const float INV_EWALD_P = (float)(1.0 / 0.3275911);
  (void)rinv;
  grij = __fmul_rn(g_ewald, r);
  expm2 = expf(-__fmul_rn(grij, grij));
  const float t = __fdiv_rn(INV_EWALD_P, __fadd_rn(INV_EWALD_P, grij));
  float p = __fadd_rn(A4, __fmul_rn(t, A5));
  p = __fadd_rn(A3, __fmul_rn(t, p));
  p = __fadd_rn(A2, __fmul_rn(t, p));
  p = __fadd_rn(A1, __fmul_rn(t, p));
  erfc = __fmul_rn(__fmul_rn(t, p), expm2);
  prefactor = __fdiv_rn(__fmul_rn(__fmul_rn(qqrd2e, qi), qj), r);
}

// j = entry & NEIGHMASK, opaque to the optimiser: otherwise the mask is folded into the 64-bit address arithmetic of
// x[j] (shift/mask/add-with-carry, 6 instructions) instead of one LOP3 + one IMAD.WIDE
template <int PACKT>
__device__ __forceinline__ int nbr_index(const int e) {
  int j;
  if (PACKT) asm("and.b32 %0, %1, 0x03FFFFFF;" : "=r"(j) : "r"(e));
  else asm("and.b32 %0, %1, 0x3FFFFFFF;" : "=r"(j) : "r"(e));
  return j;
}
// gather of one atom: a single 256-bit load in double mode (LDG.E.256, sm_100: half the L1 wavefronts of two
// LDG.128 — the L1 data pipe, not HBM, is what the gathers of this kernel load), 128-bit in mixed mode
__device__ __forceinline__ double4 ld_atom(const double4 *p) {
  double4 v;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_atom(const float4 *p) { return __ldg(p); }

struct PairView {  // device pointers of one evaluation
  int nlocal;
  const void *x;         // double4* or float4*
  const int *type;
  const int *numneigh;
  const long long *offsets;
  const int *entries;
  double4 *f;
  int packed_type;       // entries carry type(j) (NeighState::packed_type)
};

#ifndef PAIR_MINB
#define PAIR_MINB 4
#endif

// One pair evaluation: fpair (= F/r), evdwl, ecoul from rsq and the per-type-pair constants `cij`.
// GENERAL = 1: everything — Coulomb / dispersion lookup tables (INTEL_ALLOW_TABLE), special-bond scaling — written
//   with the reference's branch structure.
// GENERAL = 0: the production flavour for lists built on the device (atomic systems: no special bits) with analytic
//   kernels.  It has NO data-dependent branch: the three cut-off tests become selects, so that the two pairs of the
//   2x-unrolled loop form one basic block and their FP64 dependency chains (exp -> polynomial -> ...) interleave.
//   In-range lanes execute exactly the same operations in both flavours (results are bit-identical).
template <int STYLE, class flt_t, int EVFLAG, int GENERAL>
__device__ __forceinline__ void pair_eval(const PairConsts<flt_t> &pc, const flt_t *__restrict__ cij,
                                          const flt_t *__restrict__ ctab, const flt_t *__restrict__ dtab,
                                          const flt_t *__restrict__ s_tab, const flt_t rsq, const flt_t qtmp,
                                          const flt_t qj, const int sbindex, flt_t &fpair, flt_t &evdwl, flt_t &ecoul) {
  const bool in_cut = rsq < cij[C_CUTSQ];
  fpair = (flt_t)0; evdwl = (flt_t)0; ecoul = (flt_t)0;
  if (GENERAL && !in_cut) return;
  flt_t r, rinv, r2inv;
  m_r<STYLE == B200MD_PAIR_BUCK_COUL_LONG>(rsq, r, rinv, r2inv, pc.mc);
  flt_t forcecoul = (flt_t)0, forcebuck = (flt_t)0;

  if (STYLE == B200MD_PAIR_BUCK_COUL_CUT) {
    const bool in_coul = rsq < cij[C_CUT_COULSQ];
    if (!GENERAL || in_coul) {
      forcecoul = pc.qqrd2e * qtmp * qj * rinv;
      if (GENERAL && sbindex) forcecoul *= pc.special_coul[sbindex];
      if (!GENERAL) forcecoul = in_coul ? forcecoul : (flt_t)0;
      if (EVFLAG) ecoul = forcecoul;
    }
  }
  constexpr bool LCL = STYLE == B200MD_PAIR_BUCK_LONG_COUL_LONG || STYLE == B200MD_PAIR_LJ_LONG_COUL_LONG;
  if (STYLE == B200MD_PAIR_BUCK_COUL_LONG || (LCL && pc.order1)) {
    if (!GENERAL || !pc.coultable || rsq <= pc.tabinnersq) {
      const flt_t EWALD_F = sizeof(flt_t) == 8 ? pc.mc[MC_EWF] : (flt_t)1.12837917;
      flt_t erfc, expm2, grij, prefactor;
      ewald_real<flt_t>(pc.g_ewald, pc.qqrd2e, qtmp, qj, r, rinv, grij, expm2, erfc, prefactor, s_tab, pc.mc);
      forcecoul = prefactor * (erfc + EWALD_F * grij * expm2);
      if (EVFLAG) ecoul = prefactor * erfc;
      if (GENERAL && sbindex) {
        const flt_t adjust = ((flt_t)1.0 - pc.special_coul[sbindex]) * prefactor;
        forcecoul -= adjust;
        if (EVFLAG) ecoul -= adjust;
      }
    } else {
      const float rsq_lookup = (float)rsq;
      const int itable = (__float_as_int(rsq_lookup) & pc.ncoulmask) >> pc.ncoulshiftbits;
      const flt_t *tb = ctab + 8 * itable;  // {r,dr,f,df,e,de,c,dc}
      // {r,dr,f,df} in one vector load (256-bit in double mode): the table gathers load the L1 data pipe
      const typename V4<flt_t>::type t4 = ld_atom(reinterpret_cast<const typename V4<flt_t>::type *>(tb));
      const flt_t fraction = ((flt_t)rsq_lookup - t4.x) * t4.y;
      const flt_t qiqj = qtmp * qj;
      forcecoul = qiqj * (t4.z + fraction * t4.w);
      if (EVFLAG) ecoul = qiqj * (tb[4] + fraction * tb[5]);
      if (sbindex) {
        const flt_t prefactor = qiqj * (tb[6] + fraction * tb[7]);
        const flt_t adjust = ((flt_t)1.0 - pc.special_coul[sbindex]) * prefactor;
        forcecoul -= adjust;
        if (EVFLAG) ecoul -= adjust;
      }
    }
  }

  const bool in_lj = pc.same_cut ? true : rsq < cij[C_CUT_LJSQ];
  if (STYLE == B200MD_PAIR_LJ_LONG_COUL_LONG) {
    // pair_lj_long_coul_long_intel.cpp:620-688; C_BUCK1, C_BUCK2, C_A, C_C hold lj1, lj2, lj3, lj4
    if (!GENERAL || in_lj) {
      const flt_t r6inv = r2inv * r2inv * r2inv;
      const flt_t lj1 = cij[C_BUCK1], lj2 = cij[C_BUCK2], lj3 = cij[C_A], lj4 = cij[C_C];
      if (pc.order6) {
        if (!GENERAL || !pc.disptable || rsq <= pc.tabinnerdispsq) {
          const flt_t grij2 = pc.g2 * rsq;
          const flt_t a2 = m_rcp(grij2, pc.mc);
          const flt_t x2 = a2 * m_exp(-grij2, s_tab, pc.mc) * lj4;
          forcebuck = r6inv * r6inv * lj1 -
                      pc.g8 * x2 * rsq * ((((flt_t)6.0 * a2 + (flt_t)6.0) * a2 + (flt_t)3.0) * a2 + (flt_t)1.0);
          if (EVFLAG) evdwl = r6inv * r6inv * lj3 - pc.g6 * x2 * ((a2 + (flt_t)1.0) * a2 + (flt_t)0.5);
        } else {
          const float rsq_lookup = (float)rsq;
          const int itable = (__float_as_int(rsq_lookup) & pc.ndispmask) >> pc.ndispshiftbits;
          const flt_t *tb = dtab + 6 * itable;  // {r,dr,f,df,e,de}
          const flt_t fd = (rsq - tb[0]) * tb[1];
          forcebuck = r6inv * r6inv * lj1 - (tb[2] + fd * tb[3]) * lj4;
          if (EVFLAG) evdwl = r6inv * r6inv * lj3 - (tb[4] + fd * tb[5]) * lj4;
        }
        if (GENERAL && sbindex) {
          const flt_t t = r6inv * ((flt_t)1.0 - pc.special_lj[sbindex]);
          forcebuck += t * (lj2 - r6inv * lj1);
          if (EVFLAG) evdwl += t * (lj4 - r6inv * lj3);
        }
      } else {
        forcebuck = r6inv * (r6inv * lj1 - lj2);
        if (EVFLAG) evdwl = r6inv * (r6inv * lj3 - lj4) - cij[C_OFFSET];
        if (GENERAL && sbindex) {
          const flt_t factor_lj = pc.special_lj[sbindex];
          forcebuck *= factor_lj;
          if (EVFLAG) evdwl *= factor_lj;
        }
      }
      if (!GENERAL) {
        forcebuck = in_lj ? forcebuck : (flt_t)0;
        if (EVFLAG) evdwl = in_lj ? evdwl : (flt_t)0;
      }
    }
  } else if (!GENERAL || in_lj) {
    const flt_t r6inv = r2inv * r2inv * r2inv;
    const flt_t rexp = m_exp(-r * cij[C_RHOINV], s_tab, pc.mc);
    if (STYLE == B200MD_PAIR_BUCK_LONG_COUL_LONG && pc.order6) {
      if (!GENERAL || !pc.disptable || rsq <= pc.tabinnerdispsq) {
        const flt_t grij2 = pc.g2 * rsq;
        const flt_t a2 = m_rcp(grij2, pc.mc);
        const flt_t x2 = a2 * m_exp(-grij2, s_tab, pc.mc) * cij[C_C];
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
      if (GENERAL && sbindex) {
        const flt_t t = pc.special_lj[sbindex] - (flt_t)1.0;
        forcebuck += t * r * rexp * cij[C_BUCK1] - t * r6inv * cij[C_BUCK2];
        if (EVFLAG) evdwl += t * rexp * cij[C_A] - t * r6inv * cij[C_C];
      }
    } else {
      forcebuck = r * rexp * cij[C_BUCK1] - r6inv * cij[C_BUCK2];
      if (EVFLAG) evdwl = rexp * cij[C_A] - r6inv * cij[C_C] - cij[C_OFFSET];
      if (GENERAL && sbindex) {
        const flt_t factor_lj = pc.special_lj[sbindex];
        forcebuck *= factor_lj;
        if (EVFLAG) evdwl *= factor_lj;
      }
    }
    if (!GENERAL) {
      forcebuck = in_lj ? forcebuck : (flt_t)0;
      if (EVFLAG) evdwl = in_lj ? evdwl : (flt_t)0;
    }
  }

  fpair = (forcecoul + forcebuck) * r2inv;
  if (!GENERAL) {
    fpair = in_cut ? fpair : (flt_t)0;
    if (EVFLAG) { evdwl = in_cut ? evdwl : (flt_t)0; ecoul = in_cut ? ecoul : (flt_t)0; }
  }
}

// PACKT = 1: the entry carries type(j) in bits 26..29 (lists built on the device, internal.h) — no type[j] gather.
template <int STYLE, class flt_t, int EVFLAG, int TPA, int GENERAL, int PACKT>
__global__ void __launch_bounds__(256, PAIR_MINB)
k_pair(const int nlocal, const typename V4<flt_t>::type *__restrict__ x, const int *__restrict__ type,
       const int *__restrict__ numneigh, const long long *__restrict__ offsets,
       const int *__restrict__ entries, const PairConsts<flt_t> pc, const flt_t *__restrict__ coeff,
       const flt_t *__restrict__ ctab, const flt_t *__restrict__ dtab, const flt_t *__restrict__ exptab,
       double4 *__restrict__ f, double *__restrict__ ev_partial) {
  typedef typename V4<flt_t>::type vec4;
  __shared__ flt_t s_coeff[(B2_MAXTYPES + 1) * (B2_MAXTYPES + 1) * C_N];
  __shared__ flt_t s_tab[B2_EXP_TAB];   // 2^(k/64) for fast_exp (double instantiation only)
  __shared__ double s_ev[8][8];
  for (int k = threadIdx.x; k < pc.tp1 * pc.tp1 * C_N; k += blockDim.x) s_coeff[k] = coeff[k];
  if (sizeof(flt_t) == 8 && threadIdx.x < B2_EXP_TAB) s_tab[threadIdx.x] = exptab[threadIdx.x];
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

    // software pipeline over the two dependent loads of an iteration (list entry -> gathered atom): the entry is
    // fetched two iterations ahead, the atom one ahead, so their latencies overlap the ~65 FP64 operations of a pair
    int e_cur = 0, e_nxt = 0;
    if (sub < jnum) e_cur = jlist[sub];
    if (sub + TPA < jnum) e_nxt = jlist[sub + TPA];
    int j_nxt = nbr_index<PACKT>(e_cur);
    vec4 xj_nxt = ld_atom(x + j_nxt);
    int tj_nxt = PACKT ? 0 : type[j_nxt];

#pragma unroll 2
    for (int jj = sub; jj < jnum; jj += TPA) {
      const int e = e_cur;
      const vec4 xj = xj_nxt;
      const int tj = PACKT ? ((e >> B2_TYPESHIFT) & 15) : tj_nxt;
      e_cur = e_nxt;   // past the row's end this stays a valid (already used) entry: the loads below are harmless
      if (jj + 2 * TPA < jnum) e_nxt = jlist[jj + 2 * TPA];
      j_nxt = nbr_index<PACKT>(e_cur);
      xj_nxt = ld_atom(x + j_nxt);
      if (!PACKT) tj_nxt = type[j_nxt];
      const int sbindex = GENERAL ? (e >> B2_SBBITS) & 3 : 0;
      const flt_t *cij = ci + tj * C_N;
      const flt_t delx = xi.x - xj.x;
      const flt_t dely = xi.y - xj.y;
      const flt_t delz = xi.z - xj.z;
      const flt_t rsq = GENERAL ? m_rsq(delx, dely, delz) : m_rsq_fused(delx, dely, delz);
      flt_t fpair, evdwl, ecoul;
      pair_eval<STYLE, flt_t, EVFLAG, GENERAL>(pc, cij, ctab, dtab, s_tab, rsq, qtmp, xj.w, sbindex, fpair, evdwl, ecoul);
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
static __global__ void __launch_bounds__(256) k_ev_reduce(int nrows, const double *__restrict__ partial, double *__restrict__ out) {
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
  fill_math_consts(pc.mc);
  pc.same_cut = ps.same_cut ? 1 : 0;
  return pc;
}

static inline int pick_tpa(const b200md_ctx *ctx, int nlocal, long long total_entries) {
  // enough lanes per atom to coalesce the CSR row reads, fewer when the rows are short
  const double avg = nlocal > 0 ? (double)total_entries / nlocal : 0.0;
  (void)ctx;
  return avg >= 48.0 ? 8 : 4;
}

static inline int pair_threads() {   // tuning knob (power of two, <= 256)
  static const int t = getenv("B200MD_PAIR_THREADS") ? atoi(getenv("B200MD_PAIR_THREADS")) : 256;
  return t;
}

template <int STYLE, class flt_t, int EVFLAG>
int launch_tpa(b200md_ctx *ctx, const PairView &v, int tpa, int variant, const PairConsts<flt_t> &pc,
               const flt_t *coeff, const flt_t *ctab, const flt_t *dtab, const flt_t *exptab, double *ev_partial,
               int nblocks) {
  typedef typename V4<flt_t>::type vec4;
#define LAUNCH(T, G, P)                                                                                    \
  k_pair<STYLE, flt_t, EVFLAG, T, G, P><<<nblocks, pair_threads(), 0, ctx->stream>>>(                       \
      v.nlocal, (const vec4 *)v.x, v.type, v.numneigh, v.offsets, v.entries, pc, coeff, ctab, dtab, exptab, \
      v.f, ev_partial)
#define LAUNCH_T(T)                                \
  do {                                             \
    if (variant == 0) LAUNCH(T, 0, 1);             \
    else if (variant == 1) LAUNCH(T, 1, 1);        \
    else LAUNCH(T, 1, 0);                          \
  } while (0)
  if (tpa == 4) LAUNCH_T(4);
  else LAUNCH_T(8);
#undef LAUNCH_T
#undef LAUNCH
  KERNEL_OK(ctx, "k_pair");
  return 0;
}

// has_special: the list may carry special-bond bits (host-supplied lists); lists built on the device never do
template <class flt_t>
int launch_pair(b200md_ctx *ctx, const PairView &v, long long total_entries, int evflag, double *ev_dev,
                int has_special) {
  PairState &ps = ctx->pair;
  PairConsts<flt_t> pc = make_consts<flt_t>(ps);
  pc.qqrd2e = (flt_t)ctx->qqrd2e;
  const flt_t *coeff, *ctab, *dtab, *exptab = nullptr;
  if (sizeof(flt_t) == 8) {
    coeff = (const flt_t *)ps.coeff_d.p; ctab = (const flt_t *)ps.ctab_d.p; dtab = (const flt_t *)ps.dtab_d.p;
    exptab = (const flt_t *)ps.exptab.p;
  } else {
    coeff = (const flt_t *)ps.coeff_f.p; ctab = (const flt_t *)ps.ctab_f.p; dtab = (const flt_t *)ps.dtab_f.p;
  }
  // 0: branch-free analytic, packed type; 1: tables, packed type; 2: everything, type gathered (host lists)
  // (special-bond bits ride in bits 30-31 beside the packed type: the generic flavour reads them)
  const int variant = !v.packed_type ? 2 : ((pc.coultable || pc.disptable || has_special) ? 1 : 0);
  const int tpa = pick_tpa(ctx, v.nlocal, total_entries);
  const int nblocks = cdiv((long)v.nlocal * tpa, pair_threads());
  if (nblocks == 0) {
    if (evflag) CUDA_OK(ctx, cudaMemsetAsync(ev_dev, 0, 8 * sizeof(double), ctx->stream));
    return 0;
  }
  if (evflag) RESERVE(ctx, ctx->ev_partial, (size_t)nblocks * 8);
  double *evp = ctx->ev_partial.p;
#define STYLE_CASE(S)                                                                                            \
  case S:                                                                                                        \
    if (evflag) TRY((launch_tpa<S, flt_t, 1>(ctx, v, tpa, variant, pc, coeff, ctab, dtab, exptab, evp, nblocks))); \
    else TRY((launch_tpa<S, flt_t, 0>(ctx, v, tpa, variant, pc, coeff, ctab, dtab, exptab, evp, nblocks)));        \
    break;
  switch (ps.p.style) {
    STYLE_CASE(B200MD_PAIR_BUCK)
    STYLE_CASE(B200MD_PAIR_BUCK_COUL_CUT)
    STYLE_CASE(B200MD_PAIR_BUCK_COUL_LONG)
    STYLE_CASE(B200MD_PAIR_BUCK_LONG_COUL_LONG)
    STYLE_CASE(B200MD_PAIR_LJ_LONG_COUL_LONG)
    default: return b2_fail(ctx, B200MD_EINVAL, "unknown pair style %d", ps.p.style);
  }
#undef STYLE_CASE
  if (evflag) {
    k_ev_reduce<<<1, 256, 0, ctx->stream>>>(nblocks, evp, ev_dev);
    KERNEL_OK(ctx, "k_ev_reduce");
  }
  return 0;
}

}  // namespace pairk

// pair.cu / pair_mixed.cu
int b2_launch_pair_double(b200md_ctx *ctx, const pairk::PairView &v, long long total_entries, int evflag, double *ev_dev,
                          int has_special);
int b2_launch_pair_float(b200md_ctx *ctx, const pairk::PairView &v, long long total_entries, int evflag, double *ev_dev,
                         int has_special);
