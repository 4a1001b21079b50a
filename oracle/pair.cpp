// oracle/pair.cpp — TEST INFRASTRUCTURE (see oracle.h).  Pinned bit for bit against the reference's own compiled loops
// (oracle/_ref, tests/test_oracle_vs_ref.py).
//
// CPU restatement of the reference's four Buckingham eval<> loops, in the reference's own structure:
// OpenMP static i-range split, thread-private force arrays reduced after the loop, flt_t arithmetic with
// acc_t accumulation.
//   buck                 pair_buck_intel.cpp:215-323        (inner jj loop :241-317)
//   buck/coul/cut        pair_buck_coul_cut_intel.cpp:231-360 (inner :259-353)
//   buck/coul/long       pair_buck_coul_long_intel.cpp:247-411 (inner :275-405; erfc poly :296-307;
//                        table branch :317-340)
//   buck/long/coul/long  pair_buck_long_coul_long_intel.cpp:295-508 (Coulomb :350-409, dispersion :410-473)
// Parameter packing follows pack_force_const of each style (…:391-429, :431-476, :481-542, :573-644) and
// the stock PairBuck*::init_one / Pair::init_tables / init_bitmap.
// Cut-off test: `rsq < cutsq` (the INTEL_VMASK / upstream form).
#include <omp.h>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

constexpr int SBBITS = 30;
constexpr int NEIGHMASK = 0x3FFFFFFF;

template <class flt_t>
struct PackedParams {
  int tp1;
  std::vector<flt_t> cutsq, cut_ljsq, cut_coulsq, buck1, buck2, rhoinv, a, c, offset;
  flt_t special_lj[4], special_coul[4];
  flt_t qqrd2e, g_ewald, tabinnersq, tabinnerdispsq;
  std::vector<flt_t> rtable, drtable, ftable, dftable, etable, detable, ctable, dctable;
  std::vector<flt_t> rdisp, drdisp, fdisp, dfdisp, edisp, dedisp;
  explicit PackedParams(const orc_pair_params &p) {
    tp1 = p.ntypes + 1;
    const int n = tp1 * tp1;
    auto cp = [&](std::vector<flt_t> &dst, const double *src, int m) {
      dst.resize(m);
      for (int i = 0; i < m; i++) dst[i] = src ? (flt_t)src[i] : (flt_t)0;
    };
    cp(cutsq, p.cutsq, n);  cp(cut_ljsq, p.cut_ljsq, n);  cp(cut_coulsq, p.cut_coulsq, n);
    cp(buck1, p.buck1, n);  cp(buck2, p.buck2, n);        cp(rhoinv, p.rhoinv, n);
    cp(a, p.a, n);          cp(c, p.c, n);                cp(offset, p.offset, n);
    for (int i = 0; i < 4; i++) {
      special_lj[i] = (flt_t)p.special_lj[i];
      special_coul[i] = (flt_t)p.special_coul[i];
    }
    special_lj[0] = special_coul[0] = (flt_t)1.0;  // pair_buck_intel.cpp:414-417
    qqrd2e = (flt_t)p.qqrd2e;
    g_ewald = (flt_t)p.g_ewald;
    tabinnersq = (flt_t)p.tabinnersq;
    tabinnerdispsq = (flt_t)p.tabinnerdispsq;
    if (p.ncoultablebits) {
      const int nt = 1 << p.ncoultablebits;
      cp(rtable, p.rtable, nt);  cp(drtable, p.drtable, nt);  cp(ftable, p.ftable, nt);
      cp(dftable, p.dftable, nt); cp(etable, p.etable, nt);   cp(detable, p.detable, nt);
      cp(ctable, p.ctable, nt);  cp(dctable, p.dctable, nt);
    }
    if (p.ndisptablebits) {
      const int nt = 1 << p.ndisptablebits;
      cp(rdisp, p.rdisptable, nt);  cp(drdisp, p.drdisptable, nt);  cp(fdisp, p.fdisptable, nt);
      cp(dfdisp, p.dfdisptable, nt); cp(edisp, p.edisptable, nt);   cp(dedisp, p.dedisptable, nt);
    }
  }
};

inline uint32_t float_bits(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

template <int STYLE, int EVFLAG, int EFLAG, int NEWTON_PAIR, class flt_t, class acc_t>
void eval(const int vflag, const int eatom, const int nlocal, const int nall, const double *xd,
          const int *type, const double *qd, const int *numneigh, const long *cnumneigh,
          const int *firstneigh, const orc_pair_params &pp, double *fout, double *ev_global,
          int nthreads) {
  const PackedParams<flt_t> fc(pp);
  const int ntypes = fc.tp1;
  const int inum = nlocal;
  const int ORDER1 = pp.order1, ORDER6 = pp.order6;
  const int ncoultablebits = pp.ncoultablebits, ncoulmask = pp.ncoulmask,
            ncoulshiftbits = pp.ncoulshiftbits;
  const int ndisptablebits = pp.ndisptablebits, ndispmask = pp.ndispmask,
            ndispshiftbits = pp.ndispshiftbits;

  // IntelBuffers::thr_pack (intel_buffers.h:185-203): double -> flt_t AoS {x,y,z,w=type}, q
  struct atom_t { flt_t x, y, z; int w; };
  struct force_t { acc_t x, y, z, w; };
  std::vector<atom_t> xbuf(nall);
  std::vector<flt_t> qbuf(nall, (flt_t)0);
  for (int i = 0; i < nall; i++) {
    xbuf[i].x = (flt_t)xd[3 * (size_t)i];
    xbuf[i].y = (flt_t)xd[3 * (size_t)i + 1];
    xbuf[i].z = (flt_t)xd[3 * (size_t)i + 2];
    xbuf[i].w = type[i];
    if (qd) qbuf[i] = (flt_t)qd[i];
  }
  const atom_t *const x = xbuf.data();
  const flt_t *const q = qbuf.data();
  const flt_t qqrd2e = fc.qqrd2e;
  const flt_t g_ewald = fc.g_ewald;
  const flt_t tabinnersq = fc.tabinnersq;
  const flt_t g2 = (flt_t)(pp.g_ewald_6 * pp.g_ewald_6), g6 = g2 * g2 * g2, g8 = g6 * g2;

  const int f_stride = nall;
  std::vector<force_t> fthr((size_t)f_stride * nthreads);
  force_t *const f_start = fthr.data();

  acc_t oevdwl = 0, oecoul = 0, ov0 = 0, ov1 = 0, ov2 = 0, ov3 = 0, ov4 = 0, ov5 = 0;

#pragma omp parallel num_threads(nthreads) reduction(+ : oevdwl, oecoul, ov0, ov1, ov2, ov3, ov4, ov5)
  {
    // IP_PRE_omp_range_id
    const int tid = omp_get_thread_num();
    const int idelta = 1 + inum / nthreads;
    const int iifrom = std::min(tid * idelta, inum);
    const int iito = std::min(iifrom + idelta, inum);

    force_t *const f = f_start + (size_t)tid * f_stride;
    std::memset(f, 0, (size_t)f_stride * sizeof(force_t));

    for (int i = iifrom; i < iito; ++i) {
      const int itype = x[i].w;
      const int ptr_off = itype * ntypes;
      const int *const jlist = firstneigh + cnumneigh[i];
      const int jnum = numneigh[i];

      acc_t fxtmp = 0, fytmp = 0, fztmp = 0, fwtmp = 0;
      acc_t sevdwl = 0, secoul = 0, sv0 = 0, sv1 = 0, sv2 = 0, sv3 = 0, sv4 = 0, sv5 = 0;

      const flt_t xtmp = x[i].x, ytmp = x[i].y, ztmp = x[i].z;
      const flt_t qtmp = q[i];

      for (int jj = 0; jj < jnum; jj++) {
        flt_t forcecoul = 0, forcebuck = 0, evdwl = 0, ecoul = 0;
        const int sbindex = jlist[jj] >> SBBITS & 3;
        const int j = jlist[jj] & NEIGHMASK;

        const flt_t delx = xtmp - x[j].x;
        const flt_t dely = ytmp - x[j].y;
        const flt_t delz = ztmp - x[j].z;
        const int jtype = x[j].w;
        const int ij = ptr_off + jtype;
        const flt_t rsq = delx * delx + dely * dely + delz * delz;
        const flt_t r2inv = (flt_t)1.0 / rsq;
        // pair_buck_coul_long_intel.cpp:288 takes r = 1/sqrt(r2inv); the others sqrt(rsq)
        const flt_t r = (STYLE == ORC_BUCK_COUL_LONG) ? (flt_t)1.0 / std::sqrt(r2inv) : std::sqrt(rsq);

        if (!(rsq < fc.cutsq[ij])) continue;

        if (STYLE == ORC_BUCK_COUL_CUT) {
          if (rsq < fc.cut_coulsq[ij]) {  // pair_buck_coul_cut_intel.cpp:275-287
            forcecoul = qqrd2e * qtmp * q[j] / r;
            if (EFLAG) ecoul = forcecoul;
            if (sbindex) {
              const flt_t factor_coul = fc.special_coul[sbindex];
              forcecoul *= factor_coul;
              if (EFLAG) ecoul *= factor_coul;
            }
          }
        }
        constexpr bool LCL = STYLE == ORC_BUCK_LONG_COUL_LONG || STYLE == ORC_LJ_LONG_COUL_LONG;
        if (STYLE == ORC_BUCK_COUL_LONG || (LCL && ORDER1)) {
          // coul/long: whole Coulomb block gated by cutsq (:291); long/coul/long: same (:350)
          if (!ncoultablebits || (LCL ? rsq <= pp.tabinnersq : rsq <= tabinnersq)) {
            const flt_t A1 = 0.254829592, A2 = -0.284496736, A3 = 1.421413741;
            const flt_t A4 = -1.453152027, A5 = 1.061405429;
            const flt_t EWALD_F = 1.12837917;
            const flt_t INV_EWALD_P = 1.0 / 0.3275911;
            // long/coul/long has no flt_t local for g_ewald: the product uses the base class's double member and only
            // the assignment rounds (pair_buck_long_coul_long_intel.cpp:361); coul/long multiplies in flt_t (:167,304)
            const flt_t grij = LCL ? (flt_t)(pp.g_ewald * r) : g_ewald * r;
            const flt_t expm2 = std::exp(-grij * grij);
            const flt_t t = INV_EWALD_P / (INV_EWALD_P + grij);
            const flt_t erfc = t * (A1 + t * (A2 + t * (A3 + t * (A4 + t * A5)))) * expm2;
            const flt_t prefactor = qqrd2e * qtmp * q[j] / r;
            forcecoul = prefactor * (erfc + EWALD_F * grij * expm2);
            if (EFLAG) ecoul = prefactor * erfc;
            // coul/long applies it unconditionally (:312-315), long/coul/long under if(sbindex)
            // (:373-380); identical because special_coul[0] == 1
            const flt_t adjust = ((flt_t)1.0 - fc.special_coul[sbindex]) * prefactor;
            forcecoul -= adjust;
            if (EFLAG) ecoul -= adjust;
          } else {
            const float rsq_lookup = (float)rsq;
            const int itable = (float_bits(rsq_lookup) & ncoulmask) >> ncoulshiftbits;
            const flt_t fraction = ((flt_t)rsq_lookup - fc.rtable[itable]) * fc.drtable[itable];
            const flt_t tablet = fc.ftable[itable] + fraction * fc.dftable[itable];
            forcecoul = qtmp * q[j] * tablet;
            if (EFLAG) ecoul = qtmp * q[j] * (fc.etable[itable] + fraction * fc.detable[itable]);
            if (sbindex) {
              const flt_t table2 = fc.ctable[itable] + fraction * fc.dctable[itable];
              const flt_t prefactor = qtmp * q[j] * table2;
              const flt_t adjust = ((flt_t)1.0 - fc.special_coul[sbindex]) * prefactor;
              forcecoul -= adjust;
              if (EFLAG) ecoul -= adjust;
            }
          }
        }

        if (STYLE == ORC_LJ_LONG_COUL_LONG) {
          // pair_lj_long_coul_long_intel.cpp:620-688: buck1, buck2, a, c hold lj1, lj2, lj3, lj4
          if (rsq < fc.cut_ljsq[ij]) {
            const flt_t lj1 = fc.buck1[ij], lj2 = fc.buck2[ij], lj3 = fc.a[ij], lj4 = fc.c[ij];
            if (ORDER6) {
              const flt_t r6inv = r2inv * r2inv * r2inv;
              if (!ndisptablebits || rsq <= pp.tabinnerdispsq) {  // :622-638 (double literals: see the Buckingham twin)
                const flt_t grij2 = g2 * rsq;
                const flt_t a2 = (flt_t)1.0 / grij2;
                const flt_t x2 = a2 * std::exp(-grij2) * lj4;
                forcebuck = r6inv * r6inv * lj1 - g8 * x2 * rsq * (((6.0 * a2 + 6.0) * a2 + 3.0) * a2 + 1.0);
                if (EFLAG) evdwl = r6inv * r6inv * lj3 - g6 * x2 * ((a2 + 1.0) * a2 + 0.5);
              } else {  // :640-652
                const float rsq_lookup = (float)rsq;
                const int itable = (float_bits(rsq_lookup) & ndispmask) >> ndispshiftbits;
                const flt_t fd = (rsq - fc.rdisp[itable]) * fc.drdisp[itable];
                forcebuck = r6inv * r6inv * lj1 - (fc.fdisp[itable] + fd * fc.dfdisp[itable]) * lj4;
                if (EFLAG) evdwl = r6inv * r6inv * lj3 - (fc.edisp[itable] + fd * fc.dedisp[itable]) * lj4;
              }
              if (sbindex) {  // :632-638, :653-661
                const flt_t f = fc.special_lj[sbindex];
                const flt_t t = r6inv * (1.0 - f);
                forcebuck += t * (lj2 - r6inv * lj1);
                if (EFLAG) evdwl += t * (lj4 - r6inv * lj3);
              }
            } else {  // :664-675
              const flt_t r6inv = r2inv * r2inv * r2inv;
              forcebuck = r6inv * (r6inv * lj1 - lj2);
              if (EFLAG) evdwl = r6inv * (r6inv * lj3 - lj4) - fc.offset[ij];
              if (sbindex) {
                const flt_t factor_lj = fc.special_lj[sbindex];
                forcebuck *= factor_lj;
                if (EFLAG) evdwl *= factor_lj;
              }
            }
          }
        } else if (rsq < fc.cut_ljsq[ij]) {
          const flt_t r6inv = r2inv * r2inv * r2inv;
          const flt_t rexp = std::exp(-r * fc.rhoinv[ij]);
          if (STYLE == ORC_BUCK_LONG_COUL_LONG && ORDER6) {
            if (!ndisptablebits || rsq <= pp.tabinnerdispsq) {  // :414-431 (double member, like :351)
              const flt_t grij2 = g2 * rsq;
              const flt_t a2 = (flt_t)1.0 / grij2;
              const flt_t x2 = a2 * std::exp(-grij2) * fc.c[ij];
              // the reference writes the polynomial with plain double literals (:418-422): in mixed mode it is evaluated in
              // double and only the assignment rounds to flt_t (found by the bitwise check against oracle/_ref)
              forcebuck = r * rexp * fc.buck1[ij] - g8 * x2 * rsq * (((6.0 * a2 + 6.0) * a2 + 3.0) * a2 + 1.0);
              if (EFLAG) evdwl = rexp * fc.a[ij] - g6 * x2 * ((a2 + 1.0) * a2 + 0.5);
            } else {  // :433-444
              const float rsq_lookup = (float)rsq;
              const int itable = (float_bits(rsq_lookup) & ndispmask) >> ndispshiftbits;
              const flt_t fd = (rsq - fc.rdisp[itable]) * fc.drdisp[itable];
              forcebuck = r * rexp * fc.buck1[ij] -
                          (fc.fdisp[itable] + fd * fc.dfdisp[itable]) * fc.c[ij];
              if (EFLAG)
                evdwl = rexp * fc.a[ij] - (fc.edisp[itable] + fd * fc.dedisp[itable]) * fc.c[ij];
            }
            if (sbindex) {  // :423-431, :445-453
              const flt_t f = fc.special_lj[sbindex];
              const flt_t t = (f - 1.0);
              forcebuck += t * r * rexp * fc.buck1[ij] - t * r6inv * fc.buck2[ij];
              if (EFLAG) evdwl += t * rexp * fc.a[ij] - t * r6inv * fc.c[ij];
            }
          } else {
            forcebuck = r * rexp * fc.buck1[ij] - r6inv * fc.buck2[ij];
            if (EFLAG) evdwl = rexp * fc.a[ij] - r6inv * fc.c[ij] - fc.offset[ij];
            if (sbindex) {
              const flt_t factor_lj = fc.special_lj[sbindex];
              forcebuck *= factor_lj;
              if (EFLAG) evdwl *= factor_lj;
            }
          }
        }

        const flt_t fpair = (forcecoul + forcebuck) * r2inv;
        fxtmp += delx * fpair;
        fytmp += dely * fpair;
        fztmp += delz * fpair;
        if (NEWTON_PAIR || j < nlocal) {
          f[j].x -= delx * fpair;
          f[j].y -= dely * fpair;
          f[j].z -= delz * fpair;
        }
        if (EVFLAG) {
          flt_t ev_pre = (flt_t)0;
          if (NEWTON_PAIR || i < nlocal) ev_pre += (flt_t)0.5;
          if (NEWTON_PAIR || j < nlocal) ev_pre += (flt_t)0.5;
          if (EFLAG) {
            sevdwl += ev_pre * evdwl;
            secoul += ev_pre * ecoul;
            if (eatom) {
              if (NEWTON_PAIR || i < nlocal) fwtmp += (flt_t)0.5 * evdwl + (flt_t)0.5 * ecoul;
              if (NEWTON_PAIR || j < nlocal) f[j].w += (flt_t)0.5 * evdwl + (flt_t)0.5 * ecoul;
            }
          }
          if (vflag == 1) {  // IP_PRE_ev_tally_nbor
            sv0 += ev_pre * delx * delx * fpair;
            sv1 += ev_pre * dely * dely * fpair;
            sv2 += ev_pre * delz * delz * fpair;
            sv3 += ev_pre * delx * dely * fpair;
            sv4 += ev_pre * delx * delz * fpair;
            sv5 += ev_pre * dely * delz * fpair;
          }
        }
      }  // jj
      f[i].x += fxtmp;
      f[i].y += fytmp;
      f[i].z += fztmp;
      if (EVFLAG) {  // IP_PRE_ev_tally_atomq
        if (EFLAG) {
          f[i].w += fwtmp;
          oevdwl += sevdwl;
          oecoul += secoul;
        }
        if (vflag == 1) {
          ov0 += sv0; ov1 += sv1; ov2 += sv2; ov3 += sv3; ov4 += sv4; ov5 += sv5;
        }
      }
    }  // ii

#pragma omp barrier
    // IP_PRE_fdotr_acc_force: reduce thread-private arrays into thread 0's; f.r virial if vflag==2
    {
      const int n = NEWTON_PAIR ? nall : nlocal;
      const int delta = 1 + n / nthreads;
      const int from = std::min(tid * delta, n), to = std::min(from + delta, n);
      for (int t = 1; t < nthreads; t++) {
        const force_t *ft = f_start + (size_t)t * f_stride;
        for (int k = from; k < to; k++) {
          f_start[k].x += ft[k].x;
          f_start[k].y += ft[k].y;
          f_start[k].z += ft[k].z;
          f_start[k].w += ft[k].w;
        }
      }
      if (EVFLAG && vflag == 2) {
        for (int k = from; k < to; k++) {
          ov0 += f_start[k].x * x[k].x;
          ov1 += f_start[k].y * x[k].y;
          ov2 += f_start[k].z * x[k].z;
          ov3 += f_start[k].y * x[k].x;
          ov4 += f_start[k].z * x[k].x;
          ov5 += f_start[k].z * x[k].y;
        }
      }
    }
  }  // omp parallel

  for (int k = 0; k < 8; k++) ev_global[k] = 0.0;
  if (EVFLAG) {
    if (EFLAG) {
      ev_global[0] = (double)oevdwl;
      ev_global[1] = (double)oecoul;
    }
    if (vflag) {
      ev_global[2] = (double)ov0; ev_global[3] = (double)ov1; ev_global[4] = (double)ov2;
      ev_global[5] = (double)ov3; ev_global[6] = (double)ov4; ev_global[7] = (double)ov5;
    }
  }
  for (int i = 0; i < nall; i++) {
    fout[4 * (size_t)i] = (double)f_start[i].x;
    fout[4 * (size_t)i + 1] = (double)f_start[i].y;
    fout[4 * (size_t)i + 2] = (double)f_start[i].z;
    fout[4 * (size_t)i + 3] = (double)f_start[i].w;
  }
}

template <int STYLE, class flt_t, class acc_t>
void dispatch(int eflag, int vflag, int eatom, int newton, int nlocal, int nall, const double *x,
              const int *type, const double *q, const int *numneigh, const long *offsets,
              const int *entries, const orc_pair_params &p, double *f, double *ev, int nthreads) {
#define CALL(EV, E, N)                                                                          \
  eval<STYLE, EV, E, N, flt_t, acc_t>(vflag, eatom, nlocal, nall, x, type, q, numneigh, offsets, \
                                      entries, p, f, ev, nthreads)
  // template dispatch of compute<flt_t,acc_t> (pair_buck_intel.cpp:64-123)
  if (eflag || vflag) {
    if (eflag) { if (newton) CALL(1, 1, 1); else CALL(1, 1, 0); }
    else       { if (newton) CALL(1, 0, 1); else CALL(1, 0, 0); }
  } else       { if (newton) CALL(0, 0, 1); else CALL(0, 0, 0); }
#undef CALL
}

}  // namespace

extern "C" {

void orc_pair_init(int style, int ntypes, const double *A, const double *rho, const double *C,
                   const double *cut_lj, const double *cut_coul, int offset_flag,
                   orc_pair_params *p) {
  const int tp1 = ntypes + 1;
  p->ntypes = ntypes;
  for (int i = 1; i <= ntypes; i++)
    for (int j = 1; j <= ntypes; j++) {
      const int ij = i * tp1 + j;
      const double cl = cut_lj[ij];
      const double cc = cut_coul ? cut_coul[ij] : 0.0;
      if (style == ORC_LJ_LONG_COUL_LONG) {
        // PairLJLongCoulLong::init_one [UPSTREAM]: A = epsilon, rho = sigma
        const double eps = A[ij], sig = rho[ij];
        p->buck1[ij] = 48.0 * eps * std::pow(sig, 12.0);   // lj1
        p->buck2[ij] = 24.0 * eps * std::pow(sig, 6.0);    // lj2
        p->a[ij] = 4.0 * eps * std::pow(sig, 12.0);        // lj3
        p->c[ij] = 4.0 * eps * std::pow(sig, 6.0);         // lj4
        p->rhoinv[ij] = 0.0;
        if (offset_flag && cl > 0.0) {
          const double ratio = sig / cl;
          p->offset[ij] = 4.0 * eps * (std::pow(ratio, 12.0) - std::pow(ratio, 6.0));
        } else
          p->offset[ij] = 0.0;
        p->cut_ljsq[ij] = cl * cl;
        p->cut_coulsq[ij] = cc * cc;
        const double cut = std::max(cl, cc);
        p->cutsq[ij] = cut * cut;
        continue;
      }
      p->a[ij] = A[ij];
      p->c[ij] = C[ij];
      p->rhoinv[ij] = 1.0 / rho[ij];
      p->buck1[ij] = A[ij] / rho[ij];
      p->buck2[ij] = 6.0 * C[ij];
      if (offset_flag) {
        const double rexp = std::exp(-cl / rho[ij]);
        p->offset[ij] = A[ij] * rexp - C[ij] / std::pow(cl, 6.0);
      } else
        p->offset[ij] = 0.0;
      p->cut_ljsq[ij] = cl * cl;
      p->cut_coulsq[ij] = cc * cc;
      const double cut = style == ORC_BUCK ? cl : std::max(cl, cc);
      p->cutsq[ij] = cut * cut;
    }
}

void orc_init_coul_tables(double cut_coul, double tabinner, int nbits, double g_ewald, double qqrd2e,
                          double *rtable, double *drtable, double *ftable, double *dftable,
                          double *etable, double *detable, double *ctable, double *dctable,
                          int *ncoulmask_out, int *ncoulshiftbits_out, double *tabinnersq_out) {
  const double EWALD_F = 1.12837917;
  union u_if { int i; float f; };
  // Pair::init_bitmap
  const double inner = tabinner, outer = cut_coul;
  int nlowermin = 1;
  while (!((std::pow(2.0, nlowermin) <= inner * inner) &&
           (std::pow(2.0, nlowermin + 1) > inner * inner))) {
    if (std::pow(2.0, nlowermin) <= inner * inner) nlowermin++;
    else nlowermin--;
  }
  int nexpbits = 0;
  const double required_range = outer * outer / std::pow(2.0, nlowermin);
  double available_range = 2.0;
  while (available_range < required_range) {
    nexpbits++;
    available_range = std::pow(2.0, std::pow(2.0, nexpbits));
  }
  const int nmantbits = nbits - nexpbits;
  const int nshiftbits = FLT_MANT_DIG - (nmantbits + 1);
  int nmask = 1;
  for (int j = 0; j < nbits + nshiftbits; j++) nmask *= 2;
  nmask -= 1;
  u_if rsq_lookup;
  rsq_lookup.f = (float)(outer * outer);
  const int maskhi = rsq_lookup.i & ~(nmask);
  rsq_lookup.f = (float)(inner * inner);
  const int masklo = rsq_lookup.i & ~(nmask);

  // Pair::init_tables
  const double cut_coulsq = cut_coul * cut_coul;
  double tabinnersq = tabinner * tabinner;
  const int ntable = 1 << nbits;
  u_if minrsq_lookup;
  minrsq_lookup.i = 0 << nshiftbits;
  minrsq_lookup.i |= maskhi;
  for (int i = 0; i < ntable; i++) {
    rsq_lookup.i = i << nshiftbits;
    rsq_lookup.i |= masklo;
    if (rsq_lookup.f < tabinnersq) {
      rsq_lookup.i = i << nshiftbits;
      rsq_lookup.i |= maskhi;
    }
    const double r = sqrtf(rsq_lookup.f);
    const double grij = g_ewald * r;
    const double expm2 = std::exp(-grij * grij);
    const double derfc = std::erfc(grij);
    rtable[i] = rsq_lookup.f;
    ctable[i] = qqrd2e / r;
    ftable[i] = qqrd2e / r * (derfc + EWALD_F * grij * expm2);
    etable[i] = qqrd2e / r * derfc;
    minrsq_lookup.f = std::min(minrsq_lookup.f, rsq_lookup.f);
  }
  tabinnersq = minrsq_lookup.f;
  const int ntablem1 = ntable - 1;
  for (int i = 0; i < ntablem1; i++) {
    drtable[i] = 1.0 / (rtable[i + 1] - rtable[i]);
    dftable[i] = ftable[i + 1] - ftable[i];
    dctable[i] = ctable[i + 1] - ctable[i];
    detable[i] = etable[i + 1] - etable[i];
  }
  drtable[ntablem1] = 1.0 / (rtable[0] - rtable[ntablem1]);
  dftable[ntablem1] = ftable[0] - ftable[ntablem1];
  dctable[ntablem1] = ctable[0] - ctable[ntablem1];
  detable[ntablem1] = etable[0] - etable[ntablem1];
  int itablemin = minrsq_lookup.i & nmask;
  itablemin >>= nshiftbits;
  int itablemax = itablemin - 1;
  if (itablemin == 0) itablemax = ntablem1;
  rsq_lookup.i = itablemax << nshiftbits;
  rsq_lookup.i |= maskhi;
  if (rsq_lookup.f < cut_coulsq) {
    rsq_lookup.f = (float)cut_coulsq;
    const double r = sqrtf(rsq_lookup.f);
    const double grij = g_ewald * r;
    const double expm2 = std::exp(-grij * grij);
    const double derfc = std::erfc(grij);
    const double f_tmp = qqrd2e / r * (derfc + EWALD_F * grij * expm2);
    const double e_tmp = qqrd2e / r * derfc;
    const double c_tmp = qqrd2e / r;
    drtable[itablemax] = 1.0 / (rsq_lookup.f - rtable[itablemax]);
    dftable[itablemax] = f_tmp - ftable[itablemax];
    dctable[itablemax] = c_tmp - ctable[itablemax];
    detable[itablemax] = e_tmp - etable[itablemax];
  }
  *ncoulmask_out = nmask;
  *ncoulshiftbits_out = nshiftbits;
  *tabinnersq_out = tabinnersq;
}

void orc_pair_eval(int style, int prec, int eflag, int vflag, int eatom, int newton, int nlocal,
                   int nall, const double *x, const int *type, const double *q, const int *numneigh,
                   const long *offsets, const int *entries, const orc_pair_params *p, double *f,
                   double *ev, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#define STYLE_CASE(S)                                                                          \
  case S:                                                                                      \
    if (prec == ORC_PREC_DOUBLE)                                                               \
      dispatch<S, double, double>(eflag, vflag, eatom, newton, nlocal, nall, x, type, q,       \
                                  numneigh, offsets, entries, *p, f, ev, nthreads);            \
    else                                                                                       \
      dispatch<S, float, double>(eflag, vflag, eatom, newton, nlocal, nall, x, type, q,        \
                                 numneigh, offsets, entries, *p, f, ev, nthreads);             \
    break;
  switch (style) {
    STYLE_CASE(ORC_BUCK)
    STYLE_CASE(ORC_BUCK_COUL_CUT)
    STYLE_CASE(ORC_BUCK_COUL_LONG)
    STYLE_CASE(ORC_BUCK_LONG_COUL_LONG)
    STYLE_CASE(ORC_LJ_LONG_COUL_LONG)
  }
#undef STYLE_CASE
}

void orc_nve_dtfm(int nlocal, const int *type, const double *mass, double dt, double ftm2v,
                  double *dtfm) {
  // FixNVEIntel::reset_dt (fix_nve_intel.cpp:129-194), igroup == all, per-type mass
  const double dtf = 0.5 * dt * ftm2v;
  int n = 0;
  for (int i = 0; i < nlocal; i++) {
    dtfm[n++] = dtf / mass[type[i]];
    dtfm[n++] = dtf / mass[type[i]];
    dtfm[n++] = dtf / mass[type[i]];
  }
}

void orc_nve_initial(int nlocal, double *x, double *v, const double *f, const double *dtfm,
                     double dtv) {
  // FixNVEIntel::initial_integrate (fix_nve_intel.cpp:60-99), igroup == 0 branch
  const long n3 = 3L * nlocal;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n3; i++) {
    v[i] += dtfm[i] * f[i];
    x[i] += dtv * v[i];
  }
}

void orc_nve_dtfm_group(int nlocal, const int *type, const double *mass, const double *rmass, const int *ingroup,
                        double dt, double ftm2v, double *dtfm) {
  // FixNVEIntel::reset_dt (fix_nve_intel.cpp:147-190): rmass / per-type mass, 0 for atoms outside the group
  const double dtf = 0.5 * dt * ftm2v;
  int n = 0;
  for (int i = 0; i < nlocal; i++) {
    const double m = rmass ? rmass[i] : mass[type[i]];
    const double d = (ingroup && !ingroup[i]) ? 0.0 : dtf / m;
    dtfm[n++] = d;
    dtfm[n++] = d;
    dtfm[n++] = d;
  }
}

void orc_nve_initial_group(int nlocal, double *x, double *v, const double *f, const double *dtfm, double dtv) {
  // FixNVEIntel::initial_integrate, igroup != 0 branch (fix_nve_intel.cpp:88-97)
  const long n3 = 3L * nlocal;
  for (long i = 0; i < n3; i++) {
    if (dtfm[i] != 0.0) {
      v[i] += dtfm[i] * f[i];
      x[i] += dtv * v[i];
    }
  }
}

void orc_nve_final(int nlocal, double *v, const double *f, const double *dtfm) {
  // FixNVEIntel::final_integrate (fix_nve_intel.cpp:103-127)
  const long n3 = 3L * nlocal;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n3; i++) v[i] += dtfm[i] * f[i];
}

}  // extern "C"
