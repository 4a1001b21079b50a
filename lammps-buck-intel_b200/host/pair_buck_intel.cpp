// pair_buck_intel.cpp — host side of buck/intel.  What the reference does in
//   compute()          pair_buck_intel.cpp:48-123   (precision dispatch, ev_setup, repack, eval<> choice)
//   eval<>()           :127-365                      (the pair loop)            -> b200md_pair_compute
//   init_style()       :367-389                      ("package intel" check, neighbour request)
//   pack_force_const() :391-443                      (T x T constant tables)    -> b200md_pair_setup
// becomes parameter marshalling around the C ABI; the arithmetic lives in csrc/pair_kernel.cuh.
// PairBuck::settings/coeff/init_one restate the stock base class.
#include "pair_buck_intel.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>

using namespace LAMMPS_NS;

// ---- upstream Pair plumbing --------------------------------------------------------------------------------
void Pair::ev_setup(int eflag, int vflag) {
  eflag_either = eflag;
  eflag_global = eflag & 1;
  eflag_atom = eflag & 2;
  vflag_either = vflag;
  vflag_global = vflag & 3;
  vflag_fdotr = (vflag & 2) ? 1 : 0;
  eng_vdwl = eng_coul = 0.0;
  for (double &v : virial) v = 0.0;
  if (eflag_atom) eatom.assign(atom->nlocal, 0.0);
}

void Pair::init() {
  if (!allocated) error->all(FLERR, "All pair coeffs are not set");
  const int n = atom->ntypes;
  for (int i = 1; i <= n; i++)
    if (!setflag[i * tp1() + i]) error->all(FLERR, "All pair coeffs are not set");
  init_style();
  cutforce = 0.0;
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) {
      const double cut = init_one(i, j);
      cutsq[i * tp1() + j] = cutsq[j * tp1() + i] = cut * cut;
      cutforce = std::max(cutforce, cut);
    }
}

void PairBuck::bounds(Error *error, const char *str, int nmax, int &nlo, int &nhi) {
  const char *star = std::strchr(str, '*');
  const int n = (int)std::strlen(str);
  if (!star) nlo = nhi = std::atoi(str);
  else if (n == 1) { nlo = 1; nhi = nmax; }
  else if (star == str) { nlo = 1; nhi = std::atoi(str + 1); }
  else if (star == str + n - 1) { nlo = std::atoi(str); nhi = nmax; }
  else { nlo = std::atoi(str); nhi = std::atoi(star + 1); }
  if (nlo < 1 || nhi > nmax || nlo > nhi) error->all(FLERR, "Numeric index is out of bounds");
}

void PairBuck::allocate() {
  allocated = 1;
  const int n = tp1();
  setflag.assign((size_t)n * n, 0);
  cutsq.assign((size_t)n * n, 0.0);
  k.allocate(n);
}

void PairBuck::set_pair(int ilo, int ihi, int jlo, int jhi, double a, double rho, double c, double cut_lj,
                        double cut_coul) {
  int count = 0;
  const int n = tp1();
  for (int i = ilo; i <= ihi; i++)
    for (int j = std::max(jlo, i); j <= jhi; j++) {
      k.a[i * n + j] = a;
      k.rho[i * n + j] = rho;
      k.c[i * n + j] = c;
      k.cut_lj[i * n + j] = cut_lj;
      k.cut_coul[i * n + j] = cut_coul;
      setflag[i * n + j] = 1;
      count++;
    }
  if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
}

void PairBuck::settings(int narg, char **arg) {
  if (narg != 1) error->all(FLERR, "Illegal pair_style command");
  cut_global = std::atof(arg[0]);
  if (allocated)
    for (size_t ij = 0; ij < setflag.size(); ij++)
      if (setflag[ij]) k.cut_lj[ij] = cut_global;
}

void PairBuck::coeff(int narg, char **arg) {
  if (narg < 5 || narg > 6) error->all(FLERR, "Incorrect args for pair coefficients");
  if (!allocated) allocate();
  int ilo, ihi, jlo, jhi;
  bounds(error, arg[0], atom->ntypes, ilo, ihi);
  bounds(error, arg[1], atom->ntypes, jlo, jhi);
  const double a = std::atof(arg[2]), rho = std::atof(arg[3]), c = std::atof(arg[4]);
  if (rho <= 0) error->all(FLERR, "Incorrect args for pair coefficients");
  const double cut = narg == 6 ? std::atof(arg[5]) : cut_global;
  set_pair(ilo, ihi, jlo, jhi, a, rho, c, cut, 0.0);
}

double PairBuck::init_one(int i, int j) {
  const int n = tp1();
  if (!setflag[i * n + j]) error->all(FLERR, "All pair coeffs are not set");   // no mixing rule for Buckingham
  const int ij = i * n + j, ji = j * n + i;
  k.rhoinv[ij] = 1.0 / k.rho[ij];
  k.buck1[ij] = k.a[ij] / k.rho[ij];
  k.buck2[ij] = 6.0 * k.c[ij];
  if (offset_flag && k.cut_lj[ij] > 0.0) {
    const double rexp = std::exp(-k.cut_lj[ij] / k.rho[ij]);
    k.offset[ij] = k.a[ij] * rexp - k.c[ij] / std::pow(k.cut_lj[ij], 6.0);
  } else k.offset[ij] = 0.0;
  k.cut_ljsq[ij] = k.cut_lj[ij] * k.cut_lj[ij];
  k.cut_coulsq[ij] = k.cut_coul[ij] * k.cut_coul[ij];
  for (auto *v : {&k.a, &k.rho, &k.c, &k.cut_lj, &k.cut_coul, &k.rhoinv, &k.buck1, &k.buck2, &k.offset, &k.cut_ljsq,
                  &k.cut_coulsq})
    (*v)[ji] = (*v)[ij];
  return std::max(k.cut_lj[ij], k.cut_coul[ij]);
}

void PairBuck::init_all_pairs() {
  const int n = atom->ntypes;
  if (!allocated) error->all(FLERR, "All pair coeffs are not set");
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) {
      const double cut = init_one(i, j);
      cutsq[i * tp1() + j] = cutsq[j * tp1() + i] = cut * cut;
    }
}

FixIntel *PairBuck::require_fix_intel() {
  if (!lmp->fix_intel && lmp->dry_run) return nullptr;
  if (!lmp->fix_intel) error->all(FLERR, "The 'package intel' command is required for /intel styles");
  return lmp->fix_intel;
}

// pack_force_const of all four styles: hand init_one()'s products to the device
void PairBuck::device_setup(FixIntel *fix, int style, double g_ewald, double g_ewald_6, int ewald_order,
                            const PairTables *ctab, const PairTables *dtab) {
  if (!fix) return;   // dry run
  b200md_pair_params p;
  std::memset(&p, 0, sizeof(p));
  p.style = style;
  p.ntypes = atom->ntypes;
  p.cutsq = cutsq.data();
  p.cut_ljsq = k.cut_ljsq.data();
  p.cut_coulsq = k.cut_coulsq.data();
  p.buck1 = k.buck1.data(); p.buck2 = k.buck2.data(); p.rhoinv = k.rhoinv.data();
  p.a = k.a.data(); p.c = k.c.data(); p.offset = k.offset.data();
  for (int i = 0; i < 4; i++) { p.special_lj[i] = force->special_lj[i]; p.special_coul[i] = force->special_coul[i]; }
  p.g_ewald = g_ewald;
  p.g_ewald_6 = g_ewald_6;
  p.ewald_order = ewald_order;
  if (ctab && ctab->nbits) {
    p.ncoultablebits = ctab->nbits; p.ncoulmask = ctab->mask; p.ncoulshiftbits = ctab->shiftbits;
    p.tabinnersq = ctab->tabinnersq;
    p.rtable = ctab->r.data(); p.drtable = ctab->dr.data(); p.ftable = ctab->f.data(); p.dftable = ctab->df.data();
    p.etable = ctab->e.data(); p.detable = ctab->de.data(); p.ctable = ctab->c.data(); p.dctable = ctab->dc.data();
  }
  if (dtab && dtab->nbits) {
    p.ndisptablebits = dtab->nbits; p.ndispmask = dtab->mask; p.ndispshiftbits = dtab->shiftbits;
    p.tabinnerdispsq = dtab->tabinnersq;
    p.rdisptable = dtab->r.data(); p.drdisptable = dtab->dr.data(); p.fdisptable = dtab->f.data();
    p.dfdisptable = dtab->df.data(); p.edisptable = dtab->e.data(); p.dedisptable = dtab->de.data();
  }
  fix->check(b200md_pair_setup(fix->ctx(), &p));
}

// compute<flt_t,acc_t> + eval<>: one call; ev_global[8] comes back as {evdwl, ecoul, v0..v5}
void PairBuck::device_compute(FixIntel *fix, int eflag, int vflag) {
  if (eflag || vflag) ev_setup(eflag, vflag);
  else { eflag_either = vflag_either = eflag_global = eflag_atom = vflag_global = vflag_fdotr = 0; }
  if (!fix->resident) {   // host owns the positions (plug-in deployment): repack like IntelBuffers::thr_pack
    fix->check(b200md_atoms_set_x(fix->ctx(), atom->x.data()));
    int rebuilt = 0;
    fix->check(b200md_neigh_decide(fix->ctx(), update->ntimestep, &rebuilt));
  }
  double ev[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  fix->check(b200md_pair_compute(fix->ctx(), eflag, vflag, ev));
  if (eflag_global) { eng_vdwl = ev[0]; eng_coul = ev[1]; }
  if (vflag_global) for (int n = 0; n < 6; n++) virial[n] = ev[2 + n];
  if (!fix->resident || eflag_atom) {
    atom->f.assign((size_t)3 * atom->nlocal, 0.0);
    fix->check(b200md_atoms_download(fix->ctx(), nullptr, nullptr, atom->f.data(), eflag_atom ? eatom.data() : nullptr));
  }
}

// ---- Pair::init_bitmap / init_tables / init_tables_disp (App. A.2; consumed at pair_buck_coul_long_intel.cpp:
//      531-542 and pair_buck_long_coul_long_intel.cpp:433-454) ------------------------------------------------------
namespace {
union IntFloat { int i; float f; };
}

void PairBuck::init_bitmap(double inner, double outer, int ntablebits, int &masklo, int &maskhi, int &nmask,
                           int &nshiftbits) {
  int nlowermin = 1;
  while (!((std::pow(2.0, (double)nlowermin) <= inner * inner) && (std::pow(2.0, (double)nlowermin + 1.0) > inner * inner))) {
    if (std::pow(2.0, (double)nlowermin) <= inner * inner) nlowermin++;
    else nlowermin--;
  }
  int nexpbits = 0;
  const double required_range = outer * outer / std::pow(2.0, (double)nlowermin);
  double available_range = 2.0;
  while (available_range < required_range) {
    nexpbits++;
    available_range = std::pow(2.0, std::pow(2.0, (double)nexpbits));
  }
  const int nmantbits = ntablebits - nexpbits;
  nshiftbits = 24 - (nmantbits + 1);   // FLT_MANT_DIG - (nmantbits + 1)
  nmask = 1;
  for (int j = 0; j < ntablebits + nshiftbits; j++) nmask *= 2;
  nmask -= 1;
  IntFloat rsq_lookup;
  rsq_lookup.f = (float)(outer * outer);
  maskhi = rsq_lookup.i & ~nmask;
  rsq_lookup.f = (float)(inner * inner);
  masklo = rsq_lookup.i & ~nmask;
}


// deltas of a bitmapped table: d[i] = v[i+1] - v[i] (dr = 1/(r[i+1]-r[i])), the last entry wraps to entry 0 and the
// entry just below the minimum is re-done against the true outer cut-off when the table stops short of it
template <class F>
static void finish_table(int ntable, int nmask, int nshiftbits, int maskhi, double minrsq, double outer_sq,
                         std::vector<double> &r, std::vector<double> &dr, std::vector<std::vector<double> *> v,
                         std::vector<std::vector<double> *> dv, F eval_at) {
  const int nv = (int)v.size();
  for (int i = 0; i < ntable; i++) {
    const int nx = (i + 1) % ntable;
    dr[i] = 1.0 / (r[nx] - r[i]);
    for (int q = 0; q < nv; q++) (*dv[q])[i] = (*v[q])[nx] - (*v[q])[i];
  }
  IntFloat lk;
  lk.f = (float)minrsq;
  const int itablemin = (lk.i & nmask) >> nshiftbits;
  const int itablemax = itablemin == 0 ? ntable - 1 : itablemin - 1;
  lk.i = (itablemax << nshiftbits) | maskhi;
  if (lk.f < outer_sq) {
    lk.f = (float)outer_sq;
    std::vector<double> top(nv);
    eval_at((double)lk.f, top);
    dr[itablemax] = 1.0 / ((double)lk.f - r[itablemax]);
    for (int q = 0; q < nv; q++) (*dv[q])[itablemax] = top[q] - (*v[q])[itablemax];
  }
}

void PairBuck::init_tables(double cut_coul, double g_ewald, PairTables &t) const {
  const double EWALD_F = 1.12837917, qqrd2e = force->qqrd2e;
  t.nbits = ncoultablebits;
  if (!t.nbits) return;
  int masklo, maskhi;
  init_bitmap(tabinner, cut_coul, t.nbits, masklo, maskhi, t.mask, t.shiftbits);
  const int ntable = 1 << t.nbits;
  for (auto *v : {&t.r, &t.dr, &t.f, &t.df, &t.e, &t.de, &t.c, &t.dc}) v->assign(ntable, 0.0);
  double tabinnersq = tabinner * tabinner;
  auto eval_at = [&](double rsq, std::vector<double> &out) {   // {f, e, c} at rsq (r from a float sqrt, as upstream)
    const double r = (double)std::sqrt((float)rsq);
    const double grij = g_ewald * r, expm2 = std::exp(-grij * grij), derfc = std::erfc(grij);
    out[0] = qqrd2e / r * (derfc + EWALD_F * grij * expm2);
    out[1] = qqrd2e / r * derfc;
    out[2] = qqrd2e / r;
  };
  IntFloat lk, mn;
  mn.i = maskhi;
  std::vector<double> o(3);
  for (int i = 0; i < ntable; i++) {
    lk.i = (i << t.shiftbits) | masklo;
    if (lk.f < tabinnersq) lk.i = (i << t.shiftbits) | maskhi;
    eval_at((double)lk.f, o);
    t.r[i] = lk.f; t.f[i] = o[0]; t.e[i] = o[1]; t.c[i] = o[2];
    mn.f = std::min(mn.f, lk.f);
  }
  t.tabinnersq = mn.f;
  finish_table(ntable, t.mask, t.shiftbits, maskhi, mn.f, cut_coul * cut_coul, t.r, t.dr, {&t.f, &t.e, &t.c},
               {&t.df, &t.de, &t.dc}, eval_at);
}

void PairBuck::init_tables_disp(double cut_lj_global, double g_ewald_6, PairTables &t) const {
  t.nbits = ndisptablebits;
  if (!t.nbits) return;
  const double g2 = g_ewald_6 * g_ewald_6, g6 = g2 * g2 * g2, g8 = g6 * g2;
  int masklo, maskhi;
  init_bitmap(tabinner_disp, cut_lj_global, t.nbits, masklo, maskhi, t.mask, t.shiftbits);
  const int ntable = 1 << t.nbits;
  for (auto *v : {&t.r, &t.dr, &t.f, &t.df, &t.e, &t.de}) v->assign(ntable, 0.0);
  const double tabinnersq = tabinner_disp * tabinner_disp;
  auto eval_at = [&](double rsq, std::vector<double> &out) {   // {fdisp, edisp} at rsq
    double x2 = g2 * rsq;
    const double a2 = 1.0 / x2;
    x2 = a2 * std::exp(-x2);
    out[0] = g8 * (((6.0 * a2 + 6.0) * a2 + 3.0) * a2 + 1.0) * x2 * rsq;
    out[1] = g6 * ((a2 + 1.0) * a2 + 0.5) * x2;
  };
  IntFloat lk, mn;
  mn.i = maskhi;
  std::vector<double> o(2);
  for (int i = 0; i < ntable; i++) {
    lk.i = (i << t.shiftbits) | masklo;
    if (lk.f < tabinnersq) lk.i = (i << t.shiftbits) | maskhi;
    eval_at((double)lk.f, o);
    t.r[i] = lk.f; t.f[i] = o[0]; t.e[i] = o[1];
    mn.f = std::min(mn.f, lk.f);
  }
  t.tabinnersq = mn.f;
  finish_table(ntable, t.mask, t.shiftbits, maskhi, mn.f, cut_lj_global * cut_lj_global, t.r, t.dr, {&t.f, &t.e},
               {&t.df, &t.de}, eval_at);
}

// ---- PairBuckIntel ---------------------------------------------------------------------------------------
void PairBuckIntel::init_style() {
  PairBuck::init_style();
  fix = require_fix_intel();                                      // pair_buck_intel.cpp:372-376
  // pair_buck_intel.cpp:370 requests the "intel" neighbour list; here the device builds a full list
  init_all_pairs();
  device_setup(fix, B200MD_PAIR_BUCK, 0.0, 0.0, 0, nullptr, nullptr);   // pack_force_const, :391-443
}

void PairBuckIntel::compute(int eflag, int vflag) {
  if (!fix) error->all(FLERR, "Pair style buck/intel used before init_style()");
  device_compute(fix, eflag, vflag);
}
