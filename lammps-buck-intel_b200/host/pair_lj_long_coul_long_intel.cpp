// pair_lj_long_coul_long_intel.cpp — host side of lj/long/coul/long/intel: the stock base class's parameter logic
// (PairLJLongCoulLong::settings / coeff / init_one [UPSTREAM]) and the hand-over of pack_force_const's products
// (pair_lj_long_coul_long_intel.cpp:798-867: cutsq, cut_ljsq, lj1..lj4, offset, g_ewald, tables) to the device.
#include "pair_lj_long_coul_long_intel.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace LAMMPS_NS;

void PairLJLongCoulLong::settings(int narg, char **arg) {
  if (narg != 3 && narg != 4) error->all(FLERR, "Illegal pair_style command");
  ewald_order = ewald_off = 0;
  auto option = [&](const char *a, int order) {
    if (std::strcmp(a, "long") == 0) ewald_order |= 1 << order;
    else if (std::strcmp(a, "cut") == 0) {}
    else if (std::strcmp(a, "off") == 0) ewald_off |= 1 << order;
    else error->all(FLERR, "Illegal pair_style lj/long/coul/long command");
  };
  option(arg[0], 6);
  option(arg[1], 1);
  if (!((ewald_order ^ ewald_off) & (1 << 1)))
    error->all(FLERR, "Coulomb cut not supported in pair_style lj/long/coul/long");
  cut_global = std::atof(arg[2]);
  cut_coul = narg == 4 ? std::atof(arg[3]) : cut_global;
  ewaldflag = (ewald_order >> 1) & 1;
  dispersionflag = (ewald_order >> 6) & 1;
  if (allocated)
    for (size_t ij = 0; ij < setflag.size(); ij++)
      if (setflag[ij]) k.cut_lj[ij] = cut_global;
}

void PairLJLongCoulLong::coeff(int narg, char **arg) {
  if (narg < 4 || narg > 5) error->all(FLERR, "Incorrect args for pair coefficients");
  if (!allocated) {
    allocate();
    epsilon.assign(setflag.size(), 0.0);
    sigma.assign(setflag.size(), 0.0);
  }
  int ilo, ihi, jlo, jhi;
  bounds(error, arg[0], atom->ntypes, ilo, ihi);
  bounds(error, arg[1], atom->ntypes, jlo, jhi);
  const double eps = std::atof(arg[2]), sig = std::atof(arg[3]);
  const double cut_one = narg == 5 ? std::atof(arg[4]) : cut_global;
  set_pair(ilo, ihi, jlo, jhi, 0.0, 1.0, 0.0, cut_one, 0.0);
  const int n = tp1();
  for (int i = ilo; i <= ihi; i++)
    for (int j = std::max(jlo, i); j <= jhi; j++) {
      epsilon[i * n + j] = eps;
      sigma[i * n + j] = sig;
    }
}

// lj1 = 48 eps sigma^12, lj2 = 24 eps sigma^6, lj3 = 4 eps sigma^12, lj4 = 4 eps sigma^6; offset = 4 eps ((s/rc)^12 - (s/rc)^6)
double PairLJLongCoulLong::init_one(int i, int j) {
  const int n = tp1(), ij = i * n + j, ji = j * n + i;
  // no explicit i-j coefficients: Pair::mix_energy / mix_distance [UPSTREAM] with `pair_modify mix` (stock default of
  // this style: geometric)
  if (!setflag[ij]) {
    if (!setflag[i * n + i] || !setflag[j * n + j]) error->all(FLERR, "All pair coeffs are not set");
    epsilon[ij] = std::sqrt(epsilon[i * n + i] * epsilon[j * n + j]);
    sigma[ij] = mix_flag == ARITHMETIC ? 0.5 * (sigma[i * n + i] + sigma[j * n + j])
                                       : std::sqrt(sigma[i * n + i] * sigma[j * n + j]);
    k.cut_lj[ij] = std::max(k.cut_lj[i * n + i], k.cut_lj[j * n + j]);
    setflag[ij] = 1;
  }
  if (ewald_order & (1 << 6)) k.cut_lj[ij] = cut_global;   // long dispersion: one global cut-off
  k.cut_coul[ij] = (ewald_off & (1 << 1)) ? 0.0 : cut_coul;
  const double eps = epsilon[ij], sig = sigma[ij];
  k.buck1[ij] = 48.0 * eps * std::pow(sig, 12.0);
  k.buck2[ij] = 24.0 * eps * std::pow(sig, 6.0);
  k.a[ij] = 4.0 * eps * std::pow(sig, 12.0);
  k.c[ij] = 4.0 * eps * std::pow(sig, 6.0);
  k.rhoinv[ij] = 0.0;
  if (offset_flag && k.cut_lj[ij] > 0.0) {
    const double ratio = sig / k.cut_lj[ij];
    k.offset[ij] = 4.0 * eps * (std::pow(ratio, 12.0) - std::pow(ratio, 6.0));
  } else k.offset[ij] = 0.0;
  k.cut_ljsq[ij] = k.cut_lj[ij] * k.cut_lj[ij];
  k.cut_coulsq[ij] = k.cut_coul[ij] * k.cut_coul[ij];
  epsilon[ji] = eps;
  sigma[ji] = sig;
  for (auto *v : {&k.a, &k.c, &k.cut_lj, &k.cut_coul, &k.rhoinv, &k.buck1, &k.buck2, &k.offset, &k.cut_ljsq, &k.cut_coulsq})
    (*v)[ji] = (*v)[ij];
  return std::max(k.cut_lj[ij], k.cut_coul[ij]);
}

void PairLJLongCoulLong::init_style() {
  if (!atom->q_flag && (ewald_order & (1 << 1)))
    error->all(FLERR, "Invoking coulombic in pair style lj/long/coul/long requires atom attribute q");
  if (ewald_order & ((1 << 1) | (1 << 6))) {
    if (!force->kspace) error->all(FLERR, "Pair style requires a KSpace style");
    g_ewald = force->kspace->g_ewald;
    g_ewald_6 = force->kspace->g_ewald_6;               // pair_lj_long_coul_long_intel.cpp:479
  }
  ctab = PairTables();
  dtab = PairTables();
  if ((ewald_order & (1 << 1)) && ncoultablebits) init_tables(cut_coul, g_ewald, ctab);
  if ((ewald_order & (1 << 6)) && ndisptablebits) init_tables_disp(cut_global, g_ewald_6, dtab);
}

void *PairLJLongCoulLong::extract(const char *str, int &dim) {
  dim = 0;
  if (std::strcmp(str, "cut_coul") == 0) return &cut_coul;
  if (std::strcmp(str, "ewald_order") == 0) return &ewald_order;
  if (std::strcmp(str, "ewald_mix") == 0) return &mix_flag;
  if (std::strcmp(str, "cut_LJ") == 0) return &cut_global;
  dim = 2;   // [(ntypes+1)^2] matrices read by PPPMDisp::init_coeffs
  if (std::strcmp(str, "B") == 0) return k.c.data();          // lj4 = 4 eps sigma^6
  if (std::strcmp(str, "epsilon") == 0) return epsilon.data();
  if (std::strcmp(str, "sigma") == 0) return sigma.data();
  dim = 0;
  return nullptr;
}

void PairLJLongCoulLongIntel::init_style() {
  PairLJLongCoulLong::init_style();
  fix = require_fix_intel();
  init_all_pairs();
  device_setup(fix, B200MD_PAIR_LJ_LONG_COUL_LONG, g_ewald, g_ewald_6, ewald_order, &ctab, &dtab);
}

void PairLJLongCoulLongIntel::compute(int eflag, int vflag) {
  if (!fix) error->all(FLERR, "Pair style lj/long/coul/long/intel used before init_style()");
  device_compute(fix, eflag, vflag);
}
