// pair_buck_coul_intel.cpp — host side of buck/coul/cut/intel, buck/coul/long/intel, buck/long/coul/long/intel.
//   PairBuckCoulCutIntel       compute pair_buck_coul_cut_intel.cpp:55-130, init_style :404-429, pack :431-492
//   PairBuckCoulLongIntel      compute pair_buck_coul_long_intel.cpp:55-130, init_style :457-479, pack :481-566
//                              (g_ewald from force->kspace :507, Coulomb tables :531-542)
//   PairBuckLongCoulLongIntel  compute pair_buck_long_coul_long_intel.cpp:57-209, init_style :542-571, pack :573-646
// The eval<> loops of those files are csrc/pair_kernel.cuh; here only parameters move.
#include "pair_buck_coul_intel.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace LAMMPS_NS;

// ---- buck/coul/cut -----------------------------------------------------------------------------------------
void PairBuckCoulCut::settings(int narg, char **arg) {
  if (narg < 1 || narg > 2) error->all(FLERR, "Illegal pair_style command");
  cut_global = std::atof(arg[0]);
  cut_coul_global = narg == 2 ? std::atof(arg[1]) : cut_global;
  if (allocated)
    for (size_t ij = 0; ij < setflag.size(); ij++)
      if (setflag[ij]) { k.cut_lj[ij] = cut_global; k.cut_coul[ij] = cut_coul_global; }
}

void PairBuckCoulCut::coeff(int narg, char **arg) {
  if (narg < 5 || narg > 7) error->all(FLERR, "Incorrect args for pair coefficients");
  if (!allocated) allocate();
  int ilo, ihi, jlo, jhi;
  bounds(error, arg[0], atom->ntypes, ilo, ihi);
  bounds(error, arg[1], atom->ntypes, jlo, jhi);
  const double a = std::atof(arg[2]), rho = std::atof(arg[3]), c = std::atof(arg[4]);
  if (rho <= 0) error->all(FLERR, "Incorrect args for pair coefficients");
  double cut_lj_one = cut_global, cut_coul_one = cut_coul_global;
  if (narg >= 6) cut_coul_one = cut_lj_one = std::atof(arg[5]);
  if (narg == 7) cut_coul_one = std::atof(arg[6]);
  set_pair(ilo, ihi, jlo, jhi, a, rho, c, cut_lj_one, cut_coul_one);
}

void PairBuckCoulCut::init_style() {
  if (!atom->q_flag) error->all(FLERR, "Pair style buck/coul/cut requires atom attribute q");
}

void PairBuckCoulCutIntel::init_style() {
  PairBuckCoulCut::init_style();
  fix = require_fix_intel();
  init_all_pairs();
  device_setup(fix, B200MD_PAIR_BUCK_COUL_CUT, 0.0, 0.0, 0, nullptr, nullptr);
}

void PairBuckCoulCutIntel::compute(int eflag, int vflag) {
  if (!fix) error->all(FLERR, "Pair style buck/coul/cut/intel used before init_style()");
  device_compute(fix, eflag, vflag);
}

// ---- buck/coul/long ----------------------------------------------------------------------------------------
void PairBuckCoulLong::settings(int narg, char **arg) {
  if (narg < 1 || narg > 2) error->all(FLERR, "Illegal pair_style command");
  cut_global = std::atof(arg[0]);
  cut_coul = narg == 2 ? std::atof(arg[1]) : cut_global;
  if (allocated)
    for (size_t ij = 0; ij < setflag.size(); ij++)
      if (setflag[ij]) k.cut_lj[ij] = cut_global;
}

void PairBuckCoulLong::coeff(int narg, char **arg) {
  if (narg < 5 || narg > 6) error->all(FLERR, "Incorrect args for pair coefficients");
  if (!allocated) allocate();
  int ilo, ihi, jlo, jhi;
  bounds(error, arg[0], atom->ntypes, ilo, ihi);
  bounds(error, arg[1], atom->ntypes, jlo, jhi);
  const double a = std::atof(arg[2]), rho = std::atof(arg[3]), c = std::atof(arg[4]);
  if (rho <= 0) error->all(FLERR, "Incorrect args for pair coefficients");
  const double cut_lj_one = narg == 6 ? std::atof(arg[5]) : cut_global;
  set_pair(ilo, ihi, jlo, jhi, a, rho, c, cut_lj_one, 0.0);
}

double PairBuckCoulLong::init_one(int i, int j) {
  const int ij = i * tp1() + j;
  if (!setflag[ij]) error->all(FLERR, "All pair coeffs are not set");
  k.cut_coul[ij] = cut_coul;   // one global Coulomb cut-off
  return PairBuck::init_one(i, j);
}

void PairBuckCoulLong::init_style() {
  if (!atom->q_flag) error->all(FLERR, "Pair style buck/coul/long requires atom attribute q");
  // "insure use of KSpace long-range solver, set g_ewald" — the kspace style is initialised first
  if (!force->kspace) error->all(FLERR, "Pair style requires a KSpace style");
  g_ewald = force->kspace->g_ewald;                     // pair_buck_coul_long_intel.cpp:507
  ctab = PairTables();
  if (ncoultablebits) init_tables(cut_coul, g_ewald, ctab);
}

void *PairBuckCoulLong::extract(const char *str, int &dim) {
  dim = 0;
  if (std::strcmp(str, "cut_coul") == 0) return &cut_coul;
  return nullptr;
}

void PairBuckCoulLongIntel::init_style() {
  PairBuckCoulLong::init_style();
  fix = require_fix_intel();
  init_all_pairs();
  device_setup(fix, B200MD_PAIR_BUCK_COUL_LONG, g_ewald, 0.0, 0, &ctab, nullptr);
}

void PairBuckCoulLongIntel::compute(int eflag, int vflag) {
  if (!fix) error->all(FLERR, "Pair style buck/coul/long/intel used before init_style()");
  device_compute(fix, eflag, vflag);
}

// ---- buck/long/coul/long -----------------------------------------------------------------------------------
void PairBuckLongCoulLong::settings(int narg, char **arg) {
  if (narg != 3 && narg != 4) error->all(FLERR, "Illegal pair_style command");
  ewald_order = ewald_off = 0;
  auto option = [&](const char *a, int order) {
    if (std::strcmp(a, "long") == 0) ewald_order |= 1 << order;
    else if (std::strcmp(a, "cut") == 0) {}
    else if (std::strcmp(a, "off") == 0) ewald_off |= 1 << order;
    else error->all(FLERR, "Illegal pair_style buck/long/coul/long command");
  };
  option(arg[0], 6);
  option(arg[1], 1);
  // dispersion "cut" = plain Buckingham; Coulomb has no cut flavour in this style
  if (!((ewald_order ^ ewald_off) & (1 << 1)))
    error->all(FLERR, "Coulomb cut not supported in pair_style buck/long/coul/coul");
  cut_global = std::atof(arg[2]);
  cut_coul = narg == 4 ? std::atof(arg[3]) : cut_global;
  ewaldflag = (ewald_order >> 1) & 1;
  dispersionflag = (ewald_order >> 6) & 1;
  if (allocated)
    for (size_t ij = 0; ij < setflag.size(); ij++)
      if (setflag[ij]) k.cut_lj[ij] = cut_global;
}

void PairBuckLongCoulLong::coeff(int narg, char **arg) {
  if (narg < 5 || narg > 6) error->all(FLERR, "Incorrect args for pair coefficients");
  if (!allocated) allocate();
  int ilo, ihi, jlo, jhi;
  bounds(error, arg[0], atom->ntypes, ilo, ihi);
  bounds(error, arg[1], atom->ntypes, jlo, jhi);
  const double a = std::atof(arg[2]), rho = std::atof(arg[3]), c = std::atof(arg[4]);
  if (rho <= 0) error->all(FLERR, "Incorrect args for pair coefficients");
  const double cut_one = narg == 6 ? std::atof(arg[5]) : cut_global;
  set_pair(ilo, ihi, jlo, jhi, a, rho, c, cut_one, 0.0);
}

double PairBuckLongCoulLong::init_one(int i, int j) {
  const int ij = i * tp1() + j;
  if (!setflag[ij]) error->all(FLERR, "All pair coeffs are not set");
  // Coulomb participates only when it is not switched off
  k.cut_coul[ij] = (ewald_off & (1 << 1)) ? 0.0 : cut_coul;
  if (ewald_order & (1 << 6)) k.cut_lj[ij] = cut_global;   // long dispersion: one global cut-off
  return PairBuck::init_one(i, j);
}

void PairBuckLongCoulLong::init_style() {
  if (!atom->q_flag && (ewald_order & (1 << 1)))
    error->all(FLERR, "Invoking coulombic in pair style buck/long/coul/long requires atom attribute q");
  if (ewald_order & ((1 << 1) | (1 << 6))) {
    if (!force->kspace) error->all(FLERR, "Pair style requires a KSpace style");
    g_ewald = force->kspace->g_ewald;
    g_ewald_6 = force->kspace->g_ewald_6;               // pair_buck_long_coul_long_intel.cpp:267
  }
  ctab = PairTables();
  dtab = PairTables();
  if ((ewald_order & (1 << 1)) && ncoultablebits) init_tables(cut_coul, g_ewald, ctab);
  if ((ewald_order & (1 << 6)) && ndisptablebits) init_tables_disp(cut_global, g_ewald_6, dtab);
}

void *PairBuckLongCoulLong::extract(const char *str, int &dim) {
  dim = 0;
  if (std::strcmp(str, "cut_coul") == 0) return &cut_coul;
  if (std::strcmp(str, "ewald_order") == 0) return &ewald_order;
  if (std::strcmp(str, "cut_LJ") == 0) return &cut_global;   // the r^-6 real-space cut-off PPPMDisp sizes g_ewald_6 on
  if (std::strcmp(str, "B") == 0) { dim = 2; return k.c.data(); }   // dispersion coefficients read by pppm/disp
  return nullptr;
}

void PairBuckLongCoulLongIntel::init_style() {
  PairBuckLongCoulLong::init_style();
  fix = require_fix_intel();
  init_all_pairs();
  device_setup(fix, B200MD_PAIR_BUCK_LONG_COUL_LONG, g_ewald, g_ewald_6, ewald_order, &ctab, &dtab);
}

void PairBuckLongCoulLongIntel::compute(int eflag, int vflag) {
  if (!fix) error->all(FLERR, "Pair style buck/long/coul/long/intel used before init_style()");
  device_compute(fix, eflag, vflag);
}
