// pair_buck_coul_intel.h — the three charged Buckingham styles on the device:
//   PairStyle(buck/coul/cut/intel,PairBuckCoulCutIntel)            pair_buck_coul_cut_intel.h:19-40
//   PairStyle(buck/coul/long/intel,PairBuckCoulLongIntel)          pair_buck_coul_long_intel.h:18-40
//   PairStyle(buck/long/coul/long/intel,PairBuckLongCoulLongIntel) pair_buck_long_coul_long_intel.h:18-46
// each over its stock base class (settings / coeff / init_one), restated here.
#ifdef PAIR_CLASS

PairStyle(buck/coul/cut/intel,PairBuckCoulCutIntel)
PairStyle(buck/coul/long/intel,PairBuckCoulLongIntel)
PairStyle(buck/long/coul/long/intel,PairBuckLongCoulLongIntel)

#else

#ifndef B200MD_PAIR_BUCK_COUL_INTEL_H
#define B200MD_PAIR_BUCK_COUL_INTEL_H
#include "pair_buck_intel.h"

namespace LAMMPS_NS {

// `pair_style buck/coul/cut cut_lj [cut_coul]`, `pair_coeff i j A rho C [cut_lj [cut_coul]]`
class PairBuckCoulCut : public PairBuck {
 public:
  explicit PairBuckCoulCut(LAMMPS *l) : PairBuck(l) {}
  void settings(int narg, char **arg) override;
  void coeff(int narg, char **arg) override;
  void init_style() override;

 protected:
  double cut_coul_global = 0.0;
};

class PairBuckCoulCutIntel : public PairBuckCoulCut {
 public:
  explicit PairBuckCoulCutIntel(LAMMPS *l) : PairBuckCoulCut(l) { suffix_flag |= Suffix::INTEL; }
  void compute(int eflag, int vflag) override;
  void init_style() override;

 private:
  FixIntel *fix = nullptr;
};

// `pair_style buck/coul/long cut_lj [cut_coul]`, `pair_coeff i j A rho C [cut_lj]`
class PairBuckCoulLong : public PairBuck {
 public:
  explicit PairBuckCoulLong(LAMMPS *l) : PairBuck(l) { ewaldflag = 1; }
  void settings(int narg, char **arg) override;
  void coeff(int narg, char **arg) override;
  void init_style() override;
  double init_one(int i, int j) override;
  void *extract(const char *str, int &dim) override;
  const PairTables *coul_tables() const override { return ctab.nbits ? &ctab : nullptr; }

 protected:
  double cut_coul = 0.0, g_ewald = 0.0;
  PairTables ctab;
};

class PairBuckCoulLongIntel : public PairBuckCoulLong {
 public:
  explicit PairBuckCoulLongIntel(LAMMPS *l) : PairBuckCoulLong(l) { suffix_flag |= Suffix::INTEL; }
  void compute(int eflag, int vflag) override;
  void init_style() override;

 private:
  FixIntel *fix = nullptr;
};

// `pair_style buck/long/coul/long flag_buck flag_coul cut_buck [cut_coul]` (flags long|cut|off)
class PairBuckLongCoulLong : public PairBuck {
 public:
  explicit PairBuckLongCoulLong(LAMMPS *l) : PairBuck(l) {}
  void settings(int narg, char **arg) override;
  void coeff(int narg, char **arg) override;
  void init_style() override;
  double init_one(int i, int j) override;
  void *extract(const char *str, int &dim) override;
  const PairTables *coul_tables() const override { return ctab.nbits ? &ctab : nullptr; }
  const PairTables *disp_tables() const override { return dtab.nbits ? &dtab : nullptr; }

 protected:
  int ewald_order = 0, ewald_off = 0;   // bit1 = long Coulomb, bit6 = long dispersion (pair_buck_long_coul_long_intel.cpp:111-112)
  double cut_coul = 0.0, g_ewald = 0.0, g_ewald_6 = 0.0;
  PairTables ctab, dtab;
};

class PairBuckLongCoulLongIntel : public PairBuckLongCoulLong {
 public:
  explicit PairBuckLongCoulLongIntel(LAMMPS *l) : PairBuckLongCoulLong(l) { suffix_flag |= Suffix::INTEL; }
  void compute(int eflag, int vflag) override;
  void init_style() override;

 private:
  FixIntel *fix = nullptr;
};

}  // namespace LAMMPS_NS

#endif
#endif
