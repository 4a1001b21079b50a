// pair_lj_long_coul_long_intel.h — PairStyle(lj/long/coul/long/intel,PairLJLongCoulLongIntel) on the device
// (pair_lj_long_coul_long_intel.h:18-46 of the reference).  With `cut long` it is the lj/cut/coul/long of
// examples/in.spce:7, which the driver maps onto it.  The stock base class (settings / coeff / init_one) is restated.
#ifdef PAIR_CLASS

PairStyle(lj/long/coul/long/intel,PairLJLongCoulLongIntel)

#else

#ifndef B200MD_PAIR_LJ_LONG_COUL_LONG_INTEL_H
#define B200MD_PAIR_LJ_LONG_COUL_LONG_INTEL_H
#include "pair_buck_intel.h"

namespace LAMMPS_NS {

// `pair_style lj/long/coul/long flag_lj flag_coul cut_lj [cut_coul]`, `pair_coeff i j epsilon sigma [cut_lj]`
class PairLJLongCoulLong : public PairBuck {
 public:
  explicit PairLJLongCoulLong(LAMMPS *l) : PairBuck(l) {}
  void settings(int narg, char **arg) override;
  void coeff(int narg, char **arg) override;
  void init_style() override;
  double init_one(int i, int j) override;
  void *extract(const char *str, int &dim) override;
  const PairTables *coul_tables() const override { return ctab.nbits ? &ctab : nullptr; }
  const PairTables *disp_tables() const override { return dtab.nbits ? &dtab : nullptr; }

 protected:
  int ewald_order = 0, ewald_off = 0;   // bit1 = long Coulomb, bit6 = long dispersion (pair_lj_long_coul_long_intel.cpp:111-112)
  double cut_coul = 0.0, g_ewald = 0.0, g_ewald_6 = 0.0;
  std::vector<double> epsilon, sigma;   // as given by pair_coeff; k.buck1, k.buck2, k.a, k.c receive lj1..lj4
  PairTables ctab, dtab;
};

class PairLJLongCoulLongIntel : public PairLJLongCoulLong {
 public:
  explicit PairLJLongCoulLongIntel(LAMMPS *l) : PairLJLongCoulLong(l) { suffix_flag |= Suffix::INTEL; }
  void compute(int eflag, int vflag) override;
  void init_style() override;

 private:
  FixIntel *fix = nullptr;
};

}  // namespace LAMMPS_NS

#endif
#endif
