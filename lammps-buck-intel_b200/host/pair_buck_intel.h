// pair_buck_intel.h — PairStyle(buck/intel,PairBuckIntel) on the device.
// Mirrors pair_buck_intel.h:32-38 of the reference: same class name, same public surface
// (`compute(int,int)`, `init_style()`), `settings/coeff/init_one` from the stock base class PairBuck
// (restated here because the reference does not ship it).
#ifdef PAIR_CLASS

PairStyle(buck/intel,PairBuckIntel)

#else

#ifndef B200MD_PAIR_BUCK_INTEL_H
#define B200MD_PAIR_BUCK_INTEL_H
#include "fix_intel.h"
#include "lammps_shim.h"

namespace LAMMPS_NS {

// per-type-pair tables every Buckingham style keeps (PairBuck::allocate)
struct BuckCoeffs {
  std::vector<double> a, rho, c, cut_lj, cut_coul;          // as given by pair_coeff
  std::vector<double> rhoinv, buck1, buck2, offset, cut_ljsq, cut_coulsq;  // init_one products
  void allocate(int n) {
    for (auto *v : {&a, &rho, &c, &cut_lj, &cut_coul, &rhoinv, &buck1, &buck2, &offset, &cut_ljsq, &cut_coulsq})
      v->assign((size_t)n * n, 0.0);
  }
};

// Coulomb / dispersion lookup tables built by Pair::init_tables / init_tables_disp (App. A.2)
struct PairTables {
  int nbits = 0, mask = 0, shiftbits = 0;
  double tabinnersq = 0.0;
  std::vector<double> r, dr, f, df, e, de, c, dc;
};

// stock PairBuck: `pair_style buck cut`, `pair_coeff i j A rho C [cut]`
class PairBuck : public Pair {
 public:
  explicit PairBuck(LAMMPS *l) : Pair(l) {}
  void settings(int narg, char **arg) override;
  void coeff(int narg, char **arg) override;
  void init_style() override {}
  double init_one(int i, int j) override;
  // the lookup tables init_style built (nullptr: none) — read by `lmp_b200 -dry-run`, which prints their CRC-32 so that
  // the harness-side builders (lammps-buck-intel_b200/__init__.py) can be held bit-identical to these on the CPU
  virtual const PairTables *coul_tables() const { return nullptr; }
  virtual const PairTables *disp_tables() const { return nullptr; }

 protected:
  double cut_global = 0.0;
  BuckCoeffs k;
  void allocate();
  void set_pair(int ilo, int ihi, int jlo, int jhi, double a, double rho, double c, double cut_lj, double cut_coul);
  // shared by the four /intel classes: pack_force_const -> b200md_pair_setup, eval<> -> b200md_pair_compute
  void device_setup(FixIntel *fix, int style, double g_ewald, double g_ewald_6, int ewald_order,
                    const PairTables *ctab, const PairTables *dtab);
  void device_compute(FixIntel *fix, int eflag, int vflag);
  FixIntel *require_fix_intel();
  void init_all_pairs() override;   // pack_force_const repeats init_one for every type pair (pair_buck_intel.cpp:399-409)
  static void bounds(Error *error, const char *str, int nmax, int &nlo, int &nhi);
  static void init_bitmap(double inner, double outer, int ntablebits, int &masklo, int &maskhi, int &nmask,
                          int &nshiftbits);
  void init_tables(double cut_coul, double g_ewald, PairTables &t) const;
  void init_tables_disp(double cut_lj_global, double g_ewald_6, PairTables &t) const;
};

class PairBuckIntel : public PairBuck {
 public:
  explicit PairBuckIntel(LAMMPS *l) : PairBuck(l) { suffix_flag |= Suffix::INTEL; }
  void compute(int eflag, int vflag) override;
  void init_style() override;

 private:
  FixIntel *fix = nullptr;
};

}  // namespace LAMMPS_NS

#endif
#endif
