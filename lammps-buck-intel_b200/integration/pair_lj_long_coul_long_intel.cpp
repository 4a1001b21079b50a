// pair_lj_long_coul_long_intel.cpp, B200 build — in place of the reference's file: the class of the reference's own
// header (pair_lj_long_coul_long_intel.h:33-107, unchanged) with init_style / pack_force_const (:752-867) and compute /
// eval<..., ORDER1, ORDER6, DISPTABLE, COULTABLE> (:57-747) as two C-ABI calls.  lj1..lj4 travel in buck1, buck2, a, c
// (include/b200md.h).  The dispersion tables are handed over whenever the dispersion sum is long-ranged and
// `table/disp` is on — the reference copies them only `if (ncoultablebits)` (:839-860), which leaves them unset for the
// `long off` of examples/in.hexane.  Compile-checked against that header by tests/test_host.py.
#include "pair_lj_long_coul_long_intel.h"

#include "b200_pair_binding.h"

using namespace LAMMPS_NS;

PairLJLongCoulLongIntel::PairLJLongCoulLongIntel(LAMMPS *lmp) : PairLJLongCoulLong(lmp) {
  suffix_flag |= Suffix::INTEL;
}

PairLJLongCoulLongIntel::~PairLJLongCoulLongIntel() {}

void PairLJLongCoulLongIntel::init_style() {
  PairLJLongCoulLong::init_style();   // g_ewald, g_ewald_6 from force->kspace (:479), Pair::init_tables(_disp)
  B200_FIND_FIX_INTEL();
  B200_INIT_ALL_PAIRS();
  const int tp1 = atom->ntypes + 1;
  std::vector<double> cc((size_t)tp1 * tp1, (ewald_off & (1 << 1)) ? 0.0 : cut_coulsq);
  std::vector<double> zero((size_t)tp1 * tp1, 0.0);
  b200md_pair_params p = b200md_pair_params();
  p.style = B200MD_PAIR_LJ_LONG_COUL_LONG;
  p.ntypes = atom->ntypes;
  p.cutsq = &cutsq[0][0];
  p.cut_ljsq = &cut_ljsq[0][0];
  p.cut_coulsq = cc.data();
  p.buck1 = &lj1[0][0]; p.buck2 = &lj2[0][0]; p.a = &lj3[0][0]; p.c = &lj4[0][0];
  p.rhoinv = zero.data();
  p.offset = &offset[0][0];
  B200_PACK_SPECIAL(p);
  p.g_ewald = force->kspace->g_ewald;
  p.g_ewald_6 = force->kspace->g_ewald_6;
  p.ewald_order = ewald_order;
  if ((ewald_order & (1 << 1)) && ncoultablebits) B200_PACK_COUL_TABLES(p);
  if ((ewald_order & (1 << 6)) && ndisptablebits) B200_PACK_DISP_TABLES(p);
  B200_PAIR_SETUP(p);
}

void PairLJLongCoulLongIntel::compute(int eflag, int vflag) { B200_PAIR_COMPUTE(eflag, vflag); }

template <class flt_t>
void PairLJLongCoulLongIntel::ForceConst<flt_t>::set_ntypes(const int, const int, Memory *) {}
template void PairLJLongCoulLongIntel::ForceConst<float>::set_ntypes(const int, const int, Memory *);
template void PairLJLongCoulLongIntel::ForceConst<double>::set_ntypes(const int, const int, Memory *);
