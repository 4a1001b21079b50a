// pair_buck_long_coul_long_intel.cpp, B200 build — in place of the reference's file: the class of the reference's own
// header (pair_buck_long_coul_long_intel.h:33-107, unchanged) with init_style / pack_force_const (:541-646) and
// compute / eval<EVFLAG,EFLAG,NEWTON_PAIR,ORDER1,ORDER6,...> (:57-539) as two C-ABI calls; ORDER1 / ORDER6 travel as
// `ewald_order` (bit 1: long Coulomb, bit 6: long dispersion).  Compile-checked against that header by tests/test_host.py.
#include "pair_buck_long_coul_long_intel.h"

#include "b200_pair_binding.h"

using namespace LAMMPS_NS;

PairBuckLongCoulLongIntel::PairBuckLongCoulLongIntel(LAMMPS *lmp) : PairBuckLongCoulLong(lmp) {
  suffix_flag |= Suffix::INTEL;
}

PairBuckLongCoulLongIntel::~PairBuckLongCoulLongIntel() {}

void PairBuckLongCoulLongIntel::init_style() {
  PairBuckLongCoulLong::init_style();   // g_ewald, g_ewald_6 from force->kspace (:267), Pair::init_tables(_disp)
  B200_FIND_FIX_INTEL();
  B200_INIT_ALL_PAIRS();
  const int tp1 = atom->ntypes + 1;
  // Coulomb participates unless it is switched off (`long off`: ewald_off bit 1)
  std::vector<double> cc((size_t)tp1 * tp1, (ewald_off & (1 << 1)) ? 0.0 : cut_coulsq);
  b200md_pair_params p = b200md_pair_params();
  p.style = B200MD_PAIR_BUCK_LONG_COUL_LONG;
  p.ntypes = atom->ntypes;
  p.cutsq = &cutsq[0][0];
  p.cut_ljsq = &cut_bucksq[0][0];
  p.cut_coulsq = cc.data();
  p.buck1 = &buck1[0][0]; p.buck2 = &buck2[0][0]; p.rhoinv = &rhoinv[0][0];
  p.a = &buck_a[0][0]; p.c = &buck_c[0][0]; p.offset = &offset[0][0];
  B200_PACK_SPECIAL(p);
  p.g_ewald = force->kspace->g_ewald;
  p.g_ewald_6 = force->kspace->g_ewald_6;
  p.ewald_order = ewald_order;
  if ((ewald_order & (1 << 1)) && ncoultablebits) B200_PACK_COUL_TABLES(p);
  if ((ewald_order & (1 << 6)) && ndisptablebits) B200_PACK_DISP_TABLES(p);
  B200_PAIR_SETUP(p);
}

void PairBuckLongCoulLongIntel::compute(int eflag, int vflag) { B200_PAIR_COMPUTE(eflag, vflag); }

template <class flt_t>
void PairBuckLongCoulLongIntel::ForceConst<flt_t>::set_ntypes(const int, const int, Memory *) {}
template void PairBuckLongCoulLongIntel::ForceConst<float>::set_ntypes(const int, const int, Memory *);
template void PairBuckLongCoulLongIntel::ForceConst<double>::set_ntypes(const int, const int, Memory *);
