// pair_buck_coul_cut_intel.cpp, B200 build — in place of the reference's file: the class of the reference's own header
// (pair_buck_coul_cut_intel.h:33-93, unchanged) with init_style / pack_force_const (:404-492) and compute / eval<>
// (:55-402) as two C-ABI calls.  Compile-checked against that header by tests/test_host.py.
#include "pair_buck_coul_cut_intel.h"

#include "b200_pair_binding.h"

using namespace LAMMPS_NS;

PairBuckCoulCutIntel::PairBuckCoulCutIntel(LAMMPS *lmp) : PairBuckCoulCut(lmp) { suffix_flag |= Suffix::INTEL; }

PairBuckCoulCutIntel::~PairBuckCoulCutIntel() {}

void PairBuckCoulCutIntel::init_style() {
  PairBuckCoulCut::init_style();
  B200_FIND_FIX_INTEL();
  B200_INIT_ALL_PAIRS();
  b200md_pair_params p = b200md_pair_params();
  p.style = B200MD_PAIR_BUCK_COUL_CUT;
  p.ntypes = atom->ntypes;
  p.cutsq = &cutsq[0][0];
  p.cut_ljsq = &cut_ljsq[0][0];
  p.cut_coulsq = &cut_coulsq[0][0];   // per-pair Coulomb cut-offs (`pair_coeff i j A rho C cut_lj cut_coul`)
  p.buck1 = &buck1[0][0]; p.buck2 = &buck2[0][0]; p.rhoinv = &rhoinv[0][0];
  p.a = &a[0][0]; p.c = &c[0][0]; p.offset = &offset[0][0];
  B200_PACK_SPECIAL(p);
  B200_PAIR_SETUP(p);
}

void PairBuckCoulCutIntel::compute(int eflag, int vflag) { B200_PAIR_COMPUTE(eflag, vflag); }

template <class flt_t>
void PairBuckCoulCutIntel::ForceConst<flt_t>::set_ntypes(const int, const int, Memory *, const int) {}
template void PairBuckCoulCutIntel::ForceConst<float>::set_ntypes(const int, const int, Memory *, const int);
template void PairBuckCoulCutIntel::ForceConst<double>::set_ntypes(const int, const int, Memory *, const int);
