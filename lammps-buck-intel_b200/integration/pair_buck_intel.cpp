// pair_buck_intel.cpp, B200 build — in place of the reference's file: the class of the reference's own header
// (pair_buck_intel.h:32-80, unchanged) with init_style / pack_force_const (:367-443) and compute / eval<> (:48-365) as
// two C-ABI calls.  Compile-checked against that header by tests/test_host.py.
#include "pair_buck_intel.h"

#include "b200_pair_binding.h"

using namespace LAMMPS_NS;

PairBuckIntel::PairBuckIntel(LAMMPS *lmp) : PairBuck(lmp) { suffix_flag |= Suffix::INTEL; }

PairBuckIntel::~PairBuckIntel() {}

void PairBuckIntel::init_style() {
  PairBuck::init_style();
  B200_FIND_FIX_INTEL();
  B200_INIT_ALL_PAIRS();
  const int tp1 = atom->ntypes + 1;
  std::vector<double> cut_ljsq((size_t)tp1 * tp1, 0.0), zero((size_t)tp1 * tp1, 0.0);
  for (int i = 1; i < tp1; i++)
    for (int j = 1; j < tp1; j++) cut_ljsq[(size_t)i * tp1 + j] = cutsq[i][j];   // one cut-off per pair in this style
  b200md_pair_params p = b200md_pair_params();
  p.style = B200MD_PAIR_BUCK;
  p.ntypes = atom->ntypes;
  p.cutsq = &cutsq[0][0];          // memory->create storage is contiguous: row-major (ntypes+1)^2
  p.cut_ljsq = cut_ljsq.data();
  p.cut_coulsq = zero.data();
  p.buck1 = &buck1[0][0]; p.buck2 = &buck2[0][0]; p.rhoinv = &rhoinv[0][0];
  p.a = &a[0][0]; p.c = &c[0][0]; p.offset = &offset[0][0];
  B200_PACK_SPECIAL(p);
  B200_PAIR_SETUP(p);
}

void PairBuckIntel::compute(int eflag, int vflag) { B200_PAIR_COMPUTE(eflag, vflag); }

// the per-precision coefficient copies of the reference (:447-496) live on the device behind b200md_pair_setup
template <class flt_t>
void PairBuckIntel::ForceConst<flt_t>::set_ntypes(const int, Memory *, const int) {}
template void PairBuckIntel::ForceConst<float>::set_ntypes(const int, Memory *, const int);
template void PairBuckIntel::ForceConst<double>::set_ntypes(const int, Memory *, const int);
