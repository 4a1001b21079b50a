// pair_buck_coul_long_intel.cpp, B200 build — drops into a LAMMPS tree IN PLACE OF the reference's file of the same name.
// It implements the class that the reference's own header declares (pair_buck_coul_long_intel.h:18-93, used unchanged):
// the bodies of init_style / pack_force_const (:455-566) and compute / eval<> (:55-453) become two C-ABI calls.
// tests/test_host.py::test_integration_binding_compiles_against_the_reference_header compiles this file against the
// reference's header and stand-ins of the stock LAMMPS headers (the ones the reference's own sources compile against).
#include "pair_buck_coul_long_intel.h"

#include "b200_pair_binding.h"

using namespace LAMMPS_NS;

PairBuckCoulLongIntel::PairBuckCoulLongIntel(LAMMPS *lmp) : PairBuckCoulLong(lmp) { suffix_flag |= Suffix::INTEL; }

PairBuckCoulLongIntel::~PairBuckCoulLongIntel() {}

void PairBuckCoulLongIntel::init_style() {
  PairBuckCoulLong::init_style();   // g_ewald from force->kspace, Pair::init_tables (:507, :531-542 read their products)
  B200_FIND_FIX_INTEL();
  B200_INIT_ALL_PAIRS();
  const int tp1 = atom->ntypes + 1;
  std::vector<double> cc((size_t)tp1 * tp1, cut_coulsq);   // one global Coulomb cut-off in this style
  b200md_pair_params p = b200md_pair_params();
  p.style = B200MD_PAIR_BUCK_COUL_LONG;
  p.ntypes = atom->ntypes;
  p.cutsq = &cutsq[0][0];           // memory->create storage is contiguous: row-major (ntypes+1)^2, what the C ABI takes
  p.cut_ljsq = &cut_ljsq[0][0];
  p.cut_coulsq = cc.data();
  p.buck1 = &buck1[0][0]; p.buck2 = &buck2[0][0]; p.rhoinv = &rhoinv[0][0];
  p.a = &a[0][0]; p.c = &c[0][0]; p.offset = &offset[0][0];
  B200_PACK_SPECIAL(p);
  p.g_ewald = force->kspace->g_ewald;
  if (ncoultablebits) B200_PACK_COUL_TABLES(p);
  B200_PAIR_SETUP(p);
}

// ev_global of the reference: evdwl, ecoul, v_xx, v_yy, v_zz, v_xy, v_xz, v_yz (:337-349)
void PairBuckCoulLongIntel::compute(int eflag, int vflag) { B200_PAIR_COMPUTE(eflag, vflag); }

// the reference keeps per-precision copies of the coefficients in ForceConst (:573-646); here they live on the device
// behind b200md_pair_setup, so the members the header declares stay empty
template <class flt_t>
void PairBuckCoulLongIntel::ForceConst<flt_t>::set_ntypes(const int, const int, Memory *, const int) {}
template void PairBuckCoulLongIntel::ForceConst<float>::set_ntypes(const int, const int, Memory *, const int);
template void PairBuckCoulLongIntel::ForceConst<double>::set_ntypes(const int, const int, Memory *, const int);
