// pair_buck_coul_long_intel.cpp, B200 build — drops into a LAMMPS tree IN PLACE OF the reference's file of the same name.
// It implements the class that the reference's own header declares (pair_buck_coul_long_intel.h:18-93, used unchanged):
// the bodies of init_style / pack_force_const (:455-566) and compute / eval<> (:55-453) become two C-ABI calls.
// tests/test_host.py::test_integration_binding_compiles_against_the_reference_header compiles this file against the
// reference's header and stand-ins of the stock LAMMPS headers (the ones the reference's own sources compile against).
#include "pair_buck_coul_long_intel.h"

#include <vector>

#include "atom.h"
#include "error.h"
#include "force.h"
#include "kspace.h"
#include "modify.h"
#include "suffix.h"

#include "b200_fix_intel.h"

using namespace LAMMPS_NS;

PairBuckCoulLongIntel::PairBuckCoulLongIntel(LAMMPS *lmp) : PairBuckCoulLong(lmp) {
  suffix_flag |= Suffix::INTEL;
}

PairBuckCoulLongIntel::~PairBuckCoulLongIntel() {}

void PairBuckCoulLongIntel::init_style() {
  PairBuckCoulLong::init_style();   // g_ewald from force->kspace, Pair::init_tables (:507, :531-542 read their products)
  const int ifix = modify->find_fix("package_intel");
  if (ifix < 0) error->all(FLERR, "The 'package intel' command is required for /intel styles");
  fix = static_cast<FixIntel *>(modify->fix[ifix]);
  fix->pair_init_check();

  // pack_force_const repeats init_one for every type pair (:497-506) before it copies the coefficients
  const int tp1 = atom->ntypes + 1;
  for (int i = 1; i < tp1; i++)
    for (int j = i; j < tp1; j++)
      if (setflag[i][j] != 0 || (setflag[i][i] != 0 && setflag[j][j] != 0)) {
        const double cut = init_one(i, j);
        cutsq[i][j] = cutsq[j][i] = cut * cut;
      }

  b200md_pair_params p = b200md_pair_params();
  p.style = B200MD_PAIR_BUCK_COUL_LONG;
  p.ntypes = atom->ntypes;
  // memory->create gives contiguous [tp1][tp1] storage: &a[0][0] is the row-major table the C ABI takes
  p.cutsq = &cutsq[0][0];
  p.cut_ljsq = &cut_ljsq[0][0];
  std::vector<double> cc((size_t)tp1 * tp1, cut_coulsq);   // one global Coulomb cut-off in this style
  p.cut_coulsq = cc.data();
  p.buck1 = &buck1[0][0];
  p.buck2 = &buck2[0][0];
  p.rhoinv = &rhoinv[0][0];
  p.a = &a[0][0];
  p.c = &c[0][0];
  p.offset = &offset[0][0];
  for (int k = 0; k < 4; k++) {
    p.special_lj[k] = force->special_lj[k];
    p.special_coul[k] = force->special_coul[k];
  }
  p.g_ewald = force->kspace->g_ewald;
  p.ncoultablebits = ncoultablebits;
  p.ncoulmask = ncoulmask;
  p.ncoulshiftbits = ncoulshiftbits;
  p.tabinnersq = tabinnersq;
  p.rtable = rtable; p.drtable = drtable; p.ftable = ftable; p.dftable = dftable;
  p.etable = etable; p.detable = detable; p.ctable = ctable; p.dctable = dctable;
  if (b200md_pair_setup(b200_ctx(fix), &p)) error->all(FLERR, b200md_last_error(b200_ctx(fix)));
}

void PairBuckCoulLongIntel::compute(int eflag, int vflag) {
  if (eflag || vflag) ev_setup(eflag, vflag);
  else evflag = vflag_fdotr = 0;
  b200_positions_to_device(fix);
  double ev[8];   // ev_global of the reference: evdwl, ecoul, v_xx, v_yy, v_zz, v_xy, v_xz, v_yz (:337-349)
  if (b200md_pair_compute(b200_ctx(fix), eflag, vflag, ev)) error->one(FLERR, b200md_last_error(b200_ctx(fix)));
  if (eflag_global) {
    eng_vdwl += ev[0];
    eng_coul += ev[1];
  }
  if (vflag_global)
    for (int n = 0; n < 6; n++) virial[n] += ev[2 + n];
  b200_forces_to_host(fix);
}

// the reference keeps per-precision copies of the coefficients in ForceConst (:573-646); here they live on the device
// behind b200md_pair_setup, so the members the header declares stay empty
template <class flt_t>
void PairBuckCoulLongIntel::ForceConst<flt_t>::set_ntypes(const int, const int, Memory *, const int) {}
template void PairBuckCoulLongIntel::ForceConst<float>::set_ntypes(const int, const int, Memory *, const int);
template void PairBuckCoulLongIntel::ForceConst<double>::set_ntypes(const int, const int, Memory *, const int);
