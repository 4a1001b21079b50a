// pppm_disp_intel.cpp, B200 build — in place of the reference's file: the class of the reference's own header
// (pppm_disp_intel.h:33-40, unchanged).  PPPMDispIntel::init (:86-109) hands the two meshes the stock base class sized —
// Coulomb, and the dispersion mesh of whichever "function" the mixing rule selected — to the device;
// PPPMDispIntel::compute (:115-554: the Coulomb branch :183-243, geometric :245-313, arithmetic :315-407, no mixing
// :409-467, the energy / virial sums :470-541) is one C-ABI call that runs every mesh set up.
// Compile-checked against the reference's header and a stand-in of stock pppm_disp.h by tests/test_host.py.
#include "pppm_disp_intel.h"

#include <cstring>

#include "atom.h"
#include "error.h"
#include "force.h"
#include "modify.h"
#include "pair.h"
#include "suffix.h"

#include "b200_fix_intel.h"

using namespace LAMMPS_NS;

PPPMDispIntel::PPPMDispIntel(LAMMPS *lmp, int narg, char **arg) : PPPMDisp(lmp, narg, arg) {
  suffix_flag |= Suffix::INTEL;
}

PPPMDispIntel::~PPPMDispIntel() {}

void PPPMDispIntel::init() {
  PPPMDisp::init();   // stock: function[0..3], set_grid / set_grid_6 (both meshes, g_ewald, g_ewald_6), init_coeffs (B)
  const int ifix = modify->find_fix("package_intel");
  if (ifix < 0) error->all(FLERR, "The 'package intel' command is required for /intel styles");
  fix = static_cast<FixIntel *>(modify->fix[ifix]);
  fix->kspace_init_check();
  if (order > INTEL_P3M_MAXORDER || order_6 > INTEL_P3M_MAXORDER)
    error->all(FLERR, "PPPM order greater than supported by USER-INTEL\n");

  b200md_pppm_params p;
  if (function[0]) {   // the Coulomb mesh ('c')
    std::memset(&p, 0, sizeof(p));
    p.nx = nx_pppm; p.ny = ny_pppm; p.nz = nz_pppm;
    p.order = order;
    p.g_ewald = g_ewald;
    p.differentiation = differentiation_flag;
    p.scale = scale;
    if (b200md_pppm_setup(b200_ctx(fix), &p)) error->all(FLERR, b200md_last_error(b200_ctx(fix)));
  }
  const int rule = function[1] ? 1 : (function[2] ? 2 : (function[3] ? 3 : 0));
  if (rule) {          // the dispersion mesh: geometric ('g'), arithmetic (seven grids) or no mixing rule
    std::memset(&p, 0, sizeof(p));
    p.nx = nx_pppm_6; p.ny = ny_pppm_6; p.nz = nz_pppm_6;
    p.order = order_6;
    p.g_ewald = g_ewald_6;
    p.differentiation = differentiation_flag;
    p.scale = 1.0;
    p.dispersion = rule;
    if (rule == 3) {   // the C_ij matrix itself, as the pair style holds it; the eigen-split happens on the other side
      int dim = 0;
      double **cij = (double **)force->pair->extract("B", dim);
      if (!cij || dim != 2) error->all(FLERR, "KSpace style is incompatible with Pair style");
      p.B = &cij[0][0];
    } else p.B = B;    // init_coeffs: B[type] (geometric) or B[7 type + k] (arithmetic)
    if (b200md_pppm_setup(b200_ctx(fix), &p)) error->all(FLERR, b200md_last_error(b200_ctx(fix)));
  }
}

void PPPMDispIntel::compute(int eflag, int vflag) {
  if (eflag || vflag) ev_setup(eflag, vflag);
  else evflag = evflag_atom = eflag_global = vflag_global = eflag_atom = vflag_atom = 0;
  if (!force->pair) b200_positions_to_device(fix);   // otherwise the pair style of this step has moved them already
  double e = 0.0, v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (b200md_pppm_compute(b200_ctx(fix), eflag, vflag, &e, v)) error->one(FLERR, b200md_last_error(b200_ctx(fix)));
  if (eflag_global) energy = e;                      // energy_1 + energy_6 with all self and volume terms (:470-541)
  if (vflag_global)
    for (int n = 0; n < 6; n++) virial[n] = v[n];
  b200_forces_to_host(fix);
}
