// pppm_intel.cpp, B200 build — drops into a LAMMPS tree IN PLACE OF the reference's file of the same name and
// implements the class its header declares (pppm_intel.h:33-39, used unchanged): PPPMIntel::init (:67-98) hands the
// mesh the stock base class sized to the device, PPPMIntel::compute (:104-317: particle_map, make_rho, brick2fft,
// poisson_ik / poisson_ad, fieldforce_ik / fieldforce_ad, the energy / virial post-factors) is one C-ABI call.
// Compile-checked against the reference's header by tests/test_host.py (see pair_buck_coul_long_intel.cpp here).
#include "pppm_intel.h"

#include <cstring>

#include "atom.h"
#include "error.h"
#include "force.h"
#include "modify.h"
#include "suffix.h"

#include "b200_fix_intel.h"

using namespace LAMMPS_NS;

#define INTEL_P3M_MAXORDER 7   // the reference's limit (pppm_intel.cpp:87): the device kernels cover orders 2..7 too

PPPMIntel::PPPMIntel(LAMMPS *lmp, int narg, char **arg) : PPPM(lmp, narg, arg) {
  suffix_flag |= Suffix::INTEL;
}

PPPMIntel::~PPPMIntel() {}

void PPPMIntel::init() {
  PPPM::init();   // qsum_qsq, set_grid_global, adjust_gewald: nx/ny/nz_pppm, order, g_ewald (stock, untouched)
  const int ifix = modify->find_fix("package_intel");
  if (ifix < 0) error->all(FLERR, "The 'package intel' command is required for /intel styles");
  fix = static_cast<FixIntel *>(modify->fix[ifix]);
  if (order > INTEL_P3M_MAXORDER) error->all(FLERR, "PPPM order greater than supported by USER-INTEL");

  b200md_pppm_params p;
  std::memset(&p, 0, sizeof(p));
  p.nx = nx_pppm;
  p.ny = ny_pppm;
  p.nz = nz_pppm;
  p.order = order;
  p.g_ewald = g_ewald;
  p.differentiation = differentiation_flag;
  p.scale = scale;
  p.slab_volfactor = slabflag ? slab_volfactor : 0.0;
  if (b200md_pppm_setup(b200_ctx(fix), &p)) error->all(FLERR, b200md_last_error(b200_ctx(fix)));
}

void PPPMIntel::compute(int eflag, int vflag) {
  if (eflag || vflag) ev_setup(eflag, vflag);
  else evflag = evflag_atom = eflag_global = vflag_global = eflag_atom = vflag_atom = 0;
  // the pair style of this step has moved the positions to the device already (kspace runs after pair in Verlet::run);
  // a k-space-only model moves them here
  if (!force->pair) b200_positions_to_device(fix);
  double e = 0.0, v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (b200md_pppm_compute(b200_ctx(fix), eflag, vflag, &e, v)) error->one(FLERR, b200md_last_error(b200_ctx(fix)));
  if (eflag_global) energy = e;          // qscale, the g_ewald self term and the volume term are applied on the device
  if (vflag_global)
    for (int n = 0; n < 6; n++) virial[n] = v[n];
  b200_forces_to_host(fix);              // f += (pppm_intel.cpp:628-630)
}

// brick2fft (:642-672) is fused into the density fold on the device; the virtual stays for callers of the stock interface
void PPPMIntel::brick2fft() {}
