// pppm_disp_intel.cpp — host side of pppm/disp/intel.
//   PPPMDispIntel::init     pppm_disp_intel.cpp:86-109
//   PPPMDispIntel::compute  :115-554: Coulomb branch :183-243 and geometric branch :245-313 (particle_map<'c'|'g'>
//                           :556-630, make_rho<'c'|'g'> :633-784 with the per-atom weight B[type]),
//                           energy/virial post-factors :470-510                         -> b200md_pppm_compute
// init_coeffs / the real-space accuracy estimate restate the stock base class PPPMDisp.
#include "pppm_disp_intel.h"

#include <cmath>
#include <cstring>

using namespace LAMMPS_NS;

static const double MY_PI = 3.14159265358979323846;

double PPPMDispIntel::lj_rspace_error(double g6) const {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2];
  double rgs = cutoff_lj * g6;
  rgs *= rgs;
  const double rgs_inv = 1.0 / rgs;
  return csum / std::sqrt((double)atom->natoms * xprd * yprd * zprd * cutoff_lj) * std::sqrt(MY_PI) * std::pow(g6, 5.0) *
         std::exp(-rgs) * (1.0 + rgs_inv * (3.0 + rgs_inv * (6.0 + rgs_inv * 6.0)));
}

void PPPMDispIntel::init() {
  if (domain->triclinic) error->all(FLERR, "Cannot (yet) use PPPMDisp with triclinic box and this build");
  for (int d = 0; d < 3; d++)
    if (!domain->periodicity[d]) error->all(FLERR, "Cannot use nonperiodic boundaries with PPPMDisp");
  if (!force->pair) error->all(FLERR, "KSpace style is incompatible with Pair style");
  int itmp;
  int *p_order = (int *)force->pair->extract("ewald_order", itmp);
  double *p_cutoff = (double *)force->pair->extract("cut_coul", itmp);
  double *b = (double *)force->pair->extract("B", itmp);
  if (!p_order || !p_cutoff) error->all(FLERR, "KSpace style is incompatible with Pair style");
  const int ewald_order = *p_order;
  function[0] = (ewald_order >> 1) & 1;
  function[1] = (ewald_order >> 6) & 1;   // buck/long/coul/long mixes geometrically (ewald_mix = GEOMETRIC)
  if (!function[0] && !function[1]) error->all(FLERR, "PPPMDisp used but no parameters set, for full pppm use pppm");
  if (order_6 > 7 || order > 7) error->all(FLERR, "PPPM order greater than supported by USER-INTEL");

  if (function[0]) PPPM::init();   // Coulomb grid: qsum_qsq, set_grid_global, adjust_gewald

  if (function[1]) {
    if (!b) error->all(FLERR, "KSpace style is incompatible with Pair style");
    const int n = atom->ntypes + 1;
    B.assign(n, 0.0);
    for (int i = 1; i < n; i++) B[i] = std::sqrt(std::fabs(b[i * n + i]));   // PPPMDisp::init_coeffs, geometric
    csum = 0.0;
    double bsum = 0.0;
    for (int i = 0; i < atom->nlocal; i++) { csum += B[atom->type[i]] * B[atom->type[i]]; bsum += B[atom->type[i]]; }
    csumij = bsum * bsum;
    cutoff_lj = force->pair->cutforce;
    if (!function[0]) {
      two_charge_force = force->qqr2e * (force->qelectron * force->qelectron) / (force->angstrom * force->angstrom);
      accuracy = accuracy_absolute >= 0.0 ? accuracy_absolute : accuracy_relative * two_charge_force;
    }
    if (!gewaldflag_6) {
      // real-space error of the r^-6 sum = requested accuracy, on the decaying branch of the estimate
      double lo = std::sqrt(2.5) / cutoff_lj, hi = 12.0 / cutoff_lj;
      if (lj_rspace_error(lo) < accuracy) g_ewald_6 = lo;
      else {
        for (int it = 0; it < 200; it++) {
          const double mid = 0.5 * (lo + hi);
          if (lj_rspace_error(mid) > accuracy) lo = mid; else hi = mid;
        }
        g_ewald_6 = 0.5 * (lo + hi);
      }
    }
    if (!gridflag_6)
      error->all(FLERR, "pppm/disp/intel needs `kspace_modify mesh/disp nx ny nz` in this build (the qopt-based "
                        "sizing of the dispersion grid is not restated)");
    auto smooth = [](int v) { while (true) { int m = v; for (int f : {2, 3, 5}) while (m % f == 0) m /= f; if (m == 1) return v; v++; } };
    nx_pppm_6 = smooth(nx_pppm_6); ny_pppm_6 = smooth(ny_pppm_6); nz_pppm_6 = smooth(nz_pppm_6);
  }
  if (!lmp->fix_intel && lmp->dry_run) return;
  if (!lmp->fix_intel) error->all(FLERR, "The 'package intel' command is required for /intel styles");
  fix = lmp->fix_intel;
}

void PPPMDispIntel::setup() {
  if (!fix) return;
  b200md_pppm_params p;
  if (function[0]) {
    std::memset(&p, 0, sizeof(p));
    p.nx = nx_pppm; p.ny = ny_pppm; p.nz = nz_pppm; p.order = order; p.g_ewald = g_ewald;
    p.differentiation = differentiation_flag; p.scale = scale;
    fix->check(b200md_pppm_setup(fix->ctx(), &p));
  }
  if (function[1]) {
    std::memset(&p, 0, sizeof(p));
    p.nx = nx_pppm_6; p.ny = ny_pppm_6; p.nz = nz_pppm_6; p.order = order_6; p.g_ewald = g_ewald_6;
    p.differentiation = differentiation_flag;   // PPPMDisp uses one kspace_modify diff setting for both grids
    p.scale = 1.0; p.dispersion = 1; p.B = B.data();
    fix->check(b200md_pppm_setup(fix->ctx(), &p));
  }
}

void PPPMDispIntel::compute(int eflag, int vflag) {
  if (!fix) error->all(FLERR, "KSpace style pppm/disp/intel used before init()");
  double e = 0.0;
  energy = 0.0;
  for (double &v : virial) v = 0.0;
  fix->check(b200md_pppm_compute(fix->ctx(), eflag, vflag, &e, virial));   // energy_1 + energy_6, :540-541
  if (eflag & 1) energy = e;
  if (!fix->resident) {
    atom->f.assign((size_t)3 * atom->nlocal, 0.0);
    fix->check(b200md_atoms_download(fix->ctx(), nullptr, nullptr, atom->f.data(), nullptr));
  }
}
