// pppm_disp_intel.cpp — host side of pppm/disp/intel.
//   PPPMDispIntel::init     pppm_disp_intel.cpp:86-109
//   PPPMDispIntel::compute  :115-554: Coulomb branch :183-243, geometric branch :245-313 (particle_map<'c'|'g'>
//                           :556-630, make_rho<'c'|'g'> :633-784 with the per-atom weight B[type]),
//                           arithmetic branch :315-407, no-mixing branch :409-467,
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
  // PPPMDisp::init [UPSTREAM]: order 1 -> function[0]; order 6 -> function[1|2|3] by the pair style's mixing rule
  // (ewald_mix; buck/long/coul/long has none to offer: geometric) and `kspace_modify mix/disp pair|geom|none`
  int *p_mix = (int *)force->pair->extract("ewald_mix", itmp);
  const int ewald_mix = p_mix ? *p_mix : Pair::GEOMETRIC;
  for (int &fn : function) fn = 0;
  function[0] = (ewald_order >> 1) & 1;
  if ((ewald_order >> 6) & 1) {
    if ((ewald_mix == Pair::GEOMETRIC || ewald_mix == Pair::SIXTHPOWER || mixflag == 1) && mixflag != 2) function[1] = 1;
    else if (ewald_mix == Pair::ARITHMETIC && mixflag != 2) function[2] = 1;
    else if (mixflag == 2) function[3] = 1;
    else error->all(FLERR, "Unsupported mixing rule in kspace_style pppm/disp");
  }
  const int rule = disp_rule();
  if (!function[0] && !rule) error->all(FLERR, "PPPMDisp used but no parameters set, for full pppm use pppm");
  if (order_6 > 7 || order > 7) error->all(FLERR, "PPPM order greater than supported by USER-INTEL");

  if (function[0]) PPPM::init();   // Coulomb grid: qsum_qsq, set_grid_global, adjust_gewald

  if (rule) {
    if (!b) error->all(FLERR, "KSpace style is incompatible with Pair style");
    // PPPMDisp::init [UPSTREAM] initialises the pair style first ("to get the coefficients"): the i-j entries of B,
    // epsilon and sigma exist only after init_one has mixed them
    force->pair->init_all_pairs();
    const int n = atom->ntypes + 1;
    // PPPMDisp::init_coeffs [UPSTREAM]
    if (rule == 1) {
      B.assign(n, 0.0);
      for (int i = 1; i < n; i++) B[i] = std::sqrt(std::fabs(b[i * n + i]));
    } else if (rule == 2) {
      double *epsilon = (double *)force->pair->extract("epsilon", itmp);
      double *sigma = (double *)force->pair->extract("sigma", itmp);
      if (!epsilon || !sigma) error->all(FLERR, "Epsilon or sigma reference not set by pair style in PPPMDisp");
      // B[7 i + k] = sqrt(eps_i) / 4 * sqrt(binom(6,k)) sigma_i^k: sum_k B_i[k] B_j[6-k] = 4 eps_ij sigma_ij^6
      const double c[7] = {1.0, std::sqrt(6.0), std::sqrt(15.0), std::sqrt(20.0), std::sqrt(15.0), std::sqrt(6.0), 1.0};
      B.assign((size_t)7 * n, 0.0);
      for (int i = 1; i < n; i++) {
        const double eps_i = std::sqrt(epsilon[i * n + i]) / 4.0, sigma_i = sigma[i * n + i];
        double sigma_n = 1.0;
        for (int k = 0; k < 7; k++) { B[7 * i + k] = sigma_n * eps_i * c[k]; sigma_n *= sigma_i; }
      }
    } else {
      B.assign(b, b + (size_t)n * n);   // the eigen-split of C_ij happens behind b200md_pppm_setup
    }
    // csum = sum_i C_ii, csumij = sum_ij C_ij over the atoms (C_ij = b[i][j] for every rule but the geometric one,
    // which replaces the pair style's off-diagonal coefficients by sqrt(C_ii C_jj))
    csum = 0.0;
    csumij = 0.0;
    std::vector<double> cnt(n, 0.0);
    for (int i = 0; i < atom->nlocal; i++) cnt[atom->type[i]] += 1.0;
    for (int i = 1; i < n; i++) {
      csum += cnt[i] * std::fabs(b[i * n + i]);
      for (int j = 1; j < n; j++)
        csumij += cnt[i] * cnt[j] * (rule == 1 ? std::sqrt(std::fabs(b[i * n + i] * b[j * n + j])) : b[i * n + j]);
    }
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
  if (disp_rule()) {
    std::memset(&p, 0, sizeof(p));
    p.nx = nx_pppm_6; p.ny = ny_pppm_6; p.nz = nz_pppm_6; p.order = order_6; p.g_ewald = g_ewald_6;
    p.differentiation = differentiation_flag;   // PPPMDisp uses one kspace_modify diff setting for both grids
    p.scale = 1.0; p.dispersion = disp_rule(); p.B = B.data();
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
