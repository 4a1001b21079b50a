// pppm_disp_intel.cpp — host side of pppm/disp/intel.
//   PPPMDispIntel::init     pppm_disp_intel.cpp:86-109
//   PPPMDispIntel::compute  :115-554: Coulomb branch :183-243, geometric branch :245-313 (particle_map<'c'|'g'>
//                           :556-630, make_rho<'c'|'g'> :633-784 with the per-atom weight B[type]),
//                           arithmetic branch :315-407, no-mixing branch :409-467,
//                           energy/virial post-factors :470-510                         -> b200md_pppm_compute
// init_coeffs / the real-space accuracy estimate restate the stock base class PPPMDisp.
#include "pppm_disp_intel.h"

#include <algorithm>
#include <cmath>
#include <cstring>

using namespace LAMMPS_NS;

static const double MY_PI = 3.14159265358979323846;

double PPPMDispIntel::lj_rspace_error(double g6) const {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2] * slab_volfactor;   // zprd_slab
  double rgs = cutoff_lj * g6;
  rgs *= rgs;
  const double rgs_inv = 1.0 / rgs;
  return csum / std::sqrt((double)atom->natoms * xprd * yprd * zprd * cutoff_lj) * std::sqrt(MY_PI) * std::pow(g6, 5.0) *
         std::exp(-rgs) * (1.0 + rgs_inv * (3.0 + rgs_inv * (6.0 + rgs_inv * 6.0)));
}

// ---- mesh sizing of the stock base class PPPMDisp [UPSTREAM], restated: both meshes are sized on the
// error functional of the optimal influence function (PPPM::compute_qopt), not on PPPM's closed-form ik estimate ------

double PPPMDispIntel::df_kspace_coul() const {   // sqrt(qopt / natoms) q2 / volume on the Coulomb mesh
  return compute_df_kspace();
}

double PPPMDispIntel::df_kspace_6() const {      // the same with csum = sum_i C_ii on the dispersion mesh
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd_slab = domain->prd[2] * slab_volfactor;
  const int n[3] = {nx_pppm_6, ny_pppm_6, nz_pppm_6};
  const double prd[3] = {xprd, yprd, zprd_slab};
  const double qopt = compute_qopt(n, prd, order_6, g_ewald_6, differentiation_flag, 1);
  return std::sqrt(qopt / atom->natoms) * csum / (xprd * yprd * zprd_slab);
}

// PPPMDisp::set_grid: g_ewald from the real-space error, then a uniform spacing shrunk by 5 % a time from 4/g_ewald
void PPPMDispIntel::set_grid() {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2];
  const long natoms = atom->natoms;
  if (!gewaldflag) {
    if (accuracy <= 0.0) error->all(FLERR, "KSpace accuracy must be > 0");
    g_ewald = accuracy * std::sqrt(natoms * cutoff * xprd * yprd * zprd) / (2.0 * q2);
    if (g_ewald >= 1.0) error->all(FLERR, "KSpace accuracy too large to estimate G vector");
    g_ewald = std::sqrt(-std::log(g_ewald)) / cutoff;
  }
  if (!gridflag) {
    double h = 4.0 / g_ewald;
    for (int count = 1;; count++) {
      nx_pppm = std::max(static_cast<int>(xprd / h), 2);
      ny_pppm = std::max(static_cast<int>(yprd / h), 2);
      nz_pppm = std::max(static_cast<int>(zprd * slab_volfactor / h), 2);
      if (df_kspace_coul() <= accuracy || count > 500) break;
      h *= 0.95;
    }
  }
  while (!factorable(nx_pppm)) nx_pppm++;
  while (!factorable(ny_pppm)) ny_pppm++;
  while (!factorable(nz_pppm)) nz_pppm++;
}

double PPPMDispIntel::f_coul() const {   // PPPMDisp::f: real-space minus k-space error of the Coulomb sum
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2];
  const double df_rspace = 2.0 * q2 * std::exp(-g_ewald * g_ewald * cutoff * cutoff) /
                           std::sqrt(atom->natoms * cutoff * xprd * yprd * zprd);
  return df_rspace - df_kspace_coul();
}

void PPPMDispIntel::adjust_gewald() {    // Newton-Raphson on f, one-sided difference of 1e-6 (PPPMDisp::derivf)
  for (int i = 0; i < 10000; i++) {
    const double f1 = f_coul();
    g_ewald += 1.0e-6;
    const double f2 = f_coul();
    g_ewald -= 1.0e-6;
    g_ewald -= f1 / ((f2 - f1) / 1.0e-6);
    if (std::fabs(f_coul()) < 1.0e-5) return;
  }
  error->all(FLERR, "Could not compute g_ewald");
}

void PPPMDispIntel::final_accuracy() {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2];
  acc_coul[2] = df_kspace_coul();
  acc_coul[1] = 2.0 * q2 * std::exp(-g_ewald * g_ewald * cutoff * cutoff) / std::sqrt(atom->natoms * cutoff * xprd * yprd * zprd);
  acc_coul[0] = std::sqrt(acc_coul[1] * acc_coul[1] + acc_coul[2] * acc_coul[2]);
}

// PPPMDisp::set_init_g6: the real-space error falls monotonically with g_ewald_6; bracket the root from 1/cutoff by
// doubling / halving, then bisect to 1e-5
void PPPMDispIntel::set_init_g6() {
  const double acc_rspace = accuracy_real_6 > 0.0 ? accuracy_real_6 : accuracy;
  const int LARGE = 10000;
  const double SMALL = 0.00001;
  int counter = 0;
  double g_old = g_ewald_6 = 1.0 / cutoff_lj;
  double df_real = lj_rspace_error(g_ewald_6) - acc_rspace;
  if (df_real > 0.0)
    while (df_real > 0.0 && counter < LARGE) {
      counter++;
      g_old = g_ewald_6;
      g_ewald_6 *= 2.0;
      df_real = lj_rspace_error(g_ewald_6) - acc_rspace;
    }
  if (df_real < 0.0)
    while (df_real < 0.0 && counter < LARGE) {
      counter++;
      g_old = g_ewald_6;
      g_ewald_6 *= 0.5;
      df_real = lj_rspace_error(g_ewald_6) - acc_rspace;
    }
  if (counter >= LARGE - 1) error->all(FLERR, "Cannot compute initial g_ewald_disp");
  double gmin = std::min(g_old, g_ewald_6), gmax = std::max(g_old, g_ewald_6);
  g_ewald_6 = gmin + 0.5 * (gmax - gmin);
  counter = 0;
  while (gmax - gmin > SMALL && counter < LARGE) {
    counter++;
    df_real = lj_rspace_error(g_ewald_6) - acc_rspace;
    if (df_real < 0.0) gmax = g_ewald_6;
    else gmin = g_ewald_6;
    g_ewald_6 = gmin + 0.5 * (gmax - gmin);
  }
  if (counter >= LARGE - 1) error->all(FLERR, "Cannot compute initial g_ewald_disp");
}

// PPPMDisp::set_n_pppm_6: uniform spacing from 4/g_ewald_6, shrunk by 5 % a time until the k-space error of the
// dispersion mesh meets `kspace_modify force/disp/kspace` (else the overall accuracy)
void PPPMDispIntel::set_n_pppm_6() {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd_slab = domain->prd[2] * slab_volfactor;
  const double acc_kspace = accuracy_kspace_6 > 0.0 ? accuracy_kspace_6 : accuracy;
  double h = 4.0 / g_ewald_6;
  for (int count = 1;; count++) {
    nx_pppm_6 = std::max(static_cast<int>(xprd / h), 2);
    ny_pppm_6 = std::max(static_cast<int>(yprd / h), 2);
    nz_pppm_6 = std::max(static_cast<int>(zprd_slab / h), 2);
    if (df_kspace_6() <= acc_kspace || count > 500) break;
    h *= 0.95;
  }
}

void PPPMDispIntel::set_grid_6() {
  if (!gewaldflag_6) set_init_g6();
  if (!gridflag_6) set_n_pppm_6();
  while (!factorable(nx_pppm_6)) nx_pppm_6++;
  while (!factorable(ny_pppm_6)) ny_pppm_6++;
  while (!factorable(nz_pppm_6)) nz_pppm_6++;
}

void PPPMDispIntel::adjust_gewald_6() {   // Newton-Raphson on f_6 = real-space minus k-space error of the r^-6 sum
  auto f_6 = [&]() { return lj_rspace_error(g_ewald_6) - df_kspace_6(); };
  for (int i = 0; i < 10000; i++) {
    const double f1 = f_6();
    g_ewald_6 += 1.0e-6;
    const double f2 = f_6();
    g_ewald_6 -= 1.0e-6;
    g_ewald_6 -= f1 / ((f2 - f1) / 1.0e-6);
    if (std::fabs(f_6()) < 1.0e-5) return;
  }
  error->all(FLERR, "Could not adjust g_ewald_6");
}

void PPPMDispIntel::final_accuracy_6() {
  acc_6[1] = lj_rspace_error(g_ewald_6);
  acc_6[2] = df_kspace_6();
  acc_6[0] = std::sqrt(acc_6[1] * acc_6[1] + acc_6[2] * acc_6[2]);
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

  if (function[0]) {   // Coulomb mesh: qsum_qsq, PPPMDisp::set_grid, adjust_gewald, final_accuracy
    init_charges();
    set_grid();
    if (!gewaldflag) adjust_gewald();
    final_accuracy();
  }

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
    double *p_cutoff_lj = (double *)force->pair->extract("cut_LJ", itmp);
    if (!p_cutoff_lj) error->all(FLERR, "KSpace style is incompatible with Pair style");
    cutoff_lj = *p_cutoff_lj;
    if (!function[0]) {
      two_charge_force = force->qqr2e * (force->qelectron * force->qelectron) / (force->angstrom * force->angstrom);
      accuracy = accuracy_absolute >= 0.0 ? accuracy_absolute : accuracy_relative * two_charge_force;
    }
    set_grid_6();
    // PPPMDisp::init [UPSTREAM]: with one accuracy for both halves of the sum, g_ewald_6 is re-balanced on the mesh
    if (!gewaldflag_6 && accuracy_kspace_6 == accuracy_real_6) adjust_gewald_6();
    final_accuracy_6();
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
