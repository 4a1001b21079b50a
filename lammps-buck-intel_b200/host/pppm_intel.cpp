// pppm_intel.cpp — host side of pppm/intel.
//   PPPMIntel::init     pppm_intel.cpp:67-98   (base init, "package intel" fix, order <= INTEL_P3M_MAXORDER check)
//   PPPMIntel::compute  pppm_intel.cpp:104-317 (particle_map, make_rho, brick2fft, poisson, fieldforce, energy/virial
//                                               post-factors)                                  -> b200md_pppm_compute
// PPPM::init / set_grid_global / adjust_gewald restate the stock base class.
#include "pppm_intel.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pair_buck_intel.h"

using namespace LAMMPS_NS;

static const double MY_PI = 3.14159265358979323846;
#define INTEL_P3M_MAXORDER 7   // pppm_intel.h of the reference

void KSpace::modify_params(int narg, char **arg) {
  int i = 0;
  while (i < narg) {
    if (!std::strcmp(arg[i], "mesh") && i + 3 < narg) {
      nx_pppm = std::atoi(arg[i + 1]); ny_pppm = std::atoi(arg[i + 2]); nz_pppm = std::atoi(arg[i + 3]);
      gridflag = (nx_pppm || ny_pppm || nz_pppm) ? 1 : 0;
      i += 4;
    } else if (!std::strcmp(arg[i], "mix/disp") && i + 1 < narg) {
      // KSpace::modify_params [UPSTREAM]: which of PPPMDisp's dispersion functions serves the pair style
      if (!std::strcmp(arg[i + 1], "pair")) mixflag = 0;
      else if (!std::strcmp(arg[i + 1], "geom")) mixflag = 1;
      else if (!std::strcmp(arg[i + 1], "none")) mixflag = 2;
      else error->all(FLERR, "Illegal kspace_modify command");
      i += 2;
    } else if (!std::strcmp(arg[i], "mesh/disp") && i + 3 < narg) {
      nx_pppm_6 = std::atoi(arg[i + 1]); ny_pppm_6 = std::atoi(arg[i + 2]); nz_pppm_6 = std::atoi(arg[i + 3]);
      gridflag_6 = (nx_pppm_6 || ny_pppm_6 || nz_pppm_6) ? 1 : 0;
      i += 4;
    } else if (!std::strcmp(arg[i], "force/disp/real") && i + 1 < narg) {
      // absolute accuracies of the two halves of the dispersion sum (examples/in.hexane:16-17)
      accuracy_real_6 = std::atof(arg[i + 1]);
      i += 2;
    } else if (!std::strcmp(arg[i], "force/disp/kspace") && i + 1 < narg) {
      accuracy_kspace_6 = std::atof(arg[i + 1]);
      i += 2;
    } else if (!std::strcmp(arg[i], "force") && i + 1 < narg) {
      accuracy_absolute = std::atof(arg[i + 1]);
      i += 2;
    } else if (!std::strcmp(arg[i], "order") && i + 1 < narg) { order = std::atoi(arg[i + 1]); i += 2; }
    else if (!std::strcmp(arg[i], "order/disp") && i + 1 < narg) { order_6 = std::atoi(arg[i + 1]); i += 2; }
    else if (!std::strcmp(arg[i], "gewald") && i + 1 < narg) {
      g_ewald = std::atof(arg[i + 1]);
      gewaldflag = g_ewald != 0.0;
      i += 2;
    } else if (!std::strcmp(arg[i], "gewald/disp") && i + 1 < narg) {
      g_ewald_6 = std::atof(arg[i + 1]);
      gewaldflag_6 = g_ewald_6 != 0.0;
      i += 2;
    } else if (!std::strcmp(arg[i], "slab") && i + 1 < narg) {
      // kspace_modify slab VOLFACTOR | nozforce (KSpace::modify_params [UPSTREAM])
      if (!std::strcmp(arg[i + 1], "nozforce")) error->all(FLERR, "kspace_modify slab nozforce is not provided");
      slab_volfactor = std::atof(arg[i + 1]);
      if (slab_volfactor <= 1.0) error->all(FLERR, "Bad kspace_modify slab parameter");
      if (slab_volfactor < 2.0) error->warning(FLERR, "Kspace_modify slab param < 2.0 may cause unphysical behavior");
      slabflag = 1;
      i += 2;
    } else if (!std::strcmp(arg[i], "diff") && i + 1 < narg) {
      if (!std::strcmp(arg[i + 1], "ad")) differentiation_flag = 1;
      else if (!std::strcmp(arg[i + 1], "ik")) differentiation_flag = 0;
      else error->all(FLERR, "Illegal kspace_modify command");
      i += 2;
    } else error->all(FLERR, "Illegal kspace_modify command");
  }
}

PPPM::PPPM(LAMMPS *l, int narg, char **arg) : KSpace(l) {
  if (narg < 1) error->all(FLERR, "Illegal kspace_style pppm command");
  accuracy_relative = std::fabs(std::atof(arg[0]));
}

bool PPPM::factorable(int n) {
  for (int f : {2, 3, 5})
    while (n % f == 0) n /= f;
  return n == 1;
}

void PPPM::init() {
  init_charges();
  set_grid_global();
  // PPPM::final_accuracy [UPSTREAM] (without the table term): k-space error from the closed-form estimate (ik) or the
  // error functional of the mesh (ad), real-space error of the erfc cut-off
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2];
  const long natoms = atom->natoms;
  if (differentiation_flag == 1) acc_est[2] = compute_df_kspace();
  else {
    const double zprd_slab = zprd * slab_volfactor;
    const double lx = estimate_ik_error(xprd / nx_pppm, xprd, natoms), ly = estimate_ik_error(yprd / ny_pppm, yprd, natoms),
                 lz = estimate_ik_error(zprd_slab / nz_pppm, zprd_slab, natoms);
    acc_est[2] = std::sqrt(lx * lx + ly * ly + lz * lz) / std::sqrt(3.0);
  }
  acc_est[1] = 2.0 * q2 * std::exp(-g_ewald * g_ewald * cutoff * cutoff) / std::sqrt(natoms * cutoff * xprd * yprd * zprd);
  acc_est[0] = std::sqrt(acc_est[1] * acc_est[1] + acc_est[2] * acc_est[2]);
}

// the checks of PPPM::init and qsum_qsq, without the mesh sizing (PPPMDisp sizes its Coulomb mesh its own way)
void PPPM::init_charges() {
  if (domain->triclinic) error->all(FLERR, "Cannot (yet) use PPPM with triclinic box and this build");
  if (slabflag == 0)
    for (int d = 0; d < 3; d++)
      if (!domain->periodicity[d]) error->all(FLERR, "Cannot use nonperiodic boundaries with PPPM");
  if (slabflag && (!domain->periodicity[0] || !domain->periodicity[1] || domain->periodicity[2]))
    error->all(FLERR, "Incorrect boundaries with slab PPPM");
  if (!atom->q_flag) error->all(FLERR, "KSpace style requires atom attribute q");
  if (order < 2 || order > 7) error->all(FLERR, "PPPM order cannot be < 2 or > than 7");
  if (!force->pair) error->all(FLERR, "KSpace style is incompatible with Pair style");
  int itmp;
  double *p_cutoff = (double *)force->pair->extract("cut_coul", itmp);
  if (!p_cutoff) error->all(FLERR, "KSpace style is incompatible with Pair style");
  cutoff = *p_cutoff;
  // qsum_qsq
  qsum = qsqsum = 0.0;
  for (int i = 0; i < atom->nlocal; i++) { qsum += atom->q[i]; qsqsum += atom->q[i] * atom->q[i]; }
  if (qsqsum == 0.0) error->all(FLERR, "Cannot use kspace solver on system with no charge");
  q2 = qsqsum * force->qqrd2e;
  two_charge_force = force->qqr2e * (force->qelectron * force->qelectron) / (force->angstrom * force->angstrom);
  accuracy = accuracy_absolute >= 0.0 ? accuracy_absolute : accuracy_relative * two_charge_force;
}

// analytic rms force error of ik differentiation for grid spacing h along a box edge prd
double PPPM::estimate_ik_error(double h, double prd, long natoms) const {
  static const double acons[8][7] = {
      {0, 0, 0, 0, 0, 0, 0},
      {2.0 / 3.0, 0, 0, 0, 0, 0, 0},
      {1.0 / 50.0, 5.0 / 294.0, 0, 0, 0, 0, 0},
      {1.0 / 588.0, 7.0 / 1440.0, 21.0 / 3872.0, 0, 0, 0, 0},
      {1.0 / 4320.0, 3.0 / 1936.0, 7601.0 / 2271360.0, 143.0 / 28800.0, 0, 0, 0},
      {1.0 / 23232.0, 7601.0 / 13628160.0, 143.0 / 69120.0, 517231.0 / 106536960.0, 106640677.0 / 11737571328.0, 0, 0},
      {691.0 / 68140800.0, 13.0 / 57600.0, 47021.0 / 35512320.0, 9694607.0 / 2095994880.0,
       733191589.0 / 59609088000.0, 326190917.0 / 11700633600.0, 0},
      {1.0 / 345600.0, 3617.0 / 35512320.0, 745739.0 / 838397952.0, 56399353.0 / 12773376000.0,
       25091609.0 / 1560084480.0, 1755948832039.0 / 36229939200000.0, 4887769399.0 / 37838389248.0}};
  double sum = 0.0;
  for (int m = 0; m < order; m++) sum += acons[order][m] * std::pow(h * g_ewald, 2.0 * m);
  return q2 * std::pow(h * g_ewald, (double)order) * std::sqrt(g_ewald * prd * std::sqrt(2.0 * MY_PI) * sum / natoms) /
         (prd * prd);
}

double PPPM::newton_raphson_f() const {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2];
  const long natoms = atom->natoms;
  const double df_rspace = 2.0 * q2 * std::exp(-g_ewald * g_ewald * cutoff * cutoff) /
                           std::sqrt(natoms * cutoff * xprd * yprd * zprd);
  const double lx = estimate_ik_error(xprd / nx_pppm, xprd, natoms);
  const double ly = estimate_ik_error(yprd / ny_pppm, yprd, natoms);
  const double zprd_slab = zprd * slab_volfactor;
  const double lz = estimate_ik_error(zprd_slab / nz_pppm, zprd_slab, natoms);
  return df_rspace - std::sqrt(lx * lx + ly * ly + lz * lz) / std::sqrt(3.0);
}

// ---- the error functional of the optimal influence function -------------------------------------------------------
// Q = sum over mesh wave vectors k != 0 of  sum_m |R(k_m)|^2 - |sum_m U^2(k_m) R(k_m).D(k)|^2 / (|D(k)|^2 (sum_m U^2(k_m))^2)
// (Hockney & Eastwood eq. 8-23), k_m = k + 2 pi m n / L the aliases (m = -2..2 per dimension, as stock LAMMPS sums
// them), U = the charge-assignment window (sin x / x)^order, R = -i k phi(k) the reference force, D = i k (ik
// differentiation) or the window-weighted gradient (ad).  phi: 4 pi exp(-k^2/4g^2)/k^2 (Coulomb) or
// -(pi^1.5 g^3/3) [(1 - 2b^2) exp(-b^2) + 2 b^3 sqrt(pi) erfc(b)], b = k/2g (the r^-6 Ewald kernel, PPPMDisp).
// The summand is even in each component of k — the alias set mirrors with the wave number — except at the Nyquist
// index of an even mesh, so one octant is summed with multiplicities.
double PPPM::compute_qopt(const int n[3], const double prd[3], int order, double g, int ad, int dispersion) {
  struct Dim { std::vector<double> mult, k, q, s, w2; };   // per representative index: k; per (index, alias): q, s, w^2
  Dim D[3];
  for (int d = 0; d < 3; d++) {
    const double unitk = 2.0 * MY_PI / prd[d];
    for (int i = 0; i < n[d]; i++) {
      const int kper = i - n[d] * (2 * i / n[d]);
      if (kper < 0 && -2 * kper != n[d]) continue;   // counted with its mirror image
      D[d].mult.push_back(kper > 0 ? 2.0 : 1.0);
      D[d].k.push_back(unitk * kper);
      for (int a = -2; a <= 2; a++) {
        const double q = unitk * (kper + n[d] * a);
        const double arg = 0.5 * q * prd[d] / n[d];
        const double w = arg != 0.0 ? std::pow(std::sin(arg) / arg, order) : 1.0;
        D[d].q.push_back(q);
        D[d].s.push_back(std::exp(-0.25 * (q / g) * (q / g)));
        D[d].w2.push_back(w * w);
      }
    }
  }
  const double inv2ew = 0.5 / g, rtpi = std::sqrt(MY_PI), g3 = g * g * g;
  const int nz = (int)D[2].k.size(), ny = (int)D[1].k.size(), nx = (int)D[0].k.size();
  double qopt = 0.0;
#pragma omp parallel for reduction(+ : qopt) schedule(dynamic, 1)
  for (int m = 0; m < nz; m++)
    for (int l = 0; l < ny; l++)
      for (int k = 0; k < nx; k++) {
        const double kx = D[0].k[k], ky = D[1].k[l], kz = D[2].k[m];
        const double sqk = kx * kx + ky * ky + kz * kz;
        if (sqk == 0.0) continue;
        double sum1 = 0.0, sum2 = 0.0, sum3 = 0.0, sum4 = 0.0;
        for (int a = 0; a < 5; a++) {
          const double qx = D[0].q[5 * k + a], sx = D[0].s[5 * k + a], wx = D[0].w2[5 * k + a];
          for (int b = 0; b < 5; b++) {
            const double qy = D[1].q[5 * l + b], sy = D[1].s[5 * l + b], wy = D[1].w2[5 * l + b];
            for (int c = 0; c < 5; c++) {
              const double qz = D[2].q[5 * m + c], sz = D[2].s[5 * m + c], wz = D[2].w2[5 * m + c];
              const double dot1 = kx * qx + ky * qy + kz * qz;
              const double dot2 = qx * qx + qy * qy + qz * qz;
              const double u2 = wx * wy * wz, s3 = sx * sy * sz;
              if (!dispersion) {
                sum1 += s3 * s3 / dot2 * 16.0 * MY_PI * MY_PI;
                sum2 += ad ? s3 * u2 * 4.0 * MY_PI : u2 * s3 * 4.0 * MY_PI / dot2 * dot1;
              } else {
                const double rtdot2 = std::sqrt(dot2);
                const double term = g3 * ((1.0 - 2.0 * dot2 * inv2ew * inv2ew) * s3 +
                                          2.0 * dot2 * rtdot2 * inv2ew * inv2ew * inv2ew * rtpi * std::erfc(rtdot2 * inv2ew));
                sum1 += term * term * MY_PI * MY_PI * MY_PI / 9.0 * dot2;
                sum2 += -u2 * term * MY_PI * rtpi / 3.0 * (ad ? dot2 : dot1);
              }
              sum3 += u2;
              sum4 += dot2 * u2;
            }
          }
        }
        const double lost = ad ? sum2 * sum2 / (sum3 * sum4) : sum2 * sum2 / (sum3 * sum3 * sqk);
        qopt += D[0].mult[k] * D[1].mult[l] * D[2].mult[m] * (sum1 - lost);
      }
  return qopt;
}

double PPPM::compute_df_kspace() const {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd_slab = domain->prd[2] * slab_volfactor;
  const int n[3] = {nx_pppm, ny_pppm, nz_pppm};
  const double prd[3] = {xprd, yprd, zprd_slab};
  const double qopt = compute_qopt(n, prd, order, g_ewald, differentiation_flag, 0);
  return std::sqrt(qopt / atom->natoms) * q2 / (xprd * yprd * zprd_slab);
}

void PPPM::set_grid_global() {
  const double xprd = domain->prd[0], yprd = domain->prd[1], zprd = domain->prd[2];
  const long natoms = atom->natoms;
  if (!gewaldflag) {
    if (accuracy <= 0.0) error->all(FLERR, "KSpace accuracy must be > 0");
    g_ewald = accuracy * std::sqrt(natoms * cutoff * xprd * yprd * zprd) / (2.0 * q2);
    if (g_ewald >= 1.0) g_ewald = (1.35 - 0.15 * std::log(accuracy)) / cutoff;
    else g_ewald = std::sqrt(-std::log(g_ewald)) / cutoff;
  }
  if (!gridflag && differentiation_flag == 1) {
    // ad differentiation has no closed-form estimate: shrink a uniform spacing by 5 % until the error functional of
    // the mesh meets the accuracy (PPPM::set_grid_global [UPSTREAM])
    double h = 4.0 / g_ewald;
    for (int count = 0;; count++) {
      nx_pppm = std::max(static_cast<int>(xprd / h), 2);
      ny_pppm = std::max(static_cast<int>(yprd / h), 2);
      nz_pppm = std::max(static_cast<int>(zprd * slab_volfactor / h), 2);
      if (compute_df_kspace() <= accuracy) break;
      if (count > 500) error->all(FLERR, "Could not compute grid size");
      h *= 0.95;
    }
  } else if (!gridflag) {
    int *n[3] = {&nx_pppm, &ny_pppm, &nz_pppm};
    const double prd[3] = {xprd, yprd, zprd * slab_volfactor};   // zprd_slab
    for (int d = 0; d < 3; d++) {
      double h = 1.0 / g_ewald;
      int k = static_cast<int>(prd[d] / h) + 1;
      double err = estimate_ik_error(h, prd[d], natoms);
      while (err > accuracy) {
        err = estimate_ik_error(h, prd[d], natoms);
        k++;
        h = prd[d] / k;
      }
      *n[d] = k;
    }
  }
  while (!factorable(nx_pppm)) nx_pppm++;
  while (!factorable(ny_pppm)) ny_pppm++;
  while (!factorable(nz_pppm)) nz_pppm++;
  if (nx_pppm >= 16384 || ny_pppm >= 16384 || nz_pppm >= 16384) error->all(FLERR, "PPPM grid is too large");
  if (!gewaldflag) {   // adjust_gewald: Newton-Raphson on the error balance (the ik estimate, whatever the differentiation)
    for (int i = 0; i < 10000; i++) {
      const double f0 = newton_raphson_f();
      g_ewald += 1.0e-6;
      const double f1 = newton_raphson_f();
      g_ewald -= 1.0e-6;
      g_ewald -= f0 / ((f1 - f0) / 1.0e-6);
      if (std::fabs(newton_raphson_f()) < 1.0e-5) return;
    }
    error->all(FLERR, "Could not compute g_ewald");
  }
}

// ---- PPPMIntel -----------------------------------------------------------------------------------------------
void PPPMIntel::init() {
  PPPM::init();
  if (!lmp->fix_intel && lmp->dry_run) return;
  if (!lmp->fix_intel) error->all(FLERR, "The 'package intel' command is required for /intel styles");   // :73-74
  fix = lmp->fix_intel;
  if (order > INTEL_P3M_MAXORDER) error->all(FLERR, "PPPM order greater than supported by USER-INTEL");  // :87-88
}

void PPPMIntel::setup() {
  if (!fix) return;   // dry run
  b200md_pppm_params p;
  std::memset(&p, 0, sizeof(p));
  p.nx = nx_pppm; p.ny = ny_pppm; p.nz = nz_pppm;
  p.order = order;
  p.g_ewald = g_ewald;
  p.differentiation = differentiation_flag;
  p.scale = scale;
  p.slab_volfactor = slabflag ? slab_volfactor : 0.0;
  fix->check(b200md_pppm_setup(fix->ctx(), &p));
}

void PPPMIntel::compute(int eflag, int vflag) {
  if (!fix) error->all(FLERR, "KSpace style pppm/intel used before init()");
  if (!fix->resident) {
    // plug-in deployment: positions were refreshed by the pair style of this step; forces are accumulated on the
    // device and downloaded once, after the last force contribution
  }
  double e = 0.0;
  energy = 0.0;
  for (double &v : virial) v = 0.0;
  fix->check(b200md_pppm_compute(fix->ctx(), eflag, vflag, &e, virial));
  if (eflag & 1) energy = e;
  // eflag_atom / vflag_atom: stock poisson_peratom + fieldforce_peratom (reached at pppm_intel.cpp:224-229, :876)
  if ((eflag & 2) || (vflag & 4)) {
    if (eflag & 2) eatom.assign(atom->nlocal, 0.0);
    if (vflag & 4) vatom.assign((size_t)6 * atom->nlocal, 0.0);
    fix->check(b200md_pppm_peratom(fix->ctx(), (eflag & 2) ? eatom.data() : nullptr, (vflag & 4) ? vatom.data() : nullptr));
  }
  if (!fix->resident) {
    atom->f.assign((size_t)3 * atom->nlocal, 0.0);
    fix->check(b200md_atoms_download(fix->ctx(), nullptr, nullptr, atom->f.data(), nullptr));
  }
}
