// pppm_intel.h — KSpaceStyle(pppm/intel,PPPMIntel) and KSpaceStyle(pppm/disp/intel,PPPMDispIntel) on the device.
// Mirrors pppm_intel.h:33-39 / pppm_disp_intel.h of the reference: `PPPMIntel(LAMMPS*, int narg, char **arg)`
// (arg0 = relative accuracy), `init()`, `compute(int,int)`; `setup()` and the grid-sizing logic come from the stock
// base class PPPM, restated here because the reference does not ship it.
#ifdef KSPACE_CLASS

KSpaceStyle(pppm/intel,PPPMIntel)

#else

#ifndef B200MD_PPPM_INTEL_H
#define B200MD_PPPM_INTEL_H
#include "fix_intel.h"
#include "lammps_shim.h"

namespace LAMMPS_NS {

class PPPM : public KSpace {
 public:
  PPPM(LAMMPS *l, int narg, char **arg);
  void init() override;
  void init_charges();
  void setup() override {}
  void compute(int, int) override { error->all(FLERR, "PPPM::compute: only the /intel style is provided"); }

  // PPPM::set_grid_global + adjust_gewald: sizes nx/ny/nz_pppm and g_ewald (ik: the analytic estimate; ad: the
  // error functional of the optimal influence function, compute_qopt)
  void set_grid_global();
  double estimate_ik_error(double h, double prd, long natoms) const;
  // Q of Hockney & Eastwood summed over an n[0] x n[1] x n[2] mesh on a box prd[] (prd[2] = zprd_slab), 5 aliases per
  // dimension: stock PPPM::compute_qopt_ik / _ad and PPPMDisp::compute_qopt_ik / _ad (dispersion = 0, reference force
  // 4 pi exp(-k^2/4g^2)/k^2) and PPPMDisp::compute_qopt_6_ik / _6_ad (dispersion = 1, the r^-6 Ewald kernel) [UPSTREAM].
  // rms k-space force error = sqrt(Q / natoms) * q2 / volume (q2 -> csum for dispersion)
  static double compute_qopt(const int n[3], const double prd[3], int order, double g, int ad, int dispersion);
  double compute_df_kspace() const;   // of the Coulomb mesh as it is sized now
  double qsqsum = 0.0, qsum = 0.0;
  double acc_est[3] = {0.0, 0.0, 0.0};   // estimated absolute RMS force accuracy: total, real space, k-space
  double cutoff = 0.0;

 protected:
  double newton_raphson_f() const;
  double compute_qopt_dummy = 0.0;
  double q2 = 0.0;
  static bool factorable(int n);
};

class PPPMIntel : public PPPM {
 public:
  PPPMIntel(LAMMPS *l, int narg, char **arg) : PPPM(l, narg, arg) { suffix_flag |= Suffix::INTEL; }
  void init() override;
  void setup() override;
  void compute(int eflag, int vflag) override;

 protected:
  FixIntel *fix = nullptr;
};

}  // namespace LAMMPS_NS

#endif
#endif
