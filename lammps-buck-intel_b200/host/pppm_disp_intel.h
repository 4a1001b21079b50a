// pppm_disp_intel.h — KSpaceStyle(pppm/disp/intel,PPPMDispIntel) on the device (pppm_disp_intel.h:18-40 of the
// reference).  All four "functions" of PPPMDisp::compute run on the device: the Coulomb grid ('c') and the
// geometric-mixing dispersion grid ('g') the reference accelerates itself (pppm_disp_intel.cpp:183-313), and the
// arithmetic-mixing (seven grids, :315-407) and no-mixing (:409-467) branches it leaves to the stock members.
#ifdef KSPACE_CLASS

KSpaceStyle(pppm/disp/intel,PPPMDispIntel)

#else

#ifndef B200MD_PPPM_DISP_INTEL_H
#define B200MD_PPPM_DISP_INTEL_H
#include "pppm_intel.h"

namespace LAMMPS_NS {

class PPPMDispIntel : public PPPM {
 public:
  PPPMDispIntel(LAMMPS *l, int narg, char **arg) : PPPM(l, narg, arg) { suffix_flag |= Suffix::INTEL; }
  void init() override;
  void setup() override;
  void compute(int eflag, int vflag) override;
  int function[4] = {0, 0, 0, 0};   // Coulomb, geometric, arithmetic, none
  int disp_rule() const { return function[1] ? 1 : (function[2] ? 2 : (function[3] ? 3 : 0)); }
  // PPPMDisp::init_coeffs: geometric B[type] = sqrt(|C_ii|); arithmetic B[7 type + k]; none: the C_ij matrix itself
  // (its eigen-split is done behind b200md_pppm_setup)
  std::vector<double> B;
  double csum = 0.0, csumij = 0.0, cutoff_lj = 0.0;
  double lj_rspace_error(double g6) const;
  // mesh sizing of stock PPPMDisp (restated): Coulomb mesh and dispersion mesh, both on PPPM::compute_qopt
  void set_grid();
  void adjust_gewald();
  void final_accuracy();
  void set_grid_6();
  void set_init_g6();
  void set_n_pppm_6();
  void adjust_gewald_6();
  void final_accuracy_6();
  double df_kspace_coul() const, df_kspace_6() const, f_coul() const;
  // estimated absolute RMS force accuracy of the Coulomb sum and of the dispersion sum: total, real space, k-space
  double acc_coul[3] = {0.0, 0.0, 0.0}, acc_6[3] = {0.0, 0.0, 0.0};

 private:
  FixIntel *fix = nullptr;
};

}  // namespace LAMMPS_NS

#endif
#endif
