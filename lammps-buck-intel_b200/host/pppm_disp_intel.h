// pppm_disp_intel.h — KSpaceStyle(pppm/disp/intel,PPPMDispIntel) on the device (pppm_disp_intel.h:18-40 of the
// reference).  Of the four "functions" of PPPMDisp the reference accelerates the Coulomb grid ('c') and the
// geometric-mixing dispersion grid ('g') (pppm_disp_intel.cpp:183-313); those are the two provided here.  Arithmetic
// mixing (7 grids) and no-mixing are out of scope, as in SURVEY §2.1-6.
#pragma once
#include "pppm_intel.h"

namespace LAMMPS_NS {

class PPPMDispIntel : public PPPM {
 public:
  PPPMDispIntel(LAMMPS *l, int narg, char **arg) : PPPM(l, narg, arg) { suffix_flag |= Suffix::INTEL; }
  void init() override;
  void setup() override;
  void compute(int eflag, int vflag) override;
  int function[4] = {0, 0, 0, 0};   // Coulomb, geometric, arithmetic, none
  std::vector<double> B;            // geometric mixing: B[type] = sqrt(|C_ii|)
  double csum = 0.0, csumij = 0.0, cutoff_lj = 0.0;
  double lj_rspace_error(double g6) const;

 private:
  FixIntel *fix = nullptr;
};

}  // namespace LAMMPS_NS
