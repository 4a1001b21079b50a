// pppm_intel.h — KSpaceStyle(pppm/intel,PPPMIntel) and KSpaceStyle(pppm/disp/intel,PPPMDispIntel) on the device.
// Mirrors pppm_intel.h:33-39 / pppm_disp_intel.h of the reference: `PPPMIntel(LAMMPS*, int narg, char **arg)`
// (arg0 = relative accuracy), `init()`, `compute(int,int)`; `setup()` and the grid-sizing logic come from the stock
// base class PPPM, restated here because the reference does not ship it.
#pragma once
#include "fix_intel.h"
#include "lammps_shim.h"

namespace LAMMPS_NS {

class PPPM : public KSpace {
 public:
  PPPM(LAMMPS *l, int narg, char **arg);
  void init() override;
  void setup() override {}
  void compute(int, int) override { error->all(FLERR, "PPPM::compute: only the /intel style is provided"); }

  // PPPM::set_grid_global + adjust_gewald (ik differentiation): sizes nx/ny/nz_pppm and g_ewald
  void set_grid_global();
  double estimate_ik_error(double h, double prd, long natoms) const;
  double qsqsum = 0.0, qsum = 0.0;
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
