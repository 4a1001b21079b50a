/* pppm_disp.h — stand-in for the stock KSpace style pppm/disp (TEST INFRASTRUCTURE, see lammps_stub.h).  Only what a
 * translation unit that hands the solve to the device reads: the sizes and Ewald parameters of the two meshes, which
 * of the four "functions" are on, the dispersion coefficients of PPPMDisp::init_coeffs.  The
 * reference's own pppm_disp_intel.cpp needs the whole stock class (some 300 members) and is not built here. */
#ifndef B200MD_REF_PPPM_DISP_H
#define B200MD_REF_PPPM_DISP_H
#include "lammps_stub.h"

namespace LAMMPS_NS {

class PPPMDisp : public KSpace {
 public:
  PPPMDisp(LAMMPS *lmp, int narg, char **arg) : KSpace(lmp, narg, arg) {}
  virtual ~PPPMDisp() {}
  virtual void init() {}                 /* stock: function[], set_grid / set_grid_6, init_coeffs */
  virtual void compute(int, int) {}

  int function[4] = {0, 0, 0, 0};        /* Coulomb, geometric, arithmetic, no mixing */
  int nx_pppm = 0, ny_pppm = 0, nz_pppm = 0;
  int nx_pppm_6 = 0, ny_pppm_6 = 0, nz_pppm_6 = 0;
  double *B = nullptr;                   /* [ntypes+1] (geometric), [7 (ntypes+1)] (arithmetic), [nsplit (ntypes+1)] (none) */
  int nsplit = 0;
  double energy_1 = 0, energy_6 = 0, virial_1[6] = {0, 0, 0, 0, 0, 0}, virial_6[6] = {0, 0, 0, 0, 0, 0};
};

}  // namespace LAMMPS_NS
#endif
