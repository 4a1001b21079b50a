/* fix_nve_intel.h — class declaration for /root/reference/fix_nve_intel.cpp (the reference ships only the .cpp).
 * TEST INFRASTRUCTURE ONLY: members are exactly the ones the .cpp defines and uses (fix_nve_intel.cpp:32-199). */
#ifndef B200MD_REF_FIX_NVE_INTEL_H
#define B200MD_REF_FIX_NVE_INTEL_H
#include "fix_nve.h"
namespace LAMMPS_NS {
class FixNVEIntel : public FixNVE {
 public:
  FixNVEIntel(class LAMMPS *, int, char **);
  virtual ~FixNVEIntel();
  virtual void setup(int);
  virtual void initial_integrate(int);
  virtual void final_integrate();
  virtual void reset_dt();
  virtual double memory_usage();

 protected:
  double *_dtfm;
  int _nlocal3, _nlocal_max;
};
}
#endif
