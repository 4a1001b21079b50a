// fix_nve_intel.h — FixStyle(nve/intel,FixNVEIntel) on the device (the reference ships fix_nve_intel.cpp without its
// header; the interface is the stock FixNVE one: initial_integrate / final_integrate / reset_dt).
#ifdef FIX_CLASS

FixStyle(nve/intel,FixNVEIntel)

#else

#ifndef B200MD_FIX_NVE_INTEL_H
#define B200MD_FIX_NVE_INTEL_H
#include "fix_intel.h"

namespace LAMMPS_NS {

class FixNVEIntel : public Fix {
 public:
  explicit FixNVEIntel(LAMMPS *l) : Fix(l) { style = "nve/intel"; }
  // `fix ID group nve/intel`: the stock Fix constructor signature (the caller resolves the group ID)
  FixNVEIntel(LAMMPS *l, int narg, char **arg) : Fix(l) {
    if (narg < 3) error->all(FLERR, "Illegal fix nve command");
    id = arg[0];
    style = arg[2];
  }
  void init() override;
  void setup(int vflag) override;
  void initial_integrate(int vflag) override;   // fix_nve_intel.cpp:60-99
  void final_integrate() override;              // :103-127
  void reset_dt() override;                     // :129-194 (dtv, dtf, per-atom dtf/mass)

 private:
  FixIntel *fix = nullptr;
  double dtv = 0.0, dtf = 0.0;
};

}  // namespace LAMMPS_NS

#endif
#endif
