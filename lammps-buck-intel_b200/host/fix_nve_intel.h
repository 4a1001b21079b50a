// fix_nve_intel.h — FixStyle(nve/intel,FixNVEIntel) on the device (the reference ships fix_nve_intel.cpp without its
// header; the interface is the stock FixNVE one: initial_integrate / final_integrate / reset_dt).
#pragma once
#include "fix_intel.h"

namespace LAMMPS_NS {

class FixNVEIntel : public Fix {
 public:
  explicit FixNVEIntel(LAMMPS *l) : Fix(l) { style = "nve/intel"; }
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
