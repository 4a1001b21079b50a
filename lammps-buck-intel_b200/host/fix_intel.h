// fix_intel.h — `package intel` / fix INTEL: the object every /intel style asks for its buffers
// (pair_buck_intel.cpp:372-376 "The 'package intel' command is required for /intel styles"; precision dispatch
// :50-58; IntelBuffers ownership intel_buffers.h:272-312).  Here it owns the device context of include/b200md.h:
// the atoms live in HBM, cell-sorted, and the host arrays atom->x/v/f are mirrors that are valid after sync_host().
#pragma once
#include <string>

#include "../../include/b200md.h"
#include "lammps_shim.h"

namespace LAMMPS_NS {

class FixIntel : public Fix {
 public:
  enum { PREC_MODE_SINGLE, PREC_MODE_MIXED, PREC_MODE_DOUBLE };
  // `package intel Nphi mode double|mixed` (single is not provided on the device)
  FixIntel(LAMMPS *lmp, int device, int prec_mode);
  ~FixIntel() override;
  int precision() const { return _precision_mode; }
  b200md_ctx *ctx() const { return _ctx; }

  // resident = 1: fix nve/intel integrates on the device, atom->x/f are not exchanged every step (the fast path).
  // resident = 0: the plug-in deployment where the host integrates: pair->compute uploads x and downloads f.
  int resident = 1;

  void check(int rc) const;               // non-zero C-ABI status -> error->all with the library's message
  void upload_atoms();                    // box + units + atom->x/v/q/type/mass -> device (IntelBuffers::thr_pack)
  void setup_neighbor();                  // neighbor->skin/every/delay/dist_check -> device list parameters
  void sync_host(bool x, bool v, bool f); // device -> atom->x/v/f
  bool atoms_on_device() const { return _uploaded; }
  void invalidate() { _uploaded = false; }
  // first force evaluation of a run: builds the list; later steps call neigh_decide
  void ensure_neighbor(bool force_build);
  bool list_built = false;

 private:
  b200md_ctx *_ctx = nullptr;
  int _precision_mode;
  bool _uploaded = false;
};

}  // namespace LAMMPS_NS
