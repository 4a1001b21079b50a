// b200_fix_intel.h — what the B200 build adds to USER-INTEL's FixIntel (`package intel`; upstream LAMMPS, not part of
// the reference): one device context per fix, and the two hand-overs of the host-stepped deployment (INTEGRATION.md
// section 3).  In this repo host/fix_intel.cpp implements them (FixIntel::ctx, upload_atoms / ensure_neighbor,
// sync_host); in a LAMMPS tree they become members of the real FixIntel.
#ifndef B200MD_B200_FIX_INTEL_H
#define B200MD_B200_FIX_INTEL_H

#include "b200md.h"

namespace LAMMPS_NS {
class FixIntel;

// the context of this fix (created by `package intel` with the precision mode of the fix)
b200md_ctx *b200_ctx(FixIntel *fix);
// before the first force contribution of a step: after a re-neighbouring (neighbor->ago == 0: atoms were exchanged /
// re-sorted) the atoms are uploaded and the device list is rebuilt (b200md_atoms_upload, b200md_neigh_build), otherwise
// only the positions move (b200md_atoms_set_x, b200md_neigh_decide) — IntelBuffers::thr_pack of the reference
void b200_positions_to_device(FixIntel *fix);
// after a force contribution: atom->f += what the device added since the last call (b200md_atoms_download) — the
// add_result_array of the reference
void b200_forces_to_host(FixIntel *fix);
// ~FixIntel: destroys the context of this fix
void b200_release(FixIntel *fix);
}  // namespace LAMMPS_NS

#endif
