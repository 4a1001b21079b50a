// fix_intel.cpp — device-context owner (`package intel`), plus fix nve/intel.
#include "fix_intel.h"

#include <vector>

#include "fix_nve_intel.h"

using namespace LAMMPS_NS;

LAMMPS::~LAMMPS() {
  delete fix_intel;
  delete atom; delete force; delete domain; delete neighbor; delete update; delete error;
}

FixIntel::FixIntel(LAMMPS *l, int device, int prec_mode) : Fix(l), _precision_mode(prec_mode) {
  style = "INTEL";
  if (prec_mode == PREC_MODE_SINGLE)
    error->all(FLERR, "package intel mode single is not provided on the device (use mixed or double)");
  const int rc = b200md_ctx_create(device, prec_mode == PREC_MODE_DOUBLE ? B200MD_PREC_DOUBLE : B200MD_PREC_MIXED, &_ctx);
  if (rc != 0) error->all(FLERR, std::string("package intel: ") + b200md_last_error(nullptr));
}

FixIntel::~FixIntel() { b200md_ctx_destroy(_ctx); }

void FixIntel::check(int rc) const {
  if (rc != 0) error->all(FLERR, b200md_last_error(_ctx));   // the library carries the reference's own messages
}

void FixIntel::upload_atoms() {
  int per[3] = {domain->periodicity[0], domain->periodicity[1], domain->periodicity[2]};
  check(b200md_set_units(_ctx, force->qqrd2e, force->ftm2v));
  check(b200md_set_box(_ctx, domain->boxlo, domain->boxhi, per));
  check(b200md_atoms_upload(_ctx, atom->nlocal, atom->ntypes, atom->x.data(), atom->v.empty() ? nullptr : atom->v.data(),
                            atom->q_flag ? atom->q.data() : nullptr, atom->type.data(), atom->mass.data()));
  // molecular systems: the special-bond partners that become bits 30-31 of the device-built list entries
  if (atom->maxspecial > 0)
    check(b200md_atoms_set_special(_ctx, atom->maxspecial, atom->nspecial.data(), atom->special.data()));
  else check(b200md_atoms_set_special(_ctx, 0, nullptr, nullptr));
  _uploaded = true;
  list_built = false;
}

void FixIntel::setup_neighbor() {
  check(b200md_neigh_setup(_ctx, neighbor->skin, neighbor->every, neighbor->delay, neighbor->dist_check));
}

void FixIntel::ensure_neighbor(bool force_build) {
  if (!list_built || force_build) {
    check(b200md_neigh_build(_ctx));
    list_built = true;
  } else {
    int rebuilt = 0;
    check(b200md_neigh_decide(_ctx, update->ntimestep, &rebuilt));
  }
}

void FixIntel::sync_host(bool x, bool v, bool f) {
  const size_t n3 = (size_t)3 * atom->nlocal;
  if (x) atom->x.resize(n3);
  if (v) atom->v.resize(n3);
  if (f) atom->f.resize(n3);
  check(b200md_atoms_download(_ctx, x ? atom->x.data() : nullptr, v ? atom->v.data() : nullptr,
                              f ? atom->f.data() : nullptr, nullptr));
}

// ---- fix nve/intel ---------------------------------------------------------------------------------------------
void FixNVEIntel::init() {
  if (!lmp->fix_intel) error->all(FLERR, "The 'package intel' command is required for /intel styles");
  fix = lmp->fix_intel;
}

void FixNVEIntel::setup(int) { reset_dt(); }

void FixNVEIntel::reset_dt() {
  dtv = update->dt;
  dtf = 0.5 * update->dt * force->ftm2v;
  // mask[i] & groupbit and atom->rmass (fix_nve_intel.cpp:147-190): handed over before _dtfm is built
  std::vector<int> in;
  if (igroup != 0) {
    in.resize(atom->nlocal);
    for (int i = 0; i < atom->nlocal; i++) in[i] = (atom->mask[i] & groupbit) ? 1 : 0;
  }
  fix->check(b200md_nve_set_group(fix->ctx(), igroup != 0 ? in.data() : nullptr,
                                  atom->rmass_flag ? atom->rmass.data() : nullptr));
  fix->check(b200md_nve_setup(fix->ctx(), update->dt));   // also precomputes dtf/mass per atom (_dtfm, :173-177)
}

void FixNVEIntel::initial_integrate(int) { fix->check(b200md_nve_initial_integrate(fix->ctx())); }
void FixNVEIntel::final_integrate() { fix->check(b200md_nve_final_integrate(fix->ctx())); }
