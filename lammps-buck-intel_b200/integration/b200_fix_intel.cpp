// b200_fix_intel.cpp — the three hand-overs of b200_fix_intel.h, written against the stock LAMMPS objects (atom->x as
// double**, domain->boxlo/boxhi, neighbor->skin/every/delay/dist_check/ago, force->qqrd2e/ftm2v).  In a LAMMPS tree they
// are members of USER-INTEL's FixIntel; here they are free functions with one state record per fix, so that the upstream
// class need not change.  This is the host-stepped deployment of INTEGRATION.md section 3: LAMMPS integrates and decides
// when to re-neighbour, positions go to the device before the first force contribution of a step, forces come back
// after each contribution (the call sequence of host/pair_buck_intel.cpp: PairBuck::device_compute with
// FixIntel::resident = 0, which the GPU tests run through `lmp_b200 -host-step`).
#include "b200_fix_intel.h"

#include <map>
#include <vector>

#include "atom.h"
#include "domain.h"
#include "error.h"
#include "fix_intel.h"
#include "force.h"
#include "neighbor.h"
#include "update.h"

namespace LAMMPS_NS {

namespace {

struct B200State {
  b200md_ctx *ctx = nullptr;
  bool uploaded = false;
  int nlocal = -1;
  std::vector<double> f, fprev;   // device forces at the last hand-over (pair->compute overwrites, kspace accumulates)
  ~B200State() { if (ctx) b200md_ctx_destroy(ctx); }
};

std::map<FixIntel *, B200State> &states() {
  static std::map<FixIntel *, B200State> s;
  return s;
}

// FixIntel derives from Pointers: its LAMMPS* is protected.  As members of FixIntel the functions below read it directly.
struct FixAccess : public FixIntel {
  static LAMMPS *lmp_of(FixIntel *fix) { return fix->*(&FixAccess::lmp); }
};

void check(LAMMPS *lmp, b200md_ctx *ctx, int rc) {
  if (rc) lmp->error->all(FLERR, b200md_last_error(ctx));   // the library carries the reference's own messages
}

}  // namespace

b200md_ctx *b200_ctx(FixIntel *fix) {
  B200State &st = states()[fix];
  if (st.ctx) return st.ctx;
  LAMMPS *lmp = FixAccess::lmp_of(fix);
  if (fix->precision() == FixIntel::PREC_MODE_SINGLE)
    lmp->error->all(FLERR, "package intel mode single is not provided on the device (use mixed or double)");
  const int prec = fix->precision() == FixIntel::PREC_MODE_DOUBLE ? B200MD_PREC_DOUBLE : B200MD_PREC_MIXED;
  if (b200md_ctx_create(0, prec, &st.ctx)) {
    st.ctx = nullptr;
    lmp->error->all(FLERR, b200md_last_error(nullptr));       // no CPU fallback: no sm_100 device, no run
  }
  // what the styles' init needs before any atom is on the device: units, box, the neighbour settings (the k-space
  // brick halo is sized from skin/2)
  check(lmp, st.ctx, b200md_set_units(st.ctx, lmp->force->qqrd2e, lmp->force->ftm2v));
  check(lmp, st.ctx, b200md_set_box(st.ctx, lmp->domain->boxlo, lmp->domain->boxhi, lmp->domain->periodicity));
  check(lmp, st.ctx, b200md_neigh_setup(st.ctx, lmp->neighbor->skin, lmp->neighbor->every, lmp->neighbor->delay,
                                        lmp->neighbor->dist_check));
  return st.ctx;
}

void b200_positions_to_device(FixIntel *fix) {
  b200md_ctx *ctx = b200_ctx(fix);
  B200State &st = states()[fix];
  LAMMPS *lmp = FixAccess::lmp_of(fix);
  Atom *atom = lmp->atom;
  const int n = atom->nlocal;
  if (!st.uploaded || lmp->neighbor->ago == 0 || n != st.nlocal) {
    // LAMMPS re-neighboured: atoms may have been wrapped, exchanged and re-sorted.  Owned atoms only: the device makes
    // its own ghosts and its own (full, newton off) list
    check(lmp, ctx, b200md_set_box(ctx, lmp->domain->boxlo, lmp->domain->boxhi, lmp->domain->periodicity));
    check(lmp, ctx, b200md_atoms_upload(ctx, n, atom->ntypes, n ? atom->x[0] : nullptr, nullptr, atom->q, atom->type,
                                        atom->mass));
    check(lmp, ctx, b200md_neigh_build(ctx));
    st.uploaded = true;
    st.nlocal = n;
  } else {
    int rebuilt = 0;
    check(lmp, ctx, b200md_atoms_set_x(ctx, atom->x[0]));
    check(lmp, ctx, b200md_neigh_decide(ctx, (long)lmp->update->ntimestep, &rebuilt));
  }
  st.fprev.assign((size_t)3 * n, 0.0);   // the pair style of this step overwrites the device force array
}

void b200_forces_to_host(FixIntel *fix) {
  b200md_ctx *ctx = b200_ctx(fix);
  B200State &st = states()[fix];
  LAMMPS *lmp = FixAccess::lmp_of(fix);
  Atom *atom = lmp->atom;
  const int n = atom->nlocal;
  st.f.assign((size_t)3 * n, 0.0);
  if ((int)st.fprev.size() != 3 * n) st.fprev.assign((size_t)3 * n, 0.0);
  check(lmp, ctx, b200md_atoms_download(ctx, nullptr, nullptr, st.f.data(), nullptr));
  double **f = atom->f;
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) f[i][d] += st.f[3 * (size_t)i + d] - st.fprev[3 * (size_t)i + d];   // f +=, as add_result_array
  st.fprev.swap(st.f);
}

void b200_release(FixIntel *fix) { states().erase(fix); }

}  // namespace LAMMPS_NS
