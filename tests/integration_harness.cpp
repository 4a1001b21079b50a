// integration_harness.cpp — TEST INFRASTRUCTURE.  Drives the translation units of lammps-buck-intel_b200/integration/
// (the B200 bodies of the classes that the reference's OWN headers declare) the way oracle/ref_harness.cpp drives the
// reference's own bodies: a stand-in LAMMPS "instance" (oracle/ref_shim), the class objects, init_style() / init(),
// compute().  Built into oracle/_ref/libinteg.so by oracle/Makefile.ref where /root/reference is present; linked
// against libb200md.so.  On a machine without a B200 the first C-ABI call fails and the message comes back through
// error->all, as in LAMMPS; on a B200 the forces can be compared with oracle/_ref (tests/test_integration.py).
#include <cstring>
#include <string>
#include <vector>

#include "fix_intel.h"
#include "intel_buffers_impl.h"
#include "pair_buck_coul_long_intel.h"
#include "pppm_intel.h"

#include "b200_fix_intel.h"
#include "oracle.h"

using namespace LAMMPS_NS;

namespace {

struct World {   // one LAMMPS "instance": the objects the classes reach through Pointers
  LAMMPS lmp;
  Memory memory;
  Error error;
  Atom atom;
  Comm comm;
  Force force;
  Neighbor neighbor;
  Modify modify;
  Update update;
  Domain domain;
  Group group;
  FixIntel *fix = nullptr;
  Fix *fixes[1] = {nullptr};
  std::vector<double> xbuf, fbuf;
  std::vector<double *> xrow, frow;

  explicit World(int prec) {
    lmp.memory = &memory; lmp.error = &error; lmp.atom = &atom; lmp.comm = &comm; lmp.force = &force;
    lmp.neighbor = &neighbor; lmp.modify = &modify; lmp.update = &update; lmp.domain = &domain; lmp.group = &group;
    static char id[] = "package_intel";
    char *arg[1] = {id};
    fix = new FixIntel(&lmp, 1, arg, prec == ORC_PREC_MIXED ? FixIntel::PREC_MODE_MIXED : FixIntel::PREC_MODE_DOUBLE);
    fixes[0] = fix;
    modify.fix = fixes;
    modify.nfix = 1;
  }
  ~World() {
    b200_release(fix);
    delete fix;
  }
  static void rows(std::vector<double> &buf, std::vector<double *> &row, int n) {
    buf.assign((size_t)3 * (n + 1), 0.0);
    row.resize((size_t)n + 1);
    for (int i = 0; i <= n; i++) row[i] = buf.data() + (size_t)3 * i;
  }
};

struct StubKSpace : public KSpace {
  explicit StubKSpace(LAMMPS *lmp) : KSpace(lmp, 0, nullptr) {}
  void compute(int, int) {}
};

void fill2(double **dst, const double *src, int tp1) {
  for (int i = 0; i < tp1; i++)
    for (int j = 0; j < tp1; j++) dst[i][j] = src ? src[i * tp1 + j] : 0.0;
}

int fail(char *err, int errlen, const char *msg) {
  if (err && errlen > 0) {
    strncpy(err, msg, (size_t)errlen - 1);
    err[errlen - 1] = 0;
  }
  return 1;
}

}  // namespace

extern "C" {

// `pair_style buck/coul/long` + `kspace_style pppm` (nx = 0: no k-space style, g_ewald from p) for `nsteps` force
// evaluations with the positions displaced by `dx` per step (exercises upload, then the positions-only path), through
// PairBuckCoulLongIntel::init_style / compute and PPPMIntel::init / compute of integration/.  f [nlocal][3] receives
// atom->f of the last evaluation, ev[8] the pair tallies, ek / vk[6] the k-space energy and virial.
int integ_buck_coul_long(int prec, int nlocal, const double *x, const int *type, const double *q, int ntypes,
                         const double *mass, const double *boxlo, const double *boxhi, double skin, const double *A,
                         const double *rho, const double *C, const double *cut_lj, double cut_coul,
                         const orc_pair_params *p, int nx, int ny, int nz, int order, int diff_ad, int eflag, int vflag,
                         int nsteps, const double *dx, double *f, double *ev, double *ek, double *vk, char *err,
                         int errlen) {
  try {
    World w(prec);
    const int tp1 = ntypes + 1;
    std::vector<int> typev(type, type + nlocal);
    typev.push_back(1);
    std::vector<double> qv(q, q + nlocal), massv(mass, mass + tp1);
    qv.push_back(0.0);
    w.atom.nlocal = nlocal; w.atom.nghost = 0; w.atom.nmax = nlocal + 1; w.atom.natoms = nlocal; w.atom.ntypes = ntypes;
    World::rows(w.xbuf, w.xrow, nlocal);
    World::rows(w.fbuf, w.frow, nlocal);
    memcpy(w.xbuf.data(), x, sizeof(double) * 3 * (size_t)nlocal);
    w.atom.x = w.xrow.data(); w.atom.f = w.frow.data();
    w.atom.type = typev.data(); w.atom.q = qv.data(); w.atom.mass = massv.data();
    for (int d = 0; d < 3; d++) {
      w.domain.boxlo[d] = boxlo[d]; w.domain.boxhi[d] = boxhi[d]; w.domain.prd[d] = boxhi[d] - boxlo[d];
    }
    w.force.qqrd2e = p->qqrd2e;
    for (int k = 0; k < 4; k++) { w.force.special_lj[k] = p->special_lj[k]; w.force.special_coul[k] = p->special_coul[k]; }
    w.neighbor.skin = skin; w.neighbor.every = 1; w.neighbor.delay = 0; w.neighbor.dist_check = 1;

    // LAMMPS::init order: force->init() runs kspace->init() before pair->init()
    StubKSpace ks(&w.lmp);
    static char a0[] = "1.0e-4";
    char *karg[1] = {a0};
    PPPMIntel pp(&w.lmp, 1, karg);
    KSpace *kspace = &ks;
    if (nx > 0) {
      pp.nx_pppm = nx; pp.ny_pppm = ny; pp.nz_pppm = nz; pp.order = order;
      pp.differentiation_flag = diff_ad; pp.scale = 1.0;
      kspace = &pp;
    }
    kspace->g_ewald = p->g_ewald;
    w.force.kspace = kspace;
    if (nx > 0) pp.init();

    PairBuckCoulLongIntel pair(&w.lmp);
    pair.allocate();
    fill2(pair.a, A, tp1); fill2(pair.rho, rho, tp1); fill2(pair.c, C, tp1); fill2(pair.cut_lj, cut_lj, tp1);
    pair.cut_coul = cut_coul;
    pair.ncoultablebits = p->ncoultablebits; pair.ncoulmask = p->ncoulmask; pair.ncoulshiftbits = p->ncoulshiftbits;
    pair.tabinnersq = p->tabinnersq;
    pair.rtable = const_cast<double *>(p->rtable); pair.drtable = const_cast<double *>(p->drtable);
    pair.ftable = const_cast<double *>(p->ftable); pair.dftable = const_cast<double *>(p->dftable);
    pair.etable = const_cast<double *>(p->etable); pair.detable = const_cast<double *>(p->detable);
    pair.ctable = const_cast<double *>(p->ctable); pair.dctable = const_cast<double *>(p->dctable);
    for (int i = 1; i < tp1; i++)
      for (int j = 1; j < tp1; j++) pair.setflag[i][j] = 1;
    w.force.pair = &pair;
    pair.init_style();

    for (int s = 0; s < nsteps; s++) {
      w.neighbor.ago = s;                                   // ago == 0: LAMMPS has just re-neighboured
      w.update.ntimestep = s;
      if (s > 0 && dx)
        for (size_t k = 0; k < (size_t)3 * nlocal; k++) w.xbuf[k] += dx[k];
      std::fill(w.fbuf.begin(), w.fbuf.end(), 0.0);         // Verlet::force_clear
      pair.compute(eflag, vflag);
      if (nx > 0) pp.compute(eflag, vflag);
    }
    memcpy(f, w.fbuf.data(), sizeof(double) * 3 * (size_t)nlocal);
    ev[0] = pair.eng_vdwl; ev[1] = pair.eng_coul;
    for (int k = 0; k < 6; k++) ev[2 + k] = pair.virial[k];
    if (ek) *ek = nx > 0 ? pp.energy : 0.0;
    if (vk) for (int k = 0; k < 6; k++) vk[k] = nx > 0 ? pp.virial[k] : 0.0;
    return 0;
  } catch (const std::exception &e) {
    return fail(err, errlen, e.what());
  }
}

}  // extern "C"
