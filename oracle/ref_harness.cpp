// ref_harness.cpp — drives the reference's OWN translation units (compiled unchanged from /root/reference into
// oracle/_ref/libref.so, recipe: oracle/Makefile.ref) on caller-supplied inputs.  TEST INFRASTRUCTURE ONLY.
//
// What runs from the reference: PairBuck{,CoulCut,CoulLong,LongCoulLong}Intel and PairLJLongCoulLongIntel::init_style /
// pack_force_const / compute / eval<> (pair_buck_intel.cpp:48-443 and the four siblings), IntelBuffers::thr_pack (intel_buffers.h:185-203),
// PPPMIntel::compute / particle_map / make_rho / brick2fft / poisson_ik / poisson_ad / fieldforce_ik / fieldforce_ad
// (pppm_intel.cpp:104-1054), FixNVEIntel::setup / reset_dt / initial_integrate / final_integrate (fix_nve_intel.cpp).
// What is NOT in the reference and is supplied by ref_shim/ (stand-ins stating SURVEY.md App. A): the LAMMPS core
// classes, FixIntel, the IP_PRE_* macros, the base pair classes' init_one, GridComm / Remap / FFT3d on one rank.
// Upstream PPPM products (Green's function, vg, fk*, rho_coeff, sf_coeff) come in through orc_pppm_state.
//
// The entry points mirror oracle.h so that a test can hand the same arrays to both and compare bit for bit.
#include <pthread.h>

#include <chrono>
#include <string>

#include "fix_intel.h"
#include "intel_buffers_impl.h"
#include "fix_nve_intel.h"
#include "pair_buck_coul_cut_intel.h"
#include "pair_buck_coul_long_intel.h"
#include "pair_buck_intel.h"
#include "pair_buck_long_coul_long_intel.h"
#include "pair_lj_long_coul_long_intel.h"
#include "pppm_intel.h"

#include "oracle.h"

using namespace LAMMPS_NS;

namespace {

double g_last_seconds = 0.0;   // wall time of the last compute() call of the reference (ref_last_seconds)
struct Stopwatch {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~Stopwatch() { g_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

struct StubKSpace : public KSpace {
  StubKSpace(LAMMPS *lmp) : KSpace(lmp, 0, nullptr) {}
  void compute(int, int) {}
};

// one LAMMPS "instance": the objects the reference reaches through Pointers
struct World {
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
  NeighList list;
  FixIntel *fix = nullptr;
  Fix *fixes[2] = {nullptr, nullptr};
  std::vector<double> xbuf, vbuf, fbuf;
  std::vector<double *> xrow, vrow, frow;

  World(int prec, int nthreads) {
    lmp.memory = &memory; lmp.error = &error; lmp.atom = &atom; lmp.comm = &comm; lmp.force = &force;
    lmp.neighbor = &neighbor; lmp.modify = &modify; lmp.update = &update; lmp.domain = &domain; lmp.group = &group;
    comm.nthreads = nthreads;
    static char id[] = "package_intel";
    char *arg[1] = {id};
    const int mode = prec == ORC_PREC_MIXED ? FixIntel::PREC_MODE_MIXED : FixIntel::PREC_MODE_DOUBLE;
    fix = new FixIntel(&lmp, 1, arg, mode);
    fixes[0] = fix;
    modify.fix = fixes;
    modify.nfix = 1;
  }
  ~World() { delete fix; }

  static void rows(std::vector<double> &buf, std::vector<double *> &row, int n) {
    buf.assign((size_t)3 * (n + 1), 0.0);
    row.resize((size_t)n + 1);
    for (int i = 0; i <= n; i++) row[i] = buf.data() + (size_t)3 * i;
  }
  void set_atoms(int nlocal, int nall, const double *x, int *type, double *q, int ntypes) {
    atom.nlocal = nlocal;
    atom.nghost = nall - nlocal;
    atom.nmax = nall + 1;
    atom.natoms = nlocal;
    atom.ntypes = ntypes;
    rows(xbuf, xrow, nall);
    rows(fbuf, frow, nall);
    if (x) memcpy(xbuf.data(), x, sizeof(double) * 3 * (size_t)nall);
    atom.x = xrow.data();
    atom.f = frow.data();
    atom.type = type;
    atom.q = q;
  }
};

template <class flt_t, class acc_t>
IntelBuffers<flt_t, acc_t> *buffers_of(FixIntel *fix);
template <> IntelBuffers<double, double> *buffers_of<double, double>(FixIntel *fix) { return fix->get_double_buffers(); }
template <> IntelBuffers<float, double> *buffers_of<float, double>(FixIntel *fix) { return fix->get_mixed_buffers(); }

// what the intel neighbour build (upstream, not in the reference) leaves behind: sized buffers, x/type/q packed at
// ago = 0, and the flat packed list
template <class flt_t, class acc_t>
void stage_buffers(World &w, int nlocal, int nall, const int *numneigh, const long *offsets, const int *entries) {
  auto *b = buffers_of<flt_t, acc_t>(w.fix);
  b->grow(nall, nlocal, w.comm.nthreads, 0);
  b->zero_ev();   // FixIntel::setup does this once per run; evals without EFLAG leave ev_global[0..1] untouched
  b->thr_pack(0, nall, 0);
  if (numneigh) {
    w.list.inum = nlocal;
    w.list.numneigh = const_cast<int *>(numneigh);
    w.list.maxlocal = (int)offsets[nlocal];
    b->grow_nbor(&w.list, nlocal, w.comm.nthreads, 0);
    int *fn = b->firstneigh(&w.list), *cn = b->cnumneigh(&w.list);
    memcpy(fn, entries, sizeof(int) * (size_t)offsets[nlocal]);
    for (int i = 0; i < nlocal; i++) cn[i] = (int)offsets[i];
  }
}

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

// run fn on a thread with a large stack: PPPMIntel::make_rho puts nthreads * ngrid doubles on the stack
// (pppm_intel.cpp:418)
struct BigStack {
  void *(*fn)(void *);
  void *arg;
};
int run_big_stack(void *(*fn)(void *), void *arg, size_t bytes) {
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, bytes);
  pthread_t th;
  if (pthread_create(&th, &at, fn, arg)) return -1;
  void *ret = nullptr;
  pthread_join(th, &ret);
  pthread_attr_destroy(&at);
  return (int)(intptr_t)ret;
}

}  // namespace

extern "C" {

// PairBuck*Intel::compute on a packed CSR list (the layout of IntelBuffers::firstneigh / cnumneigh, intel_buffers.h:
// 145-146).  Same arguments as orc_pair_eval plus the raw coefficients A, rho, C, cut_lj, cut_coul ((ntypes+1)^2 each) the
// base classes' init_one derives everything from; p supplies the scalars and tables.  f is [nall][4] (w = eatom).
int ref_pair_eval(int style, int prec, int eflag, int vflag, int eatom, int newton, int nlocal, int nall,
                  const double *x, const int *type, const double *q, const int *numneigh, const long *offsets,
                  const int *entries, const double *A, const double *rho, const double *C, const double *cut_lj,
                  const double *cut_coul, int offset_flag, double skin, const orc_pair_params *p, double *f, double *ev,
                  int nthreads, char *err, int errlen) {
  try {
    if (nthreads < 1) nthreads = 1;
#if !defined(_OPENMP)
    nthreads = 1;
#else
    omp_set_num_threads(nthreads);
#endif
    World w(prec, nthreads);
    const int ntypes = p->ntypes, tp1 = ntypes + 1;
    std::vector<int> typev(type, type + nall);
    typev.push_back(1);
    std::vector<double> qv((size_t)nall + 1, 0.0);
    if (q) memcpy(qv.data(), q, sizeof(double) * (size_t)nall);
    w.set_atoms(nlocal, nall, x, typev.data(), q ? qv.data() : nullptr, ntypes);
    w.force.newton_pair = newton;
    w.force.qqrd2e = p->qqrd2e;
    for (int k = 0; k < 4; k++) { w.force.special_lj[k] = p->special_lj[k]; w.force.special_coul[k] = p->special_coul[k]; }
    w.neighbor.skin = skin;
    StubKSpace ks(&w.lmp);
    ks.g_ewald = p->g_ewald;
    ks.g_ewald_6 = p->g_ewald_6;
    w.force.kspace = &ks;

    Pair *pair = nullptr;
    auto set_tables = [&](Pair *pr) {
      pr->ncoultablebits = p->ncoultablebits;
      pr->ncoulmask = p->ncoulmask;
      pr->ncoulshiftbits = p->ncoulshiftbits;
      pr->tabinnersq = p->tabinnersq;
      pr->rtable = const_cast<double *>(p->rtable); pr->drtable = const_cast<double *>(p->drtable);
      pr->ftable = const_cast<double *>(p->ftable); pr->dftable = const_cast<double *>(p->dftable);
      pr->etable = const_cast<double *>(p->etable); pr->detable = const_cast<double *>(p->detable);
      pr->ctable = const_cast<double *>(p->ctable); pr->dctable = const_cast<double *>(p->dctable);
      pr->ndisptablebits = p->ndisptablebits;
      pr->ndispmask = p->ndispmask;
      pr->ndispshiftbits = p->ndispshiftbits;
      pr->tabinnerdispsq = p->tabinnerdispsq;
      pr->rdisptable = const_cast<double *>(p->rdisptable); pr->drdisptable = const_cast<double *>(p->drdisptable);
      pr->fdisptable = const_cast<double *>(p->fdisptable); pr->dfdisptable = const_cast<double *>(p->dfdisptable);
      pr->edisptable = const_cast<double *>(p->edisptable); pr->dedisptable = const_cast<double *>(p->dedisptable);
    };
    auto all_set = [&](Pair *pr) {
      for (int i = 1; i < tp1; i++)
        for (int j = 1; j < tp1; j++) pr->setflag[i][j] = 1;
      pr->offset_flag = offset_flag;
    };
    if (style == ORC_BUCK) {
      auto *pb = new PairBuckIntel(&w.lmp);
      pb->allocate();
      fill2(pb->a, A, tp1); fill2(pb->rho, rho, tp1); fill2(pb->c, C, tp1); fill2(pb->cut, cut_lj, tp1);
      pair = pb;
    } else if (style == ORC_BUCK_COUL_CUT) {
      auto *pb = new PairBuckCoulCutIntel(&w.lmp);
      pb->allocate();
      fill2(pb->a, A, tp1); fill2(pb->rho, rho, tp1); fill2(pb->c, C, tp1);
      fill2(pb->cut_lj, cut_lj, tp1); fill2(pb->cut_coul, cut_coul, tp1);
      pair = pb;
    } else if (style == ORC_BUCK_COUL_LONG) {
      auto *pb = new PairBuckCoulLongIntel(&w.lmp);
      pb->allocate();
      fill2(pb->a, A, tp1); fill2(pb->rho, rho, tp1); fill2(pb->c, C, tp1); fill2(pb->cut_lj, cut_lj, tp1);
      pb->cut_coul = cut_coul[tp1 + 1];   // one global Coulomb cut-off
      set_tables(pb);
      pair = pb;
    } else if (style == ORC_BUCK_LONG_COUL_LONG) {
      auto *pb = new PairBuckLongCoulLongIntel(&w.lmp);
      pb->allocate();
      fill2(pb->buck_a_read, A, tp1); fill2(pb->buck_rho_read, rho, tp1); fill2(pb->buck_c_read, C, tp1);
      fill2(pb->cut_buck_read, cut_lj, tp1);
      pb->cut_buck_global = cut_lj[tp1 + 1];
      pb->cut_coul = cut_coul ? cut_coul[tp1 + 1] : 0.0;
      pb->ewald_order = (p->order1 ? 1 << 1 : 0) | (p->order6 ? 1 << 6 : 0);
      set_tables(pb);
      pair = pb;
    } else if (style == ORC_LJ_LONG_COUL_LONG) {   // A = epsilon, rho = sigma
      auto *pb = new PairLJLongCoulLongIntel(&w.lmp);
      pb->allocate();
      fill2(pb->epsilon_read, A, tp1); fill2(pb->sigma_read, rho, tp1); fill2(pb->cut_lj_read, cut_lj, tp1);
      pb->cut_lj_global = cut_lj[tp1 + 1];
      pb->cut_coul = cut_coul ? cut_coul[tp1 + 1] : 0.0;
      pb->ewald_order = (p->order1 ? 1 << 1 : 0) | (p->order6 ? 1 << 6 : 0);
      set_tables(pb);
      pair = pb;
    } else return fail(err, errlen, "ref_pair_eval: unknown style");
    all_set(pair);
    w.force.pair = pair;
    pair->list = &w.list;
    pair->init_style();   // the reference's: PairX::init_style, FixIntel lookup, pack_force_const

    if (prec == ORC_PREC_MIXED) stage_buffers<float, double>(w, nlocal, nall, numneigh, offsets, entries);
    else stage_buffers<double, double>(w, nlocal, nall, numneigh, offsets, entries);
    w.neighbor.ago = 1;   // a step after the build: compute() repacks positions itself (thr_pack, x only)

    {
      Stopwatch sw;
      pair->compute(eflag ? (eatom ? 3 : 1) : 0, vflag);
    }

    for (int i = 0; i < nall; i++) {
      f[4 * (size_t)i + 0] = w.atom.f[i][0];
      f[4 * (size_t)i + 1] = w.atom.f[i][1];
      f[4 * (size_t)i + 2] = w.atom.f[i][2];
      f[4 * (size_t)i + 3] = (eflag && eatom) ? pair->eatom[i] : 0.0;
    }
    ev[0] = pair->eng_vdwl;
    ev[1] = pair->eng_coul;
    for (int k = 0; k < 6; k++) ev[2 + k] = pair->virial[k];
    delete pair;
    return 0;
  } catch (const std::exception &e) {
    return fail(err, errlen, e.what());
  }
}

// FixNVEIntel: which = 0 initial_integrate, 1 final_integrate (after init + setup, i.e. reset_dt for ntypes > 1).
// rmass / ingroup may be NULL (per-type masses, group all).  x, v [nlocal][3] are updated in place.
int ref_nve(int which, int nlocal, int ntypes, double *x, double *v, const double *f, const int *type,
            const double *mass /*[ntypes+1]*/, const double *rmass, const int *ingroup, double dt, double ftm2v,
            char *err, int errlen) {
  try {
    World w(ORC_PREC_DOUBLE, 1);
    std::vector<int> typev(type, type + nlocal), mask((size_t)nlocal + 1, 1);
    std::vector<double> massv(mass, mass + ntypes + 1), rm;
    w.set_atoms(nlocal, nlocal, x, typev.data(), nullptr, ntypes);
    World::rows(w.vbuf, w.vrow, nlocal);
    memcpy(w.vbuf.data(), v, sizeof(double) * 3 * (size_t)nlocal);
    memcpy(w.fbuf.data(), f, sizeof(double) * 3 * (size_t)nlocal);
    w.atom.v = w.vrow.data();
    w.atom.mass = massv.data();
    if (rmass) { rm.assign(rmass, rmass + nlocal); w.atom.rmass = rm.data(); }
    w.update.dt = dt;
    w.force.ftm2v = ftm2v;
    static char a0[] = "1", a1[] = "all", a2[] = "nve/intel";
    char *arg[3] = {a0, a1, a2};
    FixNVEIntel nve(&w.lmp, 3, arg);
    if (ingroup) {   // a sub-group: bit 1 of the mask
      nve.igroup = 1;
      nve.groupbit = 2;
      for (int i = 0; i < nlocal; i++) mask[i] = 1 | (ingroup[i] ? 2 : 0);
    }
    w.atom.mask = mask.data();
    nve.init();
    nve.setup(0);
    if (ingroup || rmass) nve.reset_dt();   // setup() only calls it for ntypes > 1 (fix_nve_intel.cpp:52)
    w.neighbor.ago = 1;
    {
      Stopwatch sw;
      if (which == 0) nve.initial_integrate(0);
      else nve.final_integrate();
    }
    memcpy(x, w.xbuf.data(), sizeof(double) * 3 * (size_t)nlocal);
    memcpy(v, w.vbuf.data(), sizeof(double) * 3 * (size_t)nlocal);
    return 0;
  } catch (const std::exception &e) {
    return fail(err, errlen, e.what());
  }
}

}  // extern "C"

// ---- PPPMIntel -----------------------------------------------------------------------------------------------------
namespace {

struct RefPPPM : public PPPMIntel {
  RefPPPM(LAMMPS *lmp, int narg, char **arg) : PPPMIntel(lmp, narg, arg) {}
  using PPPMIntel::fix;
  std::vector<double *> vgrow, rcrow, drcrow;
  std::vector<double> rc, drc, gf, vgv, kx, ky, kz;

  void load(const orc_pppm_state &s) {
    order = s.order;
    differentiation_flag = s.diff_ad;
    nx_pppm = s.nx; ny_pppm = s.ny; nz_pppm = s.nz;
    nlower = s.nlower; nupper = s.nupper;
    shift = s.shift; shiftone = s.shiftone;
    g_ewald = s.g_ewald; qqrd2e = s.qqrd2e; scale = s.scale; volume = s.volume;
    delxinv = s.delinv[0]; delyinv = s.delinv[1]; delzinv = s.delinv[2]; delvolinv = s.delvolinv;
    nxlo_in = nylo_in = nzlo_in = nxlo_fft = nylo_fft = nzlo_fft = 0;
    nxhi_in = nxhi_fft = s.nx - 1; nyhi_in = nyhi_fft = s.ny - 1; nzhi_in = nzhi_fft = s.nz - 1;
    nxlo_out = s.lo_out[0]; nylo_out = s.lo_out[1]; nzlo_out = s.lo_out[2];
    nxhi_out = s.hi_out[0]; nyhi_out = s.hi_out[1]; nzhi_out = s.hi_out[2];
    ngrid = (nxhi_out - nxlo_out + 1) * (nyhi_out - nylo_out + 1) * (nzhi_out - nzlo_out + 1);
    nfft = s.nx * s.ny * s.nz;
    nfft_both = nfft;
    memory->create3d_offset(density_brick, nzlo_out, nzhi_out, nylo_out, nyhi_out, nxlo_out, nxhi_out, "density");
    if (s.diff_ad) memory->create3d_offset(u_brick, nzlo_out, nzhi_out, nylo_out, nyhi_out, nxlo_out, nxhi_out, "u");
    else {
      memory->create3d_offset(vdx_brick, nzlo_out, nzhi_out, nylo_out, nyhi_out, nxlo_out, nxhi_out, "vdx");
      memory->create3d_offset(vdy_brick, nzlo_out, nzhi_out, nylo_out, nyhi_out, nxlo_out, nxhi_out, "vdy");
      memory->create3d_offset(vdz_brick, nzlo_out, nzhi_out, nylo_out, nyhi_out, nxlo_out, nxhi_out, "vdz");
    }
    memory->create(density_fft, nfft, "density_fft");
    memory->create(work1, 2 * nfft, "work1");
    memory->create(work2, 2 * nfft, "work2");
    gf.assign(s.greensfn, s.greensfn + nfft);
    greensfn = gf.data();
    vgv.assign(s.vg, s.vg + 6 * (size_t)nfft);
    vgrow.resize(nfft);
    for (int i = 0; i < nfft; i++) vgrow[i] = vgv.data() + 6 * (size_t)i;
    vg = vgrow.data();
    kx.assign(s.fkx, s.fkx + s.nx); ky.assign(s.fky, s.fky + s.ny); kz.assign(s.fkz, s.fkz + s.nz);
    fkx = kx.data(); fky = ky.data(); fkz = kz.data();
    rc.assign(s.rho_coeff, s.rho_coeff + order * order);
    drc.assign(s.drho_coeff, s.drho_coeff + order * order);
    rcrow.resize(order); drcrow.resize(order);
    for (int l = 0; l < order; l++) {   // rho_coeff[l][k], k in [nlower, nupper]
      rcrow[l] = rc.data() + (size_t)l * order - nlower;
      drcrow[l] = drc.data() + (size_t)l * order - nlower;
    }
    rho_coeff = rcrow.data();
    drho_coeff = drcrow.data();
    for (int k = 0; k < 6; k++) sf_coeff[k] = s.sf_coeff[k];
    fft1 = new FFT3d(s.nx, s.ny, s.nz);
    fft2 = new FFT3d(s.nx, s.ny, s.nz);
    fft1->nthreads = fft2->nthreads = comm->nthreads;
    remap = new Remap((size_t)nfft);
    cg = new GridComm();
    cg->p = this;
    natoms_original = -1;
  }
  ~RefPPPM() {
    memory->destroy3d_offset(density_brick, nzlo_out, nylo_out, nxlo_out);
    memory->destroy3d_offset(u_brick, nzlo_out, nylo_out, nxlo_out);
    memory->destroy3d_offset(vdx_brick, nzlo_out, nylo_out, nxlo_out);
    memory->destroy3d_offset(vdy_brick, nzlo_out, nylo_out, nxlo_out);
    memory->destroy3d_offset(vdz_brick, nzlo_out, nylo_out, nxlo_out);
    memory->destroy(density_fft); memory->destroy(work1); memory->destroy(work2);
    memory->destroy(part2grid);
    delete fft1; delete fft2; delete remap; delete cg;
  }
};

struct PppmCall {
  const orc_pppm_state *st;
  int prec, nlocal;
  const double *x, *q;
  int eflag, vflag;
  double *f, *energy, *virial, *density_fft_out, *field_out;
  int nthreads;
  char *err;
  int errlen;
};

void *pppm_thread(void *vp) {
  PppmCall &c = *(PppmCall *)vp;
  try {
    int nthreads = c.nthreads < 1 ? 1 : c.nthreads;
#if !defined(_OPENMP)
    nthreads = 1;
#else
    omp_set_num_threads(nthreads);
#endif
    World w(c.prec, nthreads);
    const orc_pppm_state &s = *c.st;
    std::vector<int> typev((size_t)c.nlocal + 1, 1);
    std::vector<double> qv(c.q, c.q + c.nlocal);
    qv.push_back(0.0);
    w.set_atoms(c.nlocal, c.nlocal, c.x, typev.data(), qv.data(), 1);
    w.force.qqrd2e = s.qqrd2e;
    for (int d = 0; d < 3; d++) {
      w.domain.boxlo[d] = s.boxlo[d];
      w.domain.prd[d] = s.prd[d];
      w.domain.boxhi[d] = s.boxlo[d] + s.prd[d];
    }
    static char a0[] = "1.0e-4";
    char *arg[1] = {a0};
    RefPPPM pp(&w.lmp, 1, arg);
    pp.load(s);
    pp.init();   // PPPMIntel::init: FixIntel lookup, order check
    if (c.prec == ORC_PREC_MIXED) stage_buffers<float, double>(w, c.nlocal, c.nlocal, nullptr, nullptr, nullptr);
    else stage_buffers<double, double>(w, c.nlocal, c.nlocal, nullptr, nullptr, nullptr);
    {
      Stopwatch sw;
      pp.compute(c.eflag, c.vflag);
    }
    // fieldforce adds into IntelBuffers::_f (thread 0's array): that is the k-space force
    if (c.prec == ORC_PREC_MIXED) {
      auto *fb = w.fix->get_mixed_buffers()->get_f();
      for (int i = 0; i < c.nlocal; i++) { c.f[3 * i] += fb[i].x; c.f[3 * i + 1] += fb[i].y; c.f[3 * i + 2] += fb[i].z; }
    } else {
      auto *fb = w.fix->get_double_buffers()->get_f();
      for (int i = 0; i < c.nlocal; i++) { c.f[3 * i] += fb[i].x; c.f[3 * i + 1] += fb[i].y; c.f[3 * i + 2] += fb[i].z; }
    }
    if (c.energy) *c.energy = pp.energy;
    if (c.virial) for (int k = 0; k < 6; k++) c.virial[k] = pp.virial[k];
    const long nfft = (long)s.nx * s.ny * s.nz;
    if (c.density_fft_out) memcpy(c.density_fft_out, pp.density_fft, sizeof(double) * nfft);
    if (c.field_out) {
      FFT_SCALAR ***b[3] = {s.diff_ad ? pp.u_brick : pp.vdx_brick, pp.vdy_brick, pp.vdz_brick};
      for (int d = 0; d < (s.diff_ad ? 1 : 3); d++)
        for (int k = 0; k < s.nz; k++)
          for (int j = 0; j < s.ny; j++)
            for (int i = 0; i < s.nx; i++) c.field_out[d * nfft + ((long)k * s.ny + j) * s.nx + i] = b[d][k][j][i];
    }
    return (void *)0;
  } catch (const std::exception &e) {
    fail(c.err, c.errlen, e.what());
    return (void *)1;
  }
}

}  // namespace

extern "C" {

// PPPMIntel::compute on one periodic rank.  f [nlocal][3] is accumulated into (+=), like orc_pppm_compute.
// density_fft_out [nfft] and field_out [3][nfft] (ik) / [nfft] (ad) may be NULL.
int ref_pppm_compute(const orc_pppm_state *st, int prec, int nlocal, const double *x, const double *q, int eflag,
                     int vflag, double *f, double *energy, double *virial, double *density_fft_out, double *field_out,
                     int nthreads, char *err, int errlen) {
  PppmCall c{st, prec, nlocal, x, q, eflag, vflag, f, energy, virial, density_fft_out, field_out, nthreads, err, errlen};
  const long ngrid = (long)(st->hi_out[0] - st->lo_out[0] + 1) * (st->hi_out[1] - st->lo_out[1] + 1) *
                     (st->hi_out[2] - st->lo_out[2] + 1);
  const size_t stack = (size_t)(nthreads < 1 ? 1 : nthreads) * ngrid * sizeof(double) + (64u << 20);
  const int rc = run_big_stack(pppm_thread, &c, stack);
  if (rc < 0) return fail(err, errlen, "ref_pppm_compute: could not start the worker thread");
  return rc;
}

/* seconds the reference's own compute() / initial_integrate() / final_integrate() took in the last ref_pair_eval /
 * ref_pppm_compute / ref_nve call (set-up excluded) */
double ref_last_seconds(void) { return g_last_seconds; }

int ref_has_openmp(void) {
#if defined(_OPENMP)
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"
