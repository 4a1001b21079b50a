/* lammps_stub.h — a minimal stand-in for the LAMMPS (mid-2016) core classes that the translation units of
 * /root/reference include.  TEST INFRASTRUCTURE ONLY (see ../oracle.h).
 *
 * Purpose: compile the reference's OWN sources (pair_buck*_intel.cpp, pppm_intel.cpp, fix_nve_intel.cpp, unchanged, from
 * where they lie under /root/reference) into oracle/_ref/libref.so so that the restatement in oracle/ can be checked
 * against the code it restates.  Nothing here is copied from LAMMPS: the classes carry only the members and methods the
 * reference touches, with the behaviour SURVEY.md Appendix A states for them ([UPSTREAM] contracts).  Every header name
 * the reference includes (atom.h, force.h, pair_buck.h, pppm.h, ...) is a one-line forwarder to this file.
 */
#ifndef B200MD_REF_LAMMPS_STUB_H
#define B200MD_REF_LAMMPS_STUB_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mpi.h"

#define FLERR __FILE__, __LINE__
#define SBBITS 30
#define NEIGHMASK 0x3FFFFFFF

typedef double FFT_SCALAR;   /* the reference is built without -DFFT_SINGLE */
#define MPI_FFT_SCALAR MPI_DOUBLE

namespace LAMMPS_NS {

typedef int tagint;
typedef int64_t bigint;

class Memory;
class Error;
class Atom;
class Comm;
class Force;
class Neighbor;
class Modify;
class Update;
class Domain;
class Group;
class Pair;
class KSpace;
class Fix;
class NeighList;
class NeighRequest;

class LAMMPS {
 public:
  Memory *memory = nullptr;
  Error *error = nullptr;
  Atom *atom = nullptr;
  Comm *comm = nullptr;
  Force *force = nullptr;
  Neighbor *neighbor = nullptr;
  Modify *modify = nullptr;
  Update *update = nullptr;
  Domain *domain = nullptr;
  Group *group = nullptr;
  MPI_Comm world = 0;
};

class Pointers {
 public:
  explicit Pointers(LAMMPS *ptr)
      : lmp(ptr), memory(ptr->memory), error(ptr->error), atom(ptr->atom), comm(ptr->comm), force(ptr->force),
        neighbor(ptr->neighbor), modify(ptr->modify), update(ptr->update), domain(ptr->domain), group(ptr->group),
        world(ptr->world) {}
  virtual ~Pointers() {}

 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Atom *&atom;
  Comm *&comm;
  Force *&force;
  Neighbor *&neighbor;
  Modify *&modify;
  Update *&update;
  Domain *&domain;
  Group *&group;
  MPI_Comm &world;
};

/* error->all / error->one abort the run in LAMMPS; here they unwind to the harness, which reports the message */
class Error {
 public:
  void all(const char *, int, const char *msg) { throw std::runtime_error(msg); }
  void one(const char *, int, const char *msg) { throw std::runtime_error(msg); }
  void warning(const char *, int, const char *, int = 1) {}
};

/* Memory::create / destroy: contiguous storage behind the row pointers, like the stock allocator */
class Memory {
 public:
  template <class T>
  T *create(T *&a, int n, const char *) {
    a = (T *)aligned(sizeof(T) * (size_t)(n > 0 ? n : 1));
    return a;
  }
  template <class T>
  void destroy(T *&a) {
    free(a);
    a = nullptr;
  }
  template <class T>
  T **create(T **&a, int n1, int n2, const char *) {
    T *data = (T *)aligned(sizeof(T) * (size_t)n1 * n2);
    a = (T **)malloc(sizeof(T *) * (size_t)(n1 > 0 ? n1 : 1));
    for (int i = 0; i < n1; i++) a[i] = data + (size_t)i * n2;
    return a;
  }
  template <class T>
  void destroy(T **&a) {
    if (!a) return;
    free(a[0]);
    free(a);
    a = nullptr;
  }
  /* a[n1lo..n1hi][n2lo..n2hi][n3lo..n3hi] over one contiguous block */
  template <class T>
  T ***create3d_offset(T ***&a, int n1lo, int n1hi, int n2lo, int n2hi, int n3lo, int n3hi, const char *) {
    const int n1 = n1hi - n1lo + 1, n2 = n2hi - n2lo + 1, n3 = n3hi - n3lo + 1;
    T *data = (T *)aligned(sizeof(T) * (size_t)n1 * n2 * n3);
    memset(data, 0, sizeof(T) * (size_t)n1 * n2 * n3);
    T **plane = (T **)malloc(sizeof(T *) * (size_t)n1 * n2);
    T ***arr = (T ***)malloc(sizeof(T **) * (size_t)n1);
    for (int i = 0; i < n1; i++) {
      arr[i] = plane + (size_t)i * n2 - n2lo;
      for (int j = 0; j < n2; j++) plane[(size_t)i * n2 + j] = data + ((size_t)i * n2 + j) * n3 - n3lo;
    }
    a = arr - n1lo;
    return a;
  }
  template <class T>
  void destroy3d_offset(T ***&a, int n1lo, int n2lo, int n3lo) {
    if (!a) return;
    T ***arr = a + n1lo;
    T **plane = arr[0] + n2lo;
    free(plane[0] + n3lo);
    free(plane);
    free(arr);
    a = nullptr;
  }

 private:
  static void *aligned(size_t bytes) {
    void *p = nullptr;
    if (posix_memalign(&p, 64, bytes ? bytes : 64)) throw std::bad_alloc();
    return p;
  }
};

class Atom {
 public:
  int nlocal = 0, nghost = 0, nmax = 0, ntypes = 0;
  bigint natoms = 0;
  double **x = nullptr, **v = nullptr, **f = nullptr;
  double *q = nullptr, *mass = nullptr, *rmass = nullptr;
  int *type = nullptr, *mask = nullptr;
  tagint *tag = nullptr;
  int torque = 0;          /* no torque arrays: IntelBuffers::get_stride keeps one force row per atom */
  int firstgroup = -1, nfirst = 0;
};

class Comm {
 public:
  int me = 0, nprocs = 1, nthreads = 1;
};

class Force {
 public:
  int newton = 1, newton_pair = 1;
  double special_lj[4] = {1, 0, 0, 0}, special_coul[4] = {1, 0, 0, 0};
  double qqrd2e = 1.0, qqr2e = 1.0, ftm2v = 1.0, boltz = 1.0;
  Pair *pair = nullptr;
  KSpace *kspace = nullptr;
};

class Group {
 public:
  int bitmask[2] = {1, 2};
};

class NeighRequest {
 public:
  int intel = 0, half = 1, full = 0;
};

class NeighList {
 public:
  int inum = 0;
  int *ilist = nullptr, *numneigh = nullptr;
  int **firstneigh = nullptr;
  int *stencil = nullptr;
  int maxlocal = 0;
  int get_maxlocal() { return maxlocal; }
};

class Neighbor {
 public:
  int ago = 0, oneatom = 2000, maxhead = 0;
  int every = 1, delay = 10, dist_check = 1;   /* neigh_modify (stock defaults); read by the B200 binding only */
  double skin = 0.0;
  int nrequest = 0;
  NeighRequest **requests = nullptr;
  int request(void *) {
    requests = (NeighRequest **)realloc(requests, sizeof(NeighRequest *) * (size_t)(nrequest + 1));
    requests[nrequest] = new NeighRequest();
    return nrequest++;
  }
};

class Update {
 public:
  double dt = 0.0;
  bigint ntimestep = 0;
};

class Domain {
 public:
  int triclinic = 0;
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {1, 1, 1};
  double boxlo_lamda[3] = {0, 0, 0};
  int periodicity[3] = {1, 1, 1};              /* read by the B200 binding only */
  double xprd = 1, yprd = 1, zprd = 1;
  double prd[3] = {1, 1, 1};
  void x2lamda(int) {}
  void lamda2x(int) {}
};

namespace FixConst {
enum { INITIAL_INTEGRATE = 1 << 0, FINAL_INTEGRATE = 1 << 6, PRE_REVERSE = 1 << 4 };
}

class Fix : public Pointers {
 public:
  char *id = nullptr;
  int igroup = 0, groupbit = 1;
  Fix(LAMMPS *lmp, int narg, char **arg) : Pointers(lmp) {
    if (narg > 0 && arg) id = strdup(arg[0]);
  }
  virtual ~Fix() { free(id); }
  virtual void init() {}
  virtual void setup(int) {}
  virtual void initial_integrate(int) {}
  virtual void final_integrate() {}
  virtual void reset_dt() {}
  virtual double memory_usage() { return 0.0; }
};

/* stock fix nve: the constructor and setup do nothing the reference depends on; dtv / dtf are set by reset_dt */
class FixNVE : public Fix {
 public:
  FixNVE(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg) {}
  virtual void init() {   /* stock FixNVE::init */
    dtv = update->dt;
    dtf = 0.5 * update->dt * force->ftm2v;
  }
  virtual void setup(int) {}
  virtual double memory_usage() { return 0.0; }

 protected:
  double dtv = 0.0, dtf = 0.0;
};

class Modify {
 public:
  int nfix = 0;
  Fix **fix = nullptr;
  int find_fix(const char *id) {
    for (int i = 0; i < nfix; i++)
      if (fix[i]->id && strcmp(fix[i]->id, id) == 0) return i;
    return -1;
  }
};

namespace Suffix {
enum { NONE = 0, OPT = 1 << 0, GPU = 1 << 1, OMP = 1 << 2, INTEL = 1 << 3 };
}

namespace MathConst {
static const double THIRD = 1.0 / 3.0;
static const double MY_PI = 3.14159265358979323846;
static const double MY_2PI = 6.28318530717958647692;
static const double MY_3PI = 9.42477796076937971538;
static const double MY_4PI = 12.56637061435917295384;
static const double MY_PI2 = 1.57079632679489661923;
static const double MY_PI4 = 0.78539816339744830962;
static const double MY_PIS = 1.77245385090551602729;
}

namespace MathSpecial {
inline double square(double x) { return x * x; }
}

/* ---- Pair base (SURVEY App. A.1 ev_setup, A.2) ------------------------------------------------------------------ */
class Pair : public Pointers {
 public:
  double eng_vdwl = 0, eng_coul = 0, virial[6] = {0, 0, 0, 0, 0, 0};
  double *eatom = nullptr, **vatom = nullptr;
  double **cutsq = nullptr;
  int **setflag = nullptr;
  int suffix_flag = 0;
  int evflag = 0, eflag_either = 0, eflag_global = 0, eflag_atom = 0;
  int vflag_either = 0, vflag_global = 0, vflag_atom = 0, vflag_fdotr = 0;
  int no_virial_fdotr = 0;
  int offset_flag = 0, mix_flag = 0;
  NeighList *list = nullptr;
  virtual void *extract(const char *, int &dim) { dim = 0; return nullptr; }   /* stock Pair::extract (KSpace styles read cut-offs and coefficients through it) */
  /* Coulomb / dispersion tables (Pair::init_tables products; filled by the harness from host-built tables) */
  int ncoultablebits = 0, ncoulmask = 0, ncoulshiftbits = 0;
  double tabinner = 0, tabinnersq = 0;
  double *rtable = nullptr, *drtable = nullptr, *ftable = nullptr, *dftable = nullptr, *ctable = nullptr,
         *dctable = nullptr, *etable = nullptr, *detable = nullptr;
  int ndisptablebits = 0, ndispmask = 0, ndispshiftbits = 0;
  double tabinnerdispsq = 0;
  double *rdisptable = nullptr, *drdisptable = nullptr, *fdisptable = nullptr, *dfdisptable = nullptr,
         *edisptable = nullptr, *dedisptable = nullptr;
  int maxeatom = 0;

  explicit Pair(LAMMPS *lmp) : Pointers(lmp) {}
  virtual ~Pair() {}
  virtual void compute(int, int) = 0;
  virtual void init_style() {}
  virtual double init_one(int, int) { return 0.0; }
  int fdotr_is_set() { return vflag_fdotr; }

  void ev_setup(int eflag, int vflag) {
    evflag = 1;
    eflag_either = eflag;
    eflag_global = eflag % 2;
    eflag_atom = eflag / 2;
    vflag_either = vflag;
    vflag_global = vflag % 4;
    vflag_atom = vflag / 4;
    if (eflag_atom && atom->nlocal + atom->nghost > maxeatom) {
      maxeatom = atom->nlocal + atom->nghost;
      free(eatom);
      eatom = (double *)malloc(sizeof(double) * (size_t)maxeatom);
    }
    if (eflag_global) eng_vdwl = eng_coul = 0.0;
    if (vflag_global)
      for (int i = 0; i < 6; i++) virial[i] = 0.0;
    if (eflag_atom)
      for (int i = 0; i < atom->nlocal + atom->nghost; i++) eatom[i] = 0.0;
    if (vflag_global == 2 && no_virial_fdotr == 0) vflag_fdotr = 1;
    else vflag_fdotr = 0;
  }

 protected:
  void alloc2(double **&a, int n) {
    memory->create(a, n, n, "pair");
    for (int i = 0; i < n * n; i++) a[0][i] = 0.0;
  }
  void base_allocate(int n) {
    memory->create(setflag, n, n, "pair:setflag");
    for (int i = 0; i < n * n; i++) setflag[0][i] = 0;
    alloc2(cutsq, n);
  }
};

/* pair_style buck (A.2): init_one derives rhoinv, buck1, buck2, offset; no mixing */
class PairBuck : public Pair {
 public:
  explicit PairBuck(LAMMPS *lmp) : Pair(lmp) {}
  void allocate() {
    const int n = atom->ntypes + 1;
    base_allocate(n);
    alloc2(cut, n); alloc2(a, n); alloc2(rho, n); alloc2(c, n);
    alloc2(rhoinv, n); alloc2(buck1, n); alloc2(buck2, n); alloc2(offset, n);
  }
  virtual void init_style() { neighbor->request(this); }
  virtual double init_one(int i, int j) {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    rhoinv[i][j] = 1.0 / rho[i][j];
    buck1[i][j] = a[i][j] / rho[i][j];
    buck2[i][j] = 6.0 * c[i][j];
    if (offset_flag) {
      const double rexp = exp(-cut[i][j] / rho[i][j]);
      offset[i][j] = a[i][j] * rexp - c[i][j] / pow(cut[i][j], 6.0);
    } else offset[i][j] = 0.0;
    a[j][i] = a[i][j];
    c[j][i] = c[i][j];
    rhoinv[j][i] = rhoinv[i][j];
    buck1[j][i] = buck1[i][j];
    buck2[j][i] = buck2[i][j];
    offset[j][i] = offset[i][j];
    return cut[i][j];
  }
  double cut_global = 0;
  double **cut = nullptr, **a = nullptr, **rho = nullptr, **c = nullptr;
  double **rhoinv = nullptr, **buck1 = nullptr, **buck2 = nullptr, **offset = nullptr;
};

/* pair_style buck/coul/cut (A.2) */
class PairBuckCoulCut : public Pair {
 public:
  explicit PairBuckCoulCut(LAMMPS *lmp) : Pair(lmp) {}
  void allocate() {
    const int n = atom->ntypes + 1;
    base_allocate(n);
    alloc2(cut_lj, n); alloc2(cut_ljsq, n); alloc2(cut_coul, n); alloc2(cut_coulsq, n);
    alloc2(a, n); alloc2(rho, n); alloc2(c, n);
    alloc2(rhoinv, n); alloc2(buck1, n); alloc2(buck2, n); alloc2(offset, n);
  }
  virtual void init_style() { neighbor->request(this); }
  virtual double init_one(int i, int j) {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    const double cut = cut_lj[i][j] > cut_coul[i][j] ? cut_lj[i][j] : cut_coul[i][j];
    cut_ljsq[i][j] = cut_lj[i][j] * cut_lj[i][j];
    cut_coulsq[i][j] = cut_coul[i][j] * cut_coul[i][j];
    rhoinv[i][j] = 1.0 / rho[i][j];
    buck1[i][j] = a[i][j] / rho[i][j];
    buck2[i][j] = 6.0 * c[i][j];
    if (offset_flag) {
      const double rexp = exp(-cut_lj[i][j] / rho[i][j]);
      offset[i][j] = a[i][j] * rexp - c[i][j] / pow(cut_lj[i][j], 6.0);
    } else offset[i][j] = 0.0;
    cut_ljsq[j][i] = cut_ljsq[i][j];
    cut_coulsq[j][i] = cut_coulsq[i][j];
    a[j][i] = a[i][j];
    c[j][i] = c[i][j];
    rhoinv[j][i] = rhoinv[i][j];
    buck1[j][i] = buck1[i][j];
    buck2[j][i] = buck2[i][j];
    offset[j][i] = offset[i][j];
    return cut;
  }
  double cut_lj_global = 0, cut_coul_global = 0;
  double **cut_lj = nullptr, **cut_ljsq = nullptr, **cut_coul = nullptr, **cut_coulsq = nullptr;
  double **a = nullptr, **rho = nullptr, **c = nullptr;
  double **rhoinv = nullptr, **buck1 = nullptr, **buck2 = nullptr, **offset = nullptr;
};

/* pair_style buck/coul/long (A.2): one global Coulomb cut-off, g_ewald from the kspace style */
class PairBuckCoulLong : public Pair {
 public:
  explicit PairBuckCoulLong(LAMMPS *lmp) : Pair(lmp) {}
  void allocate() {
    const int n = atom->ntypes + 1;
    base_allocate(n);
    alloc2(cut_lj, n); alloc2(cut_ljsq, n);
    alloc2(a, n); alloc2(rho, n); alloc2(c, n);
    alloc2(rhoinv, n); alloc2(buck1, n); alloc2(buck2, n); alloc2(offset, n);
  }
  virtual void init_style();   /* needs KSpace: defined after it */
  virtual double init_one(int i, int j) {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    const double cut = cut_lj[i][j] > cut_coul ? cut_lj[i][j] : cut_coul;
    cut_ljsq[i][j] = cut_lj[i][j] * cut_lj[i][j];
    rhoinv[i][j] = 1.0 / rho[i][j];
    buck1[i][j] = a[i][j] / rho[i][j];
    buck2[i][j] = 6.0 * c[i][j];
    if (offset_flag) {
      const double rexp = exp(-cut_lj[i][j] / rho[i][j]);
      offset[i][j] = a[i][j] * rexp - c[i][j] / pow(cut_lj[i][j], 6.0);
    } else offset[i][j] = 0.0;
    cut_ljsq[j][i] = cut_ljsq[i][j];
    a[j][i] = a[i][j];
    c[j][i] = c[i][j];
    rhoinv[j][i] = rhoinv[i][j];
    buck1[j][i] = buck1[i][j];
    buck2[j][i] = buck2[i][j];
    offset[j][i] = offset[i][j];
    return cut;
  }
  double cut_lj_global = 0, cut_coul = 0, cut_coulsq = 0, g_ewald = 0;
  double **cut_lj = nullptr, **cut_ljsq = nullptr;
  double **a = nullptr, **rho = nullptr, **c = nullptr;
  double **rhoinv = nullptr, **buck1 = nullptr, **buck2 = nullptr, **offset = nullptr;
};

/* pair_style buck/long/coul/long (A.2): member names as the reference uses them
 * (pair_buck_long_coul_long_intel.cpp:267,351,414,600-641) */
class PairBuckLongCoulLong : public Pair {
 public:
  explicit PairBuckLongCoulLong(LAMMPS *lmp) : Pair(lmp) {}
  void allocate() {
    const int n = atom->ntypes + 1;
    base_allocate(n);
    alloc2(cut_buck, n); alloc2(cut_bucksq, n); alloc2(cut_buck_read, n);
    alloc2(buck_a, n); alloc2(buck_c, n); alloc2(buck_rho, n);
    alloc2(buck_a_read, n); alloc2(buck_c_read, n); alloc2(buck_rho_read, n);
    alloc2(rhoinv, n); alloc2(buck1, n); alloc2(buck2, n); alloc2(offset, n);
  }
  virtual void compute(int, int) {}
  virtual void init_style();
  virtual double init_one(int i, int j) {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    if (ewald_order & (1 << 6)) cut_buck[i][j] = cut_buck_global;
    else cut_buck[i][j] = cut_buck_read[i][j];
    buck_a[i][j] = buck_a_read[i][j];
    buck_c[i][j] = buck_c_read[i][j];
    buck_rho[i][j] = buck_rho_read[i][j];
    const double cut = cut_buck[i][j] > cut_coul ? cut_buck[i][j] : cut_coul;
    cutsq[i][j] = cut * cut;
    cut_bucksq[i][j] = cut_buck[i][j] * cut_buck[i][j];
    buck1[i][j] = buck_a[i][j] / buck_rho[i][j];
    buck2[i][j] = 6.0 * buck_c[i][j];
    rhoinv[i][j] = 1.0 / buck_rho[i][j];
    if (offset_flag) {
      const double rexp = exp(-cut_buck[i][j] / buck_rho[i][j]);
      offset[i][j] = buck_a[i][j] * rexp - buck_c[i][j] / pow(cut_buck[i][j], 6.0);
    } else offset[i][j] = 0.0;
    cutsq[j][i] = cutsq[i][j];
    cut_bucksq[j][i] = cut_bucksq[i][j];
    buck_a[j][i] = buck_a[i][j];
    buck_c[j][i] = buck_c[i][j];
    rhoinv[j][i] = rhoinv[i][j];
    buck1[j][i] = buck1[i][j];
    buck2[j][i] = buck2[i][j];
    offset[j][i] = offset[i][j];
    return cut;
  }
  int ewald_order = 0, ewald_off = 0;
  double cut_buck_global = 0, cut_coul = 0, cut_coulsq = 0, g_ewald = 0, g_ewald_6 = 0;
  double **cut_buck = nullptr, **cut_bucksq = nullptr, **cut_buck_read = nullptr;
  double **buck_a = nullptr, **buck_c = nullptr, **buck_rho = nullptr;
  double **buck_a_read = nullptr, **buck_c_read = nullptr, **buck_rho_read = nullptr;
  double **rhoinv = nullptr, **buck1 = nullptr, **buck2 = nullptr, **offset = nullptr;
};

/* pair_style lj/long/coul/long (A.2 names; the reference reads ewald_order, cut_ljsq, lj1..lj4, offset, g_ewald,
 * g_ewald_6 and the tables: pair_lj_long_coul_long_intel.cpp:111-112,479,571,817-835) */
class PairLJLongCoulLong : public Pair {
 public:
  explicit PairLJLongCoulLong(LAMMPS *lmp) : Pair(lmp) {}
  void allocate() {
    const int n = atom->ntypes + 1;
    base_allocate(n);
    alloc2(cut_lj, n); alloc2(cut_ljsq, n); alloc2(cut_lj_read, n);
    alloc2(epsilon, n); alloc2(sigma, n); alloc2(epsilon_read, n); alloc2(sigma_read, n);
    alloc2(lj1, n); alloc2(lj2, n); alloc2(lj3, n); alloc2(lj4, n); alloc2(offset, n);
  }
  virtual void compute(int, int) {}
  virtual void init_style();
  virtual double init_one(int i, int j) {
    if (setflag[i][j] == 0) error->all(FLERR, "All pair coeffs are not set");
    epsilon[i][j] = epsilon_read[i][j];
    sigma[i][j] = sigma_read[i][j];
    if (ewald_order & (1 << 6)) cut_lj[i][j] = cut_lj_global;
    else cut_lj[i][j] = cut_lj_read[i][j];
    const double cut = cut_lj[i][j] > cut_coul ? cut_lj[i][j] : cut_coul;
    cutsq[i][j] = cut * cut;
    cut_ljsq[i][j] = cut_lj[i][j] * cut_lj[i][j];
    lj1[i][j] = 48.0 * epsilon[i][j] * pow(sigma[i][j], 12.0);
    lj2[i][j] = 24.0 * epsilon[i][j] * pow(sigma[i][j], 6.0);
    lj3[i][j] = 4.0 * epsilon[i][j] * pow(sigma[i][j], 12.0);
    lj4[i][j] = 4.0 * epsilon[i][j] * pow(sigma[i][j], 6.0);
    if (offset_flag && cut_lj[i][j] > 0.0) {
      const double ratio = sigma[i][j] / cut_lj[i][j];
      offset[i][j] = 4.0 * epsilon[i][j] * (pow(ratio, 12.0) - pow(ratio, 6.0));
    } else offset[i][j] = 0.0;
    cutsq[j][i] = cutsq[i][j];
    cut_ljsq[j][i] = cut_ljsq[i][j];
    lj1[j][i] = lj1[i][j];
    lj2[j][i] = lj2[i][j];
    lj3[j][i] = lj3[i][j];
    lj4[j][i] = lj4[i][j];
    offset[j][i] = offset[i][j];
    return cut;
  }
  int ewald_order = 0, ewald_off = 0;
  double cut_lj_global = 0, cut_coul = 0, cut_coulsq = 0, g_ewald = 0, g_ewald_6 = 0;
  double **cut_lj = nullptr, **cut_ljsq = nullptr, **cut_lj_read = nullptr;
  double **epsilon = nullptr, **sigma = nullptr, **epsilon_read = nullptr, **sigma_read = nullptr;
  double **lj1 = nullptr, **lj2 = nullptr, **lj3 = nullptr, **lj4 = nullptr, **offset = nullptr;
};

/* ---- KSpace / PPPM base (A.5) ------------------------------------------------------------------------------------ */
class KSpace : public Pointers {
 public:
  double energy = 0, virial[6] = {0, 0, 0, 0, 0, 0};
  double *eatom = nullptr, **vatom = nullptr;
  double g_ewald = 0, g_ewald_6 = 0, scale = 1.0, qqrd2e = 1.0;
  int order = 5, order_6 = 5;
  int differentiation_flag = 0, slabflag = 0, triclinic = 0, tip4pflag = 0;
  double slab_volfactor = 1.0;   /* kspace_modify slab (the reference reads slabflag only; the B200 binding passes this on) */
  int suffix_flag = 0;
  int evflag = 0, evflag_atom = 0, eflag_either = 0, eflag_global = 0, eflag_atom = 0;
  int vflag_either = 0, vflag_global = 0, vflag_atom = 0;
  double qsum = 0, qsqsum = 0, q2 = 0;
  KSpace(LAMMPS *lmp, int, char **) : Pointers(lmp) {}
  virtual ~KSpace() {}
  virtual void init() {}
  virtual void setup() {}
  virtual void compute(int, int) = 0;
  void ev_setup(int eflag, int vflag) {
    evflag = 1;
    eflag_either = eflag;
    eflag_global = eflag % 2;
    eflag_atom = eflag / 2;
    vflag_either = vflag;
    vflag_global = vflag % 4;
    vflag_atom = vflag / 4;
    if (eflag_atom || vflag_atom) evflag_atom = 1;
    else evflag_atom = 0;
    if (eflag_global) energy = 0.0;
    if (vflag_global)
      for (int i = 0; i < 6; i++) virial[i] = 0.0;
  }
};

inline void PairBuckCoulLong::init_style() {
  cut_coulsq = cut_coul * cut_coul;
  g_ewald = force->kspace->g_ewald;
  neighbor->request(this);
}
inline void PairLJLongCoulLong::init_style() {
  cut_coulsq = cut_coul * cut_coul;
  if (force->kspace) {
    g_ewald = force->kspace->g_ewald;
    g_ewald_6 = force->kspace->g_ewald_6;
  }
  neighbor->request(this);
}
inline void PairBuckLongCoulLong::init_style() {
  cut_coulsq = cut_coul * cut_coul;
  if (force->kspace) {
    g_ewald = force->kspace->g_ewald;
    g_ewald_6 = force->kspace->g_ewald_6;
  }
  neighbor->request(this);
}

/* FFT3d::compute(in, out, flag): unnormalised complex 3-D transform, flag = +1 is exp(+ikx) (the FFTW_BACKWARD /
 * KISS-inverse plan of stock LAMMPS; pinned by the Ewald known-answer test), -1 the other sign.  The transform itself is
 * upstream code: the stub calls the oracle's mixed-radix FFT (oracle/fft.cpp). */
extern "C" void orc_fft3d(double *data, int nx, int ny, int nz, int dir, int nthreads);
class FFT3d {
 public:
  int nx, ny, nz, nthreads = 1;
  FFT3d(int nx_, int ny_, int nz_) : nx(nx_), ny(ny_), nz(nz_) {}
  void compute(FFT_SCALAR *in, FFT_SCALAR *out, int flag) {
    if (in != out) memcpy(out, in, sizeof(FFT_SCALAR) * 2 * (size_t)nx * ny * nz);
    orc_fft3d(out, nx, ny, nz, flag, nthreads);
  }
};

/* Remap brick -> FFT decomposition: the identity on one rank */
class Remap {
 public:
  size_t n;
  explicit Remap(size_t n_) : n(n_) {}
  void perform(FFT_SCALAR *in, FFT_SCALAR *out, FFT_SCALAR *) {
    if (in != out) memcpy(out, in, sizeof(FFT_SCALAR) * n);
  }
};

class PPPM;
/* GridComm on one periodic rank: reverse = add ghost cells into their periodic owners (x, then y, then z, like the
 * staged six-way swap), forward = fill ghost cells from their owners (A.5) */
class GridComm {
 public:
  PPPM *p = nullptr;
  void ghost_notify() {}
  void setup() {}
  void reverse_comm(KSpace *, int which);
  void forward_comm(KSpace *, int which);
};

class PPPM : public KSpace {
 public:
  PPPM(LAMMPS *lmp, int narg, char **arg) : KSpace(lmp, narg, arg) {}
  virtual ~PPPM() {}
  virtual void init() {}                 /* the harness fills the state below (upstream products come from oracle/) */
  virtual void compute(int, int) {}
  virtual void brick2fft() {}

  /* state read by pppm_intel.cpp (A.5) */
  int nx_pppm = 0, ny_pppm = 0, nz_pppm = 0;
  int nlower = 0, nupper = 0;
  double shift = 0, shiftone = 0;
  double *boxlo = nullptr;
  double delxinv = 0, delyinv = 0, delzinv = 0, delvolinv = 0, volume = 0;
  int nxlo_in = 0, nylo_in = 0, nzlo_in = 0, nxhi_in = 0, nyhi_in = 0, nzhi_in = 0;
  int nxlo_out = 0, nylo_out = 0, nzlo_out = 0, nxhi_out = 0, nyhi_out = 0, nzhi_out = 0;
  int nxlo_fft = 0, nylo_fft = 0, nzlo_fft = 0, nxhi_fft = 0, nyhi_fft = 0, nzhi_fft = 0;
  int ngrid = 0, nfft = 0, nfft_both = 0;
  FFT_SCALAR ***density_brick = nullptr, ***vdx_brick = nullptr, ***vdy_brick = nullptr, ***vdz_brick = nullptr,
             ***u_brick = nullptr;
  FFT_SCALAR *density_fft = nullptr, *work1 = nullptr, *work2 = nullptr;
  double *greensfn = nullptr, **vg = nullptr, *fkx = nullptr, *fky = nullptr, *fkz = nullptr;
  FFT_SCALAR **rho_coeff = nullptr, **drho_coeff = nullptr;   /* [order][nlower..nupper] */
  double sf_coeff[6] = {0, 0, 0, 0, 0, 0};
  int **part2grid = nullptr;
  int nmax = 0;
  bigint natoms_original = -1;
  int peratom_allocate_flag = 0;
  FFT3d *fft1 = nullptr, *fft2 = nullptr;
  Remap *remap = nullptr;
  GridComm *cg = nullptr, *cg_peratom = nullptr;

 protected:
  void qsum_qsq() {
    qsum = qsqsum = 0.0;
    for (int i = 0; i < atom->nlocal; i++) {
      qsum += atom->q[i];
      qsqsum += atom->q[i] * atom->q[i];
    }
    q2 = qsqsum * force->qqrd2e;
  }
  /* the stub has no per-atom tallies, triclinic boxes or slab correction: the harness never asks for them */
  void allocate_peratom() { error->all(FLERR, "ref stub: per-atom k-space tallies are not available"); }
  void poisson_peratom() { error->all(FLERR, "ref stub: per-atom k-space tallies are not available"); }
  void fieldforce_peratom() { error->all(FLERR, "ref stub: per-atom k-space tallies are not available"); }
  void poisson_ik_triclinic() { error->all(FLERR, "ref stub: triclinic boxes are not available"); }
  void slabcorr() { error->all(FLERR, "ref stub: slab correction is not available"); }
};

inline void GridComm::reverse_comm(KSpace *, int) {
  FFT_SCALAR ***d = p->density_brick;
  const int nx = p->nx_pppm, ny = p->ny_pppm, nz = p->nz_pppm;
  auto pm = [](int a, int n) { int r = a % n; return r < 0 ? r + n : r; };
  for (int mz = p->nzlo_out; mz <= p->nzhi_out; mz++)
    for (int my = p->nylo_out; my <= p->nyhi_out; my++)
      for (int mx = p->nxlo_out; mx <= p->nxhi_out; mx++) {
        if (mx >= 0 && mx < nx) continue;
        d[mz][my][pm(mx, nx)] += d[mz][my][mx];
      }
  for (int mz = p->nzlo_out; mz <= p->nzhi_out; mz++)
    for (int my = p->nylo_out; my <= p->nyhi_out; my++) {
      if (my >= 0 && my < ny) continue;
      for (int mx = 0; mx < nx; mx++) d[mz][pm(my, ny)][mx] += d[mz][my][mx];
    }
  for (int mz = p->nzlo_out; mz <= p->nzhi_out; mz++) {
    if (mz >= 0 && mz < nz) continue;
    for (int my = 0; my < ny; my++)
      for (int mx = 0; mx < nx; mx++) d[pm(mz, nz)][my][mx] += d[mz][my][mx];
  }
}

inline void GridComm::forward_comm(KSpace *, int) {
  const int nx = p->nx_pppm, ny = p->ny_pppm, nz = p->nz_pppm;
  auto pm = [](int a, int n) { int r = a % n; return r < 0 ? r + n : r; };
  FFT_SCALAR ***bricks[3];
  int nb = 0;
  if (p->differentiation_flag == 1) bricks[nb++] = p->u_brick;
  else { bricks[nb++] = p->vdx_brick; bricks[nb++] = p->vdy_brick; bricks[nb++] = p->vdz_brick; }
  for (int b = 0; b < nb; b++)
    for (int mz = p->nzlo_out; mz <= p->nzhi_out; mz++)
      for (int my = p->nylo_out; my <= p->nyhi_out; my++)
        for (int mx = p->nxlo_out; mx <= p->nxhi_out; mx++) {
          if (mx >= 0 && mx < nx && my >= 0 && my < ny && mz >= 0 && mz < nz) continue;
          bricks[b][mz][my][mx] = bricks[b][pm(mz, nz)][pm(my, ny)][pm(mx, nx)];
        }
}

}  // namespace LAMMPS_NS

#endif
