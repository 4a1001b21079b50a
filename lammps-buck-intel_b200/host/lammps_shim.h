// lammps_shim.h — the slice of the LAMMPS object model that the /intel styles of HPAC/lammps-buck-intel touch.
//
// The reference is a plug-in: its classes derive from upstream LAMMPS base classes (Pair, KSpace, Fix) and read
// upstream singletons (atom, force, domain, neighbor, update, error).  None of those ship with the reference
// (SURVEY.md §2.2, App. A), so this header restates the members the hot path uses — same names, same meaning — for
// the stand-alone host layer of this repo.  Inside a real LAMMPS tree the real headers take their place; what carries
// over is the body of each class's C-ABI marshalling, with two mechanical differences:
// per-atom arrays here are flat std::vector<double> ([n][3] contiguous, what atom->x[0] points at in LAMMPS) and
// per-type-pair tables flat [(ntypes+1)^2] vectors instead of double**.
//
//   reference use                                         member here
//   atom->x/v/f/q/type/mass/nlocal/ntypes                 Atom        (fix_nve_intel.cpp:64-72, pair_buck_intel.cpp:86)
//   force->qqrd2e/ftm2v/special_lj/special_coul/kspace    Force       (pair_buck_coul_long_intel.cpp:157,507)
//   domain->boxlo/boxhi/prd/periodicity                   Domain      (pppm_intel.cpp:153,342)
//   neighbor->skin/every/delay/dist_check                 Neighbor    (pair_buck_intel.cpp:370,399-409)
//   update->dt/ntimestep                                  Update      (fix_nve_intel.cpp:130)
//   error->all(FLERR,msg) / error->one                    Error       (pppm_intel.cpp:73,87,342,385)
#pragma once
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace LAMMPS_NS {

#define FLERR __FILE__, __LINE__

class LAMMPSException : public std::runtime_error {
 public:
  explicit LAMMPSException(const std::string &m) : std::runtime_error(m) {}
};

class Error {
 public:
  // error->all aborts every rank, error->one the calling rank; stand-alone both raise
  [[noreturn]] void all(const char *file, int line, const std::string &msg) { fail(file, line, msg); }
  [[noreturn]] void one(const char *file, int line, const std::string &msg) { fail(file, line, msg); }
  void warning(const char *, int, const std::string &msg) { std::fprintf(stderr, "WARNING: %s\n", msg.c_str()); }

 private:
  [[noreturn]] static void fail(const char *file, int line, const std::string &msg) {
    const char *base = file;
    for (const char *p = file; *p; p++)
      if (*p == '/') base = p + 1;
    throw LAMMPSException("ERROR: " + msg + " (" + base + ":" + std::to_string(line) + ")");
  }
};

class Atom {
 public:
  int nlocal = 0, ntypes = 0;
  long natoms = 0;
  int q_flag = 0;          // atom_style charge
  int molecule_flag = 0;   // atom_style full: the data file carries a molecule id column
  // contiguous [nlocal][3] storage, as atom->x[0] (fix_nve_intel.cpp:64-66)
  std::vector<double> x, v, f;
  std::vector<double> q;
  std::vector<int> type;
  std::vector<int> molecule; // molecule id per atom (atom_style full); only `delete_atoms ... mol yes` reads it
  std::vector<double> mass;  // [ntypes+1]
  std::vector<int> mask;     // group bits per atom (bit 0 = all); empty = every atom in `all` only
  std::vector<double> rmass; // per-atom masses when rmass_flag (fix_nve_intel.cpp:148-156)
  int rmass_flag = 0;
  std::vector<int> mass_setflag;
  // bond topology of a molecular data file (atom_style full; 0-based atom pairs): only the special-bond lists derived
  // from it reach the hot path (bits 30-31 of the neighbour-list entries)
  std::vector<int> bonds;            // [nbonds][2]
  std::vector<int> nspecial, special; // [nlocal][3] cumulative 1-2 / 1-3 / 1-4 counts, [nlocal][maxspecial]
  int maxspecial = 0;
};

class Pair;
class KSpace;

class Force {
 public:
  double qqrd2e = 1.0, ftm2v = 1.0, boltz = 1.0, mvv2e = 1.0;
  double qqr2e = 1.0, qelectron = 1.0, angstrom = 1.0;   // KSpace accuracy scaling (two_charge_force)
  double special_lj[4] = {1.0, 0.0, 0.0, 0.0}, special_coul[4] = {1.0, 0.0, 0.0, 0.0};
  int newton_pair = 1;
  Pair *pair = nullptr;
  KSpace *kspace = nullptr;
};

class Domain {
 public:
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {1, 1, 1}, prd[3] = {1, 1, 1};
  int periodicity[3] = {1, 1, 1};
  int triclinic = 0;
  bool box_exist = false;
  // lattice (for `region ... block` in lattice units)
  double lattice_a = 1.0;
  void set_box(const double lo[3], const double hi[3]) {
    for (int d = 0; d < 3; d++) { boxlo[d] = lo[d]; boxhi[d] = hi[d]; prd[d] = hi[d] - lo[d]; }
    box_exist = true;
  }
};

class Neighbor {
 public:
  double skin = 0.3;
  int every = 1, delay = 10, dist_check = 1;   // stock defaults (neigh_modify)
};

class Update {
 public:
  double dt = 0.005;
  long ntimestep = 0;
  std::string unit_style = "lj";
};

class FixIntel;

class LAMMPS {
 public:
  Atom *atom;
  Force *force;
  Domain *domain;
  Neighbor *neighbor;
  Update *update;
  Error *error;
  FixIntel *fix_intel = nullptr;   // modify->fix[ifix] of style "INTEL" (`package intel`, pair_buck_intel.cpp:372-376)
  int suffix_enable = 0;
  bool dry_run = false;            // host-side initialisation only (lmp_b200 -dry-run): no device context exists
  std::string suffix;
  LAMMPS() : atom(new Atom), force(new Force), domain(new Domain), neighbor(new Neighbor), update(new Update),
             error(new Error) {}
  ~LAMMPS();
  LAMMPS(const LAMMPS &) = delete;
  LAMMPS &operator=(const LAMMPS &) = delete;
};

class Pointers {
 public:
  explicit Pointers(LAMMPS *l)
      : lmp(l), atom(l->atom), force(l->force), domain(l->domain), neighbor(l->neighbor), update(l->update),
        error(l->error) {}
  virtual ~Pointers() {}

 protected:
  LAMMPS *lmp;
  Atom *&atom;
  Force *&force;
  Domain *&domain;
  Neighbor *&neighbor;
  Update *&update;
  Error *&error;
};

namespace Suffix { enum { NONE = 0, INTEL = 1 << 3 }; }

// upstream Pair: the part of the interface the reference overrides or reads
class Pair : protected Pointers {
 public:
  double eng_vdwl = 0.0, eng_coul = 0.0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  std::vector<double> eatom;
  int suffix_flag = Suffix::NONE;
  int ewaldflag = 0, dispersionflag = 0, offset_flag = 0;
  enum { GEOMETRIC = 0, ARITHMETIC = 1, SIXTHPOWER = 2 };   // pair.h [UPSTREAM]
  int mix_flag = GEOMETRIC;       // pair_modify mix geometric|arithmetic
  int ncoultablebits = 12;        // pair_modify table 12 (stock default)
  int ndisptablebits = 12;
  double tabinner = 1.4142135623730951, tabinner_disp = 1.4142135623730951;
  double cutforce = 0.0;
  explicit Pair(LAMMPS *l) : Pointers(l) {}
  virtual void compute(int eflag, int vflag) = 0;
  virtual void settings(int narg, char **arg) = 0;
  virtual void coeff(int narg, char **arg) = 0;
  virtual void init_style() = 0;
  virtual double init_one(int i, int j) = 0;
  virtual void init();   // loops init_one over type pairs (Pair::init)
  // the init_one loop alone: what PPPMDisp::init calls pair->init() for before it reads the mixed coefficients
  virtual void init_all_pairs() {}
  virtual void *extract(const char *, int &dim) { dim = 0; return nullptr; }

 protected:
  int allocated = 0;
  std::vector<int> setflag;      // [(ntypes+1)^2]
  std::vector<double> cutsq;
  int eflag_either = 0, vflag_either = 0, eflag_global = 0, eflag_atom = 0, vflag_global = 0, vflag_fdotr = 0;
  void ev_setup(int eflag, int vflag);
  int tp1() const { return atom->ntypes + 1; }
};

class KSpace : protected Pointers {
 public:
  double energy = 0.0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  std::vector<double> eatom, vatom;   // [nlocal], [nlocal][6]: filled when compute() is asked for eflag & 2 / vflag & 4
  double g_ewald = 0.0, g_ewald_6 = 0.0;
  int order = 5, order_6 = 5;
  int nx_pppm = 0, ny_pppm = 0, nz_pppm = 0;
  int nx_pppm_6 = 0, ny_pppm_6 = 0, nz_pppm_6 = 0;
  int differentiation_flag = 0;   // kspace_modify diff ik|ad
  int gridflag = 0, gewaldflag = 0;   // kspace_modify mesh / gewald
  int gridflag_6 = 0, gewaldflag_6 = 0;
  int mixflag = 0;                // kspace_modify mix/disp pair (0) | geom (1) | none (2)
  double accuracy = 0.0, accuracy_relative = 0.0, accuracy_absolute = -1.0, two_charge_force = 0.0;
  double accuracy_real_6 = -1.0, accuracy_kspace_6 = -1.0;   // kspace_modify force/disp/real, force/disp/kspace
  double scale = 1.0;
  int slabflag = 0;                 // kspace_modify slab
  double slab_volfactor = 1.0;
  int suffix_flag = Suffix::NONE;
  explicit KSpace(LAMMPS *l) : Pointers(l) {}
  virtual void init() = 0;
  virtual void setup() = 0;
  virtual void compute(int eflag, int vflag) = 0;
  virtual void modify_params(int narg, char **arg);
};

class Fix : protected Pointers {
 public:
  std::string id, style;
  int igroup = 0, groupbit = 1;   // group all unless the fix command names another group
  explicit Fix(LAMMPS *l) : Pointers(l) {}
  virtual void init() {}
  virtual void setup(int) {}
  virtual void initial_integrate(int) {}
  virtual void final_integrate() {}
  virtual void reset_dt() {}
};

}  // namespace LAMMPS_NS
