// lmp_b200.cpp — minimal input-script driver for the commands of examples/in.buck, in.buck_big, in.buck_coul_cut and
// in.buck_coul_long, so that the shipped scripts run unchanged on the B200 styles:
//
//     lmp_b200 -in in.buck_coul_long -sf intel -pk intel 0 mode double [-var x 2] [-kspace pppm]
//
// It plays the role of the upstream LAMMPS executable around the reference's plug-in classes: Input (command parsing),
// Lattice/CreateAtoms/ReadData/Replicate, Velocity (RanPark stream), Verlet::setup/run in the stock order
// (initial_integrate -> neighbor decide -> pair -> kspace -> final_integrate) and Thermo one-style output.
// All forces, neighbour lists and integration run through the classes of this directory, i.e. through the C ABI.
//   -dry-run     parse + host-side init only (no device): prints a JSON summary of what would run
//   -styles      list the registered pair / kspace / fix styles (the PairStyle(key,Class) lines of the headers)
//   -host-step   plug-in deployment: positions/forces cross PCIe every step (FixIntel::resident = 0)
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

#include "fix_nve_intel.h"
#include "pair_buck_coul_intel.h"
#include "pair_lj_long_coul_long_intel.h"
#include "pppm_disp_intel.h"
#include "pppm_intel.h"

using namespace LAMMPS_NS;

namespace {

// ---- RanPark (Park-Miller minimal standard), as used by `velocity ... create` --------------------------------
class RanPark {
 public:
  explicit RanPark(int s) : seed(s) {}
  double uniform() {
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
    const int k = seed / IQ;
    seed = IA * (seed - k * IQ) - IR * k;
    if (seed < 0) seed += IM;
    return (1.0 / IM) * seed;
  }
  // `loop geom`: seed hashed from the atom's coordinates
  void reset(int ibase, const double *coord) {
    unsigned int hash = 0;
    const char *str = (const char *)&ibase;
    for (size_t i = 0; i < sizeof(int); i++) { hash += str[i]; hash += (hash << 10); hash ^= (hash >> 6); }
    str = (const char *)coord;
    for (size_t i = 0; i < 3 * sizeof(double); i++) { hash += str[i]; hash += (hash << 10); hash ^= (hash >> 6); }
    hash += (hash << 3);
    hash ^= (hash >> 11);
    hash += (hash << 15);
    seed = hash & 0x7ffffff;
    if (!seed) seed = 1;
    for (int i = 0; i < 5; i++) uniform();
  }

 private:
  int seed;
};

// ---- style registries, filled the way Force::Force / Modify::Modify of stock LAMMPS fill theirs: the style headers are
// included once more with PAIR_CLASS / KSPACE_CLASS / FIX_CLASS defined, so that only their PairStyle(key,Class) ...
// lines are seen (pair_buck_intel.h:18-22 of the reference) ------------------------------------------------------
typedef Pair *(*PairCreator)(LAMMPS *);
typedef KSpace *(*KSpaceCreator)(LAMMPS *, int, char **);
typedef Fix *(*FixCreator)(LAMMPS *, int, char **);
template <class T> Pair *pair_creator(LAMMPS *l) { return new T(l); }
template <class T> KSpace *kspace_creator(LAMMPS *l, int narg, char **arg) { return new T(l, narg, arg); }
template <class T> Fix *fix_creator(LAMMPS *l, int narg, char **arg) { return new T(l, narg, arg); }

struct Styles {
  std::map<std::string, PairCreator> pair;
  std::map<std::string, KSpaceCreator> kspace;
  std::map<std::string, FixCreator> fix;
  Styles() {
#define PAIR_CLASS
#define PairStyle(key, Class) pair[#key] = &pair_creator<Class>;
#include "style_pair.h"
#undef PairStyle
#undef PAIR_CLASS
#define KSPACE_CLASS
#define KSpaceStyle(key, Class) kspace[#key] = &kspace_creator<Class>;
#include "style_kspace.h"
#undef KSpaceStyle
#undef KSPACE_CLASS
#define FIX_CLASS
#define FixStyle(key, Class) fix[#key] = &fix_creator<Class>;
#include "style_fix.h"
#undef FixStyle
#undef FIX_CLASS
  }
};

struct Script {
  LAMMPS lmp;
  Styles styles;
  std::map<std::string, std::string> vars;
  std::unique_ptr<Pair> pair;
  std::unique_ptr<KSpace> kspace;
  std::unique_ptr<FixNVEIntel> nve;
  std::map<std::string, int> groups{{"all", 0}};   // group ID -> bit
  std::string pair_style_name, kspace_style_name, kspace_override;
  std::vector<std::string> kspace_args;
  int thermo_every = 0;
  bool suffix_intel = false, dry_run = false, host_step = false, echo = false;
  int device = 0, prec_mode = FixIntel::PREC_MODE_DOUBLE;
  double nktv2p = 1.0;
  std::string dir;   // directory of the input script (read_data paths are relative to it)
  // lattice
  double lattice_a = 1.0;
  std::vector<std::array<double, 3>> basis;
  struct Block { double lo[3], hi[3]; };
  std::map<std::string, Block> regions;             // region ID block ... (lattice units; scale 1 without a lattice)
  std::vector<std::string> skipped_fixes;           // -dry-run only: fix styles outside the pair / k-space / nve path
  bool warned_dump = false;
  bool dt_set = false;
  double last_thermo[6] = {0, 0, 0, 0, 0, 0};
  long nbuilds_reported = 0;

  [[noreturn]] void fail(const std::string &m) { lmp.error->all(FLERR, m); }

  // ---- variables: `variable x index 1`, `variable xx equal 20*$x`; $x and ${xx} substitution -------------------
  double eval_expr(const std::string &e) {   // products / quotients / sums of numbers (enough for the scripts)
    double acc = 0.0, term = 1.0;
    char op = '*', sign = '+';
    size_t i = 0;
    bool have = false;
    auto flush_term = [&]() { acc += sign == '+' ? term : -term; term = 1.0; op = '*'; have = false; };
    while (i < e.size()) {
      if (std::isspace((unsigned char)e[i])) { i++; continue; }
      if ((e[i] == '+' || e[i] == '-') && have) { flush_term(); sign = e[i]; i++; continue; }
      if (e[i] == '*' || e[i] == '/') { op = e[i]; i++; continue; }
      char *end;
      const double v = std::strtod(e.c_str() + i, &end);
      if (end == e.c_str() + i) fail("Invalid syntax in variable formula: " + e);
      term = op == '*' ? term * v : term / v;
      have = true;
      i = end - e.c_str();
    }
    if (have) flush_term();
    return acc;
  }
  std::string substitute(const std::string &line) {
    std::string out;
    for (size_t i = 0; i < line.size(); i++) {
      if (line[i] != '$') { out += line[i]; continue; }
      std::string name;
      if (i + 1 < line.size() && line[i + 1] == '{') {
        const size_t c = line.find('}', i);
        if (c == std::string::npos) fail("Invalid variable name");
        name = line.substr(i + 2, c - i - 2);
        i = c;
      } else if (i + 1 < line.size()) { name = line.substr(i + 1, 1); i++; }
      auto it = vars.find(name);
      if (it == vars.end()) fail("Substitution for illegal variable " + name);
      out += it->second;
    }
    return out;
  }

  void set_units(const std::string &u) {
    Force *f = lmp.force;
    lmp.update->unit_style = u;
    if (u == "lj") {
      f->boltz = 1.0; f->mvv2e = 1.0; f->ftm2v = 1.0; f->qqrd2e = f->qqr2e = 1.0; nktv2p = 1.0;
      f->qelectron = 1.0; f->angstrom = 1.0;
      if (!dt_set) lmp.update->dt = 0.005;
      lmp.neighbor->skin = 0.3;
    } else if (u == "metal") {
      f->boltz = 8.617343e-5; f->mvv2e = 1.0364269e-4; f->ftm2v = 1.0 / 1.0364269e-4;
      f->qqrd2e = f->qqr2e = 14.399645; nktv2p = 1.6021765e6;
      f->qelectron = 1.0; f->angstrom = 1.0;
      if (!dt_set) lmp.update->dt = 0.001;
      lmp.neighbor->skin = 2.0;
    } else if (u == "real") {
      f->boltz = 0.0019872067; f->mvv2e = 48.88821291 * 48.88821291; f->ftm2v = 1.0 / 48.88821291 / 48.88821291;
      f->qqrd2e = f->qqr2e = 332.06371; nktv2p = 68568.415;
      f->qelectron = 1.0; f->angstrom = 1.0;
      if (!dt_set) lmp.update->dt = 1.0;
      lmp.neighbor->skin = 2.0;
    } else fail("Illegal units command");
  }

  void create_atoms(int type) {
    Atom *a = lmp.atom;
    Domain *d = lmp.domain;
    // lattice cells overlapping the box; k (z) outer, j, i inner, basis inner-most (CreateAtoms::add_lattice)
    int lo[3], hi[3];
    for (int q = 0; q < 3; q++) {
      lo[q] = (int)std::floor(d->boxlo[q] / lattice_a) - 1;
      hi[q] = (int)std::ceil(d->boxhi[q] / lattice_a) + 1;
    }
    for (int k = lo[2]; k <= hi[2]; k++)
      for (int j = lo[1]; j <= hi[1]; j++)
        for (int i = lo[0]; i <= hi[0]; i++)
          for (const auto &b : basis) {
            const double x[3] = {(i + b[0]) * lattice_a, (j + b[1]) * lattice_a, (k + b[2]) * lattice_a};
            bool in = true;
            for (int q = 0; q < 3; q++)
              if (x[q] < d->boxlo[q] || x[q] >= d->boxhi[q]) in = false;
            if (!in) continue;
            a->x.insert(a->x.end(), x, x + 3);
            a->type.push_back(type);
          }
    a->nlocal = (int)a->type.size();
    a->natoms = a->nlocal;
    a->v.assign((size_t)3 * a->nlocal, 0.0);
    a->f.assign((size_t)3 * a->nlocal, 0.0);
    if (a->q_flag) a->q.assign(a->nlocal, 0.0);
  }

  void read_data(const std::string &file) {
    std::string path = file;
    std::ifstream in(path);
    if (!in && !dir.empty()) { path = dir + "/" + file; in.open(path); }
    if (!in) fail("Cannot open file " + file);
    Atom *a = lmp.atom;
    std::string line;
    std::getline(in, line);   // title
    long natoms = 0;
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    std::string section;
    std::vector<std::pair<long, std::array<double, 6>>> rows;   // id -> type, q, x, y, z, molecule
    std::map<long, std::array<double, 3>> vel;                  // Velocities section: id -> vx, vy, vz
    while (std::getline(in, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line = line.substr(0, hash);
      std::istringstream ss(line);
      std::vector<std::string> w;
      std::string t;
      while (ss >> t) w.push_back(t);
      if (w.empty()) continue;
      if (w.size() >= 2 && w[1] == "atoms") { natoms = std::atol(w[0].c_str()); continue; }
      if (w.size() >= 3 && w[1] == "atom" && w[2] == "types") {
        a->ntypes = std::atoi(w[0].c_str());
        a->mass.assign(a->ntypes + 1, 0.0);
        a->mass_setflag.assign(a->ntypes + 1, 0);
        continue;
      }
      if (w.size() >= 4 && w[2] == "xlo") { lo[0] = std::atof(w[0].c_str()); hi[0] = std::atof(w[1].c_str()); continue; }
      if (w.size() >= 4 && w[2] == "ylo") { lo[1] = std::atof(w[0].c_str()); hi[1] = std::atof(w[1].c_str()); continue; }
      if (w.size() >= 4 && w[2] == "zlo") { lo[2] = std::atof(w[0].c_str()); hi[2] = std::atof(w[1].c_str()); continue; }
      // sections of atom_style full that the pair / k-space path does not read (examples/data.spce) are skipped
      if (w[0] == "Masses" || w[0] == "Atoms" || w[0] == "Velocities" || w[0] == "Bonds" || w[0] == "Angles" ||
          w[0] == "Dihedrals" || w[0] == "Impropers") { section = w[0]; continue; }
      if (section == "Masses" && w.size() >= 2) {
        const int t2 = std::atoi(w[0].c_str());
        if (t2 < 1 || t2 > a->ntypes) fail("Invalid type for mass set");
        a->mass[t2] = std::atof(w[1].c_str());
        a->mass_setflag[t2] = 1;
      } else if (section == "Bonds" && w.size() >= 4) {   // id type atom1 atom2 (1-based ids)
        a->bonds.push_back(std::atoi(w[2].c_str()) - 1);
        a->bonds.push_back(std::atoi(w[3].c_str()) - 1);
      } else if (section == "Atoms") {
        // atom_style charge: id type q x y z ; atomic: id type x y z ; full: id mol type q x y z [ix iy iz]
        const size_t mol = a->molecule_flag ? 1 : 0;
        const size_t need = (a->q_flag ? 6 : 5) + mol;
        if (w.size() < need) fail("Incorrect atom format in data file");
        std::array<double, 6> r;
        r[5] = mol ? std::atof(w[1].c_str()) : 0.0;
        r[0] = std::atof(w[1 + mol].c_str());
        size_t c = 2 + mol;
        r[1] = a->q_flag ? std::atof(w[c++].c_str()) : 0.0;
        r[2] = std::atof(w[c].c_str()); r[3] = std::atof(w[c + 1].c_str()); r[4] = std::atof(w[c + 2].c_str());
        rows.push_back({std::atol(w[0].c_str()), r});
      } else if (section == "Velocities" && w.size() >= 4) {
        vel[std::atol(w[0].c_str())] = {std::atof(w[1].c_str()), std::atof(w[2].c_str()), std::atof(w[3].c_str())};
      }
    }
    if ((long)rows.size() != natoms) fail("Did not assign all atoms correctly");
    std::sort(rows.begin(), rows.end(), [](const auto &p, const auto &q) { return p.first < q.first; });
    lmp.domain->set_box(lo, hi);
    for (const auto &r : rows) {
      double x[3] = {r.second[2], r.second[3], r.second[4]};
      for (int d = 0; d < 3; d++) {   // Domain::remap into the periodic box
        const double prd = hi[d] - lo[d];
        while (x[d] < lo[d]) x[d] += prd;
        while (x[d] >= hi[d]) x[d] -= prd;
        x[d] = std::max(x[d], lo[d]);
      }
      a->x.insert(a->x.end(), x, x + 3);
      a->type.push_back((int)r.second[0]);
      if (a->q_flag) a->q.push_back(r.second[1]);
      if (a->molecule_flag) a->molecule.push_back((int)r.second[5]);
    }
    a->nlocal = (int)a->type.size();
    a->natoms = a->nlocal;
    a->v.assign((size_t)3 * a->nlocal, 0.0);
    a->f.assign((size_t)3 * a->nlocal, 0.0);
    if (!vel.empty()) {
      if (vel.size() != rows.size()) fail("Did not assign all velocities correctly");
      for (size_t i = 0; i < rows.size(); i++) {
        const auto it = vel.find(rows[i].first);
        if (it == vel.end()) fail("Invalid atom ID in Velocities section of data file");
        for (int d = 0; d < 3; d++) a->v[3 * i + d] = it->second[d];
      }
    }
    // bonds name atoms by id; ids of a data file need not be 1..N in file order, but they are dense here
    for (size_t i = 0; i < rows.size(); i++)
      if (rows[i].first != (long)i + 1) { if (!a->bonds.empty()) fail("Atom IDs of a molecular data file must run 1..N"); break; }
  }

  void replicate(int nx, int ny, int nz) {
    Atom *a = lmp.atom;
    Domain *d = lmp.domain;
    const int n = a->nlocal;
    std::vector<double> x, q, v;
    std::vector<int> type, molecule;
    int maxmol = 0;
    for (int m : a->molecule) maxmol = std::max(maxmol, m);
    // new tags run z outer, y, x inner, old atoms inner-most (atom_offset = (iz*ny*nx + iy*nx + ix)*maxtag)
    for (int iz = 0; iz < nz; iz++)
      for (int iy = 0; iy < ny; iy++)
        for (int ix = 0; ix < nx; ix++)
          for (int i = 0; i < n; i++) {
            x.push_back(a->x[3 * i] + ix * d->prd[0]);
            x.push_back(a->x[3 * i + 1] + iy * d->prd[1]);
            x.push_back(a->x[3 * i + 2] + iz * d->prd[2]);
            type.push_back(a->type[i]);
            if (a->q_flag) q.push_back(a->q[i]);
            for (int d2 = 0; d2 < 3; d2++) v.push_back(a->v[3 * i + d2]);
            // mol_offset = (iz*ny*nx + iy*nx + ix) * maxmol for molecule ids > 0 (Replicate [UPSTREAM])
            if (a->molecule_flag) molecule.push_back(a->molecule[i] > 0 ? a->molecule[i] + ((iz * ny + iy) * nx + ix) * maxmol : 0);
          }
    double hi[3] = {d->boxlo[0] + nx * d->prd[0], d->boxlo[1] + ny * d->prd[1], d->boxlo[2] + nz * d->prd[2]};
    double lo[3] = {d->boxlo[0], d->boxlo[1], d->boxlo[2]};
    d->set_box(lo, hi);
    a->x.swap(x); a->q.swap(q); a->type.swap(type); a->molecule.swap(molecule);
    if (!a->bonds.empty()) {   // every image carries the molecule's bonds (the data file keeps molecules whole)
      std::vector<int> b;
      for (int r = 0; r < nx * ny * nz; r++)
        for (int v : a->bonds) b.push_back(v + r * n);
      a->bonds.swap(b);
    }
    a->nlocal = (int)a->type.size();
    a->natoms = a->nlocal;
    a->v.swap(v);
    a->f.assign((size_t)3 * a->nlocal, 0.0);
  }

  // delete_atoms region ID [mol yes] (examples/in.spce_if, in.hexane_if): atoms inside the block go, with `mol yes`
  // every atom of a molecule that has an atom inside; bonds of deleted atoms go with them.  The survivors keep their
  // order (a molecular system keeps its ids in stock LAMMPS; here the arrays are positional)
  long delete_atoms_region(const Block &b, bool mol) {
    Atom *a = lmp.atom;
    const int n = a->nlocal;
    std::vector<char> dead(n, 0);
    for (int i = 0; i < n; i++) {
      bool in = true;
      for (int d = 0; d < 3; d++)
        if (a->x[3 * (size_t)i + d] < b.lo[d] || a->x[3 * (size_t)i + d] > b.hi[d]) in = false;
      dead[i] = in;
    }
    if (mol) {
      if (!a->molecule_flag) fail("Cannot delete_atoms mol yes for non-molecular systems");
      int maxmol = 0;
      for (int m : a->molecule) maxmol = std::max(maxmol, m);
      std::vector<char> moldead(maxmol + 1, 0);
      for (int i = 0; i < n; i++)
        if (dead[i] && a->molecule[i] > 0) moldead[a->molecule[i]] = 1;
      for (int i = 0; i < n; i++)
        if (a->molecule[i] > 0 && moldead[a->molecule[i]]) dead[i] = 1;
    }
    std::vector<int> newidx(n, -1);
    int m = 0;
    for (int i = 0; i < n; i++) {
      if (dead[i]) continue;
      newidx[i] = m;
      for (int d = 0; d < 3; d++) { a->x[3 * (size_t)m + d] = a->x[3 * (size_t)i + d]; a->v[3 * (size_t)m + d] = a->v[3 * (size_t)i + d]; }
      a->type[m] = a->type[i];
      if (a->q_flag) a->q[m] = a->q[i];
      if (a->molecule_flag) a->molecule[m] = a->molecule[i];
      if (!a->mask.empty()) a->mask[m] = a->mask[i];
      m++;
    }
    a->x.resize((size_t)3 * m); a->v.resize((size_t)3 * m); a->type.resize(m);
    if (a->q_flag) a->q.resize(m);
    if (a->molecule_flag) a->molecule.resize(m);
    if (!a->mask.empty()) a->mask.resize(m);
    std::vector<int> bonds;
    for (size_t k = 0; k + 1 < a->bonds.size(); k += 2) {
      const int i = newidx[a->bonds[k]], j = newidx[a->bonds[k + 1]];
      if (i >= 0 && j >= 0) { bonds.push_back(i); bonds.push_back(j); }
    }
    a->bonds.swap(bonds);
    a->nlocal = m;
    a->natoms = m;
    a->f.assign((size_t)3 * m, 0.0);
    return n - m;
  }

  // Special::build [UPSTREAM]: 1-2 partners are the bonded atoms, 1-3 their partners, 1-4 one hop further; an atom is
  // listed once, in its nearest class, never itself.  Stored as LAMMPS does: cumulative counts + one id list per atom.
  void build_special() {
    Atom *a = lmp.atom;
    a->maxspecial = 0;
    a->nspecial.clear();
    a->special.clear();
    if (a->bonds.empty()) return;
    const int n = a->nlocal;
    std::vector<std::vector<int>> one(n), all(n);
    for (size_t b = 0; b + 1 < a->bonds.size(); b += 2) {
      const int i = a->bonds[b], j = a->bonds[b + 1];
      if (i < 0 || j < 0 || i >= n || j >= n) fail("Bond atoms missing");
      one[i].push_back(j);
      one[j].push_back(i);
    }
    a->nspecial.assign((size_t)3 * n, 0);
    for (int i = 0; i < n; i++) {
      std::vector<int> lvl = one[i], seen = one[i];
      seen.push_back(i);
      all[i] = one[i];
      a->nspecial[3 * i] = (int)all[i].size();
      for (int hop = 1; hop < 3; hop++) {
        std::vector<int> next;
        for (int p : lvl)
          for (int q : one[p])
            if (std::find(seen.begin(), seen.end(), q) == seen.end()) { seen.push_back(q); next.push_back(q); }
        all[i].insert(all[i].end(), next.begin(), next.end());
        a->nspecial[3 * i + hop] = (int)all[i].size();
        lvl.swap(next);
      }
      a->maxspecial = std::max(a->maxspecial, (int)all[i].size());
    }
    if (a->maxspecial > 32) fail("More than 32 special neighbors per atom");
    a->special.assign((size_t)n * std::max(a->maxspecial, 1), 0);
    for (int i = 0; i < n; i++)
      for (size_t s2 = 0; s2 < all[i].size(); s2++) a->special[(size_t)i * a->maxspecial + s2] = all[i][s2];
  }

  double kinetic_energy() const {   // sum 1/2 m v^2 in energy units
    const Atom *a = lmp.atom;
    double ke = 0.0;
    for (int i = 0; i < a->nlocal; i++) {
      const double *v = &a->v[3 * (size_t)i];
      ke += a->mass[a->type[i]] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    }
    return 0.5 * lmp.force->mvv2e * ke;
  }
  double dof() const { return 3.0 * lmp.atom->natoms - 3.0; }   // compute temp: fix_dof = dimension for a periodic box
  double temperature() const { return 2.0 * kinetic_energy() / (dof() * lmp.force->boltz); }

  // velocity all create T seed [loop all|local|geom] [dist uniform] [mom yes] [rot no]
  void velocity_create(double t_desired, int seed, const std::string &loop) {
    Atom *a = lmp.atom;
    if (seed <= 0) fail("Illegal velocity create command");
    for (int t = 1; t <= a->ntypes; t++)
      if (!a->mass_setflag[t]) fail("Cannot use velocity create loop all unless atoms have IDs / masses are set");
    RanPark random(seed);
    for (int i = 0; i < a->nlocal; i++) {
      if (loop == "geom" || loop == "local") random.reset(seed, &a->x[3 * (size_t)i]);
      // dist uniform: u - 1/2 per component, then 1/sqrt(m)
      const double vx = random.uniform() - 0.5, vy = random.uniform() - 0.5, vz = random.uniform() - 0.5;
      const double factor = 1.0 / std::sqrt(a->mass[a->type[i]]);
      a->v[3 * (size_t)i] = vx * factor;
      a->v[3 * (size_t)i + 1] = vy * factor;
      a->v[3 * (size_t)i + 2] = vz * factor;
    }
    // zero linear momentum, then scale to the requested temperature
    double p[3] = {0, 0, 0}, mtot = 0.0;
    for (int i = 0; i < a->nlocal; i++) {
      const double m = a->mass[a->type[i]];
      for (int d = 0; d < 3; d++) p[d] += m * a->v[3 * (size_t)i + d];
      mtot += m;
    }
    for (int i = 0; i < a->nlocal; i++)
      for (int d = 0; d < 3; d++) a->v[3 * (size_t)i + d] -= p[d] / mtot;
    const double t = temperature();
    if (t == 0.0) fail("Attempting to rescale a 0.0 temperature");
    const double fac = std::sqrt(t_desired / t);
    for (double &v : a->v) v *= fac;
  }

  // Force::new_pair / new_kspace, Modify::add_fix [UPSTREAM]: with a suffix enabled (`-sf intel`, `suffix intel`) the
  // style name + "/intel" is looked up first.  Only the /intel classes compute (on the device): the un-suffixed
  // styles exist as their base classes, so asking for one without the suffix is an error here
  template <class Map>
  typename Map::mapped_type lookup(const Map &m, std::string &name, const char *what) {
    if (name.size() > 6 && name.substr(name.size() - 6) == "/intel") { name = name.substr(0, name.size() - 6); suffix_intel = true; }
    const auto it = m.find(name + "/intel");
    if (it == m.end()) fail(std::string("Unknown ") + what + " style " + name);
    if (!suffix_intel)
      fail(std::string(what) + " style " + name + " without -sf intel: only the /intel styles are provided");
    return it->second;
  }

  void pair_style(const std::vector<std::string> &w) {
    std::string s = w[1];
    // `lj/cut/coul/long cut_lj [cut_coul]` (examples/in.spce:7) is `lj/long/coul/long cut long cut_lj [cut_coul]`
    const bool lj_cut = s == "lj/cut/coul/long" || s == "lj/cut/coul/long/intel";
    std::string key = lj_cut ? "lj/long/coul/long" : s;
    const PairCreator make = lookup(styles.pair, key, "pair");
    pair_style_name = lj_cut ? "lj/cut/coul/long" : key;
    pair.reset(make(&lmp));
    lmp.force->pair = pair.get();
    std::vector<char *> args;
    static char a_cut[] = "cut", a_long[] = "long";
    if (lj_cut) { args.push_back(a_cut); args.push_back(a_long); }
    for (size_t i = 2; i < w.size(); i++) args.push_back(const_cast<char *>(w[i].c_str()));
    pair->settings((int)args.size(), args.data());
  }

  void kspace_style(const std::vector<std::string> &w) {
    std::string s = kspace_override.empty() ? w[1] : kspace_override;
    if (s.size() > 6 && s.substr(s.size() - 6) == "/intel") s = s.substr(0, s.size() - 6);
    std::vector<char *> args;
    for (size_t i = 2; i < w.size(); i++) args.push_back(const_cast<char *>(w[i].c_str()));
    if (s == "ewald") {
      // in.buck_coul_long asks for `ewald 1e-6`; the reference has no ewald/intel, so `-sf intel` leaves stock Ewald in
      // place.  This build has no CPU path: the mesh solver is used at the same accuracy, and says so.
      std::fprintf(stderr, "WARNING: kspace_style ewald is not provided on the device; using pppm at the same accuracy\n");
      s = "pppm";
    }
    // only the /intel k-space classes exist here; they are created whatever the suffix setting (the pair style decides)
    const auto it = styles.kspace.find(s + "/intel");
    if (it == styles.kspace.end()) fail("Unknown kspace style " + s);
    kspace_style_name = s;
    kspace.reset(it->second(&lmp, (int)args.size(), args.data()));
    lmp.force->kspace = kspace.get();
  }

  // ---- Verlet ---------------------------------------------------------------------------------------------------
  void init_styles() {
    Atom *a = lmp.atom;
    if (!pair) fail("No pair style defined");   // all scripts of the reference define one
    for (int t = 1; t <= a->ntypes; t++)
      if (!a->mass_setflag[t]) fail("All masses are not set");
    if (!lmp.fix_intel && !dry_run) lmp.fix_intel = new FixIntel(&lmp, device, prec_mode);
    if (lmp.fix_intel) lmp.fix_intel->resident = host_step ? 0 : 1;
  }

  void print_thermo(long step, const double th[16], double ke) {
    const double T = 2.0 * ke / (dof() * lmp.force->boltz);
    const double epair = th[0] + th[1] + th[8];
    const double vol = lmp.domain->prd[0] * lmp.domain->prd[1] * lmp.domain->prd[2];
    const double vsum = th[2] + th[3] + th[4] + th[9] + th[10] + th[11];
    const double press = (dof() * lmp.force->boltz * T + vsum) / 3.0 / vol * nktv2p;
    std::printf("%8ld %14.8g %18.12g %14.8g %18.12g %14.8g\n", step, T, epair, 0.0, epair + ke, press);
    last_thermo[0] = T; last_thermo[1] = epair; last_thermo[2] = epair + ke; last_thermo[3] = press;
    last_thermo[4] = th[0] + th[1]; last_thermo[5] = th[8];
    std::fflush(stdout);
  }

  void forces(int eflag, int vflag, double th[16]) {
    for (int k = 0; k < 16; k++) th[k] = 0.0;
    pair->compute(eflag, vflag);
    th[0] = pair->eng_vdwl; th[1] = pair->eng_coul;
    for (int k = 0; k < 6; k++) th[2 + k] = pair->virial[k];
    if (kspace) {
      kspace->compute(eflag, vflag);
      th[8] = kspace->energy;
      for (int k = 0; k < 6; k++) th[9 + k] = kspace->virial[k];
    }
  }

  double device_ke() {   // kinetic energy from the device-resident velocities
    lmp.fix_intel->sync_host(false, true, false);
    return kinetic_energy();
  }

  // CRC-32 (IEEE, as zlib.crc32) of a table set: every array as raw doubles, then mask, shift, tabinnersq
  static unsigned int crc32(unsigned int crc, const void *data, size_t n) {
    static unsigned int tab[256];
    if (!tab[1])
      for (unsigned int i = 0; i < 256; i++) {
        unsigned int c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        tab[i] = c;
      }
    const unsigned char *p = (const unsigned char *)data;
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = tab[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
  }
  static unsigned int table_crc(const PairTables &t) {
    unsigned int c = 0;
    for (const std::vector<double> *v : {&t.r, &t.dr, &t.f, &t.df, &t.e, &t.de, &t.c, &t.dc})
      c = crc32(c, v->data(), v->size() * sizeof(double));
    const int ms[2] = {t.mask, t.shiftbits};
    c = crc32(c, ms, sizeof(ms));
    return crc32(c, &t.tabinnersq, sizeof(double));
  }

  void dry_summary() {
    Atom *a = lmp.atom;
    std::printf("{\"dry_run\": true, \"natoms\": %ld, \"ntypes\": %d, \"units\": \"%s\", \"box\": [%.10g, %.10g, %.10g], "
                "\"pair_style\": \"%s\", \"cutforce\": %.10g, \"skin\": %g, \"every\": %d, \"delay\": %d, \"check\": %d, "
                "\"dt\": %g, \"temperature\": %.10g",
                a->natoms, a->ntypes, lmp.update->unit_style.c_str(), lmp.domain->prd[0], lmp.domain->prd[1],
                lmp.domain->prd[2], pair_style_name.c_str(), pair->cutforce, lmp.neighbor->skin, lmp.neighbor->every,
                lmp.neighbor->delay, lmp.neighbor->dist_check, lmp.update->dt, temperature());
    if (kspace)
      std::printf(", \"kspace_style\": \"%s\", \"g_ewald\": %.17g, \"grid\": [%d, %d, %d], \"order\": %d, "
                  "\"g_ewald_6\": %.17g, \"grid_6\": [%d, %d, %d]",
                  kspace_style_name.c_str(), kspace->g_ewald, kspace->nx_pppm, kspace->ny_pppm, kspace->nz_pppm,
                  kspace->order, kspace->g_ewald_6, kspace->nx_pppm_6, kspace->ny_pppm_6, kspace->nz_pppm_6);
    if (auto *pp = dynamic_cast<PPPMIntel *>(kspace.get()))
      std::printf(", \"acc\": [%.10g, %.10g, %.10g]", pp->acc_est[0], pp->acc_est[1], pp->acc_est[2]);
    if (auto *pd = dynamic_cast<PPPMDispIntel *>(kspace.get())) {
      // which functions of PPPMDisp::compute the styles select (Coulomb, geometric, arithmetic, no mixing)
      static const char *rule[4] = {"none", "geometric", "arithmetic", "no mixing rule"};
      std::printf(", \"disp_functions\": [%d, %d, %d, %d], \"dispersion_grid\": \"%s\"", pd->function[0], pd->function[1],
                  pd->function[2], pd->function[3], rule[pd->disp_rule()]);
      // estimated absolute RMS force accuracies (PPPMDisp::final_accuracy / final_accuracy_6): Coulomb sum; dispersion
      // sum total, real space, k-space
      std::printf(", \"acc_coul\": [%.10g, %.10g, %.10g], \"acc_6\": [%.10g, %.10g, %.10g], \"order_6\": %d", pd->acc_coul[0],
                  pd->acc_coul[1], pd->acc_coul[2], pd->acc_6[0], pd->acc_6[1], pd->acc_6[2], pd->order_6);
    }
    if (auto *pb = dynamic_cast<PairBuck *>(pair.get())) {
      if (const PairTables *t = pb->coul_tables()) std::printf(", \"coul_table\": [%d, %u]", t->nbits, table_crc(*t));
      if (const PairTables *t = pb->disp_tables()) std::printf(", \"disp_table\": [%d, %u]", t->nbits, table_crc(*t));
    }
    if (!skipped_fixes.empty()) {
      std::printf(", \"skipped_fixes\": [");
      for (size_t i = 0; i < skipped_fixes.size(); i++) std::printf("%s\"%s\"", i ? ", " : "", skipped_fixes[i].c_str());
      std::printf("]");
    }
    std::printf("}\n");
  }

  void run(long nsteps) {
    init_styles();
    if (dry_run) {
      // host-side initialisation only: KSpace sizing first (the pair style reads g_ewald), then init_one
      lmp.dry_run = true;
      if (kspace) kspace->init();
      pair->init();
      dry_summary();
      return;
    }
    FixIntel *fx = lmp.fix_intel;
    build_special();
    fx->upload_atoms();
    fx->setup_neighbor();
    // LAMMPS::init order: force->init() runs kspace->init() before pair->init() (pair reads g_ewald)
    if (kspace) kspace->init();
    pair->init();
    if (kspace) kspace->setup();
    if (!nve) fail("No fix nve defined: nothing to integrate");
    nve->init();
    nve->setup(1);
    // Verlet::setup
    double th[16];
    fx->ensure_neighbor(true);
    {   // the setup build is not counted in "Neighbor list builds" (stock LAMMPS resets the counter at the run start)
      long tot0 = 0, nb0 = 0;
      int ng0 = 0, mx0 = 0;
      b200md_neigh_stats(fx->ctx(), &tot0, &ng0, &mx0, &nb0);
      nbuilds_reported = nb0;
    }
    forces(1, 1, th);
    std::printf("Setting up Verlet run ...\n  Unit style    : %s\n  Current step  : %ld\n  Time step     : %g\n",
                lmp.update->unit_style.c_str(), lmp.update->ntimestep, lmp.update->dt);
    std::printf("%8s %14s %18s %14s %18s %14s\n", "Step", "Temp", "E_pair", "E_mol", "TotEng", "Press");
    print_thermo(lmp.update->ntimestep, th, device_ke());
    const long first = lmp.update->ntimestep;
    for (long s = 1; s <= nsteps; s++) {
      lmp.update->ntimestep++;
      const bool out = (thermo_every && (lmp.update->ntimestep - first) % thermo_every == 0) || s == nsteps;
      nve->initial_integrate(0);
      if (fx->resident) fx->ensure_neighbor(false);
      else fx->sync_host(true, false, false);   // host owns x in the plug-in deployment
      forces(out ? 1 : 0, out ? 1 : 0, th);
      nve->final_integrate();
      if (out) print_thermo(lmp.update->ntimestep, th, device_ke());
    }
    long tot = 0, nb = 0;
    int ng = 0, mx = 0;
    b200md_neigh_stats(fx->ctx(), &tot, &ng, &mx, &nb);
    std::printf("Loop of %ld steps with %ld atoms\nNeighbor list builds = %ld\nTotal # of neighbors = %ld\n"
                "Ave neighs/atom = %.6g\n", nsteps, lmp.atom->natoms, nb - nbuilds_reported, tot,
                lmp.atom->natoms ? (double)tot / lmp.atom->natoms : 0.0);
    nbuilds_reported = nb;
    fx->sync_host(true, true, true);
  }

  void command(const std::string &raw) {
    std::string line = raw;
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    line = substitute(line);
    std::istringstream ss(line);
    std::vector<std::string> w;
    std::string t;
    while (ss >> t) w.push_back(t);
    if (w.empty()) return;
    if (echo) std::printf("%s\n", line.c_str());
    const std::string &c = w[0];
    auto need = [&](size_t n) { if (w.size() < n) fail("Illegal " + c + " command"); };
    Atom *a = lmp.atom;
    if (c == "variable") {
      need(4);
      if (w[2] == "index") { if (!vars.count(w[1])) vars[w[1]] = w[3]; }   // -var on the command line wins
      else if (w[2] == "equal") {
        std::string e;
        for (size_t i = 3; i < w.size(); i++) e += w[i];
        char buf[64];
        std::snprintf(buf, sizeof(buf), "%.15g", eval_expr(e));
        vars[w[1]] = buf;
      } else fail("Illegal variable command");
    } else if (c == "units") { need(2); set_units(w[1]); }
    else if (c == "atom_style") {
      need(2);
      if (w[1] == "atomic") a->q_flag = 0;
      else if (w[1] == "charge") a->q_flag = 1;
      else if (w[1] == "full") { a->q_flag = 1; a->molecule_flag = 1; }   // positions and charges only (examples/in.spce)
      else fail("Unknown atom style " + w[1]);
    } else if (c == "lattice") {
      need(3);
      if (w[1] != "fcc") fail("Illegal lattice command (only fcc is provided)");
      const double scale = std::atof(w[2].c_str());
      // lj units: the value is a reduced density; other units: a lattice constant
      lattice_a = lmp.update->unit_style == "lj" ? std::pow(4.0 / scale, 1.0 / 3.0) : scale;
      basis = {{{0, 0, 0}}, {{0.5, 0.5, 0}}, {{0.5, 0, 0.5}}, {{0, 0.5, 0.5}}};
      std::printf("Lattice spacing in x,y,z = %.6g %.6g %.6g\n", lattice_a, lattice_a, lattice_a);
    } else if (c == "region") {
      need(9);
      if (w[2] != "block") fail("Illegal region command (only block is provided)");
      Block b;
      for (int d = 0; d < 3; d++) {
        b.lo[d] = std::atof(w[3 + 2 * d].c_str()) * lattice_a;
        b.hi[d] = std::atof(w[4 + 2 * d].c_str()) * lattice_a;
      }
      regions[w[1]] = b;
    } else if (c == "create_box") {
      need(3);
      const auto ri = regions.find(w[2]);
      if (ri == regions.end()) fail("Create_box region ID does not exist");
      const Block &b = ri->second;
      a->ntypes = std::atoi(w[1].c_str());
      a->mass.assign(a->ntypes + 1, 0.0);
      a->mass_setflag.assign(a->ntypes + 1, 0);
      lmp.domain->set_box(b.lo, b.hi);
      std::printf("Created orthogonal box = (%g %g %g) to (%.6g %.6g %.6g)\n", b.lo[0], b.lo[1], b.lo[2], b.hi[0], b.hi[1],
                  b.hi[2]);
    } else if (c == "delete_atoms") {
      // delete_atoms region ID [mol yes|no] [compress ...]
      need(3);
      if (w[1] != "region") fail("Illegal delete_atoms command (only `region` is provided)");
      const auto ri = regions.find(w[2]);
      if (ri == regions.end()) fail("Could not find delete_atoms region ID");
      bool mol = false;
      for (size_t i = 3; i + 1 < w.size(); i += 2) {
        if (w[i] == "mol") mol = w[i + 1] == "yes";
        else if (w[i] != "compress") fail("Illegal delete_atoms command");
      }
      const long gone = delete_atoms_region(ri->second, mol);
      std::printf("Deleted %ld atoms, new total = %d\n", gone, a->nlocal);
    } else if (c == "create_atoms") {
      need(3);
      create_atoms(std::atoi(w[1].c_str()));
      std::printf("Created %d atoms\n", a->nlocal);
    } else if (c == "read_data") {
      need(2);
      read_data(w[1]);
      std::printf("  %d atoms\n", a->nlocal);
    } else if (c == "replicate") {
      need(4);
      replicate(std::atoi(w[1].c_str()), std::atoi(w[2].c_str()), std::atoi(w[3].c_str()));
      std::printf("  %d atoms\n", a->nlocal);
    } else if (c == "mass") {
      need(3);
      int lo, hi;
      if (w[1] == "*") { lo = 1; hi = a->ntypes; } else lo = hi = std::atoi(w[1].c_str());
      if (lo < 1 || hi > a->ntypes) fail("Invalid type for mass set");
      for (int t2 = lo; t2 <= hi; t2++) { a->mass[t2] = std::atof(w[2].c_str()); a->mass_setflag[t2] = 1; }
    } else if (c == "velocity") {
      need(5);
      if (w[1] != "all" || w[2] != "create") fail("Illegal velocity command (only `all create` is provided)");
      std::string loop = "all";
      for (size_t i = 5; i + 1 < w.size(); i += 2)
        if (w[i] == "loop") loop = w[i + 1];
      velocity_create(std::atof(w[3].c_str()), std::atoi(w[4].c_str()), loop);
    } else if (c == "special_bonds") {
      // `special_bonds lj/coul a b c` | `lj a b c` | `coul a b c` (examples/in.spce:16)
      for (size_t i = 1; i < w.size();) {
        if ((w[i] == "lj/coul" || w[i] == "lj" || w[i] == "coul") && i + 3 < w.size() + 0) {
          for (int t = 0; t < 3; t++) {
            const double v = std::atof(w[i + 1 + t].c_str());
            if (w[i] != "coul") lmp.force->special_lj[1 + t] = v;
            if (w[i] != "lj") lmp.force->special_coul[1 + t] = v;
          }
          i += 4;
        } else fail("Illegal special_bonds command");
      }
    } else if (c == "bond_style" || c == "angle_style" || c == "dihedral_style" || c == "improper_style" ||
               c == "bond_coeff" || c == "angle_coeff" || c == "dihedral_coeff" || c == "improper_coeff") {
      // bonded terms are not on the pair / k-space path: the topology only feeds the special-bond lists
      if (c == "bond_style" && w.size() > 1 && w[1] != "none")
        std::fprintf(stderr, "WARNING: bonded interactions are not computed by this driver (pair + kspace only)\n");
    } else if (c == "pair_style") { need(2); pair_style(w); }
    else if (c == "pair_coeff") {
      if (!pair) fail("Pair_coeff command before pair_style is defined");
      std::vector<char *> args;
      for (size_t i = 1; i < w.size(); i++) args.push_back(const_cast<char *>(w[i].c_str()));
      pair->coeff((int)args.size(), args.data());
    } else if (c == "pair_modify") {
      if (!pair) fail("Pair_modify command before pair_style is defined");
      for (size_t i = 1; i + 1 < w.size(); i += 2) {
        if (w[i] == "table") pair->ncoultablebits = std::atoi(w[i + 1].c_str());
        else if (w[i] == "table/disp") pair->ndisptablebits = std::atoi(w[i + 1].c_str());
        else if (w[i] == "shift") pair->offset_flag = w[i + 1] == "yes";
        else if (w[i] == "mix") {
          if (w[i + 1] == "geometric") pair->mix_flag = Pair::GEOMETRIC;
          else if (w[i + 1] == "arithmetic") pair->mix_flag = Pair::ARITHMETIC;
          else fail("Illegal pair_modify command (mix geometric|arithmetic are provided)");
        }
        else fail("Illegal pair_modify command");
      }
    } else if (c == "kspace_style") { need(3); kspace_style(w); }
    else if (c == "kspace_modify") {
      if (!kspace) fail("KSpace style has not yet been set");
      std::vector<char *> args;
      for (size_t i = 1; i < w.size(); i++) args.push_back(const_cast<char *>(w[i].c_str()));
      kspace->modify_params((int)args.size(), args.data());
    } else if (c == "neighbor") {
      need(3);
      lmp.neighbor->skin = std::atof(w[1].c_str());
      if (w[2] != "bin") fail("Illegal neighbor command (only bin is provided)");
    } else if (c == "neigh_modify") {
      for (size_t i = 1; i + 1 < w.size(); i += 2) {
        if (w[i] == "every") lmp.neighbor->every = std::atoi(w[i + 1].c_str());
        else if (w[i] == "delay") lmp.neighbor->delay = std::atoi(w[i + 1].c_str());
        else if (w[i] == "check") lmp.neighbor->dist_check = w[i + 1] == "yes";
        else fail("Illegal neigh_modify command");
      }
    } else if (c == "fix") {
      need(4);
      std::string s = w[3];
      if (s.size() > 6 && s.substr(s.size() - 6) == "/intel") s = s.substr(0, s.size() - 6);
      if (s != "nve") {
        // shake, nvt, npt, rigid/small ...: integrators and constraints outside the pair / k-space / nve path.  A dry
        // run still sizes the styles of the script and lists what it skipped; a real run refuses
        if (!dry_run) fail("Unknown fix style " + w[3] + " (only nve is provided)");
        skipped_fixes.push_back(w[3]);
        return;
      }
      const auto gi = groups.find(w[2]);
      if (gi == groups.end()) fail("Could not find fix group ID " + w[2]);
      std::vector<char *> args;
      for (size_t i = 1; i < w.size(); i++) args.push_back(const_cast<char *>(w[i].c_str()));
      nve.reset(static_cast<FixNVEIntel *>(styles.fix.at(s + "/intel")(&lmp, (int)args.size(), args.data())));
      nve->igroup = gi->second;
      nve->groupbit = 1 << gi->second;
    } else if (c == "group") {
      // group ID type T1 T2 ... | group ID id N | N:M ...   (the two forms the /intel examples' variants use)
      need(4);
      if (!lmp.atom->nlocal) fail("Group command before simulation box is defined");
      if (w[1] == "all") fail("Illegal group command");
      if (!groups.count(w[1])) {
        if (groups.size() >= 31) fail("Too many groups");
        const int nb = (int)groups.size();
        groups[w[1]] = nb;
      }
      const int bit = 1 << groups[w[1]];
      if (a->mask.empty()) a->mask.assign(a->nlocal, 1);
      if (w[2] == "type") {
        for (size_t k = 3; k < w.size(); k++) {
          const int gtype = std::atoi(w[k].c_str());
          for (int i = 0; i < a->nlocal; i++)
            if (a->type[i] == gtype) a->mask[i] |= bit;
        }
      } else if (w[2] == "id") {
        for (size_t k = 3; k < w.size(); k++) {
          long lo = std::atol(w[k].c_str()), hi = lo;
          const size_t colon = w[k].find(':');
          if (colon != std::string::npos) hi = std::atol(w[k].substr(colon + 1).c_str());
          for (long id = std::max(lo, 1L); id <= hi && id <= a->nlocal; id++) a->mask[id - 1] |= bit;
        }
      } else fail("Illegal group command (type and id are provided)");
    } else if (c == "package") {
      need(2);
      if (w[1] != "intel") fail("Illegal package command");
      for (size_t i = 3; i + 1 < w.size(); i += 2)
        if (w[i] == "mode") {
          if (w[i + 1] == "double") prec_mode = FixIntel::PREC_MODE_DOUBLE;
          else if (w[i + 1] == "mixed") prec_mode = FixIntel::PREC_MODE_MIXED;
          else if (w[i + 1] == "single") prec_mode = FixIntel::PREC_MODE_SINGLE;
          else fail("Illegal package intel mode");
        }
    } else if (c == "suffix") { need(2); suffix_intel = w[1] == "intel"; }
    else if (c == "thermo") { need(2); thermo_every = std::atoi(w[1].c_str()); }
    else if (c == "timestep") { need(2); lmp.update->dt = std::atof(w[1].c_str()); dt_set = true; }
    else if (c == "run") { need(2); run(std::atol(w[1].c_str())); }
    else if (c == "write_dump") {
      // write_dump all xyz FILE: the host mirror of the positions (valid after a run: FixIntel::sync_host)
      need(4);
      if (w[1] != "all" || w[2] != "xyz") fail("Illegal write_dump command (only `all xyz` is provided)");
      std::FILE *fp = std::fopen(w[3].c_str(), "w");
      if (!fp) fail("Cannot open dump file " + w[3]);
      std::fprintf(fp, "%d\nAtoms. Timestep: %ld\n", a->nlocal, lmp.update->ntimestep);
      for (int i = 0; i < a->nlocal; i++)
        std::fprintf(fp, "%d %.17g %.17g %.17g\n", a->type[i], a->x[3 * i], a->x[3 * i + 1], a->x[3 * i + 2]);
      std::fclose(fp);
    }
    else if (c == "dump" || c == "dump_modify" || c == "undump") {
      if (!warned_dump) std::fprintf(stderr, "WARNING: dump output is not provided by this driver (%s ignored)\n", c.c_str());
      warned_dump = true;
    }
    else if (c == "thermo_style" || c == "thermo_modify" || c == "processors" || c == "newton" || c == "echo" ||
             c == "log" || c == "dimension" || c == "boundary") {
      if (c == "boundary")
        for (size_t i = 1; i < w.size() && i < 4; i++) lmp.domain->periodicity[i - 1] = w[i] == "p";
    } else fail("Unknown command: " + c);
  }
};

}  // namespace

int main(int argc, char **argv) {
  Script s;
  std::string infile;
  try {
    for (int i = 1; i < argc; i++) {
      const std::string a = argv[i];
      if ((a == "-in" || a == "-i") && i + 1 < argc) infile = argv[++i];
      else if ((a == "-sf" || a == "-suffix") && i + 1 < argc) s.suffix_intel = std::string(argv[++i]) == "intel";
      else if ((a == "-var" || a == "-v") && i + 2 < argc) { s.vars[argv[i + 1]] = argv[i + 2]; i += 2; }
      else if (a == "-pk" || a == "-package") {
        // -pk intel Nphi [mode double|mixed]
        std::string line = "package";
        while (i + 1 < argc && argv[i + 1][0] != '-') line += std::string(" ") + argv[++i];
        s.command(line);
      } else if (a == "-styles") {
        // the registered style names, one per line: what the PairStyle / KSpaceStyle / FixStyle lines of the headers add
        for (const auto &kv : s.styles.pair) std::printf("pair %s\n", kv.first.c_str());
        for (const auto &kv : s.styles.kspace) std::printf("kspace %s\n", kv.first.c_str());
        for (const auto &kv : s.styles.fix) std::printf("fix %s\n", kv.first.c_str());
        return 0;
      } else if (a == "-dry-run") s.dry_run = true;
      else if (a == "-host-step") s.host_step = true;
      else if (a == "-echo") s.echo = true;
      else if (a == "-device" && i + 1 < argc) s.device = std::atoi(argv[++i]);
      else if (a == "-kspace" && i + 1 < argc) s.kspace_override = argv[++i];
      else { std::fprintf(stderr, "lmp_b200: unknown option %s\n", a.c_str()); return 2; }
    }
    if (infile.empty()) { std::fprintf(stderr, "usage: lmp_b200 -in script [-sf intel] [-pk intel 0 mode double]\n"); return 2; }
    const size_t slash = infile.rfind('/');
    if (slash != std::string::npos) s.dir = infile.substr(0, slash);
    std::ifstream in(infile);
    if (!in) { std::fprintf(stderr, "ERROR: Cannot open input script %s\n", infile.c_str()); return 1; }
    std::string line;
    while (std::getline(in, line)) s.command(line);
  } catch (const LAMMPSException &e) {
    std::fprintf(stderr, "%s\n", e.what());
    std::printf("%s\n", e.what());
    return 1;
  }
  return 0;
}
