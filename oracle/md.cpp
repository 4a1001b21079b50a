// oracle/md.cpp — TEST INFRASTRUCTURE (see oracle.h).  Whole-timestep CPU loop used for the reported CPU
// baseline (bench.py cpu_baseline / --impl reference) and for trajectory parity tests.
//
// Restates the stock Verlet::run order the reference plugs into:
//   FixNVEIntel::initial_integrate -> Neighbor::decide -> [pbc, borders, build | forward_comm]
//   -> Pair*Intel::compute -> PPPMIntel::compute -> FixNVEIntel::final_integrate
// with the reference's structure: half neighbour list, newton on, thread-private force arrays (pair.cpp),
// thread-private density grids (pppm.cpp).
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "oracle.h"

struct orc_md {
  int nlocal, ntypes, style, prec;
  std::vector<double> x, v, f, q, mass, dtfm, xhold;
  std::vector<int> type;
  double boxlo[3], boxhi[3];
  orc_pair_params p;
  std::vector<double> parr[9];
  double skin, dt, ftm2v;
  int every, delay, check, ago;
  orc_pppm *pppm;
  // ghosts + list
  std::vector<double> xa, qa, fall;
  std::vector<int> ta, src, shift, numneigh, entries;
  std::vector<long> offsets;
  int nghost;
  double cutmax;
  bool built;
  bool lost = false;   // an atom left the box by more than one period (the model blew up): stepping stops
};

namespace {
double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void pbc(orc_md *m) {
  for (int i = 0; i < m->nlocal; i++)
    for (int d = 0; d < 3; d++) {
      double &c = m->x[3 * (size_t)i + d];
      const double prd = m->boxhi[d] - m->boxlo[d];
      if (c < m->boxlo[d]) c += prd;
      if (c >= m->boxhi[d]) {
        c -= prd;
        c = std::max(c, m->boxlo[d]);
      }
    }
}

void build(orc_md *m) {
  pbc(m);
  const int n = m->nlocal;
  for (int i = 0; i < n && !m->lost; i++)
    for (int d = 0; d < 3; d++) {
      const double c = m->x[3 * (size_t)i + d];
      if (!(c >= m->boxlo[d] && c < m->boxhi[d])) m->lost = true;   // also catches NaN
    }
  if (m->lost) return;
  const double cutneighmax = m->cutmax + m->skin;
  const int periodic[3] = {1, 1, 1};
  size_t cap = (size_t)n * 2 + 4096;
  while (true) {
    m->xa.assign(3 * cap, 0.0);
    m->ta.assign(cap, 0);
    m->qa.assign(cap, 0.0);
    m->src.assign(cap, 0);
    m->shift.assign(3 * cap, 0);
    std::copy(m->x.begin(), m->x.end(), m->xa.begin());
    std::copy(m->type.begin(), m->type.end(), m->ta.begin());
    std::copy(m->q.begin(), m->q.end(), m->qa.begin());
    const int ng = orc_make_ghosts(n, m->xa.data(), m->ta.data(), m->qa.data(), m->boxlo, m->boxhi, periodic,
                                   cutneighmax, (int)cap, m->src.data(), m->shift.data());
    if (ng >= 0) { m->nghost = ng; break; }
    cap *= 2;
  }
  const int tp1 = m->ntypes + 1;
  std::vector<double> cns((size_t)tp1 * tp1, 0.0);
  for (int i = 1; i < tp1; i++)
    for (int j = 1; j < tp1; j++) {
      const double c = std::sqrt(m->p.cutsq[i * tp1 + j]) + m->skin;
      cns[i * tp1 + j] = c * c;
    }
  const int nall = n + m->nghost;
  m->numneigh.assign(n, 0);
  m->offsets.assign(n + 1, 0);
  long capent = std::max<long>(1024, (long)m->entries.size());
  while (true) {
    m->entries.resize(capent);
    const long tot = orc_neigh_half_bin(n, nall, m->xa.data(), m->ta.data(), m->ntypes, cns.data(), m->boxlo,
                                        m->boxhi, cutneighmax, m->prec, m->numneigh.data(), m->offsets.data(),
                                        m->entries.data(), capent);
    if (tot >= 0) break;
    capent = m->offsets[n] + 1024;
  }
  m->xhold = m->x;
  m->ago = 0;
  m->built = true;
}

void forward_comm(orc_md *m) {
  // ghosts made dimension by dimension: refresh in creation order so ghosts of ghosts see updated sources
  const int n = m->nlocal;
  std::copy(m->x.begin(), m->x.end(), m->xa.begin());
  for (int g = 0; g < m->nghost; g++) {
    const int s = m->src[g];
    // the ghost differs from its source by one box length in exactly one dimension (the one it was made in)
    for (int d = 0; d < 3; d++) {
      const int ds = m->shift[3 * g + d] - (s >= n ? m->shift[3 * (s - n) + d] : 0);
      m->xa[3 * (size_t)(n + g) + d] = m->xa[3 * (size_t)s + d] + ds * (m->boxhi[d] - m->boxlo[d]);
    }
  }
}

bool check_distance(orc_md *m) {
  const double trig = 0.25 * m->skin * m->skin;
  int flag = 0;
#pragma omp parallel for reduction(| : flag) schedule(static)
  for (int i = 0; i < m->nlocal; i++) {
    const double dx = m->x[3 * (size_t)i] - m->xhold[3 * (size_t)i];
    const double dy = m->x[3 * (size_t)i + 1] - m->xhold[3 * (size_t)i + 1];
    const double dz = m->x[3 * (size_t)i + 2] - m->xhold[3 * (size_t)i + 2];
    if (dx * dx + dy * dy + dz * dz > trig) flag = 1;
  }
  return flag != 0;
}

void forces(orc_md *m, int eflag, int vflag, int nthreads, double *ev, double *ek, double *vk, double *t_pair,
            double *t_kspace) {
  const int n = m->nlocal, nall = n + m->nghost;
  double t0 = now();
  m->fall.resize(4 * (size_t)nall);
  double evl[8];
  orc_pair_eval(m->style, m->prec, eflag, vflag, 0, 1, n, nall, m->xa.data(), m->ta.data(), m->qa.data(),
                m->numneigh.data(), m->offsets.data(), m->entries.data(), &m->p, m->fall.data(), evl, nthreads);
  orc_reverse_comm(n, m->nghost, m->src.data(), m->fall.data());
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) m->f[3 * (size_t)i + d] = m->fall[4 * (size_t)i + d];
  if (ev) for (int k = 0; k < 8; k++) ev[k] = evl[k];
  double t1 = now();
  if (t_pair) *t_pair += t1 - t0;
  if (m->pppm) {
    double e = 0.0, vv[6] = {0, 0, 0, 0, 0, 0};
    orc_pppm_compute(m->pppm, n, m->x.data(), m->q.data(), eflag, vflag, m->f.data(), &e, vv, nthreads);
    if (ek) *ek = e;
    if (vk) for (int k = 0; k < 6; k++) vk[k] = vv[k];
    if (t_kspace) *t_kspace += now() - t1;
  }
}
}  // namespace

extern "C" {

orc_md *orc_md_create(int nlocal, const double *x, const double *v, const double *q, const int *type, int ntypes,
                      const double *mass, const double *boxlo, const double *boxhi, int style, int prec,
                      const orc_pair_params *p, double skin, int every, int delay, int check, double dt,
                      double ftm2v, orc_pppm *pppm) {
  orc_md *m = new orc_md();
  m->nlocal = nlocal; m->ntypes = ntypes; m->style = style; m->prec = prec;
  m->x.assign(x, x + 3 * (size_t)nlocal);
  if (v) m->v.assign(v, v + 3 * (size_t)nlocal);
  else m->v.assign(3 * (size_t)nlocal, 0.0);
  m->f.assign(3 * (size_t)nlocal, 0.0);
  if (q) m->q.assign(q, q + nlocal);
  else m->q.assign(nlocal, 0.0);
  m->type.assign(type, type + nlocal);
  m->mass.assign(mass, mass + ntypes + 1);
  for (int d = 0; d < 3; d++) { m->boxlo[d] = boxlo[d]; m->boxhi[d] = boxhi[d]; }
  m->p = *p;
  const int tp1 = ntypes + 1;
  const double *srcs[9] = {p->cutsq, p->cut_ljsq, p->cut_coulsq, p->buck1, p->buck2, p->rhoinv, p->a, p->c, p->offset};
  double **dsts[9] = {&m->p.cutsq, &m->p.cut_ljsq, &m->p.cut_coulsq, &m->p.buck1, &m->p.buck2, &m->p.rhoinv,
                      &m->p.a, &m->p.c, &m->p.offset};
  m->cutmax = 0.0;
  for (int k = 0; k < 9; k++) {
    m->parr[k].assign(srcs[k], srcs[k] + (size_t)tp1 * tp1);
    *dsts[k] = m->parr[k].data();
  }
  for (int i = 0; i < tp1 * tp1; i++) m->cutmax = std::max(m->cutmax, std::sqrt(m->p.cutsq[i]));
  m->skin = skin; m->every = every; m->delay = delay; m->check = check; m->dt = dt; m->ftm2v = ftm2v;
  m->pppm = pppm;
  m->dtfm.resize(3 * (size_t)nlocal);
  orc_nve_dtfm(nlocal, m->type.data(), m->mass.data(), dt, ftm2v, m->dtfm.data());
  m->built = false;
  m->ago = 0;
  m->nghost = 0;
  return m;
}

void orc_md_destroy(orc_md *m) { delete m; }

void orc_md_run(orc_md *m, int nsteps, int nthreads, double *timers, int *nbuilds) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  double tn = 0, tp = 0, tk = 0, tv = 0, tc = 0;
  int nb = 0;
  if (!m->built && !m->lost) {
    double t0 = now();
    build(m);
    nb++;
    tn += now() - t0;
    if (!m->lost) forces(m, 0, 0, nthreads, nullptr, nullptr, nullptr, &tp, &tk);
  }
  for (int s = 0; s < nsteps && !m->lost; s++) {
    double t0 = now();
    orc_nve_initial(m->nlocal, m->x.data(), m->v.data(), m->f.data(), m->dtfm.data(), m->dt);
    double t1 = now();
    tv += t1 - t0;
    m->ago++;
    bool rebuild = false;
    if (m->ago >= m->delay && m->ago % m->every == 0) rebuild = m->check ? check_distance(m) : true;
    if (rebuild) {
      build(m);
      nb++;
      tn += now() - t1;
      if (m->lost) break;
    } else {
      forward_comm(m);
      tc += now() - t1;
    }
    forces(m, 0, 0, nthreads, nullptr, nullptr, nullptr, &tp, &tk);
    t0 = now();
    orc_nve_final(m->nlocal, m->v.data(), m->f.data(), m->dtfm.data());
    tv += now() - t0;
  }
  if (timers) {
    timers[0] += tn; timers[1] += tp; timers[2] += tk; timers[3] += tv; timers[4] += tc;
  }
  if (nbuilds) *nbuilds += nb;
}

int orc_md_lost(const orc_md *m) { return m->lost ? 1 : 0; }

void orc_md_get(orc_md *m, double *x, double *v, double *f) {
  if (x) std::copy(m->x.begin(), m->x.end(), x);
  if (v) std::copy(m->v.begin(), m->v.end(), v);
  if (f) std::copy(m->f.begin(), m->f.end(), f);
}

void orc_md_energy(orc_md *m, int nthreads, double *ev, double *ekspace, double *ke) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (!m->built) build(m);
  else forward_comm(m);
  double vk[6];
  forces(m, 1, 1, nthreads, ev, ekspace, vk, nullptr, nullptr);
  double k = 0.0;
  for (int i = 0; i < m->nlocal; i++) {
    const double *vi = &m->v[3 * (size_t)i];
    k += 0.5 * m->mass[m->type[i]] * (vi[0] * vi[0] + vi[1] * vi[1] + vi[2] * vi[2]);
  }
  if (ke) *ke = k;
}

}  // extern "C"
