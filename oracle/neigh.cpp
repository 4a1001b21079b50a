// oracle/neigh.cpp — TEST INFRASTRUCTURE (see oracle.h).  Upstream behaviour (the reference ships no neighbour code):
// pinned by the O(N^2) known-answer pair set, not by oracle/_ref.
//
// CPU restatement of the stock-LAMMPS services the reference's pair loops consume but do not ship
// (SURVEY.md Appendix A.3): periodic ghost atoms (Comm::borders, single rank) and the binned half
// neighbour list with newton on (NPairHalfBinNewton, "intel" variant evaluating the criterion in
// flt_t — pair_buck_intel.cpp:370,399-409), plus an O(N^2) full list used as the known answer for the
// pair-set parity test.  Compile with -ffp-contract=off: rsq must round exactly like the AVX (no FMA)
// build of the reference, (dx*dx + dy*dy) + dz*dz.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

template <class flt_t>
inline flt_t rsq_of(const flt_t *xi, const flt_t *xj) {
  const flt_t delx = xi[0] - xj[0];
  const flt_t dely = xi[1] - xj[1];
  const flt_t delz = xi[2] - xj[2];
  return delx * delx + dely * dely + delz * delz;
}

template <class flt_t>
long half_bin(int nlocal, int nall, const double *xd, const int *type, int ntypes,
              const double *cutneighsq_d, const double *boxlo, const double *boxhi,
              double cutneighmax, int *numneigh, long *offsets, int *entries, long cap) {
  const int tp1 = ntypes + 1;
  std::vector<flt_t> x(3 * (size_t)nall);
  for (size_t i = 0; i < 3 * (size_t)nall; i++) x[i] = (flt_t)xd[i];
  std::vector<flt_t> cutsq(tp1 * tp1);
  for (int i = 0; i < tp1 * tp1; i++) cutsq[i] = (flt_t)cutneighsq_d[i];

  // bins tile the periodic box exactly (Neighbor::setup_bins), size ~ cutneighmax/2
  int nbin[3], mlo[3], mbins[3];
  double bininv[3], binsize[3];
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < nall; i++)
    for (int d = 0; d < 3; d++) {
      lo[d] = std::min(lo[d], xd[3 * i + d]);
      hi[d] = std::max(hi[d], xd[3 * i + d]);
    }
  const double binsize_optimal = 0.5 * cutneighmax;
  for (int d = 0; d < 3; d++) {
    const double prd = boxhi[d] - boxlo[d];
    nbin[d] = std::max(1, (int)(prd / binsize_optimal));
    binsize[d] = prd / nbin[d];
    bininv[d] = 1.0 / binsize[d];
    mlo[d] = (int)std::floor((lo[d] - boxlo[d]) * bininv[d]) - 1;
    const int mhi = (int)std::floor((hi[d] - boxlo[d]) * bininv[d]) + 1;
    mbins[d] = mhi - mlo[d] + 1;
  }
  auto coord2bin = [&](const double *p) {
    int b[3];
    for (int d = 0; d < 3; d++) {
      b[d] = (int)std::floor((p[d] - boxlo[d]) * bininv[d]) - mlo[d];
      b[d] = std::min(std::max(b[d], 0), mbins[d] - 1);
    }
    return ((long)b[2] * mbins[1] + b[1]) * mbins[0] + b[0];
  };
  const long nbins_tot = (long)mbins[0] * mbins[1] * mbins[2];
  std::vector<int> atom2bin(nall);
  std::vector<int> binstart(nbins_tot + 1, 0);
  for (int i = 0; i < nall; i++) {
    atom2bin[i] = (int)coord2bin(xd + 3 * i);
    binstart[atom2bin[i] + 1]++;
  }
  for (long b = 0; b < nbins_tot; b++) binstart[b + 1] += binstart[b];
  std::vector<int> binatoms(nall), fill(binstart.begin(), binstart.end() - 1);
  for (int i = 0; i < nall; i++) binatoms[fill[atom2bin[i]]++] = i;  // ascending index in a bin

  // half stencil (Neighbor::stencil_half_bin_3d_newton): upper-half bins within cutneighmax
  int s[3];
  for (int d = 0; d < 3; d++) {
    s[d] = (int)(cutneighmax * bininv[d]);
    if (s[d] * binsize[d] < cutneighmax) s[d]++;
  }
  auto bin_distance = [&](int i, int j, int k) {
    double dx = i > 0 ? (i - 1) * binsize[0] : (i == 0 ? 0.0 : (i + 1) * binsize[0]);
    double dy = j > 0 ? (j - 1) * binsize[1] : (j == 0 ? 0.0 : (j + 1) * binsize[1]);
    double dz = k > 0 ? (k - 1) * binsize[2] : (k == 0 ? 0.0 : (k + 1) * binsize[2]);
    return dx * dx + dy * dy + dz * dz;
  };
  std::vector<long> stencil;
  const double cmaxsq = cutneighmax * cutneighmax;
  for (int k = -s[2]; k <= s[2]; k++)
    for (int j = -s[1]; j <= s[1]; j++)
      for (int i = -s[0]; i <= s[0]; i++)
        if (k > 0 || (k == 0 && j > 0) || (k == 0 && j == 0 && i > 0))
          if (bin_distance(i, j, k) < cmaxsq)
            stencil.push_back(((long)k * mbins[1] + j) * mbins[0] + i);

  // count pass then fill pass (per-atom, so the fill is embarrassingly parallel)
  auto visit = [&](int i, int *out) -> int {
    int n = 0;
    const flt_t *xi = &x[3 * (size_t)i];
    const int itype = type[i];
    const int ibin = atom2bin[i];
    // rest of own bin: owned atoms after i; ghosts by coordinate tie-break
    for (int p = binstart[ibin]; p < binstart[ibin + 1]; p++) {
      const int j = binatoms[p];
      if (j < nlocal) {
        if (j <= i) continue;
      } else {
        const flt_t *xj = &x[3 * (size_t)j];
        if (xj[2] < xi[2]) continue;
        if (xj[2] == xi[2]) {
          if (xj[1] < xi[1]) continue;
          if (xj[1] == xi[1] && xj[0] < xi[0]) continue;
        }
      }
      const flt_t rsq = rsq_of<flt_t>(xi, &x[3 * (size_t)j]);
      if (rsq <= cutsq[itype * tp1 + type[j]]) {
        if (out) out[n] = j;
        n++;
      }
    }
    for (size_t k = 0; k < stencil.size(); k++) {
      const long jb = ibin + stencil[k];
      if (jb < 0 || jb >= nbins_tot) continue;
      for (int p = binstart[jb]; p < binstart[jb + 1]; p++) {
        const int j = binatoms[p];
        const flt_t rsq = rsq_of<flt_t>(xi, &x[3 * (size_t)j]);
        if (rsq <= cutsq[itype * tp1 + type[j]]) {
          if (out) out[n] = j;
          n++;
        }
      }
    }
    return n;
  };
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nlocal; i++) numneigh[i] = visit(i, nullptr);
  offsets[0] = 0;
  for (int i = 0; i < nlocal; i++) offsets[i + 1] = offsets[i] + numneigh[i];
  if (offsets[nlocal] > cap) return -1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nlocal; i++) visit(i, entries + offsets[i]);
  return offsets[nlocal];
}

template <class flt_t>
long full_brute(int nlocal, int nall, const double *xd, const int *type, int ntypes,
                const double *cutneighsq_d, int *numneigh, long *offsets, int *entries, long cap) {
  const int tp1 = ntypes + 1;
  std::vector<flt_t> x(3 * (size_t)nall);
  for (size_t i = 0; i < 3 * (size_t)nall; i++) x[i] = (flt_t)xd[i];
  std::vector<flt_t> cutsq(tp1 * tp1);
  for (int i = 0; i < tp1 * tp1; i++) cutsq[i] = (flt_t)cutneighsq_d[i];
  auto visit = [&](int i, int *out) -> int {
    int n = 0;
    for (int j = 0; j < nall; j++) {
      if (j == i) continue;
      const flt_t rsq = rsq_of<flt_t>(&x[3 * (size_t)i], &x[3 * (size_t)j]);
      if (rsq <= cutsq[type[i] * tp1 + type[j]]) {
        if (out) out[n] = j;
        n++;
      }
    }
    return n;
  };
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < nlocal; i++) numneigh[i] = visit(i, nullptr);
  offsets[0] = 0;
  for (int i = 0; i < nlocal; i++) offsets[i + 1] = offsets[i] + numneigh[i];
  if (offsets[nlocal] > cap) return -1;
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < nlocal; i++) visit(i, entries + offsets[i]);
  return offsets[nlocal];
}

}  // namespace

extern "C" {

int orc_make_ghosts(int nlocal, double *x, int *type, double *q, const double *boxlo,
                    const double *boxhi, const int *periodic, double cutghost, int cap, int *src,
                    int *shift) {
  // Comm::borders on one rank: per dimension, atoms (owned + ghosts so far) inside the slab
  // [lo, lo+cutghost] are imaged to +prd, atoms inside [hi-cutghost, hi] to -prd.
  int nall = nlocal;
  for (int d = 0; d < 3; d++) {
    if (!periodic[d]) continue;
    const double prd = boxhi[d] - boxlo[d];
    const int nstart = nall;  // both sweeps of this dimension scan atoms present before it
    for (int side = 0; side < 2; side++) {
      const double slo = side == 0 ? boxlo[d] : boxhi[d] - cutghost;
      const double shi = side == 0 ? boxlo[d] + cutghost : boxhi[d];
      const double add = side == 0 ? prd : -prd;
      for (int i = 0; i < nstart; i++) {
        const double c = x[3 * (size_t)i + d];
        if (c >= slo && c <= shi) {
          if (nall >= cap) return -1;
          const int g = nall - nlocal;
          for (int k = 0; k < 3; k++) x[3 * (size_t)nall + k] = x[3 * (size_t)i + k];
          x[3 * (size_t)nall + d] = c + add;
          if (type) type[nall] = type[i];
          if (q) q[nall] = q[i];
          src[g] = i;
          for (int k = 0; k < 3; k++) shift[3 * g + k] = i >= nlocal ? shift[3 * (i - nlocal) + k] : 0;
          shift[3 * g + d] += side == 0 ? 1 : -1;
          nall++;
        }
      }
    }
  }
  return nall - nlocal;
}

long orc_neigh_half_bin(int nlocal, int nall, const double *x, const int *type, int ntypes,
                        const double *cutneighsq, const double *boxlo, const double *boxhi,
                        double cutneighmax, int prec, int *numneigh, long *offsets, int *entries,
                        long cap_entries) {
  if (prec == ORC_PREC_DOUBLE)
    return half_bin<double>(nlocal, nall, x, type, ntypes, cutneighsq, boxlo, boxhi, cutneighmax,
                            numneigh, offsets, entries, cap_entries);
  return half_bin<float>(nlocal, nall, x, type, ntypes, cutneighsq, boxlo, boxhi, cutneighmax,
                         numneigh, offsets, entries, cap_entries);
}

long orc_neigh_full_brute(int nlocal, int nall, const double *x, const int *type, int ntypes,
                          const double *cutneighsq, int prec, int *numneigh, long *offsets,
                          int *entries, long cap_entries) {
  if (prec == ORC_PREC_DOUBLE)
    return full_brute<double>(nlocal, nall, x, type, ntypes, cutneighsq, numneigh, offsets, entries,
                              cap_entries);
  return full_brute<float>(nlocal, nall, x, type, ntypes, cutneighsq, numneigh, offsets, entries,
                           cap_entries);
}

void orc_reverse_comm(int nlocal, int nghost, const int *src, double *f) {
  for (int g = nghost - 1; g >= 0; g--) {
    double *fg = f + 4 * (size_t)(nlocal + g);
    double *fs = f + 4 * (size_t)src[g];
    for (int k = 0; k < 4; k++) fs[k] += fg[k];
  }
}

}  // extern "C"
