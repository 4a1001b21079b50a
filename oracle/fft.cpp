// oracle/fft.cpp — TEST INFRASTRUCTURE (see oracle.h).
//
// Stand-in for stock LAMMPS FFT3d with the default KISS back end (FFT_INC empty, Makefile.simd:59):
// unnormalised complex double 3-D FFT as three sweeps of 1-D mixed-radix (2,3,4,5,generic) transforms.
// Called where the reference calls fft1->compute(work1,work1,1) (pppm_intel.cpp:835) and
// fft2->compute(work2,work2,-1) (:903,930,958,1045).  dir=+1 is exp(-i k x).
#include <omp.h>

#include <cmath>
#include <complex>
#include <vector>

#include "oracle.h"

namespace {
typedef std::complex<double> cpx;

struct Plan {
  int n;
  std::vector<int> factors;  // pairs (radix, remaining length)
  std::vector<cpx> tw;       // exp(-2 pi i k / n)
};

Plan make_plan(int n) {
  Plan p;
  p.n = n;
  int m = n;
  while (m > 1) {
    int r;
    if (m % 4 == 0) r = 4;
    else if (m % 2 == 0) r = 2;
    else if (m % 3 == 0) r = 3;
    else if (m % 5 == 0) r = 5;
    else {
      r = 7;
      while (m % r) r += 2;
    }
    m /= r;
    p.factors.push_back(r);
    p.factors.push_back(m);
  }
  p.tw.resize(n);
  for (int k = 0; k < n; k++) {
    const long double ph = -2.0L * 3.14159265358979323846264338327950288L * k / n;
    p.tw[k] = cpx((double)cosl(ph), (double)sinl(ph));
  }
  return p;
}

// decimation-in-time recursion in the manner of kiss_fft's kf_work
void work(cpx *out, const cpx *in, size_t fstride, int in_stride, const int *factors,
          const Plan &pl, int sign) {
  const int p = factors[0], m = factors[1];
  cpx *const out_beg = out;
  if (m == 1) {
    for (int k = 0; k < p; k++) out[k] = in[(size_t)k * fstride * in_stride];
  } else {
    for (int k = 0; k < p; k++)
      work(out + (size_t)k * m, in + (size_t)k * fstride * in_stride, fstride * p, in_stride,
           factors + 2, pl, sign);
  }
  out = out_beg;
  // generic butterfly of radix p over m columns
  cpx scratch[64];
  const int n = pl.n;
  for (int u = 0; u < m; u++) {
    for (int q1 = 0; q1 < p; q1++) scratch[q1] = out[u + q1 * m];
    for (int q1 = 0; q1 < p; q1++) {
      const int k = u + q1 * m;
      size_t twidx = 0;
      cpx acc = scratch[0];
      for (int q = 1; q < p; q++) {
        twidx += fstride * k;
        twidx %= n;
        cpx t = pl.tw[twidx];
        if (sign < 0) t = std::conj(t);
        acc += scratch[q] * t;
      }
      out[k] = acc;
    }
  }
}

void fft1d(const Plan &pl, const cpx *in, int in_stride, cpx *out, int sign) {
  work(out, in, 1, in_stride, pl.factors.data(), pl, sign);
}
}  // namespace

extern "C" void orc_fft3d(double *data, int nx, int ny, int nz, int dir, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  cpx *d = reinterpret_cast<cpx *>(data);
  const Plan px = make_plan(nx), py = make_plan(ny), pz = make_plan(nz);
  const int sign = dir > 0 ? -1 : 1;  // sign > 0 selects exp(-i...) twiddles
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<cpx> buf(std::max(nx, std::max(ny, nz)));
#pragma omp for schedule(static)
    for (long l = 0; l < (long)ny * nz; l++) {
      cpx *line = d + l * nx;
      fft1d(px, line, 1, buf.data(), sign);
      for (int i = 0; i < nx; i++) line[i] = buf[i];
    }
#pragma omp for schedule(static)
    for (long l = 0; l < (long)nx * nz; l++) {
      const long k = l / nx, i = l % nx;
      cpx *line = d + k * (long)nx * ny + i;
      fft1d(py, line, nx, buf.data(), sign);
      for (int j = 0; j < ny; j++) line[(long)j * nx] = buf[j];
    }
#pragma omp for schedule(static)
    for (long l = 0; l < (long)nx * ny; l++) {
      cpx *line = d + l;
      fft1d(pz, line, nx * ny, buf.data(), sign);
      for (int k = 0; k < nz; k++) line[(long)k * nx * ny] = buf[k];
    }
  }
}
