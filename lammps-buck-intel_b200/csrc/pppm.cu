// pppm.cu — PPPM long-range electrostatics on the device (single GPU part; comm.cu adds the slab split).
//
// Replaces PPPMIntel (pppm_intel.cpp):
//   init :67-98, compute :104-317, particle_map<> :326-392, make_rho<> :403-534, brick2fft :642-672,
//   poisson_ik<> :811-977, poisson_ad<> :986-1054, fieldforce_ik<> :541-640, fieldforce_ad<> :679-804
// and the stock PPPM state those read: compute_gf_denom, compute_rho_coeffs,
// compute_gf_ik / compute_gf_ad (+ compute_sf_precoeff), setup (fkx/fky/fkz, vg), qsum_qsq.
//
// Design (B200):
//  * make_rho is a GATHER, not a scatter: atoms are counting-sorted by the grid cell of their lower-left
//    stencil corner (integer atomics only), then one thread per grid point visits the order^3 cells whose
//    atoms reach it, in a fixed order => no FP atomics, bitwise reproducible, no thread-private grids to
//    reduce (pppm_intel.cpp:422-423,521-526 disappear).  The periodic ghost-cell fold (cg->reverse_comm,
//    :185) and brick2fft (:642-672) are fused away: the gather wraps indices and writes the FFT layout.
//  * FFT: hand-written shared-memory Stockham passes (fft.cuh).  The forward z pass, the Green's-function
//    multiply with the energy/virial sums (:846-872), the -i k multiplies (:890-953) and the three inverse
//    z passes are ONE kernel; the last inverse pass stores only the real part (the unpack loops :908-970).
//  * fieldforce gathers with wrapped indices, so the ghost fill (cg->forward_comm, :219-220) is fused away.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "pppm_internal.h"

namespace {

// Stock FFT3d::compute(flag=+1) runs exp(+ikx) and flag=-1 runs exp(-ikx) (flag=1 selects the FFTW_BACKWARD /
// KISS inverse plan); fft.cuh's `s` is +1 for exp(-i...) twiddles.  Pinned by the Ewald known-answer test.
constexpr double S_FWD = -1.0;  // fft1->compute(work1,work1,1), pppm_intel.cpp:835
constexpr double S_BWD = 1.0;   // fft2->compute(work2,work2,-1), :903,930,958

// ---------------------------------------------------------------------------------------------
// setup kernels

__device__ __forceinline__ double d_square(double x) { return x * x; }
__device__ __forceinline__ double d_powsinxx(double x, int n) {
  if (x == 0.0) return 1.0;
  double yy = sin(x) / x, ww = 1.0;
  for (; n != 0; n >>= 1, yy *= yy)
    if (n & 1) ww *= yy;
  return ww;
}
__device__ __forceinline__ double d_gf_denom(const PppmConst &c, double x, double y, double z) {
  double sx = 0, sy = 0, sz = 0;
  for (int l = c.order - 1; l >= 0; l--) {
    sx = c.gf_b[l] + sx * x;
    sy = c.gf_b[l] + sy * y;
    sz = c.gf_b[l] + sz * z;
  }
  const double s = sx * sy * sz;
  return s * s;
}

// PPPM::compute_gf_ik
// (yoff, nyl): the y rows this rank holds in the z-pencil layout [z][y local][x]; (0, ny) on one GPU
__global__ void k_gf_ik(PppmConst c, int nbx, int nby, int nbz, int yoff, int nyl, double *__restrict__ greensfn) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nfft = (long)c.sx * nyl * c.nz;
  if (n >= nfft) return;
  const int k = (int)(n % c.sx), l = (int)((n / c.sx) % nyl) + yoff, m = (int)(n / ((long)c.sx * nyl));
  const double xprd = c.prd[0], yprd = c.prd[1], zprd = c.prd[2];
  const double unitkx = k2PI / xprd, unitky = k2PI / yprd, unitkz = k2PI / zprd;
  const int kper = k - c.nx * (2 * k / c.nx), lper = l - c.ny * (2 * l / c.ny), mper = m - c.nz * (2 * m / c.nz);
  const double snx = d_square(sin(0.5 * unitkx * kper * xprd / c.nx));
  const double sny = d_square(sin(0.5 * unitky * lper * yprd / c.ny));
  const double snz = d_square(sin(0.5 * unitkz * mper * zprd / c.nz));
  const double sqk = d_square(unitkx * kper) + d_square(unitky * lper) + d_square(unitkz * mper);
  double g = 0.0;
  if (sqk != 0.0) {
    const double numerator = 12.5663706 / sqk;
    const double denominator = d_gf_denom(c, snx, sny, snz);
    const int twoorder = 2 * c.order;
    double sum1 = 0.0;
    for (int ax = -nbx; ax <= nbx; ax++) {
      const double qx = unitkx * (kper + c.nx * ax);
      const double sx = exp(-0.25 * d_square(qx / c.g_ewald));
      const double wx = d_powsinxx(0.5 * qx * xprd / c.nx, twoorder);
      for (int ay = -nby; ay <= nby; ay++) {
        const double qy = unitky * (lper + c.ny * ay);
        const double sy = exp(-0.25 * d_square(qy / c.g_ewald));
        const double wy = d_powsinxx(0.5 * qy * yprd / c.ny, twoorder);
        for (int az = -nbz; az <= nbz; az++) {
          const double qz = unitkz * (mper + c.nz * az);
          const double sz = exp(-0.25 * d_square(qz / c.g_ewald));
          const double wz = d_powsinxx(0.5 * qz * zprd / c.nz, twoorder);
          const double dot1 = unitkx * kper * qx + unitky * lper * qy + unitkz * mper * qz;
          const double dot2 = qx * qx + qy * qy + qz * qz;
          sum1 += (dot1 / dot2) * sx * sy * sz * wx * wy * wz;
        }
      }
    }
    g = numerator * sum1 / denominator;
  }
  greensfn[n] = g;
}

// PPPM::compute_gf_ik_triclinic [UPSTREAM] (what PPPM::setup_triclinic calls for the box of pppm_intel.cpp:878-883): the
// wave vectors and their aliases go through Domain::x2lamdaT, the assignment-function factors stay per lamda axis
__device__ __forceinline__ void d_x2lamdaT(const PppmConst &c, double a, double b, double cc, double &ox, double &oy,
                                           double &oz) {
  ox = c.hinv[0] * a;
  oy = c.hinv[5] * a + c.hinv[1] * b;
  oz = c.hinv[4] * a + c.hinv[3] * b + c.hinv[2] * cc;
}
__global__ void k_gf_ik_tri(PppmConst c, int nbx, int nby, int nbz, double *__restrict__ greensfn) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nfft = (long)c.sx * c.ny * c.nz;
  if (n >= nfft) return;
  const int k = (int)(n % c.sx), l = (int)((n / c.sx) % c.ny), m = (int)(n / ((long)c.sx * c.ny));
  const int kper = k - c.nx * (2 * k / c.nx), lper = l - c.ny * (2 * l / c.ny), mper = m - c.nz * (2 * m / c.nz);
  const double snx = d_square(sin(kPI * kper / c.nx)), sny = d_square(sin(kPI * lper / c.ny)),
               snz = d_square(sin(kPI * mper / c.nz));
  double ux, uy, uz;
  d_x2lamdaT(c, k2PI * kper, k2PI * lper, k2PI * mper, ux, uy, uz);
  const double sqk = ux * ux + uy * uy + uz * uz;
  double g = 0.0;
  if (sqk != 0.0) {
    const double numerator = 12.5663706 / sqk;
    const double denominator = d_gf_denom(c, snx, sny, snz);
    const int twoorder = 2 * c.order;
    double sum1 = 0.0;
    for (int ax = -nbx; ax <= nbx; ax++) {
      const double wx = d_powsinxx(kPI * kper / c.nx + kPI * ax, twoorder);
      for (int ay = -nby; ay <= nby; ay++) {
        const double wy = d_powsinxx(kPI * lper / c.ny + kPI * ay, twoorder);
        for (int az = -nbz; az <= nbz; az++) {
          const double wz = d_powsinxx(kPI * mper / c.nz + kPI * az, twoorder);
          double bx, by, bz;
          d_x2lamdaT(c, k2PI * c.nx * ax, k2PI * c.ny * ay, k2PI * c.nz * az, bx, by, bz);
          const double qx = ux + bx, qy = uy + by, qz = uz + bz;
          const double sx = exp(-0.25 * d_square(qx / c.g_ewald)), sy = exp(-0.25 * d_square(qy / c.g_ewald)),
                       sz = exp(-0.25 * d_square(qz / c.g_ewald));
          const double dot1 = ux * qx + uy * qy + uz * qz;
          const double dot2 = qx * qx + qy * qy + qz * qz;
          sum1 += (dot1 / dot2) * sx * sy * sz * wx * wy * wz;
        }
      }
    }
    g = numerator * sum1 / denominator;
  }
  greensfn[n] = g;
}

// PPPM::compute_sf_precoeff + the per-point part of compute_gf_ad
// (yoff, nyl: the y rows held by this rank, [z][row][x]; the whole grid on one GPU)
// have_g: greensfn already holds the influence function (dispersion grid: k_gf_6) and only the sums are formed
// the six alias sums of PPPM::compute_sf_precoeff at the grid point with the periodic indices (kper, lper, mper)
__device__ __forceinline__ void sf_sums(const PppmConst &c, int kper, int lper, int mper, double sum[6]) {
  double wx0[5], wy0[5], wz0[5], wx1[5], wy1[5], wz1[5], wx2[5], wy2[5], wz2[5];
  for (int i = 0; i < 5; i++) {
    wx0[i] = d_powsinxx(0.5 * (k2PI * (kper + c.nx * (i - 2))) / c.nx, c.order);
    wx1[i] = d_powsinxx(0.5 * (k2PI * (kper + c.nx * (i - 1))) / c.nx, c.order);
    wx2[i] = d_powsinxx(0.5 * (k2PI * (kper + c.nx * (i))) / c.nx, c.order);
    wy0[i] = d_powsinxx(0.5 * (k2PI * (lper + c.ny * (i - 2))) / c.ny, c.order);
    wy1[i] = d_powsinxx(0.5 * (k2PI * (lper + c.ny * (i - 1))) / c.ny, c.order);
    wy2[i] = d_powsinxx(0.5 * (k2PI * (lper + c.ny * (i))) / c.ny, c.order);
    wz0[i] = d_powsinxx(0.5 * (k2PI * (mper + c.nz * (i - 2))) / c.nz, c.order);
    wz1[i] = d_powsinxx(0.5 * (k2PI * (mper + c.nz * (i - 1))) / c.nz, c.order);
    wz2[i] = d_powsinxx(0.5 * (k2PI * (mper + c.nz * (i))) / c.nz, c.order);
  }
  for (int t = 0; t < 6; t++) sum[t] = 0.0;
  for (int ax = 0; ax < 5; ax++)
    for (int ay = 0; ay < 5; ay++)
      for (int az = 0; az < 5; az++) {
        const double u0 = wx0[ax] * wy0[ay] * wz0[az];
        sum[0] += u0 * (wx1[ax] * wy0[ay] * wz0[az]);
        sum[1] += u0 * (wx2[ax] * wy0[ay] * wz0[az]);
        sum[2] += u0 * (wx0[ax] * wy1[ay] * wz0[az]);
        sum[3] += u0 * (wx0[ax] * wy2[ay] * wz0[az]);
        sum[4] += u0 * (wx0[ax] * wy0[ay] * wz1[az]);
        sum[5] += u0 * (wx0[ax] * wy0[ay] * wz2[az]);
      }
}

__global__ void k_gf_ad(PppmConst c, int yoff, int nyl, double *__restrict__ greensfn, double *__restrict__ sfpre,
                        int have_g) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nfft = (long)c.sx * nyl * c.nz;
  if (n >= nfft) return;
  const int k = (int)(n % c.sx), l = (int)((n / c.sx) % nyl) + yoff, m = (int)(n / ((long)c.sx * nyl));
  const double xprd = c.prd[0], yprd = c.prd[1], zprd = c.prd[2];
  const double unitkx = k2PI / xprd, unitky = k2PI / yprd, unitkz = k2PI / zprd;
  const int kper = k - c.nx * (2 * k / c.nx), lper = l - c.ny * (2 * l / c.ny), mper = m - c.nz * (2 * m / c.nz);
  const int twoorder = 2 * c.order;
  const double qx = unitkx * kper, qy = unitky * lper, qz = unitkz * mper;
  const double snx = d_square(sin(0.5 * qx * xprd / c.nx)), sny = d_square(sin(0.5 * qy * yprd / c.ny)),
               snz = d_square(sin(0.5 * qz * zprd / c.nz));
  const double sx = exp(-0.25 * d_square(qx / c.g_ewald)), sy = exp(-0.25 * d_square(qy / c.g_ewald)),
               sz = exp(-0.25 * d_square(qz / c.g_ewald));
  const double wx = d_powsinxx(0.5 * qx * xprd / c.nx, twoorder), wy = d_powsinxx(0.5 * qy * yprd / c.ny, twoorder),
               wz = d_powsinxx(0.5 * qz * zprd / c.nz, twoorder);
  const double sqk = qx * qx + qy * qy + qz * qz;
  double g = 0.0;
  if (have_g) g = greensfn[n];
  else {
    if (sqk != 0.0) g = (k4PI / sqk) * sx * sy * sz * wx * wy * wz / d_gf_denom(c, snx, sny, snz);
    greensfn[n] = g;
  }
  double sum[6];
  sf_sums(c, kper, lper, mper, sum);
  // half spectrum: a stored point with 0 < kx < nx / 2 also stands for its mirror image (-kx, -ky, -kz) in the sums
  // over the grid; G is even, the truncated alias sums are not exactly, so the mirror image is evaluated itself
  if (c.sx != c.nx && k != 0 && 2 * k != c.nx) {
    const int k2 = c.nx - k, l2 = (c.ny - l) % c.ny, m2 = (c.nz - m) % c.nz;
    double sum2[6];
    sf_sums(c, k2 - c.nx * (2 * k2 / c.nx), l2 - c.ny * (2 * l2 / c.ny), m2 - c.nz * (2 * m2 / c.nz), sum2);
    for (int t = 0; t < 6; t++) sum[t] += sum2[t];
  }
  for (int t = 0; t < 6; t++) sfpre[(size_t)t * nfft + n] = sum[t] * g;
}

// PPPMDisp::compute_gf_6 [UPSTREAM]: influence function of the r^-6 reciprocal sum (c.g_ewald = g_ewald_6).  Used by
// the geometric-mixing grid of PPPMDispIntel::compute (pppm_disp_intel.cpp:245-313).
__global__ void k_gf_6(PppmConst c, int yoff, int nyl, double *__restrict__ greensfn) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nfft = (long)c.sx * nyl * c.nz;
  if (n >= nfft) return;
  const int k = (int)(n % c.sx), l = (int)((n / c.sx) % nyl) + yoff, m = (int)(n / ((long)c.sx * nyl));
  const double xprd = c.prd[0], yprd = c.prd[1], zprd = c.prd[2];
  const double unitkx = k2PI / xprd, unitky = k2PI / yprd, unitkz = k2PI / zprd;
  const int kper = k - c.nx * (2 * k / c.nx), lper = l - c.ny * (2 * l / c.ny), mper = m - c.nz * (2 * m / c.nz);
  const double inv2ew = 1.0 / (2.0 * c.g_ewald);
  const double rtpi = sqrt(kPI);
  const double numerator = -kPI * rtpi * c.g_ewald * c.g_ewald * c.g_ewald / 3.0;
  const double qx = unitkx * kper, qy = unitky * lper, qz = unitkz * mper;
  const double snx2 = d_square(sin(0.5 * qx * xprd / c.nx)), sny2 = d_square(sin(0.5 * qy * yprd / c.ny)),
               snz2 = d_square(sin(0.5 * qz * zprd / c.nz));
  const double sx = exp(-qx * qx * inv2ew * inv2ew), sy = exp(-qy * qy * inv2ew * inv2ew),
               sz = exp(-qz * qz * inv2ew * inv2ew);
  const double argx = 0.5 * qx * xprd / c.nx, argy = 0.5 * qy * yprd / c.ny, argz = 0.5 * qz * zprd / c.nz;
  double wx = argx != 0.0 ? pow(sin(argx) / argx, (double)c.order) : 1.0;
  double wy = argy != 0.0 ? pow(sin(argy) / argy, (double)c.order) : 1.0;
  double wz = argz != 0.0 ? pow(sin(argz) / argz, (double)c.order) : 1.0;
  wx *= wx; wy *= wy; wz *= wz;
  const double sqk = qx * qx + qy * qy + qz * qz;
  double g = 0.0;
  if (sqk != 0.0) {
    const double rtsqk = sqrt(sqk);
    const double term = (1.0 - 2.0 * sqk * inv2ew * inv2ew) * sx * sy * sz +
                        2.0 * sqk * rtsqk * inv2ew * inv2ew * inv2ew * rtpi * erfc(rtsqk * inv2ew);
    g = numerator * term * wx * wy * wz / d_gf_denom(c, snx2, sny2, snz2);
  }
  greensfn[n] = g;
}

// fixed-order sum of `ncol` interleaved-by-column arrays: in[col*n + i] -> out[col]
__global__ void __launch_bounds__(256) k_colsum_partial(long n, int ncol, const double *__restrict__ in, double *__restrict__ partial) {
  __shared__ double s[256];
  for (int col = 0; col < ncol; col++) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    s[threadIdx.x] = i < n ? in[(size_t)col * n + i] : 0.0;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
      if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)col * gridDim.x + blockIdx.x] = s[0];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_colsum_final(int nrows, int ncol, const double *__restrict__ partial, double *__restrict__ out) {
  __shared__ double s[256];
  for (int col = 0; col < ncol; col++) {
    double a = 0.0;
    for (int r = threadIdx.x; r < nrows; r += 256) a += partial[(size_t)col * nrows + r];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
      if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[col] = s[0];
    __syncthreads();
  }
}

// qsum_qsq: columns {q, q^2}
__global__ void k_q_moments(int n, const double4 *__restrict__ xq, const int *__restrict__ type,
                            const double *__restrict__ Btype, double *__restrict__ cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double q = Btype ? Btype[type[i]] : xq[i].w;
  cols[i] = q;
  cols[(size_t)n + i] = q * q;
}

// PPPM::slabcorr [UPSTREAM], called at pppm_intel.cpp:305: dipole moments {q z, q z^2} ...
__global__ void k_slab_moments(int n, const double4 *__restrict__ xq, double *__restrict__ cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xq[i];
  cols[i] = p.w * p.z;
  cols[(size_t)n + i] = p.w * p.z * p.z;
}
// ... and the corrections: f_z += ffact q (dipole - qsum z); eatom += efact q (z dipole - (dipole_r2 + qsum z^2) / 2
// - qsum zprd^2 / 12)
__global__ void k_slabcorr(int n, const double4 *__restrict__ xq, double4 *__restrict__ f, double ffact, double dipole,
                           double qsum, double *__restrict__ eatom, double efact, double dipole_r2, double zprd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = xq[i];
  f[i].z += ffact * p.w * (dipole - qsum * p.z);
  if (eatom) eatom[i] += efact * p.w * (p.z * dipole - 0.5 * (dipole_r2 + qsum * p.z * p.z) - qsum * zprd * zprd / 12.0);
}

// ---------------------------------------------------------------------------------------------
// per-step: particle_map + sort by cell

// particle_map<flt_t,acc_t>, pppm_intel.cpp:344-372: index arithmetic in flt_t, un-fused (mul then add) like the
// reference's AVX build so the integer cell of an atom is the same on both sides
__device__ __forceinline__ int map1(double x, double lo, double xi, double fshift) {
  return static_cast<int>(__dadd_rn(__dmul_rn(__dsub_rn(x, lo), xi), fshift)) - PPPM_OFFSET;
}
__device__ __forceinline__ int map1(float x, float lo, float xi, float fshift) {
  return static_cast<int>(__fadd_rn(__fmul_rn(__fsub_rn(x, lo), xi), fshift)) - PPPM_OFFSET;
}
template <class flt_t>
__device__ __forceinline__ void map_atom(const PppmConst &c, flt_t x, flt_t y, flt_t z, int &nx, int &ny, int &nz) {
  const flt_t fshift = (flt_t)c.shift;
  nx = map1(x, (flt_t)c.boxlo[0], (flt_t)c.delinv[0], fshift);
  ny = map1(y, (flt_t)c.boxlo[1], (flt_t)c.delinv[1], fshift);
  nz = map1(z, (flt_t)c.boxlo[2], (flt_t)c.delinv[2], fshift) - c.zoff;   // local plane index (zoff = 0 on one GPU)
}
// dx = nx + fshiftone - (x - lo)*xi (pppm_intel.cpp:469-471), operands in flt_t
__device__ __forceinline__ double frac1(int n, double so, double x, double lo, double xi) {
  return __dsub_rn(__dadd_rn((double)n, so), __dmul_rn(__dsub_rn(x, lo), xi));
}
__device__ __forceinline__ double frac1(int n, float so, float x, float lo, float xi) {
  return (double)__fsub_rn(__fadd_rn((float)n, so), __fmul_rn(__fsub_rn(x, lo), xi));
}

// Domain::x2lamda (pppm_intel.cpp:156): box -> lamda coordinates of a triclinic cell, in double; q rides along
__device__ __forceinline__ double4 to_lamda(const PppmConst &c, const double4 p) {
  const double dx = p.x - c.boxlo_box[0], dy = p.y - c.boxlo_box[1], dz = p.z - c.boxlo_box[2];
  return make_double4(c.hinv[0] * dx + c.hinv[5] * dy + c.hinv[4] * dz, c.hinv[1] * dy + c.hinv[3] * dz, c.hinv[2] * dz,
                      p.w);
}

template <class flt_t>
__global__ void k_map_key(int n, const double4 *__restrict__ xq, const float4 *__restrict__ xqf, PppmConst c,
                          int *__restrict__ key, int *__restrict__ cell_count, int *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int nx, ny, nz;
  if (c.tri) {   // lamda coordinates, rounded to flt_t like the packed positions of the orthogonal path
    const double4 p = to_lamda(c, xq[i]);
    if (sizeof(flt_t) == 4) map_atom<float>(c, (float)p.x, (float)p.y, (float)p.z, nx, ny, nz);
    else map_atom<double>(c, p.x, p.y, p.z, nx, ny, nz);
  } else if (sizeof(flt_t) == 4) {
    const float4 p = xqf[i];
    map_atom<float>(c, p.x, p.y, p.z, nx, ny, nz);
  } else {
    const double4 p = xq[i];
    map_atom<double>(c, p.x, p.y, p.z, nx, ny, nz);
  }
  if (nx + c.nlower < c.lo_out[0] || nx + c.nupper > c.hi_out[0] || ny + c.nlower < c.lo_out[1] ||
      ny + c.nupper > c.hi_out[1] || nz + c.nlower < c.lo_out[2] || nz + c.nupper > c.hi_out[2])
    flags[0] = 1;  // "Out of range atoms - cannot compute PPPM" (pppm_intel.cpp:379-385)
  const int k = (wrapi(nz, c.nz) * c.ny + wrapi(ny, c.ny)) * c.nx + wrapi(nx, c.nx);
  key[i] = k;
  atomicAdd(&cell_count[k], 1);
}

__global__ void k_cell_scatter(int n, const int *__restrict__ key, int *__restrict__ cursor, int *__restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  perm[atomicAdd(&cursor[key[i]], 1)] = i;
}

__global__ void k_cell_order(long ncell, const int *__restrict__ start, int *__restrict__ perm) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ncell) return;
  const int s = start[b], e = start[b + 1];
  for (int a = s + 1; a < e; a++) {
    const int kk = perm[a];
    int k = a - 1;
    while (k >= s && perm[k] > kk) {
      perm[k + 1] = perm[k];
      k--;
    }
    perm[k + 1] = kk;
  }
}

// per sorted atom: dx,dy,dz + weight (pa_x), lower-left cell + atom index (pa_n), and for the gather the
// 3*order one-dimensional stencil weights (pa_w, Horner over rho_coeff exactly as pppm_intel.cpp:476-488; rounded
// to flt_t like the reference's `flt_t rho[3][INTEL_P3M_MAXORDER]`, :474) and the wrapped x cell (pa_cx)
template <class flt_t>
__global__ void k_fill_sorted(int n, const int *__restrict__ perm, const double4 *__restrict__ xq,
                              const float4 *__restrict__ xqf, const int *__restrict__ type,
                              const double *__restrict__ Btype, PppmConst c, double4 *__restrict__ pa_x,
                              int4 *__restrict__ pa_n, double *__restrict__ pa_w, int *__restrict__ pa_cx) {
  __shared__ double s_rc[B2_MAXORDER * B2_MAXORDER];
  for (int k = threadIdx.x; k < c.order * c.order; k += blockDim.x) s_rc[k] = c.rho_coeff[k];
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = perm[k];
  int nx, ny, nz;
  double dx, dy, dz, w;
  if (sizeof(flt_t) == 4) {
    float4 p;
    if (c.tri) {
      const double4 pl = to_lamda(c, xq[i]);
      p = make_float4((float)pl.x, (float)pl.y, (float)pl.z, (float)pl.w);
    } else p = xqf[i];
    map_atom<float>(c, p.x, p.y, p.z, nx, ny, nz);
    const float so = (float)c.shiftone;
    dx = frac1(nx, so, p.x, (float)c.boxlo[0], (float)c.delinv[0]);
    dy = frac1(ny, so, p.y, (float)c.boxlo[1], (float)c.delinv[1]);
    dz = frac1(nz + c.zoff, so, p.z, (float)c.boxlo[2], (float)c.delinv[2]);
    const float qw = Btype ? (float)Btype[type[i]] : p.w;
    w = (double)__fmul_rn((float)c.delvolinv, qw);
  } else {
    const double4 p = c.tri ? to_lamda(c, xq[i]) : xq[i];
    map_atom<double>(c, p.x, p.y, p.z, nx, ny, nz);
    dx = frac1(nx, c.shiftone, p.x, c.boxlo[0], c.delinv[0]);
    dy = frac1(ny, c.shiftone, p.y, c.boxlo[1], c.delinv[1]);
    dz = frac1(nz + c.zoff, c.shiftone, p.z, c.boxlo[2], c.delinv[2]);
    w = c.delvolinv * (Btype ? Btype[type[i]] : p.w);
  }
  pa_x[k] = make_double4(dx, dy, dz, w);
  pa_n[k] = make_int4(nx, ny, nz, i);
  pa_cx[k] = wrapi(nx, c.nx);
  if (!pa_w) return;   // the tiled make_rho evaluates the weights itself
  const int order = c.order;
  double *wk = pa_w + (size_t)k * (3 * order);
  for (int t = 0; t < order; t++) {
    double r1 = 0.0, r2 = 0.0, r3 = 0.0;
    for (int l = order - 1; l >= 0; l--) {
      r1 = s_rc[l * order + t] + r1 * dx;
      r2 = s_rc[l * order + t] + r2 * dy;
      r3 = s_rc[l * order + t] + r3 * dz;
    }
    if (sizeof(flt_t) == 4) { r1 = (double)(float)r1; r2 = (double)(float)r2; r3 = (double)(float)r3; }
    wk[t] = r1;
    wk[order + t] = r2;
    wk[2 * order + t] = r3;
  }
}

// ---------------------------------------------------------------------------------------------
// make_rho as a gather: one thread per grid point

// One thread per grid point, blocks are 8x8x4 tiles of points so that the threads of a block share atoms (L1).
// The `order` cells of a (cy,cz) row that can reach the point are consecutive keys, i.e. ONE contiguous range of
// the sorted atoms (two ranges when the row wraps around the periodic box): 2 loads per row instead of 2 per cell.
__global__ void __launch_bounds__(256)
k_make_rho(PppmConst c, const int *__restrict__ cell_start, const double4 *__restrict__ pa_x,
           const double *__restrict__ pa_w, const int *__restrict__ pa_cx, double *__restrict__ density) {
  const int tx = threadIdx.x & 7, ty = (threadIdx.x >> 3) & 7, tz = threadIdx.x >> 6;
  const int gx = blockIdx.x * 8 + tx, gy = blockIdx.y * 8 + ty, gz = blockIdx.z * 4 + tz;
  if (gx >= c.nx || gy >= c.ny || gz >= c.nz) return;
  const int order = c.order, w3 = 3 * order;
  // x cells [gx-nupper, gx-nlower]; l - nlower = gx - cx - nlower (mod nx)
  const int xlo = gx - c.nupper, xhi = gx - c.nlower;
  double rho = 0.0;
  for (int n = c.nlower; n <= c.nupper; n++) {
    const int cz = wrapi(gz - n, c.nz);
    for (int m = c.nlower; m <= c.nupper; m++) {
      const int cy = wrapi(gy - m, c.ny);
      const long row = ((long)cz * c.ny + cy) * c.nx;
      // up to two pieces: [max(xlo,0), min(xhi,nx-1)] and the wrapped remainder
      for (int piece = 0; piece < 2; piece++) {
        int a0, a1;
        if (piece == 0) { a0 = max(xlo, 0); a1 = min(xhi, c.nx - 1); }
        else if (xlo < 0) { a0 = xlo + c.nx; a1 = c.nx - 1; }
        else if (xhi >= c.nx) { a0 = 0; a1 = xhi - c.nx; }
        else break;
        const int s = cell_start[row + a0], e = cell_start[row + a1 + 1];
        for (int a = s; a < e; a++) {
          int l = gx - pa_cx[a];              // stencil offset along x, brought back into [nlower,nupper]
          if (l > c.nupper) l -= c.nx;
          else if (l < c.nlower) l += c.nx;
          const double *wk = pa_w + (size_t)a * w3;
          const double z0 = pa_x[a].w;
          rho += ((z0 * wk[2 * order + (n - c.nlower)]) * wk[order + (m - c.nlower)]) * wk[l - c.nlower];
        }
      }
    }
  }
  density[((long)gz * c.ny + gy) * c.nx + gx] = rho;
}

// ---------------------------------------------------------------------------------------------
// make_rho, tiled: the production path.  The grid is cut into 8x8x8-cell tiles.  ONE WARP owns one tile: it walks
// the tile's atoms in sorted order (per (y,z) cell row one contiguous range of the cell-sorted arrays) and adds
// each atom's order^3 stencil into a warp-private (8+order-1)^3 block of SHARED memory — the 32 lanes cover the
// stencil points in ceil(order^3/32) rounds, distinct lanes hit distinct addresses, atoms are strictly sequential,
// so there are no atomics and the summation order is fixed.  The stencil block (tile + halo) is then stored and a
// second kernel gives every grid point the sum of the <= 8 tile blocks that cover it, in a fixed order (periodic
// wrap included: this is the ghost-cell fold of cg->reverse_comm, pppm_intel.cpp:185, and brick2fft, :642-672).
#define RHO_T 8

struct TileGeom {
  int ntx, nty, ntz, E;   // tiles per dimension, E = RHO_T + order - 1
  signed char lane_pt[64];  // lane + 32 r -> (m,l) point m * order + l of the stencil's xy face, -1 = idle
};

// x pitch of the shared-memory stencil block: the smallest >= E for which rho_lane_map finds conflict-free rounds
__host__ __device__ constexpr int rho_pitch(int order) { return order == 5 ? 13 : RHO_T + order - 1; }

// x pitch and lane -> (m,l) assignment: a round touches the addresses (m * pitch + l) of its active lanes; 64-bit
// shared-memory accesses are served per half warp, 16 doubles per wavefront, so the 16 lanes of a half warp get points
// with distinct (m * pitch + l) mod 16.  Order 5: pitch 13 puts the 25 points into lanes 0..24 of one round.
static void rho_lane_map(int order, TileGeom &tg) {
  const int O2 = order * order, NR2 = (O2 + 31) / 32, EP = rho_pitch(order);
  std::vector<int> left(O2);
  for (int p = 0; p < O2; p++) left[p] = p;
  signed char *map = tg.lane_pt;
  for (int k = 0; k < 64; k++) map[k] = -1;
  for (int h = 0; h < 2 * NR2 && !left.empty(); h++) {   // half warps in order
    bool used[16] = {};
    int filled = 0;
    for (size_t k = 0; k < left.size() && filled < 16;) {
      const int p = left[k], res = ((p / order) * EP + p % order) % 16;
      if (!used[res]) { used[res] = true; map[16 * h + filled++] = (signed char)p; left.erase(left.begin() + k); }
      else k++;
    }
  }
  // not reached for order <= 7 with rho_pitch: points that would conflict fill the idle lanes
  for (int k = 0; k < 32 * NR2 && !left.empty(); k++)
    if (map[k] < 0) { map[k] = (signed char)left.back(); left.pop_back(); }
}

// ORDER is a template parameter.  A lane owns the point (m,l) of the stencil's xy face (two for orders 6 and 7) for
// the whole kernel: it keeps the Horner coefficients of rho[0][l] and rho[1][m] in registers and evaluates them per
// atom itself (pppm_intel.cpp:476-488, same operation order); lanes 0..ORDER-1 also evaluate delvolinv*q*rho[2][n],
// which every lane fetches by shuffle while it walks the ORDER planes of the stencil.  Per atom that is ORDER
// shared-memory read-modify-writes and ORDER shuffles per lane; the next atom's {dx,dy,dz,q} is prefetched.
// ROUNDF: mixed mode rounds the weights to float like the reference's `flt_t rho[3][INTEL_P3M_MAXORDER]` (:474).
template <int ORDER, int ROUNDF>
__global__ void __launch_bounds__(128)
k_rho_tiles(PppmConst c, TileGeom tg, const int *__restrict__ cell_start, const double4 *__restrict__ pa_x,
            const int *__restrict__ pa_cx, double *__restrict__ tilebuf) {
  extern __shared__ double s_tiles[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long tile = (long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long ntiles = (long)tg.ntx * tg.nty * tg.ntz;
  if (tile >= ntiles) return;
  constexpr int E = RHO_T + ORDER - 1, E3 = E * E * E, O2 = ORDER * ORDER, NR2 = (O2 + 31) / 32;
  constexpr int EP = rho_pitch(ORDER), EEP = E * EP, E3S = E * EEP;
  double *t = s_tiles + (size_t)warp * E3S;
  for (int k = lane; k < E3S; k += 32) t[k] = 0.0;
  unsigned off[NR2];
  bool act[NR2];
  double cxw[NR2][ORDER], cyw[NR2][ORDER], czw[ORDER];
#pragma unroll
  for (int r = 0; r < NR2; r++) {
    const int p = tg.lane_pt[lane + 32 * r];
    act[r] = p >= 0;
    const int m = max(p, 0) / ORDER, l = max(p, 0) - m * ORDER;
    off[r] = (m * EP + l) * 8;   // bytes
#pragma unroll
    for (int k = 0; k < ORDER; k++) {
      cxw[r][k] = c.rho_coeff[k * ORDER + l];
      cyw[r][k] = c.rho_coeff[k * ORDER + m];
    }
  }
#pragma unroll
  for (int k = 0; k < ORDER; k++) czw[k] = c.rho_coeff[k * ORDER + min(lane, ORDER - 1)];
  __syncwarp();
  const int tx = (int)(tile % tg.ntx), ty = (int)((tile / tg.ntx) % tg.nty), tz = (int)(tile / ((long)tg.ntx * tg.nty));
  const int x0 = tx * RHO_T, y0 = ty * RHO_T, z0 = tz * RHO_T;
  const int x1 = min(x0 + RHO_T, c.nx), y1 = min(y0 + RHO_T, c.ny), z1 = min(z0 + RHO_T, c.nz);
  // the tile's (y,z) cell rows, each one contiguous range of the cell-sorted atoms: all ranges are fetched up front
  // (lane r holds rows r and r + 32), so walking the tile's atoms never waits on a dependent index load
  const int nyt = y1 - y0, nrows = nyt * (z1 - z0);
  int rs[2], re[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int r = lane + 32 * k;
    rs[k] = re[k] = 0;
    if (r < nrows) {
      const long row = ((long)(z0 + r / nyt) * c.ny + (y0 + r % nyt)) * c.nx;
      rs[k] = cell_start[row + x0];
      re[k] = cell_start[row + x1];
    }
  }
  // rb: shared-memory byte address of the stencil block's corner for the current cell row (ry, rz), minus x0
  const unsigned t_sa = (unsigned)__cvta_generic_to_shared(t) - (unsigned)x0 * 8u;
  int r = -1, a = 0, e = 0, ry = -1;
  unsigned rb = t_sa - EP * 8u;
  auto advance = [&]() {   // next atom of the tile; warp-uniform
    a++;
    while (a >= e && ++r < nrows) {
      a = __shfl_sync(0xffffffffu, r < 32 ? rs[0] : rs[1], r & 31);
      e = __shfl_sync(0xffffffffu, r < 32 ? re[0] : re[1], r & 31);
      rb += EP * 8u;
      if (++ry == nyt) { ry = 0; rb += (E - nyt) * EP * 8u; }
    }
  };
  auto horner = [&](const double *cf, const double d) {
    double w = 0.0;
#pragma unroll
    for (int l = ORDER - 1; l >= 0; l--) w = cf[l] + w * d;
    if (ROUNDF) w = (double)(float)w;
    return w;
  };
  advance();
  if (r < nrows) {
    double4 nxt = pa_x[a];
    unsigned sa_nxt = rb + (unsigned)pa_cx[a] * 8u;
    while (true) {
      const double4 cur = nxt;
      const unsigned sa = sa_nxt;
      advance();
      const bool more = r < nrows;
      if (more) { nxt = pa_x[a]; sa_nxt = rb + (unsigned)pa_cx[a] * 8u; }
      // z0 = delvolinv*q*rho[2][n] on lane n, y0 = z0*rho[1][m], x0 = y0*rho[0][l]  (pppm_intel.cpp:490-501)
      const double zl = cur.w * horner(czw, cur.z);
      double wx[NR2], wy[NR2];
#pragma unroll
      for (int q = 0; q < NR2; q++) { wx[q] = horner(cxw[q], cur.x); wy[q] = horner(cyw[q], cur.y); }
      // the ORDER planes of an atom are distinct addresses: all loads first, then the adds, then all stores, so that the
      // shared-memory latencies overlap instead of forming ORDER load -> add -> store chains (order of the volatile
      // accesses = program order; the previous atom's stores precede these loads)
      double acc[ORDER][NR2];
#pragma unroll
      for (int n = 0; n < ORDER; n++)
#pragma unroll
        for (int q = 0; q < NR2; q++) {
          acc[n][q] = 0.0;
          if (act[q]) {
            const unsigned ad = sa + off[q] + (unsigned)(n * EEP * 8);   // ptxas folds the constant into the access
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(acc[n][q]) : "r"(ad) : "memory");
          }
        }
#pragma unroll
      for (int n = 0; n < ORDER; n++) {
        const double zw = __shfl_sync(0xffffffffu, zl, n);
#pragma unroll
        for (int q = 0; q < NR2; q++) acc[n][q] += (zw * wy[q]) * wx[q];
      }
#pragma unroll
      for (int n = 0; n < ORDER; n++)
#pragma unroll
        for (int q = 0; q < NR2; q++)
          if (act[q]) {
            const unsigned ad = sa + off[q] + (unsigned)(n * EEP * 8);
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(ad), "d"(acc[n][q]) : "memory");
          }
      __syncwarp();
      if (!more) break;
    }
  }
  double *out = tilebuf + (size_t)tile * E3;
  for (int k = lane; k < E3; k += 32) out[k] = t[(k / E) * EP + k % E];
}

// covering tiles of a point along one dimension: every tile (own, and up to two on either side, periodic) whose
// cells [ulo,uhi] reach the point, i.e. ulo + nlower <= g <= uhi + nupper in the point's unwrapped frame.  Returns
// the count and, per hit, the tile index and the local coordinate inside that tile's stencil block.  Evaluated on
// the host once per setup (cover_table): the fold kernel reads the result as one int4 per coordinate.
static int cover1(int g, int n, int nt, int nlower, int nupper, int *tile, int *loc) {
  const int t = g / RHO_T;
  int cnt = 0;
  for (int dt = -2; dt <= 2; dt++) {
    int tt = t + dt, shift = 0;
    if (tt < 0) { tt += nt; shift = -n; }
    else if (tt >= nt) { tt -= nt; shift = n; }
    if (tt < 0 || tt >= nt) continue;
    const int ulo = tt * RHO_T + shift, uhi = std::min(tt * RHO_T + RHO_T, n) - 1 + shift;
    if (g >= ulo + nlower && g <= uhi + nupper) {
      tile[cnt] = tt;
      loc[cnt] = g - ulo - nlower;
      cnt++;
    }
  }
  return cnt;
}

// entries of one coordinate: tile * 16 + local coordinate (E <= 14), -1 = unused, hits first
static bool cover_table(int n, int nt, int nlower, int nupper, int4 *out) {
  for (int g = 0; g < n; g++) {
    int tile[5], loc[5];
    const int cnt = cover1(g, n, nt, nlower, nupper, tile, loc);
    if (cnt > 4) return false;
    int e[4] = {-1, -1, -1, -1};
    for (int k = 0; k < cnt; k++) e[k] = tile[k] * 16 + loc[k];
    out[g] = make_int4(e[0], e[1], e[2], e[3]);
  }
  return true;
}

// block = 32 x-points by 8 y-rows of one z plane; the sum runs over the covering tiles in a fixed order (z outer,
// x inner, tiles in ascending offset), so the density does not depend on the launch geometry
__global__ void __launch_bounds__(256)
k_rho_fold(PppmConst c, TileGeom tg, const int4 *__restrict__ cover, const double *__restrict__ tilebuf,
           double *__restrict__ density) {
  const int gx = blockIdx.x * 32 + threadIdx.x, gy = blockIdx.y * 8 + threadIdx.y, gz = blockIdx.z;
  if (gx >= c.nx || gy >= c.ny) return;
  const int4 cx = cover[gx], cy = cover[c.nx + gy], cz = cover[c.nx + c.ny + gz];
  const int ex[4] = {cx.x, cx.y, cx.z, cx.w}, ey[4] = {cy.x, cy.y, cy.z, cy.w}, ez[4] = {cz.x, cz.y, cz.z, cz.w};
  const int E = tg.E;
  const size_t E3 = (size_t)E * E * E;
  double rho = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (ez[k] < 0) break;
    const size_t tz = (size_t)(ez[k] >> 4) * tg.nty;
    const int lz = (ez[k] & 15) * E;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (ey[j] < 0) break;
      const size_t tzy = (tz + (ey[j] >> 4)) * tg.ntx;
      const int lzy = (lz + (ey[j] & 15)) * E;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (ex[i] < 0) break;
        rho += tilebuf[(tzy + (ex[i] >> 4)) * E3 + lzy + (ex[i] & 15)];
      }
    }
  }
  density[((size_t)gz * c.ny + gy) * c.nx + gx] = rho;
}

// ---------------------------------------------------------------------------------------------
// FFT passes.  A line L has elements at base(L) + k*estride; base(L) = (L / inner)*outer + (L % inner).

struct PassGeom {
  long nlines;
  int inner;
  long outer;
  long estride;
  long split;   // REAL_OUT = 2: real part to out_real[g], imaginary part to out_real[g + split]
};

// TB (lines per block) and blockDim.x are powers of two: lgTB = log2(TB), lgtpl = log2(blockDim.x / TB).  A thread
// is (line t, lane within the line) for contiguous lines and (lane, line t) for strided ones, so that consecutive
// threads touch consecutive addresses either way; the 64-bit base offset of a line is computed once per block.
template <int LINE_CONTIG, int REAL_IN, int REAL_OUT>
__global__ void __launch_bounds__(512)
k_fft_pass(FftPlan1d pl, PassGeom pg, int lgTB, int LP, const double *in_real, const double2 *in, double2 *out,
           double *out_real, double s) {  // in/out may alias (in-place passes): no __restrict__
  extern __shared__ double2 smem[];
  __shared__ long s_base[32];
  const int TB = 1 << lgTB;
  const int lgtpl = 31 - __clz(blockDim.x) - lgTB, tpl = 1 << lgtpl;
  double2 *bufA = smem, *bufB = smem + (size_t)TB * LP, *tw_s = smem + 2 * (size_t)TB * LP;
  const long L0 = (long)blockIdx.x * TB;
  const int n = pl.n;
  const int nl = (int)min((long)TB, pg.nlines - L0);
  stage_twiddles(pl, tw_s);
  if (threadIdx.x < nl) {
    const long L = L0 + threadIdx.x;
    s_base[threadIdx.x] = (L / pg.inner) * pg.outer + (L % pg.inner);
  }
  __syncthreads();
  int t, k0, kstep;
  if (LINE_CONTIG) { t = threadIdx.x >> lgtpl; k0 = threadIdx.x & (tpl - 1); kstep = tpl; }
  else { t = threadIdx.x & (TB - 1); k0 = threadIdx.x >> lgTB; kstep = tpl; }
  if (t < nl) {
    const long base = s_base[t];
    for (int k = k0; k < n; k += kstep) {
      const long g = base + (long)k * pg.estride;
      bufA[t * LP + k] = REAL_IN ? make_double2(in_real[g], 0.0) : in[g];
    }
  }
  __syncthreads();
  double2 *res = block_fft(bufA, bufB, pl, LP, nl, lgtpl, s, tw_s);
  if (t < nl) {
    const long base = s_base[t];
    for (int k = k0; k < n; k += kstep) {
      const long g = base + (long)k * pg.estride;
      const double2 v = res[t * LP + k];
      if (REAL_OUT == 2) { out_real[g] = v.x; out_real[g + pg.split] = v.y; }
      else if (REAL_OUT) out_real[g] = v.x;
      else out[g] = v;
    }
  }
}

// Real-to-complex x pass.  The density is real, so rho(-k) = conj rho(k): only kx = 0 .. nx/2 is kept (hx = nx/2 + 1
// points per line) and every later pass works on half as many lines.  Two real lines a, b ride in ONE complex transform
// z = a + i b and are separated by A(k) = (Z(k) + conj Z(n-k)) / 2, B(k) = (Z(k) - conj Z(n-k)) / (2i).
// Real line L starts at in_real + L n, its half spectrum at out + L hx.
__global__ void __launch_bounds__(512)
k_fft_x_r2c(FftPlan1d pl, long nlines, int lgTB, int LP, const double *__restrict__ in_real, double2 *__restrict__ out,
            double s) {
  extern __shared__ double2 smem[];
  const int TB = 1 << lgTB;
  const int lgtpl = 31 - __clz(blockDim.x) - lgTB, tpl = 1 << lgtpl;
  double2 *bufA = smem, *bufB = smem + (size_t)TB * LP, *tw_s = smem + 2 * (size_t)TB * LP;
  const int n = pl.n, hx = n / 2 + 1;
  const long npairs = (nlines + 1) >> 1;
  const long P0 = (long)blockIdx.x * TB;
  const int nl = (int)min((long)TB, npairs - P0);
  stage_twiddles(pl, tw_s);
  const int t = threadIdx.x >> lgtpl, k0 = threadIdx.x & (tpl - 1);
  const long La = 2 * (P0 + t), Lb = La + 1;
  const bool two = Lb < nlines;
  if (t < nl) {
    const double *a = in_real + La * n, *b = in_real + Lb * n;
    for (int k = k0; k < n; k += tpl) bufA[t * LP + k] = make_double2(a[k], two ? b[k] : 0.0);
  }
  __syncthreads();
  double2 *res = block_fft(bufA, bufB, pl, LP, nl, lgtpl, s, tw_s);
  if (t < nl) {
    double2 *oa = out + La * hx, *ob = out + Lb * hx;
    for (int k = k0; k < hx; k += tpl) {
      const double2 zk = res[t * LP + k], zm = res[t * LP + (k ? n - k : 0)];
      oa[k] = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
      if (two) ob[k] = make_double2(0.5 * (zk.y + zm.y), 0.5 * (zm.x - zk.x));
    }
  }
}

// Complex-to-real x pass, the inverse of the above: the half spectra A, B of two real lines are joined,
// Z(k) = A(k) + i B(k), Z(n-k) = conj A(k) + i conj B(k), and one complex transform returns a = Re z, b = Im z.
// The self-conjugate entries (k = 0 and, on an even grid, k = n/2) of a real line are real: their imaginary parts
// (round-off, or the Nyquist plane of a gradient field: the reference keeps Re(IFFT), see k_fft_z_poisson) are dropped.
__global__ void __launch_bounds__(512)
k_fft_x_c2r(FftPlan1d pl, long nlines, int lgTB, int LP, const double2 *__restrict__ in, double *__restrict__ out_real,
            double s) {
  extern __shared__ double2 smem[];
  const int TB = 1 << lgTB;
  const int lgtpl = 31 - __clz(blockDim.x) - lgTB, tpl = 1 << lgtpl;
  double2 *bufA = smem, *bufB = smem + (size_t)TB * LP, *tw_s = smem + 2 * (size_t)TB * LP;
  const int n = pl.n, hx = n / 2 + 1;
  const long npairs = (nlines + 1) >> 1;
  const long P0 = (long)blockIdx.x * TB;
  const int nl = (int)min((long)TB, npairs - P0);
  stage_twiddles(pl, tw_s);
  const int t = threadIdx.x >> lgtpl, k0 = threadIdx.x & (tpl - 1);
  const long La = 2 * (P0 + t), Lb = La + 1;
  const bool two = Lb < nlines;
  if (t < nl) {
    const double2 *ia = in + La * hx, *ib = in + Lb * hx;
    for (int k = k0; k < hx; k += tpl) {
      double2 A = ia[k], B = two ? ib[k] : make_double2(0.0, 0.0);
      const bool self = k == 0 || 2 * k == n;
      if (self) { A.y = 0.0; B.y = 0.0; }
      bufA[t * LP + k] = make_double2(A.x - B.y, A.y + B.x);
      if (!self) bufA[t * LP + n - k] = make_double2(A.x + B.y, B.x - A.y);
    }
  }
  __syncthreads();
  double2 *res = block_fft(bufA, bufB, pl, LP, nl, lgtpl, s, tw_s);
  if (t < nl) {
    double *oa = out_real + La * n, *ob = out_real + Lb * n;
    for (int k = k0; k < n; k += tpl) {
      const double2 v = res[t * LP + k];
      oa[k] = v.x;
      if (two) ob[k] = v.y;
    }
  }
}

// forward z pass + Poisson (pppm_intel.cpp:843-872) + (-i k) multiplies (:890-953) + NCOMP inverse z passes.
// lines are along z at (y,x) = L; TB consecutive L share y (mostly) and have consecutive x.
// NCOMP = 3: ik (E-field components), NCOMP = 1: ad (potential only).  EV: energy/virial partial sums.
// Cross terms of the wave vector on a triclinic box (PPPM::setup_triclinic: k = x2lamdaT(2 pi per)):
// ky += yx[ix], kz += zx[ix] + zy[iy]; yxg is yx with the Nyquist entry zeroed (gradient of the packed Ex + i Ey
// transform, see below).  All NULL on an orthogonal box.  zy is already offset to this rank's first y row.
// HALF = 1 (orthogonal boxes): half-spectrum layout (real-to-complex x pass): nx is the stored x extent
// nxfull / 2 + 1; a point with 0 < kx < nxfull / 2 also stands for its mirror image (-kx, -ky, -kz) in the energy /
// virial sums (on a Nyquist plane the mirror image keeps the sign of that wave number: its cross terms are summed
// with their own signs); the three gradient fields are transformed one by one
// (each is Hermitian: fkxg / fkyg / fkzg have their Nyquist entries zeroed, which is what keeping Re(IFFT) of the
// reference's unsymmetric Nyquist planes amounts to) instead of Ex + i Ey sharing a transform.
struct TriWave { const double *yx, *zx, *zy, *yxg; };
template <int NCOMP, int EV, int HALF>
__global__ void __launch_bounds__(512)
k_fft_z_poisson(FftPlan1d pl, int nx, int ny, int lgTB, int LP, const double2 *__restrict__ in,
                double2 *__restrict__ out, const double *__restrict__ greensfn, const double *__restrict__ fkx,
                const double *__restrict__ fky, const double *__restrict__ fkz, const double *__restrict__ fkxg,
                const double *__restrict__ fkyg, const double *__restrict__ fkzg, int nxfull, int iy_nyq,
                double scaleinv, double g_ewald, double *__restrict__ ev_partial, int disp, TriWave tw) {
  extern __shared__ double2 smem[];
  const int TB = 1 << lgTB;
  const int lgtpl = 31 - __clz(blockDim.x) - lgTB;
  double2 *bufA = smem, *bufB = smem + (size_t)TB * LP, *bufV = smem + 2 * (size_t)TB * LP;
  double2 *tw_s = smem + 3 * (size_t)TB * LP;
  __shared__ double s_red[16][8];
  stage_twiddles(pl, tw_s);
  const long plane = (long)nx * ny;
  const long nfft = plane * pl.n;
  const long L0 = (long)blockIdx.x * TB;
  const int n = pl.n;
  const int nl = (int)min((long)TB, plane - L0);
  // thread = (z lane k0, line t): consecutive threads touch consecutive lines = consecutive addresses
  const int t = threadIdx.x & (TB - 1), k0 = threadIdx.x >> lgTB, kstep = blockDim.x >> lgTB;
  const bool live = t < nl;
  const long L = L0 + t;
  const int ix = live ? (int)(L % nx) : 0, iy = live ? (int)(L / nx) : 0;
  if (live)
    for (int k = k0; k < n; k += kstep) bufA[t * LP + k] = in[L + (long)k * plane];
  __syncthreads();
  double2 *res = block_fft(bufA, bufB, pl, LP, nl, lgtpl, S_FWD, tw_s);
  double2 *other = (res == bufA) ? bufB : bufA;
  // Green's function multiply; energy / virial tallies
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const bool tri = tw.yx != nullptr;
  const double kzb = (live && tri) ? tw.zx[ix] + tw.zy[iy] : 0.0;   // part of kz that does not depend on iz
  if (live) {
    const double kx = fkx[ix], ky = fky[iy] + (tri ? tw.yx[ix] : 0.0);
    for (int k = k0; k < n; k += kstep) {
      const long g = L + (long)k * plane;
      const double2 w = res[t * LP + k];
      const double gf = greensfn[g];
      if (EV) {
        const double wgt = (HALF && ix != 0 && 2 * ix != nxfull) ? 2.0 : 1.0;
        const double eng = wgt * scaleinv * scaleinv * gf * (w.x * w.x + w.y * w.y);
        const double kz = fkz[k] + kzb;
        const double sqk = kx * kx + ky * ky + kz * kz;
        acc[0] += eng;
        if (sqk != 0.0) {  // PPPM::setup vg[][] evaluated on the fly instead of stored (6 doubles / point)
          double vterm = -2.0 * (1.0 / sqk + 0.25 / (g_ewald * g_ewald));
          if (disp) {   // vg_6 of PPPMDisp::setup
            const double b = 0.5 * sqrt(sqk) / g_ewald, bs = b * b, bt = bs * b;
            const double erft = 2.0 * bt * sqrt(kPI) * erfc(b), expt = exp(-bs);
            const double nom = erft - 2.0 * bs * expt, denom = nom + expt;
            vterm = denom == 0.0 ? 3.0 / sqk : 3.0 * nom / (sqk * denom);
          }
          acc[1] += eng * (1.0 + vterm * kx * kx);
          acc[2] += eng * (1.0 + vterm * ky * ky);
          acc[3] += eng * (1.0 + vterm * kz * kz);
          if (HALF && wgt == 2.0) {
            // the point and its mirror image: (-kx, +-ky, +-kz), a wave number on its Nyquist plane keeps its sign
            // (iy_nyq: the LOCAL row that holds the y Nyquist plane, -1 if none: ny is this rank's row count)
            const double e1 = 0.5 * eng, sy = iy == iy_nyq ? 1.0 : -1.0, sz = 2 * k == n ? 1.0 : -1.0;
            acc[4] += e1 * (vterm * kx * ky) * (1.0 - sy);
            acc[5] += e1 * (vterm * kx * kz) * (1.0 - sz);
            acc[6] += e1 * (vterm * ky * kz) * (1.0 + sy * sz);
          } else {
            acc[4] += eng * (vterm * kx * ky);
            acc[5] += eng * (vterm * kx * kz);
            acc[6] += eng * (vterm * ky * kz);
          }
        }
      }
      const double sg = scaleinv * gf;
      bufV[t * LP + k] = make_double2(w.x * sg, w.y * sg);
    }
  }
  __syncthreads();
  if (EV) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 7; q++) {
      double v = acc[q];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) s_red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
      double sum = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) sum += s_red[w][threadIdx.x];
      ev_partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = sum;
    }
  }
  // ik: the three gradient fields are real, so two of them share one complex inverse transform:
  // pack 0 = Wx + i Wy (real part -> Ex, imaginary part -> Ey), pack 1 = Wz, with W? = (fk?*Im V, -fk?*Re V)
  // (pppm_intel.cpp:894-895, 921-922, 949-950).  Three inverse 3-D FFTs become two.
  // The reference keeps Re(IFFT(W)); on an even grid W is not Hermitian at its own Nyquist plane (fk there is
  // -n/2 * unitk, not 0) and that plane contributes exactly nothing to the real part — so dropping it (fkxg/fkyg
  // have the Nyquist entry zeroed) leaves Re(IFFT(Wx)) unchanged and makes Wx, Wy exactly Hermitian, which is what
  // the packing needs for the two fields not to leak into each other.
  constexpr int NPACK = NCOMP == 3 ? (HALF ? 3 : 2) : 1;
  for (int comp = 0; comp < NPACK; comp++) {
    double2 *a = res, *b = other;  // V lives in bufV; a/b are free ping-pong buffers
    if (live) {
      // triclinic: the Hermitian part of ky = yx[ix] + fky[iy] drops BOTH Nyquist entries; Ez has its own transform
      // whose real part is kept, exactly what stock poisson_ik_triclinic does with the raw wave vector
      const double kx = NCOMP == 3 ? fkxg[ix] : 0.0, ky = NCOMP == 3 ? fkyg[iy] + (tri ? tw.yxg[ix] : 0.0) : 0.0;
      for (int k = k0; k < n; k += kstep) {
        const double2 v = bufV[t * LP + k];
        if (NCOMP == 1) a[t * LP + k] = v;
        else if (HALF) {
          const double fk = comp == 0 ? kx : (comp == 1 ? ky : fkzg[k]);
          a[t * LP + k] = make_double2(fk * v.y, -fk * v.x);
        } else if (comp == 0) a[t * LP + k] = make_double2(kx * v.y + ky * v.x, ky * v.y - kx * v.x);
        else {
          const double fk = fkz[k] + kzb;
          a[t * LP + k] = make_double2(fk * v.y, -fk * v.x);
        }
      }
    }
    __syncthreads();
    double2 *r2 = block_fft(a, b, pl, LP, nl, lgtpl, S_BWD, tw_s);
    if (live)
      for (int k = k0; k < n; k += kstep) out[(size_t)comp * nfft + L + (long)k * plane] = r2[t * LP + k];
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------
// multi-GPU helpers: plane copies for the halos and pack/unpack for the two transposes

// dst[(p0 + p) * plane + i] (+)= src[(s0 + p) * plane + i] for p < np
template <int ADD>
__global__ void k_planes(long plane, int np, const double *__restrict__ src, int s0, double *__restrict__ dst, int d0) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= plane * np) return;
  const long p = t / plane, i = t - p * plane;
  const double v = src[(s0 + p) * plane + i];
  if (ADD) dst[(d0 + p) * plane + i] += v;
  else dst[(d0 + p) * plane + i] = v;
}

// z-slab [nzo][ny][nx] -> per-destination blocks [q][nzo][nyl_q][nx] (block q starts at boff[q] elements)
struct RankRows { int ylo[8], yhi[8]; long boff[8]; int n; };
__global__ void k_tr_pack(int nx, int ny, int nzo, RankRows rr, const double2 *__restrict__ in, double2 *__restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long tot = (long)nx * ny * nzo;
  if (t >= tot) return;
  const int x = (int)(t % nx), y = (int)((t / nx) % ny), z = (int)(t / ((long)nx * ny));
  int q = 0;
  while (y >= rr.yhi[q]) q++;
  const int nyl = rr.yhi[q] - rr.ylo[q];
  out[rr.boff[q] + ((long)z * nyl + (y - rr.ylo[q])) * nx + x] = in[t];
}
// received blocks [p][ncomp][nzo][nyl_p][nx] -> [ncomp][nzo][ny][nx]
__global__ void k_tr_unpack(int nx, int ny, int nzo, int ncomp, RankRows rr, const double2 *__restrict__ in,
                            double2 *__restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long slab = (long)nx * ny * nzo;
  if (t >= slab * ncomp) return;
  const int comp = (int)(t / slab);
  const long r = t - comp * slab;
  const int x = (int)(r % nx), y = (int)((r / nx) % ny), z = (int)(r / ((long)nx * ny));
  int p = 0;
  while (y >= rr.yhi[p]) p++;
  const int nyl = rr.yhi[p] - rr.ylo[p];
  // rr.boff[p]: start of rank p's data (all components); one component of it is nzo*nyl*nx elements
  out[t] = in[rr.boff[p] + (long)comp * nzo * nyl * nx + ((long)z * nyl + (y - rr.ylo[p])) * nx + x];
}

// ---- peer-memory transposes: the producer stores every element where its consumer (another GPU) will read it -------
struct PeerPtrs { double2 *p[8]; };
struct RankPlanes { int zlo[8], zhi[8]; int n; };
// forward: my x/y-transformed planes [nzo][ny][nx] -> rank q's z-pencil block [gnz][nyl_q][nx], plane zglob0 + z
// (threads run along x: 16 B stores, nx of them contiguous in the peer's memory)
__global__ void __launch_bounds__(256)
k_tr_scatter_fwd(int nx, int ny, int nzo, int zglob0, RankRows rr, PeerPtrs dst, const double2 *__restrict__ in) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long tot = (long)nx * ny * nzo;
  if (t >= tot) return;
  const int x = (int)(t % nx), y = (int)((t / nx) % ny), z = (int)(t / ((long)nx * ny));
  int q = 0;
  while (y >= rr.yhi[q]) q++;
  const int nyl = rr.yhi[q] - rr.ylo[q];
  dst.p[q][((long)(zglob0 + z) * nyl + (y - rr.ylo[q])) * nx + x] = in[t];
}
// backward: my pencils [npack][gnz][nyl][nx] -> rank q's [npack][nzo_q][ny][nx] (the layout the inverse y pass reads),
// rows ylo_me .. ylo_me + nyl of its planes
__global__ void __launch_bounds__(256)
k_tr_scatter_bwd(int nx, int ny, int nyl, int gnz, int ylo_me, int npack, RankPlanes rp, PeerPtrs dst,
                 const double2 *__restrict__ in) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long blk = (long)nx * nyl * gnz;
  if (t >= blk * npack) return;
  const int comp = (int)(t / blk);
  const long r = t - comp * blk;
  const int x = (int)(r % nx), row = (int)((r / nx) % nyl), z = (int)(r / ((long)nx * nyl));
  int q = 0;
  while (z >= rp.zhi[q]) q++;
  const int nzq = rp.zhi[q] - rp.zlo[q];
  dst.p[q][(((long)comp * nzq + (z - rp.zlo[q])) * ny + (ylo_me + row)) * nx + x] = in[t];
}

// per-atom tallies: out = rho(k) * G(k) / N * (1 | vg_c(k)), c = comp - 1 (PPPM::poisson_peratom; vg as in PPPM::setup)
__global__ void k_peratom_mul(PppmConst c, int comp, double scaleinv, const double2 *__restrict__ rhok,
                              const double *__restrict__ greensfn, const double *__restrict__ fkx,
                              const double *__restrict__ fky, const double *__restrict__ fkz, double2 *__restrict__ out) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nfft = (long)c.sx * c.ny * c.nz;
  if (n >= nfft) return;
  const int i = (int)(n % c.sx), j = (int)((n / c.sx) % c.ny), k = (int)(n / ((long)c.sx * c.ny));
  double w = scaleinv * greensfn[n];
  if (comp > 0) {
    const double kx = fkx[i], ky = fky[j], kz = fkz[k];
    const double sqk = kx * kx + ky * ky + kz * kz;
    double vg = 0.0;
    if (sqk != 0.0) {
      const double vterm = -2.0 * (1.0 / sqk + 0.25 / (c.g_ewald * c.g_ewald));
      // half spectrum: the reference keeps Re(IFFT) of products that are not Hermitian where exactly one of the two
      // wave numbers sits on its Nyquist plane (it does not change sign under k -> -k there); their Hermitian part is 0
      const bool half = c.sx != c.nx;
      const bool nqx = 2 * i == c.nx, nqy = 2 * j == c.ny, nqz = 2 * k == c.nz;
      switch (comp) {
        case 1: vg = 1.0 + vterm * kx * kx; break;
        case 2: vg = 1.0 + vterm * ky * ky; break;
        case 3: vg = 1.0 + vterm * kz * kz; break;
        case 4: vg = (half && nqx != nqy) ? 0.0 : vterm * kx * ky; break;
        case 5: vg = (half && nqx != nqz) ? 0.0 : vterm * kx * kz; break;
        default: vg = (half && nqy != nqz) ? 0.0 : vterm * ky * kz; break;
      }
    }
    w *= vg;
  }
  const double2 r = rhok[n];
  out[n] = make_double2(r.x * w, r.y * w);
}

// ---------------------------------------------------------------------------------------------
// host helpers

int make_plan(b200md_ctx *ctx, FftPlan1d &pl, DevBuf<double2> &twbuf, int n) {
  pl.n = n;
  pl.nfac = 0;
  int m = n;
  while (m % 4 == 0) { pl.fac[pl.nfac++] = 4; m /= 4; }
  while (m % 2 == 0) { pl.fac[pl.nfac++] = 2; m /= 2; }
  while (m % 3 == 0) { pl.fac[pl.nfac++] = 3; m /= 3; }
  while (m % 5 == 0) { pl.fac[pl.nfac++] = 5; m /= 5; }
  if (m != 1) return b2_fail(ctx, B200MD_EINVAL, "FFT length %d is not of the form 2^a 3^b 5^c", n);
  if (n > 4096) return b2_fail(ctx, B200MD_EINVAL, "FFT length %d is longer than one shared-memory line (4096)", n);
  {
    int pp = 1;
    for (int f = 0; f < pl.nfac; f++) {
      pl.sp[f] = pp;
      pl.sm[f] = n / pl.fac[f];
      pl.ststep[f] = n / (pp * pl.fac[f]);
      pl.sinvp[f] = 1.0f / (float)pp;
      pp *= pl.fac[f];
    }
  }
  std::vector<double2> tw(n);
  for (int k = 0; k < n; k++) {
    const long double ph = -2.0L * 3.14159265358979323846264338327950288L * k / n;
    tw[k] = make_double2((double)cosl(ph), (double)sinl(ph));
  }
  RESERVE(ctx, twbuf, (size_t)n);
  CUDA_OK(ctx, cudaMemcpy(twbuf.p, tw.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
  pl.tw = twbuf.p;
  return 0;
}

// tuning knobs of the FFT passes (environment overrides are for the kernel-tuning sweeps only)
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}
static int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) l++; return l; }
static int fft_threads() { static int t = env_int("B200MD_FFT_THREADS", 256); return t; }
int pick_tb(int n, int nbuf) {
  // lines per block: as many as fit in the shared-memory budget, at most TBMAX, at least 1
  static const int tbmax = env_int("B200MD_FFT_TBMAX", 8), kb = env_int("B200MD_FFT_SMEM_KB", 70);
  const int LP = n | 1;
  int tb = tbmax;
  while (tb > 1 && (size_t)nbuf * tb * LP * sizeof(double2) > (size_t)kb * 1024) tb >>= 1;
  return tb;
}

template <int LC, int RI, int RO>
int launch_pass(b200md_ctx *ctx, const FftPlan1d &pl, const PassGeom &pg, const double *in_real, const double2 *in,
                double2 *out, double *out_real, double s, int timer_id = -1) {
  const int TB = pick_tb(pl.n, 2);
  const int LP = pl.n | 1;
  const size_t smem = (2 * (size_t)TB * LP + pl.n) * sizeof(double2);   // two line buffers + the twiddle table
  if (smem > 220 * 1024) return b2_fail(ctx, B200MD_EINVAL, "FFT length %d too long for one shared-memory line", pl.n);
  auto kern = k_fft_pass<LC, RI, RO>;
  CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblk = cdiv(pg.nlines, TB);
  {
    ScopedTimer tk(ctx, timer_id >= 0 ? timer_id : T_OTHER, timer_id >= 0);
    kern<<<nblk, fft_threads(), smem, ctx->stream>>>(pl, pg, ilog2(TB), LP, in_real, in, out, out_real, s);
  }
  KERNEL_OK(ctx, "k_fft_pass");
  return 0;
}

// x passes of the half-spectrum path: nlines real lines of pl.n points <-> nlines half lines of pl.n / 2 + 1 points
int launch_x_r2c(b200md_ctx *ctx, const FftPlan1d &pl, long nlines, const double *in_real, double2 *out, int timer_id) {
  const int TB = std::min(pick_tb(pl.n, 2), 4);   // 4 line pairs per block: measured best for the contiguous x lines
  const int LP = pl.n | 1;
  const size_t smem = (2 * (size_t)TB * LP + pl.n) * sizeof(double2);
  if (smem > 220 * 1024) return b2_fail(ctx, B200MD_EINVAL, "FFT length %d too long for one shared-memory line", pl.n);
  CUDA_OK(ctx, cudaFuncSetAttribute(k_fft_x_r2c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long npairs = (nlines + 1) / 2;
  if (npairs <= 0) return 0;
  {
    ScopedTimer tk(ctx, timer_id);
    k_fft_x_r2c<<<cdiv(npairs, TB), fft_threads(), smem, ctx->stream>>>(pl, nlines, ilog2(TB), LP, in_real, out, S_FWD);
  }
  KERNEL_OK(ctx, "k_fft_x_r2c");
  return 0;
}
int launch_x_c2r(b200md_ctx *ctx, const FftPlan1d &pl, long nlines, const double2 *in, double *out_real, int timer_id) {
  const int TB = std::min(pick_tb(pl.n, 2), 4);
  const int LP = pl.n | 1;
  const size_t smem = (2 * (size_t)TB * LP + pl.n) * sizeof(double2);
  if (smem > 220 * 1024) return b2_fail(ctx, B200MD_EINVAL, "FFT length %d too long for one shared-memory line", pl.n);
  CUDA_OK(ctx, cudaFuncSetAttribute(k_fft_x_c2r, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long npairs = (nlines + 1) / 2;
  if (npairs <= 0) return 0;
  {
    ScopedTimer tk(ctx, timer_id >= 0 ? timer_id : T_OTHER, timer_id >= 0);
    k_fft_x_c2r<<<cdiv(npairs, TB), fft_threads(), smem, ctx->stream>>>(pl, nlines, ilog2(TB), LP, in, out_real, S_BWD);
  }
  KERNEL_OK(ctx, "k_fft_x_c2r");
  return 0;
}

// The fused z pass / Poisson / gradient kernel on this rank's z pencils: nyl rows starting at global row yoff, lines of
// gnz points, spectral x extent ps.c.sx.  in = x,y-transformed density [gnz][nyl][sx]; out = [npack][gnz][nyl][sx]
// (npack: 1 ad, 2 ik full spectrum (Ex + i Ey, Ez), 3 ik half spectrum).  Returns the block count (rows of ps.partial).
int launch_z_poisson(b200md_ctx *ctx, PppmState &ps, int nyl, int yoff, int gnz, const double2 *in, double2 *out,
                     int ev, int *nblk_out) {
  const PppmConst &c = ps.c;
  const bool ad = ps.p.differentiation == 1, half = c.sx != c.nx;
  const int TB = pick_tb(gnz, 3);
  const int LP = gnz | 1;
  const size_t smem = (3 * (size_t)TB * LP + gnz) * sizeof(double2);
  const int nblk = cdiv((long)c.sx * nyl, TB);
  *nblk_out = nblk;
  const double scaleinv = 1.0 / ((double)c.nx * c.ny * gnz);
  if (ev) RESERVE(ctx, ps.partial, (size_t)std::max(nblk, 1) * 8);
  if (nblk <= 0) return 0;
  const int iy_nyq = (c.ny % 2 == 0) ? c.ny / 2 - yoff : -1;   // local row of the y Nyquist plane (may be outside)
  const TriWave tw = c.tri ? TriWave{ps.fkyx.p, ps.fkzx.p, ps.fkzy.p + yoff, ps.fkyx_g.p}
                           : TriWave{nullptr, nullptr, nullptr, nullptr};
#define ZK(NC, E, H)                                                                                              \
  do {                                                                                                            \
    auto kern = k_fft_z_poisson<NC, E, H>;                                                                        \
    CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    kern<<<nblk, fft_threads(), smem, ctx->stream>>>(ps.plan[2], c.sx, nyl, ilog2(TB), LP, in, out, ps.greensfn.p, \
                                                     ps.fkx.p, ps.fky.p + yoff, ps.fkz.p, ps.fkx_g.p,            \
                                                     ps.fky_g.p + yoff, ps.fkz_g.p, c.nx, iy_nyq, scaleinv,     \
                                                     c.g_ewald,                                                   \
                                                     ps.partial.p, ps.p.dispersion, tw);                          \
  } while (0)
  {
    ScopedTimer tk(ctx, K_FFT_Z_POISSON);
    if (half) {
      if (ad) { if (ev) ZK(1, 1, 1); else ZK(1, 0, 1); }
      else { if (ev) ZK(3, 1, 1); else ZK(3, 0, 1); }
    } else {
      if (ad) { if (ev) ZK(1, 1, 0); else ZK(1, 0, 0); }
      else { if (ev) ZK(3, 1, 0); else ZK(3, 0, 0); }
    }
  }
#undef ZK
  KERNEL_OK(ctx, "k_fft_z_poisson");
  return 0;
}

int reduce_cols(b200md_ctx *ctx, PppmState &ps, long n, int ncol, const double *cols, double *host_out) {
  const int nb = cdiv(n, 256);
  RESERVE(ctx, ps.partial, (size_t)nb * ncol);
  RESERVE(ctx, ps.red, 16);
  k_colsum_partial<<<nb, 256, 0, ctx->stream>>>(n, ncol, cols, ps.partial.p);
  KERNEL_OK(ctx, "k_colsum_partial");
  k_colsum_final<<<1, 256, 0, ctx->stream>>>(nb, ncol, ps.partial.p, ps.red.p);
  KERNEL_OK(ctx, "k_colsum_final");
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ps.red.p, ncol * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < ncol; k++) host_out[k] = ctx->h_pinned[k];
  return 0;
}

void compute_rho_coeffs(PppmConst &c) {
  const int order = c.order;
  const int w = 2 * order + 1;
  std::vector<double> a((size_t)order * w, 0.0);
  auto A = [&](int l, int k) -> double & { return a[(size_t)l * w + (k + order)]; };
  A(0, 0) = 1.0;
  for (int j = 1; j < order; j++)
    for (int k = -j; k <= j; k += 2) {
      double s = 0.0;
      for (int l = 0; l < j; l++) {
        A(l + 1, k) = (A(l, k + 1) - A(l, k - 1)) / (l + 1);
        s += std::pow(0.5, (double)l + 1) * (A(l, k - 1) + std::pow(-1.0, (double)l) * A(l, k + 1)) / (l + 1);
      }
      A(0, k) = s;
    }
  for (int i = 0; i < B2_MAXORDER * B2_MAXORDER; i++) c.rho_coeff[i] = c.drho_coeff[i] = 0.0;
  int m = (1 - order) / 2;
  for (int k = -(order - 1); k < order; k += 2) {
    for (int l = 0; l < order; l++) c.rho_coeff[l * order + (m - c.nlower)] = A(l, k);
    for (int l = 1; l < order; l++) c.drho_coeff[(l - 1) * order + (m - c.nlower)] = l * A(l, k);
    m++;
  }
}

void compute_gf_denom(PppmConst &c) {
  const int order = c.order;
  for (int l = 1; l < order; l++) c.gf_b[l] = 0.0;
  c.gf_b[0] = 1.0;
  for (int m = 1; m < order; m++) {
    int l;
    for (l = m; l > 0; l--)
      c.gf_b[l] = 4.0 * (c.gf_b[l] * (l - m) * (l - m - 0.5) - c.gf_b[l - 1] * (l - m - 1) * (l - m - 1));
    c.gf_b[0] = 4.0 * (c.gf_b[0] * (l - m) * (l - m - 0.5));
  }
  long ifact = 1;
  for (int k = 1; k < 2 * order; k++) ifact *= k;
  const double gaminv = 1.0 / ifact;
  for (int l = 0; l < order; l++) c.gf_b[l] *= gaminv;
}

// density [nplanes][ny][nx] (real) -> work [nplanes][ny][sx]: x and y transformed
int fft3d_forward_xy(b200md_ctx *ctx, PppmState &ps, const double *density, double2 *work, int nplanes) {
  const PppmConst &c = ps.c;
  if (c.sx != c.nx) TRY(launch_x_r2c(ctx, ps.plan[0], (long)c.ny * nplanes, density, work, K_FFT_X_FWD));
  else {
    PassGeom gx{(long)c.ny * nplanes, 1, (long)c.nx, 1, 0};
    TRY((launch_pass<1, 1, 0>(ctx, ps.plan[0], gx, density, nullptr, work, nullptr, S_FWD, K_FFT_X_FWD)));
  }
  PassGeom gy{(long)c.sx * nplanes, c.sx, (long)c.sx * c.ny, (long)c.sx, 0};
  TRY((launch_pass<0, 0, 0>(ctx, ps.plan[1], gy, nullptr, work, work, nullptr, S_FWD, K_FFT_Y_FWD)));
  return 0;
}

// fields [npack][nplanes][ny][sx] (z already inverted) -> real bricks vd [ncomp][nplanes][ny][nx]: inverse y, inverse x
int fft3d_inverse_yx(b200md_ctx *ctx, PppmState &ps, double2 *work2, double *vd, int nplanes) {
  const PppmConst &c = ps.c;
  const bool ad = ps.p.differentiation == 1, half = c.sx != c.nx;
  const int npack = ad ? 1 : (half ? 3 : 2);
  const long nreal = (long)c.nx * c.ny * nplanes, nspec = (long)c.sx * c.ny * nplanes;
  PassGeom gy{(long)c.sx * nplanes * npack, c.sx, (long)c.sx * c.ny, (long)c.sx, 0};
  TRY((launch_pass<0, 0, 0>(ctx, ps.plan[1], gy, nullptr, work2, work2, nullptr, S_BWD, K_FFT_Y_INV)));
  if (half) {
    for (int comp = 0; comp < npack; comp++)
      TRY(launch_x_c2r(ctx, ps.plan[0], (long)c.ny * nplanes, work2 + (size_t)comp * nspec, vd + (size_t)comp * nreal,
                       K_FFT_X_INV));
  } else if (ad) {
    PassGeom gx{(long)c.ny * nplanes, 1, (long)c.nx, 1, 0};
    TRY((launch_pass<1, 0, 1>(ctx, ps.plan[0], gx, nullptr, work2, nullptr, vd, S_BWD, K_FFT_X_INV)));
  } else {
    // the x pass stores Re (and, for the first pack, Im as the second field)
    PassGeom gxy{(long)c.ny * nplanes, 1, (long)c.nx, 1, nreal};
    TRY((launch_pass<1, 0, 2>(ctx, ps.plan[0], gxy, nullptr, work2, nullptr, vd, S_BWD, K_FFT_X_INV)));
    PassGeom gz{(long)c.ny * nplanes, 1, (long)c.nx, 1, 0};
    TRY((launch_pass<1, 0, 1>(ctx, ps.plan[0], gz, nullptr, work2 + nreal, nullptr, vd + 2 * nreal, S_BWD, K_FFT_X_INV)));
  }
  return 0;
}


// Poisson solve on nranks GPUs.  In: ps.density = local brick [c.nz][ny][nx].  Out: ps.vd = [ncomp][c.nz][ny][nx]
// (the local brick of every field component, halos filled), evsum[7] = global energy / virial sums when ev.
int poisson_multi(b200md_ctx *ctx, PppmState &ps, int ev, double *evsum) {
  const PppmConst &c = ps.c;
  const int P = ps.nranks, me = ps.rank, lower = (me + P - 1) % P, upper = (me + 1) % P;
  const int nx = c.nx, ny = c.ny, gnz = ps.gnz;
  const long plane = (long)nx * ny;
  const int sx = c.sx;                        // x extent of the spectral arrays (nx / 2 + 1 with half-spectrum transforms)
  const long splane = (long)sx * ny;
  const int nzo = ps.pzhi[me] - ps.pzlo[me];
  const bool ad = ps.p.differentiation == 1;
  const int ncomp = ad ? 1 : 3;   // field bricks: Ex, Ey, Ez (ik) or the potential u (ad)
  // halo widths: planes of a rank's brick below / above its owned range
  auto lo_w = [&](int r) { return ps.pzlo[r] - ps.zoffs[r]; };
  auto hi_w = [&](int r) { return ps.zoffs[r] + ps.nbzs[r] - ps.pzhi[r]; };
  const int my_lo = lo_w(me), my_hi = hi_w(me);
  const int up_lo = lo_w(upper), low_hi = hi_w(lower);   // what I receive: upper's low halo, lower's high halo
  int nblk_z = 0;
  {
    ScopedTimer tm(ctx, T_COMM);
    // ---- density halo sum (reverse comm): send my halo planes, add the neighbours' into my owned planes ---------
    RESERVE(ctx, ps.dens_own, (size_t)plane * nzo);
    RESERVE(ctx, ps.halo_r, (size_t)plane * std::max(up_lo + low_hi, my_lo + my_hi) * ncomp + 16);
    const int nb = 256;
    k_planes<0><<<cdiv(plane * nzo, nb), nb, 0, ctx->stream>>>(plane, nzo, ps.density.p, my_lo, ps.dens_own.p, 0);
    KERNEL_OK(ctx, "k_planes");
    // to lower: my planes [0, my_lo); to upper: my planes [my_lo + nzo, c.nz) — both contiguous in the brick
    TRY(b2_comm_exchange(ctx, ps.density.p, (size_t)plane * my_lo * sizeof(double),
                         ps.density.p + plane * (my_lo + nzo), (size_t)plane * my_hi * sizeof(double),
                         ps.halo_r.p + plane * low_hi, (size_t)plane * up_lo * sizeof(double), ps.halo_r.p,
                         (size_t)plane * low_hi * sizeof(double)));
    if (low_hi > 0) {   // the lower rank's high halo covers my first low_hi owned planes
      k_planes<1><<<cdiv(plane * low_hi, nb), nb, 0, ctx->stream>>>(plane, low_hi, ps.halo_r.p, 0, ps.dens_own.p, 0);
      KERNEL_OK(ctx, "k_planes");
    }
    if (up_lo > 0) {    // the upper rank's low halo covers my last up_lo owned planes
      k_planes<1><<<cdiv(plane * up_lo, nb), nb, 0, ctx->stream>>>(plane, up_lo, ps.halo_r.p, low_hi, ps.dens_own.p,
                                                                  nzo - up_lo);
      KERNEL_OK(ctx, "k_planes");
    }
  }
  RankRows rr;
  rr.n = P;
  const int nyl = ps.yhis[me] - ps.ylos[me];
  const long nT = (long)sx * nyl * gnz;   // points of my z-pencil block
  const int npack = ad ? 1 : (sx != nx ? 3 : 2);   // transforms coming back: u | Ex, Ey, Ez | Ex + i Ey, Ez
  {
    ScopedTimer tm(ctx, T_FFT);
    // ---- forward x, y on the owned planes -------------------------------------------------------------------------
    RESERVE(ctx, ps.work1, (size_t)splane * nzo);
    TRY(fft3d_forward_xy(ctx, ps, ps.dens_own.p, ps.work1.p, nzo));
    // ---- transpose to z pencils: block for rank q = my planes x q's rows; lands at q in [z][row][x] order ---------
    RESERVE(ctx, ps.tsend, (size_t)splane * nzo);
    RESERVE(ctx, ps.workT, (size_t)nT);
    RESERVE(ctx, ps.workT2, (size_t)nT * npack);
    size_t scount[8], sdisp[8], rcount[8], rdisp[8];
    long off = 0;
    for (int q = 0; q < P; q++) {
      rr.ylo[q] = ps.ylos[q]; rr.yhi[q] = ps.yhis[q]; rr.boff[q] = off;
      const long cnt = (long)nzo * (ps.yhis[q] - ps.ylos[q]) * sx;
      scount[q] = cnt * sizeof(double2); sdisp[q] = off * sizeof(double2);
      off += cnt;
      rcount[q] = (size_t)(ps.pzhi[q] - ps.pzlo[q]) * nyl * sx * sizeof(double2);
      rdisp[q] = (size_t)ps.pzlo[q] * nyl * sx * sizeof(double2);
    }
    double2 *workT = ps.p2p ? (double2 *)ps.symT.local : ps.workT.p;
    PeerPtrs peersT, peersW;
    RankPlanes rp;
    rp.n = P;
    for (int q = 0; q < 8; q++) {
      peersT.p[q] = (double2 *)ps.symT.peer[q];
      peersW.p[q] = (double2 *)ps.symW.peer[q];
      rp.zlo[q] = q < P ? ps.pzlo[q] : 0;
      rp.zhi[q] = q < P ? ps.pzhi[q] : 0;
    }
    if (ps.p2p && ps.p2p_dma) {
      // copy engines: one strided (2-D) device-to-peer copy per destination moves my planes' rows ylo_q..yhi_q
      // straight into rank q's pencil block at plane pzlo_me — no pack kernel, no SM time (the pair kernel that runs
      // underneath on the main stream keeps the SMs), NVLink at DMA speed.  The barrier orders the copies before the z
      // pass of every rank; nobody still reads its pencil block (the second barrier of the previous step came after
      // every rank's z pass).
      for (int k = 0; k < P && nzo > 0; k++) {
        const int q = (me + k) % P;   // start with myself, then round the ring: spreads the traffic over the peers
        const int nylq = ps.yhis[q] - ps.ylos[q];
        if (nylq == 0) continue;
        CUDA_OK(ctx, cudaMemcpy2DAsync(peersT.p[q] + (size_t)ps.pzlo[me] * nylq * sx, (size_t)nylq * sx * sizeof(double2),
                                       ps.work1.p + (size_t)ps.ylos[q] * sx, (size_t)splane * sizeof(double2),
                                       (size_t)nylq * sx * sizeof(double2), (size_t)nzo, cudaMemcpyDeviceToDevice,
                                       ctx->stream));
      }
      TRY(b2_comm_barrier(ctx));
    } else if (ps.p2p) {
      // every rank stores its planes straight into the pencil blocks of their owners (NVLink stores from the kernel);
      // the barrier orders those stores before the z pass of every rank.
      if (splane * nzo > 0) {
        k_tr_scatter_fwd<<<cdiv(splane * nzo, 256), 256, 0, ctx->stream>>>(sx, ny, nzo, ps.pzlo[me], rr, peersT, ps.work1.p);
        KERNEL_OK(ctx, "k_tr_scatter_fwd");
      }
      TRY(b2_comm_barrier(ctx));
    } else {
      k_tr_pack<<<cdiv(splane * nzo, 256), 256, 0, ctx->stream>>>(sx, ny, nzo, rr, ps.work1.p, ps.tsend.p);
      KERNEL_OK(ctx, "k_tr_pack");
      TRY(b2_comm_alltoallv(ctx, ps.tsend.p, scount, sdisp, ps.workT.p, rcount, rdisp));
    }
    // ---- z pass + Green's function + gradients + inverse z on my pencils ------------------------------------------
    TRY(launch_z_poisson(ctx, ps, nyl, ps.ylos[me], gnz, workT, ps.workT2.p, ev, &nblk_z));
    // ---- transpose back, all transformed fields in one exchange: to rank q its planes (contiguous in [z][row][x]) --
    double2 *work2 = ps.p2p ? (double2 *)ps.symW.local : nullptr;
    if (ps.p2p && ps.p2p_dma) {
      // the same the other way: per destination and packed field one 2-D copy puts my rows of its planes where its
      // inverse y pass reads them ([pack][plane][y][x]: no unpack kernel).  The first barrier of this step came after
      // every rank's inverse passes of the previous step, so nobody still reads its plane block.
      for (int k = 0; k < P && nyl > 0; k++) {
        const int q = (me + k) % P;
        const int nzq = ps.pzhi[q] - ps.pzlo[q];
        if (nzq == 0) continue;
        for (int comp = 0; comp < npack; comp++)
          CUDA_OK(ctx, cudaMemcpy2DAsync(peersW.p[q] + ((size_t)comp * nzq * ny + ps.ylos[me]) * sx, (size_t)splane * sizeof(double2),
                                         ps.workT2.p + (size_t)comp * nT + (size_t)ps.pzlo[q] * nyl * sx,
                                         (size_t)nyl * sx * sizeof(double2), (size_t)nyl * sx * sizeof(double2), (size_t)nzq,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
      }
      TRY(b2_comm_barrier(ctx));
    } else if (ps.p2p) {
      // kernel flavour of the same: every rank stores its pencils into the plane blocks of their owners
      if (nT > 0) {
        k_tr_scatter_bwd<<<cdiv(nT * npack, 256), 256, 0, ctx->stream>>>(sx, ny, nyl, gnz, ps.ylos[me], npack, rp, peersW,
                                                                        ps.workT2.p);
        KERNEL_OK(ctx, "k_tr_scatter_bwd");
      }
      TRY(b2_comm_barrier(ctx));
    } else {
    RESERVE(ctx, ps.trecv, (size_t)splane * nzo * npack);
    {
      CommGroup grp(ctx);
      long roff = 0;
      for (int q = 0; q < P; q++) {
        const long scnt = (long)(ps.pzhi[q] - ps.pzlo[q]) * nyl * sx;
        const long rcnt = (long)nzo * (ps.yhis[q] - ps.ylos[q]) * sx;
        rr.boff[q] = roff;
        for (int comp = 0; comp < npack; comp++) {
          TRY(grp.send(ps.workT2.p + (size_t)comp * nT + (size_t)ps.pzlo[q] * nyl * sx, scnt * sizeof(double2), q));
          TRY(grp.recv(ps.trecv.p + roff + (size_t)comp * rcnt, rcnt * sizeof(double2), q));
        }
        roff += rcnt * npack;
      }
      TRY(grp.end());
    }
    RESERVE(ctx, ps.work2, (size_t)splane * nzo * npack);
    k_tr_unpack<<<cdiv(splane * nzo * npack, 256), 256, 0, ctx->stream>>>(sx, ny, nzo, npack, rr, ps.trecv.p, ps.work2.p);
    KERNEL_OK(ctx, "k_tr_unpack");
    work2 = ps.work2.p;
    }
    // ---- inverse y, x on the owned planes -> the real field bricks of my planes -----------------------------------
    const long nown = plane * nzo;
    RESERVE(ctx, ps.vd_own, (size_t)nown * ncomp);
    TRY(fft3d_inverse_yx(ctx, ps, work2, ps.vd_own.p, nzo));
  }
  {
    ScopedTimer tm(ctx, T_COMM);
    // ---- field halo fill (forward comm): my boundary owned planes go out, the brick is assembled ------------------
    // lower needs its high halo = my first low_hi owned planes; upper needs its low halo = my last up_lo planes
    const long bplane = plane * c.nz;   // one component of the local brick
    RESERVE(ctx, ps.vd, (size_t)bplane * ncomp);
    RESERVE(ctx, ps.halo_s, (size_t)plane * (low_hi + up_lo) * ncomp + 16);
    const int nb = 256;
    for (int comp = 0; comp < ncomp; comp++) {
      const double *own = ps.vd_own.p + (size_t)comp * plane * nzo;
      if (low_hi > 0) k_planes<0><<<cdiv(plane * low_hi, nb), nb, 0, ctx->stream>>>(plane, low_hi, own, 0, ps.halo_s.p, comp * low_hi);
      if (up_lo > 0)
        k_planes<0><<<cdiv(plane * up_lo, nb), nb, 0, ctx->stream>>>(plane, up_lo, own, nzo - up_lo,
                                                                    ps.halo_s.p + plane * low_hi * ncomp, comp * up_lo);
      k_planes<0><<<cdiv(plane * nzo, nb), nb, 0, ctx->stream>>>(plane, nzo, own, 0, ps.vd.p + (size_t)comp * bplane, my_lo);
      ctx->launches += 3;
    }
    // receive: from upper -> my high halo (my_hi planes per component), from lower -> my low halo (my_lo planes)
    TRY(b2_comm_exchange(ctx, ps.halo_s.p, (size_t)plane * low_hi * ncomp * sizeof(double),
                         ps.halo_s.p + plane * low_hi * ncomp, (size_t)plane * up_lo * ncomp * sizeof(double),
                         ps.halo_r.p + plane * my_lo * ncomp, (size_t)plane * my_hi * ncomp * sizeof(double), ps.halo_r.p,
                         (size_t)plane * my_lo * ncomp * sizeof(double)));
    for (int comp = 0; comp < ncomp; comp++) {
      double *brick = ps.vd.p + (size_t)comp * bplane;
      if (my_lo > 0) k_planes<0><<<cdiv(plane * my_lo, nb), nb, 0, ctx->stream>>>(plane, my_lo, ps.halo_r.p, comp * my_lo, brick, 0);
      if (my_hi > 0)
        k_planes<0><<<cdiv(plane * my_hi, nb), nb, 0, ctx->stream>>>(plane, my_hi, ps.halo_r.p + plane * my_lo * ncomp,
                                                                    comp * my_hi, brick, my_lo + nzo);
      ctx->launches += 2;
    }
    CUDA_OK(ctx, cudaGetLastError());
  }
  if (ev) {
    ScopedTimer tm(ctx, T_POISSON);
    RESERVE(ctx, ps.red, 16);
    if (nblk_z > 0) {
      k_colsum_final<<<1, 256, 0, ctx->stream>>>(nblk_z, 7, ps.partial.p, ps.red.p);
      KERNEL_OK(ctx, "k_colsum_final");
    } else CUDA_OK(ctx, cudaMemsetAsync(ps.red.p, 0, 8 * sizeof(double), ctx->stream));
    TRY(b2_comm_allreduce_sum(ctx, ps.red.p, 7));   // MPI_Allreduce of energy and virial, pppm_intel.cpp:260,273
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ps.red.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 7; k++) evsum[k] = ctx->h_pinned[k];
  }
  return 0;
}

// PPPM::poisson_peratom on one GPU.  ps.work1 holds the x- and y-transformed density (the fused z kernel of the normal
// path does not write its z transform back): finish the forward transform, then per brick multiply by G / N (and
// vg_c) and run one inverse 3-D FFT whose last pass stores the real part.  7 bricks: u, v0..v5.
static int peratom_fields(b200md_ctx *ctx, PppmState &ps, int do_e, int do_v) {
  const PppmConst &c = ps.c;
  const long nfft = ps.nfft, nspec = (long)c.sx * c.ny * c.nz;
  const bool half = c.sx != c.nx;
  RESERVE(ctx, ps.pa_fields, 7 * (size_t)nfft);
  RESERVE(ctx, ps.pa_work, (size_t)nspec);
  PassGeom gx{(long)c.ny * c.nz, 1, (long)c.nx, 1, 0};
  PassGeom gy{(long)c.sx * c.nz, c.sx, (long)c.sx * c.ny, (long)c.sx, 0};
  PassGeom gz{(long)c.sx * c.ny, c.sx * c.ny, 0, (long)c.sx * c.ny, 0};
  TRY((launch_pass<0, 0, 0>(ctx, ps.plan[2], gz, nullptr, ps.work1.p, ps.work1.p, nullptr, S_FWD)));
  const double scaleinv = 1.0 / ((double)c.nx * c.ny * c.nz);
  for (int comp = do_e ? 0 : 1; comp < (do_v ? 7 : 1); comp++) {
    k_peratom_mul<<<cdiv(nspec, 256), 256, 0, ctx->stream>>>(c, comp, scaleinv, ps.work1.p, ps.greensfn.p, ps.fkx.p,
                                                             ps.fky.p, ps.fkz.p, ps.pa_work.p);
    KERNEL_OK(ctx, "k_peratom_mul");
    TRY((launch_pass<0, 0, 0>(ctx, ps.plan[2], gz, nullptr, ps.pa_work.p, ps.pa_work.p, nullptr, S_BWD)));
    TRY((launch_pass<0, 0, 0>(ctx, ps.plan[1], gy, nullptr, ps.pa_work.p, ps.pa_work.p, nullptr, S_BWD)));
    if (half)
      TRY(launch_x_c2r(ctx, ps.plan[0], (long)c.ny * c.nz, ps.pa_work.p, ps.pa_fields.p + (size_t)comp * nfft, -1));
    else
      TRY((launch_pass<1, 0, 1>(ctx, ps.plan[0], gx, nullptr, ps.pa_work.p, nullptr,
                                ps.pa_fields.p + (size_t)comp * nfft, S_BWD)));
  }
  return 0;
}

template <class flt_t>
int pppm_compute_view(b200md_ctx *ctx, PppmState &ps, const PppmView &v, int eflag, int vflag, double *energy,
                      double *virial) {
  const PppmConst &c = ps.c;
  const int n = v.n;
  const long nfft = ps.nfft;
  const int eflag_global = eflag & 1, vflag_global = vflag & 3;
  if (energy) *energy = 0.0;
  if (virial) for (int k = 0; k < 6; k++) virial[k] = 0.0;

  // qsum_qsq when the atom count changed (pppm_intel.cpp:142-145).  Multi-GPU: the local count changes with every
  // migration but the global sums do not, and the reduction is collective: done once per setup.
  const int ncomp = ps.p.dispersion ? ps.ncomp : 1;
  const int tp1 = ctx->ntypes + 1;
  if (ps.nranks > 1 ? ps.q_natoms < 0 : ps.q_natoms != n) {
    RESERVE(ctx, ps.density, std::max((size_t)nfft, 2 * (size_t)n + 2));
    for (int m = 0; m < ncomp; m++) {
      if (n > 0) {
        k_q_moments<<<cdiv(n, 256), 256, 0, ctx->stream>>>(
            n, v.xq, v.type, ps.p.dispersion ? ps.Btype.p + (size_t)m * tp1 : nullptr, ps.density.p);
        KERNEL_OK(ctx, "k_q_moments");
        double mm[2];
        TRY(reduce_cols(ctx, ps, n, 2, ps.density.p, mm));
        ps.comp_qsum[m] = mm[0];
        ps.comp_qsqsum[m] = mm[1];
      } else ps.comp_qsum[m] = ps.comp_qsqsum[m] = 0.0;
    }
    if (ps.nranks > 1) {
      RESERVE(ctx, ps.red, 2 * PppmState::MAXCOMP);
      double m2[2 * PppmState::MAXCOMP];
      for (int m = 0; m < ncomp; m++) { m2[2 * m] = ps.comp_qsum[m]; m2[2 * m + 1] = ps.comp_qsqsum[m]; }
      CUDA_OK(ctx, cudaMemcpyAsync(ps.red.p, m2, 2 * ncomp * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      TRY(b2_comm_allreduce_sum(ctx, ps.red.p, 2 * ncomp));
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ps.red.p, 2 * ncomp * sizeof(double), cudaMemcpyDeviceToHost,
                                   ctx->stream));
      CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
      for (int m = 0; m < ncomp; m++) { ps.comp_qsum[m] = ctx->h_pinned[2 * m]; ps.comp_qsqsum[m] = ctx->h_pinned[2 * m + 1]; }
    }
    ps.q_natoms = n;
  }
  // "return if there are no charges" :149 - per component of a dispersion grid
  bool sorted = false;
  for (int comp = 0; comp < ncomp; comp++) {
  ps.cur = comp;
  ps.qsum = ps.comp_qsum[comp];
  ps.qsqsum = ps.comp_qsqsum[comp];
  if (ps.qsqsum == 0.0) continue;
  const double *Bcomp = ps.p.dispersion ? ps.Btype.p + (size_t)comp * tp1 : nullptr;

  // ---- particle_map + cell sort + make_rho ------------------------------------------------------
  {
    ScopedTimer tm(ctx, T_MAKE_RHO);
    RESERVE(ctx, ps.key, (size_t)n + 1);
    RESERVE(ctx, ps.perm, (size_t)n + 1);
    RESERVE(ctx, ps.pa_x, (size_t)n + 1);
    RESERVE(ctx, ps.pa_n, (size_t)n + 1);
    const bool tiled = c.nx >= 2 * RHO_T && c.ny >= 2 * RHO_T && c.nz >= 2 * RHO_T;
    if (!tiled) RESERVE(ctx, ps.pa_w, ((size_t)n + 1) * 3 * c.order);
    RESERVE(ctx, ps.pa_cx, (size_t)n + 1);
    RESERVE(ctx, ps.cell_count, (size_t)nfft + 1);
    RESERVE(ctx, ps.cell_start, (size_t)nfft + 1);
    RESERVE(ctx, ps.cursor, (size_t)nfft + 1);
    RESERVE(ctx, ps.flags, 4);
    RESERVE(ctx, ps.scan_ws, b2_scan_ws_bytes((size_t)nfft + 1));
    if (!sorted) {   // the cell sort does not depend on the weights: once for all components
      CUDA_OK(ctx, cudaMemsetAsync(ps.cell_count.p, 0, ((size_t)nfft + 1) * sizeof(int), ctx->stream));
      CUDA_OK(ctx, cudaMemsetAsync(ps.flags.p, 0, 4 * sizeof(int), ctx->stream));
      if (n > 0) {
        k_map_key<flt_t><<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, v.xq, v.xqf, c, ps.key.p, ps.cell_count.p, ps.flags.p);
        KERNEL_OK(ctx, "k_map_key");
      }
      TRY(b2_exclusive_scan_i32(ctx, ps.cell_count.p, ps.cell_start.p, (size_t)nfft, ps.scan_ws.p));
      CUDA_OK(ctx, cudaMemcpyAsync(ps.cursor.p, ps.cell_start.p, (size_t)nfft * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
      if (n > 0) {
        k_cell_scatter<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, ps.key.p, ps.cursor.p, ps.perm.p);
        KERNEL_OK(ctx, "k_cell_scatter");
        k_cell_order<<<cdiv(nfft, 256), 256, 0, ctx->stream>>>(nfft, ps.cell_start.p, ps.perm.p);
        KERNEL_OK(ctx, "k_cell_order");
      }
      sorted = true;
    }
    if (n > 0) {
      k_fill_sorted<flt_t><<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, ps.perm.p, v.xq, v.xqf, v.type, Bcomp, c, ps.pa_x.p,
                                                                   ps.pa_n.p, tiled ? nullptr : ps.pa_w.p, ps.pa_cx.p);
      KERNEL_OK(ctx, "k_fill_sorted");
    }
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ps.flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (tiled) {
      TileGeom tg{cdiv(c.nx, RHO_T), cdiv(c.ny, RHO_T), cdiv(c.nz, RHO_T), RHO_T + c.order - 1, {}};
      rho_lane_map(c.order, tg);
      const long ntiles = (long)tg.ntx * tg.nty * tg.ntz;
      const size_t E3 = (size_t)tg.E * tg.E * tg.E;
      RESERVE(ctx, ps.tilebuf, (size_t)ntiles * E3);
      static const int wpb = env_int("B200MD_RHO_WPB", 3);
      const size_t smem = (size_t)wpb * tg.E * tg.E * rho_pitch(c.order) * sizeof(double);
#define RHO_TILES(O)                                                                                              \
  case O: {                                                                                                       \
    if (sizeof(flt_t) == 4) {                                                                                     \
      CUDA_OK(ctx, cudaFuncSetAttribute(k_rho_tiles<O, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_rho_tiles<O, 1><<<cdiv(ntiles, wpb), wpb * 32, smem, ctx->stream>>>(c, tg, ps.cell_start.p, ps.pa_x.p,     \
                                                                           ps.pa_cx.p, ps.tilebuf.p);             \
    } else {                                                                                                      \
      CUDA_OK(ctx, cudaFuncSetAttribute(k_rho_tiles<O, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_rho_tiles<O, 0><<<cdiv(ntiles, wpb), wpb * 32, smem, ctx->stream>>>(c, tg, ps.cell_start.p, ps.pa_x.p,     \
                                                                           ps.pa_cx.p, ps.tilebuf.p);             \
    }                                                                                                             \
  } break;
      {
        ScopedTimer tk(ctx, K_RHO_TILES);
        switch (c.order) {
          RHO_TILES(1) RHO_TILES(2) RHO_TILES(3) RHO_TILES(4) RHO_TILES(5) RHO_TILES(6) RHO_TILES(7)
          default: return b2_fail(ctx, B200MD_EORDER, "PPPM order greater than supported by USER-INTEL");
        }
      }
#undef RHO_TILES
      KERNEL_OK(ctx, "k_rho_tiles");
      {
        ScopedTimer tk(ctx, K_RHO_FOLD);
        k_rho_fold<<<dim3(cdiv(c.nx, 32), cdiv(c.ny, 8), c.nz), dim3(32, 8), 0, ctx->stream>>>(c, tg, ps.cover.p,
                                                                                             ps.tilebuf.p, ps.density.p);
      }
      KERNEL_OK(ctx, "k_rho_fold");
    } else {  // grids smaller than two tiles per dimension: plain per-point gather
      const dim3 grid(cdiv(c.nx, 8), cdiv(c.ny, 8), cdiv(c.nz, 4));
      k_make_rho<<<grid, 256, 0, ctx->stream>>>(c, ps.cell_start.p, ps.pa_x.p, ps.pa_w.p, ps.pa_cx.p, ps.density.p);
      KERNEL_OK(ctx, "k_make_rho");
    }
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(int *)ctx->h_pinned) return b2_fail(ctx, B200MD_ERANGE, "Out of range atoms - cannot compute PPPM");
  }

  // ---- poisson: forward FFT, Green's function, gradients, inverse FFTs -------------------------------
  const int ev = (eflag_global || vflag_global) ? 1 : 0;
  int nblk_z = 0;
  double evsum[8] = {0};
  if (ps.nranks > 1) {
    TRY(poisson_multi(ctx, ps, ev, evsum));
  } else {
  {
    ScopedTimer tm(ctx, T_FFT);
    TRY(fft3d_forward_xy(ctx, ps, ps.density.p, ps.work1.p, c.nz));
    TRY(launch_z_poisson(ctx, ps, c.ny, 0, c.nz, ps.work1.p, ps.work2.p, ev, &nblk_z));
    // inverse y over the transformed fields, inverse x storing the real bricks
    TRY(fft3d_inverse_yx(ctx, ps, ps.work2.p, ps.vd.p, c.nz));
  }
  if (ev) {
    ScopedTimer tm(ctx, T_POISSON);
    RESERVE(ctx, ps.red, 16);
    k_colsum_final<<<1, 256, 0, ctx->stream>>>(nblk_z, 7, ps.partial.p, ps.red.p);
    KERNEL_OK(ctx, "k_colsum_final");
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ps.red.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 7; k++) evsum[k] = ctx->h_pinned[k];
  }
  }

  // ---- per-atom energy / virial (eflag & 2, vflag & 4): extra inverse FFTs and their own gather ------------------
  ps.pa_have_e = ps.pa_have_v = false;
  if ((eflag & 2) || (vflag & 4)) {
    TRY(peratom_fields(ctx, ps, eflag & 2, vflag & 4));
    TRY(b2_fieldforce_peratom(ctx, ps, v, eflag & 2, vflag & 4));
  }

  // ---- fieldforce -----------------------------------------------------------------------------------
  // accumulates into f, so it goes back to the main stream, behind the pair kernel (see b2_pppm_compute)
  if (ctx->stream != ctx->main_stream) {
    cudaStream_t ks = ctx->stream;
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_k, ks));
    ctx->stream = ctx->main_stream;
    CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_k, 0));
    const int rc = b2_fieldforce<flt_t>(ctx, ps, v);
    ctx->stream = ks;   // a second grid (pppm/disp) continues on the k-space stream
    if (rc) return rc;
  } else TRY(b2_fieldforce<flt_t>(ctx, ps, v));

  // ---- energy / virial post-factors (pppm_intel.cpp:256-275) -----------------------------------------
  if (ps.p.dispersion) {
    // pppm_disp_intel.cpp:486-510: csum = sum_i C_ii, csumij = sum_ij C_ij.  Per component C_ij = sign W_i W_j, so
    // csum = sign * qsqsum and csumij = sign * qsum^2 (geometric mixing: one component, W = B)
    const double g3 = c.g_ewald * c.g_ewald * c.g_ewald, sg = ps.comp_sign[comp];
    const double csum = ps.qsqsum, csumij = ps.qsum * ps.qsum;
    const double a = kPI * kPIS / (6.0 * ps.volume) * g3 * csumij;
    if (eflag_global && energy) *energy += sg * (0.5 * ps.volume * evsum[0] - a + g3 * g3 * csum / 12.0);
    if (vflag_global && virial)
      for (int k = 0; k < 6; k++) virial[k] += sg * (0.5 * ps.volume * evsum[1 + k] - (k < 3 ? a : 0.0));
    continue;
  }
  const double qscale = ctx->qqrd2e * ps.p.scale;
  if (eflag_global && energy) {
    double e = evsum[0];
    e *= 0.5 * ps.volume;
    e -= c.g_ewald * ps.qsqsum / kPIS + kPI2 * ps.qsum * ps.qsum / (c.g_ewald * c.g_ewald * ps.volume);
    e *= qscale;
    *energy = e;
  }
  if (vflag_global && virial)
    for (int k = 0; k < 6; k++) virial[k] = 0.5 * qscale * ps.volume * evsum[1 + k];
  // ---- slabcorr (pppm_intel.cpp:305): EW3DC dipole correction for `kspace_modify slab` ---------------------------
  if (ps.p.slab_volfactor > 1.0 && n > 0) {
    RESERVE(ctx, ps.slab_cols, 2 * (size_t)n);
    k_slab_moments<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, v.xq, ps.slab_cols.p);
    KERNEL_OK(ctx, "k_slab_moments");
    double m[2];
    TRY(reduce_cols(ctx, ps, n, 2, ps.slab_cols.p, m));
    const double dipole_all = m[0], dipole_r2 = m[1], zprd = ps.zprd;
    const double e_slabcorr =
        k2PI * (dipole_all * dipole_all - ps.qsum * dipole_r2 - ps.qsum * ps.qsum * zprd * zprd / 12.0) / ps.volume;
    if (eflag_global && energy) *energy += qscale * e_slabcorr;
    const double ffact = qscale * (-4.0 * kPI / ps.volume), efact = qscale * k2PI / ps.volume;
    // like fieldforce it adds to f: in overlap mode it goes to the main stream, behind the pair kernel
    cudaStream_t ks = ctx->stream;
    if (ks != ctx->main_stream) {
      CUDA_OK(ctx, cudaEventRecord(ctx->ev_k, ks));
      ctx->stream = ctx->main_stream;
      CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_k, 0));
    }
    k_slabcorr<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, v.xq, v.f, ffact, dipole_all, ps.qsum,
                                                      ps.pa_have_e ? ps.pa_out.p : nullptr, efact, dipole_r2, zprd);
    const cudaError_t slab_err = cudaGetLastError();
    ctx->stream = ks;
    if (slab_err != cudaSuccess) return b2_fail(ctx, B200MD_ECUDA, "k_slabcorr: %s", cudaGetErrorString(slab_err));
    ctx->launches++;
  }
  }   // components
  return 0;
}

}  // namespace

// Signed self-coupled components of a dispersion grid (PppmState::ncomp): W[m][type], sign[m] with
// C_ij = sum_m sign_m W_m[i] W_m[j] the r^-6 coefficient of the type pair.
//   mix 1, geometric (pppm_disp_intel.cpp:245-313): B[T+1], C_ij = B_i B_j -> one component.
//   mix 2, arithmetic (:315-407): B[7 (T+1)] in the layout of PPPMDisp::init_coeffs, B[7 i + k], C_ij = sum_k B_i[k]
//     B_j[6-k].  The reference spreads seven densities a_k and couples a_k with a_{6-k} (poisson_2s_ik); here each
//     coupled pair is rotated into two self-coupled densities, 2 <a_k, a_{6-k}> = <p, p> - <m, m> with
//     p, m = (a_k +- a_{6-k}) / sqrt 2, so that the same seven grids run through the one-density kernels: signs
//     + (a_3), + + + (p), - - - (m).  Energy, virial and forces are the same sums (tests/test_gpu_pppm.py).
//   mix 3, no mixing rule (:409-467): B[(T+1)^2] = C_ij, symmetric; C = V diag(lambda) V^T by cyclic Jacobi sweeps
//     (what PPPMDisp::init_coeffs does with its own eigen solver): W_m = sqrt|lambda_m| V[:, m], sign_m = sign lambda_m;
//     eigenvalues below 1e-12 max|lambda| are dropped (nsplit).
static int disp_components(int mix, int T, const double *B, std::vector<double> &W, std::vector<double> &sign) {
  const int tp1 = T + 1;
  W.clear(); sign.clear();
  if (mix == 1) {
    W.assign(B, B + tp1);
    sign.push_back(1.0);
    return 0;
  }
  if (mix == 2) {
    const double r = std::sqrt(0.5);
    W.assign((size_t)7 * tp1, 0.0);
    for (int i = 0; i < tp1; i++) {
      W[i] = B[7 * i + 3];
      for (int k = 0; k < 3; k++) {
        W[(size_t)(1 + k) * tp1 + i] = r * (B[7 * i + k] + B[7 * i + 6 - k]);
        W[(size_t)(4 + k) * tp1 + i] = r * (B[7 * i + k] - B[7 * i + 6 - k]);
      }
    }
    sign = {1, 1, 1, 1, -1, -1, -1};
    return 0;
  }
  if (mix != 3 || T < 1 || T > PppmState::MAXCOMP) return 1;
  // cyclic Jacobi on the T x T block of types 1..T
  std::vector<double> A((size_t)T * T), V((size_t)T * T, 0.0);
  double amax = 0.0;
  for (int i = 0; i < T; i++)
    for (int j = 0; j < T; j++) {
      A[(size_t)i * T + j] = B[(size_t)(i + 1) * tp1 + j + 1];
      amax = std::max(amax, std::fabs(A[(size_t)i * T + j]));
    }
  for (int i = 0; i < T; i++)
    for (int j = 0; j < i; j++)
      if (std::fabs(A[(size_t)i * T + j] - A[(size_t)j * T + i]) > 1e-12 * amax) return 1;
  for (int i = 0; i < T; i++) V[(size_t)i * T + i] = 1.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0.0;
    for (int i = 0; i < T; i++)
      for (int j = i + 1; j < T; j++) off += A[(size_t)i * T + j] * A[(size_t)i * T + j];
    if (off <= 1e-30 * amax * amax) break;
    for (int pI = 0; pI < T; pI++)
      for (int q = pI + 1; q < T; q++) {
        const double apq = A[(size_t)pI * T + q];
        if (apq == 0.0) continue;
        const double theta = (A[(size_t)q * T + q] - A[(size_t)pI * T + pI]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < T; k++) {   // A <- A J
          const double akp = A[(size_t)k * T + pI], akq = A[(size_t)k * T + q];
          A[(size_t)k * T + pI] = cs * akp - sn * akq;
          A[(size_t)k * T + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < T; k++) {   // A <- J^T A
          const double apk = A[(size_t)pI * T + k], aqk = A[(size_t)q * T + k];
          A[(size_t)pI * T + k] = cs * apk - sn * aqk;
          A[(size_t)q * T + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < T; k++) {   // V <- V J
          const double vkp = V[(size_t)k * T + pI], vkq = V[(size_t)k * T + q];
          V[(size_t)k * T + pI] = cs * vkp - sn * vkq;
          V[(size_t)k * T + q] = sn * vkp + cs * vkq;
        }
      }
  }
  double lmax = 0.0;
  for (int m = 0; m < T; m++) lmax = std::max(lmax, std::fabs(A[(size_t)m * T + m]));
  for (int m = 0; m < T; m++) {
    const double lam = A[(size_t)m * T + m];
    if (std::fabs(lam) <= 1e-12 * lmax) continue;
    const size_t base = W.size();
    W.resize(base + tp1, 0.0);
    for (int i = 0; i < T; i++) W[base + i + 1] = std::sqrt(std::fabs(lam)) * V[(size_t)i * T + m];
    sign.push_back(lam > 0 ? 1.0 : -1.0);
  }
  if (sign.empty()) { W.assign(tp1, 0.0); sign.push_back(1.0); }
  return 0;
}

static void free_state(b200md_ctx *ctx, PppmState *&slot) {
  PppmState *ps = slot;
  if (!ps) return;
  b2_comm_peer_free(ctx, ps->symT);
  b2_comm_peer_free(ctx, ps->symW);
  for (int d = 0; d < 3; d++) ps->tw[d].free_();
  ps->fkx_g.free_(); ps->fky_g.free_(); ps->fkyx.free_(); ps->fkzx.free_(); ps->fkzy.free_(); ps->fkyx_g.free_();
  ps->fkz_g.free_();
  ps->greensfn.free_(); ps->fkx.free_(); ps->fky.free_(); ps->fkz.free_(); ps->density.free_(); ps->vd.free_();
  ps->work1.free_(); ps->work2.free_(); ps->sf_pre.free_(); ps->Btype.free_();
  ps->key.free_(); ps->cell_count.free_(); ps->cell_start.free_(); ps->cursor.free_(); ps->perm.free_();
  ps->dens_own.free_(); ps->halo_s.free_(); ps->halo_r.free_(); ps->vd_own.free_(); ps->tsend.free_(); ps->trecv.free_();
  ps->workT.free_(); ps->workT2.free_();
  ps->flags.free_(); ps->pa_x.free_(); ps->pa_n.free_(); ps->pa_w.free_(); ps->pa_cx.free_(); ps->tilebuf.free_(); ps->cover.free_(); ps->slab_cols.free_(); ps->pa_fields.free_(); ps->pa_work.free_(); ps->pa_out.free_(); ps->scan_ws.free_(); ps->partial.free_(); ps->red.free_();
  delete ps;
  slot = nullptr;
}

void b2_pppm_free(b200md_ctx *ctx) {
  free_state(ctx, ctx->pppm);
  free_state(ctx, ctx->pppm6);
}

// The brick halo (lo_out / hi_out, PPPM::set_grid_local) is sized from skin / 2 at setup: a state set up with a smaller
// skin than the one now in force would raise spurious "Out of range atoms" errors (or, on several GPUs, miss halo
// planes), so it is dropped; a smaller or equal skin keeps it.
void b2_pppm_skin_changed(b200md_ctx *ctx, double skin) {
  if (ctx->pppm && skin > ctx->pppm->skin_setup) free_state(ctx, ctx->pppm);
  if (ctx->pppm6 && skin > ctx->pppm6->skin_setup) free_state(ctx, ctx->pppm6);
}

template <class flt_t>
static int compute_all(b200md_ctx *ctx, const PppmView &v, int eflag, int vflag, double *energy, double *virial) {
  // PPPMDispIntel::compute: function[0] (Coulomb grid) then function[1] (geometric dispersion grid); energy and
  // virial are the sums (pppm_disp_intel.cpp:183-313, 540-541)
  if (energy) *energy = 0.0;
  if (virial) for (int k = 0; k < 6; k++) virial[k] = 0.0;
  PppmState *slots[2] = {ctx->pppm, ctx->pppm6};
  for (PppmState *ps : slots) {
    if (!ps) continue;
    double e = 0.0, vv[6] = {0, 0, 0, 0, 0, 0};
    TRY(pppm_compute_view<flt_t>(ctx, *ps, v, eflag, vflag, &e, vv));
    if (energy) *energy += e;
    if (virial) for (int k = 0; k < 6; k++) virial[k] += vv[k];
  }
  return 0;
}

int b2_pppm_compute(b200md_ctx *ctx, int eflag, int vflag, double *energy, double *virial) {
  if (!ctx->pppm && !ctx->pppm6) return b2_fail(ctx, B200MD_EINVAL, "pppm compute before b200md_pppm_setup");
  // the reference hands per-atom tallies to stock poisson_peratom / fieldforce_peratom (pppm_intel.cpp:224-229,
  // 281-301); provided here for the Coulomb grid on one GPU (b200md_pppm_peratom), refused elsewhere rather than
  // returned as zeros
  if ((eflag & 2) || (vflag & 4)) {
    if (ctx->pppm6 || !ctx->pppm)
      return b2_fail(ctx, B200MD_EINVAL, "per-atom energy/virial is not provided for the dispersion grid");
    if (b2_comm_nranks(ctx) > 1)
      return b2_fail(ctx, B200MD_EINVAL, "per-atom energy/virial from PPPM is single-GPU only in this build");
    if (ctx->pppm->c.tri)
      return b2_fail(ctx, B200MD_EINVAL, "per-atom energy/virial from PPPM is not provided on a triclinic box");
  }
  PppmView v;
  v.n = ctx->nlocal;
  v.xq = ctx->xq.p;
  v.xqf = ctx->prec == B200MD_PREC_MIXED ? ctx->xqf.p : nullptr;
  v.type = ctx->type.p;
  v.f = ctx->f.p;
  // k-space overlap: when a pair kernel was just launched for these positions (ev_pre), run on the k-space stream
  const bool overlap = ctx->overlap && ctx->ev_pre_valid && ctx->kstream;
  ctx->ev_pre_valid = false;
  if (overlap) {
    CUDA_OK(ctx, cudaStreamWaitEvent(ctx->kstream, ctx->ev_pre, 0));
    ctx->stream = ctx->kstream;
  }
  const int rc = ctx->prec == B200MD_PREC_MIXED ? compute_all<float>(ctx, v, eflag, vflag, energy, virial)
                                                 : compute_all<double>(ctx, v, eflag, vflag, energy, virial);
  if (overlap) {
    // the main stream continues only after everything issued on the k-space stream (its buffers are reused next step)
    cudaEventRecord(ctx->ev_k, ctx->kstream);
    ctx->stream = ctx->main_stream;
    cudaStreamWaitEvent(ctx->stream, ctx->ev_k, 0);
  }
  return rc;
}

extern "C" {

// Pure host arithmetic (no device needed; exercised by the CPU tests): the z-slab plan of the grid for every rank.
//   owned planes [pzlo,pzhi) (FFT slab), local brick origin zoff / height nbz (owned planes + stencil and skin/2 halo,
//   PPPM::set_grid_local's nzlo_out..nzhi_out for the rank's atom slab), and the y rows [ylo,yhi) held after the
//   transpose to z pencils.  Returns non-zero when a halo would reach beyond the neighbouring rank.
int b200md_pppm_decomp(int nranks, int nz, int ny, int order, double skin, double prd_z, int *pzlo, int *pzhi, int *zoff,
                       int *nbz, int *ylo, int *yhi) {
  if (nranks < 1 || nz < 1 || ny < 1 || order < 1 || !(prd_z > 0)) return B200MD_EINVAL;
  const int P = nranks;
  const int nlower = -(order - 1) / 2, nupper = order / 2;
  const double shift = (order % 2) ? PPPM_OFFSET + 0.5 : PPPM_OFFSET;
  const double dist = 0.5 * skin;
  for (int r = 0; r < P; r++) {
    const double zlo = r * (prd_z / P), zhi = r == P - 1 ? prd_z : (r + 1) * (prd_z / P);
    const int nlo = static_cast<int>((zlo - dist) * nz / prd_z + shift) - PPPM_OFFSET;
    const int nhi = static_cast<int>((zhi + dist) * nz / prd_z + shift) - PPPM_OFFSET;
    pzlo[r] = (int)((long)r * nz / P);
    pzhi[r] = (int)((long)(r + 1) * nz / P);
    const int blo = std::min(nlo + nlower, pzlo[r]), bhi = std::max(nhi + nupper, pzhi[r] - 1);
    zoff[r] = blo;
    nbz[r] = bhi - blo + 1;
    ylo[r] = (int)((long)r * ny / P);
    yhi[r] = (int)((long)(r + 1) * ny / P);
  }
  if (P > 1)
    for (int r = 0; r < P; r++) {
      const int lo_w = pzlo[r] - zoff[r], hi_w = zoff[r] + nbz[r] - pzhi[r];
      const int lower = (r + P - 1) % P, upper = (r + 1) % P;
      if (lo_w > pzhi[lower] - pzlo[lower] || hi_w > pzhi[upper] - pzlo[upper]) return B200MD_EINVAL;
    }
  return 0;
}

int b200md_pppm_setup(b200md_ctx *ctx, const b200md_pppm_params *p) {
  if (!ctx || !p) return b2_fail(ctx, B200MD_EINVAL, "b200md_pppm_setup: NULL argument");
  cudaSetDevice(ctx->device);
  if (!ctx->box_set) return b2_fail(ctx, B200MD_EINVAL, "b200md_pppm_setup before b200md_set_box");
  if (p->order > B2_MAXORDER || p->order < 1)
    return b2_fail(ctx, B200MD_EORDER, "PPPM order greater than supported by USER-INTEL");
  if (p->nx < 2 || p->ny < 2 || p->nz < 2 || p->nx >= PPPM_OFFSET || p->ny >= PPPM_OFFSET || p->nz >= PPPM_OFFSET)
    return b2_fail(ctx, B200MD_EINVAL, "PPPM grid is too large or too small");
  if (p->nx < 2 * p->order || p->ny < 2 * p->order || p->nz < 2 * p->order)
    return b2_fail(ctx, B200MD_EINVAL, "PPPM grid must have at least 2*order points per dimension");
  if (!(p->g_ewald > 0)) return b2_fail(ctx, B200MD_EINVAL, "PPPM needs g_ewald > 0");
  if (p->dispersion && !p->B) return b2_fail(ctx, B200MD_EINVAL, "dispersion PPPM needs B[type]");
  // kspace_modify slab (PPPM::setup [UPSTREAM]): z is non-periodic, the mesh spans zprd * slab_volfactor and
  // slabcorr() (pppm_intel.cpp:305) removes the inter-slab dipole interaction
  const bool slab = p->slab_volfactor > 1.0;
  if (slab) {
    if (!ctx->periodic[0] || !ctx->periodic[1] || ctx->periodic[2])
      return b2_fail(ctx, B200MD_EINVAL, "Incorrect boundaries with slab PPPM");
    if (p->dispersion) return b2_fail(ctx, B200MD_EINVAL, "slab correction is not provided for the dispersion grid");
    if (b2_comm_nranks(ctx) > 1) return b2_fail(ctx, B200MD_EINVAL, "slab PPPM is single-GPU only in this build");
  } else {
    for (int d = 0; d < 3; d++)
      if (!ctx->periodic[d]) return b2_fail(ctx, B200MD_EINVAL, "Cannot use nonperiodic boundaries with PPPM");
  }
  // triclinic boxes (pppm_intel.cpp:151-156, 878-883): stock PPPM::init refuses what it cannot do on them
  const bool tri = ctx->triclinic;
  if (tri) {
    if (p->differentiation == 1)
      return b2_fail(ctx, B200MD_EINVAL, "Cannot (yet) use PPPM with triclinic box and kspace_modify diff ad");
    if (slab) return b2_fail(ctx, B200MD_EINVAL, "Cannot (yet) use PPPM with triclinic box and slab correction");
    if (p->dispersion) return b2_fail(ctx, B200MD_EINVAL, "Cannot (yet) use PPPMDisp with triclinic box");
    if (b2_comm_nranks(ctx) > 1) return b2_fail(ctx, B200MD_EINVAL, "triclinic PPPM is single-GPU only in this build");
  }
  PppmState *&slot = p->dispersion ? ctx->pppm6 : ctx->pppm;
  free_state(ctx, slot);
  PppmState *ps = new PppmState();
  slot = ps;
  ps->p = *p;
  ps->p.B = nullptr;
  if (ps->p.scale == 0.0) ps->p.scale = 1.0;
  PppmConst &c = ps->c;
  c.nx = p->nx; c.ny = p->ny; c.nz = p->nz; c.order = p->order;
  c.nlower = -(p->order - 1) / 2;
  c.nupper = p->order / 2;
  if (p->order % 2) { c.shift = PPPM_OFFSET + 0.5; c.shiftone = 0.0; }
  else { c.shift = PPPM_OFFSET; c.shiftone = 0.5; }
  c.g_ewald = p->g_ewald;
  const int ng[3] = {p->nx, p->ny, p->nz};
  const double dist = 0.5 * ctx->neigh.skin;  // cuthalf (no TIP4P qdist)
  for (int d = 0; d < 3; d++) {
    c.boxlo[d] = ctx->boxlo[d];
    c.prd[d] = ctx->prd[d] * (d == 2 && slab ? p->slab_volfactor : 1.0);   // zprd_slab
    c.delinv[d] = ng[d] / c.prd[d];
    // PPPM::set_grid_local extents, one rank: ghost cells reach dist beyond the box
    const int nlo = static_cast<int>((0.0 - dist) * ng[d] / c.prd[d] + c.shift) - PPPM_OFFSET;
    int nhi = static_cast<int>((ctx->prd[d] + dist) * ng[d] / c.prd[d] + c.shift) - PPPM_OFFSET;
    if (d == 2 && slab) nhi = std::max(nhi, ng[d] - 1);   // set_grid_local: the top rank owns the empty planes too
    c.lo_out[d] = nlo + c.nlower;
    c.hi_out[d] = nhi + c.nupper;
  }
  c.delvolinv = c.delinv[0] * c.delinv[1] * c.delinv[2];
  c.tri = tri ? 1 : 0;
  for (int d = 0; d < 6; d++) c.hinv[d] = 0.0;
  for (int d = 0; d < 3; d++) c.boxlo_box[d] = ctx->boxlo[d];
  double hmat[6] = {ctx->prd[0], ctx->prd[1], ctx->prd[2], ctx->tilt[2], ctx->tilt[1], ctx->tilt[0]};   // Domain::h
  if (tri) {
    // Domain::set_global_box: h_inv; PPPM::setup_triclinic: the mesh spans the unit cube of lamda coordinates;
    // PPPM::set_grid_local: the skin enters in lamda units (KSpace::kspacebbox)
    c.hinv[0] = 1.0 / hmat[0]; c.hinv[1] = 1.0 / hmat[1]; c.hinv[2] = 1.0 / hmat[2];
    c.hinv[3] = -hmat[3] / (hmat[1] * hmat[2]);
    c.hinv[4] = (hmat[3] * hmat[5] - hmat[1] * hmat[4]) / (hmat[0] * hmat[1] * hmat[2]);
    c.hinv[5] = -hmat[5] / (hmat[0] * hmat[1]);
    const double lx = hmat[0], ly = hmat[1], lz = hmat[2], yz = hmat[3], xz = hmat[4], xy = hmat[5];
    const double dl[3] = {dist * std::sqrt(ly * ly * lz * lz + ly * ly * xz * xz - 2.0 * ly * xy * xz * yz +
                                           xy * xy * yz * yz + xy * xy * lz * lz) / (lx * ly * lz),
                          dist * std::sqrt(lz * lz + yz * yz) / (ly * lz), dist / lz};
    for (int d = 0; d < 3; d++) {
      c.boxlo[d] = 0.0;
      c.delinv[d] = ng[d];
      const int nlo = static_cast<int>((0.0 - dl[d]) * ng[d] + c.shift) - PPPM_OFFSET;
      const int nhi = static_cast<int>((1.0 + dl[d]) * ng[d] + c.shift) - PPPM_OFFSET;
      c.lo_out[d] = nlo + c.nlower;
      c.hi_out[d] = nhi + c.nupper;
    }
    c.delvolinv = c.delinv[0] * c.delinv[1] * c.delinv[2] / (c.prd[0] * c.prd[1] * c.prd[2]);
  }
  c.zoff = 0;
  ps->volume = c.prd[0] * c.prd[1] * c.prd[2];   // xprd * yprd * zprd_slab
  ps->zprd = ctx->prd[2];
  ps->gnz = p->nz;
  ps->nranks = b2_comm_nranks(ctx);
  ps->rank = b2_comm_rank(ctx);
  ps->skin_setup = ctx->neigh.skin;
  // half-spectrum transforms: the default (B200MD_R2C=0 keeps the complex-to-complex passes, for comparison runs)
  {
    const char *re = getenv("B200MD_R2C");
    ps->r2c = !(re && re[0] == '0');
    // triclinic: the influence function of a Nyquist-plane point and of its mirror image differ (x2lamdaT of wave
    // numbers that do not change sign), which half a spectrum cannot hold: complex-to-complex passes
    if (tri) ps->r2c = false;
  }
  c.sx = ps->r2c ? p->nx / 2 + 1 : p->nx;
  int gf_yoff = 0, gf_nyl = p->ny;
  if (ps->nranks > 1) {
    // z-slab decomposition: owned planes, local brick (owned + stencil/skin halo) and z-pencil rows of every rank
    const int P = ps->nranks, me = ps->rank;
    ps->pzlo.resize(P); ps->pzhi.resize(P); ps->zoffs.resize(P); ps->nbzs.resize(P); ps->ylos.resize(P); ps->yhis.resize(P);
    if (b200md_pppm_decomp(P, p->nz, p->ny, p->order, ctx->neigh.skin, ctx->prd[2], ps->pzlo.data(), ps->pzhi.data(),
                           ps->zoffs.data(), ps->nbzs.data(), ps->ylos.data(), ps->yhis.data()))
      return b2_fail(ctx, B200MD_EINVAL, "PPPM grid: %d planes over %d GPUs leaves slabs thinner than the stencil halo",
                     p->nz, P);
    c.zoff = ps->zoffs[me];
    c.nz = ps->nbzs[me];
    c.lo_out[2] = 0;            // local frame: the brick IS the allowed range of stencil planes
    c.hi_out[2] = c.nz - 1;
    gf_yoff = ps->ylos[me];
    gf_nyl = ps->yhis[me] - ps->ylos[me];
    // peer-memory transposes: one z-pencil block and one plane block per rank, sized for the largest share and mapped
    // into every rank (collective; falls back to the NCCL all-to-all when peer access is unavailable or B200MD_P2P=0)
    // B200MD_P2P: 0 = NCCL all-to-all (pack kernel, grouped send/recv, unpack kernel), 1 = stores from a scatter
    // kernel, 2 = strided copy-engine copies (default)
    const char *pe = getenv("B200MD_P2P");
    ps->p2p_dma = !(pe && pe[0] == '1');
    if (!(pe && pe[0] == '0')) {
      int maxrows = 0, maxplanes = 0;
      for (int q = 0; q < P; q++) {
        maxrows = std::max(maxrows, ps->yhis[q] - ps->ylos[q]);
        maxplanes = std::max(maxplanes, ps->pzhi[q] - ps->pzlo[q]);
      }
      const int npack = p->differentiation == 1 ? 1 : (ps->r2c ? 3 : 2);
      const size_t bytesT = ((size_t)c.sx * maxrows * p->nz + 16) * sizeof(double2);
      const size_t bytesW = ((size_t)c.sx * p->ny * maxplanes * npack + 16) * sizeof(double2);
      int okT = 0, okW = 0;
      TRY(b2_comm_peer_alloc(ctx, ps->symT, bytesT, &okT));
      TRY(b2_comm_peer_alloc(ctx, ps->symW, bytesW, &okW));
      ps->p2p = okT && okW;
      if (!ps->p2p) {
        b2_comm_peer_free(ctx, ps->symT);
        b2_comm_peer_free(ctx, ps->symW);
        if (me == 0) fprintf(stderr, "b200md: no peer access between the GPUs, FFT transposes use the NCCL all-to-all\n");
      }
    }
  }
  ps->nfft = (long)p->nx * p->ny * c.nz;   // points of the local brick (= the whole grid on one GPU)
  if ((long)p->nx * p->ny * p->nz > 2000000000L) return b2_fail(ctx, B200MD_EINVAL, "PPPM grid has too many points");
  compute_rho_coeffs(c);
  compute_gf_denom(c);
  {  // make_rho fold: covering tiles of every x, y and (local) z coordinate
    std::vector<int4> cov((size_t)c.nx + c.ny + c.nz);
    if (!cover_table(c.nx, cdiv(c.nx, RHO_T), c.nlower, c.nupper, cov.data()) ||
        !cover_table(c.ny, cdiv(c.ny, RHO_T), c.nlower, c.nupper, cov.data() + c.nx) ||
        !cover_table(c.nz, cdiv(c.nz, RHO_T), c.nlower, c.nupper, cov.data() + c.nx + c.ny))
      return b2_fail(ctx, B200MD_EINVAL, "PPPM grid too small for the tiled charge assignment");
    RESERVE(ctx, ps->cover, cov.size());
    CUDA_OK(ctx, cudaMemcpy(ps->cover.p, cov.data(), cov.size() * sizeof(int4), cudaMemcpyHostToDevice));
  }
  for (int d = 0; d < 3; d++) TRY(make_plan(ctx, ps->plan[d], ps->tw[d], ng[d]));
  const long nfft = ps->nfft;
  const bool ad = p->differentiation == 1;
  PppmConst cg = c;             // global-grid view for the Green's function kernels
  cg.nz = p->nz;
  const long ngf = (long)c.sx * gf_nyl * p->nz;   // Green's function points held here ([z][y rows of this rank][sx])
  RESERVE(ctx, ps->greensfn, (size_t)std::max(ngf, 1L));
  RESERVE(ctx, ps->density, (size_t)nfft);
  RESERVE(ctx, ps->work1, (size_t)nfft);
  RESERVE(ctx, ps->work2, (size_t)nfft * (ad ? 1 : 3));
  RESERVE(ctx, ps->vd, (size_t)nfft * (ad ? 1 : 3));
  RESERVE(ctx, ps->fkx, (size_t)p->nx);
  RESERVE(ctx, ps->fky, (size_t)p->ny);
  RESERVE(ctx, ps->fkz, (size_t)p->nz);
  {
    // PPPM::setup: fkx/fky/fkz
    std::vector<double> fk;
    DevBuf<double> *dst[3] = {&ps->fkx, &ps->fky, &ps->fkz};
    for (int d = 0; d < 3; d++) {
      const double unitk = k2PI / c.prd[d];
      fk.assign(ng[d], 0.0);
      for (int i = 0; i < ng[d]; i++) fk[i] = unitk * (i - ng[d] * (2 * i / ng[d]));
      CUDA_OK(ctx, cudaMemcpy(dst[d]->p, fk.data(), ng[d] * sizeof(double), cudaMemcpyHostToDevice));
      {   // gradient copies with the Nyquist entry dropped (see k_fft_z_poisson; z: half-spectrum path only)
        DevBuf<double> &gdst = d == 0 ? ps->fkx_g : (d == 1 ? ps->fky_g : ps->fkz_g);
        RESERVE(ctx, gdst, (size_t)ng[d]);
        if (ng[d] % 2 == 0) fk[ng[d] / 2] = 0.0;
        CUDA_OK(ctx, cudaMemcpy(gdst.p, fk.data(), ng[d] * sizeof(double), cudaMemcpyHostToDevice));
      }
    }
    if (tri) {
      // PPPM::setup_triclinic: k = x2lamdaT(2 pi per_i, 2 pi per_j, 2 pi per_k).  The diagonal of h_inv is 1 / prd, so
      // fkx / fky / fkz above are its diagonal part; the off-diagonal part is three more one-dimensional arrays
      auto per = [](int i, int n) { return (double)(i - n * (2 * i / n)); };
      std::vector<double> yx(p->nx), zx(p->nx), zy(p->ny), yxg(p->nx);
      for (int i = 0; i < p->nx; i++) {
        yx[i] = c.hinv[5] * k2PI * per(i, p->nx);
        zx[i] = c.hinv[4] * k2PI * per(i, p->nx);
        yxg[i] = (p->nx % 2 == 0 && i == p->nx / 2) ? 0.0 : yx[i];
      }
      for (int j = 0; j < p->ny; j++) zy[j] = c.hinv[3] * k2PI * per(j, p->ny);
      RESERVE(ctx, ps->fkyx, yx.size()); RESERVE(ctx, ps->fkzx, zx.size());
      RESERVE(ctx, ps->fkzy, zy.size()); RESERVE(ctx, ps->fkyx_g, yxg.size());

      CUDA_OK(ctx, cudaMemcpy(ps->fkyx.p, yx.data(), yx.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(ctx, cudaMemcpy(ps->fkzx.p, zx.data(), zx.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(ctx, cudaMemcpy(ps->fkzy.p, zy.data(), zy.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(ctx, cudaMemcpy(ps->fkyx_g.p, yxg.data(), yxg.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  if (p->dispersion) {
    std::vector<double> W, sg;
    if (disp_components(p->dispersion, ctx->ntypes, p->B, W, sg))
      return b2_fail(ctx, B200MD_EINVAL, "dispersion PPPM: mixing must be 1 (geometric), 2 (arithmetic) or 3 (none) "
                                         "with a symmetric coefficient matrix");
    ps->ncomp = (int)sg.size();
    for (int m = 0; m < ps->ncomp; m++) ps->comp_sign[m] = sg[m];
    RESERVE(ctx, ps->Btype, W.size());
    CUDA_OK(ctx, cudaMemcpy(ps->Btype.p, W.data(), W.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (ngf == 0) {
  } else if (tri) {
    // compute_gf_ik_triclinic: alias counts from lamda2xT(g_ewald / (pi n) * EPS_HOC-factor)
    const double fac = std::pow(-std::log(1.0e-7), 0.25);
    const double t0 = (cg.g_ewald / (kPI * cg.nx)) * fac, t1 = (cg.g_ewald / (kPI * cg.ny)) * fac,
                 t2 = (cg.g_ewald / (kPI * cg.nz)) * fac;
    const int nbx = static_cast<int>(hmat[0] * t0);
    const int nby = static_cast<int>(hmat[5] * t0 + hmat[1] * t1);
    const int nbz = static_cast<int>(hmat[4] * t0 + hmat[3] * t1 + hmat[2] * t2);
    k_gf_ik_tri<<<cdiv(ngf, 128), 128, 0, ctx->stream>>>(cg, nbx, nby, nbz, ps->greensfn.p);
    KERNEL_OK(ctx, "k_gf_ik_tri");
  } else if (p->dispersion && !ad) {
    k_gf_6<<<cdiv(ngf, 128), 128, 0, ctx->stream>>>(cg, gf_yoff, gf_nyl, ps->greensfn.p);
    KERNEL_OK(ctx, "k_gf_6");
  } else if (!ad) {
    const double fac = std::pow(-std::log(1.0e-7), 0.25);  // EPS_HOC, pppm_intel.cpp:39
    const int nbx = static_cast<int>((cg.g_ewald * cg.prd[0] / (kPI * cg.nx)) * fac);
    const int nby = static_cast<int>((cg.g_ewald * cg.prd[1] / (kPI * cg.ny)) * fac);
    const int nbz = static_cast<int>((cg.g_ewald * cg.prd[2] / (kPI * cg.nz)) * fac);
    k_gf_ik<<<cdiv(ngf, 128), 128, 0, ctx->stream>>>(cg, nbx, nby, nbz, gf_yoff, gf_nyl, ps->greensfn.p);
    KERNEL_OK(ctx, "k_gf_ik");
  } else {
    RESERVE(ctx, ps->sf_pre, 6 * (size_t)ngf);
    if (p->dispersion) {   // PPPMDisp::compute_sf_coeff_6: the sums run over the r^-6 influence function
      k_gf_6<<<cdiv(ngf, 128), 128, 0, ctx->stream>>>(cg, gf_yoff, gf_nyl, ps->greensfn.p);
      KERNEL_OK(ctx, "k_gf_6");
    }
    k_gf_ad<<<cdiv(ngf, 128), 128, 0, ctx->stream>>>(cg, gf_yoff, gf_nyl, ps->greensfn.p, ps->sf_pre.p,
                                                      p->dispersion ? 1 : 0);
    KERNEL_OK(ctx, "k_gf_ad");
    double s[6];
    TRY(reduce_cols(ctx, *ps, ngf, 6, ps->sf_pre.p, s));
    if (ps->nranks > 1) {   // the sums run over the whole grid (MPI_Allreduce in compute_sf_precoeff)
      RESERVE(ctx, ps->red, 16);
      CUDA_OK(ctx, cudaMemcpyAsync(ps->red.p, s, 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      TRY(b2_comm_allreduce_sum(ctx, ps->red.p, 6));
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ps->red.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
      for (int k = 0; k < 6; k++) s[k] = ctx->h_pinned[k];
    }
    // compute_gf_ad: self-force coefficients
    double prex = kPI / ps->volume, prey = prex, prez = prex;
    prex *= cg.nx / cg.prd[0];
    prey *= cg.ny / cg.prd[1];
    prez *= cg.nz / cg.prd[2];
    ps->sf_coeff[0] = s[0] * prex; ps->sf_coeff[1] = s[1] * prex * 2;
    ps->sf_coeff[2] = s[2] * prey; ps->sf_coeff[3] = s[3] * prey * 2;
    ps->sf_coeff[4] = s[4] * prez; ps->sf_coeff[5] = s[5] * prez * 2;
    ps->sf_pre.free_();
  }
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  (void)k4PI;
  return 0;
}

int b200md_debug_disp_components(int mix, int ntypes, const double *B, double *W, double *sign) {
  if (!B || !W || !sign || ntypes < 1) return B200MD_EINVAL;
  std::vector<double> w, s;
  if (disp_components(mix, ntypes, B, w, s)) return B200MD_EINVAL;
  std::copy(w.begin(), w.end(), W);
  std::copy(s.begin(), s.end(), sign);
  return (int)s.size();
}

// host-only helpers of the tiled charge assignment, exposed for the CPU tests (no device needed)
int b200md_debug_rho_plan(int order, int n, int *pitch, int *lane_point /*[64]*/, int *cover /*[n][4]*/) {
  if (order < 1 || order > B2_MAXORDER || n < 1 || !pitch || !lane_point || !cover) return B200MD_EINVAL;
  TileGeom tg{};
  rho_lane_map(order, tg);
  *pitch = rho_pitch(order);
  for (int k = 0; k < 64; k++) lane_point[k] = tg.lane_pt[k];
  std::vector<int4> cov(n);
  if (!cover_table(n, cdiv(n, RHO_T), -(order - 1) / 2, order / 2, cov.data())) return B200MD_EINVAL;
  for (int g = 0; g < n; g++) {
    cover[4 * g] = cov[g].x; cover[4 * g + 1] = cov[g].y; cover[4 * g + 2] = cov[g].z; cover[4 * g + 3] = cov[g].w;
  }
  return 0;
}

int b200md_pppm_peratom(b200md_ctx *ctx, double *eatom, double *vatom) {
  if (!ctx || !ctx->pppm) return b2_fail(ctx, B200MD_EINVAL, "b200md_pppm_peratom before b200md_pppm_setup");
  cudaSetDevice(ctx->device);
  PppmState &ps = *ctx->pppm;
  const int n = ctx->nlocal;
  if ((eatom && !ps.pa_have_e) || (vatom && !ps.pa_have_v) || ps.pa_n_atoms != n)
    return b2_fail(ctx, B200MD_EINVAL, "no per-atom tallies: the last b200md_pppm_compute did not ask for them");
  if (n == 0) return 0;
  std::vector<double> out(7 * (size_t)n);
  std::vector<int> tag(n);
  CUDA_OK(ctx, cudaMemcpyAsync(out.data(), ps.pa_out.p, out.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(tag.data(), ctx->tag.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; i++) {   // resident order -> upload order
    const size_t t = (size_t)(tag[i] - ctx->first_id);
    if (eatom) eatom[t] = out[i];
    if (vatom)
      for (int k = 0; k < 6; k++) vatom[6 * t + k] = out[(size_t)(k + 1) * n + i];
  }
  return 0;
}

int b200md_pppm_compute(b200md_ctx *ctx, int eflag, int vflag, double *energy, double virial[6]) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  TRY(b2_pppm_compute(ctx, eflag, vflag, energy, virial));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

namespace {
__global__ void k_pack_xq(int n, const double *__restrict__ x, const double *__restrict__ q, double4 *__restrict__ xq,
                          float4 *__restrict__ xqf, double4 *__restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = make_double4(x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2], q[i]);
  xq[i] = p;
  if (xqf) xqf[i] = make_float4((float)p.x, (float)p.y, (float)p.z, (float)p.w);
  f[i] = make_double4(0, 0, 0, 0);
}
}  // namespace

int b200md_pppm_compute_host(b200md_ctx *ctx, int eflag, int vflag, int n, const double *x, const double *q,
                             double *f, double *energy, double virial[6]) {
  if (!ctx || !x || !q || !f || n < 0) return b2_fail(ctx, B200MD_EINVAL, "b200md_pppm_compute_host: bad arguments");
  if (!ctx->pppm) return b2_fail(ctx, B200MD_EINVAL, "pppm compute before b200md_pppm_setup");
  if (ctx->pppm->nranks > 1) return b2_fail(ctx, B200MD_EINVAL, "b200md_pppm_compute_host is single-GPU only");
  if (ctx->pppm->p.dispersion) return b2_fail(ctx, B200MD_EINVAL, "host form supports the Coulomb grid only");
  cudaSetDevice(ctx->device);
  DevBuf<double> dx, dq;
  DevBuf<double4> dxq, df;
  DevBuf<float4> dxqf;
  const bool mixed = ctx->prec == B200MD_PREC_MIXED;
  auto cleanup = [&]() { dx.free_(); dq.free_(); dxq.free_(); df.free_(); dxqf.free_(); };
  if (dx.reserve(3 * (size_t)n + 1) || dq.reserve((size_t)n + 1) || dxq.reserve((size_t)n + 1) || df.reserve((size_t)n + 1) ||
      (mixed && dxqf.reserve((size_t)n + 1))) {
    cleanup();
    return b2_fail(ctx, B200MD_ENOMEM, "out of device memory in b200md_pppm_compute_host");
  }
  cudaMemcpyAsync(dx.p, x, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dq.p, q, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (n) {
    k_pack_xq<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, dx.p, dq.p, dxq.p, mixed ? dxqf.p : nullptr, df.p);
    ctx->launches++;
  }
  PppmView v{n, dxq.p, mixed ? dxqf.p : nullptr, nullptr, df.p};
  ctx->pppm->q_natoms = -1;  // a different atom set: recompute qsum/qsqsum
  int rc = mixed ? pppm_compute_view<float>(ctx, *ctx->pppm, v, eflag, vflag, energy, virial)
                 : pppm_compute_view<double>(ctx, *ctx->pppm, v, eflag, vflag, energy, virial);
  ctx->pppm->q_natoms = -1;
  if (!rc) {
    std::vector<double> hf(4 * (size_t)n);
    cudaMemcpyAsync(hf.data(), df.p, (size_t)n * sizeof(double4), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = b2_fail(ctx, B200MD_ECUDA, "pppm host compute failed: %s", cudaGetErrorString(e));
    else
      for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) f[3 * (size_t)i + d] += hf[4 * (size_t)i + d];
  }
  cleanup();
  return rc;
}

int b200md_pppm_download(b200md_ctx *ctx, double *density_fft, double *greensfn, double *field_x, double *field_y,
                         double *field_z, double sf_coeff[6]) {
  if (!ctx || !ctx->pppm) return b2_fail(ctx, B200MD_EINVAL, "pppm not set up");
  cudaSetDevice(ctx->device);
  PppmState &ps = *ctx->pppm;
  const size_t nb = (size_t)ps.nfft * sizeof(double);
  const bool ad = ps.p.differentiation == 1;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (density_fft) CUDA_OK(ctx, cudaMemcpy(density_fft, ps.density.p, nb, cudaMemcpyDeviceToHost));
  if (greensfn) {
    const PppmConst &c = ps.c;
    if (c.sx == c.nx) CUDA_OK(ctx, cudaMemcpy(greensfn, ps.greensfn.p, nb, cudaMemcpyDeviceToHost));
    else {   // half spectrum [nz][ny][sx] -> the full array: G(-k) = G(k)
      std::vector<double> h((size_t)c.sx * c.ny * c.nz);
      CUDA_OK(ctx, cudaMemcpy(h.data(), ps.greensfn.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (int k = 0; k < c.nz; k++)
        for (int j = 0; j < c.ny; j++)
          for (int i = 0; i < c.nx; i++) {
            const bool mir = i >= c.sx;
            const int ii = mir ? c.nx - i : i, jj = mir ? (c.ny - j) % c.ny : j, kk = mir ? (c.nz - k) % c.nz : k;
            greensfn[((size_t)k * c.ny + j) * c.nx + i] = h[((size_t)kk * c.ny + jj) * c.sx + ii];
          }
    }
  }
  if (field_x) CUDA_OK(ctx, cudaMemcpy(field_x, ps.vd.p, nb, cudaMemcpyDeviceToHost));
  if (field_y) CUDA_OK(ctx, cudaMemcpy(field_y, ps.vd.p + (ad ? 0 : ps.nfft), nb, cudaMemcpyDeviceToHost));
  if (field_z) CUDA_OK(ctx, cudaMemcpy(field_z, ps.vd.p + (ad ? 0 : 2 * ps.nfft), nb, cudaMemcpyDeviceToHost));
  if (sf_coeff) for (int k = 0; k < 6; k++) sf_coeff[k] = ps.sf_coeff[k];
  return 0;
}

int b200md_fft3d_host(b200md_ctx *ctx, double *data, int nx, int ny, int nz, int dir) {
  if (!ctx || !data || nx < 1 || ny < 1 || nz < 1) return b2_fail(ctx, B200MD_EINVAL, "b200md_fft3d_host: bad arguments");
  cudaSetDevice(ctx->device);
  FftPlan1d pl[3];
  DevBuf<double2> tw[3], buf;
  const int ng[3] = {nx, ny, nz};
  const size_t nfft = (size_t)nx * ny * nz;
  int rc = 0;
  for (int d = 0; d < 3 && !rc; d++) rc = make_plan(ctx, pl[d], tw[d], ng[d]);
  if (!rc && buf.reserve(nfft)) rc = b2_fail(ctx, B200MD_ENOMEM, "out of device memory in b200md_fft3d_host");
  if (!rc) {
    cudaMemcpyAsync(buf.p, data, nfft * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream);
    const double s = dir > 0 ? S_FWD : S_BWD;
    PassGeom gx{(long)ny * nz, 1, (long)nx, 1, 0};
    PassGeom gy{(long)nx * nz, nx, (long)nx * ny, (long)nx, 0};
    PassGeom gz{(long)nx * ny, nx * ny, 0, (long)nx * ny, 0};
    rc = launch_pass<1, 0, 0>(ctx, pl[0], gx, nullptr, buf.p, buf.p, nullptr, s);
    if (!rc) rc = launch_pass<0, 0, 0>(ctx, pl[1], gy, nullptr, buf.p, buf.p, nullptr, s);
    if (!rc) rc = launch_pass<0, 0, 0>(ctx, pl[2], gz, nullptr, buf.p, buf.p, nullptr, s);
    if (!rc) {
      cudaMemcpyAsync(data, buf.p, nfft * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream);
      cudaError_t e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) rc = b2_fail(ctx, B200MD_ECUDA, "fft3d failed: %s", cudaGetErrorString(e));
    }
  }
  for (int d = 0; d < 3; d++) tw[d].free_();
  buf.free_();
  return rc;
}

}  // extern "C"
