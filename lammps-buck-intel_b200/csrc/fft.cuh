// fft.cuh — hand-written mixed-radix (2,3,4,5) Stockham FFT on lines staged in shared memory.
//
// Replaces stock LAMMPS FFT3d / KISS 1-D transforms called at pppm_intel.cpp:835 (fft1->compute(+1)) and
// :903,930,958,1045 (fft2->compute(-1)); PPPM grids are 2^a 3^b 5^c.  Double complex,
// unnormalised, forward = exp(-i k x).
//
// One block transforms TB lines at once.  The lines sit in shared memory as [line][k] with an odd padded
// pitch, so both the butterfly accesses (consecutive k) and the strided-dimension loads/stores (consecutive
// lines = consecutive x in global memory, i.e. coalesced 16 B accesses) are bank-conflict free.
// Stockham autosort: out-of-place ping-pong between two shared buffers, no bit reversal, natural order out.
#pragma once
#include <cuda_runtime.h>

struct FftPlan1d {
  int n = 0;
  int nfac = 0;
  int fac[24];
  // per stage (host-filled): p = product of the previous radices, m = n / radix, tstep = n / (p * radix), and
  // 1/p as a float for the division-free k = j mod p (exact for j < 2^12, see stockham_stage)
  int sp[24], sm[24], ststep[24];
  float sinvp[24];
  double2 *tw = nullptr;  // device: exp(-2 pi i m / n), m < n
};

__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(const double2 a, const double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(const double2 a, const double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i*s (s = +1 forward, -1 backward):  -i*(x+iy) = y - ix
__device__ __forceinline__ double2 cmul_mi(const double2 a, const double s) { return make_double2(s * a.y, -s * a.x); }

template <int R>
__device__ __forceinline__ void dft_small(double2 *u, const double s);

template <>
__device__ __forceinline__ void dft_small<2>(double2 *u, const double) {
  const double2 a = u[0], b = u[1];
  u[0] = cadd(a, b);
  u[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft_small<3>(double2 *u, const double s) {
  const double c = -0.5, sn = 0.86602540378443864676;
  const double2 t1 = cadd(u[1], u[2]);
  const double2 t2 = make_double2(u[0].x + c * t1.x, u[0].y + c * t1.y);
  const double2 d = csub(u[1], u[2]);
  const double2 t3 = cmul_mi(make_double2(sn * d.x, sn * d.y), s);  // -i s sin(60) (u1-u2)
  u[0] = cadd(u[0], t1);
  u[1] = cadd(t2, t3);
  u[2] = csub(t2, t3);
}
template <>
__device__ __forceinline__ void dft_small<4>(double2 *u, const double s) {
  const double2 a = cadd(u[0], u[2]), b = csub(u[0], u[2]);
  const double2 c = cadd(u[1], u[3]), d = cmul_mi(csub(u[1], u[3]), s);
  u[0] = cadd(a, c);
  u[2] = csub(a, c);
  u[1] = cadd(b, d);
  u[3] = csub(b, d);
}
template <>
__device__ __forceinline__ void dft_small<5>(double2 *u, const double s) {
  const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
  const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
  const double2 a1 = cadd(u[1], u[4]), a2 = cadd(u[2], u[3]);
  const double2 b1 = csub(u[1], u[4]), b2 = csub(u[2], u[3]);
  const double2 p1 = make_double2(u[0].x + c1 * a1.x + c2 * a2.x, u[0].y + c1 * a1.y + c2 * a2.y);
  const double2 p2 = make_double2(u[0].x + c2 * a1.x + c1 * a2.x, u[0].y + c2 * a1.y + c1 * a2.y);
  const double2 q1 = cmul_mi(make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y), s);
  const double2 q2 = cmul_mi(make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y), s);
  u[0] = make_double2(u[0].x + a1.x + a2.x, u[0].y + a1.y + a2.y);
  u[1] = cadd(p1, q1);
  u[4] = csub(p1, q1);
  u[2] = cadd(p2, q2);
  u[3] = csub(p2, q2);
}

// one Stockham stage of radix R over TB lines: src -> dst (both [TB][LP]).  Threads are laid out as
// (line t = tid >> lgtpl, butterfly lane j0 = tid & (tpl - 1)) with tpl = threads per line a power of two, so the
// only index arithmetic per butterfly is k = j mod p, done with the precomputed float reciprocal of p.
template <int R>
__device__ __forceinline__ void stockham_stage(const double2 *__restrict__ src, double2 *__restrict__ dst, const int LP,
                                               const int TB, const int lgtpl, const int p, const int m,
                                               const int tstep, const float invp, const double2 *__restrict__ tw,
                                               const double s) {
  const int tpl = 1 << lgtpl;
  const int j0 = threadIdx.x & (tpl - 1);
  const int tstride = blockDim.x >> lgtpl;
  for (int t = threadIdx.x >> lgtpl; t < TB; t += tstride) {
    const double2 *ls = src + t * LP;
    double2 *ldb = dst + t * LP;
    for (int j = j0; j < m; j += tpl) {
      const int q = __float2int_rd(((float)j + 0.5f) * invp);
      const int k = j - q * p;
      double2 u[R];
#pragma unroll
      for (int r = 0; r < R; r++) u[r] = ls[j + r * m];
      if (k) {
#pragma unroll
        for (int r = 1; r < R; r++) {
          double2 w = tw[r * k * tstep];
          w.y *= s;
          u[r] = cmul(u[r], w);
        }
      }
      dft_small<R>(u, s);
      double2 *ld = ldb + (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; r++) ld[r * p] = u[r];
    }
  }
}

// transforms the TB lines held in bufA in place-or-pong; returns the buffer holding the result
// tw_s: the plan's twiddle table staged in shared memory by the caller (stage_twiddles)
__device__ __forceinline__ void stage_twiddles(const FftPlan1d &pl, double2 *tw_s) {
  for (int k = threadIdx.x; k < pl.n; k += blockDim.x) tw_s[k] = pl.tw[k];
}
__device__ __forceinline__ double2 *block_fft(double2 *bufA, double2 *bufB, const FftPlan1d &pl, const int LP,
                                              const int TB, const int lgtpl, const double s, const double2 *tw_s) {
  double2 *src = bufA, *dst = bufB;
  for (int f = 0; f < pl.nfac; f++) {
    const int R = pl.fac[f];
    switch (R) {
      case 2: stockham_stage<2>(src, dst, LP, TB, lgtpl, pl.sp[f], pl.sm[f], pl.ststep[f], pl.sinvp[f], tw_s, s); break;
      case 3: stockham_stage<3>(src, dst, LP, TB, lgtpl, pl.sp[f], pl.sm[f], pl.ststep[f], pl.sinvp[f], tw_s, s); break;
      case 4: stockham_stage<4>(src, dst, LP, TB, lgtpl, pl.sp[f], pl.sm[f], pl.ststep[f], pl.sinvp[f], tw_s, s); break;
      default: stockham_stage<5>(src, dst, LP, TB, lgtpl, pl.sp[f], pl.sm[f], pl.ststep[f], pl.sinvp[f], tw_s, s); break;
    }
    __syncthreads();
    double2 *tmp = src; src = dst; dst = tmp;
  }
  return src;
}
