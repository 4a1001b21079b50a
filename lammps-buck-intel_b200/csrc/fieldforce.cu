// fieldforce.cu — PPPMIntel::fieldforce_ik<> (pppm_intel.cpp:541-640) and fieldforce_ad<> (:679-804).
// One thread per atom, atoms visited in grid-cell order so neighbouring threads read neighbouring grid
// points; indices are wrapped periodically, which fuses the ghost-cell fill (cg->forward_comm, :219-220) away.
#include "pppm_internal.h"

namespace {

// ---------------------------------------------------------------------------------------------
// fieldforce: one thread per (cell-sorted) atom

template <class flt_t, int ORDER, int AD>
__global__ void __launch_bounds__(128)
k_fieldforce(int n, PppmConst c, const double4 *__restrict__ pa_x, const int4 *__restrict__ pa_n,
             const double4 *__restrict__ xq, const float4 *__restrict__ xqf, const int *__restrict__ type,
             const double *__restrict__ Btype, const double *__restrict__ vd, double qqrd2e_scale,
             double sf0, double sf1, double sf2, double sf3, double sf4, double sf5, int gnz,
             double4 *__restrict__ f) {   // gnz: planes of the whole grid (c.nz is the local brick on several GPUs)
  __shared__ double s_rc[B2_MAXORDER * B2_MAXORDER], s_drc[B2_MAXORDER * B2_MAXORDER];
  for (int k = threadIdx.x; k < ORDER * ORDER; k += blockDim.x) {
    s_rc[k] = c.rho_coeff[k];
    s_drc[k] = c.drho_coeff[k];
  }
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const double4 p = pa_x[a];
  const int4 pn = pa_n[a];
  const long nfft = (long)c.nx * c.ny * c.nz;
  flt_t rho[3][ORDER], drho[3][ORDER];
  int ix[ORDER], iy[ORDER], iz[ORDER];
#pragma unroll
  for (int k = 0; k < ORDER; k++) {
    double r1 = s_rc[(ORDER - 1) * ORDER + k], r2 = r1, r3 = r1;
    double d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
    for (int l = ORDER - 2; l >= 0; l--) {
      r1 = s_rc[l * ORDER + k] + r1 * p.x;
      r2 = s_rc[l * ORDER + k] + r2 * p.y;
      r3 = s_rc[l * ORDER + k] + r3 * p.z;
      if (AD) {
        d1 = s_drc[l * ORDER + k] + d1 * p.x;
        d2 = s_drc[l * ORDER + k] + d2 * p.y;
        d3 = s_drc[l * ORDER + k] + d3 * p.z;
      }
    }
    rho[0][k] = (flt_t)r1; rho[1][k] = (flt_t)r2; rho[2][k] = (flt_t)r3;
    if (AD) { drho[0][k] = (flt_t)d1; drho[1][k] = (flt_t)d2; drho[2][k] = (flt_t)d3; }
    ix[k] = wrapi(pn.x + c.nlower + k, c.nx);
    iy[k] = wrapi(pn.y + c.nlower + k, c.ny);
    iz[k] = wrapi(pn.z + c.nlower + k, c.nz);
  }
  double ekx = 0.0, eky = 0.0, ekz = 0.0;
#pragma unroll
  for (int nn = 0; nn < ORDER; nn++) {
#pragma unroll
    for (int mm = 0; mm < ORDER; mm++) {
      const long row = ((long)iz[nn] * c.ny + iy[mm]) * c.nx;
      if (AD) {
        const double ekx_p = (double)rho[1][mm] * (double)rho[2][nn];
        const double eky_p = (double)drho[1][mm] * (double)rho[2][nn];
        const double ekz_p = (double)rho[1][mm] * (double)drho[2][nn];
#pragma unroll
        for (int ll = 0; ll < ORDER; ll++) {
          const double u = vd[row + ix[ll]];
          ekx += (double)drho[0][ll] * ekx_p * u;
          eky += (double)rho[0][ll] * eky_p * u;
          ekz += (double)rho[0][ll] * ekz_p * u;
        }
      } else {
        const double y0 = (double)rho[2][nn] * (double)rho[1][mm];
#pragma unroll
        for (int ll = 0; ll < ORDER; ll++) {
          const double x0 = y0 * (double)rho[0][ll];
          const long g = row + ix[ll];
          ekx -= x0 * vd[g];
          eky -= x0 * vd[nfft + g];
          ekz -= x0 * vd[2 * nfft + g];
        }
      }
    }
  }
  const int i = pn.w;
  flt_t qi, px, py, pz;
  if (sizeof(flt_t) == 4) {
    const float4 xi = xqf[i];
    qi = Btype ? (flt_t)Btype[type[i]] : (flt_t)xi.w;
    px = xi.x; py = xi.y; pz = xi.z;
  } else {
    const double4 xi = xq[i];
    qi = Btype ? (flt_t)Btype[type[i]] : (flt_t)xi.w;
    px = (flt_t)xi.x; py = (flt_t)xi.y; pz = (flt_t)xi.z;
  }
  const flt_t fq = (flt_t)qqrd2e_scale;
  const flt_t qfactor = fq * qi;
  double4 fi = f[i];
  if (AD) {
    const flt_t hx_inv = (flt_t)(c.nx / c.prd[0]), hy_inv = (flt_t)(c.ny / c.prd[1]), hz_inv = (flt_t)(gnz / c.prd[2]);
    ekx *= hx_inv; eky *= hy_inv; ekz *= hz_inv;
    const flt_t ftwo_pi = (flt_t)(kPI * 2.0), ffour_pi = (flt_t)(kPI * 4.0);
    const flt_t twoqsq = (flt_t)2.0 * qi * qi;
    const flt_t s1 = px * hx_inv, s2 = py * hy_inv, s3 = pz * hz_inv;
    flt_t sf = (flt_t)sf0 * sin(ftwo_pi * s1);
    sf += (flt_t)sf1 * sin(ffour_pi * s1);
    sf *= twoqsq;
    fi.x += qfactor * ekx - fq * sf;
    sf = (flt_t)sf2 * sin(ftwo_pi * s2);
    sf += (flt_t)sf3 * sin(ffour_pi * s2);
    sf *= twoqsq;
    fi.y += qfactor * eky - fq * sf;
    sf = (flt_t)sf4 * sin(ftwo_pi * s3);
    sf += (flt_t)sf5 * sin(ffour_pi * s3);
    sf *= twoqsq;
    fi.z += qfactor * ekz - fq * sf;
  } else {
    fi.x += qfactor * ekx;
    fi.y += qfactor * eky;
    fi.z += qfactor * ekz;
  }
  f[i] = fi;
}

}  // namespace

template <class flt_t>
int b2_fieldforce(b200md_ctx *ctx, PppmState &ps, const PppmView &v) {
  ScopedTimer tm(ctx, T_FIELDFORCE);
  const PppmConst &c = ps.c;
  const int n = v.n;
  const bool ad = ps.p.differentiation == 1;
  // dispersion: f += sign_m * W_m[type] * E_m for the component m of this pass
  const double qs = ps.p.dispersion ? ps.comp_sign[ps.cur] : ctx->qqrd2e * ps.p.scale;
  const double *B = ps.p.dispersion ? ps.Btype.p + (size_t)ps.cur * (ctx->ntypes + 1) : nullptr;
  const double *sf = ps.sf_coeff;
  if (n <= 0) return 0;
  const int nb = cdiv(n, 128);
#define FF(O)                                                                                                     \
  case O:                                                                                                         \
    if (ad) k_fieldforce<flt_t, O, 1><<<nb, 128, 0, ctx->stream>>>(n, c, ps.pa_x.p, ps.pa_n.p, v.xq, v.xqf, v.type, B, \
                                                                   ps.vd.p, qs, sf[0], sf[1], sf[2], sf[3], sf[4],  \
                                                                   sf[5], ps.gnz, v.f);                                     \
    else k_fieldforce<flt_t, O, 0><<<nb, 128, 0, ctx->stream>>>(n, c, ps.pa_x.p, ps.pa_n.p, v.xq, v.xqf, v.type, B,  \
                                                                ps.vd.p, qs, 0, 0, 0, 0, 0, 0, ps.gnz, v.f);                \
    break;
  switch (c.order) {
    FF(1) FF(2) FF(3) FF(4) FF(5) FF(6) FF(7)
    default: return b2_fail(ctx, B200MD_EORDER, "PPPM order greater than supported by USER-INTEL");
  }
#undef FF
  KERNEL_OK(ctx, "k_fieldforce");
  return 0;
}

// stock PPPM::fieldforce_peratom [UPSTREAM] (the reference calls it through the base class, pppm_intel.cpp:224-229) with
// the per-atom post-factors of PPPM::compute folded in: one thread per atom interpolates the potential brick and the
// six virial bricks (fields = [7][nfft]: u, v0..v5) with the stock double-precision weights.
//   eatom = (1/2 q u - g q^2 / sqrt(pi) - pi/2 q qsum / (g^2 V)) qscale,   vatom_c = 1/2 qscale q v_c
// out = [7][n] in the resident atom order.
template <int ORDER>
__global__ void __launch_bounds__(128)
k_fieldforce_peratom(int n, PppmConst c, const double4 *__restrict__ pa_x, const int4 *__restrict__ pa_n,
                     const double4 *__restrict__ xq, const double *__restrict__ fields, int do_e, int do_v,
                     double qscale, double self_a, double self_b, double *__restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const double4 p = pa_x[a];
  const int4 pn = pa_n[a];
  const long nfft = (long)c.nx * c.ny * c.nz;
  double rho[3][ORDER];
  int ix[ORDER], iy[ORDER], iz[ORDER];
#pragma unroll
  for (int k = 0; k < ORDER; k++) {
    double r1 = 0.0, r2 = 0.0, r3 = 0.0;   // compute_rho1d
#pragma unroll
    for (int l = ORDER - 1; l >= 0; l--) {
      r1 = c.rho_coeff[l * ORDER + k] + r1 * p.x;
      r2 = c.rho_coeff[l * ORDER + k] + r2 * p.y;
      r3 = c.rho_coeff[l * ORDER + k] + r3 * p.z;
    }
    rho[0][k] = r1; rho[1][k] = r2; rho[2][k] = r3;
    ix[k] = wrapi(pn.x + c.nlower + k, c.nx);
    iy[k] = wrapi(pn.y + c.nlower + k, c.ny);
    iz[k] = wrapi(pn.z + c.nlower + k, c.nz);
  }
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
  for (int nn = 0; nn < ORDER; nn++)
#pragma unroll 1
    for (int mm = 0; mm < ORDER; mm++) {
      const long row = ((long)iz[nn] * c.ny + iy[mm]) * c.nx;
      const double y0 = rho[2][nn] * rho[1][mm];
#pragma unroll
      for (int ll = 0; ll < ORDER; ll++) {
        const double x0 = y0 * rho[0][ll];
        const long g = row + ix[ll];
        if (do_e) acc[0] += x0 * fields[g];
        if (do_v) {
#pragma unroll
          for (int t = 1; t < 7; t++) acc[t] += x0 * fields[t * nfft + g];
        }
      }
    }
  const int i = pn.w;
  const double q = xq[i].w;
  if (do_e) out[i] = (0.5 * (q * acc[0]) - (self_a * q * q + self_b * q)) * qscale;
  if (do_v)
    for (int t = 1; t < 7; t++) out[(size_t)t * n + i] = 0.5 * qscale * (q * acc[t]);
}

int b2_fieldforce_peratom(b200md_ctx *ctx, PppmState &ps, const PppmView &v, int do_e, int do_v) {
  const PppmConst &c = ps.c;
  const int n = v.n;
  if (n <= 0) return 0;
  RESERVE(ctx, ps.pa_out, 7 * (size_t)n);
  const double qscale = ctx->qqrd2e * ps.p.scale;
  const double self_a = c.g_ewald / kPIS;                                              // g q^2 / sqrt(pi)
  const double self_b = kPI2 * ps.qsum / (c.g_ewald * c.g_ewald * ps.volume);          // pi/2 q qsum / (g^2 V)
  const int nb = cdiv(n, 128);
#define FP(O)                                                                                                   \
  case O:                                                                                                       \
    k_fieldforce_peratom<O><<<nb, 128, 0, ctx->stream>>>(n, c, ps.pa_x.p, ps.pa_n.p, v.xq, ps.pa_fields.p, do_e, do_v, \
                                                          qscale, self_a, self_b, ps.pa_out.p);                  \
    break;
  switch (c.order) {
    FP(1) FP(2) FP(3) FP(4) FP(5) FP(6) FP(7)
    default: return b2_fail(ctx, B200MD_EORDER, "PPPM order greater than supported by USER-INTEL");
  }
#undef FP
  KERNEL_OK(ctx, "k_fieldforce_peratom");
  ps.pa_n_atoms = n;
  ps.pa_have_e = do_e != 0;
  ps.pa_have_v = do_v != 0;
  return 0;
}

template int b2_fieldforce<double>(b200md_ctx *, PppmState &, const PppmView &);
template int b2_fieldforce<float>(b200md_ctx *, PppmState &, const PppmView &);
