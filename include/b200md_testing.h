/* b200md_testing.h — exports of libb200md.so that exist for the test suite and bench.py only.  They are NOT part of the
 * drop-in boundary (include/b200md.h): no reference interface corresponds to them. */
#ifndef B200MD_TESTING_H
#define B200MD_TESTING_H
#include "b200md.h"
#ifdef __cplusplus
extern "C" {
#endif

/* host-only plan of the tiled charge assignment, exposed for the CPU tests: shared-memory x pitch of a stencil block,
 * lane -> stencil-face point map (-1 = idle lane) and, for a dimension of n grid points, the covering tiles of every
 * coordinate as tile * 16 + local coordinate (-1 = unused), four entries per coordinate */
int b200md_debug_rho_plan(int order, int n, int *pitch, int *lane_point, int *cover);

/* host-only: the signed self-coupled components a dispersion grid is split into (b200md_pppm_params.dispersion = 1, 2, 3
 * and its B array; see disp_components in csrc/pppm.cu).  W[ncomp][ntypes+1], sign[ncomp]; returns ncomp (<= 16) or a
 * negative error.  sum_m sign[m] W[m][i] W[m][j] is the r^-6 coefficient of the type pair. */
int b200md_debug_disp_components(int mix, int ntypes, const double *B, double *W, double *sign);

/* roofline denominators measured on this device ("P_fp measured on the box by a microbenchmark"):
 * kind 0 = FP64 FMA TFLOP/s, 1 = FP32 FMA TFLOP/s, 2 = HBM copy GB/s (read+write bytes).  Not a reference API. */
int b200md_microbench(b200md_ctx *ctx, int kind, double *value);

#ifdef __cplusplus
}
#endif
#endif
