// pair_mixed.cu — mixed-precision (<float,double>) instantiation of the pair kernels (pair_buck_intel.cpp:50-58
// PREC_MODE_MIXED).  Built with -fmad=false, see pair_kernel.cuh.
#include "pair_kernel.cuh"

int b2_launch_pair_float(b200md_ctx *ctx, const pairk::PairView &v, long long total_entries, int evflag, double *ev_dev,
                         int has_special) {
  return pairk::launch_pair<float>(ctx, v, total_entries, evflag, ev_dev, has_special);
}
