/* intel_preprocess.h — stand-in for the USER-INTEL preprocessor header the reference includes (not shipped in
 * /root/reference).  TEST INFRASTRUCTURE ONLY.  The macros restate the host-only (no offload) variants as SURVEY.md
 * Appendix A.1 states them; the bodies of the reference's loops are NOT here — they are compiled from the reference. */
#ifndef B200MD_REF_INTEL_PREPROCESS_H
#define B200MD_REF_INTEL_PREPROCESS_H

#include <cmath>
#include <cstdint>
#include <cstring>
#if defined(_OPENMP)
#include <omp.h>
#endif

#define _alignvar(expr, val) expr __attribute__((aligned(val)))
#define _noalias __restrict
#define _use_simd_pragma(x)
#define _use_omp_pragma(x)
#define ISFINITE(x) std::isfinite(x)

#define INTEL_DATA_ALIGN 64
#define INTEL_ONEATOM_FACTOR 2
#define INTEL_MIC_VECTOR_WIDTH 16
#define INTEL_VECTOR_WIDTH 4
#define INTEL_MAX_STENCIL 256
#define INTEL_MAX_STENCIL_CHECK 4096
#define INTEL_NBOR_RATIO 100
#define INTEL_BIGP 1e15
#ifndef INTEL_P3M_MAXORDER
#define INTEL_P3M_MAXORDER 7 /* 5 in mid-2016, 7 later; 7 so that the order sweep of the tests compiles */
#endif

enum { LMP_OVERFLOW, LMP_LOCAL_MIN, LMP_LOCAL_MAX, LMP_GHOST_MIN, LMP_GHOST_MAX };
enum {
  TIME_PACK,
  TIME_HOST_NEIGHBOR,
  TIME_HOST_PAIR,
  TIME_OFFLOAD_NEIGHBOR,
  TIME_OFFLOAD_PAIR,
  TIME_OFFLOAD_WAIT,
  TIME_OFFLOAD_LATENCY,
  TIME_IMBALANCE
};

/* ICC intrinsic used by the table branch (pair_buck_coul_long_intel.cpp:320): the bits of a float */
static inline uint32_t __intel_castf32_u32(float f) {
  uint32_t u;
  memcpy(&u, &f, sizeof u);
  return u;
}

#if defined(_OPENMP)
#define IP_PRE_thread_num() omp_get_thread_num()
#else
#define IP_PRE_thread_num() 0
#endif

#define IP_PRE_omp_range(ifrom, ito, tid, inum, nthreads)    \
  {                                                          \
    const int idelta = 1 + inum / nthreads;                  \
    ifrom = tid * idelta;                                    \
    ito = ((ifrom + idelta) > inum) ? inum : ifrom + idelta; \
  }

#define IP_PRE_omp_range_id(ifrom, ito, tid, inum, nthreads) \
  {                                                          \
    tid = IP_PRE_thread_num();                               \
    IP_PRE_omp_range(ifrom, ito, tid, inum, nthreads);       \
  }

#define IP_PRE_omp_range_align(ifrom, ito, tid, inum, nthreads, datasize)                \
  {                                                                                      \
    int chunk_size = INTEL_DATA_ALIGN / datasize;                                        \
    int idelta = static_cast<int>(static_cast<float>(inum) / chunk_size / nthreads) + 1; \
    idelta *= chunk_size;                                                                \
    ifrom = tid * idelta;                                                                \
    ito = ifrom + idelta;                                                                \
    if (ito > inum) ito = inum;                                                          \
  }

#define IP_PRE_omp_range_id_align(ifrom, ito, tid, inum, nthreads, datasize) \
  {                                                                          \
    tid = IP_PRE_thread_num();                                               \
    IP_PRE_omp_range_align(ifrom, ito, tid, inum, nthreads, datasize);       \
  }

#define IP_PRE_get_stride(stride, n, datasize, torque)      \
  {                                                         \
    int blength = n;                                        \
    if (torque) blength *= 2;                               \
    const int bytes = blength * datasize;                   \
    stride = INTEL_DATA_ALIGN - (bytes % INTEL_DATA_ALIGN); \
    stride = blength + stride / datasize;                   \
  }

/* host-only build: nothing to pack separately, nothing to repack */
#define IP_PRE_pack_separate_buffers(fix, buffers, ago, offload, nlocal, nall)
#define IP_PRE_repack_for_offload(newton, separate_flag, nlocal, nall, f_stride, x, q)

#define IP_PRE_get_transfern(ago, newton, evflag, eflag, vflag, buffers, offload, fix, separate_flag, x_size, q_size, \
                             ev_size, f_stride)                                                                       \
  {                                                                                                                   \
    separate_flag = 0;                                                                                                \
    int f_length;                                                                                                     \
    if (newton) f_length = nall;                                                                                      \
    else f_length = nlocal;                                                                                           \
    f_stride = buffers->get_stride(f_length);                                                                         \
  }

#define IP_PRE_get_buffers(offload, buffers, fix, tc, f_start, ev_global) \
  {                                                                       \
    tc = comm->nthreads;                                                  \
    f_start = buffers->get_f();                                           \
    fix->start_watch(TIME_HOST_PAIR);                                     \
    ev_global = buffers->get_ev_global_host();                            \
  }

#define IP_PRE_ev_tally_nbor(vflag, ev_pre, fpair, delx, dely, delz) \
  {                                                                  \
    if (vflag == 1) {                                                \
      sv0 += ev_pre * delx * delx * fpair;                           \
      sv1 += ev_pre * dely * dely * fpair;                           \
      sv2 += ev_pre * delz * delz * fpair;                           \
      sv3 += ev_pre * delx * dely * fpair;                           \
      sv4 += ev_pre * delx * delz * fpair;                           \
      sv5 += ev_pre * dely * delz * fpair;                           \
    }                                                                \
  }

#define IP_PRE_ev_tally_atom(evflag, eflag, vflag, f, fwtmp) \
  {                                                          \
    if (evflag) {                                            \
      if (eflag) {                                           \
        f[i].w += fwtmp;                                     \
        oevdwl += sevdwl;                                    \
      }                                                      \
      if (vflag == 1) {                                      \
        ov0 += sv0;                                          \
        ov1 += sv1;                                          \
        ov2 += sv2;                                          \
        ov3 += sv3;                                          \
        ov4 += sv4;                                          \
        ov5 += sv5;                                          \
      }                                                      \
    }                                                        \
  }

#define IP_PRE_ev_tally_atomq(evflag, eflag, vflag, f, fwtmp) \
  {                                                           \
    if (evflag) {                                             \
      if (eflag) {                                            \
        f[i].w += fwtmp;                                      \
        oevdwl += sevdwl;                                     \
        oecoul += secoul;                                     \
      }                                                       \
      if (vflag == 1) {                                       \
        ov0 += sv0;                                           \
        ov1 += sv1;                                           \
        ov2 += sv2;                                           \
        ov3 += sv3;                                           \
        ov4 += sv4;                                           \
        ov5 += sv5;                                           \
      }                                                       \
    }                                                         \
  }

/* after a barrier: sum the thread-private force arrays into thread 0's, then the f.r virial over the summed forces */
#define IP_PRE_fdotr_acc_force(newton, evflag, eflag, vflag, eatom, nall, nlocal, minlocal, nthreads, f_start, \
                               f_stride, x, offload)                                                           \
  {                                                                                                            \
    int o_range;                                                                                               \
    if (newton) o_range = nall;                                                                                \
    else o_range = nlocal;                                                                                     \
    if (offload == 0) o_range -= minlocal;                                                                     \
    IP_PRE_omp_range_align(iifrom, iito, tid, o_range, nthreads, sizeof(acc_t));                               \
    int t_off = f_stride;                                                                                      \
    if (eflag && eatom) {                                                                                      \
      for (int t = 1; t < nthreads; t++) {                                                                     \
        for (int n = iifrom; n < iito; n++) {                                                                  \
          f_start[n].x += f_start[n + t_off].x;                                                                \
          f_start[n].y += f_start[n + t_off].y;                                                                \
          f_start[n].z += f_start[n + t_off].z;                                                                \
          f_start[n].w += f_start[n + t_off].w;                                                                \
        }                                                                                                      \
        t_off += f_stride;                                                                                     \
      }                                                                                                        \
    } else {                                                                                                   \
      for (int t = 1; t < nthreads; t++) {                                                                     \
        for (int n = iifrom; n < iito; n++) {                                                                  \
          f_start[n].x += f_start[n + t_off].x;                                                                \
          f_start[n].y += f_start[n + t_off].y;                                                                \
          f_start[n].z += f_start[n + t_off].z;                                                                \
        }                                                                                                      \
        t_off += f_stride;                                                                                     \
      }                                                                                                        \
    }                                                                                                          \
    if (evflag) {                                                                                              \
      if (vflag == 2) {                                                                                        \
        const ATOM_T *_noalias const xo = x + minlocal;                                                        \
        for (int n = iifrom; n < iito; n++) {                                                                  \
          ov0 += f_start[n].x * xo[n].x;                                                                       \
          ov1 += f_start[n].y * xo[n].y;                                                                       \
          ov2 += f_start[n].z * xo[n].z;                                                                       \
          ov3 += f_start[n].y * xo[n].x;                                                                       \
          ov4 += f_start[n].z * xo[n].x;                                                                       \
          ov5 += f_start[n].z * xo[n].y;                                                                       \
        }                                                                                                      \
      }                                                                                                        \
    }                                                                                                          \
  }

#endif
