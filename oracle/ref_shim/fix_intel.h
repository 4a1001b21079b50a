/* fix_intel.h — stand-in for FixIntel (`package intel`), host-only, no offload.
 * TEST INFRASTRUCTURE ONLY.  Owns the three IntelBuffers instances, records the result arrays the pair styles hand
 * over (add_result_array) and folds them into atom->f and the pair tallies where stock LAMMPS does it (pre_reverse):
 * sum of the thread-private force arrays in thread order, then f += and eng_vdwl / eng_coul / virial +=. */
#ifndef B200MD_REF_FIX_INTEL_H
#define B200MD_REF_FIX_INTEL_H

#include "lammps_stub.h"
#include "intel_preprocess.h"
#include "intel_buffers.h"   /* the reference's own header */

namespace LAMMPS_NS {

class FixIntel : public Fix {
 public:
  enum { PREC_MODE_SINGLE, PREC_MODE_MIXED, PREC_MODE_DOUBLE };

  FixIntel(LAMMPS *lmp, int narg, char **arg, int precision_mode)
      : Fix(lmp, narg, arg), _precision_mode(precision_mode) {
    _single_buffers = new IntelBuffers<float, float>(lmp);
    _mixed_buffers = new IntelBuffers<float, double>(lmp);
    _double_buffers = new IntelBuffers<double, double>(lmp);
    _overflow_flag[0] = 0;
  }
  ~FixIntel() {
    delete _single_buffers;
    delete _mixed_buffers;
    delete _double_buffers;
  }

  inline int precision() { return _precision_mode; }
  inline IntelBuffers<float, float> *get_single_buffers() { return _single_buffers; }
  inline IntelBuffers<float, double> *get_mixed_buffers() { return _mixed_buffers; }
  inline IntelBuffers<double, double> *get_double_buffers() { return _double_buffers; }

  inline void pair_init_check() {}
  inline void kspace_init_check() {}
  inline void balance_stamp() {}
  inline int host_start_pair() { return 0; }
  inline int offload_end_pair() { return 0; }
  inline int separate_buffers() { return 0; }
  inline double offload_balance() { return 0.0; }
  inline int coprocessor_number() { return -1; }
  inline int *get_off_overflow_flag() { return _overflow_flag; }
  inline double *off_watch_pair() { return &_watch; }
  inline void start_watch(const int) {}
  inline double stop_watch(const int) { return 0.0; }

  inline void get_buffern(const int, int &nlocal, int &nall, int &minlocal) {
    nall = atom->nlocal + atom->nghost;
    nlocal = atom->nlocal;
    minlocal = 0;
  }

  template <class ft, class acc_t>
  void add_result_array(ft *f_in, acc_t *ev_in, const int /*offload*/, const int eatom = 0, const int vatom = 0,
                        const int rflag = 0) {
    (void)vatom;
    const int nthreads = comm->nthreads;
    int o_range;
    if (force->newton_pair) o_range = atom->nlocal + atom->nghost;
    else o_range = atom->nlocal;
    if (rflag != 2 && nthreads > 1) {   /* reduce_results: thread copies into thread 0's, ascending thread order */
      int f_stride;
      IP_PRE_get_stride(f_stride, o_range, sizeof(ft), atom->torque);
      for (int n = 0; n < o_range; n++) {
        int t_off = f_stride;
        for (int t = 1; t < nthreads; t++) {
          f_in[n].x += f_in[n + t_off].x;
          f_in[n].y += f_in[n + t_off].y;
          f_in[n].z += f_in[n + t_off].z;
          if (eatom) f_in[n].w += f_in[n + t_off].w;
          t_off += f_stride;
        }
      }
    }
    /* add_results (done in pre_reverse upstream; nothing runs in between here) */
    double **f = atom->f;
    for (int i = 0; i < o_range; i++) {
      f[i][0] += f_in[i].x;
      f[i][1] += f_in[i].y;
      f[i][2] += f_in[i].z;
      if (eatom) force->pair->eatom[i] += f_in[i].w;
    }
    if (ev_in != 0) {
      force->pair->eng_vdwl += ev_in[0];
      force->pair->eng_coul += ev_in[1];
      for (int k = 0; k < 6; k++) force->pair->virial[k] += ev_in[2 + k];
    }
  }
  /* the `fix->add_result_array(f_start, 0, offload)` form of the no-tally evals */
  template <class ft>
  void add_result_array(ft *f_in, int /*null ev*/, const int offload) {
    add_result_array(f_in, (double *)0, offload, 0, 0, 0);
  }

 protected:
  int _precision_mode;
  IntelBuffers<float, float> *_single_buffers;
  IntelBuffers<float, double> *_mixed_buffers;
  IntelBuffers<double, double> *_double_buffers;
  int _overflow_flag[5];
  double _watch = 0.0;
};

}  // namespace LAMMPS_NS
#endif
