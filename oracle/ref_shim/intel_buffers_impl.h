/* intel_buffers_impl.h — definitions of the IntelBuffers members that /root/reference/intel_buffers.h declares but whose
 * bodies live in stock intel_buffers.cpp (not shipped).  TEST INFRASTRUCTURE ONLY: plain aligned allocations sized as
 * the reference's accessors assume (one force array of get_stride(nall) rows per thread; one padding atom). */
#ifndef B200MD_REF_INTEL_BUFFERS_IMPL_H
#define B200MD_REF_INTEL_BUFFERS_IMPL_H
#include "fix_intel.h"

namespace LAMMPS_NS {

static inline void *ib_alloc(size_t bytes) {
  void *p = nullptr;
  if (posix_memalign(&p, INTEL_DATA_ALIGN, bytes ? bytes : INTEL_DATA_ALIGN)) throw std::bad_alloc();
  memset(p, 0, bytes);
  return p;
}

template <class flt_t, class acc_t>
IntelBuffers<flt_t, acc_t>::IntelBuffers(class LAMMPS *lmp_in) : lmp(lmp_in) {
  _x = 0; _q = 0; _quat = 0; _f = 0;
  _off_threads = 0; _off_map_maxlocal = 0;
  _list_alloc_atoms = 0; _list_alloc = 0; _cnumneigh = 0; _atombin = 0; _binpacked = 0;
  _cutneighsq = 0; _ntypes = 0;
  _ccache_stride = 0; _ccachex = _ccachey = _ccachez = _ccachew = 0; _ccachei = _ccachej = 0;
  _buf_size = _buf_local_size = 0; _host_nmax = 0;
  _special_holder = 0; _nspecial_holder = 0;
}

template <class flt_t, class acc_t>
IntelBuffers<flt_t, acc_t>::~IntelBuffers() {
  free_buffers();
  free_all_nbor_buffers();
  set_ntypes(0);
}

template <class flt_t, class acc_t>
void IntelBuffers<flt_t, acc_t>::free_buffers() {
  free(_x); free(_q); free(_f);
  _x = 0; _q = 0; _f = 0;
  _buf_size = _buf_local_size = 0;
}

template <class flt_t, class acc_t>
void IntelBuffers<flt_t, acc_t>::_grow(const int nall, const int nlocal, const int nthreads, const int) {
  free_buffers();
  _buf_size = static_cast<int>(nall * 1.1 + 1);
  _buf_local_size = static_cast<int>(nlocal * 1.1 + 1);
  _x = (atom_t *)ib_alloc(sizeof(atom_t) * (size_t)(_buf_size + 1));
  _q = (flt_t *)ib_alloc(sizeof(flt_t) * (size_t)(_buf_size + 1));
  const int f_stride = get_stride(_buf_size);
  _f = (vec3_acc_t *)ib_alloc(sizeof(vec3_acc_t) * (size_t)f_stride * (size_t)nthreads);
}

template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::free_nmax() { _host_nmax = 0; }
template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::_grow_nmax(const int) { _host_nmax = lmp->atom->nmax; }
template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::free_local() { _off_map_maxlocal = 0; }
template <class flt_t, class acc_t>
void IntelBuffers<flt_t, acc_t>::_grow_local(NeighList *list, const int) { _off_map_maxlocal = list->get_maxlocal(); }
template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::free_binhead() {}
template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::_grow_binhead() {}
template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::_grow_stencil(NeighList *) {}
template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::free_ccache() {}
template <class flt_t, class acc_t> void IntelBuffers<flt_t, acc_t>::grow_ccache(const int, const int, const int) {}
template <class flt_t, class acc_t> double IntelBuffers<flt_t, acc_t>::memory_usage(const int) { return 0.0; }

template <class flt_t, class acc_t>
void IntelBuffers<flt_t, acc_t>::free_nbor_list() {
  free(_list_alloc); free(_cnumneigh);
  _list_alloc = 0; _cnumneigh = 0;
  _list_alloc_atoms = 0;
}

/* the harness sizes the packed list through NeighList::maxlocal = total entries (it fills firstneigh()/cnumneigh()
 * itself, standing in for the intel neighbour build that is not part of the reference) */
template <class flt_t, class acc_t>
void IntelBuffers<flt_t, acc_t>::_grow_nbor_list(NeighList *list, const int nlocal, const int, const int, const int) {
  free_nbor_list();
  _list_alloc_atoms = nlocal;
  _list_alloc = (int *)ib_alloc(sizeof(int) * ((size_t)list->maxlocal + 64));
  _cnumneigh = (int *)ib_alloc(sizeof(int) * ((size_t)nlocal + 1));
}

template <class flt_t, class acc_t>
void IntelBuffers<flt_t, acc_t>::set_ntypes(const int ntypes) {
  if (ntypes != _ntypes) {
    if (_ntypes > 0) lmp->memory->destroy(_cutneighsq);
    if (ntypes > 0) lmp->memory->create(_cutneighsq, ntypes, ntypes, "_cutneighsq");
    _ntypes = ntypes;
  }
}

template class IntelBuffers<float, float>;
template class IntelBuffers<float, double>;
template class IntelBuffers<double, double>;

}  // namespace LAMMPS_NS
#endif
