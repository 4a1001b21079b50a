/* mpi.h — single-process stand-in for the handful of MPI calls in the reference (pppm_intel.cpp:260,273).  TEST ONLY. */
#ifndef B200MD_REF_MPI_STUB_H
#define B200MD_REF_MPI_STUB_H
#include <cstring>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_SUM 0
#define MPI_COMM_WORLD 0
static inline int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op, MPI_Comm) {
  memcpy(r, s, (size_t)count * (size_t)t);
  return 0;
}
#endif
