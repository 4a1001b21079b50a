/* pair_lj_long_coul_long.h — forwarder: the stand-in classes live in lammps_stub.h (TEST INFRASTRUCTURE, see that file) */
#include "lammps_stub.h"
