// style_pair.h — the list LAMMPS's build generates from src/pair_*.h (Make.sh style): every header of a pair style,
// included by the driver with PAIR_CLASS defined so that only the PairStyle(key,Class) lines are seen
// (pair_buck_intel.h:18-22 of the reference and its four siblings).
#include "pair_buck_intel.h"
#include "pair_buck_coul_intel.h"
#include "pair_lj_long_coul_long_intel.h"
