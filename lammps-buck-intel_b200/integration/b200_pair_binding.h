// b200_pair_binding.h — the marshalling every Pair*Intel class of the B200 build shares.  Macros, because they expand
// inside member functions: the members they read (setflag, cutsq, the lookup tables, eflag_global ...) are protected
// in stock LAMMPS.  `fix` is the class's FixIntel*.
#ifndef B200MD_B200_PAIR_BINDING_H
#define B200MD_B200_PAIR_BINDING_H

#include <vector>

#include "atom.h"
#include "error.h"
#include "force.h"
#include "kspace.h"
#include "modify.h"
#include "suffix.h"

#include "b200_fix_intel.h"

// init_style: the `package intel` fix, with the reference's message (pair_buck_intel.cpp:372-376)
#define B200_FIND_FIX_INTEL()                                                                               \
  do {                                                                                                      \
    const int ifix_ = modify->find_fix("package_intel");                                                    \
    if (ifix_ < 0) error->all(FLERR, "The 'package intel' command is required for /intel styles");          \
    fix = static_cast<FixIntel *>(modify->fix[ifix_]);                                                      \
    fix->pair_init_check();                                                                                 \
  } while (0)

// pack_force_const repeats init_one for every type pair before it copies the coefficients (pair_buck_intel.cpp:399-409)
#define B200_INIT_ALL_PAIRS()                                                                               \
  do {                                                                                                      \
    for (int i_ = 1; i_ <= atom->ntypes; i_++)                                                              \
      for (int j_ = i_; j_ <= atom->ntypes; j_++)                                                           \
        if (setflag[i_][j_] != 0 || (setflag[i_][i_] != 0 && setflag[j_][j_] != 0)) {                       \
          const double cut_ = init_one(i_, j_);                                                             \
          cutsq[i_][j_] = cutsq[j_][i_] = cut_ * cut_;                                                      \
        }                                                                                                   \
  } while (0)

#define B200_PACK_SPECIAL(p)                                                                                \
  do {                                                                                                      \
    for (int k_ = 0; k_ < 4; k_++) {                                                                        \
      (p).special_lj[k_] = force->special_lj[k_];                                                           \
      (p).special_coul[k_] = force->special_coul[k_];                                                       \
    }                                                                                                       \
  } while (0)

// the products of Pair::init_tables / init_tables_disp, as the base class holds them
#define B200_PACK_COUL_TABLES(p)                                                                            \
  do {                                                                                                      \
    (p).ncoultablebits = ncoultablebits; (p).ncoulmask = ncoulmask; (p).ncoulshiftbits = ncoulshiftbits;    \
    (p).tabinnersq = tabinnersq;                                                                            \
    (p).rtable = rtable; (p).drtable = drtable; (p).ftable = ftable; (p).dftable = dftable;                 \
    (p).etable = etable; (p).detable = detable; (p).ctable = ctable; (p).dctable = dctable;                 \
  } while (0)

#define B200_PACK_DISP_TABLES(p)                                                                            \
  do {                                                                                                      \
    (p).ndisptablebits = ndisptablebits; (p).ndispmask = ndispmask; (p).ndispshiftbits = ndispshiftbits;    \
    (p).tabinnerdispsq = tabinnerdispsq;                                                                    \
    (p).rdisptable = rdisptable; (p).drdisptable = drdisptable; (p).fdisptable = fdisptable;                \
    (p).dfdisptable = dfdisptable; (p).edisptable = edisptable; (p).dedisptable = dedisptable;              \
  } while (0)

#define B200_PAIR_SETUP(p)                                                                                  \
  do {                                                                                                      \
    if (b200md_pair_setup(b200_ctx(fix), &(p))) error->all(FLERR, b200md_last_error(b200_ctx(fix)));        \
  } while (0)

// compute(): ev_setup, positions to the device, the kernel, the tallies of ev_global (pair_buck_intel.cpp:337-349),
// forces back (add_result_array)
#define B200_PAIR_COMPUTE(eflag, vflag)                                                                     \
  do {                                                                                                      \
    if ((eflag) || (vflag)) ev_setup(eflag, vflag);                                                         \
    else evflag = vflag_fdotr = 0;                                                                          \
    b200_positions_to_device(fix);                                                                          \
    double ev_[8];                                                                                          \
    if (b200md_pair_compute(b200_ctx(fix), eflag, vflag, ev_)) error->one(FLERR, b200md_last_error(b200_ctx(fix))); \
    if (eflag_global) { eng_vdwl += ev_[0]; eng_coul += ev_[1]; }                                           \
    if (vflag_global)                                                                                       \
      for (int n_ = 0; n_ < 6; n_++) virial[n_] += ev_[2 + n_];                                             \
    b200_forces_to_host(fix);                                                                               \
  } while (0)

#endif
