/* oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of the loops in
 * HPAC/lammps-buck-intel (and of the stock-LAMMPS behaviour those loops rely on).
 * It exists to check the CUDA product and to provide the reported CPU baseline.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (lammps-buck-intel_b200/) never links or calls it.
 *
 * PARITY PINNED against the reference's own code: oracle/Makefile.ref compiles pair_buck*_intel.cpp, pppm_intel.cpp
 * and fix_nve_intel.cpp UNCHANGED from /root/reference (against the stand-in LAMMPS headers of oracle/ref_shim/) into
 * oracle/_ref/libref.so, and tests/test_oracle_vs_ref.py asserts that this restatement and that library agree BIT FOR
 * BIT (one thread) on forces, per-atom energies, energies, both virial forms, densities, fields, x and v, in double and
 * mixed precision, analytic and table branches.  What the reference does not ship (the stock base classes: init_one,
 * Pair::init_tables, PPPM::set_grid_global / compute_gf_ik / compute_rho_coeffs, the neighbour list, FFT3d) stays
 * pinned only by the known-answer tests this repo authors (tests/test_oracle_kat.py): closed-form dimers, -dE/dr, libm
 * erfc, direct Ewald sums, the NaCl Madelung constant.  pppm_disp_intel.cpp is not in oracle/_ref (needs the whole stock
 * PPPMDisp and does not compile as shipped): the dispersion grid is KAT-pinned only.
 *
 * Conventions: atom types are 1-based (LAMMPS); per-type-pair arrays are
 * (ntypes+1)x(ntypes+1) row-major, index [itype*(ntypes+1)+jtype]; x is [n][3];
 * neighbour entries carry the special-bond index in bits 30-31 (SBBITS = 30).
 */
#ifndef B200MD_ORACLE_H
#define B200MD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_BUCK = 0, ORC_BUCK_COUL_CUT = 1, ORC_BUCK_COUL_LONG = 2, ORC_BUCK_LONG_COUL_LONG = 3,
       /* pair_lj_long_coul_long_intel.cpp (`lj/long/coul/long cut long` is the lj/cut/coul/long of in.spce):
        * A = epsilon, rho = sigma on input; buck1, buck2, a, c carry lj1, lj2, lj3, lj4 */
       ORC_LJ_LONG_COUL_LONG = 4 };
enum { ORC_PREC_DOUBLE = 0, ORC_PREC_MIXED = 1 };

/* packed per-type-pair constants, doubles; restates ForceConst<flt_t> of all four styles
 * (pair_buck_intel.h:60-80, pair_buck_coul_cut_intel.h:60-70, pair_buck_coul_long_intel.h:63-73,
 *  pair_buck_long_coul_long_intel.h) */
typedef struct {
  int ntypes;            /* number of atom types (arrays are (ntypes+1)^2) */
  double *cutsq;         /* outer cutoff^2 (max of lj/coul) */
  double *cut_ljsq;      /* Buckingham cutoff^2 */
  double *cut_coulsq;    /* Coulomb cutoff^2 (coul/cut per pair, coul/long global) */
  double *buck1;         /* A/rho */
  double *buck2;         /* 6C */
  double *rhoinv;        /* 1/rho */
  double *a;             /* A */
  double *c;             /* C */
  double *offset;        /* energy shift at cutoff */
  double special_lj[4];
  double special_coul[4];
  double qqrd2e;
  double g_ewald;        /* coul/long and long/coul/long ORDER1 */
  double g_ewald_6;      /* long/coul/long ORDER6 */
  int order1, order6;    /* long/coul/long: ewald_order bits */
  /* Coulomb tables (INTEL_ALLOW_TABLE path); ncoultablebits == 0 -> analytic */
  int ncoultablebits, ncoulmask, ncoulshiftbits;
  double tabinnersq;
  const double *rtable, *drtable, *ftable, *dftable, *etable, *detable, *ctable, *dctable;
  /* dispersion tables (long/coul/long ORDER6 table path); ndisptablebits == 0 -> analytic */
  int ndisptablebits, ndispmask, ndispshiftbits;
  double tabinnerdispsq;
  const double *rdisptable, *drdisptable, *fdisptable, *dfdisptable, *edisptable, *dedisptable;
} orc_pair_params;

/* ---- neighbour / ghosts (stock LAMMPS Comm::borders + NPairHalfBinNewton) */

/* Periodic ghost images within cutghost of each face, built dimension by dimension.
 * x holds nlocal atoms on entry and room for cap atoms; returns nghost (or -1 if cap too small).
 * src[g] = index (into nall so far) of the atom ghost g copies; shift[g][3] accumulated image. */
int orc_make_ghosts(int nlocal, double *x, int *type, double *q, const double *boxlo,
                    const double *boxhi, const int *periodic, double cutghost, int cap,
                    int *src, int *shift);

/* Half list, newton on, binned (bin = cutneighmax/2).  prec selects the arithmetic of the
 * distance test (the "intel" list evaluates it in flt_t, pair_buck_intel.cpp:399-409).
 * CSR out: numneigh[nlocal], offsets[nlocal+1], entries (cap_entries).  Returns total entries
 * or -1 on overflow. */
long orc_neigh_half_bin(int nlocal, int nall, const double *x, const int *type, int ntypes,
                        const double *cutneighsq, const double *boxlo, const double *boxhi,
                        double cutneighmax, int prec, int *numneigh, long *offsets,
                        int *entries, long cap_entries);

/* Full list, newton off, brute force O(nlocal*nall) — the known-answer pair set. */
long orc_neigh_full_brute(int nlocal, int nall, const double *x, const int *type, int ntypes,
                          const double *cutneighsq, int prec, int *numneigh, long *offsets,
                          int *entries, long cap_entries);

/* ---- pair styles */

/* PairBuck*::init_one for all type pairs: fills the derived arrays of p from
 * A, rho, C, cut_lj, cut_coul ((ntypes+1)^2 each; cut_coul may be NULL). */
void orc_pair_init(int style, int ntypes, const double *A, const double *rho, const double *C,
                   const double *cut_lj, const double *cut_coul, int offset_flag,
                   orc_pair_params *p);

/* Pair::init_tables + init_bitmap.  All out arrays have 2^nbits doubles. */
void orc_init_coul_tables(double cut_coul, double tabinner, int nbits, double g_ewald,
                          double qqrd2e, double *rtable, double *drtable, double *ftable,
                          double *dftable, double *etable, double *detable, double *ctable,
                          double *dctable, int *ncoulmask, int *ncoulshiftbits,
                          double *tabinnersq);

/* eval<EVFLAG,EFLAG,NEWTON_PAIR> of the four styles on a CSR list.
 * newton=1: half list, f[j] updated for every j (nall forces); newton=0: reference's
 * NEWTON_PAIR=0 rule (f[j] only for j<nlocal).  vflag: 0 none, 1 pair tally, 2 f.r over nall.
 * f is [nall][4] (w = per-atom energy when eatom), zeroed here.  ev[8] = {evdwl, ecoul, v0..v5}.
 * nthreads<=0 -> omp_get_max_threads(). */
void orc_pair_eval(int style, int prec, int eflag, int vflag, int eatom, int newton, int nlocal,
                   int nall, const double *x, const int *type, const double *q,
                   const int *numneigh, const long *offsets, const int *entries,
                   const orc_pair_params *p, double *f, double *ev, int nthreads);

/* Comm::reverse_comm for the ghosts made by orc_make_ghosts: f[src[g]] += f[g], last ghost first. */
void orc_reverse_comm(int nlocal, int nghost, const int *src, double *f);

/* ---- fix nve/intel (fix_nve_intel.cpp:60-127); dtfm per coordinate [3*nlocal] */
void orc_nve_dtfm(int nlocal, const int *type, const double *mass, double dt, double ftm2v,
                  double *dtfm);
void orc_nve_initial(int nlocal, double *x, double *v, const double *f, const double *dtfm,
                     double dtv);
void orc_nve_final(int nlocal, double *v, const double *f, const double *dtfm);
/* sub-group / per-atom mass branches (fix_nve_intel.cpp:88-97, 147-190); rmass and ingroup may be NULL */
void orc_nve_dtfm_group(int nlocal, const int *type, const double *mass, const double *rmass, const int *ingroup,
                        double dt, double ftm2v, double *dtfm);
void orc_nve_initial_group(int nlocal, double *x, double *v, const double *f, const double *dtfm, double dtv);

/* ---- PPPM (pppm_intel.cpp + stock PPPM) */
typedef struct orc_pppm orc_pppm;

/* accuracy-driven sizing: g_ewald, nx,ny,nz exactly as PPPM::init/set_grid_global/adjust_gewald.
 * grid[3] in/out: if all >0 on entry they are kept (kspace_modify mesh); g_ewald in/out likewise
 * (<=0 -> computed). */
void orc_pppm_size(double accuracy_relative, double two_charge_force, double qqrd2e,
                   double qsqsum, long natoms, double cutoff, const double *prd, int order,
                   int diff_ad, int *grid, double *g_ewald);

orc_pppm *orc_pppm_create(int nx, int ny, int nz, int order, double g_ewald, int diff_ad,
                          const double *boxlo, const double *boxhi, double qqrd2e, int prec);
/* PPPMDispIntel 'g' (geometric mixing) grid: q array = B[type], g_ewald = g_ewald_6 (pppm_disp_intel.cpp:245-313,
 * 486-510 + upstream PPPMDisp::compute_gf_6 / vg_6) */
orc_pppm *orc_pppm_create_disp(int nx, int ny, int nz, int order, double g_ewald_6, const double *boxlo,
                               const double *boxhi, int prec);
/* the same grid with kspace_modify diff ad (fieldforce_g_ad, compute_sf_coeff_6 [UPSTREAM]) */
orc_pppm *orc_pppm_create_disp_ad(int nx, int ny, int nz, int order, double g_ewald_6, int diff_ad,
                                  const double *boxlo, const double *boxhi, int prec);
/* PPPMDispIntel::compute function[2] / function[3] (pppm_disp_intel.cpp:315-467 + stock PPPMDisp members): arithmetic
 * mixing on seven coupled grids, w7[n][7] = B[7 type + k]; no mixing rule on nsplit eigen-grids,
 * wn[n][nsplit] = B[nsplit type + k], lam[nsplit] the eigenvalues.  p from orc_pppm_create_disp[_ad]. */
void orc_pppm_compute_arith(orc_pppm *p, int nlocal, const double *x, const double *w7, int eflag, int vflag, double *f,
                            double *energy, double *virial, int nthreads);
void orc_pppm_compute_none(orc_pppm *p, int nlocal, const double *x, int nsplit, const double *wn, const double *lam,
                           int eflag, int vflag, double *f, double *energy, double *virial, int nthreads);
/* triclinic box: PPPMIntel::compute with domain->triclinic (pppm_intel.cpp:151-156, 307-309, 878-883 + stock
 * setup_triclinic / compute_gf_ik_triclinic / poisson_ik_triclinic); x in box coordinates, ik differentiation */
orc_pppm *orc_pppm_create_tri(int nx, int ny, int nz, int order, double g_ewald, const double *boxlo, const double *boxhi,
                              double xy, double xz, double yz, double qqrd2e, int prec);
void orc_ewald_recip_tri(int n, const double *x, const double *q, const double *boxlo, const double *boxhi, double xy,
                         double xz, double yz, double g_ewald, int kmax, double qqrd2e, double *f, double *energy,
                         double *virial);
/* kspace_modify slab: mesh over zprd * slab_volfactor, PPPM::slabcorr applied (pppm_intel.cpp:305); z non-periodic */
orc_pppm *orc_pppm_create_slab(int nx, int ny, int nz, int order, double g_ewald, int diff_ad, const double *boxlo,
                               const double *boxhi, double qqrd2e, int prec, double slab_volfactor);
void orc_pppm_destroy(orc_pppm *p);
/* PPPMIntel::compute (single rank, periodic).  f[nlocal][3] is accumulated into (+=).
 * energy / virial[6] written when eflag / vflag. */
void orc_pppm_compute(orc_pppm *p, int nlocal, const double *x, const double *q, int eflag,
                      int vflag, double *f, double *energy, double *virial, int nthreads);
/* introspection for parity tests */
/* per-atom tallies of the last compute with eflag & 2 / vflag & 4 (stock poisson_peratom / fieldforce_peratom):
 * eatom[nlocal], vatom[nlocal][6] (xx,yy,zz,xy,xz,yz); Coulomb grid only */
void orc_pppm_peratom(const orc_pppm *p, double *eatom, double *vatom);
long orc_pppm_nfft(const orc_pppm *p);
const double *orc_pppm_greensfn(const orc_pppm *p);
const double *orc_pppm_density_fft(const orc_pppm *p);   /* after compute: folded density, nfft */
const double *orc_pppm_field(const orc_pppm *p, int dim);/* after compute: vdx/vdy/vdz (or u) nfft */
const double *orc_pppm_sf_coeff(const orc_pppm *p);
void orc_pppm_rho_coeff(const orc_pppm *p, double *rho_coeff /*order*order*/, double *drho_coeff);

/* everything PPPMIntel reads from its base class, exported so that the reference's own compiled
 * loops (oracle/_ref, ref_harness.cpp) can run on exactly the state the restatement uses */
typedef struct {
  int nx, ny, nz, order, diff_ad, nlower, nupper;
  int lo_out[3], hi_out[3];
  double shift, shiftone, g_ewald, qqrd2e, scale, volume;
  double boxlo[3], prd[3], delinv[3], delvolinv;
  const double *greensfn, *vg /*[nfft][6]*/, *fkx, *fky, *fkz;
  const double *rho_coeff, *drho_coeff /*[order][order], k - nlower*/;
  double sf_coeff[6];
} orc_pppm_state;
void orc_pppm_export(const orc_pppm *p, orc_pppm_state *st);

/* 3-D complex FFT used by the oracle (KISS-style mixed radix), dir=+1 is exp(+ikx) (LAMMPS flag=1), -1 is exp(-ikx),
 * unnormalised, interleaved re/im, x fastest. */
void orc_fft3d(double *data, int nx, int ny, int nz, int dir, int nthreads);

/* ---- direct Ewald sum (cross-check of PPPM + real space; kspace_style ewald semantics) */
void orc_ewald_recip(int n, const double *x, const double *q, const double *boxlo,
                     const double *boxhi, double g_ewald, int kmax, double qqrd2e, double *f,
                     double *energy, double *virial);

/* ---- whole MD step on the CPU (baseline timing): neighbour decide/build + pair + pppm + nve */
typedef struct orc_md orc_md;
orc_md *orc_md_create(int nlocal, const double *x, const double *v, const double *q,
                      const int *type, int ntypes, const double *mass, const double *boxlo,
                      const double *boxhi, int style, int prec, const orc_pair_params *p,
                      double skin, int every, int delay, int check, double dt, double ftm2v,
                      orc_pppm *pppm /* may be NULL */);
void orc_md_destroy(orc_md *m);
/* runs nsteps velocity-Verlet steps; timers[8] accumulate seconds {neigh,pair,kspace,nve,comm,..} */
void orc_md_run(orc_md *m, int nsteps, int nthreads, double *timers, int *nbuilds);
/* 1 once an atom has left the box by more than one period (or is NaN): orc_md_run stops stepping instead of binning it */
int orc_md_lost(const orc_md *m);
void orc_md_get(orc_md *m, double *x, double *v, double *f);
void orc_md_energy(orc_md *m, int nthreads, double *ev /*8*/, double *ekspace, double *ke);

#ifdef __cplusplus
}
#endif
#endif
