/* b200md.h — C ABI of the B200-native Buckingham pair / PPPM / NVE hot path.
 *
 * This is the drop-in boundary for the /intel styles of HPAC/lammps-buck-intel: every entry point
 * names the reference interface (file:line, relative to the reference tree) it replaces.  Plain C,
 * plain pointers and sizes; no torch / CUDA types.  All functions return 0 on success; on failure they
 * return a negative B200MD_E* code and b200md_last_error(ctx) holds the message — the same text the
 * reference passes to error->all()/error->one() where one exists (SURVEY.md §8b "Errors").
 *
 * There is no CPU fallback: b200md_ctx_create fails when no sm_100 device is usable.
 *
 * Conventions
 *   - atom types are 1-based; per-type-pair arrays are (ntypes+1)^2 row-major [itype*(ntypes+1)+jtype]
 *   - host arrays: x,v,f are [n][3] doubles (LAMMPS atom->x[0] layout, fix_nve_intel.cpp:64-66)
 *   - atoms live in HBM between calls, cell-sorted; "host order" always means the order of the last
 *     b200md_atoms_upload (LAMMPS local index); downloads un-permute
 *   - ev[8] = {evdwl, ecoul, v_xx, v_yy, v_zz, v_xy, v_xz, v_yz}  (ev_global, pair_buck_intel.cpp:337-349)
 *   - eflag bit0 global energy, bit1 per-atom; vflag 1|2 global virial (newton is off on the device, so
 *     both are evaluated as the per-pair tally, pair_buck_intel.cpp:93-96,312), 4 per-atom (not produced by the
 *     pair styles, as in the reference: pair_buck_intel.cpp:362; PPPM does produce it: b200md_pppm_peratom)
 */
#ifndef B200MD_H
#define B200MD_H

#include <stddef.h>
#include <stdint.h>

/* Limits (stated, and reported as errors rather than silently exceeded):
 *  - at most B200MD_MAX_TYPES = 8 atom types (the per-type-pair constants are staged in shared memory):
 *    b200md_pair_setup fails with B200MD_EINVAL beyond that;
 *  - lists built on the device carry the type of j in bits 26..29 of an entry while owned + ghost atoms of one GPU stay
 *    <= 2^26 (67 M) and ntypes < 16; beyond that the entries are plain 30-bit indices and the pair kernel gathers
 *    type[j] (slower, same results); 2^30 owned + ghost atoms per GPU is the hard limit of the entry format
 *    (NEIGHMASK, like the reference's);
 *  - a neighbour list may hold more than 2^31 entries (offsets are 64-bit). */
#define B200MD_MAX_TYPES 8

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200md_ctx b200md_ctx;

enum {
  B200MD_OK = 0,
  B200MD_EINVAL = -1,    /* bad argument / call order */
  B200MD_ECUDA = -2,     /* CUDA runtime error */
  B200MD_ENODEV = -3,    /* no usable sm_100 device */
  B200MD_ERANGE = -4,    /* "Out of range atoms - cannot compute PPPM" (pppm_intel.cpp:385) */
  B200MD_EORDER = -5,    /* "PPPM order greater than supported by USER-INTEL" (pppm_intel.cpp:87-88) */
  B200MD_ENOMEM = -6,
  B200MD_EOVERFLOW = -7, /* neighbour storage / box too small for the ghost cutoff */
  B200MD_ENONFINITE = -8,/* "Non-numeric box dimensions - simulation unstable" (pppm_intel.cpp:342) */
  B200MD_ECOMM = -9      /* NCCL error */
};

/* FixIntel::PREC_MODE_* (pair_buck_intel.cpp:50-58).  SINGLE is not provided on the device. */
enum { B200MD_PREC_DOUBLE = 0, B200MD_PREC_MIXED = 1 };

/* pair styles: PairStyle(buck/intel,…) pair_buck_intel.h:20, buck/coul/cut/intel
 * pair_buck_coul_cut_intel.h:21, buck/coul/long/intel pair_buck_coul_long_intel.h:20,
 * buck/long/coul/long/intel pair_buck_long_coul_long_intel.h:20 */
enum {
  B200MD_PAIR_BUCK = 0,
  B200MD_PAIR_BUCK_COUL_CUT = 1,
  B200MD_PAIR_BUCK_COUL_LONG = 2,
  B200MD_PAIR_BUCK_LONG_COUL_LONG = 3,
  /* lj/long/coul/long/intel, pair_lj_long_coul_long_intel.h:20 (with `cut long` it is the
   * lj/cut/coul/long of examples/in.spce:7).  In b200md_pair_params buck1, buck2, a, c carry lj1, lj2, lj3, lj4
   * (pair_lj_long_coul_long_intel.cpp:831-834); rhoinv is not read. */
  B200MD_PAIR_LJ_LONG_COUL_LONG = 4
};

/* ------------------------------------------------------------------------------------------------
 * context  — replaces FixIntel ("package intel … mode double|mixed") + IntelBuffers
 * ownership (intel_buffers.h:272-312): the context owns every device array. */
int b200md_ctx_create(int device, int precision, b200md_ctx **out);
void b200md_ctx_destroy(b200md_ctx *ctx);
const char *b200md_last_error(const b200md_ctx *ctx); /* ctx may be NULL: last create error */
int b200md_version(void);

/* units: force->qqrd2e, force->ftm2v (pair_buck_coul_long_intel.cpp:157, fix_nve_intel.cpp:131) */
int b200md_set_units(b200md_ctx *ctx, double qqrd2e, double ftm2v);
/* domain->boxlo/boxhi/periodicity (orthogonal boxes; pppm_intel.cpp:153 reads boxlo) */
int b200md_set_box(b200md_ctx *ctx, const double boxlo[3], const double boxhi[3],
                   const int periodic[3]);

/* triclinic box: domain->triclinic with the tilt factors xy, xz, yz (Domain::h), fully periodic.  Only the k-space
 * solver works on it - b200md_pppm_setup / b200md_pppm_compute / b200md_pppm_compute_host, i.e. PPPMIntel::compute
 * with its x2lamda / lamda2x bracket (pppm_intel.cpp:151-156, 307-309) and the poisson_ik_triclinic branch (:878-883),
 * ik differentiation, Coulomb grid, one GPU; neighbour lists and pair styles need an orthogonal box.  Positions stay
 * in box coordinates at the boundary.  b200md_set_box returns to an orthogonal box. */
int b200md_set_box_triclinic(b200md_ctx *ctx, const double boxlo[3], const double boxhi[3], double xy, double xz,
                             double yz);

/* ------------------------------------------------------------------------------------------------
 * atoms — replaces IntelBuffers::thr_pack (intel_buffers.h:185-203): one upload, then resident.
 * q may be NULL (atom_style atomic), v may be NULL (zeros).  mass is per type [ntypes+1]. */
int b200md_atoms_upload(b200md_ctx *ctx, int nlocal, int ntypes, const double *x, const double *v,
                        const double *q, const int *type, const double *mass);
/* refresh positions only, host order (the plug-in path: LAMMPS integrates on the host) */
int b200md_atoms_set_x(b200md_ctx *ctx, const double *x);
/* any of x,v,f may be NULL; f is [n][3]; eatom [n] (per-atom energy, f[].w of vec3_acc_t,
 * intel_buffers.h:44) may be NULL */
int b200md_atoms_download(b200md_ctx *ctx, double *x, double *v, double *f, double *eatom);
int b200md_atoms_count(const b200md_ctx *ctx, int *nlocal, int *nghost);
/* multi-GPU: the atoms this rank owns NOW (atoms migrate between the z slabs), device order, with their global ids
 * (id = index in the concatenation of all ranks' uploads, rank 0 first); arrays sized `capacity` atoms; any of
 * ids/x/v/f may be NULL.  Replaces the gather a host code would do over atom->tag. */
int b200md_atoms_download_ids(b200md_ctx *ctx, int capacity, int *n_out, int *ids, double *x, double *v,
                              double *f);

/* ------------------------------------------------------------------------------------------------
 * pair styles — replaces Pair*Intel::init_style + pack_force_const (pair_buck_intel.cpp:367-443,
 * pair_buck_coul_cut_intel.cpp:431-492, pair_buck_coul_long_intel.cpp:457-566,
 * pair_buck_long_coul_long_intel.cpp:542-646).  The caller passes what init_one() produced. */
typedef struct {
  int style;               /* B200MD_PAIR_* */
  int ntypes;
  const double *cutsq;     /* (ntypes+1)^2 each */
  const double *cut_ljsq;  /* cut_bucksq for long/coul/long */
  const double *cut_coulsq;/* coul/cut: per pair; coul/long: global value replicated; buck: NULL */
  const double *buck1, *buck2, *rhoinv, *a, *c, *offset;
  double special_lj[4], special_coul[4]; /* force->special_* ; [0] forced to 1 as in the reference */
  double g_ewald;          /* force->kspace->g_ewald (pair_buck_coul_long_intel.cpp:507) */
  double g_ewald_6;        /* pair_buck_long_coul_long_intel.cpp:267 */
  int ewald_order;         /* long/coul/long: bit1 = ORDER1, bit6 = ORDER6 (:111-112) */
  /* Coulomb tables built by Pair::init_tables on the host (pair_buck_coul_long_intel.cpp:531-542);
   * ncoultablebits = 0 selects the analytic erfc branch (:294-316) */
  int ncoultablebits, ncoulmask, ncoulshiftbits;
  double tabinnersq;
  const double *rtable, *drtable, *ftable, *dftable, *etable, *detable, *ctable, *dctable;
  /* dispersion tables (pair_buck_long_coul_long_intel.cpp:433-454); 0 bits = analytic (:414-431) */
  int ndisptablebits, ndispmask, ndispshiftbits;
  double tabinnerdispsq;
  const double *rdisptable, *drdisptable, *fdisptable, *dfdisptable, *edisptable, *dedisptable;
} b200md_pair_params;

int b200md_pair_setup(b200md_ctx *ctx, const b200md_pair_params *p);

/* special bonds for lists built on the device (molecular systems: the 1-2 / 1-3 / 1-4 partners that stock Neighbor
 * flags in bits 30-31 of a list entry, consumed at pair_buck_coul_long_intel.cpp:283,312 via special_lj / special_coul).
 * nspecial[n][3] holds LAMMPS's cumulative counts (partners [0,n0) are 1-2, [n0,n1) 1-3, [n1,n2) 1-4), special[n][maxspecial]
 * the partners as 0-based upload indices; both in upload order.  Call after b200md_atoms_upload; NULL clears.  Flagged
 * pairs stay in the list whatever their factor (stock drops factor-0 pairs when no k-space style is defined: forces are
 * identical, the pair set is not).  maxspecial <= 32.
 * Several GPUs: collective (every rank calls it with the rows of its own upload and the same maxspecial - a rank that
 * owns no atoms still passes non-NULL arrays, NULL means "clear" on every rank); the partners
 * are GLOBAL ids - rank r's k-th uploaded atom has id (atoms uploaded by ranks < r) + k, what b200md_atoms_download_ids
 * reports.  The rows are gathered into one table indexed by global id on every rank (12 + 4 maxspecial bytes per atom
 * of the whole system), so they follow atoms that migrate between the ranks; halo atoms carry their ids. */
int b200md_atoms_set_special(b200md_ctx *ctx, int maxspecial, const int *nspecial, const int *special);

/* ------------------------------------------------------------------------------------------------
 * neighbour list — replaces the "intel" NeighList request (pair_buck_intel.cpp:370) and
 * buffers->get_cutneighsq() (:399-409): on-device cell binning, periodic ghost atoms, FULL list with
 * newton off.  `neighbor skin bin`, `neigh_modify every/delay/check`. */
int b200md_neigh_setup(b200md_ctx *ctx, double skin, int every, int delay, int check);
/* wrap atoms (Domain::pbc), sort by cell, make ghosts, build the list (neighbor->build, ago = 0) */
int b200md_neigh_build(b200md_ctx *ctx);
/* Neighbor::decide at timestep `ntimestep`; *rebuilt = 1 if a build was done (else ghost positions are
 * refreshed = Comm::forward_comm) */
int b200md_neigh_decide(b200md_ctx *ctx, long ntimestep, int *rebuilt);
/* parity-test access: sizes, then the list in HOST order.  numneigh[nlocal], offsets[nlocal+1],
 * entries[total]: entry = j | special<<30 with j < nlocal an owned atom (host index) or
 * j >= nlocal a ghost; ghost_src[nghost] host index of the atom each ghost images, ghost_shift
 * [nghost][3] its periodic image. */
int b200md_neigh_stats(b200md_ctx *ctx, long *total_entries, int *nghost, int *max_numneigh,
                       long *nbuilds);
int b200md_neigh_download(b200md_ctx *ctx, int *numneigh, long *offsets, int *entries,
                          int *ghost_src, int *ghost_shift);

/* ------------------------------------------------------------------------------------------------
 * Pair*Intel::compute(eflag,vflag) — pair_buck_intel.cpp:48-123 / eval<> :127-365,
 * pair_buck_coul_cut_intel.cpp:134-402, pair_buck_coul_long_intel.cpp:55-453,
 * pair_buck_long_coul_long_intel.cpp:57-539.  Overwrites the device force array (the step's first
 * force contribution; force_clear is fused away).  ev may be NULL when eflag == vflag == 0. */
int b200md_pair_compute(b200md_ctx *ctx, int eflag, int vflag, double ev[8]);

/* The reference eval<> signature itself, host buffers in and out (H2D/D2H inside the call): packed
 * atoms x[nall][3], type[nall], q[nall] (IntelBuffers::get_x/get_q), CSR list numneigh/cnumneigh/
 * firstneigh with special bits (intel_buffers.h:145-146), f[nlocal][4] out (vec3_acc_t).  newton off:
 * the list must be FULL over owned atoms. */
int b200md_pair_eval_host(b200md_ctx *ctx, int eflag, int vflag, int nlocal, int nall,
                          const double *x, const int *type, const double *q, const int *numneigh,
                          const long *cnumneigh, const int *firstneigh, double *f, double ev[8]);

/* ------------------------------------------------------------------------------------------------
 * PPPM — replaces PPPMIntel::init (pppm_intel.cpp:67-98) + PPPM::setup state and
 * PPPMIntel::compute (:104-317): particle_map :326-392, make_rho :403-534, brick2fft :642-672,
 * poisson_ik :811-977 / poisson_ad :986-1054, fieldforce_ik :541-640 / fieldforce_ad :679-804. */
typedef struct {
  int nx, ny, nz;        /* nx_pppm … (2^a 3^b 5^c) */
  int order;             /* <= 7 */
  double g_ewald;
  int differentiation;   /* 0 = ik, 1 = ad (kspace_modify diff) */
  double scale;          /* KSpace::scale, 1.0 */
  int dispersion;        /* 0: Coulomb ('c', charges, function[0]); r^-6 grids of PPPMDispIntel::compute by mixing rule:
                            1: geometric (function[1], 'g', pppm_disp_intel.cpp:245-313 with SURVEY §2.4-2 corrected),
                            2: arithmetic (function[2], seven grids, :315-407), 3: none (function[3], :409-467) */
  const double *B;       /* the array PPPMDisp::init_coeffs builds for that rule, type index 0 unused:
                            1: B[ntypes+1], C_ij = B_i B_j;
                            2: B[7 (ntypes+1)], B[7 i + k] = sqrt(eps_i) / 4 * sqrt(binom(6,k)) * sigma_i^k
                               (C_ij = sum_k B_i[k] B_j[6-k] = 4 eps_ij sigma_ij^6, Lorentz-Berthelot);
                            3: B[(ntypes+1)^2] = C_ij itself, symmetric (split into eigen-components inside);
                            0: NULL */
  double slab_volfactor; /* kspace_modify slab: > 1 = z is non-periodic, the mesh spans zprd * slab_volfactor and
                            PPPM::slabcorr (pppm_intel.cpp:305) is applied; 0 or 1 = off.  Coulomb grid, one GPU */
} b200md_pppm_params;

int b200md_pppm_setup(b200md_ctx *ctx, const b200md_pppm_params *p);
/* forces are ACCUMULATED into the device force array (f +=, pppm_intel.cpp:628-630).
 * energy / virial[6] may be NULL.  eflag bit1 / vflag bit2 ask for the per-atom tallies (b200md_pppm_peratom). */
int b200md_pppm_compute(b200md_ctx *ctx, int eflag, int vflag, double *energy, double virial[6]);
/* Per-atom k-space energy and virial of the last b200md_pppm_compute called with eflag & 2 / vflag & 4: stock
 * PPPM::poisson_peratom + fieldforce_peratom, which the reference reaches through its base class
 * (pppm_intel.cpp:224-229, 876) and the eatom / vatom post-factors of PPPM::compute.  eatom[n], vatom[n][6]
 * (xx,yy,zz,xy,xz,yz) in upload order; either may be NULL.  Coulomb grid, one GPU. */
int b200md_pppm_peratom(b200md_ctx *ctx, double *eatom, double *vatom);
/* host-buffer form of PPPMIntel::compute: x[n][3], q[n] in, f[n][3] += out */
int b200md_pppm_compute_host(b200md_ctx *ctx, int eflag, int vflag, int n, const double *x,
                             const double *q, double *f, double *energy, double virial[6]);
/* parity-test access (host order x-fastest nfft arrays; any may be NULL) */
int b200md_pppm_download(b200md_ctx *ctx, double *density_fft, double *greensfn, double *field_x,
                         double *field_y, double *field_z, double sf_coeff[6]);


/* hand-written 3-D complex FFT used by PPPM, exposed for parity tests (replaces FFT3d::compute,
 * pppm_intel.cpp:835,903): data = nz*ny*nx interleaved re/im doubles on the HOST, transformed in place;
 * dir +1 is exp(+ikx) — what stock FFT3d does for flag=1 (pppm_intel.cpp:835) — and -1 is exp(-ikx);
 * unnormalised. */
int b200md_fft3d_host(b200md_ctx *ctx, double *data, int nx, int ny, int nz, int dir);

/* ------------------------------------------------------------------------------------------------
 * fix nve/intel — FixNVEIntel::initial_integrate / final_integrate / reset_dt
 * (fix_nve_intel.cpp:60-99, 103-127, 129-194); group all and per-type mass unless b200md_nve_set_group says otherwise. */
int b200md_nve_setup(b200md_ctx *ctx, double dt);
/* fix nve/intel on a sub-group and / or with per-atom masses: the `mask[i] & groupbit` and `atom->rmass` branches of
 * FixNVEIntel::reset_dt (fix_nve_intel.cpp:147-190: _dtfm = 0 outside the group) and the `_dtfm[i] != 0.0` branch of
 * initial_integrate (:88-97: atoms outside the group keep x and v).  ingroup[n] (0 / non-zero) and rmass[n] are in
 * upload order; either may be NULL (group all / per-type mass).  Call after b200md_atoms_upload and before
 * b200md_nve_setup.  Several GPUs: collective, every rank passes the rows of its own upload (the same arrays NULL /
 * non-NULL on every rank); they are gathered into tables indexed by global id. */
int b200md_nve_set_group(b200md_ctx *ctx, const int *ingroup, const double *rmass);
int b200md_nve_initial_integrate(b200md_ctx *ctx);
int b200md_nve_final_integrate(b200md_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * Verlet::run on the device: nsteps of initial_integrate -> decide/build|forward ->
 * pair -> kspace -> final_integrate with atoms resident.  Thermo-style energies are evaluated on the
 * last step when thermo != NULL: thermo[0..7] = pair ev, [8] = kspace energy, [9..14] kspace virial,
 * [15] = kinetic energy (sum 1/2 m v^2, mass units). */
int b200md_run(b200md_ctx *ctx, long nsteps, double *thermo /*16 or NULL*/);
/* same, bracketed by CUDA events on the library's stream: *elapsed_ms is device time of the nsteps */
int b200md_run_timed(b200md_ctx *ctx, long nsteps, double *thermo, double *elapsed_ms);
/* one timestep through HOST buffers (the plug-in deployment where LAMMPS keeps atom->x/f on the host):
 * H2D of x_in[n][3] (NULL: keep device positions), one step, D2H of x_out and f_out ([n][3], may be NULL).
 * Pass pinned memory for full PCIe speed. */
int b200md_step_host(b200md_ctx *ctx, const double *x_in, double *x_out, double *f_out);
/* the same on any number of GPUs: atoms migrate between the ranks, so results come back in DEVICE order with the atoms'
 * global ids (first id of the rank's upload + upload index).  Positions stay resident (no upload); *n_out = atoms owned
 * after this step's migration (<= capacity, else an error).  ids / x_out / f_out may be NULL. */
int b200md_step_host_ids(b200md_ctx *ctx, int capacity, int *n_out, int *ids, double *x_out /*[cap][3]*/,
                         double *f_out /*[cap][3]*/);
/* forces only (setup phase of a run: build + pair + kspace), energies in thermo like b200md_run */
int b200md_setup_forces(b200md_ctx *ctx, int eflag, int vflag, double *thermo);

/* per-phase device timers (the HPAC_TIMING / fix->start_watch hooks, pppm_intel.cpp:113-123,
 * pair_buck_intel.cpp:80,356-359): accumulated CUDA-event ms since the last reset.
 * names: see b200md_timer_name(i), i < b200md_timer_count(). */
int b200md_timers_enable(b200md_ctx *ctx, int on);
int b200md_timers_get(b200md_ctx *ctx, double *ms, long *calls, int n);
int b200md_timers_reset(b200md_ctx *ctx);
int b200md_timer_count(void);
const char *b200md_timer_name(int i);

/* number of kernels this library launched since ctx creation (bench.py's gpu_launches) */
long b200md_launch_count(const b200md_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * multi-GPU: one process per GPU, z-slab spatial decomposition of atoms and of the PPPM
 * grid.  Replaces Comm::borders/forward_comm (atom halo), GridComm, Remap and FFT3d transposes, and the
 * two MPI_Allreduce calls of pppm_intel.cpp:260,273.  nccl_unique_id is the 128-byte ncclUniqueId
 * produced by b200md_comm_unique_id on rank 0 and broadcast by the host (torch.distributed / MPI). */
/* The grid plan of the slab decomposition, pure host arithmetic (usable without a device): for every rank r the
 * owned FFT planes [pzlo,pzhi), the local brick (origin zoff, nbz planes: owned planes + stencil + skin/2 halo =
 * nzlo_out..nzhi_out of PPPM::set_grid_local for that rank's atom slab) and the y rows [ylo,yhi) it holds after the
 * transpose to z pencils.  Arrays are [nranks].  Non-zero return: a halo would reach beyond the neighbouring rank. */
int b200md_pppm_decomp(int nranks, int nz, int ny, int order, double skin, double prd_z, int *pzlo, int *pzhi,
                       int *zoff, int *nbz, int *ylo, int *yhi);
/* Call order: b200md_comm_init and b200md_neigh_setup come BEFORE b200md_pppm_setup.  b200md_comm_init drops any
 * k-space state (it holds the rank count and the slab plan), b200md_neigh_setup drops it when the new skin is larger
 * than the one its brick halo (skin/2) was sized for: a later b200md_pppm_compute then fails with "pppm compute before
 * b200md_pppm_setup" instead of solving on a stale decomposition. */
int b200md_comm_unique_id(void *id128);
int b200md_comm_init(b200md_ctx *ctx, int rank, int nranks, const void *id128);
int b200md_comm_finalize(b200md_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
