// internal.h — device context shared by the kernels behind include/b200md.h.
//
// HBM layout (all arrays device-resident, capacity-grown, never shrunk):
//   xq   double4[nall]   {x,y,z,q}      owned atoms [0,nlocal) sorted by neighbour bin, then ghosts
//                                       [nlocal,nall) sorted by bin — one 32 B sector per gathered j
//   xqf  float4[nall]    same, float    mixed mode only (IntelBuffers<float,double>::atom_t, q folded in)
//   type int[nall]
//   v    double4[nlocal] {vx,vy,vz,dtf/m}
//   f    double4[nlocal] {fx,fy,fz,eatom}   (vec3_acc_t, intel_buffers.h:44)
//   tag  int[nlocal]     host index of each sorted atom
//   neighbour list: numneigh int[nlocal], offsets int64[nlocal+1], entries int[total] (CSR, rows in
//   stencil order — the firstneigh/cnumneigh/numneigh triple of intel_buffers.h:145-146)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200md.h"
#include "../../include/b200md_testing.h"

#define B2_SBBITS 30
#define B2_NEIGHMASK 0x3FFFFFFF
// device-built lists with nall <= 2^26 carry the type of j in bits 26..29 of the entry (the pair kernel then needs no
// type[j] gather); such lists never have special-bond bits.  Exports strip the type bits.
#define B2_TYPESHIFT 26
#define B2_IDXMASK26 0x03FFFFFF
#define B2_MAXTYPES 8       // (ntypes+1) <= 9 rows staged in shared memory
#define B2_MAXORDER 7

// phase timers first (they partition the step); the K_* entries time single kernels INSIDE those phases (nested: do not
// add them to the phase sum) so that bench.py can report a roofline per kernel from the same run
enum TimerId {
  T_NEIGH = 0, T_COMM, T_PAIR, T_MAKE_RHO, T_FFT, T_POISSON, T_FIELDFORCE, T_NVE, T_OTHER,
  K_NB_MASK, K_NB_FILL, K_RHO_TILES, K_RHO_FOLD, K_FFT_X_FWD, K_FFT_Y_FWD, K_FFT_Z_POISSON, K_FFT_Y_INV, K_FFT_X_INV,
  K_NVE_INITIAL, K_NVE_FINAL, T_COUNT
};
#define T_PHASE_COUNT (T_OTHER + 1)

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  int reserve(size_t n, bool keep = false, cudaStream_t s = 0) {
    if (n <= cap) return 0;
    size_t ncap = n + n / 8 + 64;
    T *np = nullptr;
    cudaError_t e = cudaMalloc((void **)&np, ncap * sizeof(T));
    if (e != cudaSuccess) return -1;
    if (keep && p && cap) cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
    if (p) { cudaStreamSynchronize(s); cudaFree(p); }
    p = np;
    cap = ncap;
    return 0;
  }
  void free_() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PairCoeff {  // one (itype,jtype) entry, flt_t-converted on the host at setup
  double cutsq, cut_ljsq, cut_coulsq, buck1, buck2, rhoinv, a, c, offset;
};

struct PairState {
  bool ready = false;
  b200md_pair_params p{};     // scalars only; pointers are not kept
  int tp1 = 0;
  DevBuf<double> coeff_d;     // [tp1*tp1][9] doubles
  DevBuf<float> coeff_f;      // same in float (mixed)
  DevBuf<double> cutneighsq;  // [tp1*tp1] (flt_t-rounded values stored as double)
  std::vector<double> h_cutsq;
  double cutmax = 0.0;        // max cut over type pairs
  bool same_cut = false;      // cut_ljsq == cutsq for all type pairs
  // tables: 8 arrays of 2^bits (double and float copies)
  DevBuf<double> ctab_d;      // [ntable][8] {r,dr,f,df,e,de,c,dc}
  DevBuf<float> ctab_f;
  DevBuf<double> dtab_d;      // dispersion [ntable][6] {r,dr,f,df,e,de}
  DevBuf<float> dtab_f;
  DevBuf<double> exptab;      // 2^(k/64), k < 64: table of the double-precision exp kernel (pair_kernel.cuh)
};

struct NeighState {
  bool ready = false;
  double skin = 0.3;
  int every = 1, delay = 0, check = 1;
  long last_build = -1, nbuilds = 0;
  double cutneighmax = 0, cutghost = 0;
  // bins
  int nbin[3] = {0, 0, 0}, mshell[3] = {0, 0, 0}, mbin[3] = {0, 0, 0};
  double bininv[3] = {0, 0, 0};
  long nbins_tot = 0;
  DevBuf<int> bin_of, bin_sorted;  // per atom: bin id before / after the sort
  DevBuf<int> bin_count, bin_start, bin_end, bin_cursor;
  DevBuf<int> perm;           // scratch permutation
  DevBuf<int> ghost_src;      // [nghost] owned index (sorted order)
  DevBuf<int> ghost_shift;    // [nghost] packed (sx+1) + 3*(sy+1) + 9*(sz+1)
  DevBuf<int> ghost_cnt, goff, gsrc_tmp, gshift_tmp, gbin, gperm;  // scratch
  int ago = 0;                // steps since the last build (Neighbor::ago)
  DevBuf<int> numneigh;
  DevBuf<long long> offsets;
  DevBuf<int> entries;
  // multi-GPU (z slabs): halo = owned atoms within cutghost of a slab face, sent to the neighbour rank every step
  DevBuf<int> halo_idx;           // [ns_lo | ns_hi] owned indices sent to the lower / upper neighbour
  int ns_lo = 0, ns_hi = 0, nr_lo = 0, nr_hi = 0;   // sent to lower/upper, received from lower/upper
  DevBuf<double4> halo_sbuf, halo_rbuf;             // rbuf = [from lower | from upper]
  DevBuf<int> halo_stype, halo_rtype;
  DevBuf<int> halo_stag, halo_rtag;   // global ids of the halo atoms (exchanged at a rebuild when special bonds are set)
  DevBuf<int> mig_flag, mig_off;                    // migration scratch: 3 flag / 3 offset arrays
  DevBuf<unsigned char> mig_send, mig_recv;
  double slab_lo = 0, slab_hi = 0;                  // this rank's z slab
  DevBuf<int> mask_words;         // build scratch: mask words per bin
  DevBuf<long long> mask_off;     // their exclusive scan
  DevBuf<unsigned> maskbuf;       // hit masks [bin][atom][candidate word]
  long long total_entries = 0;
  int pitch = 0;                  // entries per list row (multiple of 32: rows start on 128 B lines); 0 = packed CSR rows
  int max_numneigh = 0;
  bool packed_type = false;   // entries carry type(j) << B2_TYPESHIFT
  DevBuf<double4> xhold;      // positions at last build (owned)
  DevBuf<int> flags;          // device flags: [0] displacement trigger, [1] errors
  DevBuf<unsigned char> scan_ws;
  // scratch for permuting
  DevBuf<double4> tmp4a, tmp4b;
  DevBuf<int> tmpi_a, tmpi_b;
};

struct PppmState;  // pppm.cu
struct CommState;  // comm.cu

struct b200md_ctx {
  int device = 0;
  int prec = B200MD_PREC_DOUBLE;
  cudaStream_t stream = nullptr;
  // special bonds, indexed by global id (= upload index on one GPU): [N][3] cumulative counts, [N][sp_max] partner ids
  DevBuf<int> sp_count, sp_list;
  int sp_max = 0;                  // 0: atomic system, lists carry no special bits
  DevBuf<int> nve_group;       // fix nve on a sub-group: 0 / 1 per atom, by global id (b200md_nve_set_group)
  DevBuf<double> nve_rmass;    // per-atom masses, by global id
  bool nve_grouped = false, nve_has_rmass = false;
  cudaStream_t copy_stream = nullptr;   // device->host copies that overlap the force kernels (b200md_step_host)
  cudaEvent_t ev_copy = nullptr;
  // k-space overlap: particle_map .. poisson of PPPM run on `kstream` (higher priority) underneath the FP64-bound pair
  // kernel; `ev_pre` (recorded on `stream` just before the pair launch) orders them after the position update,
  // `ev_k` hands the fields back to `stream` for fieldforce (which must follow the pair kernel: it accumulates into f)
  cudaStream_t kstream = nullptr, main_stream = nullptr;
  cudaEvent_t ev_pre = nullptr, ev_k = nullptr;
  bool overlap = false, ev_pre_valid = false;   // opt-in (B200MD_OVERLAP=1): see ctx.cu
  std::string err;
  int sm_count = 148;

  double qqrd2e = 1.0, ftm2v = 1.0;
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {1, 1, 1}, prd[3] = {1, 1, 1};
  int periodic[3] = {1, 1, 1};
  // triclinic box (b200md_set_box_triclinic): tilt factors xy, xz, yz.  Only the k-space solver works on it.
  bool triclinic = false;
  double tilt[3] = {0, 0, 0};
  bool box_set = false;

  int nlocal = 0, nghost = 0, ntypes = 0;
  bool has_q = false;
  std::vector<double> mass;
  DevBuf<double4> xq, v, f;
  DevBuf<float4> xqf;
  DevBuf<int> type, tag, inv_tag;   // tag: host index of each atom (multi-GPU: global id = first_id + host index)
  int first_id = 0;                 // multi-GPU: global id of this rank's first uploaded atom
  DevBuf<double> stage;       // staging for host transfers
  double *h_pinned = nullptr; // small pinned scratch (ev partial results, flags)
  size_t h_pinned_bytes = 0;

  PairState pair;
  NeighState neigh;
  PppmState *pppm = nullptr;    // Coulomb grid (pppm/intel; function[0] of pppm/disp/intel)
  PppmState *pppm6 = nullptr;   // geometric-mixing dispersion grid (function[1] of pppm/disp/intel)
  CommState *comm = nullptr;

  // nve
  double dt = 0.0, dtv = 0.0, dtf = 0.0;
  bool nve_ready = false;
  long ntimestep = 0;

  // ev reduction scratch
  DevBuf<double> ev_partial;  // [nblocks][8]
  DevBuf<double> ev_out;      // [32]

  // timers: event pairs are recorded on the launching stream without synchronising; elapsed times are
  // harvested when the caller asks (b200md_timers_get), so a timed region is not perturbed
  bool timers_on = false;
  double t_ms[T_COUNT] = {0};
  long t_calls[T_COUNT] = {0};
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  struct TimerRec { int id; cudaEvent_t a, b; };
  std::vector<TimerRec> t_pending;
  std::vector<cudaEvent_t> t_pool;
  long launches = 0;
  cudaEvent_t timer_event() {
    if (!t_pool.empty()) { cudaEvent_t e = t_pool.back(); t_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

// ---- error helpers ---------------------------------------------------------------------------
int b2_fail(b200md_ctx *ctx, int code, const char *fmt, ...);
#define CUDA_OK(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return b2_fail(ctx, B200MD_ECUDA, "%s failed: %s (%s:%d)", #call,                     \
                     cudaGetErrorString(e__), __FILE__, __LINE__);                          \
  } while (0)
#define KERNEL_OK(ctx, name)                                                                \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return b2_fail(ctx, B200MD_ECUDA, "launch of %s failed: %s", name,                    \
                     cudaGetErrorString(e__));                                              \
  } while (0)
#define RESERVE(ctx, buf, n)                                                                \
  do {                                                                                      \
    if ((buf).reserve((size_t)(n), false, (ctx)->stream))                                   \
      return b2_fail(ctx, B200MD_ENOMEM, "out of device memory reserving %s (%zu elements)", \
                     #buf, (size_t)(n));                                                    \
  } while (0)
#define RESERVE_KEEP(ctx, buf, n)                                                           \
  do {                                                                                      \
    if ((buf).reserve((size_t)(n), true, (ctx)->stream))                                    \
      return b2_fail(ctx, B200MD_ENOMEM, "out of device memory reserving %s (%zu elements)", \
                     #buf, (size_t)(n));                                                    \
  } while (0)
#define TRY(expr)                    \
  do {                               \
    int rc__ = (expr);               \
    if (rc__ != 0) return rc__;      \
  } while (0)

struct ScopedTimer {
  b200md_ctx *c;
  int id;
  cudaEvent_t a = nullptr;
  ScopedTimer(b200md_ctx *ctx, int id_, bool enabled = true) : c(ctx), id(id_) {
    if (c->timers_on && enabled) {
      a = c->timer_event();
      cudaEventRecord(a, c->stream);
    }
  }
  ~ScopedTimer() {
    if (a) {
      cudaEvent_t b = c->timer_event();
      cudaEventRecord(b, c->stream);
      c->t_pending.push_back({id, a, b});
    }
  }
};

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- cross-file internals ----------------------------------------------------------------------
// scan.cu
size_t b2_scan_ws_bytes(size_t n);
int b2_exclusive_scan_i32(b200md_ctx *ctx, const int *in, int *out, size_t n, void *ws);            // out[n] = total
int b2_exclusive_scan_i32_i64(b200md_ctx *ctx, const int *in, long long *out, size_t n, void *ws);  // out[n] = total
// atoms.cu
int b2_refresh_float_copy(b200md_ctx *ctx, int first, int count);
// neigh.cu
int b2_ghost_refresh(b200md_ctx *ctx);
int b2_neigh_build(b200md_ctx *ctx);
int b2_neigh_check_trigger(b200md_ctx *ctx, int *trigger);
// pair.cu
int b2_pair_compute(b200md_ctx *ctx, int eflag, int vflag, double *ev);
// pppm.cu
int b2_pppm_compute(b200md_ctx *ctx, int eflag, int vflag, double *energy, double *virial);
void b2_pppm_free(b200md_ctx *ctx);
void b2_pppm_skin_changed(b200md_ctx *ctx, double skin);
// ctx.cu
int b2_unpack_to_stage(b200md_ctx *ctx, const double4 *src, double *stage3);
// nve.cu
int b2_nve_initial(b200md_ctx *ctx);
int b2_nve_final(b200md_ctx *ctx);
int b2_kinetic_energy(b200md_ctx *ctx, double *ke);
// comm.cu
void b2_comm_free(b200md_ctx *ctx);
int b2_comm_nranks(const b200md_ctx *ctx);
int b2_comm_rank(const b200md_ctx *ctx);
int b2_comm_exchange(b200md_ctx *ctx, const void *s_lo, size_t n_lo, const void *s_hi, size_t n_hi, void *r_from_hi,
                     size_t n_from_hi, void *r_from_lo, size_t n_from_lo);
int b2_comm_exchange_counts(b200md_ctx *ctx, int n_lo, int n_hi, int *n_from_hi, int *n_from_lo);
int b2_comm_allreduce_sum(b200md_ctx *ctx, double *dev, int n);
int b2_comm_allreduce_max_int(b200md_ctx *ctx, int *dev, int n);
int b2_comm_allgather_int(b200md_ctx *ctx, int value, int *host_out /*[nranks]*/);
int b2_comm_allgatherv(b200md_ctx *ctx, const void *send, void *recv, const size_t *bytes, const size_t *disp);

// grouped point-to-point calls (one NCCL group): several messages per peer are matched in issue order
struct CommGroup {
  b200md_ctx *ctx;
  bool open = false;
  explicit CommGroup(b200md_ctx *c);
  ~CommGroup();
  int send(const void *buf, size_t bytes, int peer);
  int recv(void *buf, size_t bytes, int peer);
  int end();
};
int b2_comm_alltoallv(b200md_ctx *ctx, const void *sbuf, const size_t *scount, const size_t *sdisp, void *rbuf,
                      const size_t *rcount, const size_t *rdisp);
// peer memory over NVLink / NVSwitch: a buffer of the same size on every rank, mapped into every other rank's address
// space (CUDA IPC), so that a kernel can store straight into its peers' copies — the FFT transposes of pppm.cu write
// their blocks where the next pass reads them instead of packing, all-to-all-ing and unpacking.
struct PeerBuf {
  void *local = nullptr;
  size_t bytes = 0;
  void *peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // peer[me] == local
};
// collective over all ranks; *ok = 0 when peer access is unavailable on any rank (nothing stays allocated then)
int b2_comm_peer_alloc(b200md_ctx *ctx, PeerBuf &pb, size_t bytes, int *ok);
void b2_comm_peer_free(b200md_ctx *ctx, PeerBuf &pb);
// stream-ordered barrier over the ranks: returns (on the stream) once every rank has reached it
int b2_comm_barrier(b200md_ctx *ctx);

// Per-atom host array of this rank's upload (`per` elements per atom, upload order) -> device table of ALL atoms
// indexed by global id (= upload index on one GPU).  Tables indexed this way stay valid when atoms migrate between the
// ranks.  Collective on several GPUs (every rank calls it, also with no atoms); *nglobal receives the atom count.
template <class T>
inline int b2_atoms_global_table(b200md_ctx *ctx, const T *host_rows, size_t per, DevBuf<T> &table, long *nglobal) {
  const int P = b2_comm_nranks(ctx), me = b2_comm_rank(ctx);
  std::vector<int> counts(P, ctx->nlocal);
  if (P > 1) TRY(b2_comm_allgather_int(ctx, ctx->nlocal, counts.data()));
  std::vector<size_t> bytes(P), disp(P);
  size_t tot = 0;
  for (int r = 0; r < P; r++) {
    disp[r] = tot * per * sizeof(T);
    bytes[r] = (size_t)counts[r] * per * sizeof(T);
    tot += (size_t)counts[r];
  }
  RESERVE(ctx, table, tot * per + 1);
  if (ctx->nlocal > 0)
    CUDA_OK(ctx, cudaMemcpyAsync((char *)table.p + disp[me], host_rows, bytes[me], cudaMemcpyHostToDevice, ctx->stream));
  if (P > 1) TRY(b2_comm_allgatherv(ctx, (char *)table.p + disp[me], table.p, bytes.data(), disp.data()));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (nglobal) *nglobal = (long)tot;
  return 0;
}
