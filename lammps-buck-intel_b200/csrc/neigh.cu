// neigh.cu — on-device cell binning, periodic ghost atoms and FULL neighbour-list build (newton off).
//
// Replaces what the reference consumes but does not ship: Domain::pbc,
// Comm::borders / forward_comm on one rank, Neighbor::setup_bins / decide / check_distance, and the
// "intel" packed list requested at pair_buck_intel.cpp:370 and read at :142-144,221-222,246-247.
// Criterion: rsq <= cutneighsq[itype][jtype], evaluated in flt_t with the same un-fused rounding as the
// reference's AVX build ((dx*dx + dy*dy) + dz*dz), so the pair set is bit-exact against the oracle.
//
// Design (B200): atoms are counting-sorted by bin (bin edge = cutneighmax/2, bins tile the box exactly so a
// periodic shift is a whole number of bins); owned atoms come first, ghosts after, both in bin order, so
// the five x-adjacent stencil bins of a (y,z) row are ONE contiguous index range.  One block per bin: distances are
// evaluated once into hit bit-masks (k_nb_mask), then one warp per atom expands its masks into the row (k_nb_fill):
// 25 rows x {owned range, ghost range} in a fixed order => deterministic list; rows sit on a fixed pitch so each starts
// on a 128 B line.  All counters are integers (no FP atomics anywhere).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <vector>

#include "internal.h"

namespace {

struct BinGeom {
  double lo[3], hi[3], prd[3], bininv[3];   // the box this rank bins: whole box, or its z slab (multi-GPU)
  double wlo[3], whi[3], wprd[3];           // the global periodic box (Domain::pbc and image shifts)
  int nbin[3], m[3], mbin[3], s[3], periodic[3];
  int img[3];                               // 1: periodic images of this dimension are made locally
  double cutghost;
};

__device__ __forceinline__ int bin_coord(double x, double lo, double bininv, int nbin) {
  int b = (int)floor((x - lo) * bininv);
  return min(max(b, 0), nbin - 1);
}

__global__ void k_wrap_bin(int n, double4 *__restrict__ xq, BinGeom g, int do_wrap, int *__restrict__ bin_of,
                           int *__restrict__ bin_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 p = xq[i];
  double c[3] = {p.x, p.y, p.z};
  if (do_wrap) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (g.periodic[d]) {  // Domain::pbc
        if (c[d] < g.wlo[d]) c[d] += g.wprd[d];
        if (c[d] >= g.whi[d]) {
          c[d] -= g.wprd[d];
          c[d] = fmax(c[d], g.wlo[d]);
        }
      }
    }
    p.x = c[0]; p.y = c[1]; p.z = c[2];
    xq[i] = p;
  }
  if (!bin_of) return;
  const int bx = bin_coord(c[0], g.lo[0], g.bininv[0], g.nbin[0]) + g.m[0];
  const int by = bin_coord(c[1], g.lo[1], g.bininv[1], g.nbin[1]) + g.m[1];
  const int bz = bin_coord(c[2], g.lo[2], g.bininv[2], g.nbin[2]) + g.m[2];
  const int b = (bz * g.mbin[1] + by) * g.mbin[0] + bx;
  bin_of[i] = b;
  atomicAdd(&bin_count[b], 1);
}

__global__ void k_scatter(int n, const int *__restrict__ bin_of, int *__restrict__ cursor, int *__restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = atomicAdd(&cursor[bin_of[i]], 1);
  perm[pos] = i;
}

// restores a deterministic (ascending previous index) order inside every bin after the atomic scatter
__global__ void k_sort_bins(long nbins, const int *__restrict__ start, int *__restrict__ perm) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const int s = start[b], e = start[b + 1];
  for (int a = s + 1; a < e; a++) {
    const int key = perm[a];
    int k = a - 1;
    while (k >= s && perm[k] > key) {
      perm[k + 1] = perm[k];
      k--;
    }
    perm[k + 1] = key;
  }
}

__global__ void k_permute_atoms(int n, const int *__restrict__ perm, const double4 *__restrict__ xq_in,
                                const double4 *__restrict__ v_in, const int *__restrict__ type_in,
                                const int *__restrict__ tag_in, const int *__restrict__ bin_in,
                                double4 *__restrict__ xq_out, double4 *__restrict__ v_out,
                                int *__restrict__ type_out, int *__restrict__ tag_out, int *__restrict__ bin_out,
                                double4 *__restrict__ xhold) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = perm[k];
  const double4 p = xq_in[i];
  xq_out[k] = p;
  xhold[k] = p;
  v_out[k] = v_in[i];
  type_out[k] = type_in[i];
  tag_out[k] = tag_in[i];
  bin_out[k] = bin_in[i];
}

__device__ __forceinline__ void ghost_flags(const double4 p, const BinGeom &g, int lo[3], int hi[3]) {
  const double c[3] = {p.x, p.y, p.z};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    // Comm::borders slabs, inclusive on both ends: [lo, lo+cutghost] -> +prd, [hi-cutghost, hi] -> -prd
    lo[d] = g.img[d] && c[d] <= g.lo[d] + g.cutghost;
    hi[d] = g.img[d] && c[d] >= g.hi[d] - g.cutghost;
  }
}

// base atom b: owned (b < nlocal, xq[b]) or z-halo atom received from a neighbour rank (rbuf[b - nlocal])
__device__ __forceinline__ double4 base_pos(int b, int nlocal, const double4 *xq, const double4 *rbuf) {
  return b < nlocal ? xq[b] : rbuf[b - nlocal];
}

__global__ void k_ghost_count(int nlocal, int nbase, const double4 *__restrict__ xq, const double4 *__restrict__ rbuf,
                              BinGeom g, int *__restrict__ cnt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbase) return;
  int lo[3], hi[3];
  ghost_flags(base_pos(b, nlocal, xq, rbuf), g, lo, hi);
  // a halo atom is itself a ghost (shift 0) in addition to its images
  cnt[b] = (1 + lo[0] + hi[0]) * (1 + lo[1] + hi[1]) * (1 + lo[2] + hi[2]) - (b < nlocal ? 1 : 0);
}

__global__ void k_ghost_fill(int nlocal, int nbase, const double4 *__restrict__ xq, const double4 *__restrict__ rbuf,
                             const int *__restrict__ bin_sorted, BinGeom g, const int *__restrict__ goff,
                             int *__restrict__ gsrc, int *__restrict__ gshift, int *__restrict__ gbin,
                             int *__restrict__ gbin_count) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbase) return;
  int lo[3], hi[3];
  const double4 p = base_pos(b, nlocal, xq, rbuf);
  ghost_flags(p, g, lo, hi);
  int w = goff[b];
  int ex, ey, ez;
  if (b < nlocal) {
    const int bb = bin_sorted[b];
    ex = bb % g.mbin[0]; ey = (bb / g.mbin[0]) % g.mbin[1]; ez = bb / (g.mbin[0] * g.mbin[1]);
  } else {  // halo atom: bin from its coordinates; it may sit in the z shell
    ex = bin_coord(p.x, g.lo[0], g.bininv[0], g.nbin[0]) + g.m[0];
    ey = bin_coord(p.y, g.lo[1], g.bininv[1], g.nbin[1]) + g.m[1];
    const int bz = (int)floor((p.z - g.lo[2]) * g.bininv[2]);
    ez = min(max(bz + g.m[2], 0), g.mbin[2] - 1);
  }
  const int src = b < nlocal ? b : -1 - (b - nlocal);
  for (int sz = -1; sz <= 1; sz++) {
    if ((sz == 1 && !lo[2]) || (sz == -1 && !hi[2])) continue;
    for (int sy = -1; sy <= 1; sy++) {
      if ((sy == 1 && !lo[1]) || (sy == -1 && !hi[1])) continue;
      for (int sx = -1; sx <= 1; sx++) {
        if ((sx == 1 && !lo[0]) || (sx == -1 && !hi[0])) continue;
        if (!sx && !sy && !sz && b < nlocal) continue;
        const int gx = min(max(ex + sx * g.nbin[0], 0), g.mbin[0] - 1);
        const int gy = min(max(ey + sy * g.nbin[1], 0), g.mbin[1] - 1);
        const int gz = min(max(ez + sz * g.nbin[2], 0), g.mbin[2] - 1);
        const int gb = (gz * g.mbin[1] + gy) * g.mbin[0] + gx;
        gsrc[w] = src;
        gshift[w] = (sx + 1) + 3 * (sy + 1) + 9 * (sz + 1);
        gbin[w] = gb;
        atomicAdd(&gbin_count[gb], 1);
        w++;
      }
    }
  }
}

__global__ void k_ghost_permute(int ng, const int *__restrict__ perm, const int *__restrict__ src_in,
                                const int *__restrict__ shift_in, int *__restrict__ src_out,
                                int *__restrict__ shift_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ng) return;
  const int i = perm[k];
  src_out[k] = src_in[i];
  shift_out[k] = shift_in[i];
}

// Comm::forward_comm: ghost = source + image shift; the source is an owned atom (src >= 0) or a z-halo atom just
// received from a neighbour rank (src = -1 - k, position rbuf[k]).  set_type at build time only.
__global__ void k_ghost_refresh(int nlocal, int ng, const int *__restrict__ src, const int *__restrict__ shift,
                                double px, double py, double pz, double4 *__restrict__ xq, float4 *__restrict__ xqf,
                                int *__restrict__ type, int set_type, const double4 *__restrict__ rbuf,
                                const int *__restrict__ rtype) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ng) return;
  const int s = src[k];
  const int code = shift[k];
  double4 p = s >= 0 ? xq[s] : rbuf[-1 - s];
  const int sx = code % 3 - 1, sy = (code / 3) % 3 - 1, sz = code / 9 - 1;
  if (sx) p.x = p.x + sx * px;
  if (sy) p.y = p.y + sy * py;
  if (sz) p.z = p.z + sz * pz;
  xq[nlocal + k] = p;
  if (xqf) xqf[nlocal + k] = make_float4((float)p.x, (float)p.y, (float)p.z, (float)p.w);
  if (set_type) type[nlocal + k] = s >= 0 ? type[s] : rtype[-1 - s];
}

// ---- multi-GPU: migration (Comm::exchange) and z-halo (Comm::borders / forward_comm) ------------------------
struct MigAtom {
  double4 xq, v;
  int type, tag, pad0, pad1;
};

// dest: 0 stay, 1 to the lower rank, 2 to the upper rank (positions are already wrapped into the global box)
__global__ void k_mig_dest(int n, const double4 *__restrict__ xq, double zlo0, double slab_inv, int rank, int nranks,
                           int *__restrict__ f_stay, int *__restrict__ f_lo, int *__restrict__ f_hi,
                           int *__restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int sidx = (int)floor((xq[i].z - zlo0) * slab_inv);
  sidx = min(max(sidx, 0), nranks - 1);
  const int lower = (rank + nranks - 1) % nranks, upper = (rank + 1) % nranks;
  int d = 0;
  if (sidx != rank) {
    if (sidx == lower) d = 1;
    else if (sidx == upper) d = 2;
    else { d = 0; *err = 1; }   // moved further than one slab between two rebuilds
  }
  f_stay[i] = d == 0;
  f_lo[i] = d == 1;
  f_hi[i] = d == 2;
}

__global__ void k_mig_pack(int n, const int *__restrict__ f_stay, const int *__restrict__ f_lo,
                           const int *__restrict__ o_stay, const int *__restrict__ o_lo, const int *__restrict__ o_hi,
                           const double4 *__restrict__ xq, const double4 *__restrict__ v, const int *__restrict__ type,
                           const int *__restrict__ tag, double4 *__restrict__ xq_out, double4 *__restrict__ v_out,
                           int *__restrict__ type_out, int *__restrict__ tag_out, MigAtom *__restrict__ m_lo,
                           MigAtom *__restrict__ m_hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (f_stay[i]) {
    const int k = o_stay[i];
    xq_out[k] = xq[i]; v_out[k] = v[i]; type_out[k] = type[i]; tag_out[k] = tag[i];
  } else {
    MigAtom a;
    a.xq = xq[i]; a.v = v[i]; a.type = type[i]; a.tag = tag[i]; a.pad0 = a.pad1 = 0;
    if (f_lo[i]) m_lo[o_lo[i]] = a;
    else m_hi[o_hi[i]] = a;
  }
}

__global__ void k_mig_unpack(int n, const MigAtom *__restrict__ m, int first, double4 *__restrict__ xq,
                             double4 *__restrict__ v, int *__restrict__ type, int *__restrict__ tag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const MigAtom a = m[k];
  xq[first + k] = a.xq; v[first + k] = a.v; type[first + k] = a.type; tag[first + k] = a.tag;
}

__global__ void k_halo_flag(int n, const double4 *__restrict__ xq, double zlo_cut, double zhi_cut,
                            int *__restrict__ f_lo, int *__restrict__ f_hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double z = xq[i].z;
  f_lo[i] = z <= zlo_cut;   // inclusive slabs, like Comm::borders
  f_hi[i] = z >= zhi_cut;
}

__global__ void k_compact_idx(int n, const int *__restrict__ flag, const int *__restrict__ off, int *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[off[i]] = i;
}

__global__ void k_gather_int(int n, const int *__restrict__ idx, const int *__restrict__ src, int *__restrict__ dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[idx[k]];
}

__global__ void k_halo_pack(int ns, const int *__restrict__ idx, const double4 *__restrict__ xq, double zshift,
                            double4 *__restrict__ sbuf, const int *__restrict__ type, int *__restrict__ stype) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ns) return;
  double4 p = xq[idx[k]];
  p.z += zshift;
  sbuf[k] = p;
  if (stype) stype[k] = type[idx[k]];
}

template <class flt_t>
struct Pos;
template <>
struct Pos<double> {
  typedef double4 vec;
  static __device__ __forceinline__ double rsq(const double4 a, const double4 b) {
    const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  }
};
template <>
struct Pos<float> {
  typedef float4 vec;
  static __device__ __forceinline__ float rsq(const float4 a, const float4 b) {
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  }
};

// ---- full list build: distances ONCE, as hit bit-masks; the fill pass is pure data movement ---------------------
// One BLOCK per interior bin.  The candidates of the bin's stencil are 25 (y,z) rows x {owned range, ghost range},
// each one contiguous index range, enumerated c = 0..ncand-1 in a fixed order (row order, owned then ghost, ascending
// index => deterministic rows).
//   k_nb_ranges  per bin: ncand, mask words per atom, atoms -> sizes for the mask buffer (scanned on the device)
//   k_nb_mask    a warp takes 32 candidates (one per lane, held in registers) and loops over the bin's atoms, whose
//                coordinates are broadcast from shared memory: one ballot per (atom, 32 candidates) is the hit mask,
//                stored to HBM; per-atom counts are integer shared-memory atomics.
//   k_nb_fill    a warp takes an atom, reads its mask words and writes the row (ordered compaction: one predicated
//                store per non-empty word, rows on a fixed 128 B-aligned pitch).
// Double mode: the test runs in FP32 on coordinates relative to the bin centre with a guard band; the (very rare)
// candidates inside the band are decided by the exact un-fused FP64 expression, so the pair set stays bit-exact while
// the FP64 pipe is idle.  Mixed mode: the criterion IS the reference's float expression on absolute float coordinates.
#ifndef NB_THREADS
#define NB_THREADS 128   // 4 warps per bin-block: less waiting at the block barriers than 8 (profiles/r01_ncu_k_nb_mask.txt)
#endif
#define NB_MAXI 64          // owned atoms of a bin handled per sweep
#define NB_MAXR 64          // candidate ranges per bin (<= 2 * (2s+1)^2 with s <= 2 -> 50)

struct BinRanges {
  int start[NB_MAXR];
  int pref[NB_MAXR + 1];
  int nr;
};

// Cooperative: every thread of the block calls it (contains __syncthreads).  The <= 50 (row, kind) range loads are
// issued by as many threads at once — done by one thread they are a serial chain of dependent global loads that the
// whole block waits for (it was the top stall of both list kernels) — and thread 0 only compacts them from shared
// memory, keeping the fixed row order.
__device__ __forceinline__ void bin_ranges(const BinGeom &g, int nlocal, const int *__restrict__ lstart,
                                           const int *__restrict__ gstart, int ex, int ey, int ez, BinRanges &R,
                                           int *s_raw /* shared, 2 * NB_MAXR ints */) {
  const int nyr = 2 * g.s[1] + 1, nzr = 2 * g.s[2] + 1, nraw = 2 * nyr * nzr;
  const int x0 = max(ex - g.s[0], 0), x1 = min(ex + g.s[0], g.mbin[0] - 1);
  for (int k = threadIdx.x; k < nraw; k += blockDim.x) {
    const int kind = k & 1, rowi = k >> 1;
    const int ry = ey + (rowi % nyr) - g.s[1], rz = ez + (rowi / nyr) - g.s[2];
    int j0 = 0, len = 0;
    if (ry >= 0 && ry < g.mbin[1] && rz >= 0 && rz < g.mbin[2]) {
      const int row = (rz * g.mbin[1] + ry) * g.mbin[0];
      const int *sp = kind ? gstart : lstart;
      j0 = sp[row + x0] + (kind ? nlocal : 0);
      len = sp[row + x1 + 1] + (kind ? nlocal : 0) - j0;
    }
    s_raw[2 * k] = j0;
    s_raw[2 * k + 1] = len;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nr = 0, tot = 0;
    for (int k = 0; k < nraw; k++) {
      const int len = s_raw[2 * k + 1];
      if (len > 0) {
        R.start[nr] = s_raw[2 * k];
        R.pref[nr] = tot;
        tot += len;
        nr++;
      }
    }
    R.pref[nr] = tot;
    R.nr = nr;
  }
  __syncthreads();
}

__device__ __forceinline__ int cand_index(const BinRanges &R, int c) {
  int lo = 0, hi = R.nr - 1;  // range r with pref[r] <= c < pref[r+1]
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (R.pref[mid] <= c) lo = mid; else hi = mid - 1;
  }
  return R.start[lo] + (c - R.pref[lo]);
}

__device__ __forceinline__ void bin_of_block(const BinGeom &g, int blk, int &ex, int &ey, int &ez) {
  const int bx = blk % g.nbin[0], by = (blk / g.nbin[0]) % g.nbin[1], bz = blk / (g.nbin[0] * g.nbin[1]);
  ex = bx + g.m[0]; ey = by + g.m[1]; ez = bz + g.m[2];
}

// words of mask storage per bin = atoms(bin) * ceil(ncand(bin)/32)
__global__ void k_nb_ranges(int nblk, int nlocal, const int *__restrict__ lstart, const int *__restrict__ gstart,
                            BinGeom g, int *__restrict__ words) {
  const int blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= nblk) return;
  int ex, ey, ez;
  bin_of_block(g, blk, ex, ey, ez);
  const int binid = (ez * g.mbin[1] + ey) * g.mbin[0] + ex;
  const int ni = lstart[binid + 1] - lstart[binid];
  const int x0 = max(ex - g.s[0], 0), x1 = min(ex + g.s[0], g.mbin[0] - 1);
  int tot = 0;
  for (int dz = -g.s[2]; dz <= g.s[2]; dz++) {
    const int rz = ez + dz;
    if (rz < 0 || rz >= g.mbin[2]) continue;
    for (int dy = -g.s[1]; dy <= g.s[1]; dy++) {
      const int ry = ey + dy;
      if (ry < 0 || ry >= g.mbin[1]) continue;
      const int row = (rz * g.mbin[1] + ry) * g.mbin[0];
      tot += lstart[row + x1 + 1] - lstart[row + x0] + gstart[row + x1 + 1] - gstart[row + x0];
    }
  }
  words[blk] = ni * ((tot + 31) >> 5);
}

#define NB_MAXW 512         // mask words per atom with a precomputed word -> range table (else binary search)

// candidate c -> atom index, starting the search at the range of the word's first candidate
__device__ __forceinline__ int cand_index_from(const BinRanges &R, int r, int c) {
  while (c >= R.pref[r + 1]) r++;
  return R.start[r] + (c - R.pref[r]);
}
__device__ __forceinline__ int cand_range(const BinRanges &R, int c) {
  int lo = 0, hi = R.nr - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (R.pref[mid] <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// UCUT = 1: one cutneighsq for every type pair (the common case): thresholds live in registers.
template <class flt_t, int UCUT>
__global__ void __launch_bounds__(NB_THREADS)
k_nb_mask(int nlocal, const double4 *__restrict__ xq, const float4 *__restrict__ xqf, const int *__restrict__ type,
          const int *__restrict__ lstart, const int *__restrict__ gstart, BinGeom g, int tp1,
          const double *__restrict__ cutneighsq, const long long *__restrict__ maskoff,
          unsigned *__restrict__ maskbuf, int *__restrict__ numneigh, int *__restrict__ maxn, int prefilter) {
  constexpr int NT2 = (B2_MAXTYPES + 1) * (B2_MAXTYPES + 1);
  __shared__ BinRanges R;
  __shared__ float s_cut[NT2], s_band[NT2];
  __shared__ float4 s_xi[NB_MAXI];         // double: bin-relative coordinates; float: absolute.  w = type * tp1
  __shared__ int s_cnt[NB_MAXI];
  __shared__ unsigned char s_wr[NB_MAXW];  // range holding the first candidate of each mask word
  __shared__ int s_self;                   // enumeration index of the bin's first own atom
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NB_THREADS / 32;
  int ex, ey, ez;
  bin_of_block(g, blockIdx.x, ex, ey, ez);
  const int binid = (ez * g.mbin[1] + ey) * g.mbin[0] + ex;
  const int i0 = lstart[binid], i1 = lstart[binid + 1];
  if (i0 == i1) return;
  __shared__ int s_raw[2 * NB_MAXR];
  bin_ranges(g, nlocal, lstart, gstart, ex, ey, ez, R, s_raw);
  if (tid == 0) {
    int cs = 0;
    for (int r = 0; r < R.nr; r++)
      if (R.start[r] <= i0 && i0 < R.start[r] + (R.pref[r + 1] - R.pref[r])) cs = R.pref[r] + (i0 - R.start[r]);
    s_self = cs;
  }
  // Double mode: the test runs in FP32 on bin-relative coordinates.  |rsq - cut| >= band decides in FP32 (the FP32
  // error of rsq is < 1e-6 * cutneighsq, the band is 2e-5 * cutneighsq); a word with any candidate inside the band is
  // redone with the exact un-fused FP64 expression (rare: ~4e-3 of the words).  prefilter = 0 (non-periodic boxes:
  // atoms clamped into edge bins may sit far from the bin centre): band = inf, everything takes the exact test.
  // Mixed mode: band = 0, the float expression on absolute float coordinates IS the criterion.
  for (int k = tid; k < tp1 * tp1; k += NB_THREADS) {
    const float cut = (float)cutneighsq[k];
    s_cut[k] = cut;
    s_band[k] = sizeof(flt_t) == 8 ? (prefilter ? cut * 2.0e-5f : __int_as_float(0x7f800000)) : 0.0f;
  }
  const double cx = g.lo[0] + ((ex - g.m[0]) + 0.5) / g.bininv[0];
  const double cy = g.lo[1] + ((ey - g.m[1]) + 0.5) / g.bininv[1];
  const double cz = g.lo[2] + ((ez - g.m[2]) + 0.5) / g.bininv[2];
  __syncthreads();
  const int ncand = R.pref[R.nr];
  const int nwords = (ncand + 31) >> 5;
  for (int w = tid; w < min(nwords, NB_MAXW); w += NB_THREADS) s_wr[w] = (unsigned char)cand_range(R, w * 32);
  const int nit = i1 - i0;
  const int cself0 = s_self;
  const float ucut = s_cut[tp1 + 1], uband = s_band[tp1 + 1];
  unsigned *mbase = maskbuf + maskoff[blockIdx.x];
  for (int ib = i0; ib < i1; ib += NB_MAXI) {
    const int ni = min(NB_MAXI, i1 - ib);
    __syncthreads();
    for (int k = tid; k < ni; k += NB_THREADS) {
      s_cnt[k] = 0;
      const int tw = type[ib + k] * tp1;
      if (sizeof(flt_t) == 8) {
        const double4 p = xq[ib + k];
        s_xi[k] = make_float4((float)(p.x - cx), (float)(p.y - cy), (float)(p.z - cz), __int_as_float(tw));
      } else {
        const float4 p = xqf[ib + k];
        s_xi[k] = make_float4(p.x, p.y, p.z, __int_as_float(tw));
      }
    }
    __syncthreads();
    for (int w = warp; w < nwords; w += nwarps) {
      const int c = w * 32 + lane;
      const bool valid = c < ncand;
      const int j = !valid ? 0 : (w < NB_MAXW ? cand_index_from(R, s_wr[w], c) : cand_index(R, c));
      float xj, yj, zj;
      const int tj = type[j];
      if (sizeof(flt_t) == 8) {
        const double4 p = xq[j];
        xj = (float)(p.x - cx); yj = (float)(p.y - cy); zj = (float)(p.z - cz);
      } else {
        const float4 p = xqf[j];
        xj = p.x; yj = p.y; zj = p.z;
      }
      if (!valid) xj = 3.0e18f;   // never within any cut-off
      // lane L keeps the mask of atom ii0 + L: one coalesced store per 32 atoms (layout [word][atom of the bin])
      for (int ii0 = 0; ii0 < ni; ii0 += 32) {
        const int nii = min(32, ni - ii0);
        unsigned mymask = 0u;
        bool amb = false;
#pragma unroll 4
        for (int k = 0; k < nii; k++) {
          const float4 pi = s_xi[ii0 + k];
          float rsq;
          if (sizeof(flt_t) == 8) {
            const float dx = pi.x - xj, dy = pi.y - yj, dz = pi.z - zj;
            rsq = dx * dx + dy * dy + dz * dz;
          } else {
            rsq = Pos<float>::rsq(make_float4(pi.x, pi.y, pi.z, 0.f), make_float4(xj, yj, zj, 0.f));
          }
          float cut, band;
          if (UCUT) { cut = ucut; band = uband; }
          else { const int t2 = __float_as_int(pi.w) + tj; cut = s_cut[t2]; band = s_band[t2]; }
          const unsigned mk = __ballot_sync(0xffffffffu, rsq <= cut);
          if (sizeof(flt_t) == 8) amb = amb || (fabsf(rsq - cut) < band);
          if (lane == k) mymask = mk;
        }
        if (sizeof(flt_t) == 8 && __any_sync(0xffffffffu, amb)) {
          // exact redo of this word for the nii atoms
          const double4 pj = xq[j];
          for (int k = 0; k < nii; k++) {
            const int ti = __float_as_int(s_xi[ii0 + k].w);
            const bool hit = valid && Pos<double>::rsq(xq[ib + ii0 + k], pj) <= cutneighsq[ti + tj];
            const unsigned mk = __ballot_sync(0xffffffffu, hit);
            if (lane == k) mymask = mk;
          }
        }
        if (lane < nii) {
          // the atom itself is one of its candidates (rsq = 0): clear that bit
          const int cs = cself0 + (ib - i0) + ii0 + lane;
          if ((cs >> 5) == w) mymask &= ~(1u << (cs & 31));
          mbase[(size_t)w * nit + (ib - i0 + ii0 + lane)] = mymask;
          if (mymask) atomicAdd(&s_cnt[ii0 + lane], __popc(mymask));
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < ni; k += NB_THREADS) {
      numneigh[ib + k] = s_cnt[k];
      atomicMax(maxn, s_cnt[k]);
    }
  }
}

// Fill: a warp takes an atom.  Lane L holds the atom's mask words L, L + 32, ..; the warp then walks the non-empty
// words in order: the word is broadcast by shuffle, lane b of it (bit b set) fetches candidate b's entry from the
// block's staged table and stores it at pos + popc(bits below b) — consecutive words extend the row contiguously, so
// the partial stores merge in L2 — and pos advances by the word's popcount.
#define NB_FILLC 4096       // candidates staged at a time (128 mask words: the usual stencil fits one window)
__global__ void __launch_bounds__(NB_THREADS)
k_nb_fill(int nlocal, const int *__restrict__ type, const int *__restrict__ lstart, const int *__restrict__ gstart,
          BinGeom g, const long long *__restrict__ maskoff, const unsigned *__restrict__ maskbuf,
          const long long *__restrict__ offsets, int *__restrict__ entries, int pack_type) {
  __shared__ BinRanges R;
  __shared__ int s_j[NB_FILLC];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NB_THREADS / 32;
  int ex, ey, ez;
  bin_of_block(g, blockIdx.x, ex, ey, ez);
  const int binid = (ez * g.mbin[1] + ey) * g.mbin[0] + ex;
  const int i0 = lstart[binid], i1 = lstart[binid + 1];
  if (i0 == i1) return;
  __shared__ int s_raw[2 * NB_MAXR];
  bin_ranges(g, nlocal, lstart, gstart, ex, ey, ez, R, s_raw);
  const int ncand = R.pref[R.nr];
  const int nwords = (ncand + 31) >> 5;
  const int nit = i1 - i0;
  const unsigned *mbase = maskbuf + maskoff[blockIdx.x];
  const unsigned ltmask = (1u << lane) - 1u;
  const unsigned sj_lane = (unsigned)__cvta_generic_to_shared(s_j) + (unsigned)lane * 4u;
  for (int c0 = 0; c0 < ncand; c0 += NB_FILLC) {
    const int nc = min(NB_FILLC, ncand - c0);
    __syncthreads();
    // one binary search per warp-load of 32 candidates, then a short forward walk (ranges hold ~40 candidates)
    for (int k = tid; k < nc; k += NB_THREADS) {
      const int r0 = cand_range(R, c0 + (k & ~31));
      const int j = cand_index_from(R, r0, c0 + k);
      s_j[k] = pack_type ? (j | (type[j] << B2_TYPESHIFT)) : j;
    }
    __syncthreads();
    const int w0 = c0 >> 5, w1 = min(nwords, (c0 + NB_FILLC) >> 5);
    for (int ii = warp; ii < nit; ii += nwarps) {
      const unsigned *mcol = mbase + ii;   // word w of this atom: mcol[w * nit]
      long long wpos = offsets[i0 + ii];
      if (c0 > 0) {   // entries already written for this row by earlier candidate windows
        int before = 0;
        for (int w = lane; w < w0; w += 32) before += __popc(mcol[(size_t)w * nit]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
        wpos += before;
      }
      int *row = entries + wpos;
      asm volatile("" : "+l"(row));   // keep the row pointer in registers (else it is rebuilt from wpos per store)
      unsigned mw[NB_FILLC / 1024];
#pragma unroll
      for (int q = 0; q < NB_FILLC / 1024; q++) {
        const int w = w0 + q * 32 + lane;
        mw[q] = w < w1 ? mcol[(size_t)w * nit] : 0u;
      }
      int pos = 0;
#pragma unroll
      for (int q = 0; q < NB_FILLC / 1024; q++) {
        unsigned nz = __ballot_sync(0xffffffffu, mw[q] != 0u);
        const unsigned sq = sj_lane + (unsigned)q * 4096u;   // this lane's column of the group's 32 x 32 entries
        while (nz) {
          const int b = __ffs(nz) - 1;
          nz &= nz - 1;
          const unsigned m = __shfl_sync(0xffffffffu, mw[q], b);
          int jw;   // every lane loads (the table is fully in bounds); only the hit lanes store
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(jw) : "r"(sq + ((unsigned)b << 7)));
          const int at = pos + __popc(m & ltmask);
          if ((m >> lane) & 1u) asm volatile("st.global.u32 [%0], %1;" ::"l"(row + at), "r"(jw) : "memory");
          pos += __popc(m);
        }
      }
    }
  }
}

// Special bonds (molecular systems): a warp takes an atom that has special partners, keeps their ids one per lane and
// scans its row; an entry whose partner's id matches gets the partner's class (1-2, 1-3, 1-4) in bits 30-31 — what stock
// Neighbor writes from atom->special / nspecial (consumed at pair_buck_coul_long_intel.cpp:283).  Ids are global ids
// (upload indices on one GPU): tag[] of an owned atom, tag[ghost_src[]] of a periodic image, halo_tag[] of an atom (or
// image of an atom) received from a neighbour rank; the tables are indexed by global id.
__global__ void __launch_bounds__(128)
k_nb_mark_special(int nlocal, const int *__restrict__ tag, const int *__restrict__ ghost_src,
                  const int *__restrict__ halo_tag, const int *__restrict__ numneigh,
                  const long long *__restrict__ offsets, int *__restrict__ entries, int idxmask,
                  const int *__restrict__ sp_count, const int *__restrict__ sp_list, int sp_max) {
  const int lane = threadIdx.x & 31;
  const int i = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= nlocal) return;
  const int ti = tag[i];
  const int n12 = sp_count[3 * ti], n13 = sp_count[3 * ti + 1], ns = sp_count[3 * ti + 2];
  if (ns == 0) return;
  const int mine = lane < ns ? sp_list[(size_t)ti * sp_max + lane] : -1;
  int *row = entries + offsets[i];
  const int n = numneigh[i];
  for (int k0 = 0; k0 < n; k0 += 32) {
    const int k = k0 + lane;
    int e = 0, tj = -2;
    if (k < n) {
      e = row[k];
      const int j = e & idxmask;
      const int o = j < nlocal ? j : ghost_src[j - nlocal];
      tj = o >= 0 ? tag[o] : (halo_tag ? halo_tag[-1 - o] : -2);   // halo atom of a neighbour rank: its global id
    }
    int which = 0;
    for (int s = 0; s < ns; s++) {
      const int id = __shfl_sync(0xffffffffu, mine, s);
      if (tj == id) which = s < n12 ? 1 : (s < n13 ? 2 : 3);
    }
    if (which) row[k] = e | (which << B2_SBBITS);
  }
}

__global__ void k_row_offsets(int n, int pitch, long long *__restrict__ offsets) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) offsets[i] = (long long)i * pitch;
}

__global__ void k_check_disp(int n, const double4 *__restrict__ xq, const double4 *__restrict__ xhold,
                             double triggersq, int *__restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 a = xq[i], h = xhold[i];
  const double dx = a.x - h.x, dy = a.y - h.y, dz = a.z - h.z;
  if (dx * dx + dy * dy + dz * dz > triggersq) *flag = 1;  // Neighbor::check_distance
}

// list in host order for the parity tests
__global__ void k_export_counts(int n, const int *__restrict__ tag, const int *__restrict__ numneigh,
                                int *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[tag[i]] = numneigh[i];
}
__global__ void k_export_rows(int nlocal, const int *__restrict__ tag, const int *__restrict__ numneigh,
                              const long long *__restrict__ off_in, const int *__restrict__ entries,
                              const long long *__restrict__ off_out, int *__restrict__ out, int packed_type) {
  const int lane = threadIdx.x & 31;
  const int i = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= nlocal) return;
  const long long src = off_in[i], dst = off_out[tag[i]];
  const int n = numneigh[i];
  for (int k = lane; k < n; k += 32) {
    const int e = entries[src + k];
    const int j = e & (packed_type ? B2_IDXMASK26 : B2_NEIGHMASK);
    const int jj = j < nlocal ? tag[j] : j;  // ghosts keep their slot, owned atoms go to host index
    out[dst + k] = jj | (e & ~B2_NEIGHMASK);   // special-bond bits stay, the packed type bits do not
  }
}
__global__ void k_export_ghosts(int nlocal, int ng, const int *__restrict__ tag, const int *__restrict__ src,
                                const int *__restrict__ shift, int *__restrict__ out_src, int *__restrict__ out_shift) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ng) return;
  out_src[k] = src[k] >= 0 ? tag[src[k]] : -1;   // -1: image of a halo atom owned by a neighbour rank
  const int code = shift[k];
  out_shift[3 * k] = code % 3 - 1;
  out_shift[3 * k + 1] = (code / 3) % 3 - 1;
  out_shift[3 * k + 2] = code / 9 - 1;
}

int make_geom(b200md_ctx *ctx, BinGeom &g) {
  NeighState &ns = ctx->neigh;
  ns.cutneighmax = ctx->pair.cutmax + ns.skin;
  ns.cutghost = ns.cutneighmax;
  g.cutghost = ns.cutghost;
  const double binsize_optimal = 0.5 * ns.cutneighmax;
  const int nranks = b2_comm_nranks(ctx), rank = b2_comm_rank(ctx);
  for (int d = 0; d < 3; d++) {
    g.lo[d] = ctx->boxlo[d];
    g.hi[d] = ctx->boxhi[d];
    g.prd[d] = ctx->prd[d];
    g.periodic[d] = ctx->periodic[d];
    g.wlo[d] = ctx->boxlo[d];
    g.whi[d] = ctx->boxhi[d];
    g.wprd[d] = ctx->prd[d];
    g.img[d] = ctx->periodic[d];
    if (d == 2 && nranks > 1) {
      // this rank bins its z slab; the shell bins on either side hold the halo atoms of the neighbour ranks (the
      // z images across the global periodic boundary arrive already shifted)
      if (!ctx->periodic[2]) return b2_fail(ctx, B200MD_EINVAL, "the z-slab decomposition needs a box periodic in z");
      const double slab = ctx->prd[2] / nranks;
      g.lo[2] = ctx->boxlo[2] + rank * slab;
      g.hi[2] = rank == nranks - 1 ? ctx->boxhi[2] : ctx->boxlo[2] + (rank + 1) * slab;
      g.prd[2] = g.hi[2] - g.lo[2];
      g.img[2] = 0;
      ns.slab_lo = g.lo[2];
      ns.slab_hi = g.hi[2];
    }
    if (g.periodic[d] && g.prd[d] < ns.cutghost)
      return b2_fail(ctx, B200MD_EOVERFLOW,
                     "box length %g in dimension %d is shorter than the ghost cutoff %g (multiple periodic "
                     "images are not supported)", g.prd[d], d, ns.cutghost);
    g.nbin[d] = std::max(1, (int)(g.prd[d] / binsize_optimal));
    const double binsize = g.prd[d] / g.nbin[d];
    g.bininv[d] = g.nbin[d] / g.prd[d];
    int s = (int)(ns.cutneighmax / binsize);
    if (s * binsize < ns.cutneighmax) s++;
    g.s[d] = s;
    g.m[d] = g.periodic[d] ? s : 0;
    g.mbin[d] = g.nbin[d] + 2 * g.m[d];
    ns.nbin[d] = g.nbin[d];
    ns.mshell[d] = g.m[d];
    ns.mbin[d] = g.mbin[d];
    ns.bininv[d] = g.bininv[d];
  }
  ns.nbins_tot = (long)g.mbin[0] * g.mbin[1] * g.mbin[2];
  if (ns.nbins_tot > 2000000000L) return b2_fail(ctx, B200MD_EOVERFLOW, "too many neighbour bins");
  return 0;
}

// counting sort of n keys over nbins: returns start[] (exclusive scan, nbins+1) and perm (sorted, stable)
int counting_sort(b200md_ctx *ctx, int n, long nbins, const int *keys, int *count /*already filled*/,
                  int *start, int *cursor, int *perm) {
  NeighState &ns = ctx->neigh;
  TRY(b2_exclusive_scan_i32(ctx, count, start, (size_t)nbins, ns.scan_ws.p));
  CUDA_OK(ctx, cudaMemcpyAsync(cursor, start, (size_t)nbins * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  if (n > 0) {
    k_scatter<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, keys, cursor, perm);
    KERNEL_OK(ctx, "k_scatter");
    k_sort_bins<<<cdiv(nbins, 128), 128, 0, ctx->stream>>>(nbins, start, perm);
    KERNEL_OK(ctx, "k_sort_bins");
  }
  return 0;
}


// ---- multi-GPU host logic -------------------------------------------------------------------------------------
// Comm::exchange: atoms that left this rank's z slab go to the lower / upper neighbour (never further: the rebuild
// trigger fires long before an atom crosses a whole slab).
int migrate(b200md_ctx *ctx) {
  NeighState &ns = ctx->neigh;
  const int nranks = b2_comm_nranks(ctx), rank = b2_comm_rank(ctx);
  const int n = ctx->nlocal;
  const double slab = ctx->prd[2] / nranks;
  const size_t nn = (size_t)n + 64;
  RESERVE(ctx, ns.mig_flag, 3 * nn);
  RESERVE(ctx, ns.mig_off, 3 * (nn + 1));
  RESERVE(ctx, ns.flags, 16);
  RESERVE(ctx, ns.scan_ws, b2_scan_ws_bytes(nn + 1));
  int *f_stay = ns.mig_flag.p, *f_lo = f_stay + nn, *f_hi = f_lo + nn;
  int *o_stay = ns.mig_off.p, *o_lo = o_stay + nn + 1, *o_hi = o_lo + nn + 1;
  CUDA_OK(ctx, cudaMemsetAsync(ns.flags.p + 4, 0, sizeof(int), ctx->stream));
  int cnt[3] = {0, 0, 0};
  if (n > 0) {
    k_mig_dest<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, ctx->xq.p, ctx->boxlo[2], 1.0 / slab, rank, nranks, f_stay, f_lo,
                                                      f_hi, ns.flags.p + 4);
    KERNEL_OK(ctx, "k_mig_dest");
    TRY(b2_exclusive_scan_i32(ctx, f_stay, o_stay, (size_t)n, ns.scan_ws.p));
    TRY(b2_exclusive_scan_i32(ctx, f_lo, o_lo, (size_t)n, ns.scan_ws.p));
    TRY(b2_exclusive_scan_i32(ctx, f_hi, o_hi, (size_t)n, ns.scan_ws.p));
    int *hp = (int *)ctx->h_pinned;
    CUDA_OK(ctx, cudaMemcpyAsync(hp, o_stay + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaMemcpyAsync(hp + 1, o_lo + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaMemcpyAsync(hp + 2, o_hi + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaMemcpyAsync(hp + 3, ns.flags.p + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    cnt[0] = hp[0]; cnt[1] = hp[1]; cnt[2] = hp[2];
    if (hp[3]) return b2_fail(ctx, B200MD_EOVERFLOW, "an atom moved further than one z slab between two rebuilds");
  }
  int from_hi = 0, from_lo = 0;
  TRY(b2_comm_exchange_counts(ctx, cnt[1], cnt[2], &from_hi, &from_lo));
  const int nnew = cnt[0] + from_lo + from_hi;
  const size_t cap = (size_t)nnew + 64;
  RESERVE(ctx, ns.tmp4a, cap);
  RESERVE(ctx, ns.tmp4b, cap);
  RESERVE(ctx, ns.tmpi_a, cap);
  RESERVE(ctx, ns.tmpi_b, cap);
  RESERVE(ctx, ns.mig_send, (size_t)(cnt[1] + cnt[2] + 2) * sizeof(MigAtom));
  RESERVE(ctx, ns.mig_recv, (size_t)(from_lo + from_hi + 2) * sizeof(MigAtom));
  MigAtom *s_lo = (MigAtom *)ns.mig_send.p, *s_hi = s_lo + cnt[1];
  MigAtom *r_lo = (MigAtom *)ns.mig_recv.p, *r_hi = r_lo + from_lo;
  if (n > 0) {
    k_mig_pack<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, f_stay, f_lo, o_stay, o_lo, o_hi, ctx->xq.p, ctx->v.p, ctx->type.p,
                                                      ctx->tag.p, ns.tmp4a.p, ns.tmp4b.p, ns.tmpi_a.p, ns.tmpi_b.p, s_lo,
                                                      s_hi);
    KERNEL_OK(ctx, "k_mig_pack");
  }
  TRY(b2_comm_exchange(ctx, s_lo, (size_t)cnt[1] * sizeof(MigAtom), s_hi, (size_t)cnt[2] * sizeof(MigAtom), r_hi,
                       (size_t)from_hi * sizeof(MigAtom), r_lo, (size_t)from_lo * sizeof(MigAtom)));
  if (from_lo + from_hi > 0) {
    k_mig_unpack<<<cdiv(from_lo + from_hi, 256), 256, 0, ctx->stream>>>(from_lo + from_hi, r_lo, cnt[0], ns.tmp4a.p,
                                                                        ns.tmp4b.p, ns.tmpi_a.p, ns.tmpi_b.p);
    KERNEL_OK(ctx, "k_mig_unpack");
  }
  std::swap(ctx->xq, ns.tmp4a);
  std::swap(ctx->v, ns.tmp4b);
  std::swap(ctx->type, ns.tmpi_a);
  std::swap(ctx->tag, ns.tmpi_b);
  ctx->nlocal = nnew;
  return 0;
}

// Comm::borders along z: owned atoms within cutghost of a slab face go to that neighbour; index lists are kept for
// the per-step forward communication.  rbuf = [from lower | from upper].
int halo_setup(b200md_ctx *ctx, const BinGeom &g) {
  NeighState &ns = ctx->neigh;
  const int n = ctx->nlocal;
  const size_t nn = (size_t)n + 64;
  RESERVE(ctx, ns.mig_flag, 3 * nn);
  RESERVE(ctx, ns.mig_off, 3 * (nn + 1));
  int *f_lo = ns.mig_flag.p, *f_hi = f_lo + nn;
  int *o_lo = ns.mig_off.p, *o_hi = o_lo + nn + 1;
  ns.ns_lo = ns.ns_hi = 0;
  if (n > 0) {
    k_halo_flag<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, ctx->xq.p, g.lo[2] + g.cutghost, g.hi[2] - g.cutghost, f_lo, f_hi);
    KERNEL_OK(ctx, "k_halo_flag");
    TRY(b2_exclusive_scan_i32(ctx, f_lo, o_lo, (size_t)n, ns.scan_ws.p));
    TRY(b2_exclusive_scan_i32(ctx, f_hi, o_hi, (size_t)n, ns.scan_ws.p));
    int *hp = (int *)ctx->h_pinned;
    CUDA_OK(ctx, cudaMemcpyAsync(hp, o_lo + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaMemcpyAsync(hp + 1, o_hi + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    ns.ns_lo = hp[0];
    ns.ns_hi = hp[1];
    RESERVE(ctx, ns.halo_idx, (size_t)ns.ns_lo + ns.ns_hi + 64);
    k_compact_idx<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, f_lo, o_lo, ns.halo_idx.p);
    KERNEL_OK(ctx, "k_compact_idx");
    k_compact_idx<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, f_hi, o_hi, ns.halo_idx.p + ns.ns_lo);
    KERNEL_OK(ctx, "k_compact_idx");
  }
  TRY(b2_comm_exchange_counts(ctx, ns.ns_lo, ns.ns_hi, &ns.nr_hi, &ns.nr_lo));
  RESERVE(ctx, ns.halo_sbuf, (size_t)ns.ns_lo + ns.ns_hi + 64);
  RESERVE(ctx, ns.halo_stype, (size_t)ns.ns_lo + ns.ns_hi + 64);
  RESERVE(ctx, ns.halo_rbuf, (size_t)ns.nr_lo + ns.nr_hi + 64);
  RESERVE(ctx, ns.halo_rtype, (size_t)ns.nr_lo + ns.nr_hi + 64);
  return 0;
}

// Comm::forward_comm along z: positions (and, at build time, types) of the halo atoms; atoms sent across the global
// periodic boundary are shifted by one box length so that they arrive as the right image
int halo_exchange(b200md_ctx *ctx, int with_type) {
  NeighState &ns = ctx->neigh;
  const int nranks = b2_comm_nranks(ctx), rank = b2_comm_rank(ctx);
  const double shift_lo = rank == 0 ? ctx->prd[2] : 0.0;             // to the lower neighbour (wraps for rank 0)
  const double shift_hi = rank == nranks - 1 ? -ctx->prd[2] : 0.0;   // to the upper neighbour
  if (ns.ns_lo > 0) {
    k_halo_pack<<<cdiv(ns.ns_lo, 256), 256, 0, ctx->stream>>>(ns.ns_lo, ns.halo_idx.p, ctx->xq.p, shift_lo, ns.halo_sbuf.p,
                                                              ctx->type.p, with_type ? ns.halo_stype.p : nullptr);
    KERNEL_OK(ctx, "k_halo_pack");
  }
  if (ns.ns_hi > 0) {
    k_halo_pack<<<cdiv(ns.ns_hi, 256), 256, 0, ctx->stream>>>(ns.ns_hi, ns.halo_idx.p + ns.ns_lo, ctx->xq.p, shift_hi,
                                                              ns.halo_sbuf.p + ns.ns_lo, ctx->type.p,
                                                              with_type ? ns.halo_stype.p + ns.ns_lo : nullptr);
    KERNEL_OK(ctx, "k_halo_pack");
  }
  TRY(b2_comm_exchange(ctx, ns.halo_sbuf.p, (size_t)ns.ns_lo * sizeof(double4), ns.halo_sbuf.p + ns.ns_lo,
                       (size_t)ns.ns_hi * sizeof(double4), ns.halo_rbuf.p + ns.nr_lo, (size_t)ns.nr_hi * sizeof(double4),
                       ns.halo_rbuf.p, (size_t)ns.nr_lo * sizeof(double4)));
  if (with_type)
    TRY(b2_comm_exchange(ctx, ns.halo_stype.p, (size_t)ns.ns_lo * sizeof(int), ns.halo_stype.p + ns.ns_lo,
                         (size_t)ns.ns_hi * sizeof(int), ns.halo_rtype.p + ns.nr_lo, (size_t)ns.nr_hi * sizeof(int),
                         ns.halo_rtype.p, (size_t)ns.nr_lo * sizeof(int)));
  if (with_type && ctx->sp_max > 0) {   // molecular systems: the halo atoms' global ids, for the special-bond marks
    const int nsend = ns.ns_lo + ns.ns_hi;
    RESERVE(ctx, ns.halo_stag, (size_t)nsend + 64);
    RESERVE(ctx, ns.halo_rtag, (size_t)ns.nr_lo + ns.nr_hi + 64);
    if (nsend > 0) {
      k_gather_int<<<cdiv(nsend, 256), 256, 0, ctx->stream>>>(nsend, ns.halo_idx.p, ctx->tag.p, ns.halo_stag.p);
      KERNEL_OK(ctx, "k_gather_int");
    }
    TRY(b2_comm_exchange(ctx, ns.halo_stag.p, (size_t)ns.ns_lo * sizeof(int), ns.halo_stag.p + ns.ns_lo,
                         (size_t)ns.ns_hi * sizeof(int), ns.halo_rtag.p + ns.nr_lo, (size_t)ns.nr_hi * sizeof(int),
                         ns.halo_rtag.p, (size_t)ns.nr_lo * sizeof(int)));
  }
  return 0;
}

}  // namespace

int b2_ghost_refresh(b200md_ctx *ctx) {
  NeighState &ns = ctx->neigh;
  ctx->ev_pre_valid = false;
  const bool multi = b2_comm_nranks(ctx) > 1;
  if (multi) TRY(halo_exchange(ctx, 0));   // every rank takes part, with or without ghosts of its own
  if (ctx->nghost == 0) return 0;
  k_ghost_refresh<<<cdiv(ctx->nghost, 256), 256, 0, ctx->stream>>>(
      ctx->nlocal, ctx->nghost, ns.ghost_src.p, ns.ghost_shift.p, ctx->prd[0], ctx->prd[1], ctx->prd[2],
      ctx->xq.p, ctx->prec == B200MD_PREC_MIXED ? ctx->xqf.p : nullptr, ctx->type.p, 0,
      multi ? ns.halo_rbuf.p : nullptr, multi ? ns.halo_rtype.p : nullptr);
  KERNEL_OK(ctx, "k_ghost_refresh");
  return 0;
}

// B200MD_DEBUG_NEIGH=1: wall-clock breakdown of every build on stderr (synchronises; for diagnosis only)
namespace {
struct BuildClock {
  b200md_ctx *ctx;
  bool on;
  double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit BuildClock(b200md_ctx *c) : ctx(c) { static const bool e = getenv("B200MD_DEBUG_NEIGH") != nullptr; on = e; t0 = on ? now() : 0; }
  void mark(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const double t = now();
    fprintf(stderr, "[neigh] %-10s %8.3f ms\n", what, t - t0);
    t0 = t;
  }
};
}  // namespace

int b2_neigh_build(b200md_ctx *ctx) {
  NeighState &ns = ctx->neigh;
  BuildClock clk(ctx);
  ctx->ev_pre_valid = false;
  if (!ctx->box_set) return b2_fail(ctx, B200MD_EINVAL, "neighbour build before b200md_set_box");
  if (!ctx->pair.ready) return b2_fail(ctx, B200MD_EINVAL, "neighbour build before b200md_pair_setup");
  if (ctx->triclinic)
    return b2_fail(ctx, B200MD_EINVAL, "triclinic boxes: only the k-space solver (b200md_pppm_setup / compute) is "
                                       "provided; neighbour lists and pair styles need an orthogonal box");
  ScopedTimer tm(ctx, T_NEIGH);
  BinGeom g;
  TRY(make_geom(ctx, g));
  const bool multi = b2_comm_nranks(ctx) > 1;
  if (multi) {
    // Domain::pbc in the global box, then Comm::exchange between the z slabs
    if (ctx->nlocal > 0) {
      k_wrap_bin<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(ctx->nlocal, ctx->xq.p, g, 1, nullptr, nullptr);
      KERNEL_OK(ctx, "k_wrap_bin");
    }
    TRY(migrate(ctx));
  }
  const int n = ctx->nlocal;
  const long nb = ns.nbins_tot;
  const size_t nmax = (size_t)n + 64;
  const size_t nall_guess = nmax + (size_t)ctx->nghost;

  RESERVE(ctx, ns.bin_of, nmax);
  RESERVE(ctx, ns.bin_sorted, nmax);
  RESERVE(ctx, ns.perm, nmax);
  RESERVE(ctx, ns.bin_count, 2 * (nb + 1));  // [owned counts | ghost counts]
  RESERVE(ctx, ns.bin_start, 2 * (nb + 1));  // [lstart | gstart]
  RESERVE(ctx, ns.bin_cursor, nb + 1);
  RESERVE(ctx, ns.flags, 16);
  RESERVE(ctx, ns.scan_ws, b2_scan_ws_bytes(std::max((size_t)nb + 1, nmax + 1)));
  RESERVE(ctx, ns.tmp4a, nall_guess);
  RESERVE(ctx, ns.tmp4b, nmax);
  RESERVE(ctx, ns.tmpi_a, nall_guess);
  RESERVE(ctx, ns.tmpi_b, nmax);
  RESERVE(ctx, ns.xhold, nmax);
  RESERVE(ctx, ns.ghost_cnt, nmax + 1);
  RESERVE(ctx, ns.goff, nmax + 1);
  RESERVE(ctx, ns.numneigh, nmax);
  RESERVE(ctx, ns.offsets, nmax + 1);
  RESERVE(ctx, ctx->f, nmax);

  {
    // cutneighsq = (cut + skin)^2 (pack_force_const, pair_buck_intel.cpp:399-409): double here, rounded
    // to flt_t where it is used
    const int tp1 = ctx->pair.tp1;
    std::vector<double> cn((size_t)tp1 * tp1, 0.0);
    for (int i = 1; i < tp1; i++)
      for (int j = 1; j < tp1; j++) {
        const double cutneigh = std::sqrt(ctx->pair.h_cutsq[i * tp1 + j]) + ns.skin;
        cn[i * tp1 + j] = cutneigh * cutneigh;
      }
    RESERVE(ctx, ctx->pair.cutneighsq, cn.size());
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->pair.cutneighsq.p, cn.data(), cn.size() * sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  int *lcount = ns.bin_count.p, *gcount = ns.bin_count.p + (nb + 1);
  int *lstart = ns.bin_start.p, *gstart = ns.bin_start.p + (nb + 1);
  CUDA_OK(ctx, cudaMemsetAsync(ns.bin_count.p, 0, 2 * (nb + 1) * sizeof(int), ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(ns.bin_start.p, 0, 2 * (nb + 1) * sizeof(int), ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(ns.flags.p, 0, 16 * sizeof(int), ctx->stream));

  // 1. wrap + bin owned atoms, stable counting sort, permute the resident arrays
  if (n > 0) {
    k_wrap_bin<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, ctx->xq.p, g, multi ? 0 : 1, ns.bin_of.p, lcount);
    KERNEL_OK(ctx, "k_wrap_bin");
  }
  TRY(counting_sort(ctx, n, nb, ns.bin_of.p, lcount, lstart, ns.bin_cursor.p, ns.perm.p));
  if (n > 0) {
    k_permute_atoms<<<cdiv(n, 256), 256, 0, ctx->stream>>>(
        n, ns.perm.p, ctx->xq.p, ctx->v.p, ctx->type.p, ctx->tag.p, ns.bin_of.p, ns.tmp4a.p, ns.tmp4b.p,
        ns.tmpi_a.p, ns.tmpi_b.p, ns.bin_sorted.p, ns.xhold.p);
    KERNEL_OK(ctx, "k_permute_atoms");
    std::swap(ctx->xq, ns.tmp4a);
    std::swap(ctx->v, ns.tmp4b);
    std::swap(ctx->type, ns.tmpi_a);
    std::swap(ctx->tag, ns.tmpi_b);
  }

  clk.mark("sort");
  // 2. ghost atoms: (multi-GPU) z halo from the neighbour ranks, then periodic images of owned + halo atoms;
  //    count / scan / fill, then sort the ghosts by bin too
  int nbase = n;
  const double4 *rbuf = nullptr;
  const int *rtype = nullptr;
  if (multi) {
    TRY(halo_setup(ctx, g));
    TRY(halo_exchange(ctx, 1));
    nbase = n + ns.nr_lo + ns.nr_hi;
    rbuf = ns.halo_rbuf.p;
    rtype = ns.halo_rtype.p;
    RESERVE(ctx, ns.ghost_cnt, (size_t)nbase + 64);
    RESERVE(ctx, ns.goff, (size_t)nbase + 64);
    RESERVE(ctx, ns.scan_ws, b2_scan_ws_bytes((size_t)nbase + 64));
  }
  int ng = 0;
  if (nbase > 0) {
    k_ghost_count<<<cdiv(nbase, 256), 256, 0, ctx->stream>>>(n, nbase, ctx->xq.p, rbuf, g, ns.ghost_cnt.p);
    KERNEL_OK(ctx, "k_ghost_count");
    TRY(b2_exclusive_scan_i32(ctx, ns.ghost_cnt.p, ns.goff.p, (size_t)nbase, ns.scan_ws.p));
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ns.goff.p + nbase, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    ng = *(int *)ctx->h_pinned;
  }
  ctx->nghost = ng;
  const size_t nall = (size_t)n + ng;
  if (nall >= (size_t)B2_NEIGHMASK) return b2_fail(ctx, B200MD_EOVERFLOW, "too many atoms for 30-bit neighbour indices");
  RESERVE_KEEP(ctx, ctx->xq, nall + 64);
  RESERVE_KEEP(ctx, ctx->type, nall + 64);
  // the resident arrays and their sort scratch swap roles at every build: size both alike now, so that the second
  // build does not reallocate (a cudaFree of these buffers costs hundreds of ms with a multi-GB list resident)
  RESERVE(ctx, ns.tmp4a, nall + 64);
  RESERVE(ctx, ns.tmpi_a, nall + 64);
  if (ctx->prec == B200MD_PREC_MIXED) RESERVE(ctx, ctx->xqf, nall + 64);
  if (ng > 0) {
    RESERVE(ctx, ns.gsrc_tmp, (size_t)ng);
    RESERVE(ctx, ns.gshift_tmp, (size_t)ng);
    RESERVE(ctx, ns.gbin, (size_t)ng);
    RESERVE(ctx, ns.gperm, (size_t)ng);
    RESERVE(ctx, ns.ghost_src, (size_t)ng);
    RESERVE(ctx, ns.ghost_shift, (size_t)ng);
    k_ghost_fill<<<cdiv(nbase, 256), 256, 0, ctx->stream>>>(n, nbase, ctx->xq.p, rbuf, ns.bin_sorted.p, g, ns.goff.p,
                                                            ns.gsrc_tmp.p, ns.gshift_tmp.p, ns.gbin.p, gcount);
    KERNEL_OK(ctx, "k_ghost_fill");
  }
  TRY(counting_sort(ctx, ng, nb, ns.gbin.p, gcount, gstart, ns.bin_cursor.p, ns.gperm.p));
  if (ng > 0) {
    k_ghost_permute<<<cdiv(ng, 256), 256, 0, ctx->stream>>>(ng, ns.gperm.p, ns.gsrc_tmp.p, ns.gshift_tmp.p,
                                                            ns.ghost_src.p, ns.ghost_shift.p);
    KERNEL_OK(ctx, "k_ghost_permute");
    k_ghost_refresh<<<cdiv(ng, 256), 256, 0, ctx->stream>>>(
        n, ng, ns.ghost_src.p, ns.ghost_shift.p, ctx->prd[0], ctx->prd[1], ctx->prd[2], ctx->xq.p,
        ctx->prec == B200MD_PREC_MIXED ? ctx->xqf.p : nullptr, ctx->type.p, 1, rbuf, rtype);
    KERNEL_OK(ctx, "k_ghost_refresh");
  }
  TRY(b2_refresh_float_copy(ctx, 0, n));

  clk.mark("ghosts");
  // 3. full list: hit masks + counts, scan to 64-bit CSR offsets, fill from the masks
  long long total = 0;
  int maxn = 0;
  const int pack = (nall <= (size_t)B2_IDXMASK26 && ctx->ntypes < 16) ? 1 : 0;
  ns.packed_type = pack != 0;
  if (n > 0) {
    const int nblk = g.nbin[0] * g.nbin[1] * g.nbin[2];  // one block per interior bin
    if (2 * (2 * g.s[1] + 1) * (2 * g.s[2] + 1) > NB_MAXR)
      return b2_fail(ctx, B200MD_EOVERFLOW, "neighbour stencil too wide for the bin size");
    RESERVE(ctx, ns.mask_words, (size_t)nblk + 1);
    RESERVE(ctx, ns.mask_off, (size_t)nblk + 2);
    RESERVE(ctx, ns.scan_ws, b2_scan_ws_bytes((size_t)nblk + 1));
    k_nb_ranges<<<cdiv(nblk, 128), 128, 0, ctx->stream>>>(nblk, n, lstart, gstart, g, ns.mask_words.p);
    KERNEL_OK(ctx, "k_nb_ranges");
    TRY(b2_exclusive_scan_i32_i64(ctx, ns.mask_words.p, ns.mask_off.p, (size_t)nblk, ns.scan_ws.p));
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ns.mask_off.p + nblk, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    const long long mask_total = *(long long *)ctx->h_pinned;
    RESERVE(ctx, ns.maskbuf, (size_t)mask_total + 64);
    clk.mark("ranges");
    const int prefilter = (g.periodic[0] && g.periodic[1] && g.periodic[2]) ? 1 : 0;
    bool ucut = true;   // one cutneighsq for all type pairs?
    {
      const int tp1 = ctx->pair.tp1;
      for (int a = 1; a < tp1; a++)
        for (int b = 1; b < tp1; b++)
          if (ctx->pair.h_cutsq[a * tp1 + b] != ctx->pair.h_cutsq[tp1 + 1]) ucut = false;
    }
#define NB_MASK(F, U)                                                                                              \
  k_nb_mask<F, U><<<nblk, NB_THREADS, 0, ctx->stream>>>(n, ctx->xq.p, ctx->xqf.p, ctx->type.p, lstart, gstart, g,   \
                                                        ctx->pair.tp1, ctx->pair.cutneighsq.p, ns.mask_off.p,      \
                                                        ns.maskbuf.p, ns.numneigh.p, ns.flags.p + 2, prefilter)
    {
      ScopedTimer tk(ctx, K_NB_MASK);
      if (ctx->prec == B200MD_PREC_MIXED) { if (ucut) NB_MASK(float, 1); else NB_MASK(float, 0); }
      else { if (ucut) NB_MASK(double, 1); else NB_MASK(double, 0); }
    }
#undef NB_MASK
    KERNEL_OK(ctx, "k_nb_mask");
    clk.mark("mask");
    // rows get a fixed pitch = the longest row rounded up to whole 128 B lines: every row starts line-aligned (the
    // pair kernel's row loads are 3.5 % faster than on packed CSR rows) at the price of ~12 % more list memory.
    // offsets[n] (the packed total) is still produced: statistics and the host export use it.
    TRY(b2_exclusive_scan_i32_i64(ctx, ns.numneigh.p, ns.offsets.p, (size_t)n, ns.scan_ws.p));
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ns.offsets.p + n, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned + 1, ns.flags.p + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    total = *(long long *)ctx->h_pinned;
    maxn = *(int *)(ctx->h_pinned + 1);
    // A very inhomogeneous system (one crowded spot sets the pitch for every row) would pay for the padding in memory:
    // beyond 50 % overhead the rows stay packed at the scanned CSR offsets.  B200MD_LIST_CSR=1 forces that layout.
    const int pitch = (maxn + 31) & ~31;
    const char *csr_env = getenv("B200MD_LIST_CSR");
    const bool padded = !(csr_env && csr_env[0] == '1') && (double)pitch * n <= 1.5 * (double)total + 4.0e6;
    ns.pitch = padded ? pitch : 0;
    if (padded) {
      RESERVE(ctx, ns.entries, nmax * (size_t)pitch + 64);
      k_row_offsets<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, pitch, ns.offsets.p);
      KERNEL_OK(ctx, "k_row_offsets");
    } else {
      RESERVE(ctx, ns.entries, (size_t)total + 64);
    }
    clk.mark("scan+alloc");
    {
      ScopedTimer tk(ctx, K_NB_FILL);
      k_nb_fill<<<nblk, NB_THREADS, 0, ctx->stream>>>(n, ctx->type.p, lstart, gstart, g, ns.mask_off.p, ns.maskbuf.p,
                                                      ns.offsets.p, ns.entries.p, pack);
    }
    KERNEL_OK(ctx, "k_nb_fill");
    clk.mark("fill");
  }
  if (ctx->sp_max > 0 && n > 0) {
    k_nb_mark_special<<<cdiv((long)n * 32, 128), 128, 0, ctx->stream>>>(
        n, ctx->tag.p, ns.ghost_src.p, multi ? ns.halo_rtag.p : nullptr, ns.numneigh.p, ns.offsets.p, ns.entries.p,
        ns.packed_type ? B2_IDXMASK26 : B2_NEIGHMASK, ctx->sp_count.p, ctx->sp_list.p, ctx->sp_max);
    KERNEL_OK(ctx, "k_nb_mark_special");
  }
  ns.total_entries = total;
  ns.max_numneigh = maxn;
  ns.nbuilds++;
  ns.ago = 0;
  ns.ready = true;
  return 0;
}

int b2_neigh_check_trigger(b200md_ctx *ctx, int *trigger) {
  NeighState &ns = ctx->neigh;
  *trigger = 0;
  if (ctx->nlocal == 0 && b2_comm_nranks(ctx) == 1) return 0;
  RESERVE(ctx, ns.flags, 16);
  CUDA_OK(ctx, cudaMemsetAsync(ns.flags.p, 0, sizeof(int), ctx->stream));
  if (ctx->nlocal > 0) {
    k_check_disp<<<cdiv(ctx->nlocal, 256), 256, 0, ctx->stream>>>(ctx->nlocal, ctx->xq.p, ns.xhold.p,
                                                                  0.25 * ns.skin * ns.skin, ns.flags.p);
    KERNEL_OK(ctx, "k_check_disp");
  }
  TRY(b2_comm_allreduce_max_int(ctx, ns.flags.p, 1));   // every rank rebuilds when any rank must
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, ns.flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  *trigger = *(int *)ctx->h_pinned;
  return 0;
}

extern "C" {

int b200md_neigh_setup(b200md_ctx *ctx, double skin, int every, int delay, int check) {
  if (!ctx || skin < 0 || every < 1 || delay < 0)
    return b2_fail(ctx, B200MD_EINVAL, "b200md_neigh_setup: bad arguments");
  ctx->neigh.skin = skin;
  ctx->neigh.every = every;
  ctx->neigh.delay = delay;
  ctx->neigh.check = check;
  ctx->neigh.ready = false;
  // the PPPM brick halo is sized from skin/2 (PPPM::set_grid_local): a k-space state set up for a smaller skin is stale
  b2_pppm_skin_changed(ctx, skin);
  return 0;
}

int b200md_neigh_build(b200md_ctx *ctx) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  return b2_neigh_build(ctx);
}

int b200md_neigh_decide(b200md_ctx *ctx, long ntimestep, int *rebuilt) {
  if (!ctx) return B200MD_EINVAL;
  cudaSetDevice(ctx->device);
  NeighState &ns = ctx->neigh;
  (void)ntimestep;
  int build = 0;
  if (!ns.ready) build = 1;
  else {
    // Neighbor::decide
    ns.ago++;
    if (ns.ago >= ns.delay && ns.ago % ns.every == 0) {
      if (!ns.check) build = 1;
      else TRY(b2_neigh_check_trigger(ctx, &build));
    }
  }
  if (build) TRY(b2_neigh_build(ctx));
  else {
    ScopedTimer tm(ctx, T_COMM);
    TRY(b2_ghost_refresh(ctx));
  }
  if (rebuilt) *rebuilt = build;
  return 0;
}

int b200md_neigh_stats(b200md_ctx *ctx, long *total_entries, int *nghost, int *max_numneigh, long *nbuilds) {
  if (!ctx) return B200MD_EINVAL;
  if (total_entries) *total_entries = (long)ctx->neigh.total_entries;
  if (nghost) *nghost = ctx->nghost;
  if (max_numneigh) *max_numneigh = ctx->neigh.max_numneigh;
  if (nbuilds) *nbuilds = ctx->neigh.nbuilds;
  return 0;
}

int b200md_neigh_download(b200md_ctx *ctx, int *numneigh, long *offsets, int *entries, int *ghost_src,
                          int *ghost_shift) {
  if (!ctx || !ctx->neigh.ready) return b2_fail(ctx, B200MD_EINVAL, "no neighbour list built");
  cudaSetDevice(ctx->device);
  NeighState &ns = ctx->neigh;
  const int n = ctx->nlocal, ng = ctx->nghost;
  const size_t total = (size_t)ns.total_entries;
  DevBuf<int> cnt, ent, gs, gsh;
  DevBuf<long long> off;
  int rc = 0;
  auto cleanup = [&]() { cnt.free_(); ent.free_(); gs.free_(); gsh.free_(); off.free_(); };
  if (cnt.reserve(n + 1) || off.reserve(n + 2) || ent.reserve(total + 1) || gs.reserve(ng + 1) || gsh.reserve(3 * (size_t)ng + 1)) {
    cleanup();
    return b2_fail(ctx, B200MD_ENOMEM, "out of device memory exporting the neighbour list");
  }
  if (n == 0) cudaMemsetAsync(off.p, 0, sizeof(long long), ctx->stream);   // offsets[0] of an empty list
  if (n > 0) {
    k_export_counts<<<cdiv(n, 256), 256, 0, ctx->stream>>>(n, ctx->tag.p, ns.numneigh.p, cnt.p);
    ctx->launches++;
    rc = b2_exclusive_scan_i32_i64(ctx, cnt.p, off.p, (size_t)n, ns.scan_ws.p);
    if (!rc) {
      k_export_rows<<<cdiv((long)n * 32, 256), 256, 0, ctx->stream>>>(n, ctx->tag.p, ns.numneigh.p, ns.offsets.p,
                                                                    ns.entries.p, off.p, ent.p, ns.packed_type ? 1 : 0);
      ctx->launches++;
    }
  }
  if (!rc && ng > 0) {
    k_export_ghosts<<<cdiv(ng, 256), 256, 0, ctx->stream>>>(n, ng, ctx->tag.p, ns.ghost_src.p, ns.ghost_shift.p, gs.p, gsh.p);
    ctx->launches++;
  }
  if (!rc) {
    static_assert(sizeof(long) == sizeof(long long), "LP64 expected");
    if (numneigh && n) cudaMemcpyAsync(numneigh, cnt.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (offsets) cudaMemcpyAsync(offsets, off.p, ((size_t)n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (entries && total) cudaMemcpyAsync(entries, ent.p, total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (ghost_src && ng) cudaMemcpyAsync(ghost_src, gs.p, (size_t)ng * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (ghost_shift && ng) cudaMemcpyAsync(ghost_shift, gsh.p, 3 * (size_t)ng * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = b2_fail(ctx, B200MD_ECUDA, "neighbour export failed: %s", cudaGetErrorString(e));
  }
  cleanup();
  return rc;
}

}  // extern "C"
