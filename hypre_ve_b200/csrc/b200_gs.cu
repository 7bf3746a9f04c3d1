// b200_gs.cu -- hybrid Gauss-Seidel relaxation (hypre_BoomerAMGRelax types 3/4/6 and the l1 variants
// 8/13/14, parcsr_ls/par_relax.c:1875-2265, :3492-4091, :4340-5124) as exact, level-scheduled sweeps.
//
// "Hybrid" (par_relax.c:4400-4412): the rows of a rank are cut into T contiguous blocks
// (size = n/T, the first n%T blocks one row longer); inside a block the sweep is sequential
// Gauss-Seidel, across blocks (and across ranks) it reads the values from before the sweep.  In the
// reference T is the OpenMP thread count; here it is the parameter "GSBlocks".  For a given T the
// result is the reference's, bit for bit: a row reads the NEW value of every in-block neighbour that
// precedes it in sweep order, the OLD value of everything else, and sums its entries in storage order
// in one thread (no FMA).
//
// Scheduling.  level[i] = 1 + max(level[j] : j in the block of i, j < i, a_ij != 0 or a_ji != 0),
// found by frontier peeling (Kahn's algorithm, one small kernel per level).  Levels built on the
// symmetrised pattern serve both directions: descending level order is valid for the backward sweep.
//   * block path (blocks of <= 16384 rows): ONE CTA PER GAUSS-SEIDEL BLOCK.  The block's iterate lives in
//     shared memory, its levels are separated by __syncthreads(), everything outside the block is read
//     from the pre-sweep copy: no inter-CTA synchronisation at all.  Rows are visited in
//     (block, level, row) order; the next level's row is prefetched into registers before the barrier.
//   * global path (any block size, T = 1 included): CTAs draw chunk tickets in level order, prefetch
//     their rows and then wait on the completion counter of the previous level (a soft barrier that is
//     deadlock free without a cooperative launch because tickets are handed out in dependency order);
//     new values are read with ld.global.cg after the acquire.  This path is bound by
//     (#levels x one L2 round trip): 766 levels for the 7-pt 256^3 grid, thousands on coarse grids.
#include "b200_internal.h"
#include <algorithm>

struct b200_gs_plan_s {
  int n = 0, T = 1, size = 0, rest = 0;
  int nlevels = 0;            // depth of the deepest block
  bool block_path = false;
  bool lane_path = false;     // one thread per Gauss-Seidel block: no schedule needed
  // lane path, sliced-ELL copy of A (slice = the 32 rows the lanes of a warp visit in the same step)
  int maxm = 0;               // rows of the longest block
  int *slice_ptr = nullptr;   // [nwarps * maxm + 1] first padded entry of every slice
  int *sell_j = nullptr;      // column ids, -1 = padding
  double *sell_a = nullptr;
  int maxblock = 0;
  int *perm = nullptr;        // rows sorted by (level,row) [global path] or (block,level,row) [block path]
  // global path
  int nchunks = 0;
  int *level_off = nullptr;   // [nlevels+1]
  int *chunks = nullptr;      // int4 per chunk {pos0, pos1, level, 0}
  int *ctr = nullptr;         // [1 + nlevels] ticket counter, rows finished per level
  // block path
  int nseg = 0;
  int *seg_off = nullptr;     // [nseg+1] first position of every non-empty (block, level) segment
  int *blk_seg = nullptr;     // [T+1] first segment of every block
  double *old = nullptr;      // [n] iterate before the sweep (T > 1)
};

namespace {

constexpr int GS_NT = 128;     // global path: threads per CTA = rows per chunk
constexpr int GS_PRE = 8;      // row entries prefetched into registers
constexpr int GS_BLOCK_CAP = 16384;   // rows of one Gauss-Seidel block that fit the shared-memory iterate (128 KB)
constexpr int GS_LANE_CAP = 2048;     // blocks up to this many rows are swept by ONE THREAD each

// block of row i: [ns, ne)   (par_relax.c:4400-4412)
__device__ __forceinline__ void blk_range(int i, int size, int rest, int &ns, int &ne, int &b) {
  const int split = rest * (size + 1);
  if (i < split) { b = i / (size + 1); ns = b * (size + 1); ne = ns + size + 1; }
  else { b = rest + (i - split) / size; ns = b * size + rest; ne = ns + size; }
}

__global__ void gs_indeg_kernel(int n, int size, int rest, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                const int *__restrict__ T_i, const int *__restrict__ T_j, int *__restrict__ indeg,
                                int *__restrict__ level, int *__restrict__ frontier, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ns, ne, b;
  blk_range(i, size, rest, ns, ne, b);
  int d = 0;
  for (int jj = A_i[i]; jj < A_i[i + 1]; jj++) { const int j = A_j[jj]; d += (j < i && j >= ns); }
  if (T_i) for (int jj = T_i[i]; jj < T_i[i + 1]; jj++) { const int j = T_j[jj]; d += (j < i && j >= ns); }
  indeg[i] = d;
  if (d == 0) { level[i] = 0; frontier[atomicAdd(cnt, 1)] = i; }
}
// one peeling round: rows of level r release their in-block successors; cnt is a ring of three counters
__global__ void gs_peel_kernel(int n, int size, int rest, int r, const int *__restrict__ A_i, const int *__restrict__ A_j,
                               const int *__restrict__ T_i, const int *__restrict__ T_j, int *__restrict__ indeg,
                               int *__restrict__ level, const int *__restrict__ fin, int *__restrict__ fout,
                               int *__restrict__ cnt, int *__restrict__ nlevels) {
  const int nin = cnt[r % 3];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    cnt[(r + 2) % 3] = 0;                               // becomes the output counter of the next round
    if (nin == 0) atomicMin(nlevels, r);
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nin; idx += gridDim.x * blockDim.x) {
    const int i = fin[idx];
    int ns, ne, b;
    blk_range(i, size, rest, ns, ne, b);
    for (int jj = A_i[i]; jj < A_i[i + 1]; jj++) {
      const int j = A_j[jj];
      if (j > i && j < ne && atomicSub(&indeg[j], 1) == 1) { level[j] = r + 1; fout[atomicAdd(&cnt[(r + 1) % 3], 1)] = j; }
    }
    if (T_i)
      for (int jj = T_i[i]; jj < T_i[i + 1]; jj++) {
        const int j = T_j[jj];
        if (j > i && j < ne && atomicSub(&indeg[j], 1) == 1) { level[j] = r + 1; fout[atomicAdd(&cnt[(r + 1) % 3], 1)] = j; }
      }
  }
}
// pattern symmetry of the square block: every (i,j), j < n, must have a partner (j,i)
__global__ void gs_symcheck_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int jj = A_i[i]; jj < A_i[i + 1]; jj++) {
    const int j = A_j[jj];
    if (j >= n || j == i) continue;
    bool found = false;
    for (int kk = A_i[j]; kk < A_i[j + 1]; kk++)
      if (A_j[kk] == i) { found = true; break; }
    if (!found) { *flag = 1; return; }
  }
}
__global__ void gs_unit_rows_kernel(int n, int *__restrict__ L_i) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) L_i[i] = i;
}
__global__ void gs_block_key_kernel(int n, int size, int rest, int D, int *__restrict__ level) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ns, ne, b;
  blk_range(i, size, rest, ns, ne, b);
  level[i] = b * D + level[i];
}
__global__ void gs_seg_flag_kernel(int K, const int *__restrict__ key_off, int *__restrict__ flag) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > K) return;
  flag[k] = (k < K && key_off[k + 1] > key_off[k]) ? 1 : 0;
}
__global__ void gs_seg_scatter_kernel(int K, int T, int D, const int *__restrict__ key_off, const int *__restrict__ pos,
                                      int *__restrict__ seg_off, int *__restrict__ blk_seg) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > K) return;
  if (k == K) { seg_off[pos[K]] = key_off[K]; blk_seg[T] = pos[K]; return; }
  if (key_off[k + 1] > key_off[k]) seg_off[pos[k]] = key_off[k];
  if (k % D == 0) blk_seg[k / D] = pos[k];
}

__device__ __forceinline__ int ld_volatile(const int *p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- global path ---------------------------------------------------------------------------------------
// DIR +1 forward (types 3, 13) / -1 backward (4, 14).  CLASSIC: u_i = res / a_ii, diagonal (stored first)
// skipped; else u_i += res / l1_i over the whole row.  ZERO: the iterate is 0 on entry.  `old` = pre-sweep
// copy for out-of-block reads (T > 1), null when T == 1 (then every non-dependency is still unwritten in u).
template <int DIR, bool CLASSIC, bool ZERO>
__global__ void __launch_bounds__(GS_NT)
gs_sweep_kernel(int n, int size, int rest, int nchunks, int nlevels, const int4 *__restrict__ chunks,
                const int *__restrict__ level_off, const int *__restrict__ perm, const int *__restrict__ A_i,
                const int *__restrict__ A_j, const double *__restrict__ A_a, const double *__restrict__ f,
                const double *__restrict__ l1, const double *__restrict__ old, double *u, int *ctr) {
  __shared__ int s_ticket;
  int *ticket = ctr, *done = ctr + 1;
  while (true) {
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int t = s_ticket;
    __syncthreads();
    if (t >= nchunks) return;
    const int4 c = chunks[DIR > 0 ? t : nchunks - 1 - t];
    const int p = c.x + (int)threadIdx.x;
    const bool active = p < c.y;
    // ---- prefetch everything that does not depend on the previous level ----------------------------
    int i = 0, b = 0, e = 0, ns = 0, ne = 0, bk = 0;
    double fi = 0.0, di = 0.0, ui = 0.0;
    int pj[GS_PRE];
    double pa[GS_PRE], pu[GS_PRE];
    if (active) {
      i = perm[p];
      blk_range(i, size, rest, ns, ne, bk);
      b = A_i[i]; e = A_i[i + 1];
      fi = f[i];
      di = CLASSIC ? A_a[b] : l1[i];
      if (!ZERO && !CLASSIC) ui = __ldcg(u + i);
#pragma unroll
      for (int k = 0; k < GS_PRE; k++) {
        pj[k] = -1; pa[k] = 0.0; pu[k] = 0.0;
        if (b + k < e) {
          const int j = A_j[b + k];
          pj[k] = j; pa[k] = A_a[b + k];
          const bool dep = (DIR > 0) ? (j < i && j >= ns) : (j > i && j < ne);
          if (!dep && !ZERO) pu[k] = (old && j < n && (j < ns || j >= ne)) ? old[j] : __ldcg(u + j);
        }
      }
    }
    // ---- soft barrier on the previous level ------------------------------------------------------------
    const int wl = c.z - DIR;
    if (threadIdx.x == 0 && wl >= 0 && wl < nlevels) {
      const int target = level_off[wl + 1] - level_off[wl];
      unsigned nsleep = 20;
      while (ld_volatile(done + wl) < target) { __nanosleep(nsleep); if (nsleep < 400) nsleep += nsleep; }
      __threadfence();
    }
    __syncthreads();
    // ---- the row, entries in storage order ---------------------------------------------------------------
    if (active && di != 0.0) {
      double res = fi;
#pragma unroll
      for (int k = 0; k < GS_PRE; k++) {
        if (b + k < e && !(CLASSIC && k == 0)) {
          const int j = pj[k];
          const bool dep = (DIR > 0) ? (j < i && j >= ns) : (j > i && j < ne);
          const double uj = dep ? __ldcg(u + j) : pu[k];
          res -= pa[k] * uj;
        }
      }
      for (int jj = b + GS_PRE; jj < e; jj++) {
        const int j = A_j[jj];
        const bool dep = (DIR > 0) ? (j < i && j >= ns) : (j > i && j < ne);
        double uj = 0.0;
        if (dep) uj = __ldcg(u + j);
        else if (!ZERO) uj = (old && j < n && (j < ns || j >= ne)) ? old[j] : __ldcg(u + j);
        res -= A_a[jj] * uj;
      }
      u[i] = CLASSIC ? res / di : ui + res / di;
    } else if (active && ZERO) {
      u[i] = 0.0;                                         // skipped rows keep the (zero) iterate
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(done + c.z, c.y - c.x); }
  }
}

// ---- block path: one CTA per Gauss-Seidel block, iterate in shared memory ----------------------------------
struct RowRegs {
  int i, b, e;
  double fi, di;
  int pj[GS_PRE];
  double pa[GS_PRE];
};
template <bool CLASSIC>
__device__ __forceinline__ void gs_load_row(RowRegs &r, int p, const int *__restrict__ perm, const int *__restrict__ A_i,
                                            const int *__restrict__ A_j, const double *__restrict__ A_a,
                                            const double *__restrict__ f, const double *__restrict__ l1) {
  r.i = perm[p];
  r.b = A_i[r.i]; r.e = A_i[r.i + 1];
  r.fi = f[r.i];
  r.di = CLASSIC ? A_a[r.b] : l1[r.i];
#pragma unroll
  for (int k = 0; k < GS_PRE; k++) {
    r.pj[k] = 0; r.pa[k] = 0.0;
    if (r.b + k < r.e) { r.pj[k] = A_j[r.b + k]; r.pa[k] = A_a[r.b + k]; }
  }
}
template <bool CLASSIC, bool ZERO>
__device__ __forceinline__ void gs_do_row(const RowRegs &r, int n, int ns, int ne, double *u_s, const double *__restrict__ old,
                                          const int *__restrict__ A_j, const double *__restrict__ A_a) {
  if (r.di == 0.0) return;                                 // u_s already holds the old value (0 when ZERO)
  double res = r.fi;
#pragma unroll
  for (int k = 0; k < GS_PRE; k++) {
    if (r.b + k < r.e && !(CLASSIC && k == 0)) {
      const int j = r.pj[k];
      double uj;
      if (j >= ns && j < ne) uj = u_s[j - ns];
      else uj = ZERO ? 0.0 : old[j];
      res -= r.pa[k] * uj;
    }
  }
  for (int jj = r.b + GS_PRE; jj < r.e; jj++) {
    const int j = A_j[jj];
    double uj;
    if (j >= ns && j < ne) uj = u_s[j - ns];
    else uj = ZERO ? 0.0 : old[j];
    res -= A_a[jj] * uj;
  }
  u_s[r.i - ns] = CLASSIC ? res / r.di : u_s[r.i - ns] + res / r.di;
}
// `old`: the iterate before the relaxation call, read for everything outside the block (for T == 1 that is
// only ghost columns >= n and old == u); `u`: iterate, updated in place block by block.
template <int DIR, bool CLASSIC, bool ZERO>
__global__ void gs_block_kernel(int n, int T, int size, int rest, const int *__restrict__ blk_seg, const int *__restrict__ seg_off,
                                const int *__restrict__ perm, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                const double *__restrict__ A_a, const double *__restrict__ f, const double *__restrict__ l1,
                                const double *old, double *u) {
  extern __shared__ double u_s[];
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int bk = blockIdx.x; bk < T; bk += gridDim.x) {
    const int ns = bk < rest ? bk * (size + 1) : bk * size + rest;
    const int ne = ns + size + (bk < rest ? 1 : 0);
    const int m = ne - ns;
    if (m == 0) continue;
    for (int k = tid; k < m; k += NT) u_s[k] = ZERO ? 0.0 : u[ns + k];   // current iterate (second half of 6/8: post-forward)
    const int s0 = blk_seg[bk], s1 = blk_seg[bk + 1];
    int seg = DIR > 0 ? s0 : s1 - 1;
    RowRegs nxt;
    nxt.i = 0; nxt.b = 0; nxt.e = 0; nxt.fi = 0.0; nxt.di = 0.0;
    int o0 = seg_off[seg], o1 = seg_off[seg + 1];
    if (o0 + tid < o1) gs_load_row<CLASSIC>(nxt, o0 + tid, perm, A_i, A_j, A_a, f, l1);
    __syncthreads();
    for (int cnt = s1 - s0; cnt > 0; cnt--) {
      const RowRegs cur = nxt;
      const int c0 = o0, c1 = o1;
      if (cnt > 1) {                                       // prefetch this thread's row of the next level
        seg += DIR;
        o0 = seg_off[seg]; o1 = seg_off[seg + 1];
        if (o0 + tid < o1) gs_load_row<CLASSIC>(nxt, o0 + tid, perm, A_i, A_j, A_a, f, l1);
      }
      if (c0 + tid < c1) gs_do_row<CLASSIC, ZERO>(cur, n, ns, ne, u_s, old, A_j, A_a);
      for (int p = c0 + tid + NT; p < c1; p += NT) {       // levels wider than the CTA
        RowRegs r;
        gs_load_row<CLASSIC>(r, p, perm, A_i, A_j, A_a, f, l1);
        gs_do_row<CLASSIC, ZERO>(r, n, ns, ne, u_s, old, A_j, A_a);
      }
      __syncthreads();
    }
    for (int k = tid; k < m; k += NT) u[ns + k] = u_s[k];
    __syncthreads();
  }
}

// ---- sliced-ELL copy for the lane path ---------------------------------------------------------------------
// Warp w owns blocks 32w .. 32w+31 (lane l -> block 32w + l); in step k every lane visits row ns_l + k.
// Slice (w, k) stores those 32 rows column-major: entry e of lane l at slice_ptr[w*maxm + k] + 32*e + l,
// padded with column -1 to the longest of the 32 rows, so a warp's loads are fully coalesced.
__device__ __forceinline__ void lane_block(int bk, int T, int size, int rest, int &ns, int &m) {
  if (bk >= T) { ns = 0; m = 0; return; }
  ns = bk < rest ? bk * (size + 1) : bk * size + rest;
  m = size + (bk < rest ? 1 : 0);
}
__global__ void gs_sell_width_kernel(int nwarps, int maxm, int T, int size, int rest, const int *__restrict__ A_i,
                                     int *__restrict__ width) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= nwarps * maxm) { return; }
  const int w = s / maxm, k = s % maxm;
  int ns, m, len = 0;
  lane_block(32 * w + lane, T, size, rest, ns, m);
  if (k < m) len = A_i[ns + k + 1] - A_i[ns + k];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, off));
  if (lane == 0) width[s] = 32 * len;
  if (s == 0 && lane == 0) width[nwarps * maxm] = 0;
}
__global__ void gs_sell_fill_kernel(int nwarps, int maxm, int T, int size, int rest, const int *__restrict__ A_i,
                                    const int *__restrict__ A_j, const double *__restrict__ A_a,
                                    const int *__restrict__ slice_ptr, int *__restrict__ sj, double *__restrict__ sa) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= nwarps * maxm) return;
  const int w = s / maxm, k = s % maxm;
  int ns, m;
  lane_block(32 * w + lane, T, size, rest, ns, m);
  const int base = slice_ptr[s], W = (slice_ptr[s + 1] - base) >> 5;
  int b = 0, len = 0;
  if (k < m) { b = A_i[ns + k]; len = A_i[ns + k + 1] - b; }
  for (int e = 0; e < W; e++) {
    sj[base + 32 * e + lane] = e < len ? A_j[b + e] : -1;
    sa[base + 32 * e + lane] = e < len ? A_a[b + e] : 0.0;
  }
}
// the sweep over the sliced copy: same arithmetic, same order as gs_lane_kernel
template <int DIR, bool CLASSIC, bool ZERO>
__global__ void gs_sell_kernel(int n, int T, int size, int rest, int maxm, const int *__restrict__ slice_ptr,
                               const int *__restrict__ sj, const double *__restrict__ sa, const double *__restrict__ f,
                               const double *__restrict__ l1, const double *old, double *u) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  int ns, m;
  lane_block(32 * w + lane, T, size, rest, ns, m);
  if (32 * w >= T) return;
  const int ne = ns + m;
  const int *sp = slice_ptr + (size_t)w * maxm;
  for (int t = 0; t < m; t++) {
    const int k = DIR > 0 ? t : m - 1 - t;
    const int i = ns + k;
    const int base = sp[k], W = (sp[k + 1] - base) >> 5;
    const int *cj = sj + base + lane;
    const double *ca = sa + base + lane;
    const double di = CLASSIC ? ca[0] : l1[i];
    if (di == 0.0) { if (ZERO) u[i] = 0.0; continue; }
    double res = f[i];
    for (int e = CLASSIC ? 1 : 0; e < W; e++) {
      const int j = cj[32 * e];
      if (j < 0) break;                                     // padding: the row is over
      double uj;
      if (j >= ns && j < ne) {
        const bool swept = (DIR > 0) ? (j < i) : (j > i);
        uj = (ZERO && !swept) ? 0.0 : u[j];
      } else {
        uj = ZERO ? 0.0 : old[j];
      }
      res -= ca[32 * e] * uj;
    }
    u[i] = CLASSIC ? res / di : (ZERO ? 0.0 : u[i]) + res / di;
  }
}

// ---- lane path: one THREAD per Gauss-Seidel block ------------------------------------------------------------
// With many small blocks the reference's own parallelisation maps one to one: thread b walks the rows of
// block b in sweep order, reads its own earlier results back from u (program order makes them visible to
// itself) and everything outside the block from the pre-sweep copy.  No schedule, no barriers; the
// dependency chain of a block is hidden by the other blocks resident on the SM.
template <int DIR, bool CLASSIC, bool ZERO>
__global__ void gs_lane_kernel(int n, int T, int size, int rest, const int *__restrict__ A_i, const int *__restrict__ A_j,
                               const double *__restrict__ A_a, const double *__restrict__ f, const double *__restrict__ l1,
                               const double *old, double *u) {
  const int bk = blockIdx.x * blockDim.x + threadIdx.x;
  if (bk >= T) return;
  const int ns = bk < rest ? bk * (size + 1) : bk * size + rest;
  const int ne = ns + size + (bk < rest ? 1 : 0);
  for (int i = (DIR > 0 ? ns : ne - 1); DIR > 0 ? i < ne : i >= ns; i += DIR) {
    const int b = A_i[i], e = A_i[i + 1];
    const double di = CLASSIC ? A_a[b] : l1[i];
    if (di == 0.0) { if (ZERO) u[i] = 0.0; continue; }
    double res = f[i];
    for (int jj = b + (CLASSIC ? 1 : 0); jj < e; jj++) {
      const int j = A_j[jj];
      double uj;
      if (j >= ns && j < ne) {
        const bool swept = (DIR > 0) ? (j < i) : (j > i);
        uj = (ZERO && !swept) ? 0.0 : u[j];
      } else {
        uj = ZERO ? 0.0 : old[j];
      }
      res -= A_a[jj] * uj;
    }
    u[i] = CLASSIC ? res / di : (ZERO ? 0.0 : u[i]) + res / di;
  }
}

}  // namespace

int b200_gs_plan_destroy(b200_handle h, b200_gs_plan_s *p) {
  if (!p) return 0;
  B200_TRY(b200_dfree(h, p->perm)); B200_TRY(b200_dfree(h, p->level_off));
  B200_TRY(b200_dfree(h, p->chunks)); B200_TRY(b200_dfree(h, p->ctr));
  B200_TRY(b200_dfree(h, p->seg_off)); B200_TRY(b200_dfree(h, p->blk_seg)); B200_TRY(b200_dfree(h, p->old));
  B200_TRY(b200_dfree(h, p->slice_ptr)); B200_TRY(b200_dfree(h, p->sell_j)); B200_TRY(b200_dfree(h, p->sell_a));
  delete p;
  return 0;
}

// sort rows by key (stable in the row index): transpose of the n x nkeys pattern with one entry (i, key[i]) per row
static int sort_rows_by_key(b200_handle h, int n, int nkeys, int *key, int **perm, int **key_off) {
  b200_csr_s L;
  L.nrows = n; L.ncols = nkeys; L.nnz = n; L.owns = false;
  B200_TRY(b200_dalloc<int>(h, &L.i, (size_t)n + 1));
  gs_unit_rows_kernel<<<b200_grid((size_t)n + 1, 256), 256, 0, h->stream>>>(n, L.i);
  B200_LAUNCH_CHECK();
  L.j = key;
  b200_csr Lt = nullptr;
  B200_TRY(b200_csr_transpose(h, &L, &Lt));
  B200_TRY(b200_dfree(h, L.i));
  *perm = Lt->j; *key_off = Lt->i;                         // take ownership of the two arrays
  Lt->owns = false;
  B200_TRY(b200_csr_destroy(h, Lt));
  return 0;
}

// Level schedule of the n x n leading block of A (columns >= n are ghosts: always old values), T blocks.
int b200_gs_plan_create(b200_handle h, b200_csr A, int T, b200_gs_plan_s **out) {
  const int n = A->nrows;
  if (T < 1) B200_FAIL("gs plan: GSBlocks must be >= 1");
  b200_gs_plan_s *P = new b200_gs_plan_s();
  P->n = n; P->T = T; P->size = n / T; P->rest = n - P->size * T;
  P->maxblock = P->size + (P->rest ? 1 : 0);
  *out = P;
  if (n == 0) return 0;
  const char *fg = getenv("B200_GS_FORCE_GLOBAL");      // tests: "1" soft-barrier path, "2" CTA-per-block path
  if (P->maxblock <= GS_LANE_CAP && !(fg && (fg[0] == '1' || fg[0] == '2'))) {
    P->lane_path = true;
    if (T > 1) B200_TRY(b200_dalloc<double>(h, &P->old, (size_t)std::max(A->ncols, n)));
    const char *se = getenv("B200_GS_SELL");              // "0": plain CSR walk (tests compare both)
    // measured at 256^3 (profiles/r1_e_gs_*.txt): the sliced copy wins for short blocks (<= 128 rows: many lanes,
    // the CSR walk thrashes L1), the plain CSR walk for longer ones; "1" forces the sliced copy
    if (T >= 32 && !(se && se[0] == '0') && (P->maxblock <= 128 || (se && se[0] == '1'))) {
      const int nwarps = (T + 31) / 32, maxm = P->maxblock;
      const long long nsl = (long long)nwarps * maxm;
      if (nsl < (1LL << 30)) {
        P->maxm = maxm;
        B200_TRY(b200_dalloc<int>(h, &P->slice_ptr, (size_t)nsl + 1));
        gs_sell_width_kernel<<<b200_grid((size_t)nsl, 8), 256, 0, h->stream>>>(nwarps, maxm, T, P->size, P->rest, A->i, P->slice_ptr);
        B200_LAUNCH_CHECK();
        B200_TRY(b200_exclusive_scan_inplace(h, P->slice_ptr, (size_t)nsl + 1));
        int total = 0;
        B200_CUDA(cudaMemcpyAsync(&total, P->slice_ptr + nsl, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        B200_CUDA(cudaStreamSynchronize(h->stream));
        if (total < 0 || (double)total > 1.6 * (double)A->nnz + 1024.0) {        // padding too expensive: keep the CSR walk
          B200_TRY(b200_dfree(h, P->slice_ptr)); P->slice_ptr = nullptr; P->maxm = 0;
        } else {
          B200_TRY(b200_dalloc<int>(h, &P->sell_j, (size_t)total + 32));
          B200_TRY(b200_dalloc<double>(h, &P->sell_a, (size_t)total + 32));
          gs_sell_fill_kernel<<<b200_grid((size_t)nsl, 8), 256, 0, h->stream>>>(nwarps, maxm, T, P->size, P->rest, A->i, A->j, A->a,
                                                                                P->slice_ptr, P->sell_j, P->sell_a);
          B200_LAUNCH_CHECK();
        }
      }
    }
    return 0;
  }
  int *d_flag = nullptr, *indeg = nullptr, *level = nullptr, *fr[2] = {nullptr, nullptr}, *cnt = nullptr;
  B200_TRY(b200_dalloc<int>(h, &d_flag, 1));
  B200_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
  gs_symcheck_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, A->i, A->j, d_flag);
  B200_LAUNCH_CHECK();
  int nonsym = 0;
  B200_CUDA(cudaMemcpyAsync(&nonsym, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr Tp = nullptr;
  if (nonsym) {                                          // successors also come from the transposed pattern
    b200_csr_s pat = *A;
    pat.a = nullptr; pat.owns = false; pat.blk_row = pat.blk_ent = pat.blk_meta = nullptr; pat.gs = nullptr;
    B200_TRY(b200_csr_transpose(h, &pat, &Tp));
  }
  const int *T_i = Tp ? Tp->i : nullptr, *T_j = Tp ? Tp->j : nullptr;
  B200_TRY(b200_dalloc<int>(h, &indeg, n));
  B200_TRY(b200_dalloc<int>(h, &level, n));
  B200_TRY(b200_dalloc<int>(h, &fr[0], n));
  B200_TRY(b200_dalloc<int>(h, &fr[1], n));
  B200_TRY(b200_dalloc<int>(h, &cnt, 4));
  const int big = 0x7fffffff;
  B200_CUDA(cudaMemsetAsync(cnt, 0, 3 * sizeof(int), h->stream));
  B200_CUDA(cudaMemcpyAsync(cnt + 3, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  gs_indeg_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, P->size, P->rest, A->i, A->j, T_i, T_j, indeg, level, fr[0], cnt);
  B200_LAUNCH_CHECK();
  const int grid = std::min(b200_grid(n, 256), h->num_sm * 4);
  int nlevels = big, r = 0;
  while (nlevels == big) {
    for (int k = 0; k < 128; k++, r++) {
      gs_peel_kernel<<<grid, 256, 0, h->stream>>>(n, P->size, P->rest, r, A->i, A->j, T_i, T_j, indeg, level, fr[r & 1],
                                                  fr[(r + 1) & 1], cnt, cnt + 3);
      B200_LAUNCH_CHECK();
    }
    B200_CUDA(cudaMemcpyAsync(&nlevels, cnt + 3, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  P->nlevels = nlevels;
  const int D = nlevels;
  P->block_path = P->maxblock <= GS_BLOCK_CAP && (long long)T * D <= (1LL << 26) && !(fg && fg[0] == '1');
  if (P->block_path) {
    const int K = T * D;
    gs_block_key_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, P->size, P->rest, D, level);
    B200_LAUNCH_CHECK();
    int *key_off = nullptr, *pos = nullptr;
    B200_TRY(sort_rows_by_key(h, n, K, level, &P->perm, &key_off));
    B200_TRY(b200_dalloc<int>(h, &pos, (size_t)K + 1));
    gs_seg_flag_kernel<<<b200_grid((size_t)K + 1, 256), 256, 0, h->stream>>>(K, key_off, pos);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_exclusive_scan_inplace(h, pos, (size_t)K + 1));
    B200_CUDA(cudaMemcpyAsync(&P->nseg, pos + K, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dalloc<int>(h, &P->seg_off, (size_t)P->nseg + 1));
    B200_TRY(b200_dalloc<int>(h, &P->blk_seg, (size_t)T + 1));
    gs_seg_scatter_kernel<<<b200_grid((size_t)K + 1, 256), 256, 0, h->stream>>>(K, T, D, key_off, pos, P->seg_off, P->blk_seg);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_dfree(h, key_off)); B200_TRY(b200_dfree(h, pos));
  } else {
    B200_TRY(sort_rows_by_key(h, n, D, level, &P->perm, &P->level_off));
    std::vector<int> off((size_t)D + 1);
    B200_CUDA(cudaMemcpyAsync(off.data(), P->level_off, sizeof(int) * off.size(), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    if (off[D] != n) B200_FAIL("gs plan: level schedule does not cover every row");
    std::vector<int> ch;
    for (int l = 0; l < D; l++)
      for (int p0 = off[l]; p0 < off[l + 1]; p0 += GS_NT) {
        ch.push_back(p0); ch.push_back(std::min(p0 + GS_NT, off[l + 1])); ch.push_back(l); ch.push_back(0);
      }
    P->nchunks = (int)(ch.size() / 4);
    B200_TRY(b200_dalloc<int>(h, &P->chunks, ch.size()));
    B200_CUDA(cudaMemcpyAsync(P->chunks, ch.data(), sizeof(int) * ch.size(), cudaMemcpyHostToDevice, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dalloc<int>(h, &P->ctr, (size_t)D + 1));
  }
  if (T > 1) B200_TRY(b200_dalloc<double>(h, &P->old, (size_t)std::max(A->ncols, n)));
  B200_TRY(b200_dfree(h, d_flag)); B200_TRY(b200_dfree(h, indeg)); B200_TRY(b200_dfree(h, level));
  B200_TRY(b200_dfree(h, fr[0])); B200_TRY(b200_dfree(h, fr[1])); B200_TRY(b200_dfree(h, cnt));
  if (Tp) B200_TRY(b200_csr_destroy(h, Tp));
  return 0;
}

int b200_gs_plan_levels(b200_gs_plan_s *p) { return p ? p->nlevels : 0; }
int b200_gs_plan_blocks(b200_gs_plan_s *p) { return p ? p->T : 0; }

// dir +1 / -1; classic: types 3/4/6, else the l1 variants (l1 required); zero: u == 0 on entry.
// refresh_old: copy the iterate to the pre-sweep buffer first (T > 1).  The reference does this once per
// relaxation call, so the second half of the symmetric types 6/8 keeps reading the copy made before the
// forward half (par_relax.c:3548-3640).
int b200_gs_sweep(b200_handle h, b200_gs_plan_s *P, b200_csr A, int dir, bool classic, bool zero, bool refresh_old,
                  const double *f, const double *l1, double *u) {
  if (P->n == 0) return 0;
  if (!classic && !l1) B200_FAIL("gs sweep: l1 norms required for relax types 8/13/14");
  const double *old = u;
  if (P->T > 1) {                                        // pre-sweep copy of [owned | ghost]
    const size_t w = (size_t)std::max(A->ncols, P->n);
    if (refresh_old) {
      if (zero) B200_CUDA(cudaMemsetAsync(P->old, 0, sizeof(double) * w, h->stream));
      else B200_CUDA(cudaMemcpyAsync(P->old, u, sizeof(double) * w, cudaMemcpyDeviceToDevice, h->stream));
    }
    old = P->old;
  }
  if (P->lane_path && P->sell_j) {
    const int nwarps = (P->T + 31) / 32;
#define GS_SELL(D, C, Z)                                                                                                   \
    gs_sell_kernel<D, C, Z><<<b200_grid(nwarps, 2), 64, 0, h->stream>>>(P->n, P->T, P->size, P->rest, P->maxm, P->slice_ptr, \
                                                                         P->sell_j, P->sell_a, f, l1, old, u)
    if (dir > 0) {
      if (classic) { if (zero) GS_SELL(1, true, true); else GS_SELL(1, true, false); }
      else         { if (zero) GS_SELL(1, false, true); else GS_SELL(1, false, false); }
    } else {
      if (classic) { if (zero) GS_SELL(-1, true, true); else GS_SELL(-1, true, false); }
      else         { if (zero) GS_SELL(-1, false, true); else GS_SELL(-1, false, false); }
    }
#undef GS_SELL
    B200_LAUNCH_CHECK();
    return 0;
  }
  if (P->lane_path) {
#define GS_LANE(D, C, Z)                                                                                                   \
    gs_lane_kernel<D, C, Z><<<b200_grid(P->T, 64), 64, 0, h->stream>>>(P->n, P->T, P->size, P->rest, A->i, A->j, A->a, f,  \
                                                                        l1, old, u)
    if (dir > 0) {
      if (classic) { if (zero) GS_LANE(1, true, true); else GS_LANE(1, true, false); }
      else         { if (zero) GS_LANE(1, false, true); else GS_LANE(1, false, false); }
    } else {
      if (classic) { if (zero) GS_LANE(-1, true, true); else GS_LANE(-1, true, false); }
      else         { if (zero) GS_LANE(-1, false, true); else GS_LANE(-1, false, false); }
    }
#undef GS_LANE
    B200_LAUNCH_CHECK();
    return 0;
  }
  if (P->block_path) {
    const int NT = P->maxblock <= 2048 ? 64 : (P->maxblock <= 8192 ? 128 : 256);
    const size_t smem = sizeof(double) * (size_t)P->maxblock;
    const int grid = std::min(P->T, h->num_sm * 32);
    // columns >= n (ghosts) are read from `old` too: with T > 1 the copy holds owned rows only, so ghosts
    // must stay addressable -> multi-rank callers pass T == 1 per rank or a full-width copy
#define GS_BLOCK(D, C, Z)                                                                                                  \
    {                                                                                                                      \
      if (smem > 48 * 1024)                                                                                                \
        B200_CUDA(cudaFuncSetAttribute(gs_block_kernel<D, C, Z>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      gs_block_kernel<D, C, Z><<<grid, NT, smem, h->stream>>>(P->n, P->T, P->size, P->rest, P->blk_seg, P->seg_off, P->perm, \
                                                               A->i, A->j, A->a, f, l1, old, u);                           \
    }
    if (dir > 0) {
      if (classic) { if (zero) GS_BLOCK(1, true, true) else GS_BLOCK(1, true, false) }
      else         { if (zero) GS_BLOCK(1, false, true) else GS_BLOCK(1, false, false) }
    } else {
      if (classic) { if (zero) GS_BLOCK(-1, true, true) else GS_BLOCK(-1, true, false) }
      else         { if (zero) GS_BLOCK(-1, false, true) else GS_BLOCK(-1, false, false) }
    }
#undef GS_BLOCK
    B200_LAUNCH_CHECK();
    return 0;
  }
  B200_CUDA(cudaMemsetAsync(P->ctr, 0, sizeof(int) * ((size_t)P->nlevels + 1), h->stream));
  const int grid = std::min(P->nchunks, h->num_sm * 16);
  const double *oldg = P->T > 1 ? P->old : nullptr;
#define GS_LAUNCH(D, C, Z)                                                                                                 \
  gs_sweep_kernel<D, C, Z><<<grid, GS_NT, 0, h->stream>>>(P->n, P->size, P->rest, P->nchunks, P->nlevels,                 \
                                                          (const int4 *)P->chunks, P->level_off, P->perm, A->i, A->j,     \
                                                          A->a, f, l1, oldg, u, P->ctr)
  if (dir > 0) {
    if (classic) { if (zero) GS_LAUNCH(1, true, true); else GS_LAUNCH(1, true, false); }
    else         { if (zero) GS_LAUNCH(1, false, true); else GS_LAUNCH(1, false, false); }
  } else {
    if (classic) { if (zero) GS_LAUNCH(-1, true, true); else GS_LAUNCH(-1, true, false); }
    else         { if (zero) GS_LAUNCH(-1, false, true); else GS_LAUNCH(-1, false, false); }
  }
#undef GS_LAUNCH
  B200_LAUNCH_CHECK();
  return 0;
}

// one relaxation call of the Gauss-Seidel family in place (par_relax.c cases 3/4/6/8/13/14, relax_points 0)
int b200_gs_relax(b200_handle h, b200_gs_plan_s *P, b200_csr A, int type, bool zero, const double *f, const double *l1, double *u) {
  const bool classic = type == 3 || type == 4 || type == 6;
  bool z = zero, first = true;
  if (type == 3 || type == 13 || type == 6 || type == 8) {
    B200_TRY(b200_gs_sweep(h, P, A, +1, classic, z, first, f, l1, u));
    z = false; first = false;
  }
  if (type == 4 || type == 14 || type == 6 || type == 8) B200_TRY(b200_gs_sweep(h, P, A, -1, classic, z, first, f, l1, u));
  return 0;
}

// hypre_BoomerAMGRelax (par_relax.c:30) for the Gauss-Seidel family on one rank, relax_points 0,
// relax_weight = omega = 1, `blocks` Gauss-Seidel blocks.  The level schedule is cached on the matrix.
extern "C" int b200_relax_gs(b200_handle h, b200_csr A, int relax_type, int blocks, const double *d_f, const double *d_l1,
                             double *d_u) {
  if (!A || !A->a) B200_FAIL("relax: matrix with values required");
  if (A->nrows != A->ncols) B200_FAIL("relax: square matrix required");
  const bool classic = relax_type == 3 || relax_type == 4 || relax_type == 6;
  if (!classic && relax_type != 8 && relax_type != 13 && relax_type != 14)
    B200_FAIL("relax: Gauss-Seidel types are 3, 4, 6 (classic) and 8, 13, 14 (l1)");
  if (A->gs && b200_gs_plan_blocks(A->gs) != blocks) { B200_TRY(b200_gs_plan_destroy(h, A->gs)); A->gs = nullptr; }
  if (!A->gs) B200_TRY(b200_gs_plan_create(h, A, blocks, &A->gs));
  return b200_gs_relax(h, A->gs, A, relax_type, false, d_f, d_l1, d_u);
}
