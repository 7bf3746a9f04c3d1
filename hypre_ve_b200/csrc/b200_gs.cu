// b200_gs.cu -- hybrid Gauss-Seidel relaxation (hypre_BoomerAMGRelax types 3/4/6 and the l1 variants
// 8/13/14, parcsr_ls/par_relax.c:1875-2265, :3492-4091, :4340-5124) as an exact, level-scheduled sweep.
//
// Decomposition (what "hybrid" means, par_relax.c:4400-4412): Gauss-Seidel inside a block of rows,
// Jacobi (old values) across blocks.  Here ONE block per rank (= the reference with OMP_NUM_THREADS=1,
// one MPI rank per GPU): inside the rank the sweep is the reference's sequential loop, bit for bit --
// row i reads the NEW value of every neighbour that precedes it in sweep order and the OLD value of the
// others, and each row's sum runs over its entries in storage order in one thread.
//
// How a sequential sweep runs on 148 SMs:
//   setup  -- level[i] = 1 + max(level[j] : j < i, a_ij != 0 or a_ji != 0), by frontier peeling
//             (Kahn's algorithm, one small kernel per level; no full pass per level); rows sorted by
//             (level, row) -> perm; levels are cut into chunks of <= 128 rows.  Levels built on the
//             symmetrised pattern serve both directions: descending level order is a valid order for the
//             backward sweep, and a row never reads a neighbour that could already have been rewritten.
//   sweep  -- ONE kernel: CTAs draw chunk tickets in level order, prefetch their rows (indices, values,
//             old neighbour values) and only then wait on the completion counter of the previous level
//             (a soft barrier: ticket order makes it deadlock free without a cooperative launch); new
//             values are read with ld.global.cg (L2) after the acquire.
// The sweep is bound by the dependency chain (#levels x one L2 round trip), not by HBM: the 7-pt 256^3
// grid in lexicographic order has 766 levels.  DESIGN.md section 3 gives the model and measurements.
#include "b200_internal.h"
#include <algorithm>

struct b200_gs_plan_s {
  int n = 0, nlevels = 0, nchunks = 0;
  int *perm = nullptr;        // [n] rows sorted by (level, row)
  int *level_off = nullptr;   // [nlevels+1] first position of each level in perm
  int *chunks = nullptr;      // int4 per chunk {pos0, pos1, level, 0}
  int *ctr = nullptr;         // [1 + nlevels] ticket counter, rows finished per level
};

namespace {

constexpr int GS_NT = 128;     // threads per CTA = rows per chunk
constexpr int GS_PRE = 8;      // row entries prefetched into registers before the wait

// indeg[i] = number of sweep-order predecessors of row i on the symmetrised pattern; rows with none
// open level 0.  T_i/T_j = pattern of A^T when A's pattern is not symmetric (else null).
__global__ void gs_indeg_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const int *__restrict__ T_i,
                                const int *__restrict__ T_j, int *__restrict__ indeg, int *__restrict__ level,
                                int *__restrict__ frontier, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = 0;
  for (int jj = A_i[i]; jj < A_i[i + 1]; jj++) d += (A_j[jj] < i);
  if (T_i) for (int jj = T_i[i]; jj < T_i[i + 1]; jj++) d += (T_j[jj] < i);
  indeg[i] = d;
  if (d == 0) { level[i] = 0; frontier[atomicAdd(cnt, 1)] = i; }
}
// one peeling round: rows of level r release their successors; cnt is a ring of three counters
__global__ void gs_peel_kernel(int n, int r, const int *__restrict__ A_i, const int *__restrict__ A_j, const int *__restrict__ T_i,
                               const int *__restrict__ T_j, int *__restrict__ indeg, int *__restrict__ level,
                               const int *__restrict__ fin, int *__restrict__ fout, int *__restrict__ cnt, int *__restrict__ nlevels) {
  const int nin = cnt[r % 3];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    cnt[(r + 2) % 3] = 0;                               // becomes the output counter of the next round
    if (nin == 0) atomicMin(nlevels, r);
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nin; idx += gridDim.x * blockDim.x) {
    const int i = fin[idx];
    for (int jj = A_i[i]; jj < A_i[i + 1]; jj++) {
      const int j = A_j[jj];
      if (j > i && j < n && atomicSub(&indeg[j], 1) == 1) { level[j] = r + 1; fout[atomicAdd(&cnt[(r + 1) % 3], 1)] = j; }
    }
    if (T_i)
      for (int jj = T_i[i]; jj < T_i[i + 1]; jj++) {
        const int j = T_j[jj];
        if (j > i && j < n && atomicSub(&indeg[j], 1) == 1) { level[j] = r + 1; fout[atomicAdd(&cnt[(r + 1) % 3], 1)] = j; }
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

__device__ __forceinline__ int ld_volatile(const int *p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One sweep.  DIR = +1 forward (types 3, 13), -1 backward (4, 14).  CLASSIC: u_i = res / a_ii with the
// diagonal (stored first) skipped, else u_i += res / l1_i over the whole row.  ZERO: the iterate is
// known to be 0 on entry (first sweep of a level in a cycle): old values are not read.
template <int DIR, bool CLASSIC, bool ZERO>
__global__ void __launch_bounds__(GS_NT)
gs_sweep_kernel(int n, int nchunks, int nlevels, const int4 *__restrict__ chunks, const int *__restrict__ level_off,
                const int *__restrict__ perm, const int *__restrict__ A_i, const int *__restrict__ A_j,
                const double *__restrict__ A_a, const double *__restrict__ f, const double *__restrict__ l1,
                double *u, int *ctr) {
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
    int i = 0, b = 0, e = 0;
    double fi = 0.0, di = 0.0, ui = 0.0;
    int pj[GS_PRE];
    double pa[GS_PRE], pu[GS_PRE];
    if (active) {
      i = perm[p];
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
          const bool dep = (DIR > 0) ? (j < i) : (j > i && j < n);
          if (!dep && !ZERO) pu[k] = __ldcg(u + j);      // old value: cannot be rewritten before this row is done
        }
      }
    }
    // ---- soft barrier on the previous level ------------------------------------------------------------
    const int wl = c.z - DIR;
    if (threadIdx.x == 0 && wl >= 0 && wl < nlevels) {
      const int target = level_off[wl + 1] - level_off[wl];
      unsigned ns = 20;
      while (ld_volatile(done + wl) < target) { __nanosleep(ns); if (ns < 400) ns += ns; }
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
          const bool dep = (DIR > 0) ? (j < i) : (j > i && j < n);
          const double uj = dep ? __ldcg(u + j) : pu[k];
          res -= pa[k] * uj;
        }
      }
      for (int jj = b + GS_PRE; jj < e; jj++) {
        const int j = A_j[jj];
        const bool dep = (DIR > 0) ? (j < i) : (j > i && j < n);
        const double uj = (dep || !ZERO) ? __ldcg(u + j) : 0.0;
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

}  // namespace

int b200_gs_plan_destroy(b200_handle h, b200_gs_plan_s *p) {
  if (!p) return 0;
  B200_TRY(b200_dfree(h, p->perm)); B200_TRY(b200_dfree(h, p->level_off));
  B200_TRY(b200_dfree(h, p->chunks)); B200_TRY(b200_dfree(h, p->ctr));
  delete p;
  return 0;
}

// Level schedule of the n x n leading block of A (columns >= n are ghosts: always old values).
int b200_gs_plan_create(b200_handle h, b200_csr A, b200_gs_plan_s **out) {
  const int n = A->nrows;
  b200_gs_plan_s *P = new b200_gs_plan_s();
  P->n = n;
  *out = P;
  if (n == 0) return 0;
  int *d_flag = nullptr, *indeg = nullptr, *level = nullptr, *fr[2] = {nullptr, nullptr}, *cnt = nullptr;
  B200_TRY(b200_dalloc<int>(h, &d_flag, 1));
  B200_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
  gs_symcheck_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, A->i, A->j, d_flag);
  B200_LAUNCH_CHECK();
  int nonsym = 0;
  B200_CUDA(cudaMemcpyAsync(&nonsym, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr T = nullptr;
  if (nonsym) {                                          // successors also come from the transposed pattern
    b200_csr_s pat = *A;
    pat.a = nullptr; pat.owns = false; pat.blk_row = pat.blk_ent = pat.blk_meta = nullptr;
    B200_TRY(b200_csr_transpose(h, &pat, &T));
  }
  const int *T_i = T ? T->i : nullptr, *T_j = T ? T->j : nullptr;
  B200_TRY(b200_dalloc<int>(h, &indeg, n));
  B200_TRY(b200_dalloc<int>(h, &level, n));
  B200_TRY(b200_dalloc<int>(h, &fr[0], n));
  B200_TRY(b200_dalloc<int>(h, &fr[1], n));
  B200_TRY(b200_dalloc<int>(h, &cnt, 4));
  const int big = 0x7fffffff;
  B200_CUDA(cudaMemsetAsync(cnt, 0, 3 * sizeof(int), h->stream));
  B200_CUDA(cudaMemcpyAsync(cnt + 3, &big, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  gs_indeg_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, A->i, A->j, T_i, T_j, indeg, level, fr[0], cnt);
  B200_LAUNCH_CHECK();
  const int grid = std::min(b200_grid(n, 256), h->num_sm * 4);
  int nlevels = big, r = 0;
  while (nlevels == big) {
    for (int k = 0; k < 128; k++, r++) {
      gs_peel_kernel<<<grid, 256, 0, h->stream>>>(n, r, A->i, A->j, T_i, T_j, indeg, level, fr[r & 1], fr[(r + 1) & 1], cnt, cnt + 3);
      B200_LAUNCH_CHECK();
    }
    B200_CUDA(cudaMemcpyAsync(&nlevels, cnt + 3, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  P->nlevels = nlevels;
  // rows sorted by (level, row): transpose of the n x nlevels pattern with one entry (i, level[i]) per row
  {
    b200_csr_s L;
    L.nrows = n; L.ncols = nlevels; L.nnz = n; L.owns = false;
    B200_TRY(b200_dalloc<int>(h, &L.i, (size_t)n + 1));
    gs_unit_rows_kernel<<<b200_grid((size_t)n + 1, 256), 256, 0, h->stream>>>(n, L.i);
    B200_LAUNCH_CHECK();
    L.j = level;
    b200_csr Lt = nullptr;
    B200_TRY(b200_csr_transpose(h, &L, &Lt));
    B200_TRY(b200_dfree(h, L.i));
    P->perm = Lt->j; P->level_off = Lt->i;               // take ownership of the two arrays
    Lt->owns = false;
    B200_TRY(b200_csr_destroy(h, Lt));
  }
  std::vector<int> off((size_t)nlevels + 1);
  B200_CUDA(cudaMemcpyAsync(off.data(), P->level_off, sizeof(int) * off.size(), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  if (off[nlevels] != n) B200_FAIL("gs plan: level schedule does not cover every row (cyclic dependency?)");
  std::vector<int> ch;
  for (int l = 0; l < nlevels; l++)
    for (int p0 = off[l]; p0 < off[l + 1]; p0 += GS_NT) {
      ch.push_back(p0); ch.push_back(std::min(p0 + GS_NT, off[l + 1])); ch.push_back(l); ch.push_back(0);
    }
  P->nchunks = (int)(ch.size() / 4);
  B200_TRY(b200_dalloc<int>(h, &P->chunks, ch.size()));
  B200_CUDA(cudaMemcpyAsync(P->chunks, ch.data(), sizeof(int) * ch.size(), cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dalloc<int>(h, &P->ctr, (size_t)nlevels + 1));
  B200_TRY(b200_dfree(h, d_flag)); B200_TRY(b200_dfree(h, indeg)); B200_TRY(b200_dfree(h, level));
  B200_TRY(b200_dfree(h, fr[0])); B200_TRY(b200_dfree(h, fr[1])); B200_TRY(b200_dfree(h, cnt));
  if (T) B200_TRY(b200_csr_destroy(h, T));
  return 0;
}

int b200_gs_plan_levels(b200_gs_plan_s *p) { return p ? p->nlevels : 0; }

// dir +1 / -1; classic: types 3/4/6, else the l1 variants (d_l1 required); zero: u == 0 on entry
int b200_gs_sweep(b200_handle h, b200_gs_plan_s *P, b200_csr A, int dir, bool classic, bool zero, const double *f,
                  const double *l1, double *u) {
  if (P->n == 0) return 0;
  if (!classic && !l1) B200_FAIL("gs sweep: l1 norms required for relax types 8/13/14");
  B200_CUDA(cudaMemsetAsync(P->ctr, 0, sizeof(int) * ((size_t)P->nlevels + 1), h->stream));
  const int grid = std::min(P->nchunks, h->num_sm * 16);
#define GS_LAUNCH(D, C, Z)                                                                                               \
  gs_sweep_kernel<D, C, Z><<<grid, GS_NT, 0, h->stream>>>(P->n, P->nchunks, P->nlevels, (const int4 *)P->chunks,      \
                                                          P->level_off, P->perm, A->i, A->j, A->a, f, l1, u, P->ctr)
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

// hypre_BoomerAMGRelax (par_relax.c:30) for the Gauss-Seidel family on one rank, relax_points 0,
// relax_weight = omega = 1.  The level schedule is cached on the matrix.
extern "C" int b200_relax_gs(b200_handle h, b200_csr A, int relax_type, const double *d_f, const double *d_l1, double *d_u) {
  if (!A || !A->a) B200_FAIL("relax: matrix with values required");
  if (A->nrows != A->ncols) B200_FAIL("relax: square matrix required");
  const bool classic = relax_type == 3 || relax_type == 4 || relax_type == 6;
  if (!classic && relax_type != 8 && relax_type != 13 && relax_type != 14)
    B200_FAIL("relax: Gauss-Seidel types are 3, 4, 6 (classic) and 8, 13, 14 (l1)");
  if (!A->gs) B200_TRY(b200_gs_plan_create(h, A, &A->gs));
  if (relax_type == 3 || relax_type == 13 || relax_type == 6 || relax_type == 8)
    B200_TRY(b200_gs_sweep(h, A->gs, A, +1, classic, false, d_f, d_l1, d_u));
  if (relax_type == 4 || relax_type == 14 || relax_type == 6 || relax_type == 8)
    B200_TRY(b200_gs_sweep(h, A->gs, A, -1, classic, false, d_f, d_l1, d_u));
  return 0;
}
