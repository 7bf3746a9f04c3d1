// b200_agg.cu -- aggressive coarsening and multipass interpolation, one rank.
//
// Reference: par_amg_setup.c:1239-1256 and :1590-1605 (driver), hypre_BoomerAMGCreate2ndSHost
// (par_strength.c:1729-2918, num_paths 1), hypre_BoomerAMGCoarsenPMIS with CF_init 3 (par_coarsen.c:2322-2326,
// :2420), hypre_BoomerAMGCorrectCFMarker (par_strength.c:2957-2974), hypre_BoomerAMGBuildMultipass
// (par_multi_interp.c:16-2061, weight_option 0, no truncation).
//
// How the sequential reference maps to the device without changing a bit of the result:
//   * S2 (distance-two strength graph of the C points) is a pattern product: row ic = first-touch union, over the
//     strong neighbours i2 of C point ic, of [i2 itself if C] ++ [C points in the row of S of i2].  That is the
//     Gustavson product Sc * M of two pattern matrices, so it runs on the order-preserving SpGEMM kernels; the
//     entry ic itself (never inserted by the reference) is dropped afterwards, which leaves the order of the others.
//   * pass numbers are a breadth-first search from the C points through S: one kernel per pass.
//   * pass 1 rows: one thread walks the row of A (S is a subsequence of it), exactly the reference loop.
//   * pass p >= 2 rows are the rows of  Ap * P_{p-1}  with Ap = the entries a_ij of the strong neighbours that got
//     their formula in pass p-1: again the order-preserving SpGEMM (columns in first-touch order, products
//     added in (j, k) order); the scaling factor needs sum_C / sum_N accumulated product by product in the
//     reference's order, which a second kernel recomputes sequentially per row.
#include "b200_internal.h"
#include <algorithm>
#include <utility>

int b200_csr_multiply_ex(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C, b200_csr *out);
int b200_coarse_map(b200_handle h, int n, const int *d_cf, int **f2c_out, int *ncoarse);
int b200_pmis_rows_init(b200_handle h, b200_csr S, int seed, long long first_row, int cf_init, int *d_cf, int *iterations);

namespace {

constexpr int TBA = 256;
constexpr int MAX_PASSES = 10;          // par_multi_interp.c:102

__global__ void c2f_kernel(int n, const int *__restrict__ cf, const int *__restrict__ f2c, int *__restrict__ c2f) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cf[i] > 0) c2f[f2c[i]] = i;
}
// Sc: rows of S that belong to C points (all columns kept)
__global__ void sc_count_kernel(int nc, const int *__restrict__ c2f, const int *__restrict__ S_i, int *__restrict__ cnt) {
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic > nc) return;
  cnt[ic] = ic < nc ? S_i[c2f[ic] + 1] - S_i[c2f[ic]] : 0;
}
__global__ void sc_fill_kernel(int nc, const int *__restrict__ c2f, const int *__restrict__ S_i, const int *__restrict__ S_j,
                               const int *__restrict__ Sc_i, int *__restrict__ Sc_j, double *__restrict__ Sc_a) {
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= nc) return;
  const int b = S_i[c2f[ic]], d = Sc_i[ic], len = Sc_i[ic + 1] - d;
  for (int k = 0; k < len; k++) { Sc_j[d + k] = S_j[b + k]; Sc_a[d + k] = 1.0; }
}
// M: row i2 = [coarse id of i2 if C] ++ [coarse ids of the C points in row i2 of S]
__global__ void m_count_kernel(int n, const int *__restrict__ cf, const int *__restrict__ S_i, const int *__restrict__ S_j,
                               int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int c = 0;
  if (i < n) {
    c = cf[i] > 0 ? 1 : 0;
    for (int jj = S_i[i]; jj < S_i[i + 1]; jj++) c += cf[S_j[jj]] > 0;
  }
  cnt[i] = c;
}
__global__ void m_fill_kernel(int n, const int *__restrict__ cf, const int *__restrict__ f2c, const int *__restrict__ S_i,
                              const int *__restrict__ S_j, const int *__restrict__ M_i, int *__restrict__ M_j,
                              double *__restrict__ M_a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = M_i[i];
  if (cf[i] > 0) { M_j[d] = f2c[i]; M_a[d++] = 1.0; }
  for (int jj = S_i[i]; jj < S_i[i + 1]; jj++) {
    const int i3 = S_j[jj];
    if (cf[i3] > 0) { M_j[d] = f2c[i3]; M_a[d++] = 1.0; }
  }
}
__global__ void offdiag_count_kernel(int n, int first, const int *__restrict__ C_i, const int *__restrict__ C_j, int *__restrict__ cnt) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  int c = 0;
  if (r < n) for (int jj = C_i[r]; jj < C_i[r + 1]; jj++) c += C_j[jj] != first + r;
  cnt[r] = c;
}
__global__ void offdiag_fill_kernel(int n, int first, const int *__restrict__ C_i, const int *__restrict__ C_j,
                                    const int *__restrict__ O_i, int *__restrict__ O_j) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int d = O_i[r];
  for (int jj = C_i[r]; jj < C_i[r + 1]; jj++) if (C_j[jj] != first + r) O_j[d++] = C_j[jj];
}
// hypre_BoomerAMGCorrectCFMarker: C points of the first coarsening take the marker of the second one
__global__ void correct_cf_kernel(int n, const int *__restrict__ f2c, const int *__restrict__ cfn, int *__restrict__ cf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cf[i] > 0) cf[i] = (cf[i] == 1) ? cfn[f2c[i]] : 1;
}

// ---- multipass ---------------------------------------------------------------------------------------------
__global__ void assign_init_kernel(int n, const int *__restrict__ cf, int *__restrict__ assigned) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) assigned[i] = cf[i] == 1 ? 0 : -1;
}
// pass p: unassigned F points (cf == -1) with a strong neighbour of pass p-1 (par_multi_interp.c:404-510)
__global__ void assign_pass_kernel(int n, int p, const int *__restrict__ cf, const int *__restrict__ S_i,
                                   const int *__restrict__ S_j, const int *__restrict__ assigned_in, int *__restrict__ assigned_out,
                                   int *__restrict__ remaining) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = assigned_in[i];
  if (cf[i] == -1 && a == -1) {
    for (int jj = S_i[i]; jj < S_i[i + 1]; jj++)
      if (assigned_in[S_j[jj]] == p - 1) { a = p; break; }
    if (a == -1) atomicAdd(remaining, 1);
  }
  assigned_out[i] = a;
}
// rows of pass p: number of strong neighbours that were assigned in pass p-1 (p == 1: strong C neighbours)
__global__ void nbr_count_kernel(int n, int p, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                 const int *__restrict__ assigned, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int c = 0;
  if (i < n && assigned[i] == p)
    for (int jj = S_i[i]; jj < S_i[i + 1]; jj++) c += assigned[S_j[jj]] == p - 1;
  cnt[i] = c;
}
// pass 1 (par_multi_interp.c:1600-1660): the strong C entries of the row of A, in A's order, times
// alfa = -sum_N / (sum_C * a_ii)
__global__ void pass1_fill_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                                  const int *__restrict__ S_i, const int *__restrict__ S_j, const int *__restrict__ cf,
                                  const int *__restrict__ f2c, const int *__restrict__ assigned, const int *__restrict__ P_i,
                                  int *__restrict__ P_j, double *__restrict__ P_a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || assigned[i] != 1) return;
  int s = S_i[i];
  const int se = S_i[i + 1], d0 = P_i[i];
  int d = d0;
  double sum_C = 0, sum_N = 0;
  for (int j = A_i[i] + 1; j < A_i[i + 1]; j++) {
    const int j1 = A_j[j];
    const double a = A_a[j];
    if (cf[j1] != -3) sum_N += a;
    const bool strong = s < se && S_j[s] == j1;      // S is a subsequence of the row of A
    if (strong) s++;
    if (strong && cf[j1] == 1) { P_a[d] = a; P_j[d++] = f2c[j1]; sum_C += a; }
  }
  const double diagonal = A_a[A_i[i]];
  double alfa = 1.0;    // the reference keeps the previous row's factor when sum_C * a_ii == 0 (a degenerate row)
  if (sum_C * diagonal != 0) alfa = -sum_N / (sum_C * diagonal);
  for (int k = d0; k < d; k++) P_a[k] *= alfa;
}
// Ap for pass p: (j1, a_ij) of the strong neighbours assigned in pass p-1, in A's order
__global__ void ap_fill_kernel(int n, int p, const int *__restrict__ A_i, const int *__restrict__ A_j,
                               const double *__restrict__ A_a, const int *__restrict__ S_i, const int *__restrict__ S_j,
                               const int *__restrict__ assigned, const int *__restrict__ Ap_i, int *__restrict__ Ap_j,
                               double *__restrict__ Ap_a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || assigned[i] != p) return;
  int s = S_i[i], d = Ap_i[i];
  const int se = S_i[i + 1];
  for (int j = A_i[i] + 1; j < A_i[i + 1]; j++) {
    const int j1 = A_j[j];
    const bool strong = s < se && S_j[s] == j1;
    if (strong) s++;
    if (strong && assigned[j1] == p - 1) { Ap_j[d] = j1; Ap_a[d++] = A_a[j]; }
  }
}
// scaling of the pass-p rows (par_multi_interp.c:1805-1860): sum_C, sum_N accumulated in the reference's order
__global__ void passp_scale_kernel(int n, int p, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                   const double *__restrict__ A_a, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                   const int *__restrict__ cf, const int *__restrict__ assigned, const int *__restrict__ B_i,
                                   const double *__restrict__ B_a, const int *__restrict__ C_i, double *__restrict__ C_a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || assigned[i] != p) return;
  int s = S_i[i];
  const int se = S_i[i + 1];
  double sum_C = 0, sum_N = 0, alfa = 1.0;
  for (int j = A_i[i] + 1; j < A_i[i + 1]; j++) {
    const int j1 = A_j[j];
    const double a = A_a[j];
    const bool strong = s < se && S_j[s] == j1;
    if (strong) s++;
    if (strong && assigned[j1] == p - 1) {
      for (int k = B_i[j1]; k < B_i[j1 + 1]; k++) {
        alfa = a * B_a[k];
        sum_C += alfa;
        sum_N += alfa;
      }
    } else if (cf[j1] != -3) {
      sum_N += a;
    }
  }
  const double diagonal = A_a[A_i[i]];
  if (sum_C * diagonal != 0) alfa = -sum_N / (sum_C * diagonal);   // else: the last product, as in the reference
  for (int k = C_i[i]; k < C_i[i + 1]; k++) C_a[k] *= alfa;
}
struct PassRows {
  const int *i[MAX_PASSES];
  const int *j[MAX_PASSES];
  const double *a[MAX_PASSES];
};
__global__ void merge_count_kernel(int n, const int *__restrict__ assigned, PassRows R, int *__restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int c = 0;
  if (i < n) {
    const int p = assigned[i];
    if (p == 0) c = 1;
    else if (p > 0) c = R.i[p][i + 1] - R.i[p][i];
  }
  cnt[i] = c;
}
__global__ void merge_fill_kernel(int n, const int *__restrict__ assigned, const int *__restrict__ f2c, PassRows R,
                                  const int *__restrict__ P_i, int *__restrict__ P_j, double *__restrict__ P_a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = assigned[i], d = P_i[i];
  if (p == 0) { P_j[d] = f2c[i]; P_a[d] = 1.0; }
  else if (p > 0) {
    const int b = R.i[p][i], len = R.i[p][i + 1] - b;
    for (int k = 0; k < len; k++) { P_j[d + k] = R.j[p][b + k]; P_a[d + k] = R.a[p][b + k]; }
  }
}
__global__ void sf_to_f_kernel(int n, int *cf) {                      // par_multi_interp.c:2030-2036
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cf[i] == -3) cf[i] = -1;
}

// allocate a CSR with row pointer = exclusive scan of cnt[0..n]
int csr_from_counts(b200_handle h, int n, int ncols, int *cnt, bool with_data, b200_csr *out) {
  B200_TRY(b200_exclusive_scan_inplace(h, cnt, (size_t)n + 1));
  int nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, cnt + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr M = nullptr;
  B200_TRY(b200_csr_alloc(h, n, ncols, nnz, with_data, &M));
  B200_CUDA(cudaMemcpyAsync(M->i, cnt, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
  *out = M;
  return 0;
}

}  // namespace

// hypre_BoomerAMGCreate2ndS, num_paths 1.  General form shared with the multi-rank path:
//   S      rows [owned | ghost nodes ...] with localized columns (rows beyond the first ring may be empty),
//   cf/f2c indexed like the columns of S; f2c = coarse id of a C point in the numbering of the output columns
//          (global ids across ranks), first_coarse = id of this rank's first C point, ncoarse = number of columns.
// Output: one row per OWNED C point, column ids from f2c, pattern only, no diagonal.
int b200_create_2nd_s_ex(b200_handle h, b200_csr S, int n_owned, const int *d_cf, const int *d_f2c, int first_coarse,
                         int ncoarse, b200_csr *out) {
  const int next = S->nrows;
  int nc = 0, *c2f = nullptr, *cnt = nullptr, *lf2c = nullptr;
  // local numbering of the owned C points (rank among the owned rows)
  B200_TRY(b200_coarse_map(h, n_owned, d_cf, &lf2c, &nc));
  B200_TRY(b200_dalloc<int>(h, &c2f, (size_t)nc + 1));
  if (n_owned) { c2f_kernel<<<b200_grid(n_owned, TBA), TBA, 0, h->stream>>>(n_owned, d_cf, lf2c, c2f); B200_LAUNCH_CHECK(); }
  b200_csr Sc = nullptr, M = nullptr, C = nullptr, S2 = nullptr;
  B200_TRY(b200_dalloc<int>(h, &cnt, (size_t)std::max(next, nc) + 1));
  sc_count_kernel<<<b200_grid((size_t)nc + 1, TBA), TBA, 0, h->stream>>>(nc, c2f, S->i, cnt);
  B200_LAUNCH_CHECK();
  B200_TRY(csr_from_counts(h, nc, next, cnt, true, &Sc));
  if (nc) { sc_fill_kernel<<<b200_grid(nc, TBA), TBA, 0, h->stream>>>(nc, c2f, S->i, S->j, Sc->i, Sc->j, Sc->a); B200_LAUNCH_CHECK(); }
  m_count_kernel<<<b200_grid((size_t)next + 1, TBA), TBA, 0, h->stream>>>(next, d_cf, S->i, S->j, cnt);
  B200_LAUNCH_CHECK();
  B200_TRY(csr_from_counts(h, next, ncoarse, cnt, true, &M));
  if (next) { m_fill_kernel<<<b200_grid(next, TBA), TBA, 0, h->stream>>>(next, d_cf, d_f2c, S->i, S->j, M->i, M->j, M->a); B200_LAUNCH_CHECK(); }
  B200_TRY(b200_csr_multiply_ex(h, Sc, M, 0, 0, ncoarse, &C));
  offdiag_count_kernel<<<b200_grid((size_t)nc + 1, TBA), TBA, 0, h->stream>>>(nc, first_coarse, C->i, C->j, cnt);
  B200_LAUNCH_CHECK();
  B200_TRY(csr_from_counts(h, nc, ncoarse, cnt, false, &S2));
  if (nc) { offdiag_fill_kernel<<<b200_grid(nc, TBA), TBA, 0, h->stream>>>(nc, first_coarse, C->i, C->j, S2->i, S2->j); B200_LAUNCH_CHECK(); }
  B200_TRY(b200_csr_destroy(h, Sc)); B200_TRY(b200_csr_destroy(h, M)); B200_TRY(b200_csr_destroy(h, C));
  B200_TRY(b200_dfree(h, lf2c)); B200_TRY(b200_dfree(h, c2f)); B200_TRY(b200_dfree(h, cnt));
  *out = S2;
  return 0;
}
extern "C" int b200_create_2nd_s(b200_handle h, b200_csr S, const int *d_cf, b200_csr *out) {
  if (!S) B200_FAIL("create2ndS: null S");
  int *f2c = nullptr, nc = 0;
  B200_TRY(b200_coarse_map(h, S->nrows, d_cf, &f2c, &nc));
  B200_TRY(b200_create_2nd_s_ex(h, S, S->nrows, d_cf, f2c, 0, nc, out));
  B200_TRY(b200_dfree(h, f2c));
  return 0;
}

// hypre_BoomerAMGCorrectCFMarker on the owned rows: cfn is indexed by the local rank of the C point
int b200_correct_cf(b200_handle h, int n_owned, const int *d_cfn, int *d_cf) {
  int *f2c = nullptr, nc = 0;
  B200_TRY(b200_coarse_map(h, n_owned, d_cf, &f2c, &nc));
  if (n_owned) { correct_cf_kernel<<<b200_grid(n_owned, TBA), TBA, 0, h->stream>>>(n_owned, f2c, d_cfn, d_cf); B200_LAUNCH_CHECK(); }
  B200_TRY(b200_dfree(h, f2c));
  return 0;
}

// second stage of aggressive coarsening (par_amg_setup.c:1239-1256 + :1592): PMIS on S2 with CF_init 3, then
// hypre_BoomerAMGCorrectCFMarker.  d_cf: in = first PMIS marker, out = corrected marker
extern "C" int b200_agg_coarsen(b200_handle h, b200_csr S, int seed, int *d_cf) {
  b200_csr S2 = nullptr;
  B200_TRY(b200_create_2nd_s(h, S, d_cf, &S2));
  int *cfn = nullptr;
  B200_TRY(b200_dalloc<int>(h, &cfn, (size_t)S2->nrows + 1));
  B200_TRY(b200_pmis_rows_init(h, S2, seed, 0, 3, cfn, nullptr));
  B200_TRY(b200_correct_cf(h, S->nrows, cfn, d_cf));
  B200_TRY(b200_csr_destroy(h, S2));
  B200_TRY(b200_dfree(h, cfn));
  return 0;
}

// hypre_BoomerAMGBuildMultipass, weight_option 0, trunc_factor 0, P_max_elmts 0.  General form shared with the
// multi-rank path: A and S hold the n owned rows with localized columns [owned | ghost]; d_cf, d_f2c (coarse ids in
// the numbering of P's columns) and the pass numbers are indexed like those columns; the hooks refresh ghost tails,
// sum over ranks and append the ghost nodes' rows of a pass (all no-ops on one rank).
int b200_multipass_ex(b200_handle h, b200_csr A, b200_csr S, int n, int n_ext, int *d_cf, const int *d_f2c, int ncoarse,
                      const b200_agg_hooks *hooks, b200_csr *out) {
  int *assigned = nullptr, *assigned2 = nullptr, *d_rem = nullptr, *cnt = nullptr;
  B200_TRY(b200_dalloc<int>(h, &assigned, (size_t)n_ext + 1));
  B200_TRY(b200_dalloc<int>(h, &assigned2, (size_t)n_ext + 1));
  B200_TRY(b200_dalloc<int>(h, &d_rem, 1));
  B200_TRY(b200_dalloc<int>(h, &cnt, (size_t)n + 1));
  if (n_ext) { assign_init_kernel<<<b200_grid(n_ext, TBA), TBA, 0, h->stream>>>(n_ext, d_cf, assigned); B200_LAUNCH_CHECK(); }
  int npass = 1, remaining = 1;
  for (int p = 1; p < MAX_PASSES && remaining; p++) {                 // pass 1, then `while (remaining && pass < 10)`
    B200_CUDA(cudaMemsetAsync(d_rem, 0, sizeof(int), h->stream));
    B200_CUDA(cudaMemcpyAsync(assigned2, assigned, sizeof(int) * (size_t)n_ext, cudaMemcpyDeviceToDevice, h->stream));
    if (n) {
      assign_pass_kernel<<<b200_grid(n, TBA), TBA, 0, h->stream>>>(n, p, d_cf, S->i, S->j, assigned, assigned2, d_rem);
      B200_LAUNCH_CHECK();
    }
    std::swap(assigned, assigned2);
    if (hooks && hooks->sync_int) B200_TRY(hooks->sync_int(assigned));
    B200_CUDA(cudaMemcpyAsync(&remaining, d_rem, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    if (hooks && hooks->sum_int) B200_TRY(hooks->sum_int(&remaining));
    npass = p + 1;
  }
  b200_csr rows[MAX_PASSES] = {nullptr};
  for (int p = 1; p < npass; p++) {
    nbr_count_kernel<<<b200_grid((size_t)n + 1, TBA), TBA, 0, h->stream>>>(n, p, S->i, S->j, assigned, cnt);
    B200_LAUNCH_CHECK();
    if (p == 1) {
      B200_TRY(csr_from_counts(h, n, ncoarse, cnt, true, &rows[1]));
      if (n) {
        pass1_fill_kernel<<<b200_grid(n, TBA), TBA, 0, h->stream>>>(n, A->i, A->j, A->a, S->i, S->j, d_cf, d_f2c, assigned,
                                                                    rows[1]->i, rows[1]->j, rows[1]->a);
        B200_LAUNCH_CHECK();
      }
    } else {
      b200_csr Ap = nullptr, B = rows[p - 1];
      if (hooks && hooks->with_ghost_rows) B200_TRY(hooks->with_ghost_rows(rows[p - 1], &B));   // [owned rows ; ghost nodes' rows]
      B200_TRY(csr_from_counts(h, n, n_ext, cnt, true, &Ap));
      if (n) {
        ap_fill_kernel<<<b200_grid(n, TBA), TBA, 0, h->stream>>>(n, p, A->i, A->j, A->a, S->i, S->j, assigned, Ap->i, Ap->j, Ap->a);
        B200_LAUNCH_CHECK();
      }
      B200_TRY(b200_csr_multiply_ex(h, Ap, B, 0, 0, ncoarse, &rows[p]));
      B200_TRY(b200_csr_destroy(h, Ap));
      if (n) {
        passp_scale_kernel<<<b200_grid(n, TBA), TBA, 0, h->stream>>>(n, p, A->i, A->j, A->a, S->i, S->j, d_cf, assigned, B->i, B->a,
                                                                     rows[p]->i, rows[p]->a);
        B200_LAUNCH_CHECK();
      }
      if (B != rows[p - 1]) B200_TRY(b200_csr_destroy(h, B));
    }
  }
  PassRows R;
  for (int p = 0; p < MAX_PASSES; p++) {
    R.i[p] = rows[p] ? rows[p]->i : nullptr; R.j[p] = rows[p] ? rows[p]->j : nullptr; R.a[p] = rows[p] ? rows[p]->a : nullptr;
  }
  merge_count_kernel<<<b200_grid((size_t)n + 1, TBA), TBA, 0, h->stream>>>(n, assigned, R, cnt);
  B200_LAUNCH_CHECK();
  b200_csr P = nullptr;
  B200_TRY(csr_from_counts(h, n, ncoarse, cnt, true, &P));
  if (n) {
    merge_fill_kernel<<<b200_grid(n, TBA), TBA, 0, h->stream>>>(n, assigned, d_f2c, R, P->i, P->j, P->a);
    B200_LAUNCH_CHECK();
    sf_to_f_kernel<<<b200_grid(n, TBA), TBA, 0, h->stream>>>(n, d_cf);
    B200_LAUNCH_CHECK();
  }
  for (int p = 1; p < MAX_PASSES; p++) if (rows[p]) B200_TRY(b200_csr_destroy(h, rows[p]));
  B200_TRY(b200_dfree(h, assigned)); B200_TRY(b200_dfree(h, assigned2));
  B200_TRY(b200_dfree(h, d_rem)); B200_TRY(b200_dfree(h, cnt));
  *out = P;
  return 0;
}

// one rank.  d_cf: {1, -1, -3}; SF points (-3) are folded into F on return, as the reference does.
extern "C" int b200_multipass_interp(b200_handle h, b200_csr A, b200_csr S, int *d_cf, b200_csr *out) {
  if (!A || !A->a || !S) B200_FAIL("multipass: bad arguments");
  const int n = A->nrows;
  int *f2c = nullptr, nc = 0;
  B200_TRY(b200_coarse_map(h, n, d_cf, &f2c, &nc));
  B200_TRY(b200_multipass_ex(h, A, S, n, n, d_cf, f2c, nc, nullptr, out));
  B200_TRY(b200_dfree(h, f2c));
  return 0;
}
