// b200_ij.cu -- HYPRE_IJMatrix assembly on the device.
//
// Reference: hypre_IJMatrixSetValuesParCSR / AddToValuesParCSR (IJ_mv/IJMatrix_parcsr.c:697-1186, :1188-1700: rows kept
// in an auxiliary matrix in insertion order) and hypre_IJMatrixAssembleParCSR (:2774-3080: rows copied into the CSR
// with the LAST entry on the diagonal column moved to the front).  Entry order decides the bits of every later setup
// stage, so the device path reproduces that order instead of the column-sorted order of the reference's own device
// assembly (IJ_mv/IJMatrix_parcsr_device.c).
//
// Host side: SetValues / AddToValues only validate and append (row, column, value, block) records to a pinned chunk;
// full chunks stream to a device log while the caller keeps filling the other chunk.  A "block" is one (call, row)
// pair: the reference searches a new entry only among the entries the row had BEFORE that pair started
// (`old_size`, :941-962), so duplicates inside one block stay duplicates and later blocks hit the first match.
// Assemble: stable radix sort of the log by row (insertion order survives), one gather, then one thread per row
// replays its blocks in place (set / add / append), counts, scan, and a copy with the diagonal first.
// After the first assembly SetValues / AddToValues may only touch existing entries (:727-905); those records are
// replayed onto the CSR values by the next Assemble and a missing entry raises the reference's error.
#include <cub/cub.cuh>
#include "b200_internal.h"

struct b200_ij_s {
  int ilower = 0, iupper = -1, jlower = 0, jupper = -1;
  int diag_shift = 0;          // local row r holds its diagonal in (global) column r + diag_shift
  static constexpr size_t CHUNK = (size_t)1 << 21;      // records per pinned chunk (40 MB)
  int *h_row[2] = {nullptr, nullptr}, *h_col[2] = {nullptr, nullptr}, *h_blk[2] = {nullptr, nullptr};
  double *h_val[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool in_flight[2] = {false, false};
  int cur = 0;
  size_t fill = 0;
  int *d_row = nullptr, *d_col = nullptr, *d_blk = nullptr;   // device log
  double *d_val = nullptr;
  size_t n_log = 0, cap = 0;
  unsigned block_counter = 0;
  b200_parcsr A = nullptr;     // the assembled object (owned by the caller once returned)
  long long n_errors = 0;      // records dropped at SetValues time (row / column outside the declared ranges)
};

namespace {
__global__ void iota_kernel(size_t n, int *idx) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (int)i;
}
// ptr[r] = first position of the sorted keys with key >= r, r = 0..nrows
__global__ void row_bounds_kernel(int nrows, size_t n, const int *__restrict__ keys, int *__restrict__ ptr) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nrows) return;
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) >> 1;
    if (keys[mid] < r) lo = mid + 1; else hi = mid;
  }
  ptr[r] = (int)lo;
}
__global__ void gather_kernel(size_t n, const int *__restrict__ perm, const int *__restrict__ col, const double *__restrict__ val,
                              const int *__restrict__ blk, int *__restrict__ col_s, double *__restrict__ val_s, int *__restrict__ blk_s) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = perm[i];
  col_s[i] = col[p]; val_s[i] = val[p]; blk_s[i] = blk[p];
}
// replay of the auxiliary-matrix insertion (IJMatrix_parcsr.c:930-1000) for one row, in place on its segment
__global__ void merge_rows_kernel(int nrows, int diag_shift, const int *__restrict__ ptr, int *col, double *val, const int *__restrict__ blk,
                                  int *__restrict__ cnt, int *__restrict__ dpos) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int s = ptr[r], e = ptr[r + 1];
  int m = 0, old_size = 0, cur_blk = -1, dp = -1;
  for (int t = s; t < e; t++) {
    const int b = blk[t];
    if ((b >> 1) != cur_blk) { cur_blk = b >> 1; old_size = m; }
    const int c = col[t];
    const double v = val[t];
    bool found = false;
    for (int q = 0; q < old_size; q++)
      if (col[s + q] == c) {
        val[s + q] = (b & 1) ? val[s + q] + v : v;
        found = true;
        break;
      }
    if (!found) { col[s + m] = c; val[s + m] = v; m++; }
  }
  for (int q = 0; q < m; q++)
    if (col[s + q] - diag_shift == r) dp = q;               // the LAST entry on the diagonal column (:2974-2977)
  cnt[r] = m;
  dpos[r] = dp;
}
__global__ void fill_rows_kernel(int nrows, int jlower, const int *__restrict__ ptr, const int *__restrict__ col,
                                 const double *__restrict__ val, const int *__restrict__ dpos, const int *__restrict__ A_i,
                                 int *__restrict__ A_j, double *__restrict__ A_a) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int s = ptr[r], m = A_i[r + 1] - A_i[r], dp = dpos[r];
  int o = A_i[r];
  if (dp > -1) { A_j[o] = col[s + dp] - jlower; A_a[o] = val[s + dp]; o++; }      // :3030-3034
  for (int q = 0; q < m; q++)
    if (q != dp) { A_j[o] = col[s + q] - jlower; A_a[o] = val[s + q]; o++; }
}
// after the first assembly: replay set / add records onto existing entries (IJMatrix_parcsr.c:727-905)
__global__ void update_rows_kernel(int nrows, int jlower, const int *__restrict__ ptr, const int *__restrict__ col,
                                   const double *__restrict__ val, const int *__restrict__ blk, const int *__restrict__ A_i,
                                   const int *__restrict__ A_j, double *__restrict__ A_a, int *__restrict__ n_missing) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int a0 = A_i[r], a1 = A_i[r + 1];
  for (int t = ptr[r]; t < ptr[r + 1]; t++) {
    const int c = col[t] - jlower;
    bool found = false;
    for (int q = a0; q < a1; q++)
      if (A_j[q] == c) {
        A_a[q] = (blk[t] & 1) ? A_a[q] + val[t] : val[t];
        found = true;
        break;
      }
    if (!found) atomicAdd(n_missing, 1);
  }
}
int release_log(b200_handle h, b200_ij ij) {
  b200_dfree(h, ij->d_row); b200_dfree(h, ij->d_col); b200_dfree(h, ij->d_blk); b200_dfree(h, ij->d_val);
  ij->d_row = ij->d_col = ij->d_blk = nullptr;
  ij->d_val = nullptr;
  ij->n_log = ij->cap = 0;
  return 0;
}
// stream the current pinned chunk to the device log and switch to the other chunk
int flush_chunk(b200_handle h, b200_ij ij) {
  const size_t k = ij->fill;
  if (k == 0) return 0;
  if (ij->n_log + k > ij->cap) {
    size_t ncap = ij->cap ? ij->cap * 2 : b200_ij_s::CHUNK * 2;
    while (ncap < ij->n_log + k) ncap *= 2;
    if (ncap > 2147483647ull) ncap = 2147483647ull;
    if (ij->n_log + k > ncap) B200_FAIL("ij: more than 2^31 - 1 records between two assemblies");
    int *nr = nullptr, *nc = nullptr, *nb = nullptr;
    double *nv = nullptr;
    B200_TRY(b200_dalloc<int>(h, &nr, ncap)); B200_TRY(b200_dalloc<int>(h, &nc, ncap));
    B200_TRY(b200_dalloc<int>(h, &nb, ncap)); B200_TRY(b200_dalloc<double>(h, &nv, ncap));
    if (ij->n_log) {
      B200_CUDA(cudaMemcpyAsync(nr, ij->d_row, sizeof(int) * ij->n_log, cudaMemcpyDeviceToDevice, h->stream));
      B200_CUDA(cudaMemcpyAsync(nc, ij->d_col, sizeof(int) * ij->n_log, cudaMemcpyDeviceToDevice, h->stream));
      B200_CUDA(cudaMemcpyAsync(nb, ij->d_blk, sizeof(int) * ij->n_log, cudaMemcpyDeviceToDevice, h->stream));
      B200_CUDA(cudaMemcpyAsync(nv, ij->d_val, sizeof(double) * ij->n_log, cudaMemcpyDeviceToDevice, h->stream));
    }
    b200_dfree(h, ij->d_row); b200_dfree(h, ij->d_col); b200_dfree(h, ij->d_blk); b200_dfree(h, ij->d_val);
    ij->d_row = nr; ij->d_col = nc; ij->d_blk = nb; ij->d_val = nv;
    ij->cap = ncap;
  }
  const int c = ij->cur;
  B200_CUDA(cudaMemcpyAsync(ij->d_row + ij->n_log, ij->h_row[c], sizeof(int) * k, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(ij->d_col + ij->n_log, ij->h_col[c], sizeof(int) * k, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(ij->d_blk + ij->n_log, ij->h_blk[c], sizeof(int) * k, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(ij->d_val + ij->n_log, ij->h_val[c], sizeof(double) * k, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaEventRecord(ij->ev[c], h->stream));
  ij->in_flight[c] = true;
  ij->n_log += k;
  ij->fill = 0;
  ij->cur = c ^ 1;
  if (ij->in_flight[ij->cur]) {                         // the chunk we are about to refill must have left the host
    B200_CUDA(cudaEventSynchronize(ij->ev[ij->cur]));
    ij->in_flight[ij->cur] = false;
  }
  return 0;
}
}  // namespace

extern "C" int b200_ij_create(b200_handle h, int ilower, int iupper, int jlower, int jupper, b200_ij *out) {
  if (!h || !out) B200_FAIL("ij_create: null argument");
  if (ilower > iupper + 1 || ilower < 0 || jlower > jupper + 1 || jlower < 0) B200_FAIL("ij_create: bad row / column range");
  b200_ij ij = new b200_ij_s();
  ij->ilower = ilower; ij->iupper = iupper; ij->jlower = jlower; ij->jupper = jupper;
  ij->diag_shift = jlower;                             // one rank: local row r <-> local column r (IJMatrix_parcsr.c:2974)
  for (int c = 0; c < 2; c++) {
    B200_CUDA(cudaHostAlloc((void **)&ij->h_row[c], sizeof(int) * b200_ij_s::CHUNK, cudaHostAllocDefault));
    B200_CUDA(cudaHostAlloc((void **)&ij->h_col[c], sizeof(int) * b200_ij_s::CHUNK, cudaHostAllocDefault));
    B200_CUDA(cudaHostAlloc((void **)&ij->h_blk[c], sizeof(int) * b200_ij_s::CHUNK, cudaHostAllocDefault));
    B200_CUDA(cudaHostAlloc((void **)&ij->h_val[c], sizeof(double) * b200_ij_s::CHUNK, cudaHostAllocDefault));
    B200_CUDA(cudaEventCreateWithFlags(&ij->ev[c], cudaEventDisableTiming));
  }
  *out = ij;
  return 0;
}

// the rows [ilower, iupper] of a square operator spread over several ranks: columns stay GLOBAL (0 .. global_cols-1) in the
// assembled CSR, which is the form b200_dist_matrix_create_from_ij localizes (diag / offd split + col_map_offd of
// GenerateDiagAndOffd, par_csr_matrix.c:1634, happen there)
extern "C" int b200_ij_create_rows(b200_handle h, int ilower, int iupper, int global_cols, b200_ij *out) {
  B200_TRY(b200_ij_create(h, ilower, iupper, 0, global_cols - 1, out));
  (*out)->diag_shift = ilower;
  return 0;
}
// hand the assembled object over (b200_dist.cu): the assembler forgets it
b200_parcsr b200_ij_release_object(b200_ij ij) {
  b200_parcsr A = ij->A;
  ij->A = nullptr;
  return A;
}

extern "C" int b200_ij_destroy(b200_handle h, b200_ij ij) {
  if (!ij) return 0;
  cudaStreamSynchronize(h->stream);
  release_log(h, ij);
  for (int c = 0; c < 2; c++) {
    if (ij->h_row[c]) cudaFreeHost(ij->h_row[c]);
    if (ij->h_col[c]) cudaFreeHost(ij->h_col[c]);
    if (ij->h_blk[c]) cudaFreeHost(ij->h_blk[c]);
    if (ij->h_val[c]) cudaFreeHost(ij->h_val[c]);
    if (ij->ev[c]) cudaEventDestroy(ij->ev[c]);
  }
  delete ij;
  return 0;
}

// HYPRE_IJMatrixSetValues (add = 0) / AddToValues (add = 1): host arrays, rows[] and cols[] global indices.
// Records outside the declared ranges are dropped and counted in *n_rejected (the caller raises the error flag).
extern "C" int b200_ij_set_values(b200_handle h, b200_ij ij, int nrows, const int *ncols, const int *rows, const int *cols,
                                  const double *values, int add, int *n_rejected) {
  if (!ij || nrows < 0 || (nrows && (!ncols || !rows || !cols || !values))) B200_FAIL("ij_set_values: bad argument");
  int rejected = 0;
  size_t at = 0;
  for (int r = 0; r < nrows; r++) {
    const int row = rows[r], n = ncols[r];
    if (n < 0) B200_FAIL("ij_set_values: negative ncols");
    if (n == 0) continue;                                                                   // "empty row" (:919-922), before the ownership test
    if (row < ij->ilower || row > ij->iupper) { rejected += n; at += (size_t)n; continue; }   // off-processor row: no owner to send it to
    const int blk = (int)(((ij->block_counter++) & 0x3fffffffu) << 1) | (add ? 1 : 0);
    for (int k = 0; k < n; k++, at++) {
      const int c = cols[at];
      if (c < ij->jlower || c > ij->jupper) { rejected++; continue; }
      if (ij->fill == b200_ij_s::CHUNK) B200_TRY(flush_chunk(h, ij));
      const int cu = ij->cur;
      const size_t f = ij->fill++;
      ij->h_row[cu][f] = row - ij->ilower;
      ij->h_col[cu][f] = c;
      ij->h_blk[cu][f] = blk;
      ij->h_val[cu][f] = values[at];
    }
  }
  ij->n_errors += rejected;
  if (n_rejected) *n_rejected = rejected;
  return 0;
}

// HYPRE_IJMatrixAssemble.  First call: builds the ParCSR object (diag block = the whole matrix, one rank) and returns
// it in *A_out.  Later calls replay the records logged since then onto the existing entries; *n_missing counts the
// records whose element does not exist (the reference's " Error, element %b %b does not exist").
extern "C" int b200_ij_assemble(b200_handle h, b200_ij ij, b200_parcsr *A_out, int *n_missing_out) {
  if (!ij) B200_FAIL("ij_assemble: null argument");
  if (n_missing_out) *n_missing_out = 0;
  B200_TRY(flush_chunk(h, ij));
  const int nrows = ij->iupper - ij->ilower + 1, ncols = ij->jupper - ij->jlower + 1;
  const size_t n = ij->n_log;
  int *idx = nullptr, *perm = nullptr, *keys = nullptr, *ptr = nullptr, *col_s = nullptr, *blk_s = nullptr, *cnt = nullptr, *dpos = nullptr;
  double *val_s = nullptr;
  B200_TRY(b200_dalloc<int>(h, &ptr, (size_t)nrows + 2));
  if (n) {
    B200_TRY(b200_dalloc<int>(h, &idx, n)); B200_TRY(b200_dalloc<int>(h, &perm, n)); B200_TRY(b200_dalloc<int>(h, &keys, n));
    iota_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, idx);
    B200_LAUNCH_CHECK();
    int bits = 1;
    while (bits < 31 && (1ll << bits) < (long long)nrows) bits++;
    size_t tb = 0;
    B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, ij->d_row, keys, idx, perm, (int)n, 0, bits, h->stream));
    char *tmp = nullptr;
    B200_TRY(b200_dalloc<char>(h, &tmp, tb));
    B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, ij->d_row, keys, idx, perm, (int)n, 0, bits, h->stream));   // stable
    ++g_b200_launches;
    B200_TRY(b200_dfree(h, tmp));
    B200_TRY(b200_dfree(h, idx));
    row_bounds_kernel<<<b200_grid((size_t)nrows + 1, 256), 256, 0, h->stream>>>(nrows, n, keys, ptr);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_dfree(h, keys));
    B200_TRY(b200_dalloc<int>(h, &col_s, n)); B200_TRY(b200_dalloc<int>(h, &blk_s, n)); B200_TRY(b200_dalloc<double>(h, &val_s, n));
    gather_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, perm, ij->d_col, ij->d_val, ij->d_blk, col_s, val_s, blk_s);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_dfree(h, perm));
  } else {
    B200_CUDA(cudaMemsetAsync(ptr, 0, sizeof(int) * ((size_t)nrows + 2), h->stream));
  }
  if (!ij->A) {
    B200_TRY(b200_dalloc<int>(h, &cnt, (size_t)nrows + 1));
    B200_TRY(b200_dalloc<int>(h, &dpos, (size_t)nrows + 1));
    B200_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)nrows + 1), h->stream));
    if (nrows) {
      merge_rows_kernel<<<b200_grid(nrows, 128), 128, 0, h->stream>>>(nrows, ij->diag_shift, ptr, col_s, val_s, blk_s, cnt, dpos);
      B200_LAUNCH_CHECK();
    }
    B200_TRY(b200_exclusive_scan_inplace(h, cnt, (size_t)nrows + 1));
    int nnz = 0;
    B200_CUDA(cudaMemcpyAsync(&nnz, cnt + nrows, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    b200_parcsr A = new b200_parcsr_s();
    A->global_rows = nrows; A->global_cols = ncols;
    A->first_row = ij->ilower; A->first_col = ij->jlower;
    B200_TRY(b200_csr_alloc(h, nrows, ncols, nnz, true, &A->diag));
    B200_CUDA(cudaMemcpyAsync(A->diag->i, cnt, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, h->stream));
    if (nrows && nnz) {
      fill_rows_kernel<<<b200_grid(nrows, 128), 128, 0, h->stream>>>(nrows, ij->jlower, ptr, col_s, val_s, dpos, A->diag->i, A->diag->j, A->diag->a);
      B200_LAUNCH_CHECK();
    }
    B200_TRY(b200_csr_alloc(h, nrows, 0, 0, true, &A->offd));
    B200_CUDA(cudaMemsetAsync(A->offd->i, 0, sizeof(int) * ((size_t)nrows + 1), h->stream));
    B200_TRY(b200_csr_build_plan(h, A->diag));
    B200_TRY(b200_dfree(h, cnt)); B200_TRY(b200_dfree(h, dpos));
    ij->A = A;
  } else if (n) {
    int *miss = nullptr;
    B200_TRY(b200_dalloc<int>(h, &miss, 1));
    B200_CUDA(cudaMemsetAsync(miss, 0, sizeof(int), h->stream));
    b200_csr D = ij->A->diag;
    update_rows_kernel<<<b200_grid(nrows, 128), 128, 0, h->stream>>>(nrows, ij->jlower, ptr, col_s, val_s, blk_s, D->i, D->j, D->a, miss);
    B200_LAUNCH_CHECK();
    int hm = 0;
    B200_CUDA(cudaMemcpyAsync(&hm, miss, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dfree(h, miss));
    if (n_missing_out) *n_missing_out = hm;
    if (D->T) { B200_TRY(b200_csr_destroy(h, D->T)); D->T = nullptr; }      // cached transpose holds the old values
    B200_TRY(b200_csr_drop_dict(h, D));                                     // and so does the dictionary-compressed solve copy
    B200_TRY(b200_csr_build_dict(h, D));
  }
  b200_dfree(h, col_s); b200_dfree(h, blk_s); b200_dfree(h, val_s); b200_dfree(h, ptr);
  release_log(h, ij);
  if (A_out) *A_out = ij->A;
  return 0;
}

extern "C" long long b200_ij_num_rejected(b200_ij ij) { return ij ? ij->n_errors : 0; }
