// b200_dist.cu -- row-partitioned (multi-GPU) ParCSR: halo plans, ghost-row fetches, distributed
// setup (PMIS / ext+i / Galerkin product), V-cycle and PCG.
//
// Reference: hypre_ParCSRMatrix + hypre_ParCSRCommPkg (parcsr_mv/par_csr_matrix.h:27-95,
// par_csr_communication.h:54-82), hypre_ParCSRMatrixMatvecOutOfPlace (par_csr_matvec.c:22-359),
// hypre_ParCSRMatrixExtractBExt (par_csr_matop.c:1655), hypre_exchange_interp_data
// (parcsr_ls/aux_interp.c:552-660), hypre_ParCSRMatrixRAPKTHost multi-rank branch
// (par_csr_triplemat.c:606-871), hypre_BoomerAMGCoarseParms (par_coarse_parms.c:56-133).
//
// Design (B200-first, see include/hypre_b200.h "multi-GPU"): every rank keeps its rows as ONE CSR.
// During setup the column ids are GLOBAL; ghost rows needed by a per-row algorithm are fetched from
// their owners with their entry order intact, so the single-GPU row kernels run unchanged and the
// hierarchy is bit-identical for any number of GPUs.  For the solve phase columns are localized to
// [owned | ghosts sorted by global id] and a SpMV is: pack -> grouped ncclSend/ncclRecv straight into
// the ghost tail of x -> one streaming kernel.
#include "b200_internal.h"
#include <chrono>
#include "b200_comm.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <algorithm>
#include <cmath>
#include <map>

int b200_csr_spmv_epi(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha,
                      double beta, const double *b, const double *d);
int b200_pmis_dist(b200_handle h, b200_comm c, b200_csr S, b200_halo_s *halo, int seed, long long first_row, int *d_cf_ext);
int b200_extpi_interp_ex(b200_handle h, b200_csr A, b200_csr S, const int *d_cf, int n, const int *d_f2c_in, int ncoarse_in,
                         double trunc_factor, int max_elmts, b200_csr *out);
int b200_csr_multiply_ex(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C, b200_csr *out);
int b200_vec_dot_dev(b200_handle h, int n, const double *x, const double *y, double *d_out);
int b200_generate_stencil_global(b200_handle h, int nx, int ny, int nz, int P, int Q, int R, int p, int q, int r,
                                 int stencil, const double *vals, b200_csr *out, int *first_row);
int b200_box_first_row(int nx, int ny, int nz, int P, int Q, int R, int p, int q, int r);

struct b200_amg_s;   // parameter maps live in b200_amg.cu
int b200_amg_get_int(b200_amg a, const char *name);
int b200_pmis_dist_init(b200_handle h, b200_comm c, b200_csr S, b200_halo_s *halo, int seed, long long first_row, int cf_init,
                        int *d_cf_ext);
double b200_amg_get_real(b200_amg a, const char *name);
int b200_amg_clone_params(b200_amg src, b200_amg dst);
int b200_amg_precond(b200_handle h, b200_amg amg, const double *d_rhs, double *d_out);

// legacy diag/offd ParCSR hooks (b200_parcsr.cu): the multi-rank path is the b200_dist_* API
int b200_halo_exchange(b200_handle, b200_parcsr, const double *) { B200_FAIL("use the b200_dist_* API for multi-rank operators"); }
void b200_halo_destroy(b200_handle, b200_halo_s *) {}

struct b200_dist_matrix_s {
  int n = 0;                    // local rows
  int first_row = 0, global_rows = 0;
  int first_col = 0, n_owned_cols = 0, global_cols = 0;
  std::vector<int> row_starts, col_starts;   // ownership of rows / columns, size nranks+1
  b200_csr G = nullptr;         // local rows, GLOBAL column ids (setup form)
  b200_csr L = nullptr;         // local rows, localized columns [owned | ghosts] (solve form)
  b200_csr Ls = nullptr;        // solve-phase copy of L with every row sorted by column (coarse Galerkin operators: the x gather
                                // of neighbouring lanes then hits neighbouring lines).  L keeps the first-touch order
                                // the next level's setup depends on.
  b200_halo_s *halo = nullptr;  // ghosts of L
};

namespace {

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void pack_kernel(int n, const int *__restrict__ idx, const T *__restrict__ src, T *__restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[idx[k]];
}
__global__ void unpack_add_kernel(int n, const int *__restrict__ idx, const int *__restrict__ buf, int *__restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && buf[k]) atomicAdd(&dst[idx[k]], buf[k]);
}
__global__ void unpack_clear_kernel(int n, const int *__restrict__ idx, const int *__restrict__ buf, int *__restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && buf[k] == 0 && dst[idx[k]] > 0) dst[idx[k]] = 0;     // par_coarsen.c:2516-2519
}
__global__ void sub_kernel(int n, int *x, int v) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) x[k] -= v;
}
__global__ void rowlen_kernel(int n, const int *__restrict__ idx, const int *__restrict__ A_i, int *__restrict__ len) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  if (k == n) { len[n] = 0; return; }
  len[k] = A_i[idx[k] + 1] - A_i[idx[k]];
}
__global__ void pack_rows_kernel(int n, const int *__restrict__ idx, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                 const double *__restrict__ A_a, const int *__restrict__ off, int *__restrict__ bj,
                                 double *__restrict__ ba) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int r = idx[k], s = A_i[r], len = A_i[r + 1] - s, d = off[k];
  for (int t = 0; t < len; t++) { bj[d + t] = A_j[s + t]; if (ba) ba[d + t] = A_a[s + t]; }
}
// ghost candidates: columns outside the owned range
// global -> extended local: owned -> c - first ; ghost -> n_owned + rank in the sorted ghost list
__global__ void localize_kernel(int nnz, const int *__restrict__ jg, int first, int n_owned, int ng,
                                const int *__restrict__ ghost, int *__restrict__ jl) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int c = jg[k];
  if (c >= first && c < first + n_owned) { jl[k] = c - first; return; }
  int lo = 0, hi = ng - 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (ghost[mid] < c) lo = mid + 1; else hi = mid; }
  jl[k] = n_owned + lo;
}
// extended local -> global
__global__ void globalize_kernel(int nnz, const int *__restrict__ jl, int first, int n_owned, const int *__restrict__ ghost,
                                 int *__restrict__ jg) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int c = jl[k];
  jg[k] = c < n_owned ? first + c : ghost[c - n_owned];
}
__global__ void owner_kernel(int nnz, const int *__restrict__ jg, int nranks, const int *__restrict__ starts, int *__restrict__ owner) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int c = jg[k];
  int lo = 0, hi = nranks - 1;                        // last r with starts[r] <= c
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (starts[mid] <= c) lo = mid; else hi = mid - 1; }
  owner[k] = lo;
}
__global__ void expand_rows_g_kernel(int nrows, const int *__restrict__ A_i, int first_row, int *__restrict__ rows) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  for (int k = A_i[r]; k < A_i[r + 1]; k++) rows[k] = first_row + r;
}
__global__ void iota_kernel2(int n, int *x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = i;
}
template <class T>
__global__ void gather_kernel(int n, const int *__restrict__ perm, const T *__restrict__ src, T *__restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[perm[k]];
}
__global__ void hist_kernel(int n, const int *__restrict__ key, int *__restrict__ cnt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) atomicAdd(&cnt[key[k]], 1);
}
// histogram over a handful of bins (the owner rank of every entry): equal keys of a warp are counted with one
// shared-memory atomic, a CTA adds its bins to the result once -- millions of entries on two or three counters
// would serialise in L2 otherwise
__global__ void hist_few_bins_kernel(int n, const int *__restrict__ key, int nbins, int *__restrict__ cnt) {
  extern __shared__ int bins[];
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) bins[b] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
    const int k = base + lane;
    const int v = k < n ? key[k] : -1;
    const unsigned same = __match_any_sync(0xffffffffu, v);
    if (v >= 0 && lane == __ffs(same) - 1) atomicAdd(&bins[v], __popc(same));
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x)
    if (bins[b]) atomicAdd(&cnt[b], bins[b]);
}
__global__ void cflag2_kernel(int n, const int *__restrict__ cf, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = cf[i] >= 0 ? 1 : 0;
  if (i == n) flag[n] = 0;
}
__global__ void f2c_global_kernel(int n, const int *__restrict__ cf, int coarse_first, int *__restrict__ f2c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f2c[i] = cf[i] >= 0 ? coarse_first + f2c[i] : -1;
}
__global__ void fix_cf2_kernel(int n, int *cf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cf[i] == -3) cf[i] = -1;
}
__global__ void concat_rowptr_kernel(int n1, const int *__restrict__ i1, int n2, const int *__restrict__ i2, int *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= n1) out[k] = i1[k];
  if (k >= 1 && k <= n2) out[n1 + k] = i1[n1] + i2[k];
}
__global__ void jacobi_zero_kernel2(size_t n, double w, const double *__restrict__ f, const double *__restrict__ l1, double *__restrict__ u) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) u[i] = 0.0 + w * f[i] / l1[i];
}
__global__ void dense_rows_kernel(int n, int ncols, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                  const double *__restrict__ A_a, double *__restrict__ M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < ncols; k++) M[(size_t)i * ncols + k] = 0.0;
  for (int jj = A_i[i]; jj < A_i[i + 1]; jj++) M[(size_t)i * ncols + A_j[jj]] = A_a[jj];
}
__global__ void gselim_kernel2(int n, const double *__restrict__ A_mat, double *__restrict__ A, const double *__restrict__ f,
                               double *__restrict__ x) {   // sstruct_ls/gselim.h on a scratch copy
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < n * n; i++) A[i] = A_mat[i];
  for (int i = 0; i < n; i++) x[i] = f[i];
  if (n == 1) { if (A[0] != 0.0) x[0] = x[0] / A[0]; return; }
  for (int k = 0; k < n - 1; k++) {
    if (A[k * n + k] != 0.0) {
      double divA = 1.0 / A[k * n + k];
      for (int j = k + 1; j < n; j++) {
        if (A[j * n + k] != 0.0) {
          double factor = A[j * n + k] * divA;
          for (int m = k + 1; m < n; m++) A[j * n + m] -= factor * A[k * n + m];
          x[j] -= factor * x[k];
        }
      }
    }
  }
  for (int k = n - 1; k > 0; --k) {
    if (A[k * n + k] != 0.0) {
      x[k] /= A[k * n + k];
      for (int j = 0; j < k; j++) if (A[j * n + k] != 0.0) x[j] -= x[k] * A[j * n + k];
    }
  }
  if (A[0] != 0.0) x[0] /= A[0];
}
inline int vgrid(b200_handle h, size_t n) {
  size_t g = (n + 255) / 256, cap = (size_t)h->num_sm * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// B200_TRACE2=1: stream-synchronised wall-clock time between consecutive calls on this thread, on stderr (setup diagnosis)
void tr(b200_handle h, const char *tag) {
  static const bool on = getenv("B200_TRACE2") != nullptr;
  if (!on) return;
  static thread_local double last = 0;
  cudaStreamSynchronize(h->stream);
  const double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  if (tag) fprintf(stderr, "[b200 tr2] dev %d %-28s %8.3f ms\n", h->device, tag, t - last);
  last = t;
}

int sort_unique(b200_handle h, int *d_keys, int n, int **out, int *n_out) {
  *out = nullptr; *n_out = 0;
  if (n == 0) return 0;
  int *sorted = nullptr, *uniq = nullptr, *d_num = nullptr;
  B200_TRY(b200_dalloc<int>(h, &sorted, n));
  B200_TRY(b200_dalloc<int>(h, &uniq, n));
  B200_TRY(b200_dalloc<int>(h, &d_num, 1));
  size_t tb = 0, tb2 = 0;
  B200_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, d_keys, sorted, n, 0, 32, h->stream));
  B200_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, sorted, uniq, d_num, n, h->stream));
  char *tmp = nullptr;
  B200_TRY(b200_dalloc<char>(h, &tmp, tb > tb2 ? tb : tb2));
  B200_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, d_keys, sorted, n, 0, 32, h->stream));
  B200_CUDA(cub::DeviceSelect::Unique(tmp, tb2, sorted, uniq, d_num, n, h->stream));
  g_b200_launches += 2;
  int num = 0;
  B200_CUDA(cudaMemcpyAsync(&num, d_num, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  int *res = nullptr;
  B200_TRY(b200_dalloc<int>(h, &res, num));
  B200_CUDA(cudaMemcpyAsync(res, uniq, sizeof(int) * (size_t)num, cudaMemcpyDeviceToDevice, h->stream));
  B200_TRY(b200_dfree(h, sorted)); B200_TRY(b200_dfree(h, uniq)); B200_TRY(b200_dfree(h, d_num)); B200_TRY(b200_dfree(h, tmp));
  *out = res; *n_out = num;
  return 0;
}

// ids outside [first, first + n_owned): counted, then appended (any order: the list is sorted afterwards), one atomic per warp.
// Two read-only passes over the column array instead of flag + scan + scatter over nnz-sized scratch.
__global__ void ghost_count_kernel(int nnz, const int *__restrict__ j, int first, int n_owned, int *__restrict__ count) {
  const int stride = gridDim.x * blockDim.x;
  int mine = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const int c = j[k];
    mine += (c < first || c >= first + n_owned) ? 1 : 0;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, mine);
}
__global__ void ghost_append_kernel(int nnz, const int *__restrict__ j, int first, int n_owned, int *__restrict__ cursor,
                                    int *__restrict__ out) {
  const int stride = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int trips = (nnz + stride - 1) / stride;
  for (int t = 0; t < trips; t++) {
    const int k = t * stride + blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    bool g = false;
    if (k < nnz) { c = j[k]; g = (c < first || c >= first + n_owned); }
    const unsigned m = __ballot_sync(0xffffffffu, g);
    if (m) {
      int base = 0;
      if (lane == __ffs(m) - 1) base = atomicAdd(cursor, __popc(m));
      base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
      if (g) out[base + __popc(m & ((1u << lane) - 1u))] = c;
    }
  }
}
// sorted unique list of the column ids of G that fall outside [first, first + n_owned)
int ghost_columns(b200_handle h, const int *d_j, int nnz, int first, int n_owned, int **ghost, int *ng) {
  *ghost = nullptr; *ng = 0;
  if (nnz == 0) return 0;
  int *cnt = nullptr;
  B200_TRY(b200_dalloc<int>(h, &cnt, 2));
  B200_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * 2, h->stream));
  const int want = b200_grid(nnz, 256), cap = h->num_sm * 8;
  const int grid = want < cap ? want : cap;
  ghost_count_kernel<<<grid, 256, 0, h->stream>>>(nnz, d_j, first, n_owned, cnt);
  B200_LAUNCH_CHECK();
  int m = 0;
  B200_CUDA(cudaMemcpyAsync(&m, cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  if (m) {
    int *cand = nullptr;
    B200_TRY(b200_dalloc<int>(h, &cand, m));
    ghost_append_kernel<<<grid, 256, 0, h->stream>>>(nnz, d_j, first, n_owned, cnt + 1, cand);
    B200_LAUNCH_CHECK();
    B200_TRY(sort_unique(h, cand, m, ghost, ng));
    B200_TRY(b200_dfree(h, cand));
  }
  B200_TRY(b200_dfree(h, cnt));
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// halo plan
// ------------------------------------------------------------------------------------------------
int b200_halo_build(b200_handle h, b200_comm c, const std::vector<int> &starts, int *d_ghost, int ng, b200_halo_s **out) {
  const int R = b200_comm_size(c), me = b200_comm_rank(c);
  b200_halo_s *p = new b200_halo_s();
  p->first = starts[me]; p->n_owned = starts[me + 1] - starts[me];
  p->ng = ng; p->d_ghost_gid = d_ghost;
  p->recv_cnt.assign(R, 0); p->recv_off.assign(R + 1, 0); p->send_cnt.assign(R, 0); p->send_off.assign(R + 1, 0);
  std::vector<int> hg(ng);
  if (ng) {
    B200_CUDA(cudaMemcpyAsync(hg.data(), d_ghost, sizeof(int) * (size_t)ng, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  for (int r = 0; r < R; r++) {
    auto lo = std::lower_bound(hg.begin(), hg.end(), starts[r]), hi = std::lower_bound(hg.begin(), hg.end(), starts[r + 1]);
    p->recv_cnt[r] = (int)(hi - lo);
    p->recv_off[r] = (int)(lo - hg.begin());
  }
  p->recv_off[R] = ng;
  if (p->recv_cnt[me]) { delete p; B200_FAIL("halo: an owned id was listed as a ghost"); }
  std::vector<int> all((size_t)R * R);
  B200_TRY(b200_comm_allgather_host(h, c, p->recv_cnt.data(), sizeof(int) * R, all.data()));
  for (int r = 0; r < R; r++) { p->send_cnt[r] = all[(size_t)r * R + me]; p->send_off[r + 1] = p->send_off[r] + p->send_cnt[r]; }
  for (size_t k = 0; k < all.size(); k++) if (all[k]) p->any_traffic = true;
  p->comm = c;
  p->all_cnt = all;
  p->n_send = p->send_off[R];
  B200_TRY(b200_dalloc<int>(h, &p->d_send_idx, p->n_send));
  char *buf = nullptr;
  B200_TRY(b200_dalloc<char>(h, &buf, (size_t)8 * (p->n_send > ng ? p->n_send : ng) + 8));
  p->d_send_buf = buf;
  // requests: I send the ghost ids I need to their owners; they become the owners' send lists
  std::vector<b200_xfer> sends, recvs;
  for (int r = 0; r < R; r++) {
    if (p->recv_cnt[r]) sends.push_back({r, d_ghost + p->recv_off[r], sizeof(int) * (size_t)p->recv_cnt[r], 0});
    if (p->send_cnt[r]) recvs.push_back({r, p->d_send_idx + p->send_off[r], sizeof(int) * (size_t)p->send_cnt[r], 0});
  }
  B200_TRY(b200_comm_exchange(h, c, sends, recvs));
  if (p->n_send) {
    sub_kernel<<<b200_grid(p->n_send, 256), 256, 0, h->stream>>>(p->n_send, p->d_send_idx, p->first);
    B200_LAUNCH_CHECK();
  }
  *out = p;
  return 0;
}
static void halo_args_delete(void *a);
void b200_halo_free(b200_handle h, b200_halo_s *p) {
  if (!p) return;
  if (p->p2p_state == 1) { b200_comm_p2p_free(p->comm, p->p2p_off, p->p2p_bytes); halo_args_delete(p->p2p_args); }
  b200_dfree(h, p->d_ghost_gid); b200_dfree(h, p->d_send_idx); b200_dfree(h, p->d_send_buf);
  delete p;
}

// ---- direct halo exchange: ONE kernel packs and pushes into the neighbours' receive buffers over NVLink, waits for its own
// arrivals and copies them into the ghost tail ------------------------------------------------------------------------------
// (hypre_ParCSRCommHandleCreate job 1 + hypre_ParCSRCommHandleDestroy, par_csr_communication.c:307-631, without a message
// library on the data path.)
//
// Protocol: every double travels as two self-validating 8-byte words {32 data bits, 32-bit sequence number}, written with
// one 16-byte store.  The receiver polls the element itself until both halves carry the sequence number of this exchange: no
// flag behind the data, hence NO system-scope fence on either side (a fence.sys costs microseconds with NVLink stores in
// flight; the first version of this kernel spent most of its 20 us in them).  Exchange number k of a plan -- counted on the
// device, in the plan's own region, so every kernel argument is fixed for the life of the plan and the exchange can sit in a
// CUDA graph -- uses receive buffer k & 1 and sequence number k + 1; a sender overwrites a buffer only after the receiver
// acknowledged the exchange that used it before (k - 2).  Acknowledgements are plain 8-byte counters.
namespace {
struct HaloArgs {
  int n_push, n_pull, cap;
  int send_off[B200_P2P_MAXPEER + 1];                  // entry ranges of the peers in the send list
  uint4 *push_buf[B200_P2P_MAXPEER];                   // where my block starts in the peer's receive buffer (parity 0)
  const unsigned long long *push_ack[B200_P2P_MAXPEER];   // my acknowledgement slot for that peer
  unsigned long long *pull_ack[B200_P2P_MAXPEER];      // the source's acknowledgement slot for me
  const uint4 *rbuf;                                   // my receive buffer (parity 0)
  unsigned long long *cnt;                             // local: exchanges completed under this plan
  unsigned *ctr;                                       // local: CTA counter
};
__device__ __forceinline__ unsigned long long p2p_ld_relaxed(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_relaxed(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 p2p_ld_pair(const uint4 *p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void p2p_st_pair(uint4 *p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void p2p_timeout(unsigned long long t0, unsigned long long timeout_ns, unsigned long long *dbg, int kind,
                                            unsigned long long want, unsigned long long have) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  if (t - t0 > timeout_ns) {                      // a peer that never arrives: an error at the next synchronisation, not a hang
    if (dbg) { dbg[1] = threadIdx.x; dbg[2] = want; dbg[3] = have; dbg[0] = kind; __threadfence_system(); }
    __trap();
  }
}
// grid <= number of SMs: every CTA is resident, so a CTA that waits for a peer never keeps a sibling from being scheduled
__global__ void __launch_bounds__(256)
halo_exchange_kernel(int n_send, const int *__restrict__ idx, const double *__restrict__ src, int ng, double *__restrict__ ghost_out,
                     HaloArgs a, unsigned long long timeout_ns, unsigned long long *dbg) {
  const unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(a.cnt);
  const unsigned seq = (unsigned)(k + 1);
  const unsigned long long ack_need = (k >= 2) ? k - 1 : 0;
  const size_t par = (size_t)(k & 1) * (size_t)a.cap;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  if (n_send) {
    if ((int)threadIdx.x < a.n_push && ack_need) {
      unsigned spins = 0;
      unsigned long long have;
      while ((have = p2p_ld_relaxed(a.push_ack[threadIdx.x])) < ack_need)
        if ((++spins & 1023u) == 0) p2p_timeout(t0, timeout_ns, dbg, 1, ack_need, have);
    }
    __syncthreads();
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_send; e += gridDim.x * blockDim.x) {
      int q = 0;
      while (e >= a.send_off[q + 1]) q++;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(src[idx[e]]);
      p2p_st_pair(a.push_buf[q] + par + (e - a.send_off[q]), make_uint4((unsigned)bits, seq, (unsigned)(bits >> 32), seq));
    }
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ng; e += gridDim.x * blockDim.x) {
    const uint4 *slot = a.rbuf + par + e;
    uint4 w = p2p_ld_pair(slot);
    unsigned spins = 0;
    while (w.y != seq || w.w != seq) {
      if ((++spins & 1023u) == 0) p2p_timeout(t0, timeout_ns, dbg, 2, seq, w.y);
      w = p2p_ld_pair(slot);
    }
    ghost_out[e] = __longlong_as_double((long long)(((unsigned long long)w.z << 32) | w.x));
  }
  // acknowledge and count: the last CTA has seen every sibling read k and finish its copies (loads have returned before the
  // barrier; the device-scope fence keeps the acknowledgement behind them)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(a.ctr, 1u);
    if (t == gridDim.x - 1) {
      *a.ctr = 0;
      __threadfence();
      for (int q = 0; q < a.n_pull; q++) p2p_st_relaxed(a.pull_ack[q], k + 1);
      *reinterpret_cast<volatile unsigned long long *>(a.cnt) = k + 1;
    }
  }
}
}  // namespace

static void halo_args_delete(void *a) { delete static_cast<HaloArgs *>(a); }
static unsigned long long halo_timeout_ns() {
  static const unsigned long long t = [] { const char *e = getenv("B200_P2P_TIMEOUT_S"); int s = e ? atoi(e) : 30; return (unsigned long long)(s > 0 ? s : 30) * 1000000000ull; }();
  return t;
}

// collective and deterministic: every rank takes the same decision from the plan's (allgathered) count matrix
static void halo_p2p_enable(b200_handle h, b200_comm c, b200_halo_s *p) {
  p->p2p_state = -1;
  if (!b200_comm_p2p_ok(c)) return;
  const int R = b200_comm_size(c), me = b200_comm_rank(c);
  int cap = 0, maxpeers = 0;
  for (int r = 0; r < R; r++) {
    int ng_r = 0, nrecv = 0, nsend = 0;
    for (int s2 = 0; s2 < R; s2++) {
      ng_r += p->all_cnt[(size_t)r * R + s2];
      if (p->all_cnt[(size_t)r * R + s2]) nrecv++;
      if (p->all_cnt[(size_t)s2 * R + r]) nsend++;
    }
    cap = std::max(cap, ng_r);
    maxpeers = std::max(maxpeers, std::max(nrecv, nsend));
  }
  if (maxpeers > B200_P2P_MAXPEER) return;
  cap = (cap + 31) & ~31;
  const size_t slot = sizeof(unsigned long long) * B200_P2P_SLOT;
  const size_t bytes = sizeof(uint4) * 2 * (size_t)cap + slot * R + slot * 2;
  const size_t off = b200_comm_p2p_alloc(h, c, bytes);
  if (off == (size_t)-1) return;
  p->p2p_off = off; p->p2p_bytes = bytes; p->p2p_cap = cap; p->p2p_state = 1;
  // the kernel's argument block, fixed for the life of the plan
  const size_t off_ack = off + sizeof(uint4) * 2 * (size_t)cap, off_loc = off_ack + slot * R;
  char *mine = b200_comm_p2p_base(c, me);
  HaloArgs *a = new HaloArgs();
  memset(a, 0, sizeof *a);
  a->cap = cap;
  for (int r = 0; r < R; r++) {
    if (!p->send_cnt[r]) continue;
    char *pb = b200_comm_p2p_base(c, r);
    int roff = 0;                                              // where my block starts in r's ghost array (ghosts sorted by owner)
    for (int s2 = 0; s2 < me; s2++) roff += p->all_cnt[(size_t)r * R + s2];
    const int q = a->n_push++;
    a->push_buf[q] = reinterpret_cast<uint4 *>(pb + off) + roff;
    a->push_ack[q] = reinterpret_cast<const unsigned long long *>(mine + off_ack + slot * r);
    a->send_off[q + 1] = a->send_off[q] + p->send_cnt[r];
  }
  for (int r = 0; r < R; r++) {
    if (!p->recv_cnt[r]) continue;
    char *pb = b200_comm_p2p_base(c, r);
    a->pull_ack[a->n_pull++] = reinterpret_cast<unsigned long long *>(pb + off_ack + slot * me);
  }
  a->rbuf = reinterpret_cast<const uint4 *>(mine + off);
  a->cnt = reinterpret_cast<unsigned long long *>(mine + off_loc);
  a->ctr = reinterpret_cast<unsigned *>(mine + off_loc + slot);
  p->p2p_args = a;
}

static int halo_forward_p2p(b200_handle h, b200_comm c, b200_halo_s *p, const double *owned, double *ghost_out) {
  if (!p->n_send && !p->ng) return 0;
  const HaloArgs *a = static_cast<const HaloArgs *>(p->p2p_args);
  int grid = b200_grid(std::max(p->n_send, p->ng), 256);
  if (grid > h->num_sm) grid = h->num_sm;
  halo_exchange_kernel<<<grid, 256, 0, h->stream>>>(p->n_send, p->d_send_idx, owned, p->ng, ghost_out, *a, halo_timeout_ns(),
                                                    g_b200_p2p_dbg);
  B200_LAUNCH_CHECK();
  return 0;
}

template <class T>
static int halo_forward(b200_handle h, b200_comm c, b200_halo_s *p, const T *owned, T *ghost_out) {
  const int R = b200_comm_size(c);
  T *buf = reinterpret_cast<T *>(p->d_send_buf);
  if (p->n_send) {
    pack_kernel<T><<<b200_grid(p->n_send, 256), 256, 0, h->stream>>>(p->n_send, p->d_send_idx, owned, buf);
    B200_LAUNCH_CHECK();
  }
  std::vector<b200_xfer> sends, recvs;
  for (int r = 0; r < R; r++) {
    if (p->send_cnt[r]) sends.push_back({r, buf + p->send_off[r], sizeof(T) * (size_t)p->send_cnt[r], 0});
    if (p->recv_cnt[r]) recvs.push_back({r, ghost_out + p->recv_off[r], sizeof(T) * (size_t)p->recv_cnt[r], 0});
  }
  return b200_comm_exchange(h, c, sends, recvs);
}
int b200_halo_forward_i32(b200_handle h, b200_comm c, b200_halo_s *p, const int *o, int *g) { return halo_forward<int>(h, c, p, o, g); }
int b200_halo_forward_f64(b200_handle h, b200_comm c, b200_halo_s *p, const double *o, double *g) {
  if (p->p2p_state == 0) halo_p2p_enable(h, c, p);        // first exchange of doubles under this plan: every rank is here
  if (p->p2p_state == 1) return halo_forward_p2p(h, c, p, o, g);
  return halo_forward<double>(h, c, p, o, g);
}

static int halo_reverse_i32(b200_handle h, b200_comm c, b200_halo_s *p, const int *ghost_in, int *owned, int op) {
  const int R = b200_comm_size(c);
  int *buf = reinterpret_cast<int *>(p->d_send_buf);
  std::vector<b200_xfer> sends, recvs;
  for (int r = 0; r < R; r++) {
    if (p->recv_cnt[r]) sends.push_back({r, const_cast<int *>(ghost_in) + p->recv_off[r], sizeof(int) * (size_t)p->recv_cnt[r], 0});
    if (p->send_cnt[r]) recvs.push_back({r, buf + p->send_off[r], sizeof(int) * (size_t)p->send_cnt[r], 0});
  }
  B200_TRY(b200_comm_exchange(h, c, sends, recvs));
  if (p->n_send) {
    if (op == 0) unpack_add_kernel<<<b200_grid(p->n_send, 256), 256, 0, h->stream>>>(p->n_send, p->d_send_idx, buf, owned);
    else unpack_clear_kernel<<<b200_grid(p->n_send, 256), 256, 0, h->stream>>>(p->n_send, p->d_send_idx, buf, owned);
    B200_LAUNCH_CHECK();
  }
  return 0;
}
int b200_halo_reverse_add_i32(b200_handle h, b200_comm c, b200_halo_s *p, const int *g, int *o) { return halo_reverse_i32(h, c, p, g, o, 0); }
int b200_halo_reverse_clear_i32(b200_handle h, b200_comm c, b200_halo_s *p, const int *g, int *o) { return halo_reverse_i32(h, c, p, g, o, 1); }

// Fetch the rows of M (owned rows, any column ids) that correspond to the plan's ghosts, in ghost
// order, entry order preserved.  hypre_ParCSRMatrixExtractBExt (par_csr_matop.c:1655).
static int fetch_rows(b200_handle h, b200_comm c, b200_halo_s *p, b200_csr M, b200_csr *out) {
  const int R = b200_comm_size(c);
  const bool with_data = M->a != nullptr;
  tr(h, nullptr);
  // owner side: lengths and packed entries of the requested rows
  int *slen = nullptr;
  B200_TRY(b200_dalloc<int>(h, &slen, (size_t)p->n_send + 1));
  rowlen_kernel<<<b200_grid((size_t)p->n_send + 1, 256), 256, 0, h->stream>>>(p->n_send, p->d_send_idx, M->i, slen);
  B200_LAUNCH_CHECK();
  // receiver side: ghost row lengths
  b200_csr E = nullptr;
  int *glen = nullptr;
  B200_TRY(b200_dalloc<int>(h, &glen, (size_t)p->ng + 1));
  B200_CUDA(cudaMemsetAsync(glen + p->ng, 0, sizeof(int), h->stream));
  {
    std::vector<b200_xfer> sends, recvs;
    for (int r = 0; r < R; r++) {
      if (p->send_cnt[r]) sends.push_back({r, slen + p->send_off[r], sizeof(int) * (size_t)p->send_cnt[r], 0});
      if (p->recv_cnt[r]) recvs.push_back({r, glen + p->recv_off[r], sizeof(int) * (size_t)p->recv_cnt[r], 0});
    }
    B200_TRY(b200_comm_exchange(h, c, sends, recvs));
  }
  B200_TRY(b200_exclusive_scan_inplace(h, slen, (size_t)p->n_send + 1));
  B200_TRY(b200_exclusive_scan_inplace(h, glen, (size_t)p->ng + 1));
  std::vector<int> hs(p->n_send + 1), hg(p->ng + 1);
  B200_CUDA(cudaMemcpyAsync(hs.data(), slen, sizeof(int) * ((size_t)p->n_send + 1), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(hg.data(), glen, sizeof(int) * ((size_t)p->ng + 1), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  const int send_nnz = hs[p->n_send], recv_nnz = hg[p->ng];
  int *bj = nullptr;
  double *ba = nullptr;
  B200_TRY(b200_dalloc<int>(h, &bj, send_nnz));
  if (with_data) B200_TRY(b200_dalloc<double>(h, &ba, send_nnz));
  if (p->n_send) {
    pack_rows_kernel<<<b200_grid(p->n_send, 128), 128, 0, h->stream>>>(p->n_send, p->d_send_idx, M->i, M->j, M->a, slen, bj, ba);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_csr_alloc(h, p->ng, M->ncols, recv_nnz, with_data, &E));
  B200_CUDA(cudaMemcpyAsync(E->i, glen, sizeof(int) * ((size_t)p->ng + 1), cudaMemcpyDeviceToDevice, h->stream));
  {
    std::vector<b200_xfer> sends, recvs;
    for (int r = 0; r < R; r++) {
      const int sn = hs[p->send_off[r + 1]] - hs[p->send_off[r]], so = hs[p->send_off[r]];
      const int rn = hg[p->recv_off[r + 1]] - hg[p->recv_off[r]], ro = hg[p->recv_off[r]];
      if (sn) sends.push_back({r, bj + so, sizeof(int) * (size_t)sn, 0});
      if (rn) recvs.push_back({r, E->j + ro, sizeof(int) * (size_t)rn, 0});
      if (with_data) {                            // columns and values travel in ONE grouped exchange (second message per peer)
        if (sn) sends.push_back({r, ba + so, sizeof(double) * (size_t)sn, 1});
        if (rn) recvs.push_back({r, E->a + ro, sizeof(double) * (size_t)rn, 1});
      }
    }
    B200_TRY(b200_comm_exchange(h, c, sends, recvs));
  }
  B200_TRY(b200_dfree(h, slen)); B200_TRY(b200_dfree(h, glen)); B200_TRY(b200_dfree(h, bj)); B200_TRY(b200_dfree(h, ba));
  *out = E;
  tr(h, "  fetch_rows");
  return 0;
}

// stack two CSR blocks (rows of `top` then rows of `bot`)
static int stack_rows(b200_handle h, b200_csr top, b200_csr bot, b200_csr *out) {
  const bool with_data = top->a != nullptr;
  b200_csr S = nullptr;
  B200_TRY(b200_csr_alloc(h, top->nrows + bot->nrows, top->ncols, top->nnz + bot->nnz, with_data, &S));
  const int m = std::max(top->nrows, bot->nrows) + 1;
  concat_rowptr_kernel<<<b200_grid(m, 256), 256, 0, h->stream>>>(top->nrows, top->i, bot->nrows, bot->i, S->i);
  B200_LAUNCH_CHECK();
  if (top->nnz) B200_CUDA(cudaMemcpyAsync(S->j, top->j, sizeof(int) * (size_t)top->nnz, cudaMemcpyDeviceToDevice, h->stream));
  if (bot->nnz) B200_CUDA(cudaMemcpyAsync(S->j + top->nnz, bot->j, sizeof(int) * (size_t)bot->nnz, cudaMemcpyDeviceToDevice, h->stream));
  if (with_data) {
    if (top->nnz) B200_CUDA(cudaMemcpyAsync(S->a, top->a, sizeof(double) * (size_t)top->nnz, cudaMemcpyDeviceToDevice, h->stream));
    if (bot->nnz) B200_CUDA(cudaMemcpyAsync(S->a + top->nnz, bot->a, sizeof(double) * (size_t)bot->nnz, cudaMemcpyDeviceToDevice, h->stream));
  }
  *out = S;
  return 0;
}

// rows of `top` (columns in [owned | ghosts of the first ring], the localized form) with the ghost columns renumbered through
// pos[] (their position in a larger extended index space), then the rows of `bot` (already in that space)
namespace {
__global__ void remap_cols_kernel(int nnz, const int *__restrict__ jl, int n_owned, const int *__restrict__ pos, int *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int c = jl[k];
  out[k] = c < n_owned ? c : pos[c - n_owned];
}
}  // namespace
static int stack_rows_remap(b200_handle h, b200_csr top, int n_owned, const int *d_pos, b200_csr bot, int ncols, b200_csr *out) {
  const bool with_data = top->a != nullptr;
  b200_csr S = nullptr;
  B200_TRY(b200_csr_alloc(h, top->nrows + bot->nrows, ncols, top->nnz + bot->nnz, with_data, &S));
  const int m = std::max(top->nrows, bot->nrows) + 1;
  concat_rowptr_kernel<<<b200_grid(m, 256), 256, 0, h->stream>>>(top->nrows, top->i, bot->nrows, bot->i, S->i);
  B200_LAUNCH_CHECK();
  if (top->nnz) {
    remap_cols_kernel<<<b200_grid(top->nnz, 256), 256, 0, h->stream>>>(top->nnz, top->j, n_owned, d_pos, S->j);
    B200_LAUNCH_CHECK();
  }
  if (bot->nnz) B200_CUDA(cudaMemcpyAsync(S->j + top->nnz, bot->j, sizeof(int) * (size_t)bot->nnz, cudaMemcpyDeviceToDevice, h->stream));
  if (with_data) {
    if (top->nnz) B200_CUDA(cudaMemcpyAsync(S->a, top->a, sizeof(double) * (size_t)top->nnz, cudaMemcpyDeviceToDevice, h->stream));
    if (bot->nnz) B200_CUDA(cudaMemcpyAsync(S->a + top->nnz, bot->a, sizeof(double) * (size_t)bot->nnz, cudaMemcpyDeviceToDevice, h->stream));
  }
  *out = S;
  return 0;
}

// E has one row per first-ring ghost (sorted by gid); returns a CSR with `nslots` rows where row
// pos[k] - n_owned holds E's row k and every other row is empty.  pos is increasing, so the entry
// arrays are copied unchanged.
namespace {
__global__ void spread_len_kernel(int ng, const int *__restrict__ E_i, const int *__restrict__ pos, int n_owned, int *__restrict__ len) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < ng) len[pos[k] - n_owned] = E_i[k + 1] - E_i[k];
}
}  // namespace
static int spread_rows(b200_handle h, b200_csr E, const int *d_pos, int n_owned, int nslots, b200_csr *out) {
  b200_csr Q = nullptr;
  B200_TRY(b200_csr_alloc(h, nslots, E->ncols, E->nnz, E->a != nullptr, &Q));
  B200_CUDA(cudaMemsetAsync(Q->i, 0, sizeof(int) * ((size_t)nslots + 1), h->stream));
  if (E->nrows) {
    spread_len_kernel<<<b200_grid(E->nrows, 256), 256, 0, h->stream>>>(E->nrows, E->i, d_pos, n_owned, Q->i);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_exclusive_scan_inplace(h, Q->i, (size_t)nslots + 1));
  if (E->nnz) {
    B200_CUDA(cudaMemcpyAsync(Q->j, E->j, sizeof(int) * (size_t)E->nnz, cudaMemcpyDeviceToDevice, h->stream));
    if (E->a) B200_CUDA(cudaMemcpyAsync(Q->a, E->a, sizeof(double) * (size_t)E->nnz, cudaMemcpyDeviceToDevice, h->stream));
  }
  *out = Q;
  return 0;
}

// copy of M with its column ids mapped global -> [owned | ghost] for the given plan
static int localize_copy(b200_handle h, b200_csr M, b200_halo_s *p, int first, int n_owned, bool build_plan, b200_csr *out) {
  b200_csr L = nullptr;
  B200_TRY(b200_csr_alloc(h, M->nrows, n_owned + p->ng, M->nnz, M->a != nullptr, &L));
  B200_CUDA(cudaMemcpyAsync(L->i, M->i, sizeof(int) * ((size_t)M->nrows + 1), cudaMemcpyDeviceToDevice, h->stream));
  if (M->nnz) {
    localize_kernel<<<b200_grid(M->nnz, 256), 256, 0, h->stream>>>(M->nnz, M->j, first, n_owned, p->ng, p->d_ghost_gid, L->j);
    B200_LAUNCH_CHECK();
    if (M->a) B200_CUDA(cudaMemcpyAsync(L->a, M->a, sizeof(double) * (size_t)M->nnz, cudaMemcpyDeviceToDevice, h->stream));
  }
  if (build_plan && L->a) B200_TRY(b200_csr_build_plan(h, L));
  *out = L;
  return 0;
}

// builds M->L and M->halo from M->G.  plan = false: no SpMV plan (an operand of a product only).  consume = true: the columns
// are localized in place and M->G is gone afterwards (nobody fetches rows of this matrix by global id any more; download and
// info rebuild the global ids from the localized form) -- saves a copy of the whole matrix.
static int dist_localize(b200_handle h, b200_comm c, b200_dist_matrix M, bool plan = true, bool consume = false) {
  int *ghost = nullptr, ng = 0;
  tr(h, nullptr);
  B200_TRY(ghost_columns(h, M->G->j, M->G->nnz, M->first_col, M->n_owned_cols, &ghost, &ng));
  tr(h, "  localize: ghost_columns");
  if (!ghost) B200_TRY(b200_dalloc<int>(h, &ghost, 1));
  B200_TRY(b200_halo_build(h, c, M->col_starts, ghost, ng, &M->halo));
  tr(h, "  localize: halo_build");
  if (consume) {
    b200_csr G = M->G;
    if (G->nnz) {
      localize_kernel<<<b200_grid(G->nnz, 256), 256, 0, h->stream>>>(G->nnz, G->j, M->first_col, M->n_owned_cols, M->halo->ng,
                                                                    M->halo->d_ghost_gid, G->j);
      B200_LAUNCH_CHECK();
    }
    G->ncols = M->n_owned_cols + M->halo->ng;
    if (plan && G->a) B200_TRY(b200_csr_build_plan(h, G));
    M->L = G;
    M->G = nullptr;
  } else {
    B200_TRY(localize_copy(h, M->G, M->halo, M->first_col, M->n_owned_cols, plan, &M->L));
  }
  tr(h, "  localize: copy+plan");
  return 0;
}

static int gather_starts(b200_handle h, b200_comm c, int n_local, std::vector<int> *starts) {
  const int R = b200_comm_size(c);
  std::vector<int> all(R);
  B200_TRY(b200_comm_allgather_host(h, c, &n_local, sizeof(int), all.data()));
  starts->assign(R + 1, 0);
  for (int r = 0; r < R; r++) (*starts)[r + 1] = (*starts)[r] + all[r];
  return 0;
}

// ------------------------------------------------------------------------------------------------
// public: generator, info, download, matvec
// ------------------------------------------------------------------------------------------------
// stencil 7: values = centre, x, y, z (GenerateLaplacian); 27: centre, off-centre (GenerateLaplacian27pt);
// 70: centre, x-, y-, z-, x+, y+, z+ (GenerateDifConv, par_difconv.c:15)
static int dist_generate(b200_handle h, b200_comm c, int nx, int ny, int nz, int P, int Q, int R,
                         int stencil, const double *user_values, b200_dist_matrix *out) {
  const int nr = b200_comm_size(c), me = b200_comm_rank(c);
  if (P * Q * R != nr) B200_FAIL("process grid P*Q*R must equal the number of ranks");
  if (stencil != 7 && stencil != 27 && stencil != 70 && stencil != 72) B200_FAIL("stencil must be 7 or 27");
  double values[7] = {0, 0, 0, 0, 0, 0, 0};
  if (stencil == 7) {
    for (int k = 0; k < 4; k++) values[k] = user_values[k];
    for (int k = 1; k < 4; k++) values[k + 3] = user_values[k];
  } else {
    for (int k = 0; k < (stencil == 70 ? 7 : stencil == 72 ? 4 : 2); k++) values[k] = user_values[k];
  }
  if (stencil == 70) stencil = 7;
  const int p = me % P, q = ((me - p) / P) % Q, r = (me - p - P * q) / (P * Q);      // ij.c:7785-7787
  b200_dist_matrix M = new b200_dist_matrix_s();
  B200_TRY(b200_generate_stencil_global(h, nx, ny, nz, P, Q, R, p, q, r, stencil, values, &M->G, &M->first_row));
  M->n = M->G->nrows;
  M->global_rows = M->global_cols = nx * ny * nz;
  M->row_starts.assign(nr + 1, 0);
  for (int k = 0; k < nr; k++) {
    const int pk = k % P, qk = ((k - pk) / P) % Q, rk = (k - pk - P * qk) / (P * Q);
    M->row_starts[k] = b200_box_first_row(nx, ny, nz, P, Q, R, pk, qk, rk);
  }
  M->row_starts[nr] = nx * ny * nz;
  M->col_starts = M->row_starts;
  M->first_col = M->first_row; M->n_owned_cols = M->n;
  if (M->row_starts[me] != M->first_row || M->row_starts[me + 1] - M->row_starts[me] != M->n) B200_FAIL("partition mismatch");
  B200_TRY(dist_localize(h, c, M));
  *out = M;
  return 0;
}
// ---- a row-partitioned operator from the caller's own rows -----------------------------------------------------------
// (what hypre_IJMatrixAssembleParCSR + GenerateDiagAndOffd produce across ranks, IJ_mv/IJMatrix_parcsr.c:2774,
// parcsr_mv/par_csr_matrix.c:1634: this rank's contiguous block of rows, diagonal entry first, global column ids)
__global__ void check_rows_kernel(int n, int first_row, int global_cols, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                  int *__restrict__ bad) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int s = A_i[r], e = A_i[r + 1];
  if (e > s && A_j[s] != first_row + r) atomicAdd(&bad[0], 1);               // diagonal first (csr_matrix.h invariant)
  for (int k = s; k < e; k++)
    if (A_j[k] < 0 || A_j[k] >= global_cols) atomicAdd(&bad[1], 1);
}
static int dist_from_global_rows(b200_handle h, b200_comm c, b200_csr G, b200_dist_matrix *out) {
  const int nr = b200_comm_size(c), me = b200_comm_rank(c);
  b200_dist_matrix M = new b200_dist_matrix_s();
  M->G = G;
  M->n = G->nrows;
  B200_TRY(gather_starts(h, c, M->n, &M->row_starts));
  M->first_row = M->row_starts[me];
  M->global_rows = M->global_cols = M->row_starts[nr];
  M->col_starts = M->row_starts;
  M->first_col = M->first_row; M->n_owned_cols = M->n;
  G->ncols = M->global_cols;
  int *bad = nullptr, hb[2] = {0, 0};
  B200_TRY(b200_dalloc<int>(h, &bad, 2));
  B200_CUDA(cudaMemsetAsync(bad, 0, 2 * sizeof(int), h->stream));
  if (M->n) {
    check_rows_kernel<<<b200_grid(M->n, 256), 256, 0, h->stream>>>(M->n, M->first_row, M->global_cols, G->i, G->j, bad);
    B200_LAUNCH_CHECK();
  }
  B200_CUDA(cudaMemcpyAsync(hb, bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, bad));
  // every rank reaches the collective below even if its own rows are bad: agree on the verdict first
  int verdict[2] = {hb[0], hb[1]};
  std::vector<int> all((size_t)2 * nr);
  B200_TRY(b200_comm_allgather_host(h, c, verdict, 2 * sizeof(int), all.data()));
  long long bad_diag = 0, bad_col = 0;
  for (int r = 0; r < nr; r++) { bad_diag += all[2 * r]; bad_col += all[2 * r + 1]; }
  if (bad_diag) B200_FAIL("dist matrix: a non-empty row does not store its diagonal entry first");
  if (bad_col) B200_FAIL("dist matrix: column index outside [0, global_rows)");
  B200_TRY(dist_localize(h, c, M));
  *out = M;
  return 0;
}
extern "C" int b200_dist_matrix_create_from_host(b200_handle h, b200_comm c, int n_local, const int *h_i, const int *h_j_global,
                                                 const double *h_a, b200_dist_matrix *out) {
  if (n_local < 0 || !h_i || !out) B200_FAIL("dist_matrix_create_from_host: bad argument");
  const int nnz = h_i[n_local];
  b200_csr G = nullptr;
  B200_TRY(b200_csr_alloc(h, n_local, 0, nnz, true, &G));
  B200_CUDA(cudaMemcpyAsync(G->i, h_i, sizeof(int) * ((size_t)n_local + 1), cudaMemcpyHostToDevice, h->stream));
  if (nnz) {
    B200_CUDA(cudaMemcpyAsync(G->j, h_j_global, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
    B200_CUDA(cudaMemcpyAsync(G->a, h_a, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
  }
  B200_CUDA(cudaStreamSynchronize(h->stream));            // the host arrays may be pageable and go out of scope
  return dist_from_global_rows(h, c, G, out);
}
b200_parcsr b200_ij_release_object(b200_ij ij);           // b200_ij.cu
extern "C" int b200_dist_matrix_create_from_ij(b200_handle h, b200_comm c, b200_ij ij, b200_dist_matrix *out) {
  if (!ij || !out) B200_FAIL("dist_matrix_create_from_ij: bad argument");
  b200_parcsr A = nullptr;
  int missing = 0;
  B200_TRY(b200_ij_assemble(h, ij, &A, &missing));          // device-side merge of this rank's records (b200_ij.cu)
  if (missing) B200_FAIL("dist_matrix_create_from_ij: values were set on elements that do not exist");
  A = b200_ij_release_object(ij);
  if (!A) B200_FAIL("dist_matrix_create_from_ij: nothing assembled");
  b200_csr G = A->diag;                                     // rows of this rank, global column ids, diagonal first
  A->diag = nullptr;
  B200_TRY(b200_csr_destroy(h, A->offd));
  delete A;
  return dist_from_global_rows(h, c, G, out);
}

extern "C" int b200_dist_generate_laplacian(b200_handle h, b200_comm c, int nx, int ny, int nz, int P, int Q, int R,
                                            int stencil, const double *values, b200_dist_matrix *out) {
  if (stencil != 7 && stencil != 27) B200_FAIL("stencil must be 7 or 27");
  return dist_generate(h, c, nx, ny, nz, P, Q, R, stencil, values, out);
}
void b200_rotate7pt_values(double alpha, double eps, double *value);      // b200_parcsr.cu
extern "C" int b200_dist_generate_rotate7pt(b200_handle h, b200_comm c, int nx, int ny, int P, int Q, double alpha, double eps,
                                            b200_dist_matrix *out) {
  double v[4];
  b200_rotate7pt_values(alpha, eps, v);
  return dist_generate(h, c, nx, ny, 1, P, Q, 1, 72, v, out);                // rank -> (p, q) as ij.c:9190-9191
}
extern "C" int b200_dist_generate_difconv(b200_handle h, b200_comm c, int nx, int ny, int nz, int P, int Q, int R,
                                          const double values[7], b200_dist_matrix *out) {
  return dist_generate(h, c, nx, ny, nz, P, Q, R, 70, values, out);
}

extern "C" int b200_csr_stream_bytes_per_entry(b200_csr A);
extern "C" int b200_dist_matrix_stream_bytes_per_entry(b200_dist_matrix A) { return (A && A->L) ? b200_csr_stream_bytes_per_entry(A->L) : 0; }

extern "C" int b200_dist_matrix_destroy(b200_handle h, b200_dist_matrix M) {
  if (!M) return 0;
  B200_TRY(b200_csr_destroy(h, M->G));
  B200_TRY(b200_csr_destroy(h, M->L));
  if (M->Ls) B200_TRY(b200_csr_destroy(h, M->Ls));
  b200_halo_free(h, M->halo);
  delete M;
  return 0;
}
extern "C" int b200_dist_matrix_info(b200_dist_matrix M, int *local_rows, int *first_row, int *global_rows, int *local_nnz,
                                     int *n_ghost, int *first_col, int *global_cols) {
  if (!M) B200_FAIL("null matrix");
  if (local_rows) *local_rows = M->n;
  if (first_row) *first_row = M->first_row;
  if (global_rows) *global_rows = M->global_rows;
  if (local_nnz) *local_nnz = M->G ? M->G->nnz : (M->L ? M->L->nnz : 0);
  if (n_ghost) *n_ghost = M->halo ? M->halo->ng : 0;
  if (first_col) *first_col = M->first_col;
  if (global_cols) *global_cols = M->global_cols;
  return 0;
}
extern "C" int b200_dist_matrix_download(b200_handle h, b200_dist_matrix M, int *h_i, int *h_j, double *h_a) {
  if (!M) B200_FAIL("null matrix");
  if (M->G) return b200_csr_download(h, M->G, h_i, h_j, h_a);
  // setup form already released: rebuild global ids from the localized form
  int *jg = nullptr;
  B200_TRY(b200_dalloc<int>(h, &jg, M->L->nnz));
  if (M->L->nnz) {
    globalize_kernel<<<b200_grid(M->L->nnz, 256), 256, 0, h->stream>>>(M->L->nnz, M->L->j, M->first_col, M->n_owned_cols,
                                                                      M->halo->d_ghost_gid, jg);
    B200_LAUNCH_CHECK();
  }
  if (h_i) B200_CUDA(cudaMemcpyAsync(h_i, M->L->i, sizeof(int) * ((size_t)M->n + 1), cudaMemcpyDeviceToHost, h->stream));
  if (h_j && M->L->nnz) B200_CUDA(cudaMemcpyAsync(h_j, jg, sizeof(int) * (size_t)M->L->nnz, cudaMemcpyDeviceToHost, h->stream));
  if (h_a && M->L->nnz) B200_CUDA(cudaMemcpyAsync(h_a, M->L->a, sizeof(double) * (size_t)M->L->nnz, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, jg));
  return 0;
}

// halo of x (job 1) straight into the ghost tail, then ONE kernel over [owned | ghost]
static int dist_spmv(b200_handle h, b200_comm c, b200_dist_matrix M, double *x, double *y, int mode, double alpha, double beta,
                     const double *b, const double *d, const char *what = "spmv", int level = -1) {
  if (M->halo->any_traffic) {
    b200_prof_scope ps(h, "halo", level);
    B200_TRY(b200_halo_forward_f64(h, c, M->halo, x, x + M->n_owned_cols));
  }
  b200_prof_scope ps(h, what, level);
  return b200_csr_spmv_epi(h, M->Ls ? M->Ls : M->L, x, y, mode, alpha, beta, b, d);
}
extern "C" int b200_dist_matvec(b200_handle h, b200_comm c, double alpha, b200_dist_matrix M, double *d_x, double beta,
                                const double *d_b, double *d_y) {
  if (!M || !M->L) B200_FAIL("dist_matvec: matrix not localized");
  return dist_spmv(h, c, M, d_x, d_y, 0, alpha, beta, d_b, nullptr);
}

// hypre_BoomerAMGRelax types 8/13/14 across ranks (par_relax.c:4340-5124): halo of u (job 1), then Gauss-Seidel
// inside each of the rank's `blocks` blocks; everything outside a block -- other blocks, other ranks --
// enters with its pre-sweep value.  l1 norms: option 4 (ams.c:3560-3625), computed here.
extern "C" int b200_dist_relax_gs(b200_handle h, b200_comm c, b200_dist_matrix M, int relax_type, int blocks, const double *d_f,
                                  double *d_u) {
  if (!M || !M->L) B200_FAIL("dist_relax_gs: matrix not localized");
  if (relax_type != 8 && relax_type != 13 && relax_type != 14) B200_FAIL("dist_relax_gs: relax types 8, 13, 14");
  b200_csr A = M->L;
  if (A->gs && b200_gs_plan_blocks(A->gs) != blocks) { B200_TRY(b200_gs_plan_destroy(h, A->gs)); A->gs = nullptr; }
  if (!A->gs) B200_TRY(b200_gs_plan_create(h, A, blocks, &A->gs));
  double *l1 = nullptr;
  B200_TRY(b200_dalloc<double>(h, &l1, M->n));
  B200_TRY(b200_l1_norms_blocks(h, A, 4, blocks, l1));
  if (M->halo->any_traffic) B200_TRY(b200_halo_forward_f64(h, c, M->halo, d_u, d_u + M->n_owned_cols));
  B200_TRY(b200_gs_relax(h, A->gs, A, relax_type, false, d_f, l1, d_u));
  B200_TRY(b200_dfree(h, l1));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// distributed hierarchy
// ------------------------------------------------------------------------------------------------
struct dist_level {
  b200_dist_matrix A = nullptr;      // level 0 borrowed
  b200_dist_matrix P = nullptr;      // rows: fine (this level), cols: coarse
  b200_dist_matrix R = nullptr;      // rows: coarse, cols: fine (P^T)
  int *cf = nullptr;                 // [n]
  double *l1 = nullptr;
  double *F = nullptr, *U = nullptr, *T = nullptr;   // capacity n + max ghosts
  int n = 0, cap = 0;
};
struct b200_dist_amg_s {
  std::vector<dist_level> lv;
  double *Vtemp = nullptr;
  int vtemp_cap = 0;
  double *ge_A = nullptr, *ge_f = nullptr;   // dense coarsest matrix (n x n) + scratch + gathered rhs
  int ge_n = 0;
  std::vector<int> ge_starts;
  bool coarse_ge = false;
  double relax_wt = 1.0;
  bool gs = false;                   // l1 hybrid Gauss-Seidel (8/13/14): Gauss-Seidel inside the rank's blocks,
  int relax_down = 18, relax_up = 18, gs_blocks = 1;   // pre-sweep values across blocks and ranks (par_relax.c:4352-4372)
  double setup_ms = 0;
  // Replicated tail (hypre's seq_threshold, par_amg_setup.c:2880-2898 / par_amg.c hypre_seqAMGSetup): once a level has
  // no more than SeqThreshold rows in total, it is gathered onto EVERY rank and everything below it is ONE single-GPU
  // hierarchy built and cycled redundantly -- the same kernels on the same data give the same bits on every rank, and
  // the levels where a halo exchange costs more than the kernels need no exchange at all.
  b200_amg tail = nullptr;           // its level 0 is this hierarchy's level lv.size() - 1
  b200_parcsr tailA = nullptr;
  b200_halo_s *tail_plan = nullptr;  // brings every other rank's piece of the level's right-hand side
  double *tail_F = nullptr, *tail_U = nullptr, *tail_G = nullptr;
  int tail_N = 0, tail_first = 0, tail_rank = 0;
  std::vector<b200_dist_matrix> views;   // per-level views handed out by the level accessors (owned here)
};

// distributed transpose of P (global coarse cols) -> R rows = local coarse, cols = global fine ids,
// entries of a row ordered by ascending fine row (csr_matop.c:740-767 order, partition independent)
static int dist_transpose(b200_handle h, b200_comm c, b200_dist_matrix P, const std::vector<int> &coarse_starts, b200_csr *out) {
  const int R = b200_comm_size(c), me = b200_comm_rank(c);
  const int nnz = P->G->nnz, n = P->n;
  const int nc = coarse_starts[me + 1] - coarse_starts[me], cfirst = coarse_starts[me];
  tr(h, nullptr);
  // owner of every entry's column; stable sort by owner keeps the (row, entry) order inside a bucket
  int *owner = nullptr, *rows = nullptr, *idx = nullptr, *owner_s = nullptr, *perm = nullptr, *d_starts = nullptr;
  B200_TRY(b200_dalloc<int>(h, &owner, nnz)); B200_TRY(b200_dalloc<int>(h, &rows, nnz)); B200_TRY(b200_dalloc<int>(h, &idx, nnz));
  B200_TRY(b200_dalloc<int>(h, &owner_s, nnz)); B200_TRY(b200_dalloc<int>(h, &perm, nnz));
  B200_TRY(b200_dalloc<int>(h, &d_starts, R + 1));
  B200_CUDA(cudaMemcpyAsync(d_starts, coarse_starts.data(), sizeof(int) * (R + 1), cudaMemcpyHostToDevice, h->stream));
  int *cntr = nullptr;
  B200_TRY(b200_dalloc<int>(h, &cntr, R + 1));
  B200_CUDA(cudaMemsetAsync(cntr, 0, sizeof(int) * (R + 1), h->stream));
  if (nnz) {
    owner_kernel<<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, P->G->j, R, d_starts, owner);
    B200_LAUNCH_CHECK();
    expand_rows_g_kernel<<<b200_grid(n, 128), 128, 0, h->stream>>>(n, P->G->i, P->first_row, rows);
    B200_LAUNCH_CHECK();
    iota_kernel2<<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, idx);
    B200_LAUNCH_CHECK();
    {
      const int want = b200_grid(nnz, 256), cap = h->num_sm * 8;
      hist_few_bins_kernel<<<want < cap ? want : cap, 256, sizeof(int) * (R + 1), h->stream>>>(nnz, owner, R + 1, cntr);
      B200_LAUNCH_CHECK();
    }
    int bits = 1;
    while ((1 << bits) < R) bits++;
    size_t tb = 0;
    B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, owner, owner_s, idx, perm, nnz, 0, bits, h->stream));
    char *tmp = nullptr;
    B200_TRY(b200_dalloc<char>(h, &tmp, tb));
    B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, owner, owner_s, idx, perm, nnz, 0, bits, h->stream));
    ++g_b200_launches;
    B200_TRY(b200_dfree(h, tmp));
  }
  tr(h, " T: owner+sort");
  std::vector<int> scnt(R + 1, 0);
  B200_CUDA(cudaMemcpyAsync(scnt.data(), cntr, sizeof(int) * R, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  // bucketed triplets (fine row, coarse col, value)
  int *s_row = nullptr, *s_col = nullptr;
  double *s_val = nullptr;
  B200_TRY(b200_dalloc<int>(h, &s_row, nnz)); B200_TRY(b200_dalloc<int>(h, &s_col, nnz)); B200_TRY(b200_dalloc<double>(h, &s_val, nnz));
  if (nnz) {
    gather_kernel<int><<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, perm, rows, s_row);
    gather_kernel<int><<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, perm, P->G->j, s_col);
    gather_kernel<double><<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, perm, P->G->a, s_val);
    g_b200_launches += 3;
  }
  std::vector<int> all((size_t)R * R), rcnt(R), soff(R + 1, 0), roff(R + 1, 0);
  B200_TRY(b200_comm_allgather_host(h, c, scnt.data(), sizeof(int) * R, all.data()));
  for (int r = 0; r < R; r++) { rcnt[r] = all[(size_t)r * R + me]; soff[r + 1] = soff[r] + scnt[r]; roff[r + 1] = roff[r] + rcnt[r]; }
  const int m = roff[R];
  int *r_row = nullptr, *r_col = nullptr;
  double *r_val = nullptr;
  B200_TRY(b200_dalloc<int>(h, &r_row, m)); B200_TRY(b200_dalloc<int>(h, &r_col, m)); B200_TRY(b200_dalloc<double>(h, &r_val, m));
  {
    // the rank's own bucket (almost everything) is a device copy; rows, columns and values for the other ranks travel in ONE
    // grouped exchange (three messages per peer)
    std::vector<b200_xfer> sends, recvs;
    for (int pass = 0; pass < 3; pass++) {
      for (int r = 0; r < R; r++) {
        const size_t es = pass == 2 ? sizeof(double) : sizeof(int);
        char *sp = pass == 0 ? (char *)(s_row + soff[r]) : pass == 1 ? (char *)(s_col + soff[r]) : (char *)(s_val + soff[r]);
        char *rp = pass == 0 ? (char *)(r_row + roff[r]) : pass == 1 ? (char *)(r_col + roff[r]) : (char *)(r_val + roff[r]);
        if (r == me) {
          if (scnt[r] != rcnt[r]) B200_FAIL("transpose: own bucket size mismatch");
          if (scnt[r]) B200_CUDA(cudaMemcpyAsync(rp, sp, es * (size_t)scnt[r], cudaMemcpyDeviceToDevice, h->stream));
          continue;
        }
        if (scnt[r]) sends.push_back({r, sp, es * (size_t)scnt[r], pass});
        if (rcnt[r]) recvs.push_back({r, rp, es * (size_t)rcnt[r], pass});
      }
    }
    B200_TRY(b200_comm_exchange(h, c, sends, recvs));
  }
  tr(h, " T: bucket+exchange");
  // received triplets are ordered by ascending fine row (rank blocks ascend); stable sort by local coarse column
  b200_csr T = nullptr;
  B200_TRY(b200_csr_alloc(h, nc, P->global_rows, m, true, &T));
  B200_CUDA(cudaMemsetAsync(T->i, 0, sizeof(int) * ((size_t)nc + 1), h->stream));
  if (m) {
    sub_kernel<<<b200_grid(m, 256), 256, 0, h->stream>>>(m, r_col, cfirst);
    B200_LAUNCH_CHECK();
    hist_kernel<<<b200_grid(m, 256), 256, 0, h->stream>>>(m, r_col, T->i);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_exclusive_scan_inplace(h, T->i, (size_t)nc + 1));
    int *idx2 = nullptr, *keys2 = nullptr, *perm2 = nullptr;
    B200_TRY(b200_dalloc<int>(h, &idx2, m)); B200_TRY(b200_dalloc<int>(h, &keys2, m)); B200_TRY(b200_dalloc<int>(h, &perm2, m));
    iota_kernel2<<<b200_grid(m, 256), 256, 0, h->stream>>>(m, idx2);
    B200_LAUNCH_CHECK();
    int bits = 1;
    while ((1LL << bits) < (long long)nc) bits++;
    size_t tb = 0;
    B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, r_col, keys2, idx2, perm2, m, 0, bits, h->stream));
    char *tmp = nullptr;
    B200_TRY(b200_dalloc<char>(h, &tmp, tb));
    B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, r_col, keys2, idx2, perm2, m, 0, bits, h->stream));
    ++g_b200_launches;
    gather_kernel<int><<<b200_grid(m, 256), 256, 0, h->stream>>>(m, perm2, r_row, T->j);
    gather_kernel<double><<<b200_grid(m, 256), 256, 0, h->stream>>>(m, perm2, r_val, T->a);
    g_b200_launches += 2;
    B200_TRY(b200_dfree(h, tmp)); B200_TRY(b200_dfree(h, idx2)); B200_TRY(b200_dfree(h, keys2)); B200_TRY(b200_dfree(h, perm2));
  }
  for (void *q : {(void *)owner, (void *)rows, (void *)idx, (void *)owner_s, (void *)perm, (void *)d_starts, (void *)cntr, (void *)s_row,
                  (void *)s_col, (void *)s_val, (void *)r_row, (void *)r_col, (void *)r_val})
    B200_TRY(b200_dfree(h, q));
  *out = T;
  tr(h, " T: sort by column");
  return 0;
}

static b200_dist_matrix new_dist(int n, int first_row, int global_rows, const std::vector<int> &row_starts, int first_col,
                                 int n_owned_cols, int global_cols, const std::vector<int> &col_starts, b200_csr G) {
  b200_dist_matrix M = new b200_dist_matrix_s();
  M->n = n; M->first_row = first_row; M->global_rows = global_rows; M->row_starts = row_starts;
  M->first_col = first_col; M->n_owned_cols = n_owned_cols; M->global_cols = global_cols; M->col_starts = col_starts;
  M->G = G;
  return M;
}

namespace {
__global__ void fix_rowptr_kernel(int n, const int *__restrict__ src, int add, int *__restrict__ dst) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[k] + add;
}
__global__ void iota_skip_kernel(int ng, int first, int n, int *__restrict__ out) {      // every id except [first, first + n)
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < ng) out[k] = k < first ? k : k + n;
}
}  // namespace

// every rank receives the whole matrix (rows in global order, global column ids, entry order kept)
static int gather_csr(b200_handle h, b200_comm c, b200_dist_matrix M, b200_csr *out) {
  const int R = b200_comm_size(c);
  const int N = M->global_rows, n = M->n;
  const std::vector<int> &rs = M->row_starts;
  int my_nnz = M->G->nnz;
  std::vector<int> nz(R), nzs(R + 1, 0);
  B200_TRY(b200_comm_allgather_host(h, c, &my_nnz, sizeof(int), nz.data()));
  for (int r = 0; r < R; r++) nzs[r + 1] = nzs[r] + nz[r];
  b200_csr F = nullptr;
  B200_TRY(b200_csr_alloc(h, N, M->global_cols, nzs[R], true, &F));
  int *tmp = nullptr;
  B200_TRY(b200_dalloc<int>(h, &tmp, (size_t)N + R));
  std::vector<b200_xfer> s1, r1, s2, r2, s3, r3;
  for (int r = 0; r < R; r++) {
    const int nr = rs[r + 1] - rs[r];
    s1.push_back({r, M->G->i, sizeof(int) * ((size_t)n + 1), 0});
    r1.push_back({r, tmp + rs[r] + r, sizeof(int) * ((size_t)nr + 1), 0});
    if (my_nnz) { s2.push_back({r, M->G->j, sizeof(int) * (size_t)my_nnz, 0}); s3.push_back({r, M->G->a, sizeof(double) * (size_t)my_nnz, 0}); }
    if (nz[r]) { r2.push_back({r, F->j + nzs[r], sizeof(int) * (size_t)nz[r], 0}); r3.push_back({r, F->a + nzs[r], sizeof(double) * (size_t)nz[r], 0}); }
  }
  B200_TRY(b200_comm_exchange(h, c, s1, r1));
  B200_TRY(b200_comm_exchange(h, c, s2, r2));
  B200_TRY(b200_comm_exchange(h, c, s3, r3));
  for (int r = 0; r < R; r++) {
    const int nr = rs[r + 1] - rs[r] + (r == R - 1 ? 1 : 0);      // the last block also carries the closing pointer
    if (nr) {
      fix_rowptr_kernel<<<b200_grid(nr, 256), 256, 0, h->stream>>>(nr, tmp + rs[r] + r, nzs[r], F->i + rs[r]);
      B200_LAUNCH_CHECK();
    }
  }
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, tmp));
  *out = F;
  return 0;
}

// levels >= `level` become one single-GPU hierarchy on every rank (see b200_dist_amg_s::tail)
static int build_tail(b200_handle h, b200_comm c, b200_amg prm, b200_dist_amg amg, int level, int levels_left) {
  dist_level &L = amg->lv[level];
  b200_dist_matrix A = L.A;
  b200_csr full = nullptr;
  tr(h, nullptr);
  B200_TRY(gather_csr(h, c, A, &full));
  tr(h, " tail: gather_csr");
  b200_parcsr TA = new b200_parcsr_s();
  TA->global_rows = TA->global_cols = A->global_rows;
  TA->diag = full;
  B200_TRY(b200_csr_build_plan(h, full));
  B200_TRY(b200_csr_alloc(h, A->global_rows, 0, 0, true, &TA->offd));
  B200_CUDA(cudaMemsetAsync(TA->offd->i, 0, sizeof(int) * ((size_t)A->global_rows + 1), h->stream));
  amg->tailA = TA;
  B200_TRY(b200_amg_create(&amg->tail));
  B200_TRY(b200_amg_clone_params(prm, amg->tail));
  B200_TRY(b200_amg_set_int(amg->tail, "MaxLevels", levels_left));
  B200_TRY(b200_amg_set_int(amg->tail, "AggNumLevels", 0));
  B200_TRY(b200_amg_set_int(amg->tail, "SeqThreshold", 0));
  B200_TRY(b200_amg_set_int(amg->tail, "CoarsenType", 8));
  B200_TRY(b200_amg_set_int(amg->tail, "MaxIter", 1));
  B200_TRY(b200_amg_set_real(amg->tail, "Tol", 0.0));
  tr(h, " tail: plan+params");
  B200_TRY(b200_amg_setup(h, amg->tail, TA));
  tr(h, " tail: amg_setup");
  // gather plan for the right-hand side: every id this rank does not own is a "ghost"
  const int N = A->global_rows, n = A->n, first = A->first_row, ng = N - n;
  int *gid = nullptr;
  B200_TRY(b200_dalloc<int>(h, &gid, (size_t)ng + 1));
  if (ng) {
    iota_skip_kernel<<<b200_grid(ng, 256), 256, 0, h->stream>>>(ng, first, n, gid);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_halo_build(h, c, A->row_starts, gid, ng, &amg->tail_plan));
  B200_TRY(b200_dalloc<double>(h, &amg->tail_F, (size_t)N + 8)); B200_TRY(b200_dalloc<double>(h, &amg->tail_U, (size_t)N + 8));
  B200_TRY(b200_dalloc<double>(h, &amg->tail_G, (size_t)ng + 8));
  B200_CUDA(cudaMemsetAsync(amg->tail_U, 0, sizeof(double) * ((size_t)N + 8), h->stream));
  amg->tail_N = N; amg->tail_first = first; amg->tail_rank = b200_comm_rank(c);
  tr(h, " tail: rhs plan");
  return 0;
}

extern "C" int b200_dist_amg_setup(b200_handle h, b200_comm c, b200_amg prm, b200_dist_matrix A0, b200_dist_amg *out) {
  if (!prm || !A0) B200_FAIL("dist_amg_setup: null argument");
  if (b200_amg_get_int(prm, "CoarsenType") != 8 && b200_amg_get_int(prm, "CoarsenType") != 9)
    B200_FAIL("only PMIS (CoarsenType 8; measures are drawn as for 9 = partition independent) is implemented");
  const int rdown = b200_amg_get_int(prm, "RelaxType");
  const int rup = b200_amg_get_int(prm, "RelaxTypeUp") >= 0 ? b200_amg_get_int(prm, "RelaxTypeUp") : rdown;
  auto is_l1gs = [](int t) { return t == 8 || t == 13 || t == 14; };
  const bool jac = (rdown == 18 || rdown == 7) && rup == rdown;
  if (!(jac || (is_l1gs(rdown) && is_l1gs(rup))))
    B200_FAIL("multi-GPU RelaxType: 18 (l1-Jacobi), 7 (Jacobi) or the l1 hybrid Gauss-Seidel family 8/13/14");
  if (!jac && b200_amg_get_real(prm, "RelaxWt") != 1.0) B200_FAIL("Gauss-Seidel smoothers: only relax_weight 1 is implemented");
  if (b200_amg_get_int(prm, "InterpType") != 6 ||
      b200_amg_get_int(prm, "RelaxOrder") != 0 || b200_amg_get_int(prm, "AggNumLevels") < 0 ||
      b200_amg_get_int(prm, "NumSweeps") != 1 || b200_amg_get_int(prm, "CycleType") != 1 || b200_amg_get_int(prm, "FCycle") != 0 ||
      (b200_amg_get_int(prm, "NumSweepsDown") != -1 && b200_amg_get_int(prm, "NumSweepsDown") != 1) ||
      (b200_amg_get_int(prm, "NumSweepsUp") != -1 && b200_amg_get_int(prm, "NumSweepsUp") != 1) || b200_amg_get_int(prm, "NumSweepsCoarse") != 1 ||
      b200_amg_get_int(prm, "RAP2") != 0 || (b200_amg_get_int(prm, "ModuleRAP2") != 0 && b200_amg_get_int(prm, "ModuleRAP2") != 1))
    B200_FAIL("unsupported BoomerAMG configuration on the B200 path (see b200_amg_setup)");
  const int R = b200_comm_size(c), me = b200_comm_rank(c);
  const double theta = b200_amg_get_real(prm, "StrongThreshold"), mrs = b200_amg_get_real(prm, "MaxRowSum");
  const double trunc = b200_amg_get_real(prm, "TruncFactor");
  const int pmax = b200_amg_get_int(prm, "PMaxElmts"), max_levels = b200_amg_get_int(prm, "MaxLevels");
  const int max_coarse = b200_amg_get_int(prm, "MaxCoarseSize"), seed = b200_amg_get_int(prm, "Seed");
  const int seq_th = b200_amg_get_int(prm, "SeqThreshold");
  const int mod_rap2 = b200_amg_get_int(prm, "ModuleRAP2"), agg_nl = b200_amg_get_int(prm, "AggNumLevels");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, h->stream);
  // B200_TRACE=1: wall-clock (stream-synchronised) time of every setup phase, per level, on stderr
  const bool trace = getenv("B200_TRACE") != nullptr;
  auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tlast = 0;
  int tlevel = 0;
  auto mark = [&](const char *what) {
    if (!trace) return;
    cudaStreamSynchronize(h->stream);
    const double t = tnow();
    if (what) fprintf(stderr, "[b200 trace] rank %d level %d %-18s %8.3f ms\n", b200_comm_rank(c), tlevel, what, t - tlast);
    tlast = t;
  };
  mark(nullptr);
  const long long host_ops0 = b200_comm_host_ops(c);
  b200_dist_amg amg = new b200_dist_amg_s();
  amg->relax_wt = b200_amg_get_real(prm, "RelaxWt");
  amg->gs = !jac; amg->relax_down = rdown; amg->relax_up = rup;
  amg->gs_blocks = b200_amg_get_int(prm, "GSBlocks");
  dist_level L0;
  L0.A = A0; L0.n = A0->n;
  amg->lv.push_back(L0);
  int level = 0;
  bool not_finished = max_levels > 1;
  while (not_finished) {
    dist_level &L = amg->lv[level];
    b200_dist_matrix A = L.A;
    const int n = A->n, ng = A->halo->ng;
    const long long fine_size = A->global_rows;
    tlevel = level;
    if (R > 1 && jac && seq_th > 0 && level >= 1 && level >= b200_amg_get_int(prm, "AggNumLevels") && A->global_rows <= seq_th &&
        A->global_rows > max_coarse) {
      B200_TRY(build_tail(h, c, prm, amg, level, max_levels - level));
      mark("replicated tail");
      break;
    }
    // --- strength + PMIS on the localized operator (par_amg_setup.c:1035,:1114) -----------------
    b200_csr S = nullptr;
    B200_TRY(b200_strength(h, A->L, theta, mrs, &S));
    mark("strength");
    int *cf = nullptr;                              // [n + ng]
    B200_TRY(b200_dalloc<int>(h, &cf, (size_t)n + ng + 1));
    B200_TRY(b200_pmis_dist(h, c, S, A->halo, seed, A->first_row, cf));
    mark("pmis");
    const bool aggressive = level < agg_nl;
    // --- coarse numbering (par_coarse_parms.c:83-122): global coarse ids over [owned | ghost], -1 for F points ----
    auto number_coarse = [&](const int *cfv, int **f2c_out, int *nc_out, std::vector<int> *starts) -> int {
      int *f = nullptr;
      B200_TRY(b200_dalloc<int>(h, &f, (size_t)n + ng + 1));
      cflag2_kernel<<<b200_grid((size_t)n + 1, 256), 256, 0, h->stream>>>(n, cfv, f);
      B200_LAUNCH_CHECK();
      B200_TRY(b200_exclusive_scan_inplace(h, f, (size_t)n + 1));
      B200_CUDA(cudaMemcpyAsync(nc_out, f + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      B200_CUDA(cudaStreamSynchronize(h->stream));
      B200_TRY(gather_starts(h, c, *nc_out, starts));
      if (n) {
        f2c_global_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, cfv, (*starts)[me], f);
        B200_LAUNCH_CHECK();
      }
      B200_TRY(b200_halo_forward_i32(h, c, A->halo, f, f + n));
      *f2c_out = f;
      return 0;
    };
    // --- extended operators over [owned | U] (hypre_exchange_interp_data, aux_interp.c:552-660): needed by the
    //     ext+i weights and by the distance-two strength graph of aggressive coarsening -------------------------
    b200_csr Abig2 = nullptr, Sbig2 = nullptr;
    b200_halo_s *plan2 = nullptr;
    int nring = 0;
    auto build_extended = [&]() -> int {
    // rows of A and S for the ghost nodes, then the second ring of ghost ids they mention
    tr(h, nullptr);
    b200_csr Sg = nullptr;                          // S with global column ids (to serve fetches)
    B200_TRY(b200_csr_alloc(h, n, A->global_cols, S->nnz, false, &Sg));
    B200_CUDA(cudaMemcpyAsync(Sg->i, S->i, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
    if (S->nnz) {
      globalize_kernel<<<b200_grid(S->nnz, 256), 256, 0, h->stream>>>(S->nnz, S->j, A->first_col, A->n_owned_cols,
                                                                    A->halo->d_ghost_gid, Sg->j);
      B200_LAUNCH_CHECK();
    }
    b200_csr Aext = nullptr, Sext = nullptr;
    tr(h, " ext: globalize S");
    B200_TRY(fetch_rows(h, c, A->halo, A->G, &Aext));
    B200_TRY(fetch_rows(h, c, A->halo, Sg, &Sext));
    B200_TRY(b200_csr_destroy(h, Sg));
    // second ring: ids in the fetched rows that are neither owned nor first-ring ghosts.  Build the
    // sorted union U = ghosts1 + ring2 and a plan for it; extended index space = [owned | U].
    int *ring = nullptr;
    {
      int *cand = nullptr;
      const int tot = Aext->nnz + A->halo->ng;
      B200_TRY(b200_dalloc<int>(h, &cand, tot));
      if (Aext->nnz) B200_CUDA(cudaMemcpyAsync(cand, Aext->j, sizeof(int) * (size_t)Aext->nnz, cudaMemcpyDeviceToDevice, h->stream));
      if (A->halo->ng)
        B200_CUDA(cudaMemcpyAsync(cand + Aext->nnz, A->halo->d_ghost_gid, sizeof(int) * (size_t)A->halo->ng, cudaMemcpyDeviceToDevice, h->stream));
      B200_TRY(ghost_columns(h, cand, tot, A->first_col, A->n_owned_cols, &ring, &nring));
      B200_TRY(b200_dfree(h, cand));
      if (!ring) B200_TRY(b200_dalloc<int>(h, &ring, 1));
    }
    tr(h, " ext: ring ids");
    B200_TRY(b200_halo_build(h, c, A->col_starts, ring, nring, &plan2));
    tr(h, " ext: halo_build plan2");
    // Extended operators: rows [owned | U], columns [owned | U].  The row kernels address neighbour
    // ROWS by column id, so every id of U needs a row slot; only the first ring has entries (second-ring
    // rows are never dereferenced: the kernels visit rows of strong neighbours of owned rows only).
    // ghosts1 and U are both sorted by global id and ghosts1 is a subset of U, so the fetched entries
    // are already in U order: only the row pointer has to be spread out.
    int *pos = nullptr;                              // position (in [owned | U]) of every first-ring ghost
    B200_TRY(b200_dalloc<int>(h, &pos, (size_t)ng + 1));
    if (ng) {
      localize_kernel<<<b200_grid(ng, 256), 256, 0, h->stream>>>(ng, A->halo->d_ghost_gid, A->first_col, A->n_owned_cols, nring,
                                                               plan2->d_ghost_gid, pos);
      B200_LAUNCH_CHECK();
    }
    // Only the fetched ghost rows carry global ids: they are localized against the new plan; the rank's own rows are taken
    // from the localized operator (and from S, which was built on it) with their first-ring ghost columns renumbered
    // through pos[] -- one pass over A and one over S instead of stacking the global forms and searching every column again.
    b200_csr AextU = nullptr, SextU = nullptr, AextL = nullptr, SextL = nullptr;
    B200_TRY(spread_rows(h, Aext, pos, A->n_owned_cols, nring, &AextU));
    B200_TRY(spread_rows(h, Sext, pos, A->n_owned_cols, nring, &SextU));
    B200_TRY(b200_csr_destroy(h, Aext)); B200_TRY(b200_csr_destroy(h, Sext));
    tr(h, " ext: spread");
    B200_TRY(localize_copy(h, AextU, plan2, A->first_col, A->n_owned_cols, false, &AextL));
    B200_TRY(localize_copy(h, SextU, plan2, A->first_col, A->n_owned_cols, false, &SextL));
    B200_TRY(b200_csr_destroy(h, AextU)); B200_TRY(b200_csr_destroy(h, SextU));
    tr(h, " ext: localize ghost rows");
    B200_TRY(stack_rows_remap(h, A->L, A->n_owned_cols, pos, AextL, A->n_owned_cols + nring, &Abig2));
    tr(h, " ext: stack A");
    B200_TRY(stack_rows_remap(h, S, A->n_owned_cols, pos, SextL, A->n_owned_cols + nring, &Sbig2));
    tr(h, " ext: stack S");
    B200_TRY(b200_csr_destroy(h, AextL)); B200_TRY(b200_csr_destroy(h, SextL));
    B200_TRY(b200_dfree(h, pos));
      return 0;
    };
    auto extend_markers = [&](const int *cfv, const int *f2cv, int **cf_big_out, int **f2c_big_out) -> int {
      int *cf_big = nullptr, *f2c_big = nullptr;                 // cf / f2c over [owned | U]
      B200_TRY(b200_dalloc<int>(h, &cf_big, (size_t)n + nring + 1));
      B200_TRY(b200_dalloc<int>(h, &f2c_big, (size_t)n + nring + 1));
      B200_CUDA(cudaMemcpyAsync(cf_big, cfv, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
      B200_CUDA(cudaMemcpyAsync(f2c_big, f2cv, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
      B200_TRY(b200_halo_forward_i32(h, c, plan2, cf_big, cf_big + n));
      B200_TRY(b200_halo_forward_i32(h, c, plan2, f2c_big, f2c_big + n));
      *cf_big_out = cf_big; *f2c_big_out = f2c_big;
      return 0;
    };
    B200_TRY(build_extended());
    mark("extended A,S");
    if (aggressive) {
      // second coarsening on the distance-two strength graph of the C points (par_amg_setup.c:1239-1256), then
      // hypre_BoomerAMGCorrectCFMarker (:1592).  S2's columns carry the global ids of the FIRST coarse numbering.
      int *f2c1 = nullptr, nc1 = 0;
      std::vector<int> cs1;
      B200_TRY(number_coarse(cf, &f2c1, &nc1, &cs1));
      if (cs1[R] > 0) {
        int *cfb = nullptr, *f2cb = nullptr;
        B200_TRY(extend_markers(cf, f2c1, &cfb, &f2cb));
        b200_csr S2g = nullptr;
        B200_TRY(b200_create_2nd_s_ex(h, Sbig2, n, cfb, f2cb, cs1[me], cs1[R], &S2g));
        B200_TRY(b200_dfree(h, cfb)); B200_TRY(b200_dfree(h, f2cb));
        b200_dist_matrix S2d = new_dist(nc1, cs1[me], cs1[R], cs1, cs1[me], nc1, cs1[R], cs1, S2g);
        B200_TRY(dist_localize(h, c, S2d));
        int *cfn = nullptr;
        B200_TRY(b200_dalloc<int>(h, &cfn, (size_t)nc1 + S2d->halo->ng + 1));
        B200_TRY(b200_pmis_dist_init(h, c, S2d->L, S2d->halo, seed, cs1[me], 3, cfn));
        B200_TRY(b200_correct_cf(h, n, cfn, cf));
        B200_TRY(b200_halo_forward_i32(h, c, A->halo, cf, cf + n));
        B200_TRY(b200_dfree(h, cfn));
        B200_TRY(b200_dist_matrix_destroy(h, S2d));
      }
      B200_TRY(b200_dfree(h, f2c1));
    }
    int *f2c = nullptr, nc = 0;                     // [n + ng]: global coarse id or -1
    std::vector<int> cstarts;
    B200_TRY(number_coarse(cf, &f2c, &nc, &cstarts));
    mark(aggressive ? "2nd pass + numbering" : "coarse numbering");
    const long long coarse_size = cstarts[R];
    if (coarse_size == 0 || coarse_size == fine_size) {       // par_amg_setup.c:1487-1525
      B200_TRY(b200_csr_destroy(h, S)); B200_TRY(b200_dfree(h, cf)); B200_TRY(b200_dfree(h, f2c));
      B200_TRY(b200_csr_destroy(h, Abig2)); B200_TRY(b200_csr_destroy(h, Sbig2));
      b200_halo_free(h, plan2);
      break;
    }
    b200_csr Pg = nullptr;
    if (aggressive) {
      // hypre_BoomerAMGBuildMultipass (:1601): pass numbers, pass rows and their ghost copies move over the halo of A
      b200_agg_hooks hooks;
      hooks.sync_int = [&](int *v) -> int { return b200_halo_forward_i32(h, c, A->halo, v, v + n); };
      hooks.sum_int = [&](int *v) -> int {
        long long t = *v;
        B200_TRY(b200_comm_allreduce_sum_ll(h, c, &t, 1));
        *v = t > 0x7fffffffLL ? 0x7fffffff : (int)t;
        return 0;
      };
      hooks.with_ghost_rows = [&](b200_csr rows, b200_csr *big) -> int {
        b200_csr ext = nullptr;
        B200_TRY(fetch_rows(h, c, A->halo, rows, &ext));
        B200_TRY(stack_rows(h, rows, ext, big));
        B200_TRY(b200_csr_destroy(h, ext));
        return 0;
      };
      B200_TRY(b200_multipass_ex(h, A->L, S, n, n + ng, cf, f2c, (int)coarse_size, &hooks, &Pg));
    } else {
      // --- ext+i interpolation (par_lr_interp.c:1040-1925 with hypre_exchange_interp_data) ---------
      int *cf_big = nullptr, *f2c_big = nullptr;
      B200_TRY(extend_markers(cf, f2c, &cf_big, &f2c_big));
      B200_TRY(b200_extpi_interp_ex(h, Abig2, Sbig2, cf_big, n, f2c_big, (int)coarse_size, trunc, pmax, &Pg));
      B200_TRY(b200_dfree(h, cf_big)); B200_TRY(b200_dfree(h, f2c_big));
    }
    mark("interpolation");
    B200_TRY(b200_csr_destroy(h, Abig2)); B200_TRY(b200_csr_destroy(h, Sbig2));
    b200_halo_free(h, plan2);
    B200_TRY(b200_csr_destroy(h, S));
    if (n) {
      fix_cf2_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, cf);     // par_lr_interp.c:1888-1894
      B200_LAUNCH_CHECK();
    }
    L.cf = cf;
    B200_TRY(b200_dfree(h, f2c));
    L.P = new_dist(n, A->first_row, A->global_rows, A->row_starts, cstarts[me], nc, (int)coarse_size, cstarts, Pg);
    // --- Galerkin product (par_csr_triplemat.c:606-871): Q = A*P with ghost rows of P, C = P^T * Q ---
    b200_csr Rg = nullptr;
    B200_TRY(dist_transpose(h, c, L.P, cstarts, &Rg));
    mark("transpose");
    L.R = new_dist(nc, cstarts[me], (int)coarse_size, cstarts, A->first_row, n, A->global_rows, A->row_starts, Rg);
    b200_csr AHg = nullptr;
    if (mod_rap2) {
      // hypre_ParCSRMatrixRAPKT: Q = A*P with ghost rows of P, C = P^T * Q
      b200_csr Pext = nullptr, Pbig = nullptr, Qg = nullptr;
      B200_TRY(fetch_rows(h, c, A->halo, Pg, &Pext));             // hypre_ParCSRMatrixExtractBExt(P, A)
      B200_TRY(stack_rows(h, Pg, Pext, &Pbig));
      B200_TRY(b200_csr_destroy(h, Pext));
      B200_TRY(b200_csr_multiply_ex(h, A->L, Pbig, 0, 0, (int)coarse_size, &Qg));
      B200_TRY(b200_csr_destroy(h, Pbig));
      B200_TRY(dist_localize(h, c, L.R, true, true));             // ghosts of R = remote fine rows; nobody fetches rows of R
      b200_csr Qext = nullptr, Qbig = nullptr;
      B200_TRY(fetch_rows(h, c, L.R->halo, Qg, &Qext));
      B200_TRY(stack_rows(h, Qg, Qext, &Qbig));
      B200_TRY(b200_csr_destroy(h, Qext)); B200_TRY(b200_csr_destroy(h, Qg));
      B200_TRY(b200_csr_multiply_ex(h, L.R->L, Qbig, 1, cstarts[me], (int)coarse_size, &AHg));
      B200_TRY(b200_csr_destroy(h, Qbig));
    } else {
      // the library default, hypre_BoomerAMGBuildCoarseOperatorKT (par_rap.c): row ic of R*A first, then times P with
      // the diagonal entry created first -- (R*A)*P, rows of A and P fetched from their owners in global entry order
      B200_TRY(dist_localize(h, c, L.R, true, true));
      b200_csr Aext = nullptr, Abig = nullptr, RAg = nullptr;
      B200_TRY(fetch_rows(h, c, L.R->halo, A->G, &Aext));
      B200_TRY(stack_rows(h, A->G, Aext, &Abig));
      B200_TRY(b200_csr_destroy(h, Aext));
      mark("RA: fetch+stack");
      B200_TRY(b200_csr_multiply_ex(h, L.R->L, Abig, 0, 0, A->global_cols, &RAg));
      mark("RA: multiply");
      B200_TRY(b200_csr_destroy(h, Abig));
      b200_dist_matrix RAd = new_dist(nc, cstarts[me], (int)coarse_size, cstarts, A->first_col, A->n_owned_cols, A->global_cols,
                                      A->col_starts, RAg);
      B200_TRY(dist_localize(h, c, RAd, false, true));          // operand of the next product only: no plan, in place
      b200_csr Pext = nullptr, Pbig = nullptr;
      B200_TRY(fetch_rows(h, c, RAd->halo, Pg, &Pext));
      B200_TRY(stack_rows(h, Pg, Pext, &Pbig));
      B200_TRY(b200_csr_destroy(h, Pext));
      mark("(RA)P: localize+fetch");
      B200_TRY(b200_csr_multiply_ex(h, RAd->L, Pbig, 1, cstarts[me], (int)coarse_size, &AHg));
      mark("(RA)P: multiply");
      B200_TRY(b200_csr_destroy(h, Pbig));
      B200_TRY(b200_dist_matrix_destroy(h, RAd));
    }
    dist_level Ln;
    Ln.A = new_dist(nc, cstarts[me], (int)coarse_size, cstarts, cstarts[me], nc, (int)coarse_size, cstarts, AHg);
    B200_TRY(dist_localize(h, c, Ln.A));
    Ln.n = nc;
    B200_TRY(dist_localize(h, c, L.P));
    mark("localize A_H, P");
    amg->lv.push_back(Ln);
    ++level;
    if (level == max_levels - 1 || coarse_size <= max_coarse) not_finished = false;
    if (not_finished && (double)coarse_size >= 0.75 * (double)fine_size)
      B200_FAIL("coarsening stalled (coarse >= 0.75 fine): the reference switches to CLJP here, which is out of scope");
  }
  const int nl = (int)amg->lv.size();
  // column-sorted solve copies of the coarse Galerkin operators (Jacobi-type smoothers only: the Gauss-Seidel kernels need
  // the diagonal entry first)
  {
    static const bool no_sort = [] { const char *e = getenv("B200_NO_SORTED_COPY"); return e && e[0] == '1'; }();
    for (int l = 1; l < nl - 1 && !amg->gs && !no_sort; l++) {
      b200_dist_matrix M = amg->lv[l].A;
      if ((double)M->L->nnz > 12.0 * M->n && !M->Ls) {
        B200_TRY(b200_csr_sorted_copy(h, M->L, &M->Ls));
        B200_TRY(b200_csr_build_plan(h, M->Ls));
      }
    }
    mark("sorted solve copies");
  }
  // vectors: capacity = owned + the largest ghost set any operator reads them with
  for (int l = 0; l < nl; l++) {
    dist_level &L = amg->lv[l];
    int cap = L.n + L.A->halo->ng;
    if (L.R) cap = std::max(cap, L.n + L.R->halo->ng);                 // R reads fine vectors (Vtemp)
    if (l > 0) cap = std::max(cap, L.n + amg->lv[l - 1].P->halo->ng);   // P of the finer level reads this level's U
    L.cap = cap + 8;
    B200_TRY(b200_dalloc<double>(h, &L.F, L.cap)); B200_TRY(b200_dalloc<double>(h, &L.U, L.cap)); B200_TRY(b200_dalloc<double>(h, &L.T, L.cap));
    B200_CUDA(cudaMemsetAsync(L.F, 0, sizeof(double) * L.cap, h->stream));
    B200_CUDA(cudaMemsetAsync(L.U, 0, sizeof(double) * L.cap, h->stream));
    B200_CUDA(cudaMemsetAsync(L.T, 0, sizeof(double) * L.cap, h->stream));
    B200_TRY(b200_dalloc<double>(h, &L.l1, L.n));
    // offd entries are part of the merged row (ams.c:651-657); option 4 counts them -- and the entries
    // outside the Gauss-Seidel block of the row -- with weight 1/2 (ams.c:3560-3625)
    B200_TRY(b200_l1_norms_blocks(h, L.A->L, amg->gs ? 4 : (amg->relax_down == 7 ? 5 : 1), amg->gs ? amg->gs_blocks : 1, L.l1));
    if (amg->gs && (l < nl - 1 || L.A->global_rows > max_coarse)) {
      if (L.A->L->gs && b200_gs_plan_blocks(L.A->L->gs) != amg->gs_blocks) { B200_TRY(b200_gs_plan_destroy(h, L.A->L->gs)); L.A->L->gs = nullptr; }
      if (!L.A->L->gs) B200_TRY(b200_gs_plan_create(h, L.A->L, amg->gs_blocks, &L.A->L->gs));
    }
    amg->vtemp_cap = std::max(amg->vtemp_cap, L.cap);
    // setup-form copies are no longer needed below level 0 (level 0's belongs to the caller)
  }
  tlevel = nl;
  mark("vectors + l1 norms");
  B200_TRY(b200_dalloc<double>(h, &amg->Vtemp, amg->vtemp_cap));
  B200_CUDA(cudaMemsetAsync(amg->Vtemp, 0, sizeof(double) * amg->vtemp_cap, h->stream));
  // coarsest level: gather the dense matrix on every rank (par_gauss_elim.c:78-118)
  {
    dist_level &Lc = amg->lv[nl - 1];
    const int ncg = Lc.A->global_rows;
    amg->coarse_ge = ncg <= max_coarse && ncg > 0 && !amg->tail;
    if (amg->coarse_ge) {
      amg->ge_n = ncg;
      amg->ge_starts = Lc.A->row_starts;
      double *loc = nullptr;
      B200_TRY(b200_dalloc<double>(h, &loc, (size_t)Lc.n * ncg + 1));
      B200_TRY(b200_dalloc<double>(h, &amg->ge_A, (size_t)2 * ncg * ncg + 1));
      B200_TRY(b200_dalloc<double>(h, &amg->ge_f, (size_t)3 * ncg + 1));
      if (Lc.n) {
        dense_rows_kernel<<<b200_grid(Lc.n, 64), 64, 0, h->stream>>>(Lc.n, ncg, Lc.A->G->i, Lc.A->G->j, Lc.A->G->a, loc);
        B200_LAUNCH_CHECK();
      }
      std::vector<b200_xfer> sends, recvs;
      for (int r = 0; r < R; r++) {
        const int rn = amg->ge_starts[r + 1] - amg->ge_starts[r];
        if (Lc.n) sends.push_back({r, loc, sizeof(double) * (size_t)Lc.n * ncg, 0});
        if (rn) recvs.push_back({r, amg->ge_A + (size_t)amg->ge_starts[r] * ncg, sizeof(double) * (size_t)rn * ncg, 0});
      }
      B200_TRY(b200_comm_exchange(h, c, sends, recvs));
      B200_CUDA(cudaStreamSynchronize(h->stream));
      B200_TRY(b200_dfree(h, loc));
    }
  }
  mark("coarse gather");
  if (trace) fprintf(stderr, "[b200 trace] rank %d setup used %lld host-synchronised collectives (exchanges + gathers)\n", b200_comm_rank(c),
                     b200_comm_host_ops(c) - host_ops0);
  cudaEventRecord(e1, h->stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  amg->setup_ms = ms;
  *out = amg;
  return 0;
}

extern "C" int b200_dist_amg_destroy(b200_handle h, b200_dist_amg amg) {
  if (!amg) return 0;
  for (size_t l = 0; l < amg->lv.size(); l++) {
    dist_level &L = amg->lv[l];
    if (l > 0) B200_TRY(b200_dist_matrix_destroy(h, L.A));
    B200_TRY(b200_dist_matrix_destroy(h, L.P));
    B200_TRY(b200_dist_matrix_destroy(h, L.R));
    B200_TRY(b200_dfree(h, L.cf)); B200_TRY(b200_dfree(h, L.l1));
    B200_TRY(b200_dfree(h, L.F)); B200_TRY(b200_dfree(h, L.U)); B200_TRY(b200_dfree(h, L.T));
  }
  B200_TRY(b200_dfree(h, amg->Vtemp)); B200_TRY(b200_dfree(h, amg->ge_A)); B200_TRY(b200_dfree(h, amg->ge_f));
  for (b200_dist_matrix v : amg->views) B200_TRY(b200_dist_matrix_destroy(h, v));
  if (amg->tail) {
    B200_TRY(b200_amg_destroy(h, amg->tail));
    B200_TRY(b200_parcsr_destroy(h, amg->tailA));
    b200_halo_free(h, amg->tail_plan);
    B200_TRY(b200_dfree(h, amg->tail_F)); B200_TRY(b200_dfree(h, amg->tail_U)); B200_TRY(b200_dfree(h, amg->tail_G));
  }
  delete amg;
  return 0;
}
b200_csr b200_amg_level_A(b200_amg amg, int l);
b200_csr b200_amg_level_P(b200_amg amg, int l);
const int *b200_amg_level_CF(b200_amg amg, int l);
int b200_amg_num_levels(b200_amg amg);

// Levels of the replicated tail seen through the row-partitioned accessors: the tail's first level is partitioned like the
// distributed level it was gathered from; deeper levels are reported whole by rank 0 and empty by the other ranks (any
// contiguous partition describes a replicated matrix).
static int tail_rows(b200_dist_amg amg, int tl, int me, int *r0, int *r1) {
  const int n = b200_amg_level_A(amg->tail, tl)->nrows;
  if (tl == 0) { *r0 = amg->tail_first; *r1 = amg->tail_first + amg->lv.back().n; }
  else { *r0 = 0; *r1 = (me == 0) ? n : 0; }
  return n;
}
static int tail_view(b200_handle h, b200_dist_amg amg, b200_csr src, int r0, int r1, b200_dist_matrix *out) {
  const int n = r1 - r0;
  int e0 = 0, e1 = 0;
  if (n > 0) {
    B200_CUDA(cudaMemcpyAsync(&e0, src->i + r0, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaMemcpyAsync(&e1, src->i + r1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  b200_csr G = nullptr;
  B200_TRY(b200_csr_alloc(h, n, src->ncols, e1 - e0, true, &G));
  if (n > 0) {
    fix_rowptr_kernel<<<b200_grid(n + 1, 256), 256, 0, h->stream>>>(n + 1, src->i + r0, -e0, G->i);
    B200_LAUNCH_CHECK();
    if (e1 > e0) {
      B200_CUDA(cudaMemcpyAsync(G->j, src->j + e0, sizeof(int) * (size_t)(e1 - e0), cudaMemcpyDeviceToDevice, h->stream));
      B200_CUDA(cudaMemcpyAsync(G->a, src->a + e0, sizeof(double) * (size_t)(e1 - e0), cudaMemcpyDeviceToDevice, h->stream));
    }
  } else {
    B200_CUDA(cudaMemsetAsync(G->i, 0, sizeof(int), h->stream));
  }
  b200_dist_matrix M = new b200_dist_matrix_s();
  M->n = n; M->first_row = r0; M->global_rows = src->nrows; M->global_cols = src->ncols; M->G = G;
  amg->views.push_back(M);
  *out = M;
  return 0;
}
extern "C" int b200_dist_amg_num_levels(b200_dist_amg amg) {
  if (!amg) return 0;
  return (int)amg->lv.size() + (amg->tail ? b200_amg_num_levels(amg->tail) - 1 : 0);
}
extern "C" b200_dist_matrix b200_dist_amg_level_A(b200_dist_amg amg, int l) {
  if (!amg || l < 0) return nullptr;
  if (l < (int)amg->lv.size()) return amg->lv[l].A;
  return nullptr;                     // tail levels: b200_dist_amg_level_view
}
extern "C" b200_dist_matrix b200_dist_amg_level_P(b200_dist_amg amg, int l) {
  return (amg && l >= 0 && l < (int)amg->lv.size()) ? amg->lv[l].P : nullptr;     // tail levels: b200_dist_amg_level_view
}
// what = 0: A_l, 1: P_l, for any level including the replicated tail (views are owned by the hierarchy)
extern "C" int b200_dist_amg_level_view(b200_handle h, b200_comm c, b200_dist_amg amg, int l, int what, b200_dist_matrix *out) {
  if (!amg || !out) B200_FAIL("level_view: null argument");
  *out = nullptr;
  const int nd = (int)amg->lv.size();
  if (l >= 0 && l < nd && what == 0) { *out = amg->lv[l].A; return 0; }
  if (l >= 0 && l < nd && what == 1 && amg->lv[l].P) { *out = amg->lv[l].P; return 0; }
  if (!amg->tail) B200_FAIL("level_view: no such level");
  const int tl = l - (nd - 1);
  if (tl < 0 || tl >= b200_amg_num_levels(amg->tail)) B200_FAIL("level_view: no such level");
  b200_csr src = what == 0 ? b200_amg_level_A(amg->tail, tl) : b200_amg_level_P(amg->tail, tl);
  if (!src) B200_FAIL("level_view: the coarsest level has no interpolation");
  int r0 = 0, r1 = 0;
  tail_rows(amg, tl, b200_comm_rank(c), &r0, &r1);
  return tail_view(h, amg, src, r0, r1, out);
}
extern "C" int b200_dist_amg_level_cf(b200_handle h, b200_dist_amg amg, int l, int *h_cf) {
  if (!amg || l < 0) B200_FAIL("no CF marker on this level");
  const int nd = (int)amg->lv.size();
  if (l < nd && amg->lv[l].cf) {
    B200_CUDA(cudaMemcpyAsync(h_cf, amg->lv[l].cf, sizeof(int) * (size_t)amg->lv[l].n, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
  }
  if (!amg->tail) B200_FAIL("no CF marker on this level");
  const int tl = l - (nd - 1);
  const int *cf = b200_amg_level_CF(amg->tail, tl);
  if (tl < 0 || !cf) B200_FAIL("no CF marker on this level");
  int r0 = 0, r1 = 0;
  tail_rows(amg, tl, amg->tail_rank, &r0, &r1);
  if (r1 > r0) {
    B200_CUDA(cudaMemcpyAsync(h_cf, cf + r0, sizeof(int) * (size_t)(r1 - r0), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  return 0;
}
extern "C" int b200_dist_amg_setup_ms(b200_dist_amg amg, double *ms) { if (!amg) B200_FAIL("null"); *ms = amg->setup_ms; return 0; }

__global__ void ge_scatter_kernel(int ncg, int first, int n, const double *__restrict__ f, double *__restrict__ out) {
  const int j = threadIdx.x;
  if (j < ncg) out[j] = (j >= first && j < first + n) ? f[j - first] : 0.0;
}

// one V(1,1) cycle across ranks (par_cycle.c:255-622); u is zero on entry (PCG clears it)
static int dist_cycle(b200_handle h, b200_comm c, b200_dist_amg amg, const double *f, double *u /* capacity >= lv[0].cap */) {
  const int nl = (int)amg->lv.size();
  const double w = amg->relax_wt;
  const int me = b200_comm_rank(c), R = b200_comm_size(c);
  auto coarse_solve = [&](dist_level &L, const double *F, double *U) -> int {
    if (amg->tail) {
      // gather the level's right-hand side on every rank, cycle the replicated hierarchy, keep this rank's rows
      const int N = amg->tail_N, first = amg->tail_first, n = L.n;
      if (amg->tail_plan->any_traffic) B200_TRY(b200_halo_forward_f64(h, c, amg->tail_plan, F, amg->tail_G));
      if (first) B200_CUDA(cudaMemcpyAsync(amg->tail_F, amg->tail_G, sizeof(double) * (size_t)first, cudaMemcpyDeviceToDevice, h->stream));
      if (n) B200_CUDA(cudaMemcpyAsync(amg->tail_F + first, F, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
      if (N - first - n) B200_CUDA(cudaMemcpyAsync(amg->tail_F + first + n, amg->tail_G + first, sizeof(double) * (size_t)(N - first - n), cudaMemcpyDeviceToDevice, h->stream));
      B200_TRY(b200_amg_precond(h, amg->tail, amg->tail_F, amg->tail_U));
      if (n) B200_CUDA(cudaMemcpyAsync(U, amg->tail_U + first, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
      return 0;
    }
    if (amg->coarse_ge && amg->ge_n <= 16 && b200_comm_p2p_ok(c)) {
      // Allgatherv of f (par_gauss_elim.c:264) as a rank-ordered device-to-device sum of vectors that are zero outside the
      // rank's own rows (x + 0 is exact): one kernel, no message library, capturable in the iteration's CUDA graph
      const int ncg = amg->ge_n;
      ge_scatter_kernel<<<1, 32, 0, h->stream>>>(ncg, amg->ge_starts[me], L.n, F, amg->ge_f + 2 * ncg);
      B200_LAUNCH_CHECK();
      B200_TRY(b200_comm_allreduce_sum_dev2dev(h, c, amg->ge_f + 2 * ncg, ncg, amg->ge_f));
      gselim_kernel2<<<1, 32, 0, h->stream>>>(ncg, amg->ge_A, amg->ge_A + (size_t)ncg * ncg, amg->ge_f, amg->ge_f + ncg);
      B200_LAUNCH_CHECK();
      if (L.n) B200_CUDA(cudaMemcpyAsync(U, amg->ge_f + ncg + amg->ge_starts[me], sizeof(double) * (size_t)L.n, cudaMemcpyDeviceToDevice, h->stream));
      return 0;
    }
    if (amg->coarse_ge) {
      const int ncg = amg->ge_n;
      std::vector<b200_xfer> sends, recvs;               // Allgatherv of f (par_gauss_elim.c:264)
      for (int r = 0; r < R; r++) {
        const int rn = amg->ge_starts[r + 1] - amg->ge_starts[r];
        if (L.n) sends.push_back({r, const_cast<double *>(F), sizeof(double) * (size_t)L.n, 0});
        if (rn) recvs.push_back({r, amg->ge_f + amg->ge_starts[r], sizeof(double) * (size_t)rn, 0});
      }
      B200_TRY(b200_comm_exchange(h, c, sends, recvs));
      gselim_kernel2<<<1, 32, 0, h->stream>>>(ncg, amg->ge_A, amg->ge_A + (size_t)ncg * ncg, amg->ge_f, amg->ge_f + ncg);
      B200_LAUNCH_CHECK();
      if (L.n) B200_CUDA(cudaMemcpyAsync(U, amg->ge_f + ncg + amg->ge_starts[me], sizeof(double) * (size_t)L.n, cudaMemcpyDeviceToDevice, h->stream));
      return 0;
    }
    if (amg->gs) {
      if (L.A->L->gs) {
        if (L.A->halo->ng) B200_CUDA(cudaMemsetAsync(U + L.A->n_owned_cols, 0, sizeof(double) * (size_t)L.A->halo->ng, h->stream));
        return b200_gs_relax(h, L.A->L->gs, L.A->L, amg->relax_down, true, F, L.l1, U);
      }
      return 0;
    }
    if (L.n) {
      jacobi_zero_kernel2<<<vgrid(h, L.n), 256, 0, h->stream>>>((size_t)L.n, w, F, L.l1, U);
      B200_LAUNCH_CHECK();
    }
    return 0;
  };
  if (nl == 1) return coarse_solve(amg->lv[0], f, u);
  std::vector<const double *> F(nl);
  std::vector<double *> U(nl);
  F[0] = f;
  if (amg->gs) {
    // in-place hybrid Gauss-Seidel smoothing: halo of u (old values for everything off rank), then the sweep
    auto relax = [&](dist_level &L, int type, const double *Fl, double *Ul, bool zero) -> int {
      b200_dist_matrix M = L.A;
      if (!zero && M->halo->any_traffic) B200_TRY(b200_halo_forward_f64(h, c, M->halo, Ul, Ul + M->n_owned_cols));
      // zero iterate: the reference's Vext is all zeros for BOTH halves of a symmetric sweep (par_relax.c:3540-3570 builds
      // it once per call); the backward half reads the ghost tail, which still holds the previous cycle's halo -> clear it
      if (zero && M->halo->ng) B200_CUDA(cudaMemsetAsync(Ul + M->n_owned_cols, 0, sizeof(double) * (size_t)M->halo->ng, h->stream));
      return b200_gs_relax(h, M->L->gs, M->L, type, zero, Fl, L.l1, Ul);
    };
    U[0] = u;
    for (int l = 1; l < nl; l++) { F[l] = amg->lv[l].F; U[l] = amg->lv[l].U; }
    for (int l = 0; l < nl - 1; l++) {
      dist_level &L = amg->lv[l];
      B200_TRY(relax(L, amg->relax_down, F[l], U[l], true));                                   // iterate is 0 on entry
      B200_TRY(dist_spmv(h, c, L.A, U[l], amg->Vtemp, 0, -1.0, 1.0, F[l], nullptr));
      B200_TRY(dist_spmv(h, c, L.R, amg->Vtemp, amg->lv[l + 1].F, 0, 1.0, 0.0, nullptr, nullptr));
    }
    B200_TRY(coarse_solve(amg->lv[nl - 1], F[nl - 1], U[nl - 1]));
    for (int l = nl - 2; l >= 0; l--) {
      dist_level &L = amg->lv[l];
      B200_TRY(dist_spmv(h, c, L.P, U[l + 1], U[l], 0, 1.0, 1.0, U[l], nullptr));
      B200_TRY(relax(L, amg->relax_up, F[l], U[l], false));
    }
    return 0;
  }
  // Every level keeps two buffers with fixed roles -- U: the pre-smoothed iterate, later the level's final iterate; T: the
  // iterate after the coarse-grid correction -- so the kernel arguments of a cycle never change (CUDA-graph replay in PCG).
  for (int l = 1; l < nl; l++) F[l] = amg->lv[l].F;
  for (int l = 0; l < nl - 1; l++) {
    dist_level &L = amg->lv[l];
    dist_level &Lc = amg->lv[l + 1];
    if (L.n) {
      b200_prof_scope ps(h, "presmooth", l);
      jacobi_zero_kernel2<<<vgrid(h, L.n), 256, 0, h->stream>>>((size_t)L.n, w, F[l], L.l1, L.U);
      B200_LAUNCH_CHECK();
    }
    U[l] = L.U;
    B200_TRY(dist_spmv(h, c, L.A, L.U, amg->Vtemp, 0, -1.0, 1.0, F[l], nullptr, "residual", l));        // Vtemp = F - A U
    B200_TRY(dist_spmv(h, c, L.R, amg->Vtemp, Lc.F, 0, 1.0, 0.0, nullptr, nullptr, "restrict", l));      // F_{l+1} = R Vtemp
  }
  {
    dist_level &L = amg->lv[nl - 1];
    b200_prof_scope ps(h, "coarse solve", nl - 1);
    B200_TRY(coarse_solve(L, L.F, L.U));
    U[nl - 1] = L.U;
  }
  for (int l = nl - 2; l >= 0; l--) {
    dist_level &L = amg->lv[l];
    B200_TRY(dist_spmv(h, c, L.P, U[l + 1], L.T, 0, 1.0, 1.0, L.U, nullptr, "prolong", l));              // T_l = U_l + P U_{l+1}
    B200_TRY(dist_spmv(h, c, L.A, L.T, (l == 0) ? u : L.U, 1, w, 0.0, F[l], L.l1, "postsmooth", l));     // l1-Jacobi post-sweep
  }
  return 0;
}

// hypre_ParVectorInnerProd (par_vector.c:481-501): local partial on the device, one all-gather, ranks added in order
static int dist_dot(b200_handle h, b200_comm c, int n, const double *x, const double *y, double *d_scratch, double *result) {
  B200_TRY(b200_vec_dot_dev(h, n, x, y, d_scratch));
  return b200_comm_allreduce_sum_dev(h, c, d_scratch, 1, result);
}
// <x1,y1> and <x2,y2> with ONE exchange (the two inner products PCG needs at the same point of an iteration)
static int dist_dot2(b200_handle h, b200_comm c, int n, const double *x1, const double *y1, const double *x2, const double *y2,
                     double *d_scratch, double *r1, double *r2) {
  B200_TRY(b200_vec_dot_dev(h, n, x1, y1, d_scratch));
  B200_TRY(b200_vec_dot_dev(h, n, x2, y2, d_scratch + 1));
  double out[2] = {0, 0};
  B200_TRY(b200_comm_allreduce_sum_dev(h, c, d_scratch, 2, out));
  *r1 = out[0]; *r2 = out[1];
  return 0;
}

namespace {
// device-resident PCG scalars: sc[0] gamma, sc[1] i_prod = <r,r>, sc[2] <s,p>, sc[3] gamma_old, sc[4] alpha, sc[5] beta
__global__ void dpcg_alpha_kernel(double *sc) {
  sc[4] = (sc[2] != 0.0) ? sc[0] / sc[2] : 0.0;   // alpha = gamma / <s,p> (pcg.c:522); <s,p> = 0 is the error path, x and r stay intact
  sc[3] = sc[0];                                  // gamma_old = gamma (pcg.c:530)
}
__global__ void dpcg_beta_kernel(double *sc) { sc[5] = sc[0] / sc[3]; }   // beta = gamma / gamma_old (pcg.c:729)
__global__ void dpcg_update_xr_kernel(size_t n, const double *__restrict__ sc, const double *__restrict__ p,
                                      const double *__restrict__ s, double *__restrict__ x, double *__restrict__ r) {
  const double alpha = sc[4];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    x[i] += alpha * p[i];         // pcg.c:534
    r[i] += -alpha * s[i];        // pcg.c:539
  }
}
__global__ void dpcg_update_p_kernel(size_t n, const double *__restrict__ sc, const double *__restrict__ s, double *__restrict__ p) {
  const double beta = sc[5];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = beta * p[i] + 1.0 * s[i];   // Scale then Axpy (pcg.c:734-735)
}
}  // namespace

// hypre_PCGSolve across ranks (krylov/pcg.c:271-757, two_norm 1 as ij.c:3892-3897 sets it).  The three inner products of an
// iteration are reduced device to device (rank-ordered sums over NVLink, b200_comm_allreduce_sum_dev2dev), alpha and beta are
// formed on the device and the vector updates are fused, exactly like the single-GPU loop (b200_pcg_solve_ex): ONE host
// synchronisation per iteration, for the convergence test.
extern "C" int b200_dist_pcg_solve(b200_handle h, b200_comm c, b200_dist_matrix A, b200_dist_amg amg, const double *d_b,
                                   double *d_x, double tol, int max_iter, int *iters_out, double *final_rel_res, double *h_norms) {
  if (!A || !A->L) B200_FAIL("dist_pcg: matrix not localized");
  const int n = A->n;
  int cap = n + A->halo->ng + 8;
  if (amg) cap = std::max(cap, amg->lv[0].cap);
  double *p = nullptr, *s = nullptr, *r = nullptr, *xx = nullptr, *sc = nullptr;
  B200_TRY(b200_dalloc<double>(h, &p, cap)); B200_TRY(b200_dalloc<double>(h, &s, cap)); B200_TRY(b200_dalloc<double>(h, &r, cap));
  B200_TRY(b200_dalloc<double>(h, &xx, cap)); B200_TRY(b200_dalloc<double>(h, &sc, 16));
  double *lp = sc + 8;                       // this rank's partial sums
  double *hs = h->h_pinned;
  B200_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * cap, h->stream));
  B200_CUDA(cudaMemsetAsync(s, 0, sizeof(double) * cap, h->stream));
  B200_CUDA(cudaMemsetAsync(sc, 0, sizeof(double) * 16, h->stream));
  B200_CUDA(cudaMemcpyAsync(xx, d_x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
  auto precond = [&](const double *rhs, double *out) -> int {
    if (amg) return dist_cycle(h, c, amg, rhs, out);
    return b200_vec_copy(h, n, rhs, out);
  };
  auto fetch = [&]() -> int {                 // the iteration's one host synchronisation
    B200_CUDA(cudaMemcpyAsync(hs, sc, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
  };
  const int vg = vgrid(h, n);
  int rc = 0, i = 0;
  double bi_prod = 0, i_prod = 0, eps = tol * tol;
  bool x_is_b = false;
  do {
    if ((rc = b200_vec_dot_dev(h, n, d_b, d_b, lp))) break;
    if ((rc = b200_comm_allreduce_sum_dev2dev(h, c, lp, 1, sc + 1))) break;
    if ((rc = fetch())) break;
    bi_prod = hs[1];
    if (bi_prod != 0. && !(bi_prod / bi_prod == bi_prod / bi_prod)) {       // pcg.c:359-381
      rc = b200_set_error(__FILE__, __LINE__, "hypre_PCGSolve: INFs and/or NaNs detected in input"); break;
    }
    if (!(bi_prod > 0.0)) {                                                 // pcg.c:403-416: b == 0 -> x = b
      rc = b200_vec_copy(h, n, d_b, d_x); x_is_b = true; if (h_norms) h_norms[0] = 0.0; break;
    }
    if ((rc = dist_spmv(h, c, A, xx, r, 0, -1.0, 1.0, d_b, nullptr))) break;       // r = b - A x
    if ((rc = precond(r, p))) break;
    if ((rc = b200_vec_dot_dev(h, n, r, p, lp))) break;                            // gamma = <r,p>
    if ((rc = b200_vec_dot_dev(h, n, r, r, lp + 1))) break;
    if ((rc = b200_comm_allreduce_sum_dev2dev(h, c, lp, 2, sc))) break;
    if ((rc = fetch())) break;
    if (hs[0] != 0. && !(hs[0] / hs[0] == hs[0] / hs[0])) {                 // pcg.c:440-462: INF -> NaN conversion on gamma
      rc = b200_set_error(__FILE__, __LINE__, "hypre_PCGSolve: INFs and/or NaNs detected in input"); break;
    }
    if (h_norms) h_norms[0] = std::sqrt(hs[1]);
    // One iteration = [beta, p update] of the previous one + [halo, s = A p, <s,p>, reduction, alpha, x/r update, the cycle,
    // <r,s>, <r,r>, reduction, copy of the scalars to the host].  When every exchange of iteration 1 stayed on the device-only
    // path (direct halos and reductions; a single rank trivially), the sequence is a fixed list of kernels with fixed arguments:
    // iteration 2 is captured into a CUDA graph, iterations 3.. replay it.
    auto body = [&](bool with_beta) -> int {
      if (with_beta) {
        b200_prof_scope ps(h, "pcg update p");
        dpcg_beta_kernel<<<1, 1, 0, h->stream>>>(sc);
        ++g_b200_launches;
        dpcg_update_p_kernel<<<vg, 256, 0, h->stream>>>((size_t)n, sc, s, p);
        ++g_b200_launches;
      }
      B200_TRY(dist_spmv(h, c, A, p, s, 0, 1.0, 0.0, nullptr, nullptr, "pcg s=Ap"));     // s = A p (pcg.c:512)
      { b200_prof_scope ps(h, "pcg dot");
        B200_TRY(b200_vec_dot_dev(h, n, s, p, lp)); }                                    // <s,p> (pcg.c:515)
      { b200_prof_scope ps(h, "pcg allreduce");
        B200_TRY(b200_comm_allreduce_sum_dev2dev(h, c, lp, 1, sc + 2)); }
      { b200_prof_scope ps(h, "pcg update x,r");
        dpcg_alpha_kernel<<<1, 1, 0, h->stream>>>(sc);
        ++g_b200_launches;
        dpcg_update_xr_kernel<<<vg, 256, 0, h->stream>>>((size_t)n, sc, p, s, xx, r);
        ++g_b200_launches; }
      { b200_prof_scope ps(h, "pcg precond (whole cycle)");
        B200_TRY(precond(r, s)); }                                                       // s = C r (pcg.c:568-569)
      { b200_prof_scope ps(h, "pcg dot");
        B200_TRY(b200_vec_dot2_dev(h, n, r, s, lp, lp + 1)); }                          // gamma = <r,s> (pcg.c:572), i_prod = <r,r> (:590)
      { b200_prof_scope ps(h, "pcg allreduce");
        B200_TRY(b200_comm_allreduce_sum_dev2dev(h, c, lp, 2, sc)); }
      B200_CUDA(cudaMemcpyAsync(hs, sc, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      return 0;
    };
    cudaGraphExec_t gexec = nullptr;
    bool graph_ok = b200_graph_enabled() && (!amg || !amg->gs), graph_tried = false;
    long long launches_per_graph = 0;
    while ((i + 1) <= max_iter) {
      i++;
      if (i >= 2 && gexec) {
        if ((rc = b200_graph_launch(h, gexec))) break;
        g_b200_launches += launches_per_graph;
      } else if (i == 2 && graph_ok && !graph_tried) {
        graph_tried = true;
        const long long l0 = g_b200_launches.load();
        if ((rc = b200_graph_begin(h))) break;
        rc = body(true);
        cudaGraphExec_t ge = nullptr;
        int rc2 = b200_graph_end(h, &ge);
        if (rc || rc2) { if (!rc) rc = rc2; break; }
        launches_per_graph = g_b200_launches.load() - l0;
        if (ge) { gexec = ge; if ((rc = b200_graph_launch(h, gexec))) break; }
        else { g_b200_launches -= launches_per_graph; if ((rc = body(true))) break; }      // capture refused: run it eagerly
      } else {
        const long long ops0 = b200_comm_host_ops(c);
        if ((rc = body(i > 1))) break;
        if (b200_comm_host_ops(c) != ops0) graph_ok = false;                 // an exchange left the device-only path
      }
      { b200_prof_scope ps(h, "pcg fetch+sync");
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = b200_set_error(__FILE__, __LINE__, "pcg sync failed"); break; } }
      const double gamma = hs[0], sdotp = hs[2];
      i_prod = hs[1];
      if (sdotp == 0.0) { rc = b200_set_error(__FILE__, __LINE__, "Zero sdotp value in PCG"); break; }   // pcg.c:516-521
      if (h_norms) h_norms[i] = std::sqrt(i_prod);
      if (i_prod / bi_prod < eps) break;
      if (!(gamma > 2.2250738585072014e-308)) { rc = b200_set_error(__FILE__, __LINE__, "Subnormal gamma value in PCG"); break; }
    }
    b200_graph_destroy(gexec);
  } while (0);
  b200_prof_report(h, "b200_dist_pcg_solve");
  if (!rc) {
    if (!x_is_b) cudaMemcpyAsync(d_x, xx, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream);
    if (iters_out) *iters_out = i;
    if (final_rel_res) *final_rel_res = bi_prod > 0.0 ? std::sqrt(i_prod / bi_prod) : 0.0;
  }
  b200_dfree(h, p); b200_dfree(h, s); b200_dfree(h, r); b200_dfree(h, xx); b200_dfree(h, sc);
  return rc;
}

// ---- GMRES / BiCGSTAB across ranks: the loops of b200_krylov.cu over the row-partitioned operator ----------------------
// (hypre_GMRESSolve / hypre_BiCGSTABSolve with the hypre_ParKrylov* callbacks, parcsr_ls/par_krylov_func.c: matvec =
// halo exchange + one kernel, inner product = local partial + one rank-ordered reduction, preconditioner = the
// distributed cycle from a zero guess).  Every inner product of the Gram-Schmidt sweep is a reduction over the ranks.
static int dist_krylov_ops(b200_handle h, b200_comm c, b200_dist_matrix A, b200_dist_amg amg, int precond, b200_krylov_ops *ops) {
  if (!A || !A->L) B200_FAIL("dist krylov: matrix not localized");
  if (precond != 0 && precond != 1) B200_FAIL("dist krylov: precond must be 0 (none) or 1 (BoomerAMG)");
  if (precond == 1 && !amg) B200_FAIL("dist krylov: precond 1 needs a set-up BoomerAMG hierarchy");
  const int n = A->n;
  int cap = n + A->halo->ng + 8;
  if (amg) cap = std::max(cap, amg->lv[0].cap);
  ops->n = n;
  ops->cap = cap;
  ops->device_mgs = false;
  ops->matvec = [h, c, A](double alpha, const double *x, double beta, const double *b, double *y) {
    return dist_spmv(h, c, A, const_cast<double *>(x), y, 0, alpha, beta, b, nullptr);    // x is a work vector with a ghost tail
  };
  ops->precond = [h, c, amg, precond, n](const double *rhs, double *out) {
    if (precond == 1) return dist_cycle(h, c, amg, rhs, out);
    return b200_vec_copy(h, n, rhs, out);
  };
  ops->reduce = [h, c](const double *d_partials, int k, double *out) {
    return b200_comm_allreduce_sum_dev(h, c, d_partials, k, out);
  };
  return 0;
}
extern "C" int b200_dist_gmres_solve(b200_handle h, b200_comm c, b200_dist_matrix A, b200_dist_amg amg, const b200_gmres_params *prm,
                                     const double *d_b, double *d_x, int *iters, double *final_rel_res, double *h_norms,
                                     int *converged) {
  if (!prm) B200_FAIL("dist_gmres: null argument");
  b200_krylov_ops ops;
  B200_TRY(dist_krylov_ops(h, c, A, amg, prm->precond, &ops));
  return b200_gmres_core(h, &ops, prm, d_b, d_x, iters, final_rel_res, h_norms, converged);
}
extern "C" int b200_dist_bicgstab_solve(b200_handle h, b200_comm c, b200_dist_matrix A, b200_dist_amg amg,
                                        const b200_bicgstab_params *prm, const double *d_b, double *d_x, int *iters,
                                        double *final_rel_res, double *h_norms, int *converged) {
  if (!prm) B200_FAIL("dist_bicgstab: null argument");
  b200_krylov_ops ops;
  B200_TRY(dist_krylov_ops(h, c, A, amg, prm->precond, &ops));
  return b200_bicgstab_core(h, &ops, prm, d_b, d_x, iters, final_rel_res, h_norms, converged);
}
