// b200_csr.cu -- device CSR container and the streaming CSR SpMV family.
//
// Reference behaviour reproduced: hypre_CSRMatrixMatvecOutOfPlaceHost
// (seq_mv/csr_matvec.c:24-376): y = alpha*A*x + beta*b, row sums taken over the row's
// entries in storage order.
//
// B200 design (not a translation of the reference's row loop / cuSPARSE call):
//   * a per-matrix *row-block plan* cuts the nonzero stream into tiles of ~TILE entries
//     aligned to row boundaries (blk_row[b] = first row whose first entry is >= b*TILE);
//   * each CTA streams its tile of (col,val) with aligned 128-bit loads (int4 / double2),
//     gathers x through the read-only path, and parks the products in shared memory;
//   * rows are then reduced out of shared memory by G cooperating threads (G chosen from the
//     mean row length: 1 for stencil rows, up to 32 for dense coarse rows) and a fused epilogue
//     writes y (axpby), the l1-Jacobi update, or the residual.
//   HBM sees every matrix byte exactly once, fully coalesced, independent of row length.
#include "b200_internal.h"
#include <algorithm>

namespace {

constexpr int NT = B200_SPMV_NT;              // threads per CTA
constexpr int CAP = 4096;                     // shared-memory product slots per CTA (32 KB)
}
constexpr int MAX_TILE_DEFAULT = B200_SPMV_MAX_TILE;
namespace {

__global__ void max_row_kernel(int n, const int *__restrict__ A_i, int *__restrict__ out_max) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int len = i < n ? A_i[i + 1] - A_i[i] : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) len = max(len, __shfl_down_sync(0xffffffffu, len, off));
  if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out_max, len);
}

// Gather-cost model for choosing G (lanes per row): for a sample of 32-row windows, count the
// distinct 128-byte lines of x one warp-wide gather would touch under each candidate G -- that is
// the number of L1 wavefronts the load costs, and L1 wavefronts are what bound the SpMV once the
// matrix stream is staged by bulk copies.  Banded stencils come out cheapest at G = 1 (lanes own
// consecutive rows: their p-th columns are adjacent), irregular coarse operators at G = 16/32.
__global__ void gather_cost_kernel(int nrows, const int *__restrict__ A_i, const int *__restrict__ A_j, int nwin,
                                   int stride, int gi_lo, int gi_hi, unsigned long long *__restrict__ cost /* [6][2] */) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= nwin) return;
  const int rbase = (int)(((long long)w * stride) * 32);
  for (int gi = gi_lo; gi <= gi_hi; gi++) {         // only the lane counts the plan may pick
    const int G = 1 << gi, rows_per = 32 / G;
    unsigned long long wf = 0, ent = 0;
    for (int rb = 0; rb < G; rb++) {               // the 32 rows are covered by G passes of 32/G rows
      const int row = rbase + rb * rows_per + lane / G;
      const int ln = lane % G;
      int s0 = 0, s1 = 0;
      if (row < nrows) { s0 = A_i[row]; s1 = A_i[row + 1]; }
      int maxlen = s1 - s0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));
      for (int p0 = 0; p0 < maxlen; p0 += G) {
        const int e = s0 + p0 + ln;
        const bool valid = e < s1;
        const int line = valid ? (A_j[e] >> 4) : (-1 - lane);
        const unsigned m = __match_any_sync(0xffffffffu, line);
        const unsigned lead = __ballot_sync(0xffffffffu, valid && (__ffs(m) - 1 == lane));
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        wf += __popc(lead);
        ent += __popc(vm);
      }
    }
    if (lane == 0) { atomicAdd(&cost[2 * gi], wf); atomicAdd(&cost[2 * gi + 1], ent); }
  }
}

__global__ void plan_kernel(const int *__restrict__ A_i, int nrows, int tile, int nblk, int *blk_row, int *blk_ent) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nblk) return;
  if (b == nblk) { blk_row[b] = nrows; blk_ent[b] = A_i[nrows]; return; }
  long long target = (long long)b * tile;
  int lo = 0, hi = nrows;          // first r in [0,nrows] with A_i[r] >= target
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((long long)A_i[mid] >= target) hi = mid; else lo = mid + 1;
  }
  blk_row[b] = lo;
  blk_ent[b] = A_i[lo];
}

__global__ void meta_kernel(int nblk, const int *__restrict__ blk_row, const int *__restrict__ blk_ent, int4 *__restrict__ meta) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nblk) meta[b] = make_int4(blk_row[b], blk_row[b + 1], blk_ent[b], blk_ent[b + 1]);
}

struct Epi {
  int mode;            // 0: y = alpha*s + beta*b    1: y = x[r] + w*(b[r]-s)/d[r]  (l1-Jacobi)
  double alpha, beta;  // mode 1: alpha = relaxation weight w
  const double *b;     // mode 0: b (may be null when beta == 0); mode 1: f
  const double *d;     // mode 1: l1 norms
};

__device__ __forceinline__ void epilogue(const Epi &e, int r, double s, const double *__restrict__ x,
                                         double *__restrict__ y) {
  if (e.mode == 0) {
    double v = e.alpha * s;
    if (e.beta != 0.0) v += e.beta * e.b[r];
    y[r] = v;
  } else {
    y[r] = x[r] + e.alpha * (e.b[r] - s) / e.d[r];
  }
}

template <int G>
__global__ void __launch_bounds__(NT)
spmv_stream_kernel(const int *__restrict__ A_i, const int *__restrict__ A_j,
                   const double *__restrict__ A_a, const int *__restrict__ blk_row,
                   const double *__restrict__ x, double *__restrict__ y, Epi epi, int vec_ok) {
  __shared__ double p[CAP + 4];
  const int tid = threadIdx.x;
  const int r0 = blk_row[blockIdx.x], r1 = blk_row[blockIdx.x + 1];
  if (r0 >= r1) return;
  const int e0 = A_i[r0], e1 = A_i[r1];
  const int a0 = e0 & ~3;                      // 16B/32B aligned start of the tile
  if (e1 - a0 <= CAP) {
    if (vec_ok) {
#pragma unroll 2
      for (int k = a0 + 4 * tid; k < e1; k += 4 * NT) {
        const int4 c = *reinterpret_cast<const int4 *>(A_j + k);
        const double2 v0 = *reinterpret_cast<const double2 *>(A_a + k);
        const double2 v1 = *reinterpret_cast<const double2 *>(A_a + k + 2);
        double *q = p + (k - a0);
        q[0] = (k     >= e0 && k     < e1) ? v0.x * __ldg(x + c.x) : 0.0;
        q[1] = (k + 1 >= e0 && k + 1 < e1) ? v0.y * __ldg(x + c.y) : 0.0;
        q[2] = (k + 2 >= e0 && k + 2 < e1) ? v1.x * __ldg(x + c.z) : 0.0;
        q[3] = (k + 3 >= e0 && k + 3 < e1) ? v1.y * __ldg(x + c.w) : 0.0;
      }
    } else {
#pragma unroll 4
      for (int k = e0 + tid; k < e1; k += NT) p[k - a0] = A_a[k] * __ldg(x + A_j[k]);
    }
    __syncthreads();
    const int sub = tid / G, lane = tid % G;
    for (int base = r0; base < r1; base += NT / G) {
      const int r = base + sub;
      double s = 0.0;
      if (r < r1) {
        const int s0 = A_i[r] - a0, s1 = A_i[r + 1] - a0;
        for (int k = s0 + lane; k < s1; k += G) s += p[k];
      }
      if (G > 1) {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off, G);
      }
      if (r < r1 && lane == 0) epilogue(epi, r, s, x, y);
    }
  } else {
    // a row longer than the shared tile: warp-per-row straight from global memory
    const int warp = tid >> 5, ln = tid & 31;
    for (int r = r0 + warp; r < r1; r += NT / 32) {
      double s = 0.0;
      for (int k = A_i[r] + ln; k < A_i[r + 1]; k += 32) s += A_a[k] * __ldg(x + A_j[k]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
      if (ln == 0) epilogue(epi, r, s, x, y);
    }
  }
}

template <int G>
int launch_stream(b200_handle h, b200_csr A, const double *x, double *y, const Epi &epi) {
  int vec_ok = A->owns ? 1 : 0;
  spmv_stream_kernel<G><<<A->nblk, NT, 0, h->stream>>>(A->i, A->j, A->a, A->blk_row, x, y, epi, vec_ok);
  B200_LAUNCH_CHECK();
  return 0;
}

__global__ void scale_copy_kernel(int n, double beta, const double *__restrict__ b, double *__restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (beta == 0.0) ? 0.0 : beta * b[i];
}

}  // namespace

bool b200_spmv_pipe_ok(b200_csr A);
bool b200_spmv_dict_ok(b200_csr A);
int b200_csr_build_dict(b200_handle h, b200_csr A);
int b200_csr_drop_dict(b200_handle h, b200_csr A);
int b200_csr_spmv_dict(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha, double beta,
                       const double *b, const double *d);
int b200_csr_spmv_pipe(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha, double beta,
                       const double *b, const double *d);

int b200_csr_spmv_epi(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha,
                      double beta, const double *b, const double *d) {
  if (A->nrows == 0) return 0;
  if (!A->blk_row) B200_TRY(b200_csr_build_plan(h, A));
  static const bool force_v1 = [] { const char *e = getenv("B200_SPMV_TWO_PHASE"); return e && e[0] == '1'; }();
  if (!force_v1 && b200_spmv_dict_ok(A)) return b200_csr_spmv_dict(h, A, x, y, mode, alpha, beta, b, d);
  if (!force_v1 && b200_spmv_pipe_ok(A)) return b200_csr_spmv_pipe(h, A, x, y, mode, alpha, beta, b, d);
  Epi e{mode, alpha, beta, b, d};
  switch (A->group) {
    case 1:  return launch_stream<1>(h, A, x, y, e);
    case 2:  return launch_stream<2>(h, A, x, y, e);
    case 4:  return launch_stream<4>(h, A, x, y, e);
    case 8:  return launch_stream<8>(h, A, x, y, e);
    case 16: return launch_stream<16>(h, A, x, y, e);
    default: return launch_stream<32>(h, A, x, y, e);
  }
}

int b200_csr_alloc(b200_handle h, int nrows, int ncols, int nnz, bool with_data, b200_csr *out) {
  b200_csr A = new b200_csr_s();
  A->nrows = nrows; A->ncols = ncols; A->nnz = nnz; A->owns = true;
  B200_TRY(b200_dalloc<int>(h, &A->i, (size_t)nrows + 1 + 4));      // + 4: the bulk copy of a tile's row pointers is rounded up to 16 bytes
  B200_TRY(b200_dalloc<int>(h, &A->j, (size_t)nnz + B200_PAD));
  if (with_data) B200_TRY(b200_dalloc<double>(h, &A->a, (size_t)nnz + B200_PAD));
  *out = A;
  return 0;
}

int b200_csr_build_plan(b200_handle h, b200_csr A) {
  B200_TRY(b200_csr_drop_dict(h, A));
  if (A->blk_row) { B200_TRY(b200_dfree(h, A->blk_row)); A->blk_row = nullptr; }
  if (A->blk_ent) { B200_TRY(b200_dfree(h, A->blk_ent)); A->blk_ent = nullptr; }
  if (A->blk_meta) { B200_TRY(b200_dfree(h, A->blk_meta)); A->blk_meta = nullptr; }
  double avg = A->nrows ? (double)A->nnz / A->nrows : 0.0;
  static const int tile_env = [] { const char *e = getenv("B200_SPMV_TILE"); return e ? atoi(e) : 0; }();
  const int MAX_TILE = tile_env > 0 ? tile_env : ::MAX_TILE_DEFAULT;
  int G = 1;
  if (A->nnz > 0 && A->nrows >= 64) {
    // pick the lanes-per-row that minimises gather wavefronts per entry on a sample of the matrix
    const int windows = A->nrows / 32;
    const int nwin = windows < 2048 ? windows : 2048;
    const int stride = windows / nwin;
    unsigned long long *d_cost = nullptr, h_cost[12];
    B200_TRY(b200_dalloc<unsigned long long>(h, &d_cost, 12));
    B200_CUDA(cudaMemsetAsync(d_cost, 0, sizeof(h_cost), h->stream));
    // lanes-per-row must keep the CTA busy: a tile of MAX_TILE entries holds MAX_TILE/avg rows, so fewer
    // than avg*NT/MAX_TILE lanes per row would leave threads idle
    int gmin = 1;
    while (gmin < 32 && gmin * MAX_TILE < avg * NT) gmin <<= 1;
    int gi_lo = 0, gi_hi = 0;
    while ((1 << gi_lo) < gmin) gi_lo++;
    gi_hi = gi_lo;
    while (gi_hi < 5 && (2 << gi_hi) <= 2 * avg) gi_hi++;      // the selection loop below stops at the first g > gmin with g > 2 avg
    static const bool eval_all = [] { const char *e = getenv("B200_PLAN_EVAL_ALL"); return e && e[0] == '1'; }();   // diagnosis
    gather_cost_kernel<<<b200_grid((size_t)nwin * 32, 128), 128, 0, h->stream>>>(A->nrows, A->i, A->j, nwin, stride, eval_all ? 0 : gi_lo,
                                                                             eval_all ? 5 : gi_hi, d_cost);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(h_cost, d_cost, sizeof(h_cost), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dfree(h, d_cost));
    double best = 1e300;
    static const bool dbg2 = [] { const char *e = getenv("B200_DEBUG_PLAN"); return e && e[0] == '1'; }();
    for (int gi = eval_all ? 0 : gi_lo; gi <= (eval_all ? 5 : gi_hi); gi++) {
      const int g = 1 << gi;
      if (eval_all && dbg2) fprintf(stderr, "[b200 plan]   (all) G=%d wavefronts/entry=%.3f\n", g, (double)h_cost[2 * gi] / (double)(h_cost[2 * gi + 1] ? h_cost[2 * gi + 1] : 1));
      if (gi < gi_lo || gi > gi_hi) continue;
      if (g < gmin) continue;
      if (g > gmin && g > 2 * avg) break;          // more lanes than a row has entries: mostly idle
      const double c = (double)h_cost[2 * gi] / (double)(h_cost[2 * gi + 1] ? h_cost[2 * gi + 1] : 1);
      if (dbg2) fprintf(stderr, "[b200 plan]   G=%d wavefronts/entry=%.3f\n", g, c);
      if (c < best * 0.85) { best = c; G = g; }    // widen only for a clear (15%) saving in L1 wavefronts
    }
    if (G < gmin) G = gmin;
  }
  // tile: a whole number of row passes (NT/G rows each), close to MAX_TILE entries
  double per_pass = avg * (NT / G);
  int passes = per_pass > 0 ? (int)(MAX_TILE / per_pass) : 1;
  if (passes < 1) passes = 1;
  static const int tile_rows = [] { const char *e = getenv("B200_SPMV_TILE_ROWS"); const int v = e ? atoi(e) : 160; return v < 32 ? 32 : (v > B200_SPMV_RCAP - 32 ? B200_SPMV_RCAP - 32 : v); }();
  if (passes * (NT / G) > tile_rows) passes = tile_rows / (NT / G) > 0 ? tile_rows / (NT / G) : 1;   // row pointers staged per tile
  int tile = (int)(per_pass * passes * 0.97);
  if (tile > MAX_TILE) tile = MAX_TILE;
  if (tile < 128) tile = 128;
  static const bool dbg = [] { const char *e = getenv("B200_DEBUG_PLAN"); return e && e[0] == '1'; }();
  if (dbg) fprintf(stderr, "[b200 plan] rows=%d nnz=%d avg=%.1f G=%d passes=%d tile=%d\n", A->nrows, A->nnz, avg, G, passes, tile);
  A->group = G;
  A->tile = tile;
  A->nblk = A->nnz > 0 ? (int)(((long long)A->nnz + tile - 1) / tile) : 1;
  B200_TRY(b200_dalloc<int>(h, &A->blk_row, (size_t)A->nblk + 1));
  B200_TRY(b200_dalloc<int>(h, &A->blk_ent, (size_t)A->nblk + 1));
  plan_kernel<<<b200_grid((size_t)A->nblk + 1, 256), 256, 0, h->stream>>>(A->i, A->nrows, tile, A->nblk, A->blk_row,
                                                                          A->blk_ent);
  B200_LAUNCH_CHECK();
  B200_TRY(b200_dalloc<int>(h, &A->blk_meta, (size_t)4 * A->nblk));
  meta_kernel<<<b200_grid((size_t)A->nblk, 256), 256, 0, h->stream>>>(A->nblk, A->blk_row, A->blk_ent,
                                                                      reinterpret_cast<int4 *>(A->blk_meta));
  B200_LAUNCH_CHECK();
  A->max_row = 0;
  if (A->nrows) {
    int *d_max = nullptr;
    B200_TRY(b200_dalloc<int>(h, &d_max, 1));
    B200_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), h->stream));
    max_row_kernel<<<b200_grid(A->nrows, 256), 256, 0, h->stream>>>(A->nrows, A->i, d_max);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(&A->max_row, d_max, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dfree(h, d_max));
  }
  B200_TRY(b200_csr_build_dict(h, A));       // stencil-structured operators: codes instead of columns / values (b200_spmv_dict.cu)
  return 0;
}

extern "C" int b200_csr_create(b200_handle h, int nrows, int ncols, int nnz, const int *d_i,
                               const int *d_j, const double *d_a, int copy, b200_csr *out) {
  if (nrows < 0 || ncols < 0 || nnz < 0) B200_FAIL("negative dimension");
  b200_csr A = nullptr;
  if (copy) {
    B200_TRY(b200_csr_alloc(h, nrows, ncols, nnz, d_a != nullptr, &A));
    B200_CUDA(cudaMemcpyAsync(A->i, d_i, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, h->stream));
    if (nnz) B200_CUDA(cudaMemcpyAsync(A->j, d_j, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
    if (nnz && d_a) B200_CUDA(cudaMemcpyAsync(A->a, d_a, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
  } else {
    A = new b200_csr_s();
    A->nrows = nrows; A->ncols = ncols; A->nnz = nnz; A->owns = false;
    A->i = const_cast<int *>(d_i); A->j = const_cast<int *>(d_j); A->a = const_cast<double *>(d_a);
  }
  if (A->a) B200_TRY(b200_csr_build_plan(h, A));
  *out = A;
  return 0;
}

extern "C" int b200_csr_create_from_host(b200_handle h, int nrows, int ncols, int nnz, const int *h_i,
                                         const int *h_j, const double *h_a, b200_csr *out) {
  if (nrows < 0 || ncols < 0 || nnz < 0) B200_FAIL("negative dimension");
  b200_csr A = nullptr;
  B200_TRY(b200_csr_alloc(h, nrows, ncols, nnz, h_a != nullptr, &A));
  B200_CUDA(cudaMemcpyAsync(A->i, h_i, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyHostToDevice, h->stream));
  if (nnz) B200_CUDA(cudaMemcpyAsync(A->j, h_j, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
  if (nnz && h_a) B200_CUDA(cudaMemcpyAsync(A->a, h_a, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
  if (A->a) B200_TRY(b200_csr_build_plan(h, A));
  *out = A;
  return 0;
}

extern "C" int b200_csr_destroy(b200_handle h, b200_csr A) {
  if (!A) return 0;
  if (A->owns) {
    B200_TRY(b200_dfree(h, A->i));
    B200_TRY(b200_dfree(h, A->j));
    B200_TRY(b200_dfree(h, A->a));
  }
  B200_TRY(b200_dfree(h, A->blk_row));
  B200_TRY(b200_dfree(h, A->blk_ent));
  B200_TRY(b200_dfree(h, A->blk_meta));
  B200_TRY(b200_csr_drop_dict(h, A));
  B200_TRY(b200_gs_plan_destroy(h, A->gs));
  if (A->T) B200_TRY(b200_csr_destroy(h, A->T));
  delete A;
  return 0;
}

extern "C" int b200_csr_dims(b200_csr A, int *nrows, int *ncols, int *nnz) {
  if (!A) B200_FAIL("null matrix");
  if (nrows) *nrows = A->nrows;
  if (ncols) *ncols = A->ncols;
  if (nnz) *nnz = A->nnz;
  return 0;
}

extern "C" int b200_csr_download(b200_handle h, b200_csr A, int *h_i, int *h_j, double *h_a) {
  if (!A) B200_FAIL("null matrix");
  if (h_i) B200_CUDA(cudaMemcpyAsync(h_i, A->i, sizeof(int) * ((size_t)A->nrows + 1), cudaMemcpyDeviceToHost, h->stream));
  if (h_j && A->nnz) B200_CUDA(cudaMemcpyAsync(h_j, A->j, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToHost, h->stream));
  if (h_a && A->nnz && A->a) B200_CUDA(cudaMemcpyAsync(h_a, A->a, sizeof(double) * (size_t)A->nnz, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int b200_csr_matvec(b200_handle h, double alpha, b200_csr A, const double *d_x, double beta,
                               const double *d_b, double *d_y) {
  if (!A || !A->a) B200_FAIL("matvec needs a matrix with values");
  if (d_x == d_y) B200_FAIL("matvec: x must not alias y (csr_matvec_device.c:56-120 semantics)");
  if (A->nrows == 0) return 0;
  if (alpha == 0.0 || A->nnz == 0) {   // csr_matvec.c:82-95
    scale_copy_kernel<<<b200_grid(A->nrows, 256), 256, 0, h->stream>>>(A->nrows, beta, d_b, d_y);
    B200_LAUNCH_CHECK();
    return 0;
  }
  return b200_csr_spmv_epi(h, A, d_x, d_y, 0, alpha, beta, d_b, nullptr);
}

// y = alpha*A^T*x + beta*b.  hypre_CSRMatrixMatvecT (seq_mv/csr_matvec.c:424-668) scatters row by row; here
// the transpose is formed once (stable counting-sort order = ascending source row, the order in which the
// sequential scatter adds to y_j) and cached on the matrix, as hypre_ParCSRMatrixMatvecT does with
// diagT / offdT when they exist (par_csr_matvec.c:553-597); the product is then the streaming SpMV.
extern "C" int b200_csr_matvecT(b200_handle h, double alpha, b200_csr A, const double *d_x, double beta,
                                const double *d_b, double *d_y) {
  if (!A || !A->a) B200_FAIL("matvecT needs a matrix with values");
  if (d_x == d_y) B200_FAIL("matvecT: x must not alias y");
  if (!A->T) B200_TRY(b200_csr_transpose(h, A, &A->T));
  return b200_csr_matvec(h, alpha, A->T, d_x, beta, d_b, d_y);
}

// ---- solve-phase copy with the entries of every row sorted by column --------------------------------------
// The coarse Galerkin operators keep the reference's first-touch entry order (the next level's setup depends
// on it, and the parity tests compare it).  For the solve phase that order is poison for the x gather: the
// lanes working on neighbouring entries hit unrelated 128-byte lines and the L1 data pipe saturates at ~45 %
// of the HBM roofline (profiles/r1_c: l1tex lsu wavefronts 92 %).  Sorting each row by column makes
// neighbouring lanes read neighbouring x.  One warp per row, rank sort in shared memory.
namespace {
constexpr int SORT_CAP = 2048;     // longest row sorted; longer rows are copied as they are
__global__ void __launch_bounds__(128)
row_sort_copy_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                     int *__restrict__ S_j, double *__restrict__ S_a) {
  __shared__ int keys[4][SORT_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 4 + warp; r < n; r += gridDim.x * 4) {
    const int b = A_i[r], len = A_i[r + 1] - b;
    if (len > SORT_CAP) {
      for (int k = lane; k < len; k += 32) { S_j[b + k] = A_j[b + k]; S_a[b + k] = A_a[b + k]; }
      continue;
    }
    for (int k = lane; k < len; k += 32) keys[warp][k] = A_j[b + k];
    __syncwarp();
    for (int k = lane; k < len; k += 32) {
      const int key = keys[warp][k];
      int rank = 0;
      for (int t = 0; t < len; t++) rank += (keys[warp][t] < key) || (keys[warp][t] == key && t < k);
      S_j[b + rank] = key;
      S_a[b + rank] = A_a[b + k];
    }
    __syncwarp();
  }
}
}  // namespace

extern "C" int b200_csr_sorted_copy(b200_handle h, b200_csr A, b200_csr *out) {
  if (!A || !A->a) B200_FAIL("sorted copy: matrix with values required");
  b200_csr S = nullptr;
  B200_TRY(b200_csr_alloc(h, A->nrows, A->ncols, A->nnz, true, &S));
  B200_CUDA(cudaMemcpyAsync(S->i, A->i, sizeof(int) * ((size_t)A->nrows + 1), cudaMemcpyDeviceToDevice, h->stream));
  if (A->nrows) {
    const int grid = std::min(b200_grid(A->nrows, 4), h->num_sm * 16);
    row_sort_copy_kernel<<<grid, 128, 0, h->stream>>>(A->nrows, A->i, A->j, A->a, S->j, S->a);
    B200_LAUNCH_CHECK();
  }
  *out = S;
  return 0;
}

// ---- per-matrix statistics of hypre_BoomerAMGSetupStats (parcsr_ls/par_stats.c:575-606 for A_l, :866-925 for P_l) ----
// entries per row (min, max), row sums (min, max; each row summed in storage order like the reference) and, for
// interpolation matrices, min weight / max weight over the entries != 1.  One thread per row, block reduction, the
// handful of per-block partials is finished on the host.
namespace {
struct RowStats { int min_e, max_e; double min_rs, max_rs, min_w, max_w; };
__global__ void __launch_bounds__(256)
row_stats_kernel(int n, const int *__restrict__ A_i, const double *__restrict__ A_a, RowStats *__restrict__ part) {
  __shared__ RowStats sh[256];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  RowStats s;
  s.min_e = 0x7fffffff; s.max_e = 0; s.min_rs = 1e300; s.max_rs = -1e300; s.min_w = 1e300; s.max_w = 0.0;
  if (r < n) {
    const int b = A_i[r], e = A_i[r + 1];
    double sum = 0.0;
    for (int k = b; k < e; k++) {
      const double v = A_a[k];
      sum += v;
      s.min_w = v < s.min_w ? v : s.min_w;
      if (v != 1.0) s.max_w = v > s.max_w ? v : s.max_w;
    }
    s.min_e = s.max_e = e - b;
    s.min_rs = s.max_rs = sum;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      RowStats &a = sh[threadIdx.x];
      const RowStats &c = sh[threadIdx.x + off];
      a.min_e = min(a.min_e, c.min_e); a.max_e = max(a.max_e, c.max_e);
      a.min_rs = fmin(a.min_rs, c.min_rs); a.max_rs = fmax(a.max_rs, c.max_rs);
      a.min_w = fmin(a.min_w, c.min_w); a.max_w = fmax(a.max_w, c.max_w);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
}  // namespace

extern "C" int b200_csr_row_stats(b200_handle h, b200_csr A, int *min_entries, int *max_entries, double *min_rowsum,
                                  double *max_rowsum, double *min_weight, double *max_weight) {
  if (!A || !A->a) B200_FAIL("row stats: matrix with values required");
  int mn = 0, mx = 0;
  double rs0 = 0.0, rs1 = 0.0, w0 = 1.0, w1 = 0.0;                    // the reference's values for an empty matrix
  if (A->nrows > 0) {
    const int nb = b200_grid(A->nrows, 256);
    RowStats *d = nullptr;
    B200_TRY(b200_dalloc<RowStats>(h, &d, nb));
    row_stats_kernel<<<nb, 256, 0, h->stream>>>(A->nrows, A->i, A->a, d);
    B200_LAUNCH_CHECK();
    std::vector<RowStats> p(nb);
    B200_CUDA(cudaMemcpyAsync(p.data(), d, sizeof(RowStats) * nb, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dfree(h, d));
    mn = p[0].min_e; mx = p[0].max_e; rs0 = p[0].min_rs; rs1 = p[0].max_rs; w0 = p[0].min_w; w1 = p[0].max_w;
    for (int k = 1; k < nb; k++) {
      mn = std::min(mn, p[k].min_e); mx = std::max(mx, p[k].max_e);
      rs0 = std::min(rs0, p[k].min_rs); rs1 = std::max(rs1, p[k].max_rs);
      w0 = std::min(w0, p[k].min_w); w1 = std::max(w1, p[k].max_w);
    }
    if (A->nnz == 0) w0 = 1.0;
  }
  if (min_entries) *min_entries = mn;
  if (max_entries) *max_entries = mx;
  if (min_rowsum) *min_rowsum = rs0;
  if (max_rowsum) *max_rowsum = rs1;
  if (min_weight) *min_weight = w0;
  if (max_weight) *max_weight = w1;
  return 0;
}
