// b200_internal.h -- shared declarations of libhypre_b200 (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <atomic>
#include <functional>
#include <string>
#include <vector>
#include "../../include/hypre_b200.h"

#define B200_NUM_SM_FALLBACK 148
// streaming SpMV geometry shared by the row-block plan and both kernels
#define B200_SPMV_NT 128          // threads per CTA
#define B200_SPMV_RCAP 192        // row pointers staged per tile by the pipelined SpMV kernels (tiles of up to RCAP - 32 rows)
#define B200_SPMV_MAX_TILE 1024   // upper bound on the entries of one tile

struct b200_pool_s;   // slab sub-allocator (b200_runtime.cu)
struct b200_prof_s;   // region profiler (b200_runtime.cu)

struct b200_handle_s {
  int device = 0;
  b200_pool_s *pool = nullptr;
  int num_sm = B200_NUM_SM_FALLBACK;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // scratch for reductions (dot products): partial sums + pinned host landing zone
  double *d_partials = nullptr;
  double *h_pinned = nullptr;   // 1024 doubles, pinned
  int     n_partials = 0;
  b200_prof_s *prof = nullptr;
};

// Device CSR block.  Arrays are over-allocated by B200_PAD entries so kernels may issue aligned
// 128-bit loads that straddle the logical end.
#define B200_PAD 8
struct b200_gs_plan_s;   // level schedule of the Gauss-Seidel sweeps (b200_gs.cu)
struct b200_csr_s {
  int nrows = 0, ncols = 0, nnz = 0;
  int *i = nullptr;       // [nrows+1]
  int *j = nullptr;       // [nnz (+pad)]
  double *a = nullptr;    // [nnz (+pad)]  (nullptr for pattern-only matrices such as S)
  bool owns = true;
  // streaming-SpMV plan: block b owns rows [blk_row[b], blk_row[b+1])
  int *blk_row = nullptr;
  int *blk_ent = nullptr; // first entry of each tile: i[blk_row[b]]  [nblk+1]
  int *blk_meta = nullptr; // int4 per tile {row0,row1,ent0,ent1} for the pipelined kernel [4*nblk]
  int  nblk = 0;
  int  group = 1;         // threads cooperating on one row in the reduce phase
  int  max_row = 0;
  int  tile = 0;          // entries per tile the plan was cut with
  // dictionary-compressed solve copy (b200_spmv_dict.cu): one byte per entry instead of the column / the value
  unsigned char *jc = nullptr, *ac = nullptr;   // [nnz (+pad)] codes into off_tab (column - row) / val_tab
  int *off_tab = nullptr;                       // [256]
  double *val_tab = nullptr;                    // [256]
  int  dict_state = 0;    // 0: not looked at, 1: at least one dictionary exists, -1: none applies
  b200_gs_plan_s *gs = nullptr;   // built on first use by b200_relax_gs / the AMG setup
  b200_csr_s *T = nullptr;        // explicit transpose, built on first b200_csr_matvecT (diagT/offdT of the reference)
};

struct b200_halo_s;   // multi-rank halo plan (b200_parcsr.cu)

struct b200_parcsr_s {
  int global_rows = 0, global_cols = 0;
  int first_row = 0, first_col = 0;          // this rank's first global row / col
  b200_csr diag = nullptr;
  b200_csr offd = nullptr;                   // may have 0 cols
  int *col_map_offd = nullptr;               // device, sorted global ids [ncols_offd]
  std::vector<int> h_col_map_offd;
  b200_halo_s *halo = nullptr;
  double *x_ghost = nullptr;                 // [ncols_offd]
};

extern thread_local std::string g_b200_err;
extern std::atomic<long long> g_b200_launches;   // incremented from several rank threads in the threads-as-ranks backend

int b200_set_error(const char *file, int line, const char *msg);
extern unsigned long long *g_b200_p2p_dbg;

#define B200_FAIL(msg) return b200_set_error(__FILE__, __LINE__, (msg))
#define B200_CUDA(call)                                                         \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) return b200_set_error(__FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)
#define B200_TRY(call)                  \
  do {                                  \
    int rc__ = (call);                  \
    if (rc__) return rc__;              \
  } while (0)
#define B200_LAUNCH_CHECK()                                                      \
  do {                                                                           \
    ++g_b200_launches;                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) return b200_set_error(__FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

// B200_PROF=1: CUDA-event timing of labelled regions on the handle's stream, summed per label and printed by
// b200_prof_report (stderr).  Off by default: one predictable branch per region.
struct b200_prof_s;
void b200_prof_begin(b200_handle h, const char *label, int level);
void b200_prof_end(b200_handle h);
void b200_prof_report(b200_handle h, const char *title);
struct b200_prof_scope {
  b200_handle h;
  b200_prof_scope(b200_handle h_, const char *label, int level = -1) : h(h_) { b200_prof_begin(h, label, level); }
  ~b200_prof_scope() { b200_prof_end(h); }
};

// CUDA-graph replay of a fixed kernel sequence on the handle's stream (the PCG iteration: ~100 launches, a third of them on
// levels with a few thousand rows where a launch costs more than the kernel).  begin/end bracket the sequence once, in relaxed
// capture mode; launch replays it.  Off with B200_GRAPH=0 and whenever the region profiler is on.
bool b200_graph_enabled();
int b200_graph_begin(b200_handle h);
int b200_graph_end(b200_handle h, cudaGraphExec_t *exec);      // *exec = nullptr when the capture was invalidated (caller falls back)
int b200_graph_launch(b200_handle h, cudaGraphExec_t exec);
void b200_graph_destroy(cudaGraphExec_t exec);

// device allocation: slab sub-allocator owned by the handle.  All work of a handle runs on one
// stream, so a freed block may be handed out again immediately (stream order keeps it safe).
int b200_pool_alloc(b200_handle h, void **p, size_t bytes);
int b200_pool_free(b200_handle h, void *p);
template <class T>
static inline int b200_dalloc(b200_handle h, T **p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  return b200_pool_alloc(h, (void **)p, count * sizeof(T));
}
static inline int b200_dfree(b200_handle h, void *p) {
  if (p) return b200_pool_free(h, p);
  return 0;
}
static inline int b200_grid(size_t n, int block) { return (int)((n + block - 1) / block); }

// internal cross-file entry points
int b200_csr_alloc(b200_handle h, int nrows, int ncols, int nnz, bool with_data, b200_csr *A);
int b200_csr_build_dict(b200_handle h, b200_csr A);     // b200_spmv_dict.cu; called by b200_csr_build_plan
int b200_csr_drop_dict(b200_handle h, b200_csr A);      // after the values of A changed in place
int b200_csr_build_plan(b200_handle h, b200_csr A);
int b200_exclusive_scan_inplace(b200_handle h, int *d_data, size_t n);   // d_data[n] entries, in place
int b200_gs_plan_create(b200_handle h, b200_csr A, int blocks, b200_gs_plan_s **out);
int b200_gs_plan_destroy(b200_handle h, b200_gs_plan_s *p);
int b200_gs_plan_levels(b200_gs_plan_s *p);
int b200_gs_plan_blocks(b200_gs_plan_s *p);
int b200_gs_relax(b200_handle h, b200_gs_plan_s *P, b200_csr A, int type, bool zero, const double *f, const double *l1, double *u);
// aggressive coarsening building blocks shared by the single-rank and the row-partitioned setup (b200_agg.cu)
struct b200_agg_hooks {
  std::function<int(int *)> sync_int;                        // refresh the ghost tail of an int array indexed like A's columns
  std::function<int(int *)> sum_int;                         // sum one host int over the ranks, in place
  std::function<int(b200_csr, b200_csr *)> with_ghost_rows;  // [rows of the owned nodes ; rows of the ghost nodes]
};
int b200_create_2nd_s_ex(b200_handle h, b200_csr S, int n_owned, const int *d_cf, const int *d_f2c, int first_coarse,
                         int ncoarse, b200_csr *out);
int b200_correct_cf(b200_handle h, int n_owned, const int *d_cfn, int *d_cf);
int b200_multipass_ex(b200_handle h, b200_csr A, b200_csr S, int n, int n_ext, int *d_cf, const int *d_f2c, int ncoarse,
                      const b200_agg_hooks *hooks, b200_csr *out);
// Chebyshev smoother, relax 16 (b200_cheby.cu)
struct b200_cheby_s;
int b200_cheby_setup(b200_handle h, b200_csr A, int eig_est, int order, double fraction, int variant, int scale,
                     b200_cheby_s **out);
int b200_cheby_solve(b200_handle h, b200_cheby_s *C, b200_csr As, bool zero, const double *f, double *u);
int b200_cheby_destroy(b200_handle h, b200_cheby_s *c);
int b200_reduce_sum_int(b200_handle h, const int *d_data, size_t n, long long *h_out);
int b200_vec_dot2_dev(b200_handle h, int n, const double *x, const double *y, double *d_xy, double *d_xx);   // <x,y> and <x,x>, one pass
// GMRES / BiCGSTAB written once over these operations (b200_krylov.cu); single GPU there, row-partitioned in b200_dist.cu
struct b200_krylov_ops {
  int n = 0, cap = 0;          // owned entries; allocation length of a work vector (n + the ghost tail the operator reads)
  bool device_mgs = true;      // inner products complete on this device: Gram-Schmidt coefficients never visit the host
  std::function<int(double, const double *, double, const double *, double *)> matvec;   // y = alpha A x + beta b
  std::function<int(const double *, double *)> precond;                                  // out = M^{-1} rhs from a zero guess
  std::function<int(const double *, int, double *)> reduce;      // k device partial sums -> host, summed over the ranks in order
};
int b200_gmres_core(b200_handle h, const b200_krylov_ops *ops, const b200_gmres_params *prm, const double *d_b, double *d_x,
                    int *iters, double *final_rel_res, double *h_norms, int *converged);
int b200_bicgstab_core(b200_handle h, const b200_krylov_ops *ops, const b200_bicgstab_params *prm, const double *d_b, double *d_x,
                       int *iters, double *final_rel_res, double *h_norms, int *converged);
