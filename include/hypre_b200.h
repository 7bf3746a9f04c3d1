/* hypre_b200.h -- C-ABI of libhypre_b200.so: the B200-native (sm_100a) implementation of
 * BoomerAMG's data-parallel hot path.
 *
 * Plain C, plain pointers and sizes: this is what a host build of hypre binds in place of its
 * own `XDevice` operator seam (boundary B3).  Each entry point cites the reference
 * function (file:line under /root/reference/src) whose behaviour it reproduces.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error (b200_last_error() has the text);
 *    there is NO CPU fallback: if the device or the kernels are unavailable the call fails.
 *  - "d_" pointers are device pointers, "h_" pointers are host pointers.
 *  - all reals are FP64 (HYPRE_Real=double), all indices int32 (HYPRE_Int=HYPRE_BigInt=int),
 *    matching the reference configuration (utilities/HYPRE_utilities.h:48-49).
 *  - all work is enqueued on the handle's CUDA stream; calls that return host scalars synchronise.
 */
#ifndef HYPRE_B200_H
#define HYPRE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_handle_s *b200_handle;
typedef struct b200_csr_s    *b200_csr;     /* device CSR block            (seq_mv/csr_matrix.h:25-56)   */
typedef struct b200_parcsr_s *b200_parcsr;  /* row-partitioned ParCSR      (parcsr_mv/par_csr_matrix.h:27-95) */
typedef struct b200_amg_s    *b200_amg;     /* BoomerAMG hierarchy + parms (parcsr_ls/par_amg.h:18-271)  */
typedef struct b200_comm_s   *b200_comm;    /* communicator: one rank per GPU (MPI_Comm of the reference)  */
typedef struct b200_comm_group_s *b200_comm_group;   /* in-process rank group (test backend)             */
typedef struct b200_dist_matrix_s *b200_dist_matrix; /* row-partitioned operator + halo plan (ParCSR + CommPkg) */
typedef struct b200_dist_amg_s *b200_dist_amg;       /* row-partitioned BoomerAMG hierarchy                */

/* ---- runtime (utilities/hypre_general.c:128 HYPRE_Init, hypre_memory.c) -------------------- */
int         b200_init(int device, b200_handle *h);
int         b200_finalize(b200_handle h);
const char *b200_last_error(void);
void       *b200_stream(b200_handle h);                 /* cudaStream_t */
int         b200_sync(b200_handle h);
int         b200_malloc(b200_handle h, void **d_ptr, size_t bytes);
int         b200_free(b200_handle h, void *d_ptr);
int         b200_memcpy_h2d(b200_handle h, void *d_dst, const void *h_src, size_t bytes);
int         b200_memcpy_d2h(b200_handle h, void *h_dst, const void *d_src, size_t bytes);
int         b200_memcpy_d2d(b200_handle h, void *d_dst, const void *d_src, size_t bytes);
int         b200_memset(b200_handle h, void *d_dst, int byte, size_t bytes);
/* device memory comes from slabs the handle reserves (64 MiB growing to 1 GiB each; hypre_TAlloc(HYPRE_MEMORY_DEVICE),
 * utilities/hypre_memory.c): trim returns the slabs that are entirely free to the driver (synchronises the device),
 * stats reports reserved / in-use / peak bytes */
int         b200_pool_trim(b200_handle h, size_t *bytes_released);
int         b200_pool_stats(b200_handle h, size_t *reserved, size_t *in_use, size_t *peak);
/* number of kernels this library has launched since b200_init (bench.py gpu_launches) */
long long   b200_launch_count(void);
/* device-side elapsed-time helpers on the handle's stream (cudaEvent based) */
int         b200_timer_start(b200_handle h);
int         b200_timer_stop_ms(b200_handle h, double *ms);

/* ---- sequential CSR kernels (seq_mv) ------------------------------------------------------- */
/* wrap/copy device CSR arrays; builds the row-block plan the streaming SpMV uses */
int b200_csr_create(b200_handle h, int nrows, int ncols, int nnz,
                    const int *d_i, const int *d_j, const double *d_a, int copy, b200_csr *A);
int b200_csr_create_from_host(b200_handle h, int nrows, int ncols, int nnz,
                              const int *h_i, const int *h_j, const double *h_a, b200_csr *A);
int b200_csr_destroy(b200_handle h, b200_csr A);
int b200_csr_dims(b200_csr A, int *nrows, int *ncols, int *nnz);
/* bytes per entry the SpMV of A streams: 12 (int32 column + FP64 value), or 9 / 5 / 2 when the dictionary-compressed solve copy
 * of a stencil-structured operator holds one byte in place of the column offset and / or the value (csrc/b200_spmv_dict.cu);
 * same arithmetic as hypre_CSRMatrixMatvecOutOfPlace (seq_mv/csr_matvec.c:24-376), bit-identical results */
int b200_csr_stream_bytes_per_entry(b200_csr A);
int b200_csr_download(b200_handle h, b200_csr A, int *h_i, int *h_j, double *h_a);
/* y = alpha*A*x + beta*b, x != y.   hypre_CSRMatrixMatvecOutOfPlace (seq_mv/csr_matvec.c:24-412),
 * device seam hypre_CSRMatrixMatvecDevice (seq_mv/csr_matvec_device.c:56-120). b may equal y. */
int b200_csr_matvec(b200_handle h, double alpha, b200_csr A, const double *d_x,
                    double beta, const double *d_b, double *d_y);
/* y = alpha*A^T*x + beta*b.   hypre_CSRMatrixMatvecT (seq_mv/csr_matvec.c:424-668), device seam
 * hypre_CSRMatrixMatvecDevice(trans = 1).  The explicit transpose is built on first use and cached on A
 * (the reference's diagT / offdT, par_csr_matvec.c:553-597). */
int b200_csr_matvecT(b200_handle h, double alpha, b200_csr A, const double *d_x,
                     double beta, const double *d_b, double *d_y);
/* the per-matrix numbers of hypre_BoomerAMGSetupStats (parcsr_ls/par_stats.c:575-606, :866-925): entries per row,
 * row sums (each row summed in storage order), and min weight / max weight over the entries != 1 (for P) */
int b200_csr_row_stats(b200_handle h, b200_csr A, int *min_entries, int *max_entries, double *min_rowsum,
                       double *max_rowsum, double *min_weight, double *max_weight);
/* copy of A with the entries of every row sorted by column (hypre_CSRMatrixReorder's job for all columns):
 * the solve phase applies the coarse Galerkin operators from such a copy, see DESIGN.md section 3 */
int b200_csr_sorted_copy(b200_handle h, b200_csr A, b200_csr *S);
/* AT = A^T with rows ordered by source row (stable counting sort semantics).
 * hypre_CSRMatrixTransposeHost (seq_mv/csr_matop.c:578-779) */
int b200_csr_transpose(b200_handle h, b200_csr A, b200_csr *AT);
/* C = A*B, Gustavson row order: columns in first-touch order, values summed left to right,
 * diagonal placed first when C is square.  hypre_CSRMatrixMultiplyHost (seq_mv/csr_matop.c:295-473) */
int b200_csr_multiply(b200_handle h, b200_csr A, b200_csr B, b200_csr *C);

/* ---- vector kernels (seq_mv/vector.c:238,:321,:394,:451,:511) ------------------------------ */
int b200_vec_fill(b200_handle h, int n, double value, double *d_x);           /* SetConstantValues */
int b200_vec_copy(b200_handle h, int n, const double *d_x, double *d_y);      /* Copy   */
int b200_vec_scale(b200_handle h, int n, double alpha, double *d_y);          /* Scale  */
int b200_vec_axpy(b200_handle h, int n, double alpha, const double *d_x, double *d_y); /* Axpy */
int b200_vec_dot(b200_handle h, int n, const double *d_x, const double *d_y, double *h_result); /* InnerProd */

/* ---- problem generators (parcsr_ls/par_laplace.c:15-357, par_laplace_27pt.c:15) ------------- */
/* Row partition (P,Q,R) process grid, this rank at (p,q,r); same entry order as the reference:
 * diagonal first, then z-,y-,x-,x+,y+,z+ (7-pt) resp. the 27-pt lexicographic order. */
int b200_generate_laplacian(b200_handle h, int nx, int ny, int nz, int P, int Q, int R,
                            int p, int q, int r, const double values[4], b200_parcsr *A);
int b200_generate_laplacian27(b200_handle h, int nx, int ny, int nz, int P, int Q, int R,
                              int p, int q, int r, const double values[2], b200_parcsr *A);
/* GenerateDifConv (parcsr_ls/par_difconv.c:15-365): same pattern and entry order as the 7-point Laplacian with
 * seven coefficients values[7] = centre, x-, y-, z-, x+, y+, z+ (ij.c:8270-8420 computes them from -c, -a, -atype);
 * nonsymmetric unless values[k] == values[k+3]. */
int b200_generate_difconv(b200_handle h, int nx, int ny, int nz, int P, int Q, int R,
                          int p, int q, int r, const double values[7], b200_parcsr *A);
/* GenerateRotate7pt (parcsr_ls/par_rotate_7pt.c:15-400): the 2-D rotated anisotropic diffusion operator
 * -(c^2 + eps s^2) u_xx - 2 (1 - eps) s c u_xy - (s^2 + eps c^2) u_yy, c = cos(alpha), s = sin(alpha), alpha in degrees,
 * 7-point stencil with entry order centre, (-1,-1), (0,-1), (-1,0), (+1,0), (0,+1), (+1,+1); P x Q process grid */
int b200_generate_rotate7pt(b200_handle h, int nx, int ny, int P, int Q, int p, int q, double alpha, double eps,
                            b200_parcsr *A);

/* ---- IJ assembly on the device (IJ_mv/IJMatrix_parcsr.c:697-1186 SetValues, :1188 AddToValues, :2774-3080 Assemble) ----
 * SetValues / AddToValues take HOST arrays (global row / column indices) and only append validated records to a pinned
 * log that streams to the device; Assemble sorts the log by row (stable), replays every row's insertions on the device
 * with the reference's rules (an entry is matched only against what its row held before the current (call,row) pair;
 * unmatched entries are appended; the last entry on the diagonal column moves to the front) and returns the ParCSR
 * object.  After the first Assemble only existing entries may be set / added; *n_missing counts the others. */
typedef struct b200_ij_s *b200_ij;
int b200_ij_create(b200_handle h, int ilower, int iupper, int jlower, int jupper, b200_ij *ij);
int b200_ij_destroy(b200_handle h, b200_ij ij);
int b200_ij_set_values(b200_handle h, b200_ij ij, int nrows, const int *ncols, const int *rows, const int *cols,
                       const double *values, int add, int *n_rejected);
int b200_ij_assemble(b200_handle h, b200_ij ij, b200_parcsr *A, int *n_missing);
long long b200_ij_num_rejected(b200_ij ij);
/* assembler for the rows [ilower, iupper] of a square operator spread over several ranks: columns stay global
 * (0 .. global_cols-1); hand it to b200_dist_matrix_create_from_ij */
int b200_ij_create_rows(b200_handle h, int ilower, int iupper, int global_cols, b200_ij *ij);

/* ---- ParCSR (parcsr_mv) -------------------------------------------------------------------- */
/* single-rank ParCSR from a host CSR (diag block = whole matrix); diagonal entry must be first
 * in each row as in the reference's diag block (csr_matrix.h, relied on by relax/strength). */
int b200_parcsr_create_from_host(b200_handle h, int nrows, int ncols, int nnz,
                                 const int *h_i, const int *h_j, const double *h_a, b200_parcsr *A);
int b200_parcsr_destroy(b200_handle h, b200_parcsr A);
int b200_parcsr_local_rows(b200_parcsr A, int *nrows, int *nnz_diag, int *nnz_offd, int *ncols_offd);
b200_csr b200_parcsr_diag(b200_parcsr A);
b200_csr b200_parcsr_offd(b200_parcsr A);
/* y = alpha*A*x + beta*b over diag+offd with halo exchange of x.
 * hypre_ParCSRMatrixMatvecOutOfPlace (parcsr_mv/par_csr_matvec.c:22-359) */
int b200_parcsr_matvec(b200_handle h, double alpha, b200_parcsr A, const double *d_x,
                       double beta, const double *d_b, double *d_y);

/* ---- BoomerAMG (parcsr_ls) ------------------------------------------------------------------ */
int b200_amg_create(b200_amg *amg);                      /* HYPRE_BoomerAMGCreate  (HYPRE_parcsr_amg.c:15)  */
int b200_amg_destroy(b200_handle h, b200_amg amg);       /* HYPRE_BoomerAMGDestroy (:32) */
/* integer / real parameters by the reference setter's name without the HYPRE_BoomerAMGSet prefix:
 * "CoarsenType","InterpType","PMaxElmts","RelaxType","MaxLevels","MaxCoarseSize","NumSweeps",
 * "AggNumLevels","ModuleRAP2","KeepTranspose","RelaxOrder","MaxIter","MinIter","RelaxTypeUp" (up-cycle
 * smoother when it differs from "RelaxType", e.g. 13 down / 14 up), "GSBlocks" (Gauss-Seidel blocks per rank),
 * "ChebyOrder" (1-4, default 2), "ChebyEigEst" (CG steps, default 10), "ChebyVariant" (0), "ChebyScale" (1)
 * for RelaxType 16 (par_cheby.c:41-345; real parameter "ChebyFraction", default 0.3), "CycleType" (1 = V, 2 = W),
 * "FCycle", "NumSweeps" (down = up), "NumSweepsDown" / "NumSweepsUp" (-1 = follow NumSweeps), "NumSweepsCoarse"
 * (num_grid_sweeps[1..3], par_amg.c:1934-2030; cycle control par_cycle.c:180-622) / "StrongThreshold",
 * "MaxRowSum","TruncFactor","RelaxWt","Tol".  Unsupported values are rejected at setup. */
int b200_amg_set_int(b200_amg amg, const char *name, int value);
int b200_amg_set_real(b200_amg amg, const char *name, double value);
/* hypre_BoomerAMGSetup (par_amg_setup.c:27-3518) for the in-scope configuration */
int b200_amg_setup(b200_handle h, b200_amg amg, b200_parcsr A);
/* one application of the preconditioner: hypre_BoomerAMGSolve with MaxIter=1, Tol=0
 * (par_amg_solve.c:21) = one hypre_BoomerAMGCycle (par_cycle.c:22-641); u is overwritten */
int b200_amg_solve(b200_handle h, b200_amg amg, const double *d_f, double *d_u);
/* BoomerAMG as a solver: cycles until ||f - A u||_2 / ||f||_2 < "Tol" or "MaxIter" cycles
 * (par_amg_solve.c:21-380, converge_type 0).  Returns 256 (HYPRE_ERROR_CONV) when MaxIter was reached
 * with Tol > 0, as the reference flags it; u then holds the last iterate. */
int b200_amg_solve_ex(b200_handle h, b200_amg amg, b200_parcsr A, const double *d_f, double *d_u,
                      int *num_iterations, double *final_rel_res);
int b200_amg_num_levels(b200_amg amg);
/* hierarchy access for parity tests (device objects owned by amg) */
b200_csr b200_amg_level_A(b200_amg amg, int level);
b200_csr b200_amg_level_P(b200_amg amg, int level);
b200_csr b200_amg_level_S(b200_amg amg, int level);      /* kept only if "KeepS" = 1 */
const int    *b200_amg_level_CF(b200_amg amg, int level); /* device int[nrows_l]  */
const double *b200_amg_level_l1(b200_amg amg, int level); /* device double[nrows_l] */
/* per-phase device times of the last setup, ms: strength, pmis, interp, trunc, transpose, rap, other */
int b200_amg_setup_times(b200_amg amg, double times[8]);

/* individual setup stages, exposed so each can be parity-tested against the reference */
/* hypre_BoomerAMGCreateSHost (par_strength.c:80-530): S has no diagonal, A's column order */
int b200_strength(b200_handle h, b200_csr A, double theta, double max_row_sum, b200_csr *S);
/* hypre_BoomerAMGCoarsenPMISHost (par_coarsen.c:2031-2738) + IndepSetInit (par_indepset.c:32-63)
 * + hypre_Rand (utilities/random.c:49-106); writes CF marker {1,-1,-3} */
int b200_pmis(b200_handle h, b200_csr S, int seed, int *d_cf);
/* hypre_BoomerAMGCoarsenHMIS (par_coarsen.c:2774-2797; coarsen_type 10, the library and driver default): the
 * Ruge-Stueben first pass (:1046-1330, sequential by definition -- one device thread walks the reference's measure lists)
 * followed by PMIS seeded with its C points (CF_init 1).  d_cf receives 1 (C) / -1 (F) / -3 (isolated). */
int b200_hmis(b200_handle h, b200_csr S, int seed, int *d_cf);
/* hypre_BoomerAMGBuildExtPIInterpHost (par_lr_interp.c:1040-1925) followed by
 * hypre_BoomerAMGInterpTruncation (par_interp.c:2718 -> par_csr_matrix.c:2671-3060) */
int b200_extpi_interp(b200_handle h, b200_csr A, b200_csr S, const int *d_cf,
                      double trunc_factor, int max_elmts, b200_csr *P);
/* aggressive coarsening (par_amg_setup.c:1239-1256, :1590-1605; AMG parameter "AggNumLevels"):
 * hypre_BoomerAMGCreate2ndS, num_paths 1 (par_strength.c:1729-2918): S2 = distance-two strength graph on the C points
 * of d_cf, pattern only, first-touch column order */
int b200_create_2nd_s(b200_handle h, b200_csr S, const int *d_cf, b200_csr *S2);
/* second coarsening: PMIS on S2 with CF_init 3 (par_coarsen.c:2322-2326, :2420) + hypre_BoomerAMGCorrectCFMarker
 * (par_strength.c:2957-2974); d_cf in: marker of the first PMIS, out: corrected marker */
int b200_agg_coarsen(b200_handle h, b200_csr S, int seed, int *d_cf);
/* hypre_BoomerAMGBuildMultipass (par_multi_interp.c:16-2061), weight_option 0, no truncation; SF points (-3) in d_cf
 * are folded into F on return as the reference does */
int b200_multipass_interp(b200_handle h, b200_csr A, b200_csr S, int *d_cf, b200_csr *P);
/* hypre_ParCSRComputeL1Norms (ams.c:571-760), options 1 and 4 */
int b200_l1_norms(b200_handle h, b200_csr A, int option, double *d_l1);
/* the same with the reference's thread blocks (hypre_ParCSRComputeL1NormsThreads, ams.c:3398-3650): option 4
 * adds half of every entry that lies outside the Gauss-Seidel block of its row */
int b200_l1_norms_blocks(b200_handle h, b200_csr A, int option, int blocks, double *d_l1);
/* option 1 with a C/F marker: only the entries whose column carries the row's own marker are summed -- the l1 norms of a
 * cycle that relaxes in C/F order (relax_order 1: par_amg_setup.c:3047-3050, hypre_CSRMatrixComputeRowSum csr_matop.c:1326-1352) */
int b200_l1_norms_cf(b200_handle h, b200_csr A, const int *d_cf, double *d_l1);

/* hypre_BoomerAMGRelax (par_relax.c:30-5264) for the hybrid Gauss-Seidel family on one rank, relax_points 0,
 * relax_weight = omega = 1: types 3 / 4 / 6 (forward / backward / symmetric, u_i = res / a_ii) and the l1
 * variants 13 / 14 / 8 (u_i += res / l1_i, d_l1 from b200_l1_norms option 4).  `blocks` = number of
 * Gauss-Seidel blocks the rows are cut into (the reference's OpenMP thread count, par_relax.c:4400-4412):
 * sequential inside a block, pre-sweep values across blocks.  For a given block count the result is the
 * reference loop's bit for bit.  (AMG parameter: "GSBlocks".) */
int b200_relax_gs(b200_handle h, b200_csr A, int relax_type, int blocks, const double *d_f, const double *d_l1,
                  double *d_u);

/* ---- PCG (krylov/pcg.c:271-757 via HYPRE_ParCSRPCGSolve) ------------------------------------- */
/* Solves A x = b with BoomerAMG-preconditioned CG (amg may be NULL: unpreconditioned).
 * two_norm=1 convergence test ||r||_2/||b||_2 < tol as ij.c:3894 sets.  h_norms (may be NULL)
 * receives ||r_k||_2 for k=0..iters (needs max_iter+1 doubles). */
int b200_pcg_solve(b200_handle h, b200_parcsr A, b200_amg amg, const double *d_b, double *d_x,
                   double tol, int max_iter, int *iters, double *final_rel_res, double *h_norms);
/* the same solver with the hypre_PCGData fields that matter on this path (krylov/pcg.h:190-230):
 * two_norm 0 = energy norm <C r, r>; precond 0 none, 1 BoomerAMG (amg), 2 HYPRE_ParCSRDiagScale */
typedef struct {
  double tol, a_tol;
  int max_iter, two_norm, rel_change, recompute_residual, precond;
} b200_pcg_params;
int b200_pcg_solve_ex(b200_handle h, b200_parcsr A, b200_amg amg, const b200_pcg_params *params, const double *d_b,
                      double *d_x, int *iters, double *final_rel_res, double *h_norms);
/* x = y ./ diag(A)   HYPRE_ParCSRDiagScale (parcsr_ls/HYPRE_parcsr_pcg.c:228-258) */
int b200_parcsr_diag_scale(b200_handle h, b200_parcsr A, const double *d_y, double *d_x);

/* ---- GMRES / BiCGSTAB (krylov/gmres.c:226-800, krylov/bicgstab.c:207-530; ij -solver 3 / -solver 9) ----------
 * The Krylov drivers for nonsymmetric operators (GenerateDifConv) over the same SpMV / BLAS-1 kernels as PCG.
 * Fields = the hypre_GMRESData / hypre_BiCGSTABData members that matter here (gmres.h:84-111, bicgstab.h);
 * precond 0 none, 1 BoomerAMG (one cycle from a zero guess, right preconditioning), 2 HYPRE_ParCSRDiagScale.
 * rel_change and cf_tol > 0 are rejected.  h_norms (may be NULL) receives the residual-norm history norms[0..iters]
 * (max_iter + 1 doubles); *converged mirrors the `converged` member. */
typedef struct {
  double tol, a_tol, cf_tol;
  int max_iter, min_iter, k_dim, rel_change, skip_real_r_check, precond;
} b200_gmres_params;
int b200_gmres_solve(b200_handle h, b200_parcsr A, b200_amg amg, const b200_gmres_params *params, const double *d_b,
                     double *d_x, int *iters, double *final_rel_res, double *h_norms, int *converged);
typedef struct {
  double tol, a_tol, cf_tol;
  int max_iter, min_iter, stop_crit, precond;
} b200_bicgstab_params;
int b200_bicgstab_solve(b200_handle h, b200_parcsr A, b200_amg amg, const b200_bicgstab_params *params, const double *d_b,
                        double *d_x, int *iters, double *final_rel_res, double *h_norms, int *converged);

/* ---- multi-GPU: row-partitioned ParCSR over NVLink ---------------------------------
 * One process per GPU.  Rows are partitioned contiguously like hypre's ParCSR layout
 * (par_csr_matrix.h:27-95); each rank stores its rows as ONE CSR whose columns are the owned
 * range followed by the ghost columns (sorted by global id), so a SpMV is a halo exchange
 * (hypre_ParCSRCommHandleCreate job 1, par_csr_communication.c:307-580) + one kernel over
 * [x_owned | x_ghost].  Setup runs the same per-row algorithms as the single-GPU path on rows
 * fetched from their owners, in the global row's entry order, with PMIS measures drawn from the
 * global row index (the reference's partition-independent `-pmis1`, par_indepset.c:44-55): the
 * hierarchy is therefore bit-identical for every number of GPUs.                                    */
int b200_comm_create_single(b200_comm *c);
int b200_comm_group_create(int nranks, b200_comm_group *g);     /* N ranks = N host threads, one GPU   */
int b200_comm_group_destroy(b200_comm_group g);
/* a rank that failed calls this: peers blocked in (or arriving at) an exchange return an error instead of waiting
 * (MPI_Abort of the reference's hypre_error handler, utilities/hypre_error.c); a peer that does not arrive within
 * B200_COMM_TIMEOUT_S seconds (default 300) aborts the group the same way */
int b200_comm_group_abort(b200_comm_group g);
int b200_comm_create_threads(b200_comm_group g, int rank, b200_comm *c);
/* rank threads: switch the direct peer-to-peer path on (flags and receive buffers of the other rank threads are plain
 * pointers on the same device); collective over the group.  The NCCL communicator enables it by itself (CUDA IPC over
 * NVLink) unless B200_P2P=0.  Test hook for the protocol on a single-GPU box. */
int b200_comm_threads_enable_p2p(b200_handle h, b200_comm c);
int b200_comm_nccl_unique_id(char *id128);                       /* rank 0, then broadcast by the host  */
int b200_comm_create_nccl(b200_handle h, int nranks, int rank, const char *id128, b200_comm *c);
int b200_comm_destroy(b200_handle h, b200_comm c);
/* tear the communicator down after a local failure without waiting for the peers (ncclCommAbort / group abort) */
int b200_comm_abort(b200_comm c);
int b200_comm_rank(b200_comm c);
int b200_comm_size(b200_comm c);

/* GenerateLaplacian / GenerateLaplacian27pt on a P x Q x R process grid, rank -> (p,q,r) as ij.c:7785-7787 */
int b200_dist_generate_laplacian(b200_handle h, b200_comm c, int nx, int ny, int nz, int P, int Q, int R,
                                 int stencil, const double *values, b200_dist_matrix *A);
/* GenerateDifConv on the process grid (par_difconv.c:15): values[7] = centre, x-, y-, z-, x+, y+, z+ */
int b200_dist_generate_difconv(b200_handle h, b200_comm c, int nx, int ny, int nz, int P, int Q, int R,
                               const double values[7], b200_dist_matrix *A);
/* GenerateRotate7pt on a P x Q process grid, rank -> (p, q) as ij.c:9190-9191 */
int b200_dist_generate_rotate7pt(b200_handle h, b200_comm c, int nx, int ny, int P, int Q, double alpha, double eps,
                                 b200_dist_matrix *A);
/* a row-partitioned operator from the caller's own rows (what hypre_IJMatrixAssembleParCSR + GenerateDiagAndOffd give
 * across ranks, IJ_mv/IJMatrix_parcsr.c:2774, parcsr_mv/par_csr_matrix.c:1634): rank r owns the next n_local rows after
 * those of the ranks before it; rows carry GLOBAL column ids and their diagonal entry first.  Collective.
 * _from_host: host CSR arrays;  _from_ij: the records of a b200_ij_create_rows assembler, merged on the device. */
int b200_dist_matrix_create_from_host(b200_handle h, b200_comm c, int n_local, const int *h_i, const int *h_j_global,
                                      const double *h_a, b200_dist_matrix *A);
int b200_dist_matrix_create_from_ij(b200_handle h, b200_comm c, b200_ij ij, b200_dist_matrix *A);
int b200_dist_matrix_destroy(b200_handle h, b200_dist_matrix A);
/* b200_csr_stream_bytes_per_entry of this rank's localized block (the operand of hypre_ParCSRMatrixMatvec, par_csr_matvec.c:22-359) */
int b200_dist_matrix_stream_bytes_per_entry(b200_dist_matrix A);
int b200_dist_matrix_info(b200_dist_matrix A, int *local_rows, int *first_row, int *global_rows, int *local_nnz,
                          int *n_ghost, int *first_col, int *global_cols);
/* local rows with GLOBAL column ids, entry order preserved (for parity tests) */
int b200_dist_matrix_download(b200_handle h, b200_dist_matrix A, int *h_i, int *h_j_global, double *h_a);
/* y = alpha*A*x + beta*b; d_x must have room for local_rows + n_ghost doubles (ghosts are received in place) */
int b200_dist_matvec(b200_handle h, b200_comm c, double alpha, b200_dist_matrix A, double *d_x, double beta,
                     const double *d_b, double *d_y);
/* hypre_BoomerAMGRelax types 8 / 13 / 14 across ranks (par_relax.c:4340-5124): halo of u, then Gauss-Seidel
 * inside each of the rank's `blocks` blocks with option-4 l1 norms (ams.c:3560-3625); d_u has room for
 * local_rows + n_ghost doubles */
int b200_dist_relax_gs(b200_handle h, b200_comm c, b200_dist_matrix A, int relax_type, int blocks, const double *d_f,
                       double *d_u);
/* hypre_BoomerAMGSetup across ranks; parameters are taken from a b200_amg object */
int b200_dist_amg_setup(b200_handle h, b200_comm c, b200_amg params, b200_dist_matrix A, b200_dist_amg *amg);
int b200_dist_amg_destroy(b200_handle h, b200_dist_amg amg);
int b200_dist_amg_num_levels(b200_dist_amg amg);
b200_dist_matrix b200_dist_amg_level_A(b200_dist_amg amg, int level);
b200_dist_matrix b200_dist_amg_level_P(b200_dist_amg amg, int level);
int b200_dist_amg_level_cf(b200_handle h, b200_dist_amg amg, int level, int *h_cf);
/* A_l (what 0) or P_l (what 1) of ANY level as this rank's block of rows with global column ids.  With SeqThreshold > 0
 * (HYPRE_BoomerAMGSetSeqThreshold, par_amg.c; par_amg_setup.c:2880-2898) the levels whose global size is at most the
 * threshold are gathered onto every rank and built / cycled redundantly as one single-GPU hierarchy; those levels are
 * reported through views: the first replicated level in the partition of the distributed level it came from, deeper ones
 * whole on rank 0 and empty elsewhere.  Views belong to the hierarchy. */
int b200_dist_amg_level_view(b200_handle h, b200_comm c, b200_dist_amg amg, int level, int what, b200_dist_matrix *M);
int b200_dist_amg_setup_ms(b200_dist_amg amg, double *ms);
/* hypre_PCGSolve across ranks (dot products = deterministic rank-ordered sums) */
int b200_dist_pcg_solve(b200_handle h, b200_comm c, b200_dist_matrix A, b200_dist_amg amg, const double *d_b,
                        double *d_x, double tol, int max_iter, int *iters, double *final_rel_res, double *h_norms);
/* hypre_GMRESSolve / hypre_BiCGSTABSolve across ranks (krylov/gmres.c:226, bicgstab.c:207 over the hypre_ParKrylov*
 * callbacks of parcsr_ls/par_krylov_func.c): same loops as b200_gmres_solve / b200_bicgstab_solve, every inner product
 * one rank-ordered reduction; precond 0 none, 1 the distributed BoomerAMG cycle */
int b200_dist_gmres_solve(b200_handle h, b200_comm c, b200_dist_matrix A, b200_dist_amg amg, const b200_gmres_params *params,
                          const double *d_b, double *d_x, int *iters, double *final_rel_res, double *h_norms, int *converged);
int b200_dist_bicgstab_solve(b200_handle h, b200_comm c, b200_dist_matrix A, b200_dist_amg amg,
                             const b200_bicgstab_params *params, const double *d_b, double *d_x, int *iters,
                             double *final_rel_res, double *h_norms, int *converged);

#ifdef __cplusplus
}
#endif
#endif /* HYPRE_B200_H */
