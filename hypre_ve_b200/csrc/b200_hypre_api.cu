// b200_hypre_api.cu -- the reference's public C API (boundary B1 of SURVEY.md 8b) for the hot path,
// implemented on top of the b200_* C-ABI.  Host code only: every numerical step is a b200_* call into
// the CUDA kernels of this library, there is no CPU arithmetic path.
//
// Declarations: include/HYPRE_b200.h (each group cites the reference file it mirrors).
// Behaviour kept from the reference:
//   * every function returns the global accumulated error flag (utilities/hypre_error.c:17-40);
//     argument errors set HYPRE_ERROR_ARG | arg<<3 (hypre_error.h:32);
//   * handles are opaque pointers; HYPRE_PCGSetPrecond takes two function pointers + a solver
//     (krylov/HYPRE_pcg.c:260) and HYPRE_PCGSetup calls precond_setup(precond_solver, A, b, x)
//     (krylov/pcg.c:214-262);
//   * IJ matrices assembled through the auxiliary-row path keep the diagonal first and the other
//     entries in insertion order (IJ_mv/IJMatrix_parcsr.c:2925-2952); SetValues overwrites an
//     existing column, AddToValues accumulates (:697-1213, :1215-1745);
//   * BoomerAMG setters only store; what the B200 path cannot do is rejected at Setup with
//     HYPRE_ERROR_GENERIC and a message on stderr (the reference would run it on the CPU).
// Single rank per process here (MPI_Comm is the sequential-stub int); the row-partitioned
// multi-GPU path is reached through b200_dist_* (INTEGRATION.md shows the MPI-build binding).
#include "b200_internal.h"
#include "../../include/HYPRE_b200.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <utility>
#include <vector>

int b200_amg_get_int(b200_amg a, const char *name);

namespace {

HYPRE_Int g_error_flag = 0;             // hypre__global_error (utilities/hypre_error.c:17)
b200_handle g_handle = nullptr;

HYPRE_Int err(HYPRE_Int code) { g_error_flag |= code; return g_error_flag; }
HYPRE_Int err_arg(int k) { return err(HYPRE_ERROR_ARG | (k << 3)); }
HYPRE_Int err_b200(const char *where) {
  fprintf(stderr, "hypre_b200: %s: %s\n", where, b200_last_error());
  return err(HYPRE_ERROR_GENERIC);
}
b200_handle handle() {
  if (!g_handle) {
    const char *dev = getenv("HYPRE_B200_DEVICE");
    if (b200_init(dev ? atoi(dev) : 0, &g_handle)) {
      fprintf(stderr, "hypre_b200: no usable CUDA device: %s\n", b200_last_error());
      g_handle = nullptr;
      err(HYPRE_ERROR_GENERIC);
    }
  }
  return g_handle;
}
#define NEED_HANDLE() b200_handle h = handle(); if (!h) return g_error_flag
#define CALL(expr, where) do { if ((expr)) return err_b200(where); } while (0)

enum SolverKind { KIND_AMG = 0x414d47, KIND_PCG = 0x504347, KIND_GMRES = 0x474d52, KIND_BICGSTAB = 0x424943 };

}  // namespace

struct hypre_ParCSRMatrix_struct {
  b200_parcsr A = nullptr;
  int global_rows = 0, global_cols = 0;
  long long nnz = 0;
  bool owned = true;
};
struct hypre_ParVector_struct {
  double *d = nullptr;
  int n = 0;
  bool initialized = false;
};
struct hypre_IJMatrix_struct {
  int ilower = 0, iupper = -1, jlower = 0, jupper = -1;
  int object_type = -1;
  bool initialized = false, assembled = false;
  b200_ij ij = nullptr;                                       // device-side assembler (b200_ij.cu): record log + per-row replay
  hypre_ParCSRMatrix_struct *object = nullptr;
};
struct hypre_IJVector_struct {
  int jlower = 0, jupper = -1;
  int object_type = -1;
  bool initialized = false, dirty = false, assembled_once = false;
  std::vector<double> host;
  hypre_ParVector_struct *object = nullptr;
};
struct hypre_Solver_struct {
  int kind = 0;
  // BoomerAMG
  b200_amg amg = nullptr;
  std::map<std::string, double> stored;        // every setter's last value (inert ones included)
  int num_iterations = 0;
  double rel_res = 0.0;
  // PCG (krylov/pcg.h:190-230 defaults from hypre_PCGCreate, pcg.c:60-104)
  double tol = 1e-6, a_tol = 0.0;
  int max_iter = 1000, two_norm = 0, rel_change = 0, recompute_residual = 0, print_level = 0, logging = 0;
  HYPRE_PtrToSolverFcn precond = nullptr, precond_setup = nullptr;
  HYPRE_Solver precond_solver = nullptr;
  std::vector<double> norms;
  double setup_s = 0.0, solve_s = 0.0;
  // GMRES / BiCGSTAB (hypre_GMRESCreate gmres.c:60-100, hypre_BiCGSTABCreate bicgstab.c:55-95)
  int k_dim = 5, min_iter = 0, stop_crit = 0, skip_real_r_check = 0, converged = 0;
  double cf_tol = 0.0;
};

namespace {

bool is_amg(HYPRE_Solver s) { return s && s->kind == KIND_AMG; }
bool is_pcg(HYPRE_Solver s) { return s && s->kind == KIND_PCG; }
bool is_gmres(HYPRE_Solver s) { return s && s->kind == KIND_GMRES; }
bool is_bicgstab(HYPRE_Solver s) { return s && s->kind == KIND_BICGSTAB; }

struct Neutral { const char *name; double value; const char *what; };
// parameters that switch on something outside the B200 path: only the neutral value is accepted
const Neutral kNeutral[] = {
    {"PostInterpType", 0, "Jacobi interpolation"},     {"SmoothNumLevels", 0, "complex smoothers (Schwarz/Pilut/ParaSails/Euclid)"},
    {"Additive", -1, "additive cycles"},               {"MultAdditive", -1, "mult-additive cycles"},
    {"Simple", -1, "simple additive cycles"},          {"Nodal", 0, "nodal systems coarsening"},
    {"NonGalerkinTol", 0, "non-Galerkin coarse operators"},
    {"NumCPoints", 0, "user C-points"},                {"NumFPoints", 0, "user F-points"},
    {"NumIsolatedFPoints", 0, "isolated F-points"},    {"NumInterpVectors", 0, "interpolation vectors (GSMG/RBM)"},
    {"GSMG", 0, "GSMG"},                               {"CoarsenCutFactor", 0, "coarsening cut factor"},
    {"Redundant", 0, "redundant coarse solves"},       {"SeqThreshold", 0, "sequential coarse AMG"},
    {"ConvergeType", 0, "convergence on the relative residual change"},   {"Restriction", 0, "AIR restriction"},
    {"ADropTol", 0, "dropping in the coarse operators"},                  {"InterpVecVariant", 0, "interpolation vectors (GSMG/RBM)"},
    {"LevelNonGalerkinTolSet", 0, "non-Galerkin coarse operators"},       {"GridRelaxPointsSet", 0, "user relaxation point lists"},
};

int ndigits(long long number) {            // utilities/hypre_printf.c:115
  int d = 0;
  while (number) { number /= 10; d++; }
  return d;
}
// the two tables and the complexity lines of hypre_BoomerAMGSetupStats (parcsr_ls/par_stats.c:22-1060), one rank
int setup_stats(b200_handle h, HYPRE_Solver s) {
  b200_amg amg = s->amg;
  auto &st = s->stored;
  const int nl = b200_amg_num_levels(amg);
  std::vector<long long> rows(nl), nnz(nl);
  for (int l = 0; l < nl; l++) {
    int nr = 0, nc = 0, nz = 0;
    b200_csr_dims(b200_amg_level_A(amg, l), &nr, &nc, &nz);
    rows[l] = nr; nnz[l] = nz;
  }
  printf("\n\n Num MPI tasks = 1\n\n Num OpenMP threads = 1\n\n\nBoomerAMG SETUP PARAMETERS:\n\n");
  printf(" Max levels = %d\n Num levels = %d\n\n", (int)st["MaxLevels"], nl);
  printf(" Strength Threshold = %f\n", st["StrongThreshold"]);
  printf(" Interpolation Truncation Factor = %f\n", st["TruncFactor"]);
  printf(" Maximum Row Sum Threshold for Dependency Weakening = %f\n\n", st["MaxRowSum"]);
  printf(" Coarsening Type = PMIS \n");
  const int agg = (int)st["AggNumLevels"];
  if (agg > 0) printf("\n No. of levels of aggressive coarsening: %d\n\n Interpolation on agg. levels= multipass interpolation\n", agg);
  printf(" measures are determined locally\n\n\n No global partition option chosen.\n\n");
  printf(" Interpolation = extended+i interpolation\n");
  printf("\nOperator Matrix Information:\n\n");
  int nd[4];
  nd[0] = std::max(7, ndigits(rows[0])); nd[1] = std::max(8, ndigits(nnz[0])); nd[2] = 4;
  for (int l = 0; l < nl; l++) nd[2] = std::max(ndigits(rows[l] ? nnz[l] / rows[l] : 0), nd[2]);
  nd[2] += 2; nd[3] = nd[0] + nd[1] + nd[2];
  printf("%*s", nd[0] + 13, "nonzero"); printf("%*s", nd[1] + 15, "entries/row"); printf("%18s\n", "row sums");
  printf("%s %*s ", "lev", nd[0], "rows"); printf("%*s", nd[1], "entries");
  printf("%7s %5s %4s", "sparse", "min", "max"); printf("%*s %8s %11s\n", nd[2] + 2, "avg", "min", "max");
  for (int i = 0; i < 49 + nd[3]; i++) printf("=");
  printf("\n");
  double num_mem = 0, num_coeffs = 0, num_vars = 0;
  for (int l = 0; l < nl; l++) {
    int mn = 0, mx = 0;
    double r0 = 0, r1 = 0;
    CALL(b200_csr_row_stats(h, b200_amg_level_A(amg, l), &mn, &mx, &r0, &r1, nullptr, nullptr), "setup stats");
    const double sparse = (double)nnz[l] / ((double)rows[l] * (double)rows[l]);
    printf("%3d %*lld %*.0f  %0.3f  %4d %4d", l, nd[0], rows[l], nd[1], (double)nnz[l], sparse, mn, mx);
    printf("  %*.1f  %10.3e  %10.3e\n", nd[2], (double)nnz[l] / (double)rows[l], r0, r1);
    num_mem += (double)nnz[l]; num_coeffs += (double)nnz[l]; num_vars += (double)rows[l];
  }
  nd[0] = 5;                                                       // par_stats.c:698-709
  if (nl > 1) nd[0] = std::max(ndigits(rows[0]), nd[0]);
  printf("\n\nInterpolation Matrix Information:\n");
  printf("%*s ", 2 * nd[0] + 21, "entries/row"); printf("%10s %10s %19s\n", "min", "max", "row sums");
  printf("lev %*s x %-*s min  max  avgW", nd[0], "rows", nd[0], "cols"); printf("%11s %11s %9s %11s\n", "weight", "weight", "min", "max");
  for (int i = 0; i < 70 + 2 * nd[0]; i++) printf("=");
  printf("\n");
  for (int l = 0; l + 1 < nl; l++) {
    b200_csr P = b200_amg_level_P(amg, l);
    int nr = 0, nc = 0, nz = 0, mn = 0, mx = 0;
    double r0 = 0, r1 = 0, w0 = 0, w1 = 0;
    b200_csr_dims(P, &nr, &nc, &nz);
    CALL(b200_csr_row_stats(h, P, &mn, &mx, &r0, &r1, &w0, &w1), "setup stats");
    printf("%3d %*d x %-*d %3d  %3d", l, nd[0], nr, nd[0], nc, mn, mx);
    printf("  %4.1f  %10.3e  %10.3e  %10.3e  %10.3e\n", (double)(nz - nc) / (double)(nr - nc), w0, w1, r0, r1);
    num_mem += (double)nz;
  }
  printf("\n\n     Complexity:    grid = %f\n", num_vars / (double)rows[0]);
  printf("                operator = %f\n", num_coeffs / (double)nnz[0]);
  printf("                memory = %f\n\n", num_mem / (double)nnz[0]);
  return 0;
}

}  // namespace

extern "C" {

// ---- utilities ------------------------------------------------------------------------------
HYPRE_Int HYPRE_Init(void) { handle(); return g_error_flag; }
HYPRE_Int HYPRE_Finalize(void) {
  if (g_handle) { b200_finalize(g_handle); g_handle = nullptr; }
  return g_error_flag;
}
HYPRE_Int HYPRE_GetError(void) { return g_error_flag; }
HYPRE_Int HYPRE_ClearAllErrors(void) { g_error_flag = 0; return 0; }
HYPRE_Int HYPRE_ClearError(HYPRE_Int code) { g_error_flag &= ~code; return g_error_flag; }
HYPRE_Int HYPRE_CheckError(HYPRE_Int ierr, HYPRE_Int code) { return ierr & code; }
HYPRE_Int HYPRE_GetErrorArg(void) { return (g_error_flag >> 3) & 31; }
void HYPRE_DescribeError(HYPRE_Int ierr, char *msg) {         // utilities/hypre_error.c:55-80
  if (ierr == 0) sprintf(msg, "[No error] ");
  if (ierr & HYPRE_ERROR_GENERIC) sprintf(msg, "[Generic error] ");
  if (ierr & HYPRE_ERROR_MEMORY) sprintf(msg, "[Memory error] ");
  if (ierr & HYPRE_ERROR_ARG) sprintf(msg, "[Error in argument %d] ", (ierr >> 3) & 31);
  if (ierr & HYPRE_ERROR_CONV) sprintf(msg, "[Method did not converge] ");
}

// ---- IJ matrix ------------------------------------------------------------------------------
HYPRE_Int HYPRE_IJMatrixCreate(MPI_Comm, HYPRE_BigInt ilower, HYPRE_BigInt iupper, HYPRE_BigInt jlower, HYPRE_BigInt jupper,
                               HYPRE_IJMatrix *matrix) {
  if (ilower > iupper + 1 || ilower < 0) return err_arg(2);          // HYPRE_IJMatrix.c:57-73
  if (iupper < -1) return err_arg(3);
  if (jlower > jupper + 1 || jlower < 0) return err_arg(4);
  if (jupper < -1) return err_arg(5);
  if (!matrix) return err_arg(6);
  hypre_IJMatrix_struct *m = new hypre_IJMatrix_struct();
  m->ilower = ilower; m->iupper = iupper; m->jlower = jlower; m->jupper = jupper;
  *matrix = m;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRMatrixDestroy(HYPRE_ParCSRMatrix A);
HYPRE_Int HYPRE_IJMatrixDestroy(HYPRE_IJMatrix m) {
  if (!m) return err_arg(1);
  if (m->ij) { b200_handle h = handle(); if (h) b200_ij_destroy(h, m->ij); }
  if (m->object) HYPRE_ParCSRMatrixDestroy(m->object);
  delete m;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJMatrixSetObjectType(HYPRE_IJMatrix m, HYPRE_Int type) {
  if (!m) return err_arg(1);
  m->object_type = type;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJMatrixInitialize(HYPRE_IJMatrix m);
HYPRE_Int HYPRE_IJMatrixInitialize_v2(HYPRE_IJMatrix m, HYPRE_Int /* HYPRE_MemoryLocation */) { return HYPRE_IJMatrixInitialize(m); }
HYPRE_Int HYPRE_IJMatrixGetLocalRange(HYPRE_IJMatrix m, HYPRE_BigInt *ilower, HYPRE_BigInt *iupper, HYPRE_BigInt *jlower,
                                      HYPRE_BigInt *jupper) {             // IJ_mv/HYPRE_IJMatrix.c:1042
  if (!m) return err_arg(1);
  if (!ilower || !iupper || !jlower || !jupper) return err_arg(2);
  *ilower = m->ilower; *iupper = m->iupper; *jlower = m->jlower; *jupper = m->jupper;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJMatrixInitialize(HYPRE_IJMatrix m) {
  if (!m) return err_arg(1);
  if (m->object_type != HYPRE_PARCSR) return err_arg(1);             // HYPRE_IJMatrix.c:303-311
  NEED_HANDLE();
  if (m->ij) { b200_ij_destroy(h, m->ij); m->ij = nullptr; }         // a fresh auxiliary matrix (IJMatrix_parcsr.c:178-230)
  CALL(b200_ij_create(h, m->ilower, m->iupper, m->jlower, m->jupper, &m->ij), "HYPRE_IJMatrixInitialize");
  m->initialized = true;
  m->assembled = false;
  return g_error_flag;
}
// SetValues / AddToValues: validate and append to the assembler's record log (pinned chunks streamed to the device);
// the merge itself happens on the device at Assemble (b200_ij.cu)
static HYPRE_Int ij_set(HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows, const HYPRE_BigInt *cols,
                        const HYPRE_Complex *values, bool add) {
  if (!m) return err_arg(1);
  if (nrows == 0) return g_error_flag;
  if (nrows < 0) return err_arg(2);
  if (!rows) return err_arg(4);
  if (!cols) return err_arg(5);
  if (!values) return err_arg(6);
  if (!m->initialized || !m->ij) return err_arg(1);
  NEED_HANDLE();
  int rejected = 0;
  std::vector<int> ones;
  if (!ncols) { ones.assign((size_t)nrows, 1); ncols = ones.data(); }      // NULL = one entry per row (IJMatrix_parcsr.c:916)
  CALL(b200_ij_set_values(h, m->ij, nrows, ncols, rows, cols, values, add ? 1 : 0, &rejected), "HYPRE_IJMatrixSetValues");
  if (rejected) {
    // the reference stashes off-processor rows for the owner (IJMatrix_parcsr.c:1395-1450); with one rank per
    // process there is no owner to send them to, and columns outside [jlower, jupper] do not exist
    fprintf(stderr, "hypre_b200: IJMatrix: %d entries outside rows [%d, %d] / columns [%d, %d] were dropped\n", rejected, m->ilower,
            m->iupper, m->jlower, m->jupper);
    err(HYPRE_ERROR_GENERIC);
  }
  return g_error_flag;
}
HYPRE_Int HYPRE_IJMatrixSetValues(HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows,
                                  const HYPRE_BigInt *cols, const HYPRE_Complex *values) {
  return ij_set(m, nrows, ncols, rows, cols, values, false);
}
HYPRE_Int HYPRE_IJMatrixAddToValues(HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int *ncols, const HYPRE_BigInt *rows,
                                    const HYPRE_BigInt *cols, const HYPRE_Complex *values) {
  return ij_set(m, nrows, ncols, rows, cols, values, true);
}
HYPRE_Int HYPRE_IJMatrixAssemble(HYPRE_IJMatrix m) {
  if (!m) return err_arg(1);
  if (!m->initialized || !m->ij) return err_arg(1);
  NEED_HANDLE();
  const bool first = !m->assembled;
  b200_parcsr A = nullptr;
  int missing = 0;
  if (b200_ij_assemble(h, m->ij, &A, &missing)) return err_b200("HYPRE_IJMatrixAssemble");
  if (first) {
    if (m->object) { HYPRE_ParCSRMatrixDestroy(m->object); m->object = nullptr; }     // object of an earlier Initialize cycle
    hypre_ParCSRMatrix_struct *P = new hypre_ParCSRMatrix_struct();
    P->A = A;
    int nr = 0, nd = 0, no = 0, nco = 0;
    b200_parcsr_local_rows(A, &nr, &nd, &no, &nco);
    P->global_rows = m->iupper - m->ilower + 1; P->global_cols = m->jupper - m->jlower + 1; P->nnz = (long long)nd + no;
    m->object = P;
  } else if (missing) {
    fprintf(stderr, "hypre_b200: IJMatrix: %d values were set on elements that do not exist in the assembled matrix\n", missing);
    err(HYPRE_ERROR_GENERIC);                                   // " Error, element %b %b does not exist" (IJMatrix_parcsr.c:836-842)
  }
  m->assembled = true;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJMatrixGetObject(HYPRE_IJMatrix m, void **object) {
  if (!m) return err_arg(1);
  *object = m->object;
  return g_error_flag;
}

// ---- IJ vector ------------------------------------------------------------------------------
HYPRE_Int HYPRE_IJVectorCreate(MPI_Comm, HYPRE_BigInt jlower, HYPRE_BigInt jupper, HYPRE_IJVector *vector) {
  if (jlower > jupper + 1 || jlower < 0) return err_arg(2);          // HYPRE_IJVector.c:45-57
  if (jupper < -1) return err_arg(3);
  hypre_IJVector_struct *v = new hypre_IJVector_struct();
  v->jlower = jlower; v->jupper = jupper;
  *vector = v;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorDestroy(HYPRE_ParVector v);
HYPRE_Int HYPRE_IJVectorDestroy(HYPRE_IJVector v) {
  if (!v) return err_arg(1);
  if (v->object) HYPRE_ParVectorDestroy(v->object);
  delete v;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorSetObjectType(HYPRE_IJVector v, HYPRE_Int type) {
  if (!v) return err_arg(1);
  v->object_type = type;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorInitialize(HYPRE_IJVector v) {
  if (!v) return err_arg(1);
  if (v->object_type != HYPRE_PARCSR) return err_arg(1);
  NEED_HANDLE();
  const int n = v->jupper - v->jlower + 1;
  v->host.assign((size_t)n, 0.0);
  if (!v->object) {
    v->object = new hypre_ParVector_struct();
    v->object->n = n;
    CALL(b200_malloc(h, (void **)&v->object->d, sizeof(double) * (size_t)(n > 0 ? n : 1)), "HYPRE_IJVectorInitialize");
    v->object->initialized = true;
  }
  CALL(b200_vec_fill(h, n, 0.0, v->object->d), "HYPRE_IJVectorInitialize");
  v->initialized = true;
  v->dirty = false;
  return g_error_flag;
}
// The device copy is the truth once a solver has written it: before the first Set / AddTo after an Assemble the host mirror
// is refreshed from the device, so a partial update does not revert the untouched entries to their pre-solve values.
static HYPRE_Int ijvec_refresh_mirror(HYPRE_IJVector v) {
  if (v->dirty || !v->assembled_once || v->host.empty() || !v->object) return 0;
  NEED_HANDLE();
  CALL(b200_memcpy_d2h(h, v->host.data(), v->object->d, sizeof(double) * v->host.size()), "HYPRE_IJVectorSetValues");
  return 0;
}
HYPRE_Int HYPRE_IJVectorInitialize_v2(HYPRE_IJVector v, HYPRE_Int /* HYPRE_MemoryLocation: the object lives in HBM */) {
  return HYPRE_IJVectorInitialize(v);
}
HYPRE_Int hypre_IJVectorZeroValues(HYPRE_IJVector v) {                 // IJ_mv/IJVector_parcsr.c: SetConstantValues(0)
  if (!v || !v->initialized || !v->object) return err_arg(1);
  NEED_HANDLE();
  std::fill(v->host.begin(), v->host.end(), 0.0);
  CALL(b200_vec_fill(h, v->object->n, 0.0, v->object->d), "hypre_IJVectorZeroValues");
  v->dirty = false;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorSetValues(HYPRE_IJVector v, HYPRE_Int nvalues, const HYPRE_BigInt *indices, const HYPRE_Complex *values) {
  if (!v) return err_arg(1);
  if (nvalues == 0) return g_error_flag;
  if (nvalues < 0) return err_arg(2);
  if (!values) return err_arg(4);
  if (!v->initialized) return err_arg(1);
  if (ijvec_refresh_mirror(v)) return g_error_flag;
  for (int k = 0; k < nvalues; k++) {
    const int g = indices ? indices[k] : v->jlower + k;               // NULL indices = contiguous from jlower (IJVector_parcsr.c:374-390)
    if (g < v->jlower || g > v->jupper) continue;                     // off-rank entries are dropped on one rank
    v->host[(size_t)(g - v->jlower)] = values[k];
  }
  v->dirty = true;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorAddToValues(HYPRE_IJVector v, HYPRE_Int nvalues, const HYPRE_BigInt *indices, const HYPRE_Complex *values) {
  if (!v) return err_arg(1);
  if (nvalues == 0) return g_error_flag;
  if (nvalues < 0) return err_arg(2);
  if (!values) return err_arg(4);
  if (!v->initialized) return err_arg(1);
  if (ijvec_refresh_mirror(v)) return g_error_flag;                     // a solve may have written the device copy since
  for (int k = 0; k < nvalues; k++) {
    const int g = indices ? indices[k] : v->jlower + k;
    if (g < v->jlower || g > v->jupper) continue;
    v->host[(size_t)(g - v->jlower)] += values[k];
  }
  v->dirty = true;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorAssemble(HYPRE_IJVector v) {
  if (!v) return err_arg(1);
  if (!v->initialized) return err_arg(1);
  NEED_HANDLE();
  if (v->dirty && !v->host.empty())
    CALL(b200_memcpy_h2d(h, v->object->d, v->host.data(), sizeof(double) * v->host.size()), "HYPRE_IJVectorAssemble");
  v->dirty = false;
  v->assembled_once = true;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorGetValues(HYPRE_IJVector v, HYPRE_Int nvalues, const HYPRE_BigInt *indices, HYPRE_Complex *values) {
  if (!v) return err_arg(1);
  if (nvalues == 0) return g_error_flag;
  if (nvalues < 0) return err_arg(2);
  if (!values) return err_arg(4);
  if (!v->initialized || !v->object) return err_arg(1);
  NEED_HANDLE();
  if (!v->host.empty())       // the device copy is the truth (solvers write it); refresh the host mirror
    CALL(b200_memcpy_d2h(h, v->host.data(), v->object->d, sizeof(double) * v->host.size()), "HYPRE_IJVectorGetValues");
  for (int k = 0; k < nvalues; k++) {
    const int g = indices ? indices[k] : v->jlower + k;
    if (g < v->jlower || g > v->jupper) return err_arg(3);           // IJVector_parcsr.c:583-590
    values[k] = v->host[(size_t)(g - v->jlower)];
  }
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorGetObject(HYPRE_IJVector v, void **object) {
  if (!v) return err_arg(1);
  *object = v->object;
  return g_error_flag;
}

// ---- on-disk IJ format (HYPRE_IJMatrixRead / Print, HYPRE_IJVectorRead / Print; ij -fromfile / -print) -----------
// matrix file "<name>.00000": "ilower iupper jlower jupper" then one "i j %.14e" line per entry in storage order
// (IJ_mv/HYPRE_IJMatrix.c:1144-1300); vector file: "jlower jupper" then "j %.14e" lines (HYPRE_IJVector.c:590-710).
static HYPRE_Int print_csr_ij(b200_handle h, b200_parcsr A, int ilower, int iupper, int jlower, int jupper, const char *filename) {
  char name[512];
  snprintf(name, sizeof name, "%s.%05d", filename, 0);
  FILE *f = fopen(name, "w");
  if (!f) return err_arg(2);
  int nr = 0, nd = 0, no = 0, nco = 0;
  b200_parcsr_local_rows(A, &nr, &nd, &no, &nco);
  std::vector<int> I((size_t)nr + 1, 0), J((size_t)(nd ? nd : 1));
  std::vector<double> V((size_t)(nd ? nd : 1));
  if (b200_csr_download(h, b200_parcsr_diag(A), I.data(), J.data(), V.data())) { fclose(f); return err_b200("HYPRE_IJMatrixPrint"); }
  fprintf(f, "%d %d %d %d\n", ilower, iupper, jlower, jupper);
  for (int i = 0; i < nr; i++)
    for (int k = I[i]; k < I[i + 1]; k++) fprintf(f, "%d %d %.14e\n", ilower + i, jlower + J[k], V[k]);
  fclose(f);
  return g_error_flag;
}
HYPRE_Int HYPRE_IJMatrixPrint(HYPRE_IJMatrix m, const char *filename) {
  if (!m || m->object_type != HYPRE_PARCSR || !m->object) return err_arg(1);
  if (!filename) return err_arg(2);
  NEED_HANDLE();
  return print_csr_ij(h, m->object->A, m->ilower, m->iupper, m->jlower, m->jupper, filename);
}
HYPRE_Int HYPRE_IJMatrixRead(const char *filename, MPI_Comm comm, HYPRE_Int type, HYPRE_IJMatrix *matrix_ptr) {
  if (!filename) return err_arg(1);
  if (!matrix_ptr) return err_arg(4);
  char name[512];
  snprintf(name, sizeof name, "%s.%05d", filename, 0);
  FILE *f = fopen(name, "r");
  if (!f) return err_arg(1);
  int ilower, iupper, jlower, jupper;
  if (fscanf(f, "%d %d %d %d", &ilower, &iupper, &jlower, &jupper) != 4) { fclose(f); fprintf(stderr, "hypre_b200: Error in IJ matrix input file.\n"); return err(HYPRE_ERROR_GENERIC); }
  HYPRE_IJMatrix m = nullptr;
  HYPRE_IJMatrixCreate(comm, ilower, iupper, jlower, jupper, &m);
  HYPRE_IJMatrixSetObjectType(m, type);
  if (HYPRE_IJMatrixInitialize(m) || !m->ij) { fclose(f); return g_error_flag; }
  int II, JJ, ncols = 1, ret;
  double value;
  while ((ret = fscanf(f, "%d %d%*[ \t]%le", &II, &JJ, &value)) != EOF) {       // one SetValues per line, as the reference
    if (ret != 3) { fclose(f); fprintf(stderr, "hypre_b200: Error in IJ matrix input file.\n"); return err(HYPRE_ERROR_GENERIC); }
    HYPRE_IJMatrixSetValues(m, 1, &ncols, &II, &JJ, &value);                      // (rows outside [ilower, iupper] are rejected there)
  }
  fclose(f);
  HYPRE_IJMatrixAssemble(m);
  *matrix_ptr = m;
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorPrint(HYPRE_IJVector v, const char *filename) {
  if (!v || !v->initialized || !v->object) return err_arg(1);
  if (!filename) return err_arg(2);
  NEED_HANDLE();
  char name[512];
  snprintf(name, sizeof name, "%s.%05d", filename, 0);
  FILE *f = fopen(name, "w");
  if (!f) return err_arg(2);
  if (!v->host.empty() && b200_memcpy_d2h(h, v->host.data(), v->object->d, sizeof(double) * v->host.size())) { fclose(f); return err_b200("HYPRE_IJVectorPrint"); }
  fprintf(f, "%d %d\n", v->jlower, v->jupper);
  for (int j = v->jlower; j <= v->jupper; j++) fprintf(f, "%d %.14e\n", j, v->host[(size_t)(j - v->jlower)]);
  fclose(f);
  return g_error_flag;
}
HYPRE_Int HYPRE_IJVectorRead(const char *filename, MPI_Comm comm, HYPRE_Int type, HYPRE_IJVector *vector_ptr) {
  if (!filename) return err_arg(1);
  if (!vector_ptr) return err_arg(4);
  char name[512];
  snprintf(name, sizeof name, "%s.%05d", filename, 0);
  FILE *f = fopen(name, "r");
  if (!f) return err_arg(1);
  int jlower, jupper;
  if (fscanf(f, "%d %d", &jlower, &jupper) != 2) { fclose(f); fprintf(stderr, "hypre_b200: Error in IJ vector input file.\n"); return err(HYPRE_ERROR_GENERIC); }
  HYPRE_IJVector v = nullptr;
  HYPRE_IJVectorCreate(comm, jlower, jupper, &v);
  HYPRE_IJVectorSetObjectType(v, type);
  if (HYPRE_IJVectorInitialize(v)) { fclose(f); return g_error_flag; }
  int j, ret;
  double value;
  while ((ret = fscanf(f, "%d%*[ \t]%le", &j, &value)) != EOF) {
    if (ret != 2) { fclose(f); fprintf(stderr, "hypre_b200: Error in IJ vector input file.\n"); return err(HYPRE_ERROR_GENERIC); }
    if (j < jlower || j > jupper) HYPRE_IJVectorAddToValues(v, 1, &j, &value);
    else HYPRE_IJVectorSetValues(v, 1, &j, &value);
  }
  fclose(f);
  HYPRE_IJVectorAssemble(v);
  *vector_ptr = v;
  return g_error_flag;
}

// ---- ParCSR matrix / vector -----------------------------------------------------------------
static HYPRE_ParCSRMatrix wrap_generated(b200_parcsr A, long long gr) {
  hypre_ParCSRMatrix_struct *P = new hypre_ParCSRMatrix_struct();
  P->A = A;
  int nr = 0, nd = 0, no = 0, nco = 0;
  b200_parcsr_local_rows(A, &nr, &nd, &no, &nco);
  P->global_rows = P->global_cols = (int)gr;
  P->nnz = (long long)nd + no;
  return P;
}
HYPRE_ParCSRMatrix GenerateLaplacian(MPI_Comm, HYPRE_BigInt nx, HYPRE_BigInt ny, HYPRE_BigInt nz, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int R,
                                     HYPRE_Int p, HYPRE_Int q, HYPRE_Int r, HYPRE_Real *value) {
  b200_handle h = handle();
  if (!h) return nullptr;
  if (P * Q * R != 1) {
    fprintf(stderr, "hypre_b200: GenerateLaplacian through the HYPRE API is single-rank; use b200_dist_generate_laplacian\n");
    err(HYPRE_ERROR_GENERIC);
    return nullptr;
  }
  b200_parcsr A = nullptr;
  if (b200_generate_laplacian(h, nx, ny, nz, P, Q, R, p, q, r, value, &A)) { err_b200("GenerateLaplacian"); return nullptr; }
  return wrap_generated(A, (long long)nx * ny * nz);
}
HYPRE_ParCSRMatrix GenerateRotate7pt(MPI_Comm, HYPRE_BigInt nx, HYPRE_BigInt ny, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int p, HYPRE_Int q,
                                     HYPRE_Real alpha, HYPRE_Real eps) {                             // par_rotate_7pt.c:15
  b200_handle h = handle();
  if (!h) return nullptr;
  if (P * Q != 1) {
    fprintf(stderr, "hypre_b200: GenerateRotate7pt through the HYPRE API is single-rank; use b200_dist_generate_rotate7pt\n");
    err(HYPRE_ERROR_GENERIC);
    return nullptr;
  }
  b200_parcsr A = nullptr;
  if (b200_generate_rotate7pt(h, nx, ny, P, Q, p, q, alpha, eps, &A)) { err_b200("GenerateRotate7pt"); return nullptr; }
  return wrap_generated(A, (long long)nx * ny);
}
HYPRE_ParCSRMatrix GenerateDifConv(MPI_Comm, HYPRE_BigInt nx, HYPRE_BigInt ny, HYPRE_BigInt nz, HYPRE_Int P, HYPRE_Int Q, HYPRE_Int R,
                                   HYPRE_Int p, HYPRE_Int q, HYPRE_Int r, HYPRE_Real *value) {     // par_difconv.c:15
  b200_handle h = handle();
  if (!h) return nullptr;
  if (P * Q * R != 1) {
    fprintf(stderr, "hypre_b200: GenerateDifConv through the HYPRE API is single-rank; use b200_dist_generate_difconv\n");
    err(HYPRE_ERROR_GENERIC);
    return nullptr;
  }
  b200_parcsr A = nullptr;
  if (b200_generate_difconv(h, nx, ny, nz, P, Q, R, p, q, r, value, &A)) { err_b200("GenerateDifConv"); return nullptr; }
  return wrap_generated(A, (long long)nx * ny * nz);
}
HYPRE_ParCSRMatrix GenerateLaplacian27pt(MPI_Comm, HYPRE_BigInt nx, HYPRE_BigInt ny, HYPRE_BigInt nz, HYPRE_Int P, HYPRE_Int Q,
                                         HYPRE_Int R, HYPRE_Int p, HYPRE_Int q, HYPRE_Int r, HYPRE_Real *value) {
  b200_handle h = handle();
  if (!h) return nullptr;
  if (P * Q * R != 1) {
    fprintf(stderr, "hypre_b200: GenerateLaplacian27pt through the HYPRE API is single-rank; use b200_dist_generate_laplacian\n");
    err(HYPRE_ERROR_GENERIC);
    return nullptr;
  }
  b200_parcsr A = nullptr;
  if (b200_generate_laplacian27(h, nx, ny, nz, P, Q, R, p, q, r, value, &A)) { err_b200("GenerateLaplacian27pt"); return nullptr; }
  return wrap_generated(A, (long long)nx * ny * nz);
}
// hypre_ParCSRMatrixPrintIJ (parcsr_mv/par_csr_matrix.c:696-830): the same file format from a ParCSR object
HYPRE_Int hypre_ParCSRMatrixPrintIJ(HYPRE_ParCSRMatrix A, HYPRE_Int base_i, HYPRE_Int base_j, const char *filename) {
  if (!A || !A->A) return err_arg(1);
  if (!filename) return err_arg(4);
  NEED_HANDLE();
  return print_csr_ij(h, A->A, base_i, base_i + A->global_rows - 1, base_j, base_j + A->global_cols - 1, filename);
}
HYPRE_Int HYPRE_ParCSRMatrixDestroy(HYPRE_ParCSRMatrix A) {
  if (!A) return err_arg(1);
  NEED_HANDLE();
  if (A->owned && A->A) b200_parcsr_destroy(h, A->A);
  delete A;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRMatrixGetDims(HYPRE_ParCSRMatrix A, HYPRE_BigInt *M, HYPRE_BigInt *N) {
  if (!A) return err_arg(1);
  *M = A->global_rows; *N = A->global_cols;
  return g_error_flag;
}
// hypre_ParCSRMatrixMigrate / hypre_ParVectorMigrate (parcsr_mv/par_csr_matrix.c, par_vector.c): the reference driver moves
// the system "to the wanted memory space" before the solve (test/ij.c:3193-3195); objects of this library are born in HBM
HYPRE_Int hypre_ParCSRMatrixMigrate(HYPRE_ParCSRMatrix A, HYPRE_Int) { return A ? g_error_flag : err_arg(1); }
HYPRE_Int hypre_ParVectorMigrate(HYPRE_ParVector v, HYPRE_Int) { return v ? g_error_flag : err_arg(1); }
HYPRE_Int HYPRE_ParCSRMatrixGetLocalRange(HYPRE_ParCSRMatrix A, HYPRE_BigInt *row_start, HYPRE_BigInt *row_end,
                                          HYPRE_BigInt *col_start, HYPRE_BigInt *col_end) {
  if (!A) return err_arg(1);
  *row_start = 0; *row_end = A->global_rows - 1; *col_start = 0; *col_end = A->global_cols - 1;
  return g_error_flag;
}
HYPRE_Int HYPRE_b200_ParCSRMatrixGetNumNonzeros(HYPRE_ParCSRMatrix A, long long *nnz) {
  if (!A) return err_arg(1);
  *nnz = A->nnz;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRMatrixMatvec(HYPRE_Complex alpha, HYPRE_ParCSRMatrix A, HYPRE_ParVector x, HYPRE_Complex beta, HYPRE_ParVector y) {
  if (!A) return err_arg(2);
  if (!x) return err_arg(3);
  if (!y) return err_arg(5);
  NEED_HANDLE();
  if (x == y) { fprintf(stderr, "hypre_b200: Matvec needs x != y\n"); return err(HYPRE_ERROR_GENERIC); }   // par_csr_matvec.c:68-76
  if (x->n != A->global_cols || y->n != A->global_rows) return err(HYPRE_ERROR_GENERIC);                   // size check :85-98
  CALL(b200_parcsr_matvec(h, alpha, A->A, x->d, beta, y->d, y->d), "HYPRE_ParCSRMatrixMatvec");
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRMatrixMatvecT(HYPRE_Complex alpha, HYPRE_ParCSRMatrix A, HYPRE_ParVector x, HYPRE_Complex beta, HYPRE_ParVector y) {
  if (!A) return err_arg(2);
  if (!x) return err_arg(3);
  if (!y) return err_arg(5);
  NEED_HANDLE();
  if (x == y) { fprintf(stderr, "hypre_b200: MatvecT needs x != y\n"); return err(HYPRE_ERROR_GENERIC); }
  if (x->n != A->global_rows || y->n != A->global_cols) return err(HYPRE_ERROR_GENERIC);      // par_csr_matvec.c:420-440
  CALL(b200_csr_matvecT(h, alpha, b200_parcsr_diag(A->A), x->d, beta, y->d, y->d), "HYPRE_ParCSRMatrixMatvecT");
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorCreate(MPI_Comm, HYPRE_BigInt global_size, HYPRE_BigInt *, HYPRE_ParVector *vector) {
  if (global_size < 0) return err_arg(2);
  hypre_ParVector_struct *v = new hypre_ParVector_struct();
  v->n = global_size;
  *vector = v;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorInitialize(HYPRE_ParVector v) {
  if (!v) return err_arg(1);
  NEED_HANDLE();
  if (!v->d) {
    CALL(b200_malloc(h, (void **)&v->d, sizeof(double) * (size_t)(v->n > 0 ? v->n : 1)), "HYPRE_ParVectorInitialize");
    CALL(b200_vec_fill(h, v->n, 0.0, v->d), "HYPRE_ParVectorInitialize");
  }
  v->initialized = true;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorSetConstantValues(HYPRE_ParVector v, HYPRE_Complex value) {
  if (!v || !v->d) return err_arg(1);
  NEED_HANDLE();
  CALL(b200_vec_fill(h, v->n, value, v->d), "HYPRE_ParVectorSetConstantValues");
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorDestroy(HYPRE_ParVector v) {
  if (!v) return err_arg(1);
  NEED_HANDLE();
  if (v->d) b200_free(h, v->d);
  delete v;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorInnerProd(HYPRE_ParVector x, HYPRE_ParVector y, HYPRE_Real *prod) {
  if (!x || !x->d) return err_arg(1);
  if (!y || !y->d) return err_arg(2);
  NEED_HANDLE();
  CALL(b200_vec_dot(h, x->n, x->d, y->d, prod), "HYPRE_ParVectorInnerProd");
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorCopy(HYPRE_ParVector x, HYPRE_ParVector y) {
  if (!x || !x->d) return err_arg(1);
  if (!y || !y->d) return err_arg(2);
  NEED_HANDLE();
  CALL(b200_vec_copy(h, x->n, x->d, y->d), "HYPRE_ParVectorCopy");
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorScale(HYPRE_Complex value, HYPRE_ParVector x) {
  if (!x || !x->d) return err_arg(2);
  NEED_HANDLE();
  CALL(b200_vec_scale(h, x->n, value, x->d), "HYPRE_ParVectorScale");
  return g_error_flag;
}
HYPRE_Int HYPRE_ParVectorAxpy(HYPRE_Complex alpha, HYPRE_ParVector x, HYPRE_ParVector y) {
  if (!x || !x->d) return err_arg(2);
  if (!y || !y->d) return err_arg(3);
  NEED_HANDLE();
  CALL(b200_vec_axpy(h, x->n, alpha, x->d, y->d), "HYPRE_ParVectorAxpy");
  return g_error_flag;
}
HYPRE_Int HYPRE_b200_ParVectorGetHostValues(HYPRE_ParVector v, HYPRE_Complex *host_out) {
  if (!v || !v->d) return err_arg(1);
  NEED_HANDLE();
  CALL(b200_memcpy_d2h(h, host_out, v->d, sizeof(double) * (size_t)v->n), "HYPRE_b200_ParVectorGetHostValues");
  return g_error_flag;
}
HYPRE_Int HYPRE_b200_ParVectorSetHostValues(HYPRE_ParVector v, const HYPRE_Complex *host_in) {
  if (!v || !v->d) return err_arg(1);
  NEED_HANDLE();
  CALL(b200_memcpy_h2d(h, v->d, host_in, sizeof(double) * (size_t)v->n), "HYPRE_b200_ParVectorSetHostValues");
  return g_error_flag;
}

// ---- BoomerAMG ------------------------------------------------------------------------------
HYPRE_Int HYPRE_BoomerAMGCreate(HYPRE_Solver *solver) {
  if (!solver) return err_arg(1);
  hypre_Solver_struct *s = new hypre_Solver_struct();
  s->kind = KIND_AMG;
  b200_amg_create(&s->amg);
  // library defaults of hypre_BoomerAMGCreate (par_amg.c:139-230) where they differ from b200_amg's
  // ij-driver defaults: coarsen 10 (HMIS), relax 13/14 hybrid GS, mod_rap2 0, tol 1e-7, max_iter 20
  s->stored = {{"CoarsenType", 10}, {"InterpType", 6}, {"PMaxElmts", 4}, {"RelaxType", -1}, {"ModuleRAP2", 0}, {"RAP2", 0},
               {"KeepTranspose", 0}, {"Tol", 1e-7}, {"MaxIter", 20}, {"MinIter", 0}, {"RelaxOrder", 0}, {"NumSweeps", 1},
               {"CycleType", 1}, {"MaxLevels", 25}, {"MaxCoarseSize", 9}, {"MinCoarseSize", 0}, {"AggNumLevels", 0},
               {"NumFunctions", 1}, {"StrongThreshold", 0.25}, {"MaxRowSum", 0.9}, {"TruncFactor", 0.0}, {"RelaxWt", 1.0},
               {"OuterWt", 1.0}, {"PrintLevel", 0}, {"Logging", 0}, {"GSBlocks", 1}, {"FCycle", 0},
               {"ChebyOrder", 2}, {"ChebyEigEst", 10}, {"ChebyVariant", 0}, {"ChebyScale", 1}, {"ChebyFraction", 0.3}};   // par_amg.c:215-220
  for (const Neutral &nv : kNeutral) s->stored[nv.name] = nv.value;
  *solver = s;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGDestroy(HYPRE_Solver s) {
  if (!is_amg(s)) return err_arg(1);
  NEED_HANDLE();
  b200_amg_destroy(h, s->amg);
  delete s;
  return g_error_flag;
}

#define AMG_SETTER(Name, Type, Check)                                                 \
  HYPRE_Int HYPRE_BoomerAMGSet##Name(HYPRE_Solver s, Type v) {                        \
    if (!is_amg(s)) return err_arg(1);                                                \
    if (!(Check)) return err_arg(2);                                                  \
    s->stored[#Name] = (double)v;                                                     \
    return g_error_flag;                                                              \
  }
// argument checks as in par_amg.c (e.g. :1032 max_levels < 1, :1166 strong_threshold range, :1440 tol range)
AMG_SETTER(CoarsenType, HYPRE_Int, true)
AMG_SETTER(InterpType, HYPRE_Int, (v >= 0 && v <= 25) || v == 100)
AMG_SETTER(PMaxElmts, HYPRE_Int, v >= 0)
AMG_SETTER(TruncFactor, HYPRE_Real, v >= 0 && v < 1.0)
AMG_SETTER(StrongThreshold, HYPRE_Real, v >= 0 && v <= 1.0)
AMG_SETTER(MaxRowSum, HYPRE_Real, v > 0 && v <= 1.0)
AMG_SETTER(RelaxType, HYPRE_Int, v >= 0)
AMG_SETTER(RelaxOrder, HYPRE_Int, true)
AMG_SETTER(RelaxWt, HYPRE_Real, true)
AMG_SETTER(OuterWt, HYPRE_Real, true)
HYPRE_Int HYPRE_BoomerAMGSetNumSweeps(HYPRE_Solver s, HYPRE_Int v) {          // par_amg.c:1934-1962: down = up = v, coarse = 1
  if (!is_amg(s)) return err_arg(1);
  if (v < 1) return err_arg(2);
  s->stored["NumSweeps"] = v;
  for (char k = '1'; k <= '3'; k++) s->stored.erase(std::string("CycleNumSweeps") + k);
  return g_error_flag;
}
AMG_SETTER(CycleType, HYPRE_Int, v >= 1 && v <= 2)
AMG_SETTER(MaxLevels, HYPRE_Int, v >= 1)
AMG_SETTER(MaxCoarseSize, HYPRE_Int, v >= 1)
AMG_SETTER(MinCoarseSize, HYPRE_Int, v >= 0)
AMG_SETTER(MaxIter, HYPRE_Int, v >= 0)
AMG_SETTER(MinIter, HYPRE_Int, true)
AMG_SETTER(Tol, HYPRE_Real, v >= 0 && v <= 1.0)
AMG_SETTER(AggNumLevels, HYPRE_Int, v >= 0)
AMG_SETTER(NumFunctions, HYPRE_Int, v >= 1)
AMG_SETTER(RAP2, HYPRE_Int, true)
AMG_SETTER(ModuleRAP2, HYPRE_Int, true)
AMG_SETTER(KeepTranspose, HYPRE_Int, true)
AMG_SETTER(PrintLevel, HYPRE_Int, true)
AMG_SETTER(Logging, HYPRE_Int, true)
AMG_SETTER(DebugFlag, HYPRE_Int, true)
// features outside the B200 path: stored, and rejected at Setup unless left at the neutral value
AMG_SETTER(PostInterpType, HYPRE_Int, true)
AMG_SETTER(SmoothNumLevels, HYPRE_Int, true)
AMG_SETTER(Additive, HYPRE_Int, true)
AMG_SETTER(MultAdditive, HYPRE_Int, true)
AMG_SETTER(Simple, HYPRE_Int, true)
AMG_SETTER(Nodal, HYPRE_Int, true)
AMG_SETTER(FCycle, HYPRE_Int, true)
AMG_SETTER(NonGalerkinTol, HYPRE_Real, v >= 0)
AMG_SETTER(GSMG, HYPRE_Int, true)
AMG_SETTER(CoarsenCutFactor, HYPRE_Int, true)
AMG_SETTER(Redundant, HYPRE_Int, true)
AMG_SETTER(SeqThreshold, HYPRE_Int, true)
// parameters that are inert unless one of the switches above is on (the reference only stores them too)
AMG_SETTER(CGCIts, HYPRE_Int, true)
AMG_SETTER(NumSamples, HYPRE_Int, true)
AMG_SETTER(MeasureType, HYPRE_Int, true)
AMG_SETTER(JacobiTruncThreshold, HYPRE_Real, true)
AMG_SETTER(SCommPkgSwitch, HYPRE_Real, true)
AMG_SETTER(ISType, HYPRE_Int, true)
AMG_SETTER(NumCRRelaxSteps, HYPRE_Int, true)
AMG_SETTER(CRRate, HYPRE_Real, true)
AMG_SETTER(CRStrongTh, HYPRE_Real, true)
AMG_SETTER(CRUseCG, HYPRE_Int, true)
AMG_SETTER(AddRelaxType, HYPRE_Int, true)
AMG_SETTER(AddRelaxWt, HYPRE_Real, true)
AMG_SETTER(AddLastLvl, HYPRE_Int, true)
AMG_SETTER(MultAddPMaxElmts, HYPRE_Int, true)
AMG_SETTER(MultAddTruncFactor, HYPRE_Real, true)
AMG_SETTER(ChebyOrder, HYPRE_Int, true)
AMG_SETTER(ChebyFraction, HYPRE_Real, true)
AMG_SETTER(ChebyEigEst, HYPRE_Int, true)
AMG_SETTER(ChebyVariant, HYPRE_Int, true)
AMG_SETTER(ChebyScale, HYPRE_Int, true)
AMG_SETTER(SmoothType, HYPRE_Int, true)
AMG_SETTER(SmoothNumSweeps, HYPRE_Int, true)
AMG_SETTER(AggInterpType, HYPRE_Int, true)
AMG_SETTER(AggTruncFactor, HYPRE_Real, true)
AMG_SETTER(AggP12TruncFactor, HYPRE_Real, true)
AMG_SETTER(AggPMaxElmts, HYPRE_Int, true)
AMG_SETTER(AggP12MaxElmts, HYPRE_Int, true)
AMG_SETTER(NumPaths, HYPRE_Int, true)
AMG_SETTER(NodalDiag, HYPRE_Int, true)
AMG_SETTER(Variant, HYPRE_Int, true)
AMG_SETTER(Overlap, HYPRE_Int, true)
AMG_SETTER(DomainType, HYPRE_Int, true)
AMG_SETTER(SchwarzUseNonSymm, HYPRE_Int, true)
AMG_SETTER(SchwarzRlxWeight, HYPRE_Real, true)
AMG_SETTER(EuLevel, HYPRE_Int, true)
AMG_SETTER(EuBJ, HYPRE_Int, true)
AMG_SETTER(EuSparseA, HYPRE_Real, true)
// setters the unmodified reference driver (test/ij.c) calls for features outside this path: stored; the ones in kNeutral are
// rejected at Setup unless left at the neutral value, the others are inert without their switch (par_amg.c stores them too)
AMG_SETTER(ConvergeType, HYPRE_Int, true)
AMG_SETTER(Restriction, HYPRE_Int, true)
AMG_SETTER(StrongThresholdR, HYPRE_Real, v >= 0 && v <= 1.0)
AMG_SETTER(ADropTol, HYPRE_Real, v >= 0)
AMG_SETTER(ADropType, HYPRE_Int, true)
AMG_SETTER(ILUType, HYPRE_Int, true)
AMG_SETTER(ILULevel, HYPRE_Int, true)
AMG_SETTER(ILUMaxRowNnz, HYPRE_Int, true)
AMG_SETTER(ILUMaxIter, HYPRE_Int, true)
AMG_SETTER(ILUDroptol, HYPRE_Real, true)
AMG_SETTER(InterpVecVariant, HYPRE_Int, true)
AMG_SETTER(InterpVecQMax, HYPRE_Int, true)
AMG_SETTER(InterpVecAbsQTrunc, HYPRE_Real, true)
AMG_SETTER(CoordDim, HYPRE_Int, true)
AMG_SETTER(PlotGrids, HYPRE_Int, true)
#undef AMG_SETTER
HYPRE_Int HYPRE_BoomerAMGSetCoordinates(HYPRE_Solver s, float *) { return is_amg(s) ? g_error_flag : err_arg(1); }
HYPRE_Int HYPRE_BoomerAMGSetPlotFileName(HYPRE_Solver s, const char *) { return is_amg(s) ? g_error_flag : err_arg(1); }
HYPRE_Int HYPRE_BoomerAMGSetLevelNonGalerkinTol(HYPRE_Solver s, HYPRE_Real tol, HYPRE_Int) {
  if (!is_amg(s)) return err_arg(1);
  if (tol != 0.0) s->stored["LevelNonGalerkinTolSet"] = 1;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetGridRelaxPoints(HYPRE_Solver s, HYPRE_Int **points) {
  if (!is_amg(s)) return err_arg(1);
  if (points) s->stored["GridRelaxPointsSet"] = 1;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetInterpVectors(HYPRE_Solver s, HYPRE_Int num, HYPRE_ParVector *) {
  if (!is_amg(s)) return err_arg(1);
  s->stored["NumInterpVectors"] = num;
  return g_error_flag;
}

// number of Gauss-Seidel blocks per rank for the hybrid smoothers 3/4/6/8/13/14: what the OpenMP thread
// count is to the reference (par_relax.c:4400-4412, hypre_NumThreads()); default 1
HYPRE_Int HYPRE_b200_BoomerAMGSetGSBlocks(HYPRE_Solver s, HYPRE_Int blocks) {
  if (!is_amg(s)) return err_arg(1);
  if (blocks < 1) return err_arg(2);
  s->stored["GSBlocks"] = blocks;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetOldDefault(HYPRE_Solver s) {               // HYPRE_parcsr_amg.c:1500-1508
  if (!is_amg(s)) return err_arg(1);
  s->stored["CoarsenType"] = 6; s->stored["InterpType"] = 0; s->stored["PMaxElmts"] = 0;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetPrintFileName(HYPRE_Solver s, const char *) { return is_amg(s) ? g_error_flag : err_arg(1); }
HYPRE_Int HYPRE_BoomerAMGSetCycleRelaxType(HYPRE_Solver s, HYPRE_Int relax_type, HYPRE_Int k) {
  if (!is_amg(s)) return err_arg(1);
  if (k < 1 || k > 3) return err_arg(3);                               // par_amg.c:1690-1694
  if (relax_type < 0) return err_arg(2);
  s->stored[std::string("CycleRelaxType") + char('0' + k)] = relax_type;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetCycleNumSweeps(HYPRE_Solver s, HYPRE_Int num_sweeps, HYPRE_Int k) {
  if (!is_amg(s)) return err_arg(1);
  if (k < 1 || k > 3) return err_arg(3);
  if (num_sweeps < 0) return err_arg(2);
  s->stored[std::string("CycleNumSweeps") + char('0' + k)] = num_sweeps;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetLevelRelaxWt(HYPRE_Solver s, HYPRE_Real w, HYPRE_Int level) {
  if (!is_amg(s)) return err_arg(1);
  (void)level;
  s->stored["LevelRelaxWtSet"] = 1; s->stored["LevelRelaxWt"] = w;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetLevelOuterWt(HYPRE_Solver s, HYPRE_Real w, HYPRE_Int level) {
  if (!is_amg(s)) return err_arg(1);
  (void)level;
  s->stored["LevelOuterWtSet"] = 1; s->stored["LevelOuterWt"] = w;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetCPoints(HYPRE_Solver s, HYPRE_Int, HYPRE_Int n, HYPRE_BigInt *) {
  if (!is_amg(s)) return err_arg(1);
  s->stored["NumCPoints"] = n;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetFPoints(HYPRE_Solver s, HYPRE_Int n, HYPRE_BigInt *) {
  if (!is_amg(s)) return err_arg(1);
  s->stored["NumFPoints"] = n;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetIsolatedFPoints(HYPRE_Solver s, HYPRE_Int n, HYPRE_BigInt *) {
  if (!is_amg(s)) return err_arg(1);
  s->stored["NumIsolatedFPoints"] = n;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSetDofFunc(HYPRE_Solver s, HYPRE_Int *) { return is_amg(s) ? g_error_flag : err_arg(1); }

HYPRE_Int HYPRE_BoomerAMGSetup(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector, HYPRE_ParVector) {
  if (!is_amg(s)) return err_arg(1);
  if (!A) return err_arg(2);
  NEED_HANDLE();
  auto &st = s->stored;
  for (const Neutral &nv : kNeutral)
    if (st[nv.name] != nv.value) {
      fprintf(stderr, "hypre_b200: BoomerAMGSetup: %s (%s = %g) is not on the B200 path\n", nv.what, nv.name, st[nv.name]);
      return err(HYPRE_ERROR_GENERIC);
    }
  // relax type: SetRelaxType sets all three grid-relax slots, coarse defaults to 9 (par_amg.c:1650-1672);
  // the cycle-specific setters override the slots
  int rdown = (int)st["RelaxType"], rup = rdown, rcoarse = 9;
  if (st["RelaxType"] < 0) { rdown = 13; rup = 14; }                    // library default (par_amg.c:206-209)
  if (st.count("CycleRelaxType1")) rdown = (int)st["CycleRelaxType1"];
  if (st.count("CycleRelaxType2")) rup = (int)st["CycleRelaxType2"];
  if (st.count("CycleRelaxType3")) rcoarse = (int)st["CycleRelaxType3"];
  if (rcoarse != 9) { fprintf(stderr, "hypre_b200: BoomerAMGSetup: coarsest-grid relax type %d (only 9 = Gaussian elimination)\n", rcoarse); return err(HYPRE_ERROR_GENERIC); }
  if (st.count("LevelRelaxWtSet") || st.count("LevelOuterWtSet")) { fprintf(stderr, "hypre_b200: BoomerAMGSetup: per-level relaxation weights are not on the B200 path\n"); return err(HYPRE_ERROR_GENERIC); }
  static const char *ints[] = {"CoarsenType", "InterpType", "PMaxElmts", "MaxLevels", "MaxCoarseSize", "MinCoarseSize", "NumSweeps",
                               "AggNumLevels", "ModuleRAP2", "RAP2", "KeepTranspose", "RelaxOrder", "MaxIter", "MinIter", "CycleType",
                               "NumFunctions", "PrintLevel", "GSBlocks", "ChebyOrder", "ChebyEigEst", "ChebyVariant", "ChebyScale"};
  static const char *reals[] = {"StrongThreshold", "MaxRowSum", "TruncFactor", "RelaxWt", "OuterWt", "Tol", "ChebyFraction"};
  for (const char *k : ints) CALL(b200_amg_set_int(s->amg, k, (int)st[k]), "HYPRE_BoomerAMGSetup");
  for (const char *k : reals) CALL(b200_amg_set_real(s->amg, k, st[k]), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_set_int(s->amg, "RelaxType", rdown), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_set_int(s->amg, "UserRelaxType", st["RelaxType"] < 0 ? -1 : rdown), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_set_int(s->amg, "RelaxTypeUp", rup), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_set_int(s->amg, "NumSweepsDown", st.count("CycleNumSweeps1") ? (int)st["CycleNumSweeps1"] : -1), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_set_int(s->amg, "NumSweepsUp", st.count("CycleNumSweeps2") ? (int)st["CycleNumSweeps2"] : -1), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_set_int(s->amg, "NumSweepsCoarse", st.count("CycleNumSweeps3") ? (int)st["CycleNumSweeps3"] : 1), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_set_int(s->amg, "FCycle", (int)st["FCycle"]), "HYPRE_BoomerAMGSetup");
  CALL(b200_amg_setup(h, s->amg, A->A), "HYPRE_BoomerAMGSetup");
  const int pl = (int)st["PrintLevel"];
  if (pl == 1 || pl == 3) setup_stats(h, s);
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGSolve(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!is_amg(s)) return err_arg(1);
  if (!A) return err_arg(2);
  if (!b || !b->d) return err_arg(3);
  if (!x || !x->d) return err_arg(4);
  NEED_HANDLE();
  const int rc = b200_amg_solve_ex(h, s->amg, A->A, b->d, x->d, &s->num_iterations, &s->rel_res);
  if (rc == HYPRE_ERROR_CONV) return err(HYPRE_ERROR_CONV);
  if (rc) return err_b200("HYPRE_BoomerAMGSolve");
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGGetNumIterations(HYPRE_Solver s, HYPRE_Int *n) {
  if (!is_amg(s)) return err_arg(1);
  *n = s->num_iterations;
  return g_error_flag;
}
HYPRE_Int HYPRE_BoomerAMGGetFinalRelativeResidualNorm(HYPRE_Solver s, HYPRE_Real *r) {
  if (!is_amg(s)) return err_arg(1);
  *r = s->rel_res;
  return g_error_flag;
}
HYPRE_Int HYPRE_b200_BoomerAMGGetNumLevels(HYPRE_Solver s, HYPRE_Int *n) {
  if (!is_amg(s)) return err_arg(1);
  *n = b200_amg_num_levels(s->amg);
  return g_error_flag;
}
// rows and nonzeros of A_level (the columns of the reference's setup statistics table, par_stats.c:22)
HYPRE_Int HYPRE_b200_BoomerAMGGetLevelSize(HYPRE_Solver s, HYPRE_Int level, HYPRE_Int *rows, HYPRE_Int *nnz) {
  if (!is_amg(s)) return err_arg(1);
  b200_csr A = b200_amg_level_A(s->amg, level);
  if (!A) return err_arg(2);
  int nr = 0, nc = 0, nz = 0;
  b200_csr_dims(A, &nr, &nc, &nz);
  *rows = nr; *nnz = nz;
  return g_error_flag;
}

// ---- diagonal scaling -----------------------------------------------------------------------
HYPRE_Int HYPRE_ParCSRDiagScaleSetup(HYPRE_Solver, HYPRE_ParCSRMatrix, HYPRE_ParVector, HYPRE_ParVector) { return 0; }
HYPRE_Int HYPRE_ParCSRDiagScale(HYPRE_Solver, HYPRE_ParCSRMatrix A, HYPRE_ParVector y, HYPRE_ParVector x) {
  if (!A) return err_arg(2);
  if (!y || !y->d) return err_arg(3);
  if (!x || !x->d) return err_arg(4);
  NEED_HANDLE();
  CALL(b200_parcsr_diag_scale(h, A->A, y->d, x->d), "HYPRE_ParCSRDiagScale");
  return g_error_flag;
}

// ---- PCG --------------------------------------------------------------------------------------
HYPRE_Int HYPRE_ParCSRPCGCreate(MPI_Comm, HYPRE_Solver *solver) {
  if (!solver) return err_arg(2);
  hypre_Solver_struct *s = new hypre_Solver_struct();
  s->kind = KIND_PCG;
  *solver = s;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRPCGDestroy(HYPRE_Solver s) {
  if (!is_pcg(s)) return err_arg(1);
  delete s;
  return g_error_flag;
}
HYPRE_Int HYPRE_PCGSetTol(HYPRE_Solver s, HYPRE_Real v) { if (!is_pcg(s)) return err_arg(1); s->tol = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetAbsoluteTol(HYPRE_Solver s, HYPRE_Real v) { if (!is_pcg(s)) return err_arg(1); s->a_tol = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetMaxIter(HYPRE_Solver s, HYPRE_Int v) { if (!is_pcg(s)) return err_arg(1); s->max_iter = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetTwoNorm(HYPRE_Solver s, HYPRE_Int v) { if (!is_pcg(s)) return err_arg(1); s->two_norm = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetRelChange(HYPRE_Solver s, HYPRE_Int v) { if (!is_pcg(s)) return err_arg(1); s->rel_change = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetRecomputeResidual(HYPRE_Solver s, HYPRE_Int v) { if (!is_pcg(s)) return err_arg(1); s->recompute_residual = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetPrintLevel(HYPRE_Solver s, HYPRE_Int v) { if (!is_pcg(s)) return err_arg(1); s->print_level = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetLogging(HYPRE_Solver s, HYPRE_Int v) { if (!is_pcg(s)) return err_arg(1); s->logging = v; return g_error_flag; }
HYPRE_Int HYPRE_PCGSetPrecond(HYPRE_Solver s, HYPRE_PtrToSolverFcn precond, HYPRE_PtrToSolverFcn precond_setup, HYPRE_Solver precond_solver) {
  if (!is_pcg(s)) return err_arg(1);
  s->precond = precond; s->precond_setup = precond_setup; s->precond_solver = precond_solver;
  return g_error_flag;
}
HYPRE_Int HYPRE_PCGGetPrecond(HYPRE_Solver s, HYPRE_Solver *precond_data) {
  if (!is_pcg(s)) return err_arg(1);
  *precond_data = s->precond_solver;
  return g_error_flag;
}
HYPRE_Int HYPRE_PCGGetNumIterations(HYPRE_Solver s, HYPRE_Int *n) { if (!is_pcg(s)) return err_arg(1); *n = s->num_iterations; return g_error_flag; }
HYPRE_Int HYPRE_PCGGetFinalRelativeResidualNorm(HYPRE_Solver s, HYPRE_Real *r) { if (!is_pcg(s)) return err_arg(1); *r = s->rel_res; return g_error_flag; }
// ParCSR-typed aliases (parcsr_ls/HYPRE_parcsr_pcg.c:14-220)
HYPRE_Int HYPRE_ParCSRPCGSetTol(HYPRE_Solver s, HYPRE_Real v) { return HYPRE_PCGSetTol(s, v); }
HYPRE_Int HYPRE_ParCSRPCGSetAbsoluteTol(HYPRE_Solver s, HYPRE_Real v) { return HYPRE_PCGSetAbsoluteTol(s, v); }
HYPRE_Int HYPRE_ParCSRPCGSetMaxIter(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_PCGSetMaxIter(s, v); }
HYPRE_Int HYPRE_ParCSRPCGSetTwoNorm(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_PCGSetTwoNorm(s, v); }
HYPRE_Int HYPRE_ParCSRPCGSetRelChange(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_PCGSetRelChange(s, v); }
HYPRE_Int HYPRE_ParCSRPCGSetPrintLevel(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_PCGSetPrintLevel(s, v); }
HYPRE_Int HYPRE_ParCSRPCGSetLogging(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_PCGSetLogging(s, v); }
HYPRE_Int HYPRE_ParCSRPCGSetPrecond(HYPRE_Solver s, HYPRE_PtrToParSolverFcn precond, HYPRE_PtrToParSolverFcn precond_setup, HYPRE_Solver ps) {
  return HYPRE_PCGSetPrecond(s, (HYPRE_PtrToSolverFcn)precond, (HYPRE_PtrToSolverFcn)precond_setup, ps);
}
HYPRE_Int HYPRE_ParCSRPCGGetNumIterations(HYPRE_Solver s, HYPRE_Int *n) { return HYPRE_PCGGetNumIterations(s, n); }
HYPRE_Int HYPRE_ParCSRPCGGetFinalRelativeResidualNorm(HYPRE_Solver s, HYPRE_Real *r) { return HYPRE_PCGGetFinalRelativeResidualNorm(s, r); }

HYPRE_Int HYPRE_ParCSRPCGSetup(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!is_pcg(s)) return err_arg(1);
  if (!A) return err_arg(2);
  NEED_HANDLE();
  b200_timer_start(h);
  if (s->precond_setup)                                         // hypre_PCGSetup, pcg.c:250
    s->precond_setup(s->precond_solver, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
  double ms = 0;
  b200_timer_stop_ms(h, &ms);
  s->setup_s = ms * 1e-3;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRPCGSolve(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!is_pcg(s)) return err_arg(1);
  if (!A) return err_arg(2);
  if (!b || !b->d) return err_arg(3);
  if (!x || !x->d) return err_arg(4);
  NEED_HANDLE();
  b200_pcg_params prm;
  prm.tol = s->tol; prm.a_tol = s->a_tol; prm.max_iter = s->max_iter; prm.two_norm = s->two_norm;
  prm.rel_change = s->rel_change; prm.recompute_residual = s->recompute_residual; prm.precond = 0;
  b200_amg amg = nullptr;
  if (s->precond == (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSolve) {
    if (!is_amg(s->precond_solver)) return err_arg(1);
    amg = s->precond_solver->amg;
    // the preconditioner is ONE cycle: the reference driver sets MaxIter 1 / Tol 0 on it (ij.c:3930, :3909)
    if (b200_amg_get_int(amg, "MaxIter") != 1 || s->precond_solver->stored["Tol"] != 0.0) {
      fprintf(stderr, "hypre_b200: PCG: the BoomerAMG preconditioner must have MaxIter 1 and Tol 0\n");
      return err(HYPRE_ERROR_GENERIC);
    }
    prm.precond = 1;
  } else if (s->precond == (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScale) {
    prm.precond = 2;
  } else if (s->precond != nullptr) {
    fprintf(stderr, "hypre_b200: PCG: only HYPRE_BoomerAMGSolve, HYPRE_ParCSRDiagScale or no preconditioner run on the B200 path\n");
    return err(HYPRE_ERROR_GENERIC);
  }
  s->norms.assign((size_t)s->max_iter + 2, 0.0);
  b200_timer_start(h);
  const int rc = b200_pcg_solve_ex(h, A->A, amg, &prm, b->d, x->d, &s->num_iterations, &s->rel_res, s->norms.data());
  double ms = 0;
  b200_timer_stop_ms(h, &ms);
  s->solve_s = ms * 1e-3;
  if (rc) return err_b200("HYPRE_ParCSRPCGSolve");
  if (s->print_level > 1) {                                    // residual table of pcg.c:472-491, :606-627
    double b2 = 0;
    b200_vec_dot(h, b->n, b->d, b->d, &b2);
    const char *nm = s->two_norm ? "2" : "C";
    printf("\n\nIters       ||r||_%s     conv.rate  ||r||_%s/||b||_%s\n", nm, nm, nm);
    printf("-----    ------------   ---------  ------------ \n");
    for (int i = 1; i <= s->num_iterations; i++)
      printf("% 5d    %e    %f    %e\n", i, s->norms[i], s->norms[i] / s->norms[i - 1],
             (s->two_norm && b2 > 0) ? s->norms[i] / std::sqrt(b2) : 0.0);
    printf("\n\n");
  }
  const double eps = std::fmax(s->tol * s->tol, 0.0);
  if (s->num_iterations >= s->max_iter && s->rel_res * s->rel_res >= eps && eps > 0) err(HYPRE_ERROR_CONV);   // pcg.c:741-745
  return g_error_flag;
}
HYPRE_Int HYPRE_PCGSetup(HYPRE_Solver s, HYPRE_Matrix A, HYPRE_Vector b, HYPRE_Vector x) {
  return HYPRE_ParCSRPCGSetup(s, (HYPRE_ParCSRMatrix)A, (HYPRE_ParVector)b, (HYPRE_ParVector)x);
}
HYPRE_Int HYPRE_PCGSolve(HYPRE_Solver s, HYPRE_Matrix A, HYPRE_Vector b, HYPRE_Vector x) {
  return HYPRE_ParCSRPCGSolve(s, (HYPRE_ParCSRMatrix)A, (HYPRE_ParVector)b, (HYPRE_ParVector)x);
}
HYPRE_Int HYPRE_b200_PCGGetTimes(HYPRE_Solver s, HYPRE_Real *setup_s, HYPRE_Real *solve_s) {
  if (!is_pcg(s)) return err_arg(1);
  if (setup_s) *setup_s = s->setup_s;
  if (solve_s) *solve_s = s->solve_s;
  return g_error_flag;
}
HYPRE_Int HYPRE_b200_PCGGetResidualNorms(HYPRE_Solver s, HYPRE_Int n, HYPRE_Real *norms) {
  if (!is_pcg(s)) return err_arg(1);
  for (int i = 0; i < n && i < (int)s->norms.size(); i++) norms[i] = s->norms[i];
  return g_error_flag;
}

// ---- GMRES (krylov/HYPRE_gmres.c:21-320, parcsr_ls/HYPRE_parcsr_gmres.c:15-230) and BiCGSTAB
//      (krylov/HYPRE_bicgstab.c:25-210, parcsr_ls/HYPRE_parcsr_bicgstab.c:15-220): ij -solver 3 / -solver 9 ----------
static HYPRE_Int krylov_create(HYPRE_Solver *solver, int kind) {
  if (!solver) return err_arg(2);
  hypre_Solver_struct *s = new hypre_Solver_struct();
  s->kind = kind;
  *solver = s;
  return g_error_flag;
}
static HYPRE_Int krylov_setup(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!A) return err_arg(2);
  NEED_HANDLE();
  b200_timer_start(h);
  if (s->precond_setup)                                         // hypre_GMRESSetup gmres.c:201, hypre_BiCGSTABSetup bicgstab.c:180
    s->precond_setup(s->precond_solver, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
  double ms = 0;
  b200_timer_stop_ms(h, &ms);
  s->setup_s = ms * 1e-3;
  return g_error_flag;
}
// which device preconditioner the two function pointers stand for: 0 none, 1 BoomerAMG (one cycle), 2 diagonal scaling
static int krylov_precond(HYPRE_Solver s, const char *who, b200_amg *amg) {
  *amg = nullptr;
  if (s->precond == (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSolve) {
    if (!is_amg(s->precond_solver)) { err_arg(1); return -1; }
    *amg = s->precond_solver->amg;
    if (b200_amg_get_int(*amg, "MaxIter") != 1 || s->precond_solver->stored["Tol"] != 0.0) {     // ij.c:5345, :5339
      fprintf(stderr, "hypre_b200: %s: the BoomerAMG preconditioner must have MaxIter 1 and Tol 0\n", who);
      err(HYPRE_ERROR_GENERIC);
      return -1;
    }
    return 1;
  }
  if (s->precond == (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScale) return 2;
  if (s->precond != nullptr) {
    fprintf(stderr, "hypre_b200: %s: only HYPRE_BoomerAMGSolve, HYPRE_ParCSRDiagScale or no preconditioner run on the B200 path\n", who);
    err(HYPRE_ERROR_GENERIC);
    return -1;
  }
  return 0;
}
static void krylov_print_norms(HYPRE_Solver s, b200_handle h, HYPRE_ParVector b) {   // gmres.c:516-526, bicgstab.c:453-461
  double b2 = 0;
  b200_vec_dot(h, b->n, b->d, b->d, &b2);
  const double b_norm = std::sqrt(b2);
  printf("L2 norm of b: %e\n", b_norm);
  printf("Initial L2 norm of residual: %e\n", s->norms[0]);
  printf("=============================================\n\n");
  printf("Iters     resid.norm     conv.rate  rel.res.norm\n");
  printf("-----    ------------    ---------- ------------\n");
  for (int i = 1; i <= s->num_iterations; i++)
    printf("% 5d    %e    %f   %e\n", i, s->norms[i], s->norms[i] / s->norms[i - 1], b_norm > 0 ? s->norms[i] / b_norm : 0.0);
  printf("\n\n");
}

HYPRE_Int HYPRE_ParCSRGMRESCreate(MPI_Comm, HYPRE_Solver *solver) { return krylov_create(solver, KIND_GMRES); }
HYPRE_Int HYPRE_ParCSRGMRESDestroy(HYPRE_Solver s) {
  if (!is_gmres(s)) return err_arg(1);
  delete s;
  return g_error_flag;
}
#define GMRES_SET(NAME, TYPE, FIELD) \
  HYPRE_Int HYPRE_GMRESSet##NAME(HYPRE_Solver s, TYPE v) { if (!is_gmres(s)) return err_arg(1); s->FIELD = v; return g_error_flag; } \
  HYPRE_Int HYPRE_GMRESGet##NAME(HYPRE_Solver s, TYPE *v) { if (!is_gmres(s)) return err_arg(1); *v = s->FIELD; return g_error_flag; }
GMRES_SET(KDim, HYPRE_Int, k_dim)
GMRES_SET(Tol, HYPRE_Real, tol)
GMRES_SET(AbsoluteTol, HYPRE_Real, a_tol)
GMRES_SET(ConvergenceFactorTol, HYPRE_Real, cf_tol)
GMRES_SET(MinIter, HYPRE_Int, min_iter)
GMRES_SET(MaxIter, HYPRE_Int, max_iter)
GMRES_SET(StopCrit, HYPRE_Int, stop_crit)
GMRES_SET(RelChange, HYPRE_Int, rel_change)
GMRES_SET(SkipRealResidualCheck, HYPRE_Int, skip_real_r_check)
GMRES_SET(PrintLevel, HYPRE_Int, print_level)
GMRES_SET(Logging, HYPRE_Int, logging)
#undef GMRES_SET
HYPRE_Int HYPRE_GMRESSetPrecond(HYPRE_Solver s, HYPRE_PtrToSolverFcn precond, HYPRE_PtrToSolverFcn precond_setup, HYPRE_Solver precond_solver) {
  if (!is_gmres(s)) return err_arg(1);
  s->precond = precond; s->precond_setup = precond_setup; s->precond_solver = precond_solver;
  return g_error_flag;
}
HYPRE_Int HYPRE_GMRESGetPrecond(HYPRE_Solver s, HYPRE_Solver *precond_data) {
  if (!is_gmres(s)) return err_arg(1);
  *precond_data = s->precond_solver;
  return g_error_flag;
}
HYPRE_Int HYPRE_GMRESGetNumIterations(HYPRE_Solver s, HYPRE_Int *n) { if (!is_gmres(s)) return err_arg(1); *n = s->num_iterations; return g_error_flag; }
HYPRE_Int HYPRE_GMRESGetConverged(HYPRE_Solver s, HYPRE_Int *c) { if (!is_gmres(s)) return err_arg(1); *c = s->converged; return g_error_flag; }
HYPRE_Int HYPRE_GMRESGetFinalRelativeResidualNorm(HYPRE_Solver s, HYPRE_Real *r) { if (!is_gmres(s)) return err_arg(1); *r = s->rel_res; return g_error_flag; }
HYPRE_Int HYPRE_ParCSRGMRESSetup(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!is_gmres(s)) return err_arg(1);
  return krylov_setup(s, A, b, x);
}
HYPRE_Int HYPRE_ParCSRGMRESSolve(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!is_gmres(s)) return err_arg(1);
  if (!A) return err_arg(2);
  if (!b || !b->d) return err_arg(3);
  if (!x || !x->d) return err_arg(4);
  NEED_HANDLE();
  b200_amg amg = nullptr;
  const int pk = krylov_precond(s, "GMRES", &amg);
  if (pk < 0) return g_error_flag;
  b200_gmres_params prm;
  prm.tol = s->tol; prm.a_tol = s->a_tol; prm.cf_tol = s->cf_tol; prm.max_iter = s->max_iter; prm.min_iter = s->min_iter;
  prm.k_dim = s->k_dim; prm.rel_change = s->rel_change; prm.skip_real_r_check = s->skip_real_r_check; prm.precond = pk;
  s->norms.assign((size_t)s->max_iter + 2, 0.0);
  b200_timer_start(h);
  const int rc = b200_gmres_solve(h, A->A, amg, &prm, b->d, x->d, &s->num_iterations, &s->rel_res, s->norms.data(), &s->converged);
  double ms = 0;
  b200_timer_stop_ms(h, &ms);
  s->solve_s = ms * 1e-3;
  if (rc) return err_b200("HYPRE_ParCSRGMRESSolve");
  if (s->print_level > 1) krylov_print_norms(s, h, b);
  if (s->num_iterations >= s->max_iter && !s->converged && s->rel_res > s->tol && s->tol > 0) err(HYPRE_ERROR_CONV);   // gmres.c:785-787
  return g_error_flag;
}
HYPRE_Int HYPRE_GMRESSetup(HYPRE_Solver s, HYPRE_Matrix A, HYPRE_Vector b, HYPRE_Vector x) {
  return HYPRE_ParCSRGMRESSetup(s, (HYPRE_ParCSRMatrix)A, (HYPRE_ParVector)b, (HYPRE_ParVector)x);
}
HYPRE_Int HYPRE_GMRESSolve(HYPRE_Solver s, HYPRE_Matrix A, HYPRE_Vector b, HYPRE_Vector x) {
  return HYPRE_ParCSRGMRESSolve(s, (HYPRE_ParCSRMatrix)A, (HYPRE_ParVector)b, (HYPRE_ParVector)x);
}
// ParCSR-typed aliases (parcsr_ls/HYPRE_parcsr_gmres.c:87-220)
HYPRE_Int HYPRE_ParCSRGMRESSetKDim(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_GMRESSetKDim(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetTol(HYPRE_Solver s, HYPRE_Real v) { return HYPRE_GMRESSetTol(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetAbsoluteTol(HYPRE_Solver s, HYPRE_Real v) { return HYPRE_GMRESSetAbsoluteTol(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetMinIter(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_GMRESSetMinIter(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetMaxIter(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_GMRESSetMaxIter(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetStopCrit(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_GMRESSetStopCrit(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetLogging(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_GMRESSetLogging(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetPrintLevel(HYPRE_Solver s, HYPRE_Int v) { return HYPRE_GMRESSetPrintLevel(s, v); }
HYPRE_Int HYPRE_ParCSRGMRESSetPrecond(HYPRE_Solver s, HYPRE_PtrToParSolverFcn precond, HYPRE_PtrToParSolverFcn precond_setup, HYPRE_Solver ps) {
  return HYPRE_GMRESSetPrecond(s, (HYPRE_PtrToSolverFcn)precond, (HYPRE_PtrToSolverFcn)precond_setup, ps);
}
HYPRE_Int HYPRE_ParCSRGMRESGetPrecond(HYPRE_Solver s, HYPRE_Solver *p) { return HYPRE_GMRESGetPrecond(s, p); }
HYPRE_Int HYPRE_ParCSRGMRESGetNumIterations(HYPRE_Solver s, HYPRE_Int *n) { return HYPRE_GMRESGetNumIterations(s, n); }
HYPRE_Int HYPRE_ParCSRGMRESGetFinalRelativeResidualNorm(HYPRE_Solver s, HYPRE_Real *r) { return HYPRE_GMRESGetFinalRelativeResidualNorm(s, r); }

HYPRE_Int HYPRE_ParCSRBiCGSTABCreate(MPI_Comm, HYPRE_Solver *solver) { return krylov_create(solver, KIND_BICGSTAB); }
HYPRE_Int HYPRE_ParCSRBiCGSTABDestroy(HYPRE_Solver s) {
  if (!is_bicgstab(s)) return err_arg(1);
  delete s;
  return g_error_flag;
}
HYPRE_Int HYPRE_BiCGSTABDestroy(HYPRE_Solver s) { return HYPRE_ParCSRBiCGSTABDestroy(s); }
#define BICG_SET(NAME, TYPE, FIELD) \
  HYPRE_Int HYPRE_BiCGSTABSet##NAME(HYPRE_Solver s, TYPE v) { if (!is_bicgstab(s)) return err_arg(1); s->FIELD = v; return g_error_flag; } \
  HYPRE_Int HYPRE_ParCSRBiCGSTABSet##NAME(HYPRE_Solver s, TYPE v) { return HYPRE_BiCGSTABSet##NAME(s, v); }
BICG_SET(Tol, HYPRE_Real, tol)
BICG_SET(AbsoluteTol, HYPRE_Real, a_tol)
BICG_SET(MinIter, HYPRE_Int, min_iter)
BICG_SET(MaxIter, HYPRE_Int, max_iter)
BICG_SET(StopCrit, HYPRE_Int, stop_crit)
BICG_SET(Logging, HYPRE_Int, logging)
BICG_SET(PrintLevel, HYPRE_Int, print_level)
#undef BICG_SET
HYPRE_Int HYPRE_BiCGSTABSetConvergenceFactorTol(HYPRE_Solver s, HYPRE_Real v) { if (!is_bicgstab(s)) return err_arg(1); s->cf_tol = v; return g_error_flag; }
HYPRE_Int HYPRE_BiCGSTABSetPrecond(HYPRE_Solver s, HYPRE_PtrToSolverFcn precond, HYPRE_PtrToSolverFcn precond_setup, HYPRE_Solver precond_solver) {
  if (!is_bicgstab(s)) return err_arg(1);
  s->precond = precond; s->precond_setup = precond_setup; s->precond_solver = precond_solver;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRBiCGSTABSetPrecond(HYPRE_Solver s, HYPRE_PtrToParSolverFcn precond, HYPRE_PtrToParSolverFcn precond_setup, HYPRE_Solver ps) {
  return HYPRE_BiCGSTABSetPrecond(s, (HYPRE_PtrToSolverFcn)precond, (HYPRE_PtrToSolverFcn)precond_setup, ps);
}
HYPRE_Int HYPRE_BiCGSTABGetPrecond(HYPRE_Solver s, HYPRE_Solver *precond_data) {
  if (!is_bicgstab(s)) return err_arg(1);
  *precond_data = s->precond_solver;
  return g_error_flag;
}
HYPRE_Int HYPRE_ParCSRBiCGSTABGetPrecond(HYPRE_Solver s, HYPRE_Solver *p) { return HYPRE_BiCGSTABGetPrecond(s, p); }
HYPRE_Int HYPRE_BiCGSTABGetNumIterations(HYPRE_Solver s, HYPRE_Int *n) { if (!is_bicgstab(s)) return err_arg(1); *n = s->num_iterations; return g_error_flag; }
HYPRE_Int HYPRE_BiCGSTABGetFinalRelativeResidualNorm(HYPRE_Solver s, HYPRE_Real *r) { if (!is_bicgstab(s)) return err_arg(1); *r = s->rel_res; return g_error_flag; }
HYPRE_Int HYPRE_ParCSRBiCGSTABGetNumIterations(HYPRE_Solver s, HYPRE_Int *n) { return HYPRE_BiCGSTABGetNumIterations(s, n); }
HYPRE_Int HYPRE_ParCSRBiCGSTABGetFinalRelativeResidualNorm(HYPRE_Solver s, HYPRE_Real *r) { return HYPRE_BiCGSTABGetFinalRelativeResidualNorm(s, r); }
HYPRE_Int HYPRE_ParCSRBiCGSTABSetup(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!is_bicgstab(s)) return err_arg(1);
  return krylov_setup(s, A, b, x);
}
HYPRE_Int HYPRE_ParCSRBiCGSTABSolve(HYPRE_Solver s, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x) {
  if (!is_bicgstab(s)) return err_arg(1);
  if (!A) return err_arg(2);
  if (!b || !b->d) return err_arg(3);
  if (!x || !x->d) return err_arg(4);
  NEED_HANDLE();
  b200_amg amg = nullptr;
  const int pk = krylov_precond(s, "BiCGSTAB", &amg);
  if (pk < 0) return g_error_flag;
  b200_bicgstab_params prm;
  prm.tol = s->tol; prm.a_tol = s->a_tol; prm.cf_tol = s->cf_tol; prm.max_iter = s->max_iter; prm.min_iter = s->min_iter;
  prm.stop_crit = s->stop_crit; prm.precond = pk;
  s->norms.assign((size_t)s->max_iter + 2, 0.0);
  b200_timer_start(h);
  const int rc = b200_bicgstab_solve(h, A->A, amg, &prm, b->d, x->d, &s->num_iterations, &s->rel_res, s->norms.data(), &s->converged);
  double ms = 0;
  b200_timer_stop_ms(h, &ms);
  s->solve_s = ms * 1e-3;
  if (rc) return err_b200("HYPRE_ParCSRBiCGSTABSolve");
  if (s->print_level > 0) krylov_print_norms(s, h, b);
  if (s->num_iterations >= s->max_iter && !s->converged && s->rel_res > s->tol && s->tol > 0) err(HYPRE_ERROR_CONV);   // bicgstab.c:528
  return g_error_flag;
}
HYPRE_Int HYPRE_BiCGSTABSetup(HYPRE_Solver s, HYPRE_Matrix A, HYPRE_Vector b, HYPRE_Vector x) {
  return HYPRE_ParCSRBiCGSTABSetup(s, (HYPRE_ParCSRMatrix)A, (HYPRE_ParVector)b, (HYPRE_ParVector)x);
}
HYPRE_Int HYPRE_BiCGSTABSolve(HYPRE_Solver s, HYPRE_Matrix A, HYPRE_Vector b, HYPRE_Vector x) {
  return HYPRE_ParCSRBiCGSTABSolve(s, (HYPRE_ParCSRMatrix)A, (HYPRE_ParVector)b, (HYPRE_ParVector)x);
}
// times and residual history of the last GMRES / BiCGSTAB solve (same accessors as the PCG ones)
HYPRE_Int HYPRE_b200_KrylovGetTimes(HYPRE_Solver s, HYPRE_Real *setup_s, HYPRE_Real *solve_s) {
  if (!is_gmres(s) && !is_bicgstab(s) && !is_pcg(s)) return err_arg(1);
  if (setup_s) *setup_s = s->setup_s;
  if (solve_s) *solve_s = s->solve_s;
  return g_error_flag;
}
HYPRE_Int HYPRE_b200_KrylovGetResidualNorms(HYPRE_Solver s, HYPRE_Int n, HYPRE_Real *norms) {
  if (!is_gmres(s) && !is_bicgstab(s) && !is_pcg(s)) return err_arg(1);
  for (int i = 0; i < n && i < (int)s->norms.size(); i++) norms[i] = s->norms[i];
  return g_error_flag;
}

}  // extern "C"
