// b200_amg.cu -- BoomerAMG hierarchy driver, V-cycle and PCG on the device.
//
// Reference: hypre_BoomerAMGSetup (parcsr_ls/par_amg_setup.c:27-3518), hypre_BoomerAMGCycle
// (par_cycle.c:22-641), l1-Jacobi hypre_ParCSRRelax type 1 (ams.c:41-100), hypre_GaussElimSetup /
// Solve (par_gauss_elim.c:20-330, gselim.h), hypre_PCGSolve (krylov/pcg.c:271-757).
//
// In-scope configuration (everything else is rejected loudly at setup):
//   coarsen_type 8 (PMIS), interp_type 6 (ext+i) with trunc_factor / P_max_elmts,
//   Galerkin product by two SpGEMMs (hypre_ParCSRMatrixRAPKT, mod_rap2 path),
//   relax_type 18 (l1-Jacobi, relax_order 0) on all levels, relax 9 (Gaussian elimination)
//   on the coarsest, V(1,1) cycle, explicit restriction R = P^T (keepTranspose semantics).
#include "b200_internal.h"
#include <chrono>
#include <cmath>
#include <map>

int b200_csr_spmv_epi(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha,
                      double beta, const double *b, const double *d);
int b200_pmis_rows(b200_handle h, b200_csr S, int seed, long long first_row, int *d_cf, int *iterations);
int b200_vec_dot_dev(b200_handle h, int n, const double *x, const double *y, double *d_out);
extern "C" int b200_agg_coarsen(b200_handle h, b200_csr S, int seed, int *d_cf);
extern "C" int b200_multipass_interp(b200_handle h, b200_csr A, b200_csr S, int *d_cf, b200_csr *P);

struct b200_level {
  b200_csr A = nullptr;     // owned except level 0 (borrowed from the caller's ParCSR diag block)
  b200_csr As = nullptr;    // solve-phase operator: A itself on level 0, a column-sorted copy on the coarse levels
  b200_parcsr Apar = nullptr;   // one-rank ParCSR view of A for the CG smoother (relax 15)
  b200_csr P = nullptr;     // interpolation to this level from the next coarser one
  b200_csr R = nullptr;     // P^T
  b200_csr S = nullptr;     // kept only when KeepS
  int *cf = nullptr;        // CF marker {1,-1}
  double *l1 = nullptr;     // l1 norms
  b200_cheby_s *cheby = nullptr;   // Chebyshev smoother data (relax 16)
  double *F = nullptr, *U = nullptr, *T = nullptr;   // rhs, iterate, ping-pong iterate (levels >= 1; T also level 0)
  int n = 0;
};

struct b200_amg_s {
  std::map<std::string, int> ip;
  std::map<std::string, double> rp;
  std::vector<b200_level> lv;
  double *Vtemp = nullptr;        // residual scratch, size of level 0
  double *ge_A = nullptr;         // dense coarsest matrix (row-major n x n) + work copy + rhs
  int ge_n = 0;
  bool coarse_ge = false;
  bool gs = false;                // Gauss-Seidel family smoother (relax 3/4/6/8/13/14) instead of l1-Jacobi
  bool cf_l1 = false;             // l1-Jacobi in C/F order (relax 18, relax_order 1: par_cycle.c:397-416)
  int relax_down = 18, relax_up = 18;
  int ns[4] = {1, 1, 1, 1};       // num_grid_sweeps[1..3]: down, up, coarsest (par_amg.c:1934-2030)
  int cycle_type = 1, fcycle = 0; // 1 = V, 2 = W (par_cycle.c:199-210); F-cycle flag
  bool general_cycle = false;     // anything other than V(1,1) with one coarse sweep
  double times[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool is_setup = false;
  b200_amg_s() {
    // library defaults that matter on this path = the values ij.c passes for `-solver 1`
    // (test/ij.c:203-330, :1181-1205) with the north-star choices -pmis -rlx 18 -mod_rap2 1
    ip = {{"CoarsenType", 8}, {"InterpType", 6}, {"PMaxElmts", 4}, {"RelaxType", 18}, {"MaxLevels", 25},
          {"MaxCoarseSize", 9}, {"MinCoarseSize", 0}, {"NumSweeps", 1}, {"AggNumLevels", 0}, {"ModuleRAP2", 1},
          {"RAP2", 0}, {"KeepTranspose", 1}, {"RelaxOrder", 0}, {"MaxIter", 1}, {"CycleType", 1},
          {"NumFunctions", 1}, {"MinIter", 0}, {"RelaxTypeUp", -1}, {"GSBlocks", 1}, {"ChebyOrder", 2}, {"ChebyEigEst", 10},
          {"ChebyVariant", 0}, {"ChebyScale", 1}, {"KeepS", 0}, {"PrintLevel", 0}, {"Seed", 2747},
          {"NumSweepsDown", -1}, {"NumSweepsUp", -1}, {"NumSweepsCoarse", 1}, {"FCycle", 0}, {"SeqThreshold", 0},
          {"UserRelaxType", 0}};   // -1: the caller never chose a smoother (hypre_ParAMGDataUserRelaxType, par_amg.c:233)
    rp = {{"StrongThreshold", 0.25}, {"MaxRowSum", 1.0}, {"TruncFactor", 0.0}, {"RelaxWt", 1.0},
          {"OuterWt", 1.0}, {"Tol", 0.0}, {"ChebyFraction", 0.3}};
  }
};

namespace {

__global__ void fix_cf_kernel(int n, int *cf) {                     // par_lr_interp.c:1888-1894
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cf[i] == -3) cf[i] = -1;
}
__global__ void count_c_kernel(int n, const int *__restrict__ cf, int *count) {   // par_coarse_parms.c:83-86
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int g = (i < n && cf[i] == 1) ? 1 : 0;
  unsigned b = __ballot_sync(0xffffffffu, g);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}
__global__ void dense_from_csr_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                      const double *__restrict__ A_a, double *__restrict__ M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < n; k++) M[i * n + k] = 0.0;
  for (int jj = A_i[i]; jj < A_i[i + 1]; jj++) M[i * n + A_j[jj]] = A_a[jj];     // par_gauss_elim.c:100-115
}
// hypre_gselim (sstruct_ls/gselim.h) on a scratch copy; one thread, n <= MaxCoarseSize
__global__ void gselim_kernel(int n, const double *__restrict__ A_mat, double *__restrict__ A, const double *__restrict__ f,
                              double *__restrict__ x) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < n * n; i++) A[i] = A_mat[i];
  for (int i = 0; i < n; i++) x[i] = f[i];
  if (n == 1) {
    if (A[0] != 0.0) x[0] = x[0] / A[0];
    return;
  }
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
      for (int j = 0; j < k; j++)
        if (A[j * n + k] != 0.0) x[j] -= x[k] * A[j * n + k];
    }
  }
  if (A[0] != 0.0) x[0] /= A[0];
}
// u = f / l1  (l1-Jacobi sweep from a zero iterate: u + w*(f - A*0)/l1, exact)
__global__ void jacobi_zero_kernel(size_t n, double w, const double *__restrict__ f, const double *__restrict__ l1,
                                   double *__restrict__ u) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) u[i] = 0.0 + w * f[i] / l1[i];
}
// PCG vector updates with device-resident scalars (no host round trip between kernels)
// sc[0]=gamma sc[1]=sdotp sc[2]=gamma_old sc[3]=i_prod sc[4]=alpha sc[5]=beta
__global__ void pcg_alpha_kernel(double *sc) {
  sc[4] = (sc[1] != 0.0) ? sc[0] / sc[1] : 0.0;   // alpha = gamma / <s,p> (pcg.c:522); <s,p> = 0 is the error path (:516-521): x, r stay intact
  sc[2] = sc[0];                  // gamma_old = gamma             (pcg.c:530)
}
__global__ void pcg_update_xr_kernel(size_t n, const double *__restrict__ sc, const double *__restrict__ p,
                                     const double *__restrict__ s, double *__restrict__ x, double *__restrict__ r) {
  const double alpha = sc[4];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    x[i] += alpha * p[i];         // pcg.c:534
    r[i] += -alpha * s[i];        // pcg.c:539
  }
}
__global__ void pcg_beta_kernel(double *sc) { sc[5] = sc[0] / sc[2]; }   // beta = gamma / gamma_old (pcg.c:729)
__global__ void pcg_update_p_kernel(size_t n, const double *__restrict__ sc, const double *__restrict__ s,
                                    double *__restrict__ p) {
  const double beta = sc[5];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = beta * p[i] + 1.0 * s[i];   // Scale then Axpy (pcg.c:734-735)
}

inline int vgrid(b200_handle h, size_t n) {
  size_t g = (n + 255) / 256, cap = (size_t)h->num_sm * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

struct PhaseTimer {
  b200_handle h;
  cudaEvent_t a, b;
  PhaseTimer(b200_handle hh) : h(hh) { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~PhaseTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
  void start() { cudaEventRecord(a, h->stream); }
  double stop() {
    cudaEventRecord(b, h->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  }
};

}  // namespace

extern "C" int b200_amg_create(b200_amg *out) {
  *out = new b200_amg_s();
  return 0;
}

static int free_levels(b200_handle h, b200_amg amg) {
  for (size_t l = 0; l < amg->lv.size(); l++) {
    b200_level &L = amg->lv[l];
    if (L.As && L.As != L.A) B200_TRY(b200_csr_destroy(h, L.As));
    if (L.Apar) { B200_TRY(b200_csr_destroy(h, L.Apar->offd)); delete L.Apar; L.Apar = nullptr; }
    if (l > 0) B200_TRY(b200_csr_destroy(h, L.A));
    B200_TRY(b200_csr_destroy(h, L.P));
    B200_TRY(b200_csr_destroy(h, L.R));
    B200_TRY(b200_csr_destroy(h, L.S));
    B200_TRY(b200_dfree(h, L.cf)); B200_TRY(b200_dfree(h, L.l1));
    B200_TRY(b200_cheby_destroy(h, L.cheby));
    B200_TRY(b200_dfree(h, L.F)); B200_TRY(b200_dfree(h, L.U)); B200_TRY(b200_dfree(h, L.T));
  }
  amg->lv.clear();
  B200_TRY(b200_dfree(h, amg->Vtemp)); amg->Vtemp = nullptr;
  B200_TRY(b200_dfree(h, amg->ge_A)); amg->ge_A = nullptr;
  amg->is_setup = false;
  return 0;
}

extern "C" int b200_amg_destroy(b200_handle h, b200_amg amg) {
  if (!amg) return 0;
  B200_TRY(free_levels(h, amg));
  delete amg;
  return 0;
}

extern "C" int b200_amg_set_int(b200_amg amg, const char *name, int value) {
  if (!amg) B200_FAIL("null amg");
  if (!amg->ip.count(name)) B200_FAIL((std::string("unknown integer parameter ") + name).c_str());
  amg->ip[name] = value;
  return 0;
}
extern "C" int b200_amg_set_real(b200_amg amg, const char *name, double value) {
  if (!amg) B200_FAIL("null amg");
  if (!amg->rp.count(name)) B200_FAIL((std::string("unknown real parameter ") + name).c_str());
  amg->rp[name] = value;
  return 0;
}

int b200_amg_get_int(b200_amg a, const char *name) { return a->ip.count(name) ? a->ip[name] : 0; }
// copy every parameter of src into dst (the replicated coarse-level hierarchy of the row-partitioned setup, b200_dist.cu)
int b200_amg_clone_params(b200_amg src, b200_amg dst) { dst->ip = src->ip; dst->rp = src->rp; return 0; }
double b200_amg_get_real(b200_amg a, const char *name) { return a->rp.count(name) ? a->rp[name] : 0.0; }

extern "C" int b200_amg_setup(b200_handle h, b200_amg amg, b200_parcsr Apar) {
  if (!amg || !Apar) B200_FAIL("amg_setup: null argument");
  if (Apar->offd->ncols > 0) B200_FAIL("amg_setup: multi-rank setup not built yet (offd block must be empty)");
  auto &ip = amg->ip;
  auto &rp = amg->rp;
  if (ip["CoarsenType"] != 8 && ip["CoarsenType"] != 10)
    B200_FAIL("only CoarsenType 8 (PMIS) and 10 (HMIS) are implemented on the B200 path");
  if (ip["InterpType"] != 6) B200_FAIL("only InterpType 6 (extended+i) is implemented on the B200 path");
  // grid_relax_type[1] / [2] (par_amg.c:1650-1672); the coarsest grid is always Gaussian elimination (9)
  int rdown = ip["RelaxType"], rup = ip["RelaxTypeUp"] >= 0 ? ip["RelaxTypeUp"] : ip["RelaxType"];
  auto is_gs = [](int t) { return t == 3 || t == 4 || t == 6 || t == 8 || t == 13 || t == 14; };
  auto is_l1gs = [](int t) { return t == 8 || t == 13 || t == 14; };
  auto is_jac = [](int t) { return t == 18 || t == 7; };
  bool cheby = rdown == 16 && rup == 16;
  // smoothers built from masked Jacobi half-sweeps or from PCG itself (par_cycle.c:397-460): FCF-Jacobi 17, the CG smoother 15,
  // and l1-Jacobi in C/F order (relax 18 with RelaxOrder 1)
  const bool fcf = rdown == 17 && rup == 17, cgs = rdown == 15 && rup == 15;
  const bool cf_l1 = rdown == 18 && rup == 18 && ip["RelaxOrder"] == 1;
  if (fcf || cgs || cf_l1) {
    // handled below: in-place smoothers of the amg_cycle_gs family
  } else
  if (!((is_jac(rdown) && rup == rdown) || cheby || (is_gs(rdown) && is_gs(rup) && is_l1gs(rdown) == is_l1gs(rup))))
    B200_FAIL("RelaxType: the B200 path implements 18 (l1-Jacobi), 7 (weighted Jacobi), 16 (Chebyshev), the l1 hybrid "
              "Gauss-Seidel family 8/13/14 and the hybrid Gauss-Seidel family 3/4/6 (down and up sweeps from the same family)");
  if (!is_jac(rdown) && !cheby && !fcf && !cgs && rp["RelaxWt"] != 1.0) B200_FAIL("Gauss-Seidel smoothers: only relax_weight 1 is implemented");
  if (ip["GSBlocks"] < 1) B200_FAIL("GSBlocks must be >= 1");
  amg->gs = !is_jac(rdown) || cf_l1;         // in-place smoothers (Gauss-Seidel family, Chebyshev, 15, 17, C/F l1-Jacobi) use amg_cycle_gs
  amg->relax_down = rdown; amg->relax_up = rup;
  amg->cf_l1 = cf_l1;
  if (ip["RelaxOrder"] != 0 && !cf_l1 && !fcf && !cgs && !is_jac(rdown))
    B200_FAIL("RelaxOrder 1 (C/F relaxation) is implemented for the l1-Jacobi smoother (RelaxType 18) only");
  if (ip["AggNumLevels"] < 0) B200_FAIL("AggNumLevels must be >= 0");
  // aggressive levels: second PMIS on the distance-two graph + multipass interpolation (agg_interp_type 4, the
  // reference default; agg_trunc_factor = agg_P_max_elmts = 0, num_paths 1), par_amg_setup.c:1239-1256, :1590-1605
  if (ip["NumSweeps"] < 1) B200_FAIL("NumSweeps must be >= 1");                      // par_amg.c:1947-1951
  if (ip["CycleType"] < 1) B200_FAIL("CycleType must be >= 1 (1 = V, 2 = W)");
  amg->ns[1] = ip["NumSweepsDown"] >= 0 ? ip["NumSweepsDown"] : ip["NumSweeps"];
  amg->ns[2] = ip["NumSweepsUp"] >= 0 ? ip["NumSweepsUp"] : ip["NumSweeps"];
  amg->ns[3] = ip["NumSweepsCoarse"];
  if (amg->ns[3] < 0) B200_FAIL("NumSweepsCoarse must be >= 0");
  amg->cycle_type = ip["CycleType"]; amg->fcycle = ip["FCycle"] ? 1 : 0;
  amg->general_cycle = amg->ns[1] != 1 || amg->ns[2] != 1 || amg->ns[3] != 1 || amg->cycle_type != 1 || amg->fcycle;
  if (ip["NumFunctions"] != 1) B200_FAIL("only scalar problems (NumFunctions 1)");
  if (ip["RAP2"] != 0 || (ip["ModuleRAP2"] != 0 && ip["ModuleRAP2"] != 1))
    B200_FAIL("Galerkin product: ModuleRAP2 1 (hypre_ParCSRMatrixRAPKT, R(AP)) or ModuleRAP2 0 (the fused "
              "hypre_BoomerAMGBuildCoarseOperatorKT order, (RA)P) with RAP2 0");
  B200_TRY(free_levels(h, amg));
  for (double &t : amg->times) t = 0;
  PhaseTimer tm(h), total(h);
  total.start();

  const double theta = rp["StrongThreshold"], mrs = rp["MaxRowSum"], trunc = rp["TruncFactor"];
  const int pmax = ip["PMaxElmts"], max_levels = ip["MaxLevels"], max_coarse = ip["MaxCoarseSize"];
  const int min_coarse = ip["MinCoarseSize"], keepS = ip["KeepS"], seed = ip["Seed"];

  b200_level L0;
  L0.A = Apar->diag;
  L0.n = Apar->diag->nrows;
  amg->lv.push_back(L0);
  int level = 0;
  bool not_finished = max_levels > 1, stalled = false;
  int *d_count = nullptr;
  B200_TRY(b200_dalloc<int>(h, &d_count, 1));
  const bool trace = getenv("B200_TRACE") != nullptr;
  auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tlevel0 = 0;
  if (trace) { cudaStreamSynchronize(h->stream); tlevel0 = wall(); }
  while (not_finished) {                                   // par_amg_setup.c:889
    if (trace && level > 0) {
      cudaStreamSynchronize(h->stream);
      const double t = wall();
      fprintf(stderr, "[b200 trace] single-GPU setup level %d -> %d: %.3f ms (rows %d)\n", level - 1, level, t - tlevel0, amg->lv[level - 1].n);
      tlevel0 = t;
    }
    b200_level &L = amg->lv[level];
    const int fine_size = L.n;
    b200_csr S = nullptr;
    tm.start();
    B200_TRY(b200_strength(h, L.A, theta, mrs, &S));       // :1035
    amg->times[0] += tm.stop();
    int *cf = nullptr;
    B200_TRY(b200_dalloc<int>(h, &cf, fine_size));
    tm.start();
    if (ip["CoarsenType"] == 10) B200_TRY(b200_hmis(h, S, seed, cf));   // :1107 (sequential first pass, see b200_hmis.cu)
    else B200_TRY(b200_pmis_rows(h, S, seed, 0, cf, nullptr));          // :1114
    const bool aggressive = level < ip["AggNumLevels"];
    if (aggressive && ip["CoarsenType"] == 10) {
      // HMIS second coarsening (measure_type + 3) is not built; a first pass that found no C point at all (no strong
      // connections: TEST_ij/coarsening.jobs job 14) never gets there, so only a real second pass is refused
      B200_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), h->stream));
      count_c_kernel<<<b200_grid(fine_size, 256), 256, 0, h->stream>>>(fine_size, cf, d_count);
      B200_LAUNCH_CHECK();
      int c1 = 0;
      B200_CUDA(cudaMemcpyAsync(&c1, d_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      B200_CUDA(cudaStreamSynchronize(h->stream));
      if (c1 > 0) B200_FAIL("CoarsenType 10 (HMIS) with aggressive levels is not built on the B200 path: use CoarsenType 8 (PMIS)");
    } else if (aggressive) B200_TRY(b200_agg_coarsen(h, S, seed, cf));   // :1239-1256 + CorrectCFMarker :1592
    amg->times[1] += tm.stop();
    B200_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), h->stream));
    count_c_kernel<<<b200_grid(fine_size, 256), 256, 0, h->stream>>>(fine_size, cf, d_count);
    B200_LAUNCH_CHECK();
    int coarse_size = 0;
    B200_CUDA(cudaMemcpyAsync(&coarse_size, d_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    if (coarse_size == 0 || coarse_size == fine_size || coarse_size < min_coarse) {   // :1487-1560
      // no coarse grid: stop coarsening, "and set the coarsest solve to be a single sweep of default smoother or smoother
      // set by user" (:1484-1497: grid_relax_type[3] = grid_relax_type[0], one sweep) instead of Gaussian elimination
      if (coarse_size == 0 || coarse_size == fine_size) stalled = true;
      B200_TRY(b200_csr_destroy(h, S));
      B200_TRY(b200_dfree(h, cf));
      break;
    }
    tm.start();
    b200_csr P = nullptr;
    if (aggressive) B200_TRY(b200_multipass_interp(h, L.A, S, cf, &P));   // :1601
    else B200_TRY(b200_extpi_interp(h, L.A, S, cf, trunc, pmax, &P));     // :1989
    amg->times[2] += tm.stop();
    fix_cf_kernel<<<b200_grid(fine_size, 256), 256, 0, h->stream>>>(fine_size, cf);
    B200_LAUNCH_CHECK();
    L.cf = cf;
    L.P = P;
    if (keepS) L.S = S; else B200_TRY(b200_csr_destroy(h, S));
    // Galerkin product.  ModuleRAP2 1: hypre_ParCSRMatrixRAPKTHost single-rank branch (par_csr_triplemat.c:872-888):
    //   Q = A*P ; RT = P^T ; C = RT*Q.   ModuleRAP2 0 (library default): hypre_BoomerAMGBuildCoarseOperatorKT forms
    //   row ic of R*A first (par_rap.c:1640-1700) and multiplies it by P with the diagonal entry created first
    //   (:1546-1553, :1790-1857), i.e. the same two Gustavson products associated the other way: C = (RT*A)*P
    tm.start();
    b200_csr R = nullptr;
    B200_TRY(b200_csr_transpose(h, P, &R));
    amg->times[4] += tm.stop();
    tm.start();
    b200_csr Q = nullptr, AH = nullptr;
    if (ip["ModuleRAP2"] == 1) {
      B200_TRY(b200_csr_multiply(h, L.A, P, &Q));
      B200_TRY(b200_csr_multiply(h, R, Q, &AH));
    } else {
      B200_TRY(b200_csr_multiply(h, R, L.A, &Q));
      B200_TRY(b200_csr_multiply(h, Q, P, &AH));
    }
    B200_TRY(b200_csr_destroy(h, Q));
    amg->times[5] += tm.stop();
    L.R = R;
    b200_level Ln;
    Ln.A = AH;
    Ln.n = AH->nrows;
    amg->lv.push_back(Ln);
    ++level;
    if (level == max_levels - 1 || coarse_size <= max_coarse) not_finished = false;   // :2884-2888
    if (not_finished && (double)coarse_size >= 0.75 * (double)fine_size)              // :2873-2877
      B200_FAIL("coarsening stalled (coarse >= 0.75 fine): the reference switches to CLJP here, which is out of scope");
  }
  B200_TRY(b200_dfree(h, d_count));

  const int nl = (int)amg->lv.size();
  // coarsest level: Gaussian elimination if small enough, else fall back to the smoother (:2909-2921)
  b200_level &Lc = amg->lv[nl - 1];
  amg->coarse_ge = Lc.n <= max_coarse && Lc.n > 0 && !stalled;
  if (nl == 1) {
    // "If no coarsening occurred, apply a simple smoother once ... use the user relax type (instead of 0)", 6 when the user
    // chose none (par_cycle.c:289-300): never Gaussian elimination on a one-level hierarchy
    const int t = ip["UserRelaxType"] == -1 ? 6 : rdown;
    rdown = rup = t;
    amg->relax_down = amg->relax_up = t;
    amg->gs = !is_jac(t);
    amg->coarse_ge = false;
    amg->general_cycle = false;
    cheby = (t == 16);
  }
  if (amg->coarse_ge) {
    amg->ge_n = Lc.n;
    B200_TRY(b200_dalloc<double>(h, &amg->ge_A, (size_t)2 * Lc.n * Lc.n));
    dense_from_csr_kernel<<<b200_grid(Lc.n, 128), 128, 0, h->stream>>>(Lc.n, Lc.A->i, Lc.A->j, Lc.A->a, amg->ge_A);
    B200_LAUNCH_CHECK();
  }
  // l1 norms: option 1 for relax 18, option 4 for 8/13/14 (:3018-3060); Gauss-Seidel level schedules
  tm.start();
  static const bool no_sort = [] { const char *e = getenv("B200_NO_SORTED_COPY"); return e && e[0] == '1'; }();
  for (int l = 0; l < nl; l++) {
    b200_level &L = amg->lv[l];
    L.As = L.A;
    if (l > 0 && l < nl - 1 && !no_sort && (double)L.A->nnz > 12.0 * L.n) B200_TRY(b200_csr_sorted_copy(h, L.A, &L.As));
    // streaming-SpMV plans for exactly the operators the cycle applies (setup products build none)
    if (!L.As->blk_row) B200_TRY(b200_csr_build_plan(h, L.As));
    if (L.P && !L.P->blk_row) B200_TRY(b200_csr_build_plan(h, L.P));
    if (L.R && !L.R->blk_row) B200_TRY(b200_csr_build_plan(h, L.R));
    if (l < nl - 1 || !amg->coarse_ge || rdown == 7) {
      if (!amg->gs || is_l1gs(rdown) || amg->cf_l1) {  // option 1: relax 18, 4: relax 8/13/14, 5 (= the diagonal): relax 7
        const bool l1gs = amg->gs && !amg->cf_l1;
        B200_TRY(b200_dalloc<double>(h, &L.l1, L.n));
        if (amg->cf_l1 && l < nl - 1 && L.cf) B200_TRY(b200_l1_norms_cf(h, L.A, L.cf, L.l1));      // par_amg_setup.c:3047-3050
        else B200_TRY(b200_l1_norms_blocks(h, L.A, l1gs ? 4 : (rdown == 7 ? 5 : 1), l1gs ? ip["GSBlocks"] : 1, L.l1));
      }
    }
    if (cheby && (l < nl - 1 || !amg->coarse_ge))           // par_amg_setup.c:3137-3160
      B200_TRY(b200_cheby_setup(h, L.A, ip["ChebyEigEst"], ip["ChebyOrder"], rp["ChebyFraction"], ip["ChebyVariant"],
                                ip["ChebyScale"], &L.cheby));
    if (rdown == 15 && (l < nl - 1 || !amg->coarse_ge)) {       // CG smoother: PCG on this level's operator (par_amg_setup.c:3165-3180)
      L.Apar = new b200_parcsr_s();
      L.Apar->global_rows = L.Apar->global_cols = L.n;
      L.Apar->diag = L.A;
      B200_TRY(b200_csr_alloc(h, L.n, 0, 0, true, &L.Apar->offd));
      B200_CUDA(cudaMemsetAsync(L.Apar->offd->i, 0, sizeof(int) * ((size_t)L.n + 1), h->stream));
      if (!L.A->blk_row) B200_TRY(b200_csr_build_plan(h, L.A));
    }
    const bool masked = rdown == 17 || rdown == 15 || amg->cf_l1;
    if ((l < nl - 1 || !amg->coarse_ge) && !cheby && !masked) {
      if (amg->gs) {
        if (L.A->gs && b200_gs_plan_blocks(L.A->gs) != ip["GSBlocks"]) { B200_TRY(b200_gs_plan_destroy(h, L.A->gs)); L.A->gs = nullptr; }
        if (!L.A->gs) B200_TRY(b200_gs_plan_create(h, L.A, ip["GSBlocks"], &L.A->gs));
      }
    }
    if (l > 0) B200_TRY(b200_dalloc<double>(h, &L.F, L.n));
    B200_TRY(b200_dalloc<double>(h, &L.U, L.n));
    B200_TRY(b200_dalloc<double>(h, &L.T, L.n));
  }
  B200_TRY(b200_dalloc<double>(h, &amg->Vtemp, amg->lv[0].n));
  amg->times[6] += tm.stop();
  amg->times[7] = total.stop();
  amg->is_setup = true;
  return 0;
}

extern "C" int b200_amg_num_levels(b200_amg amg) { return amg ? (int)amg->lv.size() : 0; }
extern "C" b200_csr b200_amg_level_A(b200_amg amg, int l) { return (amg && l >= 0 && l < (int)amg->lv.size()) ? amg->lv[l].A : nullptr; }
extern "C" b200_csr b200_amg_level_P(b200_amg amg, int l) { return (amg && l >= 0 && l < (int)amg->lv.size()) ? amg->lv[l].P : nullptr; }
extern "C" b200_csr b200_amg_level_S(b200_amg amg, int l) { return (amg && l >= 0 && l < (int)amg->lv.size()) ? amg->lv[l].S : nullptr; }
extern "C" const int *b200_amg_level_CF(b200_amg amg, int l) { return (amg && l >= 0 && l < (int)amg->lv.size()) ? amg->lv[l].cf : nullptr; }
extern "C" const double *b200_amg_level_l1(b200_amg amg, int l) { return (amg && l >= 0 && l < (int)amg->lv.size()) ? amg->lv[l].l1 : nullptr; }
extern "C" int b200_amg_setup_times(b200_amg amg, double *t) {
  if (!amg) B200_FAIL("null amg");
  for (int i = 0; i < 8; i++) t[i] = amg->times[i];
  return 0;
}

// One l1-Jacobi sweep u_out = u_in + w (f - A u_in) / l1   (ams.c:72-92, fused into one pass over A)
static int jacobi(b200_handle h, b200_level &L, double w, const double *f, const double *u_in, double *u_out) {
  return b200_csr_spmv_epi(h, L.As, u_in, u_out, 1, w, 0.0, f, L.l1);
}

// One V(1,1) cycle (par_cycle.c:255-622). u_zero: the caller guarantees u == 0 on entry
// (PCG clears the vector before every preconditioner application, pcg.c:434,:568), which lets
// the first sweep on every level skip its SpMV: u + (f - A*0)/l1 == f/l1 exactly.
// one in-place relaxation call: Gauss-Seidel family or Chebyshev
// One Jacobi half-sweep over the rows whose CF marker equals `pt` (pt = 0: every row), reading the iterate from before the
// sweep (`old`, the reference's Vtemp copy) and updating u in place; one thread per row, the row summed in storage order.
//   mode 0, weighted Jacobi (hypre_BoomerAMGRelax type 0, par_relax.c:139-248): u_i = (1 - w) u_i + w (f_i - sum_{j != i} a_ij old_j) / a_ii
//   mode 1, l1-Jacobi (hypre_ParCSRRelax_L1_Jacobi, par_relax_more.c:991-1161): u_i += w (f_i - sum_j a_ij old_j) / l1_i
__global__ void masked_jacobi_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                                     const int *__restrict__ cf, int pt, int mode, double w, const double *__restrict__ f,
                                     const double *__restrict__ l1, const double *__restrict__ old, double *__restrict__ u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pt != 0 && cf[i] != pt) return;
  const int b = A_i[i], e = A_i[i + 1];
  const double diag = A_a[b];
  if (diag == 0.0) return;
  double res = f[i];
  for (int jj = b + (mode == 0 ? 1 : 0); jj < e; jj++) res -= A_a[jj] * old[A_j[jj]];
  if (mode == 0) {
    double v = u[i] * (1.0 - w);
    v += w * res / diag;
    u[i] = v;
  } else {
    u[i] += (w * res) / l1[i];
  }
}

static int masked_sweep(b200_handle h, b200_level &L, int pt, int mode, double w, const double *f, double *u) {
  B200_CUDA(cudaMemcpyAsync(L.T, u, sizeof(double) * (size_t)L.n, cudaMemcpyDeviceToDevice, h->stream));      // Vtemp = u
  masked_jacobi_kernel<<<b200_grid(L.n, 128), 128, 0, h->stream>>>(L.n, L.A->i, L.A->j, L.A->a, L.cf, pt, mode, w, f, L.l1, L.T, u);
  B200_LAUNCH_CHECK();
  return 0;
}

// phase: 1 down sweep, 2 up sweep, 3 coarsest grid (par_cycle.c cycle_param)
static int gs_relax(b200_handle h, b200_amg amg, b200_level &L, int type, const double *f, double *u, bool zero, int phase) {
  if (type == 16) return b200_cheby_solve(h, L.cheby, L.As, zero, f, u);      // par_cycle.c:440-452
  if (type == 17 || type == 15 || (type == 18 && amg->cf_l1)) {
    const double w = amg->rp["RelaxWt"];
    if (zero) B200_CUDA(cudaMemsetAsync(u, 0, sizeof(double) * (size_t)L.n, h->stream));
    if (type == 17) {                                                         // :451-464, hypre_BoomerAMGRelax_FCFJacobi
      if (phase == 3 || !L.cf) return masked_sweep(h, L, 0, 0, w, f, u);      // coarsest grid: one plain Jacobi sweep
      B200_TRY(masked_sweep(h, L, -1, 0, w, f, u));
      B200_TRY(masked_sweep(h, L, 1, 0, w, f, u));
      return masked_sweep(h, L, -1, 0, w, f, u);
    }
    if (type == 18) {                                                         // :397-416: C then F going down, F then C going up
      if (phase == 3 || !L.cf) return masked_sweep(h, L, 0, 1, w, f, u);
      B200_TRY(masked_sweep(h, L, phase == 1 ? 1 : -1, 1, w, f, u));
      return masked_sweep(h, L, phase == 1 ? -1 : 1, 1, w, f, u);
    }
    // CG smoother (:439-445, hypre_ParCSRRelax_CG): num_sweep (= 1) iterations of unpreconditioned PCG in the 2-norm from u
    b200_pcg_params prm;
    prm.tol = 0.0; prm.a_tol = 0.0; prm.max_iter = 1; prm.two_norm = 1; prm.rel_change = 0; prm.recompute_residual = 0; prm.precond = 0;
    int its = 0;
    double rel = 0;
    const int rc = b200_pcg_solve_ex(h, L.Apar, nullptr, &prm, f, u, &its, &rel, nullptr);
    if (rc) {      // an iterate that is already exact stops the inner recurrence ("Zero sdotp" / "Subnormal gamma": the reference
      const std::string msg = b200_last_error();          // flags it and carries on with u as it is, pcg.c:516-521, :680-690)
      if (msg.find("sdotp") != std::string::npos || msg.find("gamma") != std::string::npos) return 0;
    }
    return rc;
  }
  return b200_gs_relax(h, L.A->gs, L.A, type, zero, f, L.l1, u);
}

// V(1,1) cycle with in-place Gauss-Seidel smoothing (par_cycle.c:255-622)
static int amg_cycle_gs(b200_handle h, b200_amg amg, const double *f, double *u, bool u_zero) {
  const int nl = (int)amg->lv.size();
  std::vector<const double *> F(nl);
  std::vector<double *> U(nl);
  F[0] = f; U[0] = u;
  for (int l = 1; l < nl; l++) { F[l] = amg->lv[l].F; U[l] = amg->lv[l].U; }
  for (int l = 0; l < nl - 1; l++) {
    b200_level &L = amg->lv[l];
    B200_TRY(gs_relax(h, amg, L, amg->relax_down, F[l], U[l], l > 0 || u_zero, 1));
    B200_TRY(b200_csr_spmv_epi(h, L.As, U[l], amg->Vtemp, 0, -1.0, 1.0, F[l], nullptr));             // :549
    B200_TRY(b200_csr_spmv_epi(h, L.R, amg->Vtemp, amg->lv[l + 1].F, 0, 1.0, 0.0, nullptr, nullptr)); // :566
  }
  {
    b200_level &L = amg->lv[nl - 1];
    if (amg->coarse_ge) {
      gselim_kernel<<<1, 32, 0, h->stream>>>(amg->ge_n, amg->ge_A, amg->ge_A + (size_t)amg->ge_n * amg->ge_n, F[nl - 1], U[nl - 1]);
      B200_LAUNCH_CHECK();
    } else {
      B200_TRY(gs_relax(h, amg, L, amg->relax_down, F[nl - 1], U[nl - 1], nl > 1 || u_zero, 3));
    }
  }
  for (int l = nl - 2; l >= 0; l--) {
    b200_level &L = amg->lv[l];
    B200_TRY(b200_csr_spmv_epi(h, L.P, U[l + 1], U[l], 0, 1.0, 1.0, U[l], nullptr));                 // :602
    B200_TRY(gs_relax(h, amg, L, amg->relax_up, F[l], U[l], false, 2));
  }
  return 0;
}

// Any other cycle shape: the reference's level-counter state machine (par_cycle.c:180-622) -- lev_counter[k] visits
// per level (cycle_type: 1 = V, 2 = W; F-cycle flag), num_grid_sweeps[cycle_param] sweeps per visit (1 down, 2 up,
// 3 coarsest).  Jacobi sweeps are out of place (ping-pong between U and T), the other smoothers in place.
static int amg_cycle_general(b200_handle h, b200_amg amg, const double *f, double *u, bool u_zero) {
  const int nl = (int)amg->lv.size();
  const double w = amg->rp["RelaxWt"];
  std::vector<const double *> F(nl);
  std::vector<double *> U(nl), alt(nl);
  std::vector<char> zero(nl, 0);
  std::vector<int> lev_counter(nl);
  F[0] = f; U[0] = u; alt[0] = amg->lv[0].T; zero[0] = u_zero;
  for (int l = 1; l < nl; l++) { F[l] = amg->lv[l].F; U[l] = amg->lv[l].U; alt[l] = amg->lv[l].T; }
  lev_counter[0] = 1;
  for (int k = 1; k < nl; k++) lev_counter[k] = amg->fcycle ? 1 : amg->cycle_type;
  int fcycle_lev = nl - 2, level = 0, cycle_param = 1;
  while (true) {
    b200_level &L = amg->lv[level];
    const int num_sweep = amg->ns[cycle_param];
    const int type = cycle_param == 2 ? amg->relax_up : amg->relax_down;
    for (int j = 0; j < num_sweep; j++) {
      if (level == nl - 1 && amg->coarse_ge) {                       // grid_relax_type[3] = 9
        gselim_kernel<<<1, 32, 0, h->stream>>>(amg->ge_n, amg->ge_A, amg->ge_A + (size_t)amg->ge_n * amg->ge_n, F[level], U[level]);
        B200_LAUNCH_CHECK();
      } else if (amg->gs) {
        B200_TRY(gs_relax(h, amg, L, type, F[level], U[level], zero[level] != 0, cycle_param));
      } else if (zero[level]) {
        jacobi_zero_kernel<<<vgrid(h, L.n), 256, 0, h->stream>>>((size_t)L.n, w, F[level], L.l1, U[level]);
        B200_LAUNCH_CHECK();
      } else {
        B200_TRY(jacobi(h, L, w, F[level], U[level], alt[level]));
        std::swap(U[level], alt[level]);
      }
      zero[level] = 0;
    }
    if (zero[level]) {                                               // no sweep was asked for: the iterate is the zero vector
      B200_CUDA(cudaMemsetAsync(U[level], 0, sizeof(double) * (size_t)L.n, h->stream));
      zero[level] = 0;
    }
    --lev_counter[level];
    if (lev_counter[level] >= 0 && level != nl - 1) {                // :534-591 residual, restriction, coarse iterate = 0
      B200_TRY(b200_csr_spmv_epi(h, L.As, U[level], amg->Vtemp, 0, -1.0, 1.0, F[level], nullptr));
      B200_TRY(b200_csr_spmv_epi(h, L.R, amg->Vtemp, amg->lv[level + 1].F, 0, 1.0, 0.0, nullptr, nullptr));
      ++level;
      zero[level] = 1;
      lev_counter[level] = std::max(lev_counter[level], amg->cycle_type);
      cycle_param = (level == nl - 1) ? 3 : 1;
    } else if (level != 0) {                                         // :592-625 interpolation
      B200_TRY(b200_csr_spmv_epi(h, amg->lv[level - 1].P, U[level], U[level - 1], 0, 1.0, 1.0, U[level - 1], nullptr));
      --level;
      cycle_param = 2;
      if (amg->fcycle && fcycle_lev == level) { lev_counter[level] = std::max(lev_counter[level], 1); fcycle_lev--; }
    } else {
      break;
    }
  }
  for (int l = 1; l < nl; l++) { amg->lv[l].U = U[l]; amg->lv[l].T = alt[l]; }
  if (U[0] != u) B200_TRY(b200_vec_copy(h, amg->lv[0].n, U[0], u));
  return 0;
}

static int amg_cycle(b200_handle h, b200_amg amg, const double *f, double *u, bool u_zero) {
  if (amg->general_cycle && amg->lv.size() > 1) return amg_cycle_general(h, amg, f, u, u_zero);
  if (amg->gs) return amg_cycle_gs(h, amg, f, u, u_zero);
  const int nl = (int)amg->lv.size();
  const double w = amg->rp["RelaxWt"];
  // level 0 buffers: the final post-smoothing sweep must land in the caller's u
  std::vector<const double *> F(nl);
  std::vector<double *> U(nl);          // current iterate per level after pre-smoothing
  F[0] = f;
  for (int l = 1; l < nl; l++) F[l] = amg->lv[l].F;
  if (nl == 1) {
    b200_level &L = amg->lv[0];
    if (amg->coarse_ge) {
      gselim_kernel<<<1, 32, 0, h->stream>>>(amg->ge_n, amg->ge_A, amg->ge_A + (size_t)amg->ge_n * amg->ge_n, f, u);
      B200_LAUNCH_CHECK();
      return 0;
    }
    B200_TRY(jacobi(h, L, w, f, u, L.T));
    B200_TRY(b200_vec_copy(h, L.n, L.T, u));
    return 0;
  }
  // Every level keeps two buffers with fixed roles -- U: the pre-smoothed iterate, later the level's final iterate; T: the
  // iterate after the coarse-grid correction -- so that the kernel arguments of a cycle never change from one application to
  // the next (the PCG loop replays the whole iteration as a CUDA graph).
  for (int l = 0; l < nl - 1; l++) {
    b200_level &L = amg->lv[l];
    b200_level &Lc = amg->lv[l + 1];
    const bool zero = (l > 0) || u_zero;          // coarse iterates start at 0 (par_cycle.c:538)
    if (zero) {
      jacobi_zero_kernel<<<vgrid(h, L.n), 256, 0, h->stream>>>((size_t)L.n, w, F[l], L.l1, L.U);
      B200_LAUNCH_CHECK();
    } else {
      B200_TRY(jacobi(h, L, w, f, u, L.U));
    }
    U[l] = L.U;
    // Vtemp = F - A U (par_cycle.c:549) ; F_{l+1} = R Vtemp (:566)
    B200_TRY(b200_csr_spmv_epi(h, L.As, L.U, amg->Vtemp, 0, -1.0, 1.0, F[l], nullptr));
    B200_TRY(b200_csr_spmv_epi(h, L.R, amg->Vtemp, Lc.F, 0, 1.0, 0.0, nullptr, nullptr));
  }
  // coarsest level
  {
    b200_level &L = amg->lv[nl - 1];
    if (amg->coarse_ge) {
      gselim_kernel<<<1, 32, 0, h->stream>>>(amg->ge_n, amg->ge_A, amg->ge_A + (size_t)amg->ge_n * amg->ge_n, L.F, L.U);
      B200_LAUNCH_CHECK();
    } else {
      jacobi_zero_kernel<<<vgrid(h, L.n), 256, 0, h->stream>>>((size_t)L.n, w, L.F, L.l1, L.U);
      B200_LAUNCH_CHECK();
    }
    U[nl - 1] = L.U;
  }
  for (int l = nl - 2; l >= 0; l--) {
    b200_level &L = amg->lv[l];
    // T_l = U_l + P U_{l+1} (:602)
    B200_TRY(b200_csr_spmv_epi(h, L.P, U[l + 1], L.T, 0, 1.0, 1.0, L.U, nullptr));
    // post-smoothing sweep: the level's final iterate
    B200_TRY(jacobi(h, L, w, F[l], L.T, (l == 0) ? u : L.U));
  }
  return 0;
}

// the plain V(1,1) l1-Jacobi / Jacobi cycle above launches a fixed kernel sequence with fixed arguments
static bool amg_cycle_is_static(b200_amg amg) { return !(amg->general_cycle && amg->lv.size() > 1) && !amg->gs; }

__global__ void diag_scale_kernel(size_t n, const int *__restrict__ A_i, const double *__restrict__ A_a,
                                  const double *__restrict__ y, double *__restrict__ x) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = y[i] / A_a[A_i[i]];     // HYPRE_parcsr_pcg.c:239-255 (diagonal is stored first)
}

// HYPRE_ParCSRDiagScale (parcsr_ls/HYPRE_parcsr_pcg.c:228-258): x = y ./ diag(A)
extern "C" int b200_parcsr_diag_scale(b200_handle h, b200_parcsr A, const double *d_y, double *d_x) {
  if (!A) B200_FAIL("diag_scale: null matrix");
  const int n = A->diag->nrows;
  diag_scale_kernel<<<vgrid(h, n), 256, 0, h->stream>>>((size_t)n, A->diag->i, A->diag->a, d_y, d_x);
  B200_LAUNCH_CHECK();
  return 0;
}

// hypre_BoomerAMGSolve (par_amg_solve.c:21-380): V-cycles until ||f - A u|| / ||f|| < Tol (converge_type 0)
// or MaxIter cycles.  MaxIter 1 / Tol 0 is the preconditioner configuration (one cycle, no norms).
extern "C" int b200_amg_solve_ex(b200_handle h, b200_amg amg, b200_parcsr A, const double *d_f, double *d_u,
                                 int *num_iterations, double *final_rel_res) {
  if (!amg || !amg->is_setup) B200_FAIL("amg_solve: setup has not been called");
  const int max_iter = amg->ip["MaxIter"], min_iter = amg->ip["MinIter"];
  const double tol = amg->rp["Tol"];
  const int n = amg->lv[0].n;
  double resid_nrm = 1.0, rhs_norm = 0.0, relative_resid = 1.0, t = 0.0;
  b200_csr A0 = A ? A->diag : amg->lv[0].A;
  if (tol > 0.) {                                                                  // :143-214
    B200_TRY(b200_csr_spmv_epi(h, A0, d_u, amg->Vtemp, 0, -1.0, 1.0, d_f, nullptr));
    B200_TRY(b200_vec_dot(h, n, amg->Vtemp, amg->Vtemp, &t));
    resid_nrm = std::sqrt(t);
    if (resid_nrm != 0. && !(resid_nrm / resid_nrm == resid_nrm / resid_nrm))
      B200_FAIL("hypre_BoomerAMGSolve: INFs and/or NaNs detected in input");
    B200_TRY(b200_vec_dot(h, n, d_f, d_f, &t));
    rhs_norm = std::sqrt(t);
    relative_resid = rhs_norm ? resid_nrm / rhs_norm : resid_nrm;
  }
  int cycle_count = 0;
  while ((relative_resid >= tol || cycle_count < min_iter) && cycle_count < max_iter) {   // :236
    B200_TRY(amg_cycle(h, amg, d_f, d_u, false));
    if (tol > 0.) {
      B200_TRY(b200_csr_spmv_epi(h, A0, d_u, amg->Vtemp, 0, -1.0, 1.0, d_f, nullptr));
      B200_TRY(b200_vec_dot(h, n, amg->Vtemp, amg->Vtemp, &t));
      resid_nrm = std::sqrt(t);
      relative_resid = rhs_norm ? resid_nrm / rhs_norm : resid_nrm;
    }
    ++cycle_count;
  }
  if (num_iterations) *num_iterations = cycle_count;
  if (final_rel_res) *final_rel_res = relative_resid;
  if (cycle_count == max_iter && tol > 0.) return 256;       // HYPRE_ERROR_CONV (:307-311); the solution is still valid
  return 0;
}

// ClearVector(out); precond(A, rhs, out): one cycle from a zero guess, for the Krylov drivers in b200_krylov.cu
int b200_amg_precond(b200_handle h, b200_amg amg, const double *d_rhs, double *d_out) {
  if (!amg || !amg->is_setup) B200_FAIL("krylov: preconditioner has not been set up");
  return amg_cycle(h, amg, d_rhs, d_out, true);
}

extern "C" int b200_amg_solve(b200_handle h, b200_amg amg, const double *d_f, double *d_u) {
  if (!amg || !amg->is_setup) B200_FAIL("amg_solve: setup has not been called");
  if (amg->ip["MaxIter"] != 1 || amg->rp["Tol"] != 0.0)
    B200_FAIL("amg_solve: preconditioner entry point needs MaxIter 1 / Tol 0 (use b200_amg_solve_ex as a solver)");
  return amg_cycle(h, amg, d_f, d_u, false);
}

// hypre_PCGSolve (krylov/pcg.c:271-757).  Supported: two_norm 0/1, tol, a_tol, max_iter; preconditioner =
// BoomerAMG (amg != NULL), diagonal scaling (precond 2) or none.  rel_change / recompute_residual /
// stop_crit / cf_tol / rtol are rejected (ij.c passes 0 for all of them).
extern "C" int b200_pcg_solve_ex(b200_handle h, b200_parcsr A, b200_amg amg, const b200_pcg_params *prm, const double *d_b,
                                 double *d_x, int *iters_out, double *final_rel_res, double *h_norms) {
  if (!A || !prm) B200_FAIL("pcg: null argument");
  if (A->offd->ncols > 0) B200_FAIL("pcg: this entry point is single-rank; use b200_dist_pcg_solve");
  if (amg && !amg->is_setup) B200_FAIL("pcg: preconditioner has not been set up");
  if (prm->rel_change || prm->recompute_residual) B200_FAIL("pcg: rel_change / recompute_residual are not implemented");
  const double tol = prm->tol, a_tol = prm->a_tol;
  const int max_iter = prm->max_iter, two_norm = prm->two_norm;
  const int n = A->diag->nrows;
  double *p = nullptr, *s = nullptr, *r = nullptr, *sc = nullptr;
  B200_TRY(b200_dalloc<double>(h, &p, n));
  B200_TRY(b200_dalloc<double>(h, &s, n));
  B200_TRY(b200_dalloc<double>(h, &r, n));
  B200_TRY(b200_dalloc<double>(h, &sc, 8));
  double *hs = h->h_pinned;
  auto precond = [&](const double *rhs, double *out) -> int {
    if (amg) return amg_cycle(h, amg, rhs, out, true);    // ClearVector + precond (pcg.c:434-435,:568-569)
    if (prm->precond == 2) return b200_parcsr_diag_scale(h, A, rhs, out);
    return b200_vec_copy(h, n, rhs, out);                 // identity preconditioner (hypre_ParKrylovIdentity)
  };
  int rc = 0, i = 0;
  long long launches_per_graph = 0;
  double bi_prod = 0, i_prod = 0, eps = 0;
  do {
    if (two_norm) {
      if ((rc = b200_vec_dot(h, n, d_b, d_b, &bi_prod))) break;             // :347
    } else {
      if ((rc = precond(d_b, p))) break;                                    // :351-354  <C*b, b>
      if ((rc = b200_vec_dot(h, n, p, d_b, &bi_prod))) break;
    }
    if (bi_prod != 0. && !(bi_prod / bi_prod == bi_prod / bi_prod)) {       // :359-381
      rc = b200_set_error(__FILE__, __LINE__, "hypre_PCGSolve: INFs and/or NaNs detected in input"); break;
    }
    eps = tol * tol;                                                        // :383
    if (!(bi_prod > 0.0)) {                                                 // :403-416  b == 0 -> x = b
      if ((rc = b200_vec_copy(h, n, d_b, d_x))) break;
      if (h_norms) h_norms[0] = 0.0;
      break;
    }
    eps = std::fmax(tol * tol, a_tol * a_tol / bi_prod);                    // :393-400 default criterion
    // r = b - A x (:428-430)
    if ((rc = b200_parcsr_matvec(h, -1.0, A, d_x, 1.0, d_b, r))) break;
    if ((rc = precond(r, p))) break;                                        // p = C r
    if ((rc = b200_vec_dot_dev(h, n, r, p, sc + 0))) break;                 // gamma = <r,p> (:438)
    {                                                                       // :440-462: INF -> NaN conversion on gamma
      B200_CUDA(cudaMemcpyAsync(hs, sc, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      B200_CUDA(cudaStreamSynchronize(h->stream));
      const double gamma0 = hs[0];
      if (gamma0 != 0. && !(gamma0 / gamma0 == gamma0 / gamma0)) {
        rc = b200_set_error(__FILE__, __LINE__, "hypre_PCGSolve: INFs and/or NaNs detected in input"); break;
      }
    }
    if (h_norms) {
      double i_prod_0 = 0;
      if (two_norm) { if ((rc = b200_vec_dot(h, n, r, r, &i_prod_0))) break; }   // :466
      else { if ((rc = b200_vec_dot(h, n, r, p, &i_prod_0))) break; }
      h_norms[0] = std::sqrt(i_prod_0);
    }
    // One iteration = [beta, p update] of the previous one + [s = A p, <s,p>, alpha, x/r update, s = C r, <r,s>, <r,r>, copy of
    // the scalars to the host]: a fixed kernel sequence with fixed arguments.  Iteration 1 runs eagerly (everything lazily
    // built is built), iteration 2 is captured into a CUDA graph while it runs, iterations 3.. replay it: one launch and one
    // synchronisation per iteration instead of ~100 launches, a third of them on levels where the launch outlasts the kernel.
    const bool graph_ok = b200_graph_enabled() && (!amg || amg_cycle_is_static(amg));
    cudaGraphExec_t gexec = nullptr;
    bool graph_tried = false;
    auto body = [&](bool with_beta) -> int {
      if (with_beta) {
        pcg_beta_kernel<<<1, 1, 0, h->stream>>>(sc);
        ++g_b200_launches;
        pcg_update_p_kernel<<<vgrid(h, n), 256, 0, h->stream>>>((size_t)n, sc, s, p);
        ++g_b200_launches;
      }
      B200_TRY(b200_parcsr_matvec(h, 1.0, A, p, 0.0, nullptr, s));          // s = A p (:512)
      B200_TRY(b200_vec_dot_dev(h, n, s, p, sc + 1));                       // sdotp (:515)
      pcg_alpha_kernel<<<1, 1, 0, h->stream>>>(sc);
      ++g_b200_launches;
      pcg_update_xr_kernel<<<vgrid(h, n), 256, 0, h->stream>>>((size_t)n, sc, p, s, d_x, r);
      ++g_b200_launches;
      B200_TRY(precond(r, s));                                              // s = C r (:568-569)
      if (two_norm) B200_TRY(b200_vec_dot2_dev(h, n, r, s, sc + 0, sc + 3)); // gamma = <r,s> (:572), i_prod = <r,r> (:590): one pass over r
      else B200_TRY(b200_vec_dot_dev(h, n, r, s, sc + 0));
      B200_CUDA(cudaMemcpyAsync(hs, sc, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      return 0;
    };
    while ((i + 1) <= max_iter) {                                           // :498
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
        if ((rc = body(i > 1))) break;
      }
      if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = b200_set_error(__FILE__, __LINE__, "pcg sync failed"); break; }
      const double gamma = hs[0], sdotp = hs[1];
      i_prod = two_norm ? hs[3] : gamma;                                    // :589-592
      if (sdotp == 0.0) { rc = b200_set_error(__FILE__, __LINE__, "Zero sdotp value in PCG"); break; }   // :516-521
      if (h_norms) h_norms[i] = std::sqrt(i_prod);
      if (i_prod / bi_prod < eps) break;                                    // converged (:634, :672-676)
      if (!(gamma > 2.2250738585072014e-308)) { rc = b200_set_error(__FILE__, __LINE__, "Subnormal gamma value in PCG"); break; }
    }
    b200_graph_destroy(gexec);
  } while (0);
  if (!rc) {
    if (iters_out) *iters_out = i;
    if (final_rel_res) *final_rel_res = bi_prod > 0.0 ? std::sqrt(i_prod / bi_prod) : 0.0;   // :751-754
  }
  b200_dfree(h, p); b200_dfree(h, s); b200_dfree(h, r); b200_dfree(h, sc);
  return rc;
}

extern "C" int b200_pcg_solve(b200_handle h, b200_parcsr A, b200_amg amg, const double *d_b, double *d_x, double tol,
                              int max_iter, int *iters_out, double *final_rel_res, double *h_norms) {
  b200_pcg_params prm;
  prm.tol = tol; prm.a_tol = 0.0; prm.max_iter = max_iter; prm.two_norm = 1;            // ij.c:3892-3897
  prm.rel_change = 0; prm.recompute_residual = 0; prm.precond = amg ? 1 : 0;
  return b200_pcg_solve_ex(h, A, amg, &prm, d_b, d_x, iters_out, final_rel_res, h_norms);
}
