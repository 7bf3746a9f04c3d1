// b200_hmis.cu -- HMIS coarsening (coarsen_type 10, the library and driver default).
//
// Reference: hypre_BoomerAMGCoarsenHMIS (parcsr_ls/par_coarsen.c:2774-2797) = the Ruge-Stueben first pass
// (hypre_BoomerAMGCoarsenRuge with coarsen_type 10 -> 11, f_pnt = Z_PT, :1046-1330) followed by PMIS seeded with the
// first pass's C points (hypre_BoomerAMGCoarsenPMISHost with CF_init 1, :2279-2309, :2420).
//
// The first pass is sequential BY DEFINITION: every step takes the oldest point of the largest measure from the
// reference's list of lists (utilities/amg_linklist.c) and the measures of its second-ring neighbours change before the
// next step.  There is no parallel formulation with the same result, so the device runs it the way the reference runs it
// on one rank -- one thread, the lists kept as one FIFO per measure value in global memory -- and only the PMIS part is
// parallel.  It exists so that the LITERAL default configuration (`ij -solver 1` without -pmis, BASELINE.json configs[0])
// builds its hierarchy on the device bit for bit; it costs a dependent L2 access per list operation (seconds at 10^5
// rows), and `-pmis` stays the configuration to use at scale.  Every loop below has a static bound: a corrupted list
// ends the kernel with a status code instead of spinning.
#include "b200_internal.h"
#include "b200_hmis_body.h"

int b200_pmis_rows_init(b200_handle h, b200_csr S, int seed, long long first_row, int cf_init, int *d_cf, int *iterations);

namespace {
__global__ void colcount_max_kernel(int n, const int *__restrict__ T_i, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMax(out, T_i[i + 1] - T_i[i]);
}
__global__ void fill_int_kernel(int n, int v, int *__restrict__ x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

__global__ void ruge_first_pass_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ S_j, const int *__restrict__ T_i,
                                       const int *__restrict__ T_j, int agg2, int nb, int *cf, int *meas, int *next, int *prev,
                                       int *head, int *tail, int *status) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  *status = b200_ruge_first_pass_body(n, S_i, S_j, T_i, T_j, agg2, nb, cf, meas, next, prev, head, tail);
}
}  // namespace

// Ruge-Stueben first pass on the strength pattern S (no diagonal, par_strength.c output); d_cf receives
// C_PT 1 / F_PT -1 / Z_PT -2 / SF_PT -3.  agg2 = the second coarsening of an aggressive level (measure_type + 3).
int b200_ruge_first_pass(b200_handle h, b200_csr S, int agg2, int *d_cf) {
  if (!S || !d_cf) B200_FAIL("ruge: null argument");
  const int n = S->nrows;
  if (n == 0) return 0;
  if (S->ncols != n) B200_FAIL("ruge: the strength pattern must be square (the first pass walks the rows of S^T as well)");
  if (n > 1000000)       // about a minute of one thread's dependent L2 accesses; beyond that the run looks like a hang
    B200_FAIL("HMIS: the Ruge-Stueben first pass is sequential; above 1 000 000 rows per level use CoarsenType 8 (PMIS, ij -pmis)");
  b200_csr T = nullptr;
  B200_TRY(b200_csr_transpose(h, S, &T));                         // S^T with rows ordered by source row (:1014-1043)
  int *d_max = nullptr, hmax = 0;
  B200_TRY(b200_dalloc<int>(h, &d_max, 2));
  B200_CUDA(cudaMemsetAsync(d_max, 0, 2 * sizeof(int), h->stream));
  colcount_max_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, T->i, d_max);
  B200_LAUNCH_CHECK();
  B200_CUDA(cudaMemcpyAsync(&hmax, d_max, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  const int nb = 2 * hmax + 4;        // a measure starts at its column count and gains at most one point per dependent point
  int *meas = nullptr, *next = nullptr, *prev = nullptr, *head = nullptr, *tail = nullptr;
  B200_TRY(b200_dalloc<int>(h, &meas, n)); B200_TRY(b200_dalloc<int>(h, &next, n)); B200_TRY(b200_dalloc<int>(h, &prev, n));
  B200_TRY(b200_dalloc<int>(h, &head, nb)); B200_TRY(b200_dalloc<int>(h, &tail, nb));
  fill_int_kernel<<<b200_grid(nb, 256), 256, 0, h->stream>>>(nb, -1, head);
  B200_LAUNCH_CHECK();
  fill_int_kernel<<<b200_grid(nb, 256), 256, 0, h->stream>>>(nb, -1, tail);
  B200_LAUNCH_CHECK();
  ruge_first_pass_kernel<<<1, 32, 0, h->stream>>>(n, S->i, S->j, T->i, T->j, agg2 ? 1 : 0, nb, d_cf, meas, next, prev, head, tail,
                                                  d_max + 1);
  B200_LAUNCH_CHECK();
  int status = 0;
  B200_CUDA(cudaMemcpyAsync(&status, d_max + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, meas)); B200_TRY(b200_dfree(h, next)); B200_TRY(b200_dfree(h, prev));
  B200_TRY(b200_dfree(h, head)); B200_TRY(b200_dfree(h, tail)); B200_TRY(b200_dfree(h, d_max));
  B200_TRY(b200_csr_destroy(h, T));
  if (status) B200_FAIL("ruge: the measure lists became inconsistent (internal error)");
  return 0;
}

// hypre_BoomerAMGCoarsenHMIS on one rank: d_cf receives C_PT 1 / F_PT -1 / SF_PT -3
extern "C" int b200_hmis(b200_handle h, b200_csr S, int seed, int *d_cf) {
  if (!S || !d_cf) B200_FAIL("hmis: null argument");
  B200_TRY(b200_ruge_first_pass(h, S, 0, d_cf));
  return b200_pmis_rows_init(h, S, seed, 0, 1, d_cf, nullptr);
}
