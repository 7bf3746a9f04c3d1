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

struct Lists {                  // one FIFO per measure value: the order hypre_enter_on_lists / hypre_remove_point keep
  int *head, *tail, *next, *prev;
  int nb, maxm, bad;
  __device__ void enter(int m, int i) {
    if (m < 0 || m >= nb) { bad = 1; return; }
    next[i] = -1;
    prev[i] = tail[m];
    if (tail[m] >= 0) next[tail[m]] = i; else head[m] = i;
    tail[m] = i;
    if (m > maxm) maxm = m;
  }
  __device__ void remove(int m, int i) {
    if (m < 0 || m >= nb) { bad = 1; return; }
    const int p = prev[i], q = next[i];
    if (p >= 0) next[p] = q; else head[m] = q;
    if (q >= 0) prev[q] = p; else tail[m] = p;
    for (int guard = 0; guard < nb && maxm > 0 && head[maxm] < 0; guard++) maxm--;
  }
};

// markers as in par_coarsen.c:860-865: C_PT 1, F_PT -1, Z_PT -2, SF_PT -3, SC_PT 3, UNDECIDED 0
__global__ void ruge_first_pass_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ S_j, const int *__restrict__ T_i,
                                       const int *__restrict__ T_j, int agg2, int nb, int *cf, int *meas, int *next, int *prev,
                                       int *head, int *tail, int *status) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Lists L{head, tail, next, prev, nb, 0, 0};
  int num_left = 0;
  for (int j = 0; j < n; j++) {                                   // :1130-1158, measures = row sums of S^T (:1056-1059)
    meas[j] = T_i[j + 1] - T_i[j];
    if (S_i[j + 1] - S_i[j] == 0) { cf[j] = agg2 ? 3 : -3; meas[j] = 0; }
    else { cf[j] = 0; num_left++; }
  }
  for (int j = 0; j < n; j++) {                                   // :1179-1222
    const int measure = meas[j];
    if (cf[j] == -3 || cf[j] == 3) continue;
    if (measure > 0) { L.enter(measure, j); continue; }
    cf[j] = -2;                                                   // nothing depends on j: f_pnt = Z_PT
    for (int k = S_i[j]; k < S_i[j + 1]; k++) {
      const int nabor = S_j[k];
      if (cf[nabor] == -3 || cf[nabor] == 3) continue;
      if (nabor < j) {
        int nm = meas[nabor];
        if (nm > 0) L.remove(nm, nabor);
        nm = ++meas[nabor];
        L.enter(nm, nabor);
      } else {
        ++meas[nabor];
      }
    }
    --num_left;
  }
  for (int step = 0; step < n && num_left > 0; step++) {          // :1245-1320, at most one C point per step
    const int index = head[L.maxm];
    if (index < 0 || index >= n) { L.bad = 2; break; }
    const int measure = meas[index];
    cf[index] = 1;
    meas[index] = 0;
    --num_left;
    L.remove(measure, index);
    for (int j = T_i[index]; j < T_i[index + 1]; j++) {           // the points that depend on the new C point become F
      const int nabor = T_j[j];
      if (cf[nabor] != 0) continue;
      cf[nabor] = -1;
      L.remove(meas[nabor], nabor);
      --num_left;
      for (int k = S_i[nabor]; k < S_i[nabor + 1]; k++) {         // ... and what they depend on gains a measure point
        const int n2 = S_j[k];
        if (cf[n2] != 0) continue;
        L.remove(meas[n2], n2);
        ++meas[n2];
        L.enter(meas[n2], n2);
      }
    }
    for (int j = S_i[index]; j < S_i[index + 1]; j++) {           // what the C point depends on loses a measure point
      const int nabor = S_j[j];
      if (cf[nabor] != 0) continue;
      int m2 = meas[nabor];
      L.remove(m2, nabor);
      meas[nabor] = --m2;
      if (m2 > 0) { L.enter(m2, nabor); continue; }
      cf[nabor] = -1;
      --num_left;
      for (int k = S_i[nabor]; k < S_i[nabor + 1]; k++) {
        const int n2 = S_j[k];
        if (cf[n2] != 0) continue;
        L.remove(meas[n2], n2);
        ++meas[n2];
        L.enter(meas[n2], n2);
      }
    }
  }
  for (int i = 0; i < n; i++)
    if (cf[i] == 3) cf[i] = 1;                                    // :1337-1343 SC_PT -> C_PT
  *status = L.bad ? L.bad : (num_left > 0 ? 3 : 0);
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
