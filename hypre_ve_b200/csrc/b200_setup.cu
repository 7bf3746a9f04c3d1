// b200_setup.cu -- BoomerAMG setup kernels: strength, PMIS, ext+i interpolation, truncation,
// transpose, Gustavson SpGEMM, l1 norms.   COMPILED WITH -fmad=false (see build.py).
//
// Parity contract (BASELINE.json north_star): CF splitting, interpolation sparsity and
// coarse-grid structure are bit-exact against the reference CPU build, coarse values to 1e-12.
// We go further and keep every floating-point operation in the reference's order, one strictly
// ordered lane per row where the order matters, so the hierarchy is bit-identical:
//   * no FMA contraction (the reference is gcc -O2 x86-64: separate mul and add);
//   * interpolation weights are accumulated in the reference's loop order
//     (par_lr_interp.c:1661-1800) and truncated by a replay of its quicksort
//     (utilities/hypre_qsort.c:367-387);
//   * SpGEMM emits columns in first-touch order and sums products left to right
//     (seq_mv/csr_matop.c:438-468).
// Parallelism comes from rows: one thread per row with a private open-addressing hash table in
// HBM scratch replacing the reference's dense per-thread marker arrays; rows are processed in
// chunks so the scratch stays bounded.
#include "b200_internal.h"
#include "b200_comm.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

constexpr int TB = 128;                         // threads per CTA for row kernels
constexpr long long HASH_BUDGET = 1LL << 28;    // hash slots per chunk (2 x 1 GiB of int scratch)

// ------------------------------------------------------------------------------------------
// per-row open addressing table (keys = column ids, -1 empty)
// ------------------------------------------------------------------------------------------
constexpr int NOTFOUND = -1;
constexpr int STRONG_F = -2;

struct Tab {
  int *k, *v;
  unsigned mask;
};
__device__ __forceinline__ unsigned hmix(int key) { return (unsigned)key * 0x9E3779B1u; }
__device__ __forceinline__ int tab_find(const Tab &t, int key) {
  unsigned s = (hmix(key) >> 7) & t.mask;
  while (true) {
    int kk = t.k[s];
    if (kk == key) return t.v[s];
    if (kk == -1) return NOTFOUND;
    s = (s + 1) & t.mask;
  }
}
// insert (key,val) if absent; returns true when newly inserted
__device__ __forceinline__ bool tab_insert(const Tab &t, int key, int val) {
  unsigned s = (hmix(key) >> 7) & t.mask;
  while (true) {
    int kk = t.k[s];
    if (kk == key) return false;
    if (kk == -1) { t.k[s] = key; t.v[s] = val; return true; }
    s = (s + 1) & t.mask;
  }
}
__device__ __forceinline__ int cap_for(int ub) {   // power of two >= 2*ub, 0 when nothing to store
  if (ub <= 0) return 0;
  int c = 4;
  while (c < 2 * ub) c <<= 1;
  return c;
}

// ------------------------------------------------------------------------------------------
// chunk planning: rows [bounds[c], bounds[c+1]) need at most ~budget hash slots
// ------------------------------------------------------------------------------------------
__global__ void widen_kernel(int n, const int *__restrict__ in, long long *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
  if (i == n) out[n] = 0;
}
__global__ void bounds_kernel(int n, const long long *__restrict__ scan, long long budget, int nchunks, int *bounds) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nchunks) return;
  if (c == nchunks) { bounds[c] = n; return; }
  long long target = (long long)c * budget;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (scan[mid] >= target) hi = mid; else lo = mid + 1;
  }
  bounds[c] = lo;
}

struct ChunkPlan {
  long long *scan = nullptr;        // device, exclusive scan of capacities [n+1]
  std::vector<int> bounds;          // host, chunk row boundaries
  std::vector<long long> base;      // host, scan[bounds[c]]
  long long max_slots = 0;          // largest chunk
};

int plan_chunks(b200_handle h, int n, const int *d_cap, ChunkPlan *plan) {
  B200_TRY(b200_dalloc<long long>(h, &plan->scan, (size_t)n + 1));
  widen_kernel<<<b200_grid((size_t)n + 1, 256), 256, 0, h->stream>>>(n, d_cap, plan->scan);
  B200_LAUNCH_CHECK();
  size_t tb = 0;
  B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, plan->scan, plan->scan, n + 1, h->stream));
  char *tmp = nullptr;
  B200_TRY(b200_dalloc<char>(h, &tmp, tb));
  B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, plan->scan, plan->scan, n + 1, h->stream));
  ++g_b200_launches;
  B200_TRY(b200_dfree(h, tmp));
  long long total = 0;
  B200_CUDA(cudaMemcpyAsync(&total, plan->scan + n, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  int nchunks = (int)((total + HASH_BUDGET - 1) / HASH_BUDGET);
  if (nchunks < 1) nchunks = 1;
  int *d_bounds = nullptr;
  B200_TRY(b200_dalloc<int>(h, &d_bounds, (size_t)nchunks + 1));
  bounds_kernel<<<b200_grid((size_t)nchunks + 1, 64), 64, 0, h->stream>>>(n, plan->scan, HASH_BUDGET, nchunks, d_bounds);
  B200_LAUNCH_CHECK();
  plan->bounds.resize(nchunks + 1);
  B200_CUDA(cudaMemcpyAsync(plan->bounds.data(), d_bounds, sizeof(int) * ((size_t)nchunks + 1), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, d_bounds));
  plan->base.resize(nchunks + 1);
  for (int c = 0; c <= nchunks; c++) {
    B200_CUDA(cudaMemcpyAsync(&plan->base[c], plan->scan + plan->bounds[c], sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  }
  B200_CUDA(cudaStreamSynchronize(h->stream));
  plan->max_slots = 1;
  for (int c = 0; c < nchunks; c++) {
    long long s = plan->base[c + 1] - plan->base[c];
    if (s > plan->max_slots) plan->max_slots = s;
  }
  return 0;
}

// ==========================================================================================
// Strength of connection  (par_strength.c:231-504, num_functions == 1)
// ==========================================================================================
template <bool FILL>
__global__ void strength_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                const double *__restrict__ A_a, double theta, double max_row_sum,
                                int *__restrict__ S_i, int *__restrict__ S_j) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = A_i[i], e = A_i[i + 1];
  if (b == e) { if (!FILL) S_i[i] = 0; return; }
  const double diag = A_a[b];
  double row_scale = 0.0, row_sum = diag;
  if (diag < 0) {
    for (int jA = b + 1; jA < e; jA++) { double v = A_a[jA]; row_scale = row_scale < v ? v : row_scale; row_sum += v; }
  } else {
    for (int jA = b + 1; jA < e; jA++) { double v = A_a[jA]; row_scale = row_scale < v ? row_scale : v; row_sum += v; }
  }
  int cnt = 0;
  int pos = FILL ? S_i[i] : 0;
  if ((fabs(row_sum) > fabs(diag) * max_row_sum) && (max_row_sum < 1.0)) {
    /* all dependencies weak */
  } else {
    const double thr = theta * row_scale;
    if (diag < 0) {
      for (int jA = b + 1; jA < e; jA++)
        if (!(A_a[jA] <= thr)) { if (FILL) S_j[pos + cnt] = A_j[jA]; cnt++; }
    } else {
      for (int jA = b + 1; jA < e; jA++)
        if (!(A_a[jA] >= thr)) { if (FILL) S_j[pos + cnt] = A_j[jA]; cnt++; }
    }
  }
  if (!FILL) S_i[i] = cnt;
}

// ==========================================================================================
// PMIS (par_coarsen.c:2031-2738) with the sequential Park-Miller stream of hypre_Rand
// (utilities/random.c:49-106) evaluated by jump-ahead: Seed_k = seed * 16807^k mod (2^31-1).
// ==========================================================================================
__device__ __forceinline__ unsigned long long mulmod31(unsigned long long a, unsigned long long b) {
  return (a * b) % 2147483647ULL;
}
__device__ __forceinline__ int lcg_at(int seed, unsigned long long k) {   // state after k calls of hypre_RandI
  unsigned long long base = 16807ULL, acc = (unsigned long long)seed;
  while (k) {
    if (k & 1ULL) acc = mulmod31(acc, base);
    base = mulmod31(base, base);
    k >>= 1;
  }
  return (int)acc;
}
__global__ void colcount_kernel(int nnz, const int *__restrict__ S_j, int *__restrict__ cnt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) atomicAdd(&cnt[S_j[k]], 1);
}
__global__ void pmis_init_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ colcnt, int seed,
                                 long long first_row, double *__restrict__ measure, int *__restrict__ cf, int cf_init = 0) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // measure = #influences + hypre_Rand()   (par_indepset.c:56-59: i-th local row takes the (i+1)-th draw)
  int s = lcg_at(seed, (unsigned long long)(first_row + i + 1));
  double m = (double)colcnt[i] + ((double)s / 2147483647.0);
  if (cf_init == 1) {               // HMIS: cf holds the Ruge-Stueben first pass (par_coarsen.c:2279-2309)
    int c = cf[i];
    if (c == -3) {
      m = 0.0;                      // special F points stay out of the graph
    } else {
      if (c == -1) c = 0;           // every F point is undecided again
      if (c == -2) c = (m >= 1.0 || S_i[i + 1] - S_i[i] > 0) ? 0 : -1;   // Z points stay F only if nothing connects them
    }                               // C points (1) stay: they are the first independent set
    cf[i] = c;
    measure[i] = m;
    return;
  }
  if (S_i[i + 1] - S_i[i] == 0) {   // isolated point: SF_PT, measure 0; C_PT when CF_init is 3 (par_coarsen.c:2316-2328)
    cf[i] = (cf_init == 3) ? 1 : -3;
    m = 0.0;
  } else {
    cf[i] = 0;
  }
  measure[i] = m;
}
__global__ void pmis_mark_kernel(int n, const double *__restrict__ measure, int *__restrict__ cf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (cf[i] == 0 && measure[i] > 1) cf[i] = 1;       // :2430-2437 (graph nodes are exactly cf==0)
}
// graph flags: a node is in the graph this sweep iff ingraph[i] != 0
__global__ void pmis_remove_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                   const double *__restrict__ measure, const int *__restrict__ ingraph,
                                   int *__restrict__ cf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !ingraph[i]) return;
  const double mi = measure[i];
  if (mi > 1) {                                     // :2457-2478
    for (int jS = S_i[i]; jS < S_i[i + 1]; jS++) {
      int j = S_j[jS];
      double mj = measure[j];
      if (mj > 1) {
        if (mi > mj) cf[j] = 0;
        else if (mj > mi) cf[i] = 0;
      }
    }
  }
}
__global__ void pmis_setcf_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                  const int *__restrict__ ingraph, double *__restrict__ measure,
                                  const int *__restrict__ cf_in, int *__restrict__ cf_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cf_in[i];
  if (ingraph[i]) {                                 // :2543-2595
    if (measure[i] < 1) c = -1;
    if (c > 0) {
      c = 1;
    } else {
      for (int jS = S_i[i]; jS < S_i[i + 1]; jS++)
        if (cf_in[S_j[jS]] > 0) c = -1;
    }
    if (c != 0) measure[i] = 0;                     // :2643-2647
  }
  cf_out[i] = c;
}
// The same sweep with the NEXT sweep's graph flags and node count produced on the way (a node stays in the graph iff its new
// marker is 0): saves the separate pass that rebuilt them.  ingraph[i] is read and written by thread i only.
__global__ void pmis_setcf_graph_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                        int *__restrict__ ingraph, double *__restrict__ measure,
                                        const int *__restrict__ cf_in, int *__restrict__ cf_out, int *__restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int g = 0;
  if (i < n) {
    int c = cf_in[i];
    if (ingraph[i]) {                                 // :2543-2595
      if (measure[i] < 1) c = -1;
      if (c > 0) {
        c = 1;
      } else {
        for (int jS = S_i[i]; jS < S_i[i + 1]; jS++)
          if (cf_in[S_j[jS]] > 0) c = -1;
      }
      if (c != 0) measure[i] = 0;                     // :2643-2647
    }
    cf_out[i] = c;
    g = (c == 0) ? 1 : 0;
    ingraph[i] = g;
  }
  unsigned b = __ballot_sync(0xffffffffu, g);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}
// First sweep of PMIS with CF_init 1 (HMIS): the graph holds the undecided nodes AND the first pass's C points; no
// independent set is picked (`if (!CF_init || iter)`, :2420).  The reference updates CF_marker in place in index order
// (:2543-2595), and here that order is visible: a C point whose measure is below 1 is turned into an F point when its
// turn comes, so it still counts as a C neighbour for the nodes before it and no longer for the nodes after it.
__global__ void pmis_setcf_first_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                        const double *__restrict__ measure, const int *__restrict__ cf_in,
                                        int *__restrict__ cf_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cf_in[i];
  if (c == 0 || c == 1) {
    if (measure[i] < 1) c = -1;
    if (c > 0) {
      c = 1;
    } else {
      for (int jS = S_i[i]; jS < S_i[i + 1]; jS++) {
        const int j = S_j[jS];
        if (cf_in[j] > 0 && !(measure[j] < 1 && j < i)) c = -1;
      }
    }
  }
  cf_out[i] = c;
}
// nodes that left the graph in that sweep lose their measure (:2643-2647)
__global__ void pmis_clear_first_kernel(int n, const int *__restrict__ cf_before, const int *__restrict__ cf_after,
                                        double *__restrict__ measure) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (cf_before[i] == 0 || cf_before[i] == 1) && cf_after[i] != 0) measure[i] = 0;
}
__global__ void pmis_graph_kernel(int n, const int *__restrict__ cf, int *__restrict__ ingraph, int *__restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int g = (i < n && cf[i] == 0) ? 1 : 0;
  if (i < n) ingraph[i] = g;
  unsigned b = __ballot_sync(0xffffffffu, g);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// ==========================================================================================
// coarse numbering (par_coarse_parms.c:56-133 + fine_to_coarse of par_lr_interp.c:1307-1313)
// ==========================================================================================
__global__ void cflag_kernel(int n, const int *__restrict__ cf, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = cf[i] >= 0 ? 1 : 0;
  if (i == n) flag[n] = 0;
}

// ==========================================================================================
// extended+i interpolation (par_lr_interp.c:1040-1925, num_procs == 1, num_functions == 1)
// ==========================================================================================
__global__ void extpi_ub_kernel(int n, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                const int *__restrict__ cf, int *__restrict__ cap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { cap[n] = 0; return; }
  int ub = 0;
  if (cf[i] < 0 && cf[i] != -3) {
    for (int jj = S_i[i]; jj < S_i[i + 1]; jj++) {
      int i1 = S_j[jj];
      ub += 1;
      if (cf[i1] < 0 && cf[i1] != -3) ub += S_i[i1 + 1] - S_i[i1];
    }
  }
  cap[i] = cap_for(ub);
}

// pass 1: discover C-hat_i in the reference's order and count it (:1301-1416)
__global__ void extpi_discover_kernel(int r0, int r1, const int *__restrict__ S_i, const int *__restrict__ S_j,
                                      const int *__restrict__ cf, const long long *__restrict__ scan, long long base,
                                      int *__restrict__ keys, int *__restrict__ vals, int *__restrict__ cnt_full) {
  int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  int c = cf[i];
  if (c >= 0) { cnt_full[i - r0] = 1; return; }
  if (c == -3) { cnt_full[i - r0] = 0; return; }
  long long off = scan[i] - base;
  int cap = (int)(scan[i + 1] - scan[i]);
  if (cap == 0) { cnt_full[i - r0] = 0; return; }
  Tab t{keys + off, vals + off, (unsigned)(cap - 1)};
  int cnt = 0;
  for (int jj = S_i[i]; jj < S_i[i + 1]; jj++) {
    int i1 = S_j[jj];
    int c1 = cf[i1];
    if (c1 >= 0) {
      if (tab_insert(t, i1, cnt)) cnt++;
    } else if (c1 != -3) {
      tab_insert(t, i1, STRONG_F);
      for (int kk = S_i[i1]; kk < S_i[i1 + 1]; kk++) {
        int k1 = S_j[kk];
        if (cf[k1] >= 0) {
          if (tab_insert(t, k1, cnt)) cnt++;
        }
      }
    }
  }
  cnt_full[i - r0] = cnt;
}

// pass 2: weights, in the reference's accumulation order (:1523-1803)
__global__ void extpi_weights_kernel(int r0, int r1, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                     const double *__restrict__ A_a, const int *__restrict__ cf,
                                     const int *__restrict__ f2c, const long long *__restrict__ scan, long long base,
                                     const int *__restrict__ keys, const int *__restrict__ vals,
                                     const int *__restrict__ Pf_i, int *__restrict__ Pf_j, double *__restrict__ Pf_a) {
  int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  const int start = Pf_i[i - r0], end = Pf_i[i - r0 + 1];
  int c = cf[i];
  if (c >= 0) { Pf_j[start] = f2c[i]; Pf_a[start] = 1.0; return; }
  if (c == -3 || start == end) {
    // no interpolatory set: the reference leaves an empty row (diagonal division has nothing to scale)
    return;
  }
  long long off = scan[i] - base;
  int cap = (int)(scan[i + 1] - scan[i]);
  Tab t{const_cast<int *>(keys) + off, const_cast<int *>(vals) + off, (unsigned)(cap - 1)};
  for (int s = 0; s < cap; s++) {
    int key = t.k[s];
    if (key != -1) {
      int v = t.v[s];
      if (v >= 0) { Pf_j[start + v] = f2c[key]; Pf_a[start + v] = 0.0; }
    }
  }
  double diagonal = A_a[A_i[i]];
  for (int jj = A_i[i] + 1; jj < A_i[i + 1]; jj++) {
    const int i1 = A_j[jj];
    const int m1 = tab_find(t, i1);
    if (m1 >= 0) {
      Pf_a[start + m1] += A_a[jj];
    } else if (m1 == STRONG_F) {
      double sum = 0.0;
      int sgn = 1;
      if (A_a[A_i[i1]] < 0) sgn = -1;
      for (int jj1 = A_i[i1] + 1; jj1 < A_i[i1 + 1]; jj1++) {
        const int i2 = A_j[jj1];
        const double a = A_a[jj1];
        if ((sgn * a) < 0 && (i2 == i || tab_find(t, i2) >= 0)) sum += a;
      }
      if (sum != 0) {
        const double distribute = A_a[jj] / sum;
        for (int jj1 = A_i[i1] + 1; jj1 < A_i[i1 + 1]; jj1++) {
          const int i2 = A_j[jj1];
          const double a = A_a[jj1];
          if ((sgn * a) < 0) {
            const int m2 = tab_find(t, i2);
            if (m2 >= 0) Pf_a[start + m2] += distribute * a;
            if (i2 == i) diagonal += distribute * a;
          }
        }
      } else {
        diagonal += A_a[jj];
      }
    } else if (cf[i1] != -3) {
      diagonal += A_a[jj];
    }
  }
  if (diagonal) {
    for (int jj = start; jj < end; jj++) Pf_a[jj] /= -diagonal;
  }
}

// replay of hypre_qsort2_abs (utilities/hypre_qsort.c:367-387): middle pivot, strict '>' on |w|.
// Sub-ranges are independent, so an explicit stack (larger range pushed, smaller iterated) gives
// the same final order as the recursion with bounded depth.
__device__ void qsort2_abs(int *v, double *w, int left, int right) {
  int stack_l[40], stack_r[40];
  int sp = 0;
  while (true) {
    while (left < right) {
      int mid = (left + right) / 2;
      int tv = v[left]; v[left] = v[mid]; v[mid] = tv;
      double tw = w[left]; w[left] = w[mid]; w[mid] = tw;
      int last = left;
      const double piv = fabs(w[left]);
      for (int i = left + 1; i <= right; i++) {
        if (fabs(w[i]) > piv) {
          ++last;
          tv = v[last]; v[last] = v[i]; v[i] = tv;
          tw = w[last]; w[last] = w[i]; w[i] = tw;
        }
      }
      tv = v[left]; v[left] = v[last]; v[last] = tv;
      tw = w[left]; w[left] = w[last]; w[last] = tw;
      // ranges (left,last-1) and (last+1,right)
      int l1 = left, r1 = last - 1, l2 = last + 1, r2 = right;
      if (r1 - l1 > r2 - l2) {           // push the larger, continue with the smaller
        if (l1 < r1) { stack_l[sp] = l1; stack_r[sp] = r1; sp++; }
        left = l2; right = r2;
      } else {
        if (l2 < r2) { stack_l[sp] = l2; stack_r[sp] = r2; sp++; }
        left = l1; right = r1;
      }
    }
    if (sp == 0) break;
    --sp;
    left = stack_l[sp]; right = stack_r[sp];
  }
}

// hypre_ParCSRMatrixTruncate (par_csr_matrix.c:2671-3060) on one row segment, in place.
// Writes the kept length; kept entries occupy the head of the segment.
__global__ void truncate_kernel(int nrows, const int *__restrict__ Pf_i, int *__restrict__ Pf_j,
                                double *__restrict__ Pf_a, double tol, int max_elmts, int *__restrict__ P_cnt) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int start = Pf_i[r];
  int len = Pf_i[r + 1] - start;
  int *v = Pf_j + start;
  double *w = Pf_a + start;
  if (tol > 0) {                                     // :2768-2890, nrm_type 0, rescale 1
    double row_nrm = 0;
    for (int j = 0; j < len; j++) row_nrm = (row_nrm < fabs(w[j])) ? fabs(w[j]) : row_nrm;
    const double drop = tol * row_nrm;
    double row_sum = 0, scale = 0;
    int keep = 0;
    for (int j = 0; j < len; j++) {
      row_sum += w[j];
      if (!(fabs(w[j]) < drop)) { scale += w[j]; w[keep] = w[j]; v[keep] = v[j]; keep++; }
    }
    len = keep;
    if (scale != 0.) {
      if (scale != row_sum) {
        scale = row_sum / scale;
        for (int j = 0; j < len; j++) w[j] *= scale;
      }
    }
  }
  if (max_elmts > 0 && len > max_elmts) {            // :2906-3020
    double row_sum = 0;
    for (int j = 0; j < len; j++) row_sum += w[j];
    qsort2_abs(v, w, 0, len - 1);
    double scale = 0;
    for (int j = 0; j < max_elmts; j++) scale += w[j];
    len = max_elmts;
    if (scale != 0.) {
      if (scale != row_sum) {
        scale = row_sum / scale;
        for (int j = 0; j < len; j++) w[j] *= scale;
      }
    }
  }
  P_cnt[r] = len;
}

__global__ void compact_rows_kernel(int nrows, const int *__restrict__ src_i, const int *__restrict__ src_j,
                                    const double *__restrict__ src_a, const int *__restrict__ dst_i,
                                    int *__restrict__ dst_j, double *__restrict__ dst_a) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  int s = src_i[r], d = dst_i[r], len = dst_i[r + 1] - d;
  for (int k = 0; k < len; k++) { dst_j[d + k] = src_j[s + k]; if (dst_a) dst_a[d + k] = src_a[s + k]; }
}

// ==========================================================================================
// transpose (csr_matop.c:578-779): stable counting sort by column == stable radix sort by column key
// ==========================================================================================
__global__ void expand_rows_kernel(int nrows, const int *__restrict__ A_i, int *__restrict__ rows) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  for (int k = A_i[r]; k < A_i[r + 1]; k++) rows[k] = r;
}
__global__ void iota_kernel(int n, int *x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = i;
}
__global__ void transpose_fill_kernel(int nnz, const int *__restrict__ perm, const int *__restrict__ rows,
                                      const double *__restrict__ A_a, int *__restrict__ T_j, double *__restrict__ T_a) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  int src = perm[k];
  T_j[k] = rows[src];
  if (T_a) T_a[k] = A_a[src];
}

// ==========================================================================================
// SpGEMM C = A*B (csr_matop.c:295-473)
// ==========================================================================================
__global__ void spgemm_ub_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                 const int *__restrict__ B_i, int allsquare, int ncols_B, int *__restrict__ cap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { cap[n] = 0; return; }
  long long ub = allsquare ? 1 : 0;
  for (int ia = A_i[i]; ia < A_i[i + 1]; ia++) { int ja = A_j[ia]; ub += B_i[ja + 1] - B_i[ja]; }
  if (ub > ncols_B) ub = ncols_B;
  cap[i] = cap_for((int)ub);
}
__global__ void spgemm_count_kernel(int r0, int r1, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                    const int *__restrict__ B_i, const int *__restrict__ B_j, int allsquare, int diag_base,
                                    const long long *__restrict__ scan, long long base, int *__restrict__ keys,
                                    int *__restrict__ vals, int *__restrict__ cnt) {
  int ic = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= r1) return;
  long long off = scan[ic] - base;
  int cap = (int)(scan[ic + 1] - scan[ic]);
  if (cap == 0) { cnt[ic - r0] = 0; return; }
  Tab t{keys + off, vals + off, (unsigned)(cap - 1)};
  int n = 0;
  if (allsquare) { tab_insert(t, diag_base + ic, n); n++; }      // diagonal first (:384-388, :442-448)
  for (int ia = A_i[ic]; ia < A_i[ic + 1]; ia++) {
    int ja = A_j[ia];
    for (int ib = B_i[ja]; ib < B_i[ja + 1]; ib++) {
      if (tab_insert(t, B_j[ib], n)) n++;
    }
  }
  cnt[ic - r0] = n;
}
__global__ void spgemm_fill_kernel(int r0, int r1, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                   const double *__restrict__ A_a, const int *__restrict__ B_i,
                                   const int *__restrict__ B_j, const double *__restrict__ B_a,
                                   const long long *__restrict__ scan, long long base, const int *__restrict__ keys,
                                   const int *__restrict__ vals, const int *__restrict__ C_i, int *__restrict__ C_j,
                                   double *__restrict__ C_a) {
  int ic = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= r1) return;
  long long off = scan[ic] - base;
  int cap = (int)(scan[ic + 1] - scan[ic]);
  if (cap == 0) return;
  const int start = C_i[ic - r0];
  Tab t{const_cast<int *>(keys) + off, const_cast<int *>(vals) + off, (unsigned)(cap - 1)};
  for (int s = 0; s < cap; s++) {
    int key = t.k[s];
    if (key != -1) { C_j[start + t.v[s]] = key; C_a[start + t.v[s]] = 0.0; }
  }
  for (int ia = A_i[ic]; ia < A_i[ic + 1]; ia++) {
    const int ja = A_j[ia];
    const double a = A_a[ia];
    for (int ib = B_i[ja]; ib < B_i[ja + 1]; ib++) {
      const int pos = tab_find(t, B_j[ib]);
      C_a[start + pos] += a * B_a[ib];               // first touch: 0 + a*b == a*b
    }
  }
}

// ==========================================================================================
// l1 norms (ams.c:571-790), single-rank (offd empty)
// ==========================================================================================
// hypre_ParCSRComputeL1Norms (ams.c:571-760) / ...L1NormsThreads (ams.c:3398-3650) on one rank.  (size, rest)
// describe the reference's thread blocks (= Gauss-Seidel blocks); one block: size = n, rest = 0.
__global__ void l1_kernel(int n, int size, int rest, const int *__restrict__ A_i, const int *__restrict__ A_j,
                          const double *__restrict__ A_a, int option, double *__restrict__ l1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = 0.0;
  if (option == 5) {                                // relax 7: a_ii itself, 1 where it is zero, no sign handling (ams.c:704-725)
    for (int j = A_i[i]; j < A_i[i + 1]; j++)
      if (A_j[j] == i) { v = A_a[j]; break; }
    l1[i] = (v == 0.0) ? 1.0 : v;
    return;
  }
  if (option == 1) {
    for (int j = A_i[i]; j < A_i[i + 1]; j++) v += 1.0 * fabs(A_a[j]);   // ComputeRowSum type 1, scal 1.0
  } else {                                          // option 4: |a_ii| + 0.5 * off-block entries, Remark 6.2
    const int split = rest * (size + 1);
    int ns, ne;
    if (i < split) { ns = (i / (size + 1)) * (size + 1); ne = ns + size + 1; }
    else { ns = (rest + (i - split) / size) * size + rest; ne = ns + size; }
    double diag = 0.0;
    for (int j = A_i[i]; j < A_i[i + 1]; j++) {
      const int ii = A_j[j];
      if (ii == i) { diag = fabs(A_a[j]); v += fabs(A_a[j]); }
      else if (ii < ns || ii >= ne) v += 0.5 * fabs(A_a[j]);
    }
    if (v <= 4.0 / 3.0 * diag) v = diag;
  }
  if (A_i[i] < A_i[i + 1] && A_a[A_i[i]] < 0.0) v = -v;   // negative diagonal (stored first), ams.c:727-735
  l1[i] = v;
}

// option 1 restricted to the entries whose column carries the row's own C/F marker: what hypre_ParCSRComputeL1Norms builds when
// the cycle relaxes in C/F order (relax_order 1, par_amg_setup.c:3047-3050; hypre_CSRMatrixComputeRowSum with CF_i / CF_j,
// csr_matop.c:1326-1352; the threaded variant ams.c:3495-3507 sums the same entries in the same order)
__global__ void l1_cf_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                             const int *__restrict__ cf, double *__restrict__ l1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cf[i];
  double v = 0.0;
  for (int j = A_i[i]; j < A_i[i + 1]; j++)
    if (cf[A_j[j]] == c) v += 1.0 * fabs(A_a[j]);
  if (A_i[i] < A_i[i + 1] && A_a[A_i[i]] < 0.0) v = -v;
  l1[i] = v;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------
static int alloc_hash(b200_handle h, long long slots, int **keys, int **vals) {
  B200_TRY(b200_dalloc<int>(h, keys, (size_t)slots));
  B200_TRY(b200_dalloc<int>(h, vals, (size_t)slots));
  return 0;
}
static int clear_keys(b200_handle h, int *keys, long long slots) {
  B200_CUDA(cudaMemsetAsync(keys, 0xFF, sizeof(int) * (size_t)slots, h->stream));   // -1
  return 0;
}

extern "C" int b200_strength(b200_handle h, b200_csr A, double theta, double max_row_sum, b200_csr *out) {
  if (!A || !A->a) B200_FAIL("strength: matrix with values required");
  const int n = A->nrows;
  int *S_i = nullptr;
  B200_TRY(b200_dalloc<int>(h, &S_i, (size_t)n + 1));
  B200_CUDA(cudaMemsetAsync(S_i + n, 0, sizeof(int), h->stream));
  if (n) {
    strength_kernel<false><<<b200_grid(n, TB), TB, 0, h->stream>>>(n, A->i, A->j, A->a, theta, max_row_sum, S_i, nullptr);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_exclusive_scan_inplace(h, S_i, (size_t)n + 1));
  int nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, S_i + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr S = nullptr;
  B200_TRY(b200_csr_alloc(h, n, A->ncols, nnz, false, &S));
  B200_CUDA(cudaMemcpyAsync(S->i, S_i, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
  if (n) {
    strength_kernel<true><<<b200_grid(n, TB), TB, 0, h->stream>>>(n, A->i, A->j, A->a, theta, max_row_sum, S->i, S->j);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_dfree(h, S_i));
  *out = S;
  return 0;
}

int b200_pmis_rows_init(b200_handle h, b200_csr S, int seed, long long first_row, int cf_init, int *d_cf, int *iterations);
int b200_pmis_rows(b200_handle h, b200_csr S, int seed, long long first_row, int *d_cf, int *iterations) {
  return b200_pmis_rows_init(h, S, seed, first_row, 0, d_cf, iterations);
}
// cf_init 0: hypre_BoomerAMGCoarsenPMIS(S, A, 0, ...); cf_init 3: the second coarsening of aggressive coarsening
// (par_amg_setup.c:1253): isolated rows become C points and the first sweep does not pick an independent set
// (`if (!CF_init || iter)`, par_coarsen.c:2420)
int b200_pmis_rows_init(b200_handle h, b200_csr S, int seed, long long first_row, int cf_init, int *d_cf, int *iterations) {
  if (cf_init != 0 && cf_init != 1 && cf_init != 3) B200_FAIL("pmis: CF_init 0, 1 or 3");
  const int n = S->nrows;
  if (iterations) *iterations = 0;
  if (n == 0) return 0;
  int *colcnt = nullptr, *ingraph = nullptr, *d_count = nullptr, *cf2 = nullptr;
  double *measure = nullptr;
  B200_TRY(b200_dalloc<int>(h, &colcnt, n));
  B200_TRY(b200_dalloc<int>(h, &ingraph, n));
  B200_TRY(b200_dalloc<int>(h, &cf2, n));
  B200_TRY(b200_dalloc<int>(h, &d_count, 2));
  B200_TRY(b200_dalloc<double>(h, &measure, n));
  B200_CUDA(cudaMemsetAsync(colcnt, 0, sizeof(int) * (size_t)n, h->stream));
  if (S->nnz) {
    colcount_kernel<<<b200_grid(S->nnz, 256), 256, 0, h->stream>>>(S->nnz, S->j, colcnt);
    B200_LAUNCH_CHECK();
  }
  pmis_init_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, S->i, colcnt, seed, first_row, measure, d_cf, cf_init);
  B200_LAUNCH_CHECK();
  int iter = 0;
  if (cf_init == 1) {                               // the sweep seeded by the Ruge-Stueben C points (see the kernel)
    pmis_setcf_first_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, S->i, S->j, measure, d_cf, cf2);
    B200_LAUNCH_CHECK();
    pmis_clear_first_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, d_cf, cf2, measure);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(d_cf, cf2, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
    iter = 1;
  }
  // graph flags and node count of the first sweep; every later sweep gets them from the sweep before it
  B200_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int) * 2, h->stream));
  pmis_graph_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, d_cf, ingraph, d_count);
  B200_LAUNCH_CHECK();
  int *cur = d_cf, *nxt = cf2;                      // markers ping-pong between the caller's array and the scratch
  for (int k = 0;; k++) {
    int count = 0;
    B200_CUDA(cudaMemcpyAsync(&count, d_count + (k & 1), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    if (count == 0) break;                          // :2399-2407
    if (!cf_init || iter) {
      pmis_mark_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, measure, cur);
      B200_LAUNCH_CHECK();
      pmis_remove_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, S->i, S->j, measure, ingraph, cur);
      B200_LAUNCH_CHECK();
    }
    B200_CUDA(cudaMemsetAsync(d_count + ((k + 1) & 1), 0, sizeof(int), h->stream));
    pmis_setcf_graph_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, S->i, S->j, ingraph, measure, cur, nxt, d_count + ((k + 1) & 1));
    B200_LAUNCH_CHECK();
    std::swap(cur, nxt);
    if (++iter > 1000) B200_FAIL("PMIS did not terminate");
  }
  if (cur != d_cf) B200_CUDA(cudaMemcpyAsync(d_cf, cur, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
  if (iterations) *iterations = iter;
  B200_TRY(b200_dfree(h, colcnt)); B200_TRY(b200_dfree(h, ingraph)); B200_TRY(b200_dfree(h, cf2));
  B200_TRY(b200_dfree(h, d_count)); B200_TRY(b200_dfree(h, measure));
  return 0;
}

namespace {
__global__ void pmis_ghost_measure_kernel(int ng, const int *__restrict__ cf_ghost, double *__restrict__ m_ghost) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ng && cf_ghost[i] != 0) m_ghost[i] = 0;          // par_coarsen.c:2654-2662
}
}  // namespace

// Row-partitioned PMIS (par_coarsen.c:2031-2738 with num_procs > 1).  S has n local rows and columns in
// the extended local space [owned | ghosts of `halo`]; measures use the GLOBAL row index for the random
// draw (seq_rand, par_indepset.c:44-55), so the result does not depend on the partition.
int b200_pmis_dist_init(b200_handle h, b200_comm c, b200_csr S, b200_halo_s *halo, int seed, long long first_row, int cf_init,
                        int *d_cf_ext);
int b200_pmis_dist(b200_handle h, b200_comm c, b200_csr S, b200_halo_s *halo, int seed, long long first_row, int *d_cf_ext) {
  return b200_pmis_dist_init(h, c, S, halo, seed, first_row, 0, d_cf_ext);
}
// cf_init 3: see b200_pmis_rows_init
int b200_pmis_dist_init(b200_handle h, b200_comm c, b200_csr S, b200_halo_s *halo, int seed, long long first_row, int cf_init,
                        int *d_cf_ext) {
  if (cf_init != 0 && cf_init != 3) B200_FAIL("pmis: CF_init 0 or 3");
  const int n = S->nrows, ng = halo ? halo->ng : 0, ne = n + ng;
  int *colcnt = nullptr, *ingraph = nullptr, *d_count = nullptr, *cf2 = nullptr;
  double *measure = nullptr;
  B200_TRY(b200_dalloc<int>(h, &colcnt, ne));
  B200_TRY(b200_dalloc<int>(h, &ingraph, n));
  B200_TRY(b200_dalloc<int>(h, &cf2, n));
  B200_TRY(b200_dalloc<int>(h, &d_count, 1));
  B200_TRY(b200_dalloc<double>(h, &measure, ne));
  B200_CUDA(cudaMemsetAsync(colcnt, 0, sizeof(int) * (size_t)(ne ? ne : 1), h->stream));
  if (S->nnz) {
    colcount_kernel<<<b200_grid(S->nnz, 256), 256, 0, h->stream>>>(S->nnz, S->j, colcnt);
    B200_LAUNCH_CHECK();
  }
  if (halo) B200_TRY(b200_halo_reverse_add_i32(h, c, halo, colcnt + n, colcnt));     // :2187-2228 (job 2)
  if (n) {
    pmis_init_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, S->i, colcnt, seed, first_row, measure, d_cf_ext, cf_init);
    B200_LAUNCH_CHECK();
  }
  if (halo) {
    B200_TRY(b200_halo_forward_f64(h, c, halo, measure, measure + n));               // :2357-2372 (job 1)
    B200_TRY(b200_halo_forward_i32(h, c, halo, d_cf_ext, d_cf_ext + n));
  }
  int iter = 0;
  while (true) {
    B200_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), h->stream));
    if (n) {
      pmis_graph_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, d_cf_ext, ingraph, d_count);
      B200_LAUNCH_CHECK();
    }
    int count = 0;
    B200_CUDA(cudaMemcpyAsync(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    long long total = count;
    B200_TRY(b200_comm_allreduce_sum_ll(h, c, &total, 1));                           // :2399
    if (total == 0) break;
    if (!cf_init || iter) {                                                          // :2420
      if (ne) {
        pmis_mark_kernel<<<b200_grid(ne, 256), 256, 0, h->stream>>>(ne, measure, d_cf_ext);   // local and ghost nodes (:2430-2449)
        B200_LAUNCH_CHECK();
      }
      if (n) {
        pmis_remove_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, S->i, S->j, measure, ingraph, d_cf_ext);
        B200_LAUNCH_CHECK();
      }
      if (halo) {
        B200_TRY(b200_halo_reverse_clear_i32(h, c, halo, d_cf_ext + n, d_cf_ext));     // job 12 + :2509-2526
        B200_TRY(b200_halo_forward_i32(h, c, halo, d_cf_ext, d_cf_ext + n));           // job 11 (:2530)
      }
    }
    if (n) {
      pmis_setcf_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, S->i, S->j, ingraph, measure, d_cf_ext, cf2);
      B200_LAUNCH_CHECK();
      B200_CUDA(cudaMemcpyAsync(d_cf_ext, cf2, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
    }
    if (halo) {
      B200_TRY(b200_halo_forward_i32(h, c, halo, d_cf_ext, d_cf_ext + n));           // :2603-2617
      if (ng) {
        pmis_ghost_measure_kernel<<<b200_grid(ng, 256), 256, 0, h->stream>>>(ng, d_cf_ext + n, measure + n);
        B200_LAUNCH_CHECK();
      }
    }
    if (++iter > 1000) B200_FAIL("PMIS did not terminate");
  }
  B200_TRY(b200_dfree(h, colcnt)); B200_TRY(b200_dfree(h, ingraph)); B200_TRY(b200_dfree(h, cf2));
  B200_TRY(b200_dfree(h, d_count)); B200_TRY(b200_dfree(h, measure));
  return 0;
}

extern "C" int b200_pmis(b200_handle h, b200_csr S, int seed, int *d_cf) {
  if (!S) B200_FAIL("pmis: null S");
  return b200_pmis_rows(h, S, seed, 0, d_cf, nullptr);
}

// fine_to_coarse numbering; returns the number of coarse points
int b200_coarse_map(b200_handle h, int n, const int *d_cf, int **f2c_out, int *ncoarse) {
  int *f2c = nullptr;
  B200_TRY(b200_dalloc<int>(h, &f2c, (size_t)n + 1));
  cflag_kernel<<<b200_grid((size_t)n + 1, 256), 256, 0, h->stream>>>(n, d_cf, f2c);
  B200_LAUNCH_CHECK();
  B200_TRY(b200_exclusive_scan_inplace(h, f2c, (size_t)n + 1));
  B200_CUDA(cudaMemcpyAsync(ncoarse, f2c + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  *f2c_out = f2c;
  return 0;
}

int b200_extpi_interp_warp(b200_handle h, b200_csr A, b200_csr S, const int *d_cf, int n, const int *d_f2c, int ncoarse,
                           double trunc_factor, int max_elmts, b200_csr *out, int *done);
int b200_csr_multiply_warp(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C,
                           b200_csr *out, int *done);
static bool force_general() {
  const char *e = getenv("B200_FORCE_GENERAL_SETUP");   // test hook: exercise the HBM-scratch kernels
  return e && e[0] == '1';
}

int b200_extpi_interp_ex(b200_handle h, b200_csr A, b200_csr S, const int *d_cf, int n, const int *d_f2c_in, int ncoarse_in,
                         double trunc_factor, int max_elmts, b200_csr *out);

extern "C" int b200_extpi_interp(b200_handle h, b200_csr A, b200_csr S, const int *d_cf, double trunc_factor,
                                 int max_elmts, b200_csr *out) {
  if (!A || !A->a || !S) B200_FAIL("interp: bad arguments");
  return b200_extpi_interp_ex(h, A, S, d_cf, A->nrows, nullptr, 0, trunc_factor, max_elmts, out);
}

// General form used by the multi-rank path: A and S may carry extra (ghost) rows after the first n,
// cf / f2c are indexed in the same extended space, and f2c holds the (global) coarse column ids.
int b200_extpi_interp_ex(b200_handle h, b200_csr A, b200_csr S, const int *d_cf, int n, const int *d_f2c_in, int ncoarse_in,
                         double trunc_factor, int max_elmts, b200_csr *out) {
  if (!force_general()) {
    int done = 0;
    B200_TRY(b200_extpi_interp_warp(h, A, S, d_cf, n, d_f2c_in, ncoarse_in, trunc_factor, max_elmts, out, &done));
    if (done) return 0;
  }
  int *f2c = nullptr, ncoarse = ncoarse_in;
  if (d_f2c_in) f2c = const_cast<int *>(d_f2c_in);
  else B200_TRY(b200_coarse_map(h, n, d_cf, &f2c, &ncoarse));
  int *cap = nullptr;
  B200_TRY(b200_dalloc<int>(h, &cap, (size_t)n + 1));
  extpi_ub_kernel<<<b200_grid((size_t)n + 1, TB), TB, 0, h->stream>>>(n, S->i, S->j, d_cf, cap);
  B200_LAUNCH_CHECK();
  ChunkPlan plan;
  B200_TRY(plan_chunks(h, n, cap, &plan));
  B200_TRY(b200_dfree(h, cap));
  int *keys = nullptr, *vals = nullptr;
  B200_TRY(alloc_hash(h, plan.max_slots, &keys, &vals));
  const int nchunks = (int)plan.bounds.size() - 1;
  int *P_cnt = nullptr;                       // final row lengths, then row pointer
  B200_TRY(b200_dalloc<int>(h, &P_cnt, (size_t)n + 1));
  B200_CUDA(cudaMemsetAsync(P_cnt + n, 0, sizeof(int), h->stream));
  struct Chunk { int r0, r1; int *Pf_i; int *Pf_j; double *Pf_a; };
  std::vector<Chunk> chunks;
  for (int c = 0; c < nchunks; c++) {
    const int r0 = plan.bounds[c], r1 = plan.bounds[c + 1], nr = r1 - r0;
    if (nr == 0) continue;
    const long long slots = plan.base[c + 1] - plan.base[c];
    if (slots) B200_TRY(clear_keys(h, keys, slots));
    Chunk ck{r0, r1, nullptr, nullptr, nullptr};
    B200_TRY(b200_dalloc<int>(h, &ck.Pf_i, (size_t)nr + 1));
    B200_CUDA(cudaMemsetAsync(ck.Pf_i + nr, 0, sizeof(int), h->stream));
    extpi_discover_kernel<<<b200_grid(nr, TB), TB, 0, h->stream>>>(r0, r1, S->i, S->j, d_cf, plan.scan, plan.base[c],
                                                                   keys, vals, ck.Pf_i);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_exclusive_scan_inplace(h, ck.Pf_i, (size_t)nr + 1));
    int total = 0;
    B200_CUDA(cudaMemcpyAsync(&total, ck.Pf_i + nr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dalloc<int>(h, &ck.Pf_j, (size_t)total));
    B200_TRY(b200_dalloc<double>(h, &ck.Pf_a, (size_t)total));
    extpi_weights_kernel<<<b200_grid(nr, TB), TB, 0, h->stream>>>(r0, r1, A->i, A->j, A->a, d_cf, f2c, plan.scan,
                                                                  plan.base[c], keys, vals, ck.Pf_i, ck.Pf_j, ck.Pf_a);
    B200_LAUNCH_CHECK();
    truncate_kernel<<<b200_grid(nr, TB), TB, 0, h->stream>>>(nr, ck.Pf_i, ck.Pf_j, ck.Pf_a, trunc_factor, max_elmts,
                                                             P_cnt + r0);
    B200_LAUNCH_CHECK();
    chunks.push_back(ck);
  }
  B200_TRY(b200_exclusive_scan_inplace(h, P_cnt, (size_t)n + 1));
  int nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, P_cnt + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr P = nullptr;
  B200_TRY(b200_csr_alloc(h, n, ncoarse, nnz, true, &P));
  B200_CUDA(cudaMemcpyAsync(P->i, P_cnt, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
  for (auto &ck : chunks) {
    const int nr = ck.r1 - ck.r0;
    compact_rows_kernel<<<b200_grid(nr, TB), TB, 0, h->stream>>>(nr, ck.Pf_i, ck.Pf_j, ck.Pf_a, P->i + ck.r0, P->j, P->a);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_dfree(h, ck.Pf_i)); B200_TRY(b200_dfree(h, ck.Pf_j)); B200_TRY(b200_dfree(h, ck.Pf_a));
  }
  B200_TRY(b200_dfree(h, keys)); B200_TRY(b200_dfree(h, vals));
  B200_TRY(b200_dfree(h, plan.scan)); B200_TRY(b200_dfree(h, P_cnt));
  if (!d_f2c_in) B200_TRY(b200_dfree(h, f2c));
  *out = P;
  return 0;
}

extern "C" int b200_csr_transpose(b200_handle h, b200_csr A, b200_csr *out) {
  if (!A) B200_FAIL("transpose: null matrix");
  const int n = A->nrows, m = A->ncols, nnz = A->nnz;
  b200_csr T = nullptr;
  B200_TRY(b200_csr_alloc(h, m, n, nnz, A->a != nullptr, &T));
  B200_CUDA(cudaMemsetAsync(T->i, 0, sizeof(int) * ((size_t)m + 1), h->stream));
  if (nnz) {
    int *rows = nullptr, *idx = nullptr, *keys_out = nullptr, *perm = nullptr;
    B200_TRY(b200_dalloc<int>(h, &rows, nnz));
    B200_TRY(b200_dalloc<int>(h, &idx, nnz));
    B200_TRY(b200_dalloc<int>(h, &keys_out, nnz));
    B200_TRY(b200_dalloc<int>(h, &perm, nnz));
    expand_rows_kernel<<<b200_grid(n, TB), TB, 0, h->stream>>>(n, A->i, rows);
    B200_LAUNCH_CHECK();
    iota_kernel<<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, idx);
    B200_LAUNCH_CHECK();
    colcount_kernel<<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, A->j, T->i);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_exclusive_scan_inplace(h, T->i, (size_t)m + 1));
    int bits = 1;
    while ((1LL << bits) < (long long)m) bits++;
    size_t tb = 0;
    B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, A->j, keys_out, idx, perm, nnz, 0, bits, h->stream));
    char *tmp = nullptr;
    B200_TRY(b200_dalloc<char>(h, &tmp, tb));
    B200_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, A->j, keys_out, idx, perm, nnz, 0, bits, h->stream));
    ++g_b200_launches;
    transpose_fill_kernel<<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, perm, rows, A->a, T->j, T->a);
    B200_LAUNCH_CHECK();
    B200_TRY(b200_dfree(h, rows)); B200_TRY(b200_dfree(h, idx)); B200_TRY(b200_dfree(h, keys_out));
    B200_TRY(b200_dfree(h, perm)); B200_TRY(b200_dfree(h, tmp));
  }
  *out = T;
  return 0;
}

int b200_csr_multiply_ex(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C, b200_csr *out);

extern "C" int b200_csr_multiply(b200_handle h, b200_csr A, b200_csr B, b200_csr *out) {
  if (!A || !B || !A->a || !B->a) B200_FAIL("multiply: matrices with values required");
  if (A->ncols != B->nrows) B200_FAIL("multiply: incompatible matrix dimensions");   // csr_matop.c:334-338
  return b200_csr_multiply_ex(h, A, B, (A->nrows == B->ncols) ? 1 : 0, 0, B->ncols, out);
}

// General form used by the multi-rank path: A's column ids index the rows of B (B may hold ghost
// rows), B's column ids may be global, `allsquare` and the diagonal's column id (diag_base + row)
// are given by the caller.
int b200_csr_multiply_ex(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C, b200_csr *out) {
  if (!force_general()) {
    int done = 0;
    B200_TRY(b200_csr_multiply_warp(h, A, B, allsquare, diag_base, ncols_C, out, &done));
    if (done) return 0;
  }
  const int n = A->nrows;
  int *cap = nullptr;
  B200_TRY(b200_dalloc<int>(h, &cap, (size_t)n + 1));
  spgemm_ub_kernel<<<b200_grid((size_t)n + 1, TB), TB, 0, h->stream>>>(n, A->i, A->j, B->i, allsquare, ncols_C, cap);
  B200_LAUNCH_CHECK();
  ChunkPlan plan;
  B200_TRY(plan_chunks(h, n, cap, &plan));
  B200_TRY(b200_dfree(h, cap));
  int *keys = nullptr, *vals = nullptr;
  B200_TRY(alloc_hash(h, plan.max_slots, &keys, &vals));
  const int nchunks = (int)plan.bounds.size() - 1;
  int *C_cnt = nullptr;
  B200_TRY(b200_dalloc<int>(h, &C_cnt, (size_t)n + 1));
  B200_CUDA(cudaMemsetAsync(C_cnt, 0, sizeof(int) * ((size_t)n + 1), h->stream));
  struct Chunk { int r0, r1; int *Ci; int *Cj; double *Ca; };
  std::vector<Chunk> chunks;
  for (int c = 0; c < nchunks; c++) {
    const int r0 = plan.bounds[c], r1 = plan.bounds[c + 1], nr = r1 - r0;
    if (nr == 0) continue;
    const long long slots = plan.base[c + 1] - plan.base[c];
    if (slots) B200_TRY(clear_keys(h, keys, slots));
    Chunk ck{r0, r1, nullptr, nullptr, nullptr};
    B200_TRY(b200_dalloc<int>(h, &ck.Ci, (size_t)nr + 1));
    B200_CUDA(cudaMemsetAsync(ck.Ci + nr, 0, sizeof(int), h->stream));
    spgemm_count_kernel<<<b200_grid(nr, TB), TB, 0, h->stream>>>(r0, r1, A->i, A->j, B->i, B->j, allsquare, diag_base, plan.scan,
                                                                 plan.base[c], keys, vals, ck.Ci);
    B200_LAUNCH_CHECK();
    B200_CUDA(cudaMemcpyAsync(C_cnt + r0, ck.Ci, sizeof(int) * (size_t)nr, cudaMemcpyDeviceToDevice, h->stream));
    B200_TRY(b200_exclusive_scan_inplace(h, ck.Ci, (size_t)nr + 1));
    int total = 0;
    B200_CUDA(cudaMemcpyAsync(&total, ck.Ci + nr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dalloc<int>(h, &ck.Cj, (size_t)total));
    B200_TRY(b200_dalloc<double>(h, &ck.Ca, (size_t)total));
    spgemm_fill_kernel<<<b200_grid(nr, TB), TB, 0, h->stream>>>(r0, r1, A->i, A->j, A->a, B->i, B->j, B->a, plan.scan,
                                                                plan.base[c], keys, vals, ck.Ci, ck.Cj, ck.Ca);
    B200_LAUNCH_CHECK();
    chunks.push_back(ck);
  }
  B200_TRY(b200_exclusive_scan_inplace(h, C_cnt, (size_t)n + 1));
  int nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, C_cnt + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr Cm = nullptr;
  B200_TRY(b200_csr_alloc(h, n, ncols_C, nnz, true, &Cm));
  B200_CUDA(cudaMemcpyAsync(Cm->i, C_cnt, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
  for (auto &ck : chunks) {
    const int nr = ck.r1 - ck.r0;
    if (chunks.size() == 1) {
      // single chunk: rows are already contiguous in row order
      B200_CUDA(cudaMemcpyAsync(Cm->j, ck.Cj, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
      B200_CUDA(cudaMemcpyAsync(Cm->a, ck.Ca, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
    } else {
      compact_rows_kernel<<<b200_grid(nr, TB), TB, 0, h->stream>>>(nr, ck.Ci, ck.Cj, ck.Ca, Cm->i + ck.r0, Cm->j, Cm->a);
      B200_LAUNCH_CHECK();
    }
    B200_TRY(b200_dfree(h, ck.Ci)); B200_TRY(b200_dfree(h, ck.Cj)); B200_TRY(b200_dfree(h, ck.Ca));
  }
  B200_TRY(b200_dfree(h, keys)); B200_TRY(b200_dfree(h, vals));
  B200_TRY(b200_dfree(h, plan.scan)); B200_TRY(b200_dfree(h, C_cnt));
  *out = Cm;
  return 0;
}

extern "C" int b200_l1_norms_blocks(b200_handle h, b200_csr A, int option, int blocks, double *d_l1) {
  if (!A || !A->a) B200_FAIL("l1 norms: matrix with values required");
  if (option != 1 && option != 4 && option != 5) B200_FAIL("l1 norms: options 1, 4 and 5 are supported");
  if (blocks < 1) B200_FAIL("l1 norms: blocks must be >= 1");
  if (A->nrows == 0) return 0;
  const int size = A->nrows / blocks, rest = A->nrows - size * blocks;
  l1_kernel<<<b200_grid(A->nrows, TB), TB, 0, h->stream>>>(A->nrows, size, rest, A->i, A->j, A->a, option, d_l1);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_l1_norms_cf(b200_handle h, b200_csr A, const int *d_cf, double *d_l1) {
  if (!A || !A->a || !d_cf) B200_FAIL("l1_norms_cf: matrix with values and a C/F marker required");
  if (A->nrows == 0) return 0;
  l1_cf_kernel<<<b200_grid(A->nrows, 256), 256, 0, h->stream>>>(A->nrows, A->i, A->j, A->a, d_cf, d_l1);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_l1_norms(b200_handle h, b200_csr A, int option, double *d_l1) {
  return b200_l1_norms_blocks(h, A, option, 1, d_l1);
}
