// b200_spmv_dict.cu -- dictionary-compressed solve copy of a stencil-structured CSR operator and its SpMV kernel (sm_100a).
//
// The streaming SpMV (b200_spmv_pipe.cu) sits on the HBM roof for the finest-level operator: 12 bytes per entry (int32 column +
// FP64 value) is what it moves, and three passes over A_0 are a third of every PCG iteration.  Operators that come from a grid
// have two properties a general CSR kernel leaves on the table:
//   * the column of an entry is its row plus one of a FEW offsets (7 for the 7-point operator, 27 for the 27-point one; a
//     row-partitioned block adds a constant offset per neighbour rank, because ghost columns are numbered in the order of the
//     boundary rows that reference them);
//   * with constant coefficients the entries take a FEW distinct values.
// b200_csr_build_dict scans the whole matrix once, collects the distinct (column - row) offsets and the distinct value bit
// patterns (exact: every entry is looked at, 64-bit patterns are compared, nothing is rounded), and if either set has at most 255
// members stores one BYTE per entry in its place: 12 -> 9, 5 or 2 bytes per entry.  The kernel below is spmv_pipe_kernel's
// G = 1 form (one lane per row, rows consumed straight from the staged stream, TMA bulk copies + mbarrier ring) reading codes and
// two small shared-memory tables; the products and their order are those of the uncompressed kernel, so the result is
// bit-identical to it.  Matrices that do not compress (the coarse Galerkin operators: thousands of offsets) keep the plain path;
// a matrix whose values are updated in place drops its dictionary (b200_csr_drop_dict) and gets a new one on the next product.
// Reference semantics: hypre_CSRMatrixMatvecOutOfPlaceHost (seq_mv/csr_matvec.c:24-376).
#include "b200_internal.h"
#include <algorithm>
#include <atomic>
#include <vector>

namespace {

constexpr int NT = B200_SPMV_NT;
constexpr int DICT_MAX = 255;          // members per dictionary (codes are bytes)
constexpr int DSLOTS = 1024;           // hash slots of the collecting sets
constexpr int EMPTY_OFF = (int)0x80000000;
constexpr unsigned long long EMPTY_VAL = 0xFFF8B200DEAD0001ull;   // a NaN payload no operator carries; if one does, no dictionary
constexpr int RCAP = B200_SPMV_RCAP;   // row pointers staged per tile (tiles with more rows read A_i from global)
constexpr int DSCAP_MAX = 2048;

__device__ __forceinline__ unsigned dhash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 29;
  return (unsigned)k & (DSLOTS - 1);
}
// insert into an open-addressing set; returns false when the set is full beyond `limit` members
template <class K>
__device__ __forceinline__ bool set_insert(K *keys, K empty, K key, int *count, int limit) {
  unsigned s = dhash((unsigned long long)key);
  for (int probe = 0; probe < DSLOTS; probe++) {
    K cur = keys[s];
    if (cur == key) return true;
    if (cur == empty) {
      const K prev = atomicCAS(&keys[s], empty, key);
      if (prev == empty) return atomicAdd(count, 1) < limit;
      if (prev == key) return true;
    }
    s = (s + 1) & (DSLOTS - 1);
  }
  return false;
}

__device__ __forceinline__ bool t_all_bad(const int *s_bad) {
  return *reinterpret_cast<const volatile int *>(&s_bad[0]) && *reinterpret_cast<const volatile int *>(&s_bad[1]);
}
// distinct (col - row) offsets and distinct value patterns of the whole matrix: CTA-local sets in shared memory, merged into the
// global sets at the end (a few hundred atomics per CTA instead of one per entry)
__global__ void __launch_bounds__(256)
dict_collect_kernel(int nrows, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                    int *__restrict__ g_off, unsigned long long *__restrict__ g_val, int *__restrict__ g_cnt /* [0] offsets, [1] values, [2] flags */) {
  __shared__ int s_off[DSLOTS];
  __shared__ unsigned long long s_val[DSLOTS];
  __shared__ int s_cnt[2], s_bad[2];
  for (int s = threadIdx.x; s < DSLOTS; s += blockDim.x) { s_off[s] = EMPTY_OFF; s_val[s] = EMPTY_VAL; }
  if (threadIdx.x < 2) { s_cnt[threadIdx.x] = 0; s_bad[threadIdx.x] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  // one warp per 32 consecutive rows, lanes over the entries of the chunk (coalesced reads of A_j / A_a); every lane runs the
  // same number of trips (the row search shuffles across the whole warp)
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); (long long)w * 32 < nrows; w += nwarps) {
    const int stop = __shfl_sync(0xffffffffu, (lane == 0) ? *reinterpret_cast<volatile int *>(&g_cnt[2]) : 0, 0);
    if ((stop & 3) == 3) break;                             // neither dictionary can exist any more
    const int r0 = w * 32, r1 = min(nrows, r0 + 32);
    const int e0 = A_i[r0], e1 = A_i[r1];
    const int myrow_end = (r0 + lane < r1) ? A_i[r0 + lane + 1] : 0x7fffffff;      // lane l knows where row r0 + l ends
    const int trips = (e1 - e0 + 31) / 32;
    for (int t = 0; t < trips; t++) {
      const int e = e0 + t * 32 + lane;
      // row of entry e: the first row whose end lies beyond e (32 candidates held across the lanes)
      int r = r0;
#pragma unroll
      for (int l = 0; l < 32; l++) r += (__shfl_sync(0xffffffffu, myrow_end, l) <= e) ? 1 : 0;
      if (e >= e1) continue;
      if (!s_bad[0] && !set_insert<int>(s_off, EMPTY_OFF, A_j[e] - r, &s_cnt[0], DICT_MAX)) s_bad[0] = 1;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(A_a[e]);
      if (!s_bad[1] && (bits == EMPTY_VAL || !set_insert<unsigned long long>(s_val, EMPTY_VAL, bits, &s_cnt[1], DICT_MAX))) s_bad[1] = 1;
    }
    if (t_all_bad(s_bad)) {
      if (lane == 0) atomicOr(&g_cnt[2], 3);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_bad[0]) atomicOr(&g_cnt[2], 1);
    if (s_bad[1]) atomicOr(&g_cnt[2], 2);
  }
  for (int s = threadIdx.x; s < DSLOTS; s += blockDim.x) {
    if (!s_bad[0] && s_off[s] != EMPTY_OFF && !set_insert<int>(g_off, EMPTY_OFF, s_off[s], &g_cnt[0], DICT_MAX)) atomicOr(&g_cnt[2], 1);
    if (!s_bad[1] && s_val[s] != EMPTY_VAL && !set_insert<unsigned long long>(g_val, EMPTY_VAL, s_val[s], &g_cnt[1], DICT_MAX)) atomicOr(&g_cnt[2], 2);
  }
}

// codes = position of the entry's offset / value pattern in the SORTED tables (binary search in shared memory)
__global__ void __launch_bounds__(256)
dict_encode_kernel(int nrows, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                   const int *__restrict__ off_tab, int n_off, const unsigned long long *__restrict__ val_tab, int n_val,
                   unsigned char *__restrict__ jc, unsigned char *__restrict__ ac) {
  __shared__ int s_off[256];
  __shared__ unsigned long long s_val[256];
  for (int s = threadIdx.x; s < 256; s += blockDim.x) {
    s_off[s] = (jc && s < n_off) ? off_tab[s] : 0x7fffffff;
    s_val[s] = (ac && s < n_val) ? val_tab[s] : ~0ull;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); (long long)w * 32 < nrows; w += nwarps) {
    const int r0 = w * 32, r1 = min(nrows, r0 + 32);
    const int e0 = A_i[r0], e1 = A_i[r1];
    const int myrow_end = (r0 + lane < r1) ? A_i[r0 + lane + 1] : 0x7fffffff;
    const int trips = (e1 - e0 + 31) / 32;
    for (int t = 0; t < trips; t++) {
      const int e = e0 + t * 32 + lane;
      int r = r0;
#pragma unroll
      for (int l = 0; l < 32; l++) r += (__shfl_sync(0xffffffffu, myrow_end, l) <= e) ? 1 : 0;
      if (e >= e1) continue;
      if (jc) {
        const int off = A_j[e] - r;
        int lo = 0, hi = n_off - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_off[mid] < off) lo = mid + 1; else hi = mid; }
        jc[e] = (unsigned char)lo;
      }
      if (ac) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(A_a[e]);
        int lo = 0, hi = n_val - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_val[mid] < bits) lo = mid + 1; else hi = mid; }
        ac[e] = (unsigned char)lo;
      }
    }
  }
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------------
struct Epi {
  int mode;            // 0: y = alpha*s + beta*b    1: y = x[r] + w*(b[r]-s)/d[r]  (l1-Jacobi)
  double alpha, beta;
  const double *b;
  const double *d;
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DDONE;\n"
      "bra DWAIT_LOOP;\n"
      "DDONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// DC: columns as offset codes;  DV: values as codes.  One lane per row (the plan's G is 1 for these operators).
template <int NSTAGE, bool DC, bool DV>
__global__ void __launch_bounds__(NT)
spmv_dict_kernel(const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                 const unsigned char *__restrict__ jc, const unsigned char *__restrict__ ac,
                 const int *__restrict__ off_tab, const double *__restrict__ val_tab,
                 const int4 *__restrict__ blk_meta, int nblk, int SCAP,
                 const double *__restrict__ x, double *__restrict__ y, Epi epi) {
  constexpr int VB = DV ? 1 : 8, CB = DC ? 1 : 4;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *vals_b = smem;                                             // [NSTAGE][SCAP * VB]
  unsigned char *cols_b = smem + (size_t)SCAP * VB * NSTAGE;                // [NSTAGE][SCAP * CB]
  int *rps = reinterpret_cast<int *>(cols_b + (size_t)SCAP * CB * NSTAGE);  // [NSTAGE][RCAP]
  unsigned long long *full = reinterpret_cast<unsigned long long *>(rps + RCAP * NSTAGE);
  __shared__ int s_off[256];
  __shared__ double s_val[256];
  const int tid = threadIdx.x;
  if (DC) for (int s = tid; s < 256; s += NT) s_off[s] = off_tab[s];
  if (DV) for (int s = tid; s < 256; s += NT) s_val[s] = val_tab[s];
  const int ntiles = (nblk - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int k, const int4 m) {      // thread 0 only; m = {r0, r1, e0, e1} of tile k
    const int stage = k % NSTAGE;
    const int a0 = m.z & ~15, len = m.w - a0;      // 16-entry alignment: 16 bytes of codes
    const int ra = m.x & ~3, rlen = m.y - ra + 1;
    const unsigned bv = (m.w > m.z) ? (((unsigned)len * VB + 15u) & ~15u) : 0u;
    const unsigned bc = (m.w > m.z) ? (((unsigned)len * CB + 15u) & ~15u) : 0u;
    const unsigned br = (rlen <= RCAP) ? (((unsigned)rlen * 4u + 15u) & ~15u) : 0u;
    if (bv + bc + br == 0) { mbar_arrive(&full[stage]); return; }
    mbar_expect_tx(&full[stage], bv + bc + br);
    if (bv) {
      if (DV) bulk_g2s(vals_b + (size_t)stage * SCAP * VB, ac + a0, bv, &full[stage]);
      else bulk_g2s(vals_b + (size_t)stage * SCAP * VB, A_a + a0, bv, &full[stage]);
      if (DC) bulk_g2s(cols_b + (size_t)stage * SCAP * CB, jc + a0, bc, &full[stage]);
      else bulk_g2s(cols_b + (size_t)stage * SCAP * CB, A_j + a0, bc, &full[stage]);
    }
    if (br) bulk_g2s(rps + (size_t)stage * RCAP, A_i + ra, br, &full[stage]);
  };
  if (tid == 0) {
    for (int k = 0; k < NSTAGE && k < ntiles; k++) issue(k, blk_meta[blockIdx.x + k * gridDim.x]);
  }

  int4 m = blk_meta[blockIdx.x];
  for (int k = 0; k < ntiles; k++) {
    const int stage = k % NSTAGE;
    const unsigned parity = (unsigned)((k / NSTAGE) & 1);
    const int r0 = m.x, r1 = m.y, e0 = m.z;
    const int a0 = e0 & ~15, ra = r0 & ~3;
    int4 m_next = m, m_refill = m;
    if (k + 1 < ntiles) m_next = blk_meta[blockIdx.x + (k + 1) * gridDim.x];
    if (tid == 0 && k + NSTAGE < ntiles) m_refill = blk_meta[blockIdx.x + (k + NSTAGE) * gridDim.x];
    const unsigned char *pvb = vals_b + (size_t)stage * SCAP * VB;
    const unsigned char *pcb = cols_b + (size_t)stage * SCAP * CB;
    const double *pv = reinterpret_cast<const double *>(pvb);
    const int *pc = reinterpret_cast<const int *>(pcb);
    const int *rp = rps + (size_t)stage * RCAP;
    const bool rp_smem = (r1 - ra + 1) <= RCAP;
    const int rfirst = r0 + tid;
    double bf = 0.0, df = 1.0, xf = 0.0;
    if (rfirst < r1) {
      if (epi.mode == 0) { if (epi.beta != 0.0) bf = epi.b[rfirst]; }
      else { bf = epi.b[rfirst]; df = epi.d[rfirst]; xf = x[rfirst]; }
    }
    mbar_wait(&full[stage], parity);
    for (int base = r0; base < r1; base += NT) {
      const int r = base + tid;
      double s = 0.0;
      if (r < r1) {
        int s0, s1;
        if (rp_smem) { s0 = rp[r - ra]; s1 = rp[r - ra + 1]; } else { s0 = A_i[r]; s1 = A_i[r + 1]; }
        s0 -= a0; s1 -= a0;
        auto col = [&](int p) { return DC ? r + s_off[pcb[p]] : pc[p]; };
        auto val = [&](int p) { return DV ? s_val[pvb[p]] : pv[p]; };
        int p = s0;
        for (; p + 3 < s1; p += 4) {                  // four gathers in flight per lane; products added in storage order
          const int c0 = col(p), c1 = col(p + 1), c2 = col(p + 2), c3 = col(p + 3);
          const double v0 = val(p), v1 = val(p + 1), v2 = val(p + 2), v3 = val(p + 3);
          const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
          s += v0 * x0; s += v1 * x1; s += v2 * x2; s += v3 * x3;
        }
        if (p < s1) {
          const bool k1 = p + 1 < s1, k2 = p + 2 < s1;
          const int c0 = col(p), c1 = k1 ? col(p + 1) : c0, c2 = k2 ? col(p + 2) : c0;
          const double v0 = val(p), v1 = k1 ? val(p + 1) : 0.0, v2 = k2 ? val(p + 2) : 0.0;
          const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2);
          s += v0 * x0;
          if (k1) s += v1 * x1;
          if (k2) s += v2 * x2;
        }
        double bb = bf, dd = df, xx = xf;
        if (base != r0) {
          if (epi.mode == 0) { if (epi.beta != 0.0) bb = epi.b[r]; }
          else { bb = epi.b[r]; dd = epi.d[r]; xx = x[r]; }
        }
        if (epi.mode == 0) {
          double v = epi.alpha * s;
          if (epi.beta != 0.0) v += epi.beta * bb;
          y[r] = v;
        } else {
          y[r] = xx + epi.alpha * (bb - s) / dd;
        }
      }
    }
    __syncthreads();
    if (tid == 0 && k + NSTAGE < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(k + NSTAGE, m_refill);
    }
    m = m_next;
  }
}

template <bool DC, bool DV, int NSTAGE>
int launch_dict_n(b200_handle h, b200_csr A, const double *x, double *y, const Epi &epi) {
  constexpr int VB = DV ? 1 : 8, CB = DC ? 1 : 4;
  int scap = (A->tile + A->max_row + 16 + 31) & ~31;
  const size_t bytes = ((size_t)(VB + CB) * scap + sizeof(int) * RCAP) * NSTAGE + sizeof(unsigned long long) * NSTAGE;
  static std::atomic<unsigned long long> attr_set{0};
  const unsigned long long bit = 1ull << (h->device & 63);
  if (!(attr_set.load(std::memory_order_acquire) & bit)) {
    const size_t maxb = ((size_t)(VB + CB) * DSCAP_MAX + sizeof(int) * RCAP) * NSTAGE + 64;
    B200_CUDA(cudaFuncSetAttribute(spmv_dict_kernel<NSTAGE, DC, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxb));
    attr_set.fetch_or(bit, std::memory_order_release);
  }
  static std::atomic<int> occ_cache[128];
  std::atomic<int> &slot = occ_cache[(scap / 32) & 127];
  int occ = slot.load(std::memory_order_relaxed);
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_dict_kernel<NSTAGE, DC, DV>, NT, bytes) != cudaSuccess || occ < 1) {
      cudaGetLastError();
      occ = 8;
    }
    slot.store(occ, std::memory_order_relaxed);
  }
  int grid = h->num_sm * occ;
  if (grid > A->nblk) grid = A->nblk;
  spmv_dict_kernel<NSTAGE, DC, DV><<<grid, NT, bytes, h->stream>>>(A->i, A->j, A->a, A->jc, A->ac, A->off_tab, A->val_tab,
                                                                  reinterpret_cast<const int4 *>(A->blk_meta), A->nblk, scap, x, y, epi);
  B200_LAUNCH_CHECK();
  return 0;
}

// stages of the ring: a tile of codes is under 2 KB, so the ring can be deep without costing occupancy (B200_DICT_STAGES)
template <bool DC, bool DV>
int launch_dict(b200_handle h, b200_csr A, const double *x, double *y, const Epi &epi) {
  static const int stages = [] { const char *e = getenv("B200_DICT_STAGES"); return e ? atoi(e) : 3; }();   // 256^3 7-pt: 2 -> 0.206, 3 -> 0.186, 4 -> 0.188, 6 -> 0.198 ms
  if (stages >= 6) return launch_dict_n<DC, DV, 6>(h, A, x, y, epi);
  if (stages >= 4) return launch_dict_n<DC, DV, 4>(h, A, x, y, epi);
  if (stages == 3) return launch_dict_n<DC, DV, 3>(h, A, x, y, epi);
  return launch_dict_n<DC, DV, 2>(h, A, x, y, epi);
}

}  // namespace

int b200_csr_drop_dict(b200_handle h, b200_csr A) {
  if (!A) return 0;
  B200_TRY(b200_dfree(h, A->jc)); B200_TRY(b200_dfree(h, A->ac));
  B200_TRY(b200_dfree(h, A->off_tab)); B200_TRY(b200_dfree(h, A->val_tab));
  A->jc = A->ac = nullptr; A->off_tab = nullptr; A->val_tab = nullptr;
  A->dict_state = 0;
  return 0;
}

// Tries to build the dictionaries of A (needs the SpMV plan).  dict_state: 1 = at least one dictionary exists, -1 = none applies.
int b200_csr_build_dict(b200_handle h, b200_csr A) {
  static const int enabled = [] { const char *e = getenv("B200_SPMV_DICT"); return e ? atoi(e) : 3; }();   // bit 0: columns, bit 1: values
  static const int min_nnz = [] { const char *e = getenv("B200_SPMV_DICT_MIN_NNZ"); return e ? atoi(e) : (1 << 20); }();
  A->dict_state = -1;
  if (!enabled || !A->a || !A->blk_meta || A->group != 1 || A->nnz < min_nnz || A->max_row <= 0) return 0;
  if (A->tile + A->max_row + 16 > DSCAP_MAX) return 0;
  int *g_off = nullptr, *g_cnt = nullptr;
  unsigned long long *g_val = nullptr;
  B200_TRY(b200_dalloc<int>(h, &g_off, DSLOTS));
  B200_TRY(b200_dalloc<unsigned long long>(h, &g_val, DSLOTS));
  B200_TRY(b200_dalloc<int>(h, &g_cnt, 4));
  std::vector<int> init_off(DSLOTS, EMPTY_OFF);
  std::vector<unsigned long long> init_val(DSLOTS, EMPTY_VAL);
  B200_CUDA(cudaMemcpyAsync(g_off, init_off.data(), sizeof(int) * DSLOTS, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(g_val, init_val.data(), sizeof(unsigned long long) * DSLOTS, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemsetAsync(g_cnt, 0, sizeof(int) * 4, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));                     // the staging vectors go out of scope below
  const int grid = h->num_sm * 4;
  dict_collect_kernel<<<grid, 256, 0, h->stream>>>(A->nrows, A->i, A->j, A->a, g_off, g_val, g_cnt);
  B200_LAUNCH_CHECK();
  std::vector<int> h_off(DSLOTS);
  std::vector<unsigned long long> h_val(DSLOTS);
  int h_cnt[4] = {0, 0, 0, 0};
  B200_CUDA(cudaMemcpyAsync(h_off.data(), g_off, sizeof(int) * DSLOTS, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(h_val.data(), g_val, sizeof(unsigned long long) * DSLOTS, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(h_cnt, g_cnt, sizeof(int) * 4, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, g_off)); B200_TRY(b200_dfree(h, g_val)); B200_TRY(b200_dfree(h, g_cnt));
  std::vector<int> offs;
  std::vector<unsigned long long> vals;
  for (int s = 0; s < DSLOTS; s++) {
    if (h_off[s] != EMPTY_OFF) offs.push_back(h_off[s]);
    if (h_val[s] != EMPTY_VAL) vals.push_back(h_val[s]);
  }
  const bool dc = (enabled & 1) && !(h_cnt[2] & 1) && !offs.empty() && (int)offs.size() <= DICT_MAX;
  const bool dv = (enabled & 2) && !(h_cnt[2] & 2) && !vals.empty() && (int)vals.size() <= DICT_MAX;
  if (!dc && !dv) return 0;
  std::sort(offs.begin(), offs.end());
  std::sort(vals.begin(), vals.end());
  int n_off = 0, n_val = 0;
  unsigned long long *d_valbits = nullptr;
  if (dc) {
    n_off = (int)offs.size();
    offs.resize(256, 0);
    B200_TRY(b200_dalloc<int>(h, &A->off_tab, 256));
    B200_CUDA(cudaMemcpyAsync(A->off_tab, offs.data(), sizeof(int) * 256, cudaMemcpyHostToDevice, h->stream));
    B200_TRY(b200_dalloc<unsigned char>(h, &A->jc, (size_t)A->nnz + 64));
  }
  if (dv) {
    n_val = (int)vals.size();
    vals.resize(256, 0);
    B200_TRY(b200_dalloc<double>(h, &A->val_tab, 256));
    B200_CUDA(cudaMemcpyAsync(A->val_tab, vals.data(), sizeof(double) * 256, cudaMemcpyHostToDevice, h->stream));   // same bits
    d_valbits = reinterpret_cast<unsigned long long *>(A->val_tab);
    B200_TRY(b200_dalloc<unsigned char>(h, &A->ac, (size_t)A->nnz + 64));
  }
  B200_CUDA(cudaStreamSynchronize(h->stream));                     // host tables are read by the copies above
  dict_encode_kernel<<<grid, 256, 0, h->stream>>>(A->nrows, A->i, A->j, A->a, A->off_tab, n_off, d_valbits, n_val, A->jc, A->ac);
  B200_LAUNCH_CHECK();
  A->dict_state = 1;
  static const bool dbg = [] { const char *e = getenv("B200_DEBUG_PLAN"); return e && e[0] == '1'; }();
  if (dbg) fprintf(stderr, "[b200 dict] rows=%d nnz=%d offsets=%d values=%d -> %d bytes per entry\n", A->nrows, A->nnz, n_off, n_val,
                   (dc ? 1 : 4) + (dv ? 1 : 8));
  return 0;
}

extern "C" int b200_csr_stream_bytes_per_entry(b200_csr A) {
  if (!A) return 0;
  if (A->dict_state != 1) return 12;
  return (A->jc ? 1 : 4) + (A->ac ? 1 : 8);
}

bool b200_spmv_dict_ok(b200_csr A) { return A->dict_state == 1 && (A->jc || A->ac); }

int b200_csr_spmv_dict(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha, double beta,
                       const double *b, const double *d) {
  Epi e{mode, alpha, beta, b, d};
  if (A->jc && A->ac) return launch_dict<true, true>(h, A, x, y, e);
  if (A->jc) return launch_dict<true, false>(h, A, x, y, e);
  return launch_dict<false, true>(h, A, x, y, e);
}
