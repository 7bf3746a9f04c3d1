// b200_spmv_pipe.cu -- persistent, bulk-copy (TMA engine) pipelined CSR SpMV for sm_100a.
//
// Same row-block plan and fused epilogues as the two-phase kernel in b200_csr.cu, but the matrix
// stream never touches the register file on its way in:
//   * each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... of the plan;
//   * one thread issues `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` copies
//     of the tile's (aligned) value and column ranges into a ring of NSTAGE shared-memory stages,
//     arming an mbarrier with the byte count; the copies for tiles k+1.. are in flight while tile k
//     is multiplied and reduced, so HBM streaming is decoupled from the x-gather latency;
//   * all threads wait on the stage's mbarrier; G lanes per row then consume the row STRAIGHT from the staged
//     stream: each lane multiply-adds its entries val*x[col] in registers (x gathered through the read-only path,
//     L2-resident for stencil matrices), the G partial sums are combined with shuffles and the epilogue
//     (axpby / l1-Jacobi) is applied -- no product round trip through shared memory.
// Reference semantics: hypre_CSRMatrixMatvecOutOfPlaceHost (seq_mv/csr_matvec.c:24-376).
#include "b200_internal.h"

namespace {

constexpr int NT = B200_SPMV_NT;
constexpr int SCAP_MAX = 1536;   // largest stage (entries) the kernel is launched with

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
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int RCAP = B200_SPMV_RCAP;   // row pointers staged per tile (tiles with more rows read A_i from global)

// Per-tile latency chain kept short: tile metadata (one int4 per tile, prefetched one tile ahead),
// row pointers (third bulk copy) and the epilogue operands of each thread's first row (prefetched
// before the mbarrier wait) are all off the critical path; what remains after the data lands is
// LDS -> x gather -> DMUL -> BAR -> shared-memory row reduction -> store.
template <int G, int NSTAGE>
__global__ void __launch_bounds__(NT)
spmv_pipe_kernel(const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                 const int4 *__restrict__ blk_meta, int nblk, int SCAP,
                 const double *__restrict__ x, double *__restrict__ y, Epi epi) {
  // SCAP (entries per stage, multiple of 32) is sized per matrix: tile + longest row, so that the
  // shared memory of an SM holds as many in-flight tiles as possible (bytes in flight = bandwidth x latency)
  extern __shared__ __align__(128) unsigned char smem[];
  double *vals = reinterpret_cast<double *>(smem);                                  // [NSTAGE][SCAP]
  int *cols = reinterpret_cast<int *>(smem + sizeof(double) * SCAP * NSTAGE);       // [NSTAGE][SCAP]
  int *rps = cols + SCAP * NSTAGE;                                                  // [NSTAGE][RCAP]
  unsigned long long *full = reinterpret_cast<unsigned long long *>(rps + RCAP * NSTAGE);
  const int tid = threadIdx.x;
  const int ntiles = (nblk - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int k, const int4 m) {      // thread 0 only; m = {r0, r1, e0, e1} of tile k
    const int stage = k % NSTAGE;
    const int a0 = m.z & ~3, len = m.w - a0;
    const int ra = m.x & ~3, rlen = m.y - ra + 1;
    const unsigned bv = (m.w > m.z) ? (((unsigned)len * 8u + 15u) & ~15u) : 0u;
    const unsigned bc = (m.w > m.z) ? (((unsigned)len * 4u + 15u) & ~15u) : 0u;
    const unsigned br = (rlen <= RCAP) ? (((unsigned)rlen * 4u + 15u) & ~15u) : 0u;
    if (bv + bc + br == 0) { mbar_arrive(&full[stage]); return; }
    mbar_expect_tx(&full[stage], bv + bc + br);
    if (bv) {
      bulk_g2s(vals + (size_t)stage * SCAP, A_a + a0, bv, &full[stage]);
      bulk_g2s(cols + (size_t)stage * SCAP, A_j + a0, bc, &full[stage]);
    }
    if (br) bulk_g2s(rps + (size_t)stage * RCAP, A_i + ra, br, &full[stage]);
  };
  if (tid == 0) {
    for (int k = 0; k < NSTAGE && k < ntiles; k++) issue(k, blk_meta[blockIdx.x + k * gridDim.x]);
  }

  const int sub = tid / G, lane = tid % G;
  int4 m = blk_meta[blockIdx.x];
  for (int k = 0; k < ntiles; k++) {
    const int stage = k % NSTAGE;
    const unsigned parity = (unsigned)((k / NSTAGE) & 1);
    const int r0 = m.x, r1 = m.y, e0 = m.z;
    const int a0 = e0 & ~3, ra = r0 & ~3;
    // prefetches for later: next tile's metadata (all threads) and the refill tile's metadata (thread 0)
    int4 m_next = m, m_refill = m;
    if (k + 1 < ntiles) m_next = blk_meta[blockIdx.x + (k + 1) * gridDim.x];
    if (tid == 0 && k + NSTAGE < ntiles) m_refill = blk_meta[blockIdx.x + (k + NSTAGE) * gridDim.x];
    double *pv = vals + (size_t)stage * SCAP;
    const int *pc = cols + (size_t)stage * SCAP;
    const int *rp = rps + (size_t)stage * RCAP;
    const bool rp_smem = (r1 - ra + 1) <= RCAP;
    // epilogue operands of this thread's first row, requested before the wait
    const int rfirst = r0 + sub;
    double bf = 0.0, df = 1.0, xf = 0.0;
    if (rfirst < r1 && lane == 0) {
      if (epi.mode == 0) { if (epi.beta != 0.0) bf = epi.b[rfirst]; }
      else { bf = epi.b[rfirst]; df = epi.d[rfirst]; xf = x[rfirst]; }
    }
    mbar_wait(&full[stage], parity);
    // Rows are consumed straight out of the staged (col,val) stream: G lanes per row multiply-add in
    // registers, so products never make a round trip through shared memory (the L1 data pipe was the
    // limiter: profiles/README.md r1_c).  With G = 1 consecutive lanes own consecutive rows, which makes
    // the x gathers of a banded stencil fully coalesced (lane l reads x[col_p(row0 + l)]), and the row
    // sum is taken in storage order like the reference's loop (csr_matvec.c:210-216).
    for (int base = r0; base < r1; base += NT / G) {
      const int r = base + sub;
      double s = 0.0;
      if (r < r1) {
        int s0, s1;
        if (rp_smem) { s0 = rp[r - ra]; s1 = rp[r - ra + 1]; } else { s0 = A_i[r]; s1 = A_i[r + 1]; }
        s0 -= a0; s1 -= a0;
        int p = s0 + lane;
        for (; p + 3 * G < s1; p += 4 * G) {          // four gathers in flight per lane
          const int c0 = pc[p], c1 = pc[p + G], c2 = pc[p + 2 * G], c3 = pc[p + 3 * G];
          const double v0 = pv[p], v1 = pv[p + G], v2 = pv[p + 2 * G], v3 = pv[p + 3 * G];
          const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
          s += v0 * x0; s += v1 * x1; s += v2 * x2; s += v3 * x3;
        }
        if (p < s1) {                                  // tail of up to three entries, also issued together
          const bool k1 = p + G < s1, k2 = p + 2 * G < s1;
          const int c0 = pc[p], c1 = k1 ? pc[p + G] : c0, c2 = k2 ? pc[p + 2 * G] : c0;
          const double v0 = pv[p], v1 = k1 ? pv[p + G] : 0.0, v2 = k2 ? pv[p + 2 * G] : 0.0;
          const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2);
          s += v0 * x0;
          if (k1) s += v1 * x1;
          if (k2) s += v2 * x2;
        }
      }
      if (G > 1) {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off, G);
      }
      if (r < r1 && lane == 0) {
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
    __syncthreads();                 // every thread is done with this stage
    if (tid == 0 && k + NSTAGE < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes before the async refill
      issue(k + NSTAGE, m_refill);
    }
    m = m_next;
  }
}

template <int G, int NSTAGE>
int launch_pipe2(b200_handle h, b200_csr A, const double *x, double *y, const Epi &epi) {
  int scap = (A->tile + A->max_row + 8 + 31) & ~31;
  if (scap > SCAP_MAX) scap = SCAP_MAX;
  const size_t bytes = ((sizeof(double) + sizeof(int)) * scap + sizeof(int) * RCAP) * NSTAGE + sizeof(unsigned long long) * NSTAGE;
  // the >48 KB opt-in is a per-device attribute of the function: one bit per device ordinal, set atomically (rank threads
  // of the threads-as-ranks backend and processes that open several devices both come through here)
  static std::atomic<unsigned long long> attr_set{0};
  const unsigned long long bit = 1ull << (h->device & 63);
  if (!(attr_set.load(std::memory_order_acquire) & bit)) {
    const size_t maxb = ((sizeof(double) + sizeof(int)) * SCAP_MAX + sizeof(int) * RCAP) * NSTAGE + 64;
    B200_CUDA(cudaFuncSetAttribute(spmv_pipe_kernel<G, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxb));
    attr_set.fetch_or(bit, std::memory_order_release);
  }
  // Operators with an irregular x gather (G > 1: the coarse Galerkin operators, R) are bound by the L2 -> SM sector traffic of
  // the gather, not by HBM: with the default carve-out (228 KB shared, 28 KB L1) every gathered sector comes from L2.  Asking
  // for the 196 KB configuration leaves 60 KB of L1 for x at the price of one resident CTA (measured on the 256^3 hierarchy:
  // A_2 229 -> 171 us, A_1 510 -> 496 us; profiles/README.md r2).  B200_SPMV_CARVEOUT=<percent> overrides (0 = default).
  static const int carve_env = [] { const char *e = getenv("B200_SPMV_CARVEOUT"); return e ? atoi(e) : 75; }();
  size_t smem_budget = (size_t)(227 * 1024);
  if (carve_env > 0 && G > 1) {
    static std::atomic<unsigned long long> carve_set{0};
    if (!(carve_set.load(std::memory_order_acquire) & bit)) {
      B200_CUDA(cudaFuncSetAttribute(spmv_pipe_kernel<G, NSTAGE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve_env));
      carve_set.fetch_or(bit, std::memory_order_release);
    }
    // supported shared-memory configurations (KB per SM); the driver rounds the preference up to the next one
    static const int cfg[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};
    const int want = 256 * carve_env / 100;
    int kb = 228;
    for (int q : cfg) if (q >= want) { kb = q; break; }
    smem_budget = (size_t)(kb - 1) * 1024;
  }
  int per_sm = (int)(smem_budget / (bytes + 1024));   // +1 KB: per-CTA reservation of the runtime
  if (per_sm > 2048 / NT) per_sm = 2048 / NT;
  {
    // persistent grid = exactly the CTAs that are resident at once: registers (40 per thread: 12 CTAs) bind before shared
    // memory for the short-row operators (P, R), and a grid larger than one wave leaves a second, under-filled wave
    static std::atomic<int> occ_cache[64];                // by stage size / 32 entries (scap is a multiple of 32, <= 1536)
    std::atomic<int> &slot = occ_cache[(scap / 32) & 63];
    int occ = slot.load(std::memory_order_relaxed);
    if (occ == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_pipe_kernel<G, NSTAGE>, NT, bytes) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = per_sm;
      }
      slot.store(occ, std::memory_order_relaxed);
    }
    if (per_sm > occ) per_sm = occ;
  }
  if (per_sm < 1) per_sm = 1;
  int grid = h->num_sm * per_sm;
  if (grid > A->nblk) grid = A->nblk;
  spmv_pipe_kernel<G, NSTAGE><<<grid, NT, bytes, h->stream>>>(A->i, A->j, A->a, reinterpret_cast<const int4 *>(A->blk_meta),
                                                              A->nblk, scap, x, y, epi);
  B200_LAUNCH_CHECK();
  return 0;
}
template <int G>
int launch_pipe(b200_handle h, b200_csr A, const double *x, double *y, const Epi &epi) {
  static const int stages = [] { const char *e = getenv("B200_SPMV_STAGES"); return e ? atoi(e) : 2; }();
  if (stages == 3) return launch_pipe2<G, 3>(h, A, x, y, epi);
  if (stages == 4) return launch_pipe2<G, 4>(h, A, x, y, epi);
  return launch_pipe2<G, 2>(h, A, x, y, epi);
}

}  // namespace

// returns 1 if this matrix can go through the pipelined kernel
bool b200_spmv_pipe_ok(b200_csr A) {
  return A->owns && A->blk_meta != nullptr && A->max_row > 0 && A->tile + A->max_row + 8 <= SCAP_MAX;
}

int b200_csr_spmv_pipe(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha, double beta,
                       const double *b, const double *d) {
  Epi e{mode, alpha, beta, b, d};
  switch (A->group) {
    case 1:  return launch_pipe<1>(h, A, x, y, e);
    case 2:  return launch_pipe<2>(h, A, x, y, e);
    case 4:  return launch_pipe<4>(h, A, x, y, e);
    case 8:  return launch_pipe<8>(h, A, x, y, e);
    case 16: return launch_pipe<16>(h, A, x, y, e);
    default: return launch_pipe<32>(h, A, x, y, e);
  }
}
