// b200_setup_warp.cu -- warp-per-row, shared-memory versions of the two heavy setup stages:
// ext+i interpolation (+ truncation) and the Gustavson SpGEMM.   Compiled with -fmad=false.
//
// Why these stay bit-identical to the reference while 32 lanes cooperate on a row:
//   * the reference's outer loops (over the entries of row i of S / A) are kept sequential;
//   * its inner loops run over the entries of ONE other row (S_{i1}, A_{i1} or B_{ja}), whose
//     column indices are distinct, so the 32 lanes of a chunk touch distinct hash keys and distinct
//     accumulators: no two additions to the same target are reordered;
//   * first-touch positions inside a chunk are handed out by ballot + prefix popcount, i.e. in
//     entry order, exactly the order the sequential loop would have assigned;
//   * the one true reduction (the `sum` of par_lr_interp.c:1673-1693) is accumulated in lane
//     order with 32 shuffles, adding 0.0 for non-qualifying entries (an exact no-op).
// Each warp owns a private hash table (column -> slot) and the row under construction in shared
// memory; a row that does not fit raises a device flag and the caller redoes the stage with the
// general HBM-scratch kernels of b200_setup.cu.
#include "b200_internal.h"
#include <map>
#include <mutex>

namespace {

constexpr int WPB = 4;                 // warps per CTA
// extpi_warp_kernel: 56 registers unconstrained = 9 CTAs per SM; capped at 48 (10 CTAs, 20 bytes of spills) the coarse-level
// interpolation is 1.3 ms faster at 256^3, at 40 (12 CTAs) the spills cost more than the occupancy brings (+3.5 ms)
#define EXTPI_MIN_CTAS 10
constexpr int NOTFOUND = -1;
constexpr int STRONG_F = -2;
constexpr int SELF = -3;
constexpr int PENDING = -1;            // row not produced yet (per-row overflow protocol)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned hslot(int key, int cap) { return (((unsigned)key * 0x9E3779B1u) >> 9) & (unsigned)(cap - 1); }

template <int CAP>
__device__ __forceinline__ int wt_find(const int *keys, const int *vals, int key) {
  unsigned s = hslot(key, CAP);
  while (true) {
    int kk = keys[s];
    if (kk == key) return vals[s];
    if (kk == -1) return NOTFOUND;
    s = (s + 1) & (CAP - 1);
  }
}
template <int CAP>
__device__ __forceinline__ int wt_insert(int *keys, int key, bool *is_new) {
  unsigned s = hslot(key, CAP);
  while (true) {
    int prev = atomicCAS(&keys[s], -1, key);
    if (prev == -1) { *is_new = true; return (int)s; }
    if (prev == key) { *is_new = false; return (int)s; }
    s = (s + 1) & (CAP - 1);
  }
}

// lane-0 replay of hypre_qsort2_abs (utilities/hypre_qsort.c:367-387) on shared-memory arrays
__device__ void qsort2_abs_smem(int *v, double *w, int left, int right) {
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
      int l1 = left, r1 = last - 1, l2 = last + 1, r2 = right;
      if (r1 - l1 > r2 - l2) {
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

// ------------------------------------------------------------------------------------------
// ext+i interpolation + truncation, one warp per fine row, output rows of <= pmax entries
// written with stride pmax (par_lr_interp.c:1301-1416, :1523-1803; par_csr_matrix.c:2768-3020)
// ------------------------------------------------------------------------------------------
// Latency structure: everything that depends only on the entries of row i (neighbour ids, their CF
// markers, row extents, diagonal signs, hash lookups) is fetched by the 32 lanes at once, one lane per
// entry, and handed to the sequential outer loop by shuffles; the first 32-entry chunk of the NEXT strong
// F neighbour's row is already in flight (software pipeline) while the current one is folded in.  The
// arithmetic and its order are unchanged.
template <int CAP>
__global__ void __launch_bounds__(32 * WPB, EXTPI_MIN_CTAS)
extpi_warp_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                  const int *__restrict__ S_i, const int *__restrict__ S_j, const int *__restrict__ cf,
                  const int *__restrict__ f2c, double trunc_tol, int pmax, int *__restrict__ out_j,
                  double *__restrict__ out_a, int *__restrict__ out_cnt, int *__restrict__ overflow,
                  const int *__restrict__ rows) {
  constexpr int LIMIT = CAP / 2;      // max occupied slots (C-hat members + strong-F markers)
  extern __shared__ unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-warp layout: ra[LIMIT] doubles | sumbuf[32] doubles | keys[CAP] | vals[CAP] | rj[LIMIT]
  unsigned char *base_p = smem_raw + (size_t)warp * (sizeof(double) * (LIMIT + 32) + sizeof(int) * (2 * CAP + LIMIT));
  double *ra = reinterpret_cast<double *>(base_p);
  double *sumbuf = ra + LIMIT;
  int *keys = reinterpret_cast<int *>(base_p + sizeof(double) * (LIMIT + 32));
  int *vals = keys + CAP;
  int *rj = vals + CAP;
  const unsigned ltmask = (1u << lane) - 1u;
  const int nwarps = gridDim.x * WPB;
  for (int idx = blockIdx.x * WPB + warp; idx < n; idx += nwarps) {
    const int i = rows ? rows[idx] : idx;         // second pass: list of rows still pending
    if (out_cnt[i] != PENDING) continue;          // finished by an earlier (smaller-table) pass
    const int c = cf[i];
    if (c >= 0) {
      if (lane == 0) { out_j[(size_t)i * pmax] = f2c[i]; out_a[(size_t)i * pmax] = 1.0; out_cnt[i] = 1; }
      continue;
    }
    if (c == -3) { if (lane == 0) out_cnt[i] = 0; continue; }
    for (int s = lane; s < CAP; s += 32) keys[s] = -1;
    __syncwarp();
    int count = 0, used = 0;
    bool over = false;
    // ---- discovery of C-hat_i in the reference's order -----------------------------------
    const int sb = S_i[i], se = S_i[i + 1];
    for (int base = sb; base < se && !over; base += 32) {
      const int nv = min(32, se - base);
      int i1_l = -1, c1_l = 0, b1_l = 0, e1_l = 0;
      if (lane < nv) {
        i1_l = S_j[base + lane];
        c1_l = cf[i1_l];
        if (c1_l < 0 && c1_l != -3) { b1_l = S_i[i1_l]; e1_l = S_i[i1_l + 1]; }
      }
      const unsigned fmask = __ballot_sync(FULL, lane < nv && c1_l < 0 && c1_l != -3);
      // two-stage pipeline over the strong F neighbours: stage A holds k1 = S_j[..] of neighbour tA,
      // stage B holds k1 and cf[k1] of neighbour tB (the next one to be consumed)
      int tB = fmask ? __ffs(fmask) - 1 : -1, tA = -1, kA = -1, kB = -1, cB = -1;
      if (tB >= 0) {
        const int bb = __shfl_sync(FULL, b1_l, tB), ee = __shfl_sync(FULL, e1_l, tB);
        if (bb + lane < ee) kB = S_j[bb + lane];
        const unsigned m = (tB >= 31) ? 0u : (fmask & ~((2u << tB) - 1u));
        tA = m ? __ffs(m) - 1 : -1;
        if (tA >= 0) {
          const int b2 = __shfl_sync(FULL, b1_l, tA), e2 = __shfl_sync(FULL, e1_l, tA);
          if (b2 + lane < e2) kA = S_j[b2 + lane];
        }
        if (kB >= 0) cB = cf[kB];
      }
      for (int t = 0; t < nv && !over; t++) {
        const int i1 = __shfl_sync(FULL, i1_l, t);
        const int c1 = __shfl_sync(FULL, c1_l, t);
        if (c1 >= 0) {
          if (used + 1 > LIMIT) { over = true; break; }
          int isn = 0;
          if (lane == 0) {
            bool nw; int slot = wt_insert<CAP>(keys, i1, &nw);
            if (nw) { vals[slot] = count; rj[count] = f2c[i1]; ra[count] = 0.0; isn = 1; }
          }
          isn = __shfl_sync(FULL, isn, 0);
          count += isn; used += isn;
          __syncwarp();
        } else if (c1 != -3) {
          const int b1 = __shfl_sync(FULL, b1_l, t), e1 = __shfl_sync(FULL, e1_l, t);
          const int cur_k = kB, cur_c = cB;                 // first chunk of this neighbour (t == tB)
          // advance the pipeline before touching the hash table: B <- A, A <- following neighbour
          tB = tA; kB = kA; cB = -1; tA = -1; kA = -1;
          if (tB >= 0) {
            if (kB >= 0) cB = cf[kB];
            const unsigned m = (tB >= 31) ? 0u : (fmask & ~((2u << tB) - 1u));
            tA = m ? __ffs(m) - 1 : -1;
            if (tA >= 0) {
              const int b2 = __shfl_sync(FULL, b1_l, tA), e2 = __shfl_sync(FULL, e1_l, tA);
              if (b2 + lane < e2) kA = S_j[b2 + lane];
            }
          }
          if (used + 1 > LIMIT) { over = true; break; }
          int isn = 0;
          if (lane == 0) {
            bool nw; int slot = wt_insert<CAP>(keys, i1, &nw);
            if (nw) { vals[slot] = STRONG_F; isn = 1; }
          }
          isn = __shfl_sync(FULL, isn, 0);
          used += isn;
          __syncwarp();
          for (int kk0 = b1; kk0 < e1; kk0 += 32) {
            if (used + min(32, e1 - kk0) > LIMIT) { over = true; break; }
            int k1 = -1;
            bool want = false;
            if (kk0 == b1) { k1 = cur_k; want = (k1 >= 0) && (cur_c >= 0); }
            else {
              const int kk = kk0 + lane;
              if (kk < e1) { k1 = S_j[kk]; want = cf[k1] >= 0; }
            }
            bool nw = false;
            int slot = 0;
            if (want) slot = wt_insert<CAP>(keys, k1, &nw);
            const unsigned newmask = __ballot_sync(FULL, nw);
            if (nw) {
              const int pos = count + __popc(newmask & ltmask);
              vals[slot] = pos; rj[pos] = f2c[k1]; ra[pos] = 0.0;
            }
            const int nn = __popc(newmask);
            count += nn; used += nn;
            __syncwarp();
          }
        }
      }
    }
    if (over) {                                   // stays PENDING for the next, larger pass
      if (lane == 0) atomicExch(overflow, 1);
      __syncwarp();
      continue;
    }
    if (count == 0) { if (lane == 0) out_cnt[i] = 0; __syncwarp(); continue; }
    // ---- weights, reference accumulation order ---------------------------------------------
    double diagonal = A_a[A_i[i]];
    const int eA = A_i[i + 1];
    for (int base = A_i[i] + 1; base < eA; base += 32) {
      const int nv = min(32, eA - base);
      int i1_l = -1, m1_l = NOTFOUND, cf1_l = 0, b1_l = 0, e1_l = 0, sg_l = 1;
      double aij_l = 0.0;
      if (lane < nv) {
        i1_l = A_j[base + lane];
        aij_l = A_a[base + lane];
        m1_l = wt_find<CAP>(keys, vals, i1_l);
        if (m1_l == STRONG_F) {
          const int r0 = A_i[i1_l];
          b1_l = r0 + 1; e1_l = A_i[i1_l + 1];
          sg_l = (A_a[r0] < 0) ? -1 : 1;
        } else if (m1_l == NOTFOUND) {
          cf1_l = cf[i1_l];
        }
      }
      const unsigned fmask = __ballot_sync(FULL, lane < nv && m1_l == STRONG_F);
      // prefetch the first chunk of the first strong F neighbour's row
      int tP = fmask ? __ffs(fmask) - 1 : -1, pI = -1;
      double pA = 0.0;
      if (tP >= 0) {
        const int bb = __shfl_sync(FULL, b1_l, tP), ee = __shfl_sync(FULL, e1_l, tP);
        if (bb + lane < ee) { pI = A_j[bb + lane]; pA = A_a[bb + lane]; }
      }
      for (int t = 0; t < nv; t++) {
        const int m1 = __shfl_sync(FULL, m1_l, t);
        const double aij = __shfl_sync(FULL, aij_l, t);
        if (m1 >= 0) {
          if (lane == 0) ra[m1] += aij;
        } else if (m1 == STRONG_F) {
          const int b1 = __shfl_sync(FULL, b1_l, t), e1 = __shfl_sync(FULL, e1_l, t);
          const int sgn = __shfl_sync(FULL, sg_l, t);
          const int cI = pI;
          const double cA = pA;
          {                                                   // next neighbour's first chunk goes in flight now
            const unsigned m = (t >= 31) ? 0u : (fmask & ~((2u << t) - 1u));
            tP = m ? __ffs(m) - 1 : -1;
            pI = -1; pA = 0.0;
            if (tP >= 0) {
              const int bb = __shfl_sync(FULL, b1_l, tP), ee = __shfl_sync(FULL, e1_l, tP);
              if (bb + lane < ee) { pI = A_j[bb + lane]; pA = A_a[bb + lane]; }
            }
          }
          double sum = 0.0;
          int q0 = NOTFOUND;
          for (int k0 = b1; k0 < e1; k0 += 32) {
            const int k = k0 + lane;
            double contrib = 0.0, a = 0.0;
            int i2 = -1;
            const bool have = k < e1;
            if (k0 == b1) { i2 = cI; a = cA; }
            else if (have) { a = A_a[k]; i2 = A_j[k]; }
            int q = NOTFOUND;
            if (have && (sgn * a) < 0) {
              q = (i2 == i) ? SELF : wt_find<CAP>(keys, vals, i2);
              if (q == SELF || q >= 0) contrib = a;
            }
            if (k0 == b1) q0 = q;
            // lane order == entry order; lanes with contrib == 0 would add an exact 0.0, so only the contributing
            // values are summed: compacted into the warp's scratch in lane order, then added by every lane from
            // broadcast reads (2 instructions per term instead of a find-first-set + two shuffles + bookkeeping)
            const unsigned cmask = __ballot_sync(FULL, contrib != 0.0);
            if (cmask) {
              if (contrib != 0.0) sumbuf[__popc(cmask & ltmask)] = contrib;
              __syncwarp();
              const int nc = __popc(cmask);
              for (int j = 0; j < nc; j++) sum += sumbuf[j];
              __syncwarp();
            }
          }
          if (sum != 0) {
            const double distribute = aij / sum;
            for (int k0 = b1; k0 < e1; k0 += 32) {
              const int k = k0 + lane;
              int q = NOTFOUND;
              double a = 0.0;
              if (k0 == b1) { q = q0; a = cA; }
              else if (k < e1) {
                a = A_a[k];
                if ((sgn * a) < 0) {
                  const int i2 = A_j[k];
                  q = (i2 == i) ? SELF : wt_find<CAP>(keys, vals, i2);
                }
              }
              if (q >= 0) ra[q] += distribute * a;            // distinct q across the lanes of one row
              const unsigned selfmask = __ballot_sync(FULL, q == SELF);
              if (selfmask) {
                const double d = __shfl_sync(FULL, distribute * a, __ffs(selfmask) - 1);
                diagonal += d;
              }
            }
          } else {
            diagonal += aij;
          }
        } else if (__shfl_sync(FULL, cf1_l, t) != -3) {
          diagonal += aij;
        }
        __syncwarp();
      }
    }
    if (diagonal) {
      for (int p = lane; p < count; p += 32) ra[p] /= -diagonal;
    }
    __syncwarp();
    // ---- truncation ------------------------------------------------------------------------
    // Fast path (no drop tolerance, at most 64 candidates): when the pmax largest |w| are pairwise distinct and distinct from
    // every other entry, hypre_qsort2_abs (descending |w|) leaves exactly those at the front, in descending order, whatever it
    // does with ties further down -- pmax warp-wide arg-max rounds instead of a one-lane quicksort.  Any tie that could touch
    // the result (or a NaN) falls through to the exact replay below.
    if (trunc_tol <= 0 && count > pmax && count <= 64 && pmax <= 32) {
      const bool h0 = lane < count, h1 = lane + 32 < count;
      const double v0 = h0 ? ra[lane] : 0.0, v1 = h1 ? ra[lane + 32] : 0.0;
      const int k0 = h0 ? rj[lane] : 0, k1 = h1 ? rj[lane + 32] : 0;
      double a0 = h0 ? fabs(v0) : -1.0, a1 = h1 ? fabs(v1) : -1.0;
      bool ok = !__any_sync(FULL, (h0 && v0 != v0) || (h1 && v1 != v1));
      double myv = 0.0;
      int myk = 0;
      for (int r = 0; r < pmax && ok; r++) {
        double m = fmax(a0, a1);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, off));
        const unsigned b0 = __ballot_sync(FULL, a0 == m), b1 = __ballot_sync(FULL, a1 == m);
        if (__popc(b0) + __popc(b1) != 1) { ok = false; break; }
        const int src = b0 ? __ffs(b0) - 1 : __ffs(b1) - 1;
        const double sv = __shfl_sync(FULL, b0 ? v0 : v1, src);
        const int sk = __shfl_sync(FULL, b0 ? k0 : k1, src);
        if (lane == r) { myv = sv; myk = sk; }
        if (lane == src) { if (b0) a0 = -1.0; else a1 = -1.0; }
      }
      if (ok) {
        double row_sum = 0;
        if (lane == 0) for (int j = 0; j < count; j++) row_sum += ra[j];      // storage order (par_csr_matrix.c:2981-2990)
        row_sum = __shfl_sync(FULL, row_sum, 0);
        double scale = 0;
        for (int r = 0; r < pmax; r++) scale += __shfl_sync(FULL, myv, r);    // kept entries in their sorted order
        if (scale != 0. && scale != row_sum) myv *= row_sum / scale;
        if (lane < pmax) { out_j[(size_t)i * pmax + lane] = myk; out_a[(size_t)i * pmax + lane] = myv; }
        if (lane == 0) out_cnt[i] = pmax;
        __syncwarp();
        continue;
      }
    }
    // exact replay: lane 0 runs the sequential algorithm on the shared row
    int len = count;
    if (lane == 0) {
      if (trunc_tol > 0) {
        double row_nrm = 0;
        for (int j = 0; j < len; j++) row_nrm = (row_nrm < fabs(ra[j])) ? fabs(ra[j]) : row_nrm;
        const double drop = trunc_tol * row_nrm;
        double row_sum = 0, scale = 0;
        int keep = 0;
        for (int j = 0; j < len; j++) {
          row_sum += ra[j];
          if (!(fabs(ra[j]) < drop)) { scale += ra[j]; ra[keep] = ra[j]; rj[keep] = rj[j]; keep++; }
        }
        len = keep;
        if (scale != 0. && scale != row_sum) {
          scale = row_sum / scale;
          for (int j = 0; j < len; j++) ra[j] *= scale;
        }
      }
      if (len > pmax) {
        double row_sum = 0;
        for (int j = 0; j < len; j++) row_sum += ra[j];
        qsort2_abs_smem(rj, ra, 0, len - 1);
        double scale = 0;
        for (int j = 0; j < pmax; j++) scale += ra[j];
        len = pmax;
        if (scale != 0. && scale != row_sum) {
          scale = row_sum / scale;
          for (int j = 0; j < len; j++) ra[j] *= scale;
        }
      }
      out_cnt[i] = len;
    }
    len = __shfl_sync(FULL, len, 0);
    __syncwarp();
    for (int p = lane; p < len; p += 32) { out_j[(size_t)i * pmax + p] = rj[p]; out_a[(size_t)i * pmax + p] = ra[p]; }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// ext+i interpolation + truncation for STENCIL-SIZED rows (the finest level): one thread per fine row, the
// reference's own sequential loops (par_lr_interp.c:1301-1416, :1523-1803) on thread-private arrays -- C-hat_i
// (first-touch order) and the strong F neighbours are short lists searched linearly, so no hash table, no HBM
// scratch and no separate truncation pass.  A row that outgrows the lists raises `overflow` and the caller
// falls back to the general kernels.
// ------------------------------------------------------------------------------------------
template <int CH, int SF>
__global__ void __launch_bounds__(128)
extpi_thread_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                    const int *__restrict__ S_i, const int *__restrict__ S_j, const int *__restrict__ cf,
                    const int *__restrict__ f2c, double trunc_tol, int pmax, int *__restrict__ out_j,
                    double *__restrict__ out_a, int *__restrict__ out_cnt, int *__restrict__ overflow,
                    const int *__restrict__ rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = rows ? rows[t] : t;              // list of the F rows: every lane of a warp has work
  const int c = cf[i];
  if (c >= 0) { out_j[(size_t)i * pmax] = f2c[i]; out_a[(size_t)i * pmax] = 1.0; out_cnt[i] = 1; return; }
  if (c == -3) { out_cnt[i] = 0; return; }
  int key[CH], sf[SF];
  double val[CH];
  int nkey = 0, nsf = 0;
  bool over = false;
  auto find = [&](int k) { for (int p = 0; p < nkey; p++) if (key[p] == k) return p; return -1; };
  auto touch = [&](int k) {
    if (find(k) >= 0) return;
    if (nkey == CH) { over = true; return; }
    key[nkey++] = k;
  };
  for (int jj = S_i[i]; jj < S_i[i + 1] && !over; jj++) {
    const int i1 = S_j[jj];
    const int c1 = cf[i1];
    if (c1 >= 0) touch(i1);
    else if (c1 != -3) {
      if (nsf == SF) { over = true; break; }
      sf[nsf++] = i1;
      for (int kk = S_i[i1]; kk < S_i[i1 + 1] && !over; kk++) {
        const int k1 = S_j[kk];
        if (cf[k1] >= 0) touch(k1);
      }
    }
  }
  if (over) { atomicExch(overflow, 1); return; }          // stays PENDING: the warp-per-row passes take it
  if (nkey == 0) { out_cnt[i] = 0; return; }
  for (int p = 0; p < nkey; p++) val[p] = 0.0;
  double diagonal = A_a[A_i[i]];
  for (int jj = A_i[i] + 1; jj < A_i[i + 1]; jj++) {
    const int i1 = A_j[jj];
    const double aij = A_a[jj];
    const int m1 = find(i1);
    if (m1 >= 0) { val[m1] += aij; continue; }
    bool strong_f = false;
    for (int q = 0; q < nsf; q++) strong_f |= (sf[q] == i1);
    if (strong_f) {
      const int b1 = A_i[i1], e1 = A_i[i1 + 1];
      const int sgn = (A_a[b1] < 0) ? -1 : 1;
      double sum = 0.0;
      for (int k = b1 + 1; k < e1; k++) {
        const double a = A_a[k];
        if ((sgn * a) < 0) { const int i2 = A_j[k]; if (i2 == i || find(i2) >= 0) sum += a; }
      }
      if (sum != 0) {
        const double distribute = aij / sum;
        for (int k = b1 + 1; k < e1; k++) {
          const double a = A_a[k];
          if ((sgn * a) < 0) {
            const int i2 = A_j[k];
            const int m2 = find(i2);
            if (m2 >= 0) val[m2] += distribute * a;
            if (i2 == i) diagonal += distribute * a;
          }
        }
      } else {
        diagonal += aij;
      }
    } else if (cf[i1] != -3) {
      diagonal += aij;
    }
  }
  if (diagonal) for (int p = 0; p < nkey; p++) val[p] /= -diagonal;
  for (int p = 0; p < nkey; p++) key[p] = f2c[key[p]];
  // truncation (par_csr_matrix.c:2768-3020)
  int len = nkey;
  if (trunc_tol > 0) {
    double row_nrm = 0;
    for (int j = 0; j < len; j++) row_nrm = (row_nrm < fabs(val[j])) ? fabs(val[j]) : row_nrm;
    const double drop = trunc_tol * row_nrm;
    double row_sum = 0, scale = 0;
    int keep = 0;
    for (int j = 0; j < len; j++) {
      row_sum += val[j];
      if (!(fabs(val[j]) < drop)) { scale += val[j]; val[keep] = val[j]; key[keep] = key[j]; keep++; }
    }
    len = keep;
    if (scale != 0. && scale != row_sum) {
      scale = row_sum / scale;
      for (int j = 0; j < len; j++) val[j] *= scale;
    }
  }
  if (len > pmax) {
    double row_sum = 0;
    for (int j = 0; j < len; j++) row_sum += val[j];
    qsort2_abs_smem(key, val, 0, len - 1);
    double scale = 0;
    for (int j = 0; j < pmax; j++) scale += val[j];
    len = pmax;
    if (scale != 0. && scale != row_sum) {
      scale = row_sum / scale;
      for (int j = 0; j < len; j++) val[j] *= scale;
    }
  }
  out_cnt[i] = len;
  for (int p = 0; p < len; p++) { out_j[(size_t)i * pmax + p] = key[p]; out_a[(size_t)i * pmax + p] = val[p]; }
}

// rows that need no work: C points interpolate from themselves, isolated F points (-3) get an empty row; F rows stay PENDING
__global__ void extpi_trivial_rows_kernel(int n, const int *__restrict__ cf, const int *__restrict__ f2c, int pmax,
                                          int *__restrict__ out_j, double *__restrict__ out_a, int *__restrict__ out_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cf[i];
  if (c >= 0) { out_j[(size_t)i * pmax] = f2c[i]; out_a[(size_t)i * pmax] = 1.0; out_cnt[i] = 1; }
  else if (c == -3) out_cnt[i] = 0;
}

__global__ void strided_to_csr_kernel(int n, int stride, const int *__restrict__ P_i, const int *__restrict__ sj,
                                      const double *__restrict__ sa, int *__restrict__ P_j, double *__restrict__ P_a) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int d = P_i[r], len = P_i[r + 1] - d;
  for (int k = 0; k < len; k++) { P_j[d + k] = sj[(size_t)r * stride + k]; P_a[d + k] = sa[(size_t)r * stride + k]; }
}

// ------------------------------------------------------------------------------------------
// SpGEMM, one warp per row of C (csr_matop.c:375-468).
// GB lanes serve one entry a_{ic,ja} and walk row ja of B; 32/GB entries of A are in flight per
// step.  GB < 32 is used only when every row of B has at most GB entries, so a step holds whole
// B rows and the products of a step are in the reference's (ia, ib) order across lanes.
// Products of one step that hit the same column are found with __match_any_sync and applied in
// lane order; new columns take their first-touch positions from a ballot prefix.
// ------------------------------------------------------------------------------------------
// MODE 0: symbolic (row counts only);  MODE 1: numeric into an exactly sized CSR (counts known);
// MODE 2: ONE pass -- numeric into a scratch row of upper-bound length, count written afterwards; a row that
//         outgrows the table stays PENDING for the next, larger-table pass.
template <int CAP, int GB, int MODE>
__global__ void __launch_bounds__(32 * WPB)
spgemm_warp_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j, const double *__restrict__ A_a,
                   const int *__restrict__ B_i, const int *__restrict__ B_j, const double *__restrict__ B_a,
                   int allsquare, int diag_base, const int *__restrict__ rows, int *__restrict__ cnt,
                   const int *__restrict__ C_i, int *__restrict__ C_j, double *__restrict__ C_a,
                   int *__restrict__ overflow, int skip_ub = 0) {
  constexpr bool NUMERIC = MODE != 0;
  constexpr int LIMIT = CAP / 2;
  constexpr int APS = 32 / GB;                      // A entries per step
  extern __shared__ unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per_warp = NUMERIC ? (sizeof(double) * LIMIT + sizeof(int) * (2 * CAP + LIMIT)) : (sizeof(int) * CAP);
  unsigned char *base = smem_raw + (size_t)warp * per_warp;
  double *acc = reinterpret_cast<double *>(base);
  int *keys = NUMERIC ? reinterpret_cast<int *>(base + sizeof(double) * LIMIT) : reinterpret_cast<int *>(base);
  int *vals = keys + CAP;
  int *cols = vals + CAP;
  const unsigned ltmask = (1u << lane) - 1u;
  const int nwarps = gridDim.x * WPB;
  const int sub = lane % GB, grp_id = lane / GB;
  for (int idx = blockIdx.x * WPB + warp; idx < n; idx += nwarps) {
    const int ic = rows ? rows[idx] : idx;          // rows of one size class / rows still pending
    if (MODE != 1 && cnt[ic] != PENDING) continue;
    if (MODE == 2 && skip_ub > 0 && C_i[ic + 1] - C_i[ic] > skip_ub) {   // almost surely too long for this table: next pass
      if (lane == 0) atomicExch(overflow, 1);
      continue;
    }
    for (int s = lane; s < CAP; s += 32) keys[s] = -1;
    __syncwarp();
    int count = 0;
    bool over = false;
    if (allsquare) {                                // diagonal first (:384-388, :442-448)
      if (lane == 0) {
        bool nw; int slot = wt_insert<CAP>(keys, diag_base + ic, &nw);
        if (NUMERIC) { vals[slot] = 0; cols[0] = diag_base + ic; acc[0] = 0.0; }
      }
      count = 1;
      __syncwarp();
    }
    const int eA = A_i[ic + 1];
    // one step: up to 32 (entry of B, product) pairs, in the reference's (ia, ib) order across the lanes
    auto fold = [&](bool valid, int jb, double prod) {
      const unsigned vm = __ballot_sync(FULL, valid);
      if (vm == 0) return;
      if (MODE != 1 && count + __popc(vm) > LIMIT) { over = true; return; }   // MODE 1 tables are sized from the symbolic count
      unsigned grp = 0;
      if (valid) grp = __match_any_sync(vm, jb);
      const int leader = valid ? (__ffs(grp) - 1) : lane;
      const bool is_leader = valid && leader == lane;
      bool nw = false;
      int slot = 0;
      if (is_leader) slot = wt_insert<CAP>(keys, jb, &nw);
      const unsigned newmask = __ballot_sync(FULL, nw);
      if (NUMERIC) {
        if (nw) {
          const int pos = count + __popc(newmask & ltmask);
          vals[slot] = pos; cols[pos] = jb; acc[pos] = 0.0;      // 0 + a*b == a*b: same as the first-touch store
        }
        slot = __shfl_sync(FULL, slot, leader);
        __syncwarp();
        const int rank = __popc(grp & ltmask);
        int maxrank = valid ? __popc(grp) : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) maxrank = max(maxrank, __shfl_xor_sync(FULL, maxrank, off));
        for (int r = 0; r < maxrank; r++) {
          if (valid && rank == r) acc[vals[slot]] += prod;
          __syncwarp();
        }
      }
      count += __popc(newmask);
      __syncwarp();
    };
    if (GB == 32) {
      // rows of B longer than 16 entries: one A entry per step.  The 32 lanes fetch (ja, a, row extent of B)
      // for 32 entries of A at once; the first chunk of the NEXT entry's B row is in flight while the
      // current one is folded into the table.
      for (int base = A_i[ic]; base < eA && !over; base += 32) {
        const int nv = min(32, eA - base);
        int bB_l = 0, eB_l = 0;
        double a_l = 0.0;
        if (lane < nv) {
          const int ja = A_j[base + lane];
          bB_l = B_i[ja]; eB_l = B_i[ja + 1];
          if (NUMERIC) a_l = A_a[base + lane];
        }
        int pJ = -1;
        double pB = 0.0;
        {
          const int bb = __shfl_sync(FULL, bB_l, 0), ee = __shfl_sync(FULL, eB_l, 0);
          if (bb + lane < ee) { pJ = B_j[bb + lane]; if (NUMERIC) pB = B_a[bb + lane]; }
        }
        for (int t = 0; t < nv && !over; t++) {
          const int bB = __shfl_sync(FULL, bB_l, t), eB = __shfl_sync(FULL, eB_l, t);
          const double a = __shfl_sync(FULL, a_l, t);
          const int cJ = pJ;
          const double cB = pB;
          pJ = -1; pB = 0.0;
          if (t + 1 < nv) {
            const int bb = __shfl_sync(FULL, bB_l, t + 1), ee = __shfl_sync(FULL, eB_l, t + 1);
            if (bb + lane < ee) { pJ = B_j[bb + lane]; if (NUMERIC) pB = B_a[bb + lane]; }
          }
          for (int c0 = bB; c0 < eB && !over; c0 += 32) {
            const int ib = c0 + lane;
            const bool valid = ib < eB;
            int jb = -1;
            double prod = 0.0;
            if (c0 == bB) { jb = cJ; if (NUMERIC) prod = a * cB; }
            else if (valid) { jb = B_j[ib]; if (NUMERIC) prod = a * B_a[ib]; }
            fold(valid, jb, prod);
          }
        }
      }
    } else {
      for (int ia0 = A_i[ic]; ia0 < eA && !over; ia0 += APS) {
        const int ia = ia0 + grp_id;
        int bB = 0, eB = 0;
        double a = 0.0;
        if (ia < eA) { const int ja = A_j[ia]; bB = B_i[ja]; eB = B_i[ja + 1]; if (NUMERIC) a = A_a[ia]; }
        const int ib = bB + sub;                          // every row of B has at most GB entries
        const bool valid = ib < eB;
        int jb = -1;
        double prod = 0.0;
        if (valid) { jb = B_j[ib]; if (NUMERIC) prod = a * B_a[ib]; }
        fold(valid, jb, prod);
      }
    }
    if (over) {
      if (lane == 0) atomicExch(overflow, 1);       // symbolic: row stays PENDING; numeric never overflows (sized by cnt)
      __syncwarp();
      continue;
    }
    if (MODE == 0) {
      if (lane == 0) cnt[ic] = count;
    } else {
      const int start = C_i[ic];                    // MODE 2: offset of the row's upper-bound slot in the scratch
      for (int p = lane; p < count; p += 32) { C_j[start + p] = cols[p]; C_a[start + p] = acc[p]; }
      if (MODE == 2 && lane == 0) cnt[ic] = count;
    }
    __syncwarp();
  }
}

// upper bound of the length of row ic of A*B: the products it forms (+1 for the diagonal slot), at most `cap`
__global__ void spgemm_row_bound_kernel(int n, const int *__restrict__ A_i, const int *__restrict__ A_j,
                                        const int *__restrict__ B_i, int allsquare, int cap, int *__restrict__ ub) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { ub[n] = 0; return; }
  long long s = allsquare ? 1 : 0;
  for (int k = A_i[i]; k < A_i[i + 1]; k++) { const int ja = A_j[k]; s += B_i[ja + 1] - B_i[ja]; }
  ub[i] = (int)(s < cap ? s : cap);
}
// scratch rows (upper-bound offsets) -> CSR rows, one warp per row
__global__ void spgemm_compact_kernel(int n, const int *__restrict__ off, const int *__restrict__ sj, const double *__restrict__ sa,
                                      const int *__restrict__ C_i, int *__restrict__ C_j, double *__restrict__ C_a) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nwarps) {
    const int src = off[r], dst = C_i[r], len = C_i[r + 1] - dst;
    for (int p = lane; p < len; p += 32) { C_j[dst + p] = sj[src + p]; C_a[dst + p] = sa[src + p]; }
  }
}

__global__ void max_rowlen_kernel(int n, const int *__restrict__ A_i, int *__restrict__ out_max) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int len = i < n ? A_i[i + 1] - A_i[i] : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) len = max(len, __shfl_down_sync(FULL, len, off));
  if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out_max, len);
}
__global__ void fill_int_kernel(size_t n, int v, int *x) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = v;
}

// flag[i] = 1 when lo < cnt[i] <= hi  (pending rows: lo = hi = PENDING selects cnt == PENDING)
__global__ void class_flag_kernel(int n, const int *__restrict__ cnt, int lo, int hi, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { flag[n] = 0; return; }
  const int c = cnt[i];
  flag[i] = (lo == PENDING && hi == PENDING) ? (c == PENDING) : (c > lo && c <= hi);
}
__global__ void class_scatter_kernel(int n, const int *__restrict__ cnt, int lo, int hi, const int *__restrict__ pos,
                                     int *__restrict__ list) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cnt[i];
  const bool in = (lo == PENDING && hi == PENDING) ? (c == PENDING) : (c > lo && c <= hi);
  if (in) list[pos[i]] = i;
}

template <class K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return b200_set_error(__FILE__, __LINE__, cudaGetErrorString(e));
  }
  return 0;
}

inline int warp_grid(b200_handle h, int n, int blocks_per_sm) {
  long long need = ((long long)n + WPB - 1) / WPB;
  long long cap = (long long)h->num_sm * blocks_per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}
// Grid of a grid-stride, warp-per-row kernel = exactly the CTAs that are resident at once (occupancy of THIS kernel with THIS
// much dynamic shared memory, asked of the runtime and remembered).  A grid sized from a guessed blocks-per-SM constant that
// exceeds the real occupancy (registers: 9 for extpi_warp_kernel<256>, 12 for spgemm_warp_kernel<256,32,2>) runs as one full
// wave plus an under-filled one: ncu showed sm__warps_active 51 % on the level-0 (RA)P product.
template <class K>
int occ_grid(b200_handle h, K kernel, int n, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, size_t>, int> cache;
  int occ = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_pair((const void *)kernel, bytes * 64 + (size_t)h->device);
    auto it = cache.find(key);
    if (it != cache.end()) occ = it->second;
    else {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 32 * WPB, bytes) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 8; }
      cache[key] = occ;
    }
  }
  return warp_grid(h, n, occ);
}
inline int fill_grid(b200_handle h, size_t n) {
  size_t g = (n + 255) / 256, cap = (size_t)h->num_sm * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// order-preserving list of the rows whose count lies in (lo, hi] (or is PENDING)
int build_row_list(b200_handle h, int n, const int *cnt, int lo, int hi, int **list_out, int *count_out) {
  int *pos = nullptr;
  B200_TRY(b200_dalloc<int>(h, &pos, (size_t)n + 1));
  class_flag_kernel<<<b200_grid((size_t)n + 1, 256), 256, 0, h->stream>>>(n, cnt, lo, hi, pos);
  B200_LAUNCH_CHECK();
  B200_TRY(b200_exclusive_scan_inplace(h, pos, (size_t)n + 1));
  int m = 0;
  B200_CUDA(cudaMemcpyAsync(&m, pos + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  int *list = nullptr;
  if (m > 0) {
    B200_TRY(b200_dalloc<int>(h, &list, m));
    class_scatter_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, cnt, lo, hi, pos, list);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_dfree(h, pos));
  *list_out = list;
  *count_out = m;
  return 0;
}

}  // namespace

int b200_coarse_map(b200_handle h, int n, const int *d_cf, int **f2c_out, int *ncoarse);

// returns 0 and *done = 1 when the warp path produced P; *done = 0 -> caller must use the general path
int b200_extpi_interp_warp(b200_handle h, b200_csr A, b200_csr S, const int *d_cf, int n, const int *d_f2c_in, int ncoarse_in,
                           double trunc_factor, int max_elmts, b200_csr *out, int *done) {
  *done = 0;
  if (max_elmts <= 0) return 0;                       // unbounded rows: general path
  if (n == 0) return 0;
  // stencil-sized rows (the finest level): one thread per row on private lists (extpi_thread_kernel)
  static const double thread_avg = [] { const char *e = getenv("B200_EXTPI_THREAD_AVG"); return e ? atof(e) : 10.0; }();
  const double avg_row = (double)A->nnz / (A->nrows ? A->nrows : 1);
  const bool stencil_rows = avg_row <= thread_avg;
  static const bool no_thread_rows = [] { const char *e = getenv("B200_EXTPI_NO_THREAD_ROWS"); return e && e[0] == '1'; }();
  if (stencil_rows && no_thread_rows) return 0;
  int *d_flag = nullptr;
  B200_TRY(b200_dalloc<int>(h, &d_flag, 1));
  int *f2c = nullptr, ncoarse = ncoarse_in;
  if (d_f2c_in) f2c = const_cast<int *>(d_f2c_in);
  else B200_TRY(b200_coarse_map(h, n, d_cf, &f2c, &ncoarse));
  int *sj = nullptr, *cnt = nullptr;
  double *sa = nullptr;
  B200_TRY(b200_dalloc<int>(h, &sj, (size_t)n * max_elmts));
  B200_TRY(b200_dalloc<double>(h, &sa, (size_t)n * max_elmts));
  B200_TRY(b200_dalloc<int>(h, &cnt, (size_t)n + 1));
  fill_int_kernel<<<fill_grid(h, n), 256, 0, h->stream>>>((size_t)n, PENDING, cnt);
  B200_LAUNCH_CHECK();
  B200_CUDA(cudaMemsetAsync(cnt + n, 0, sizeof(int), h->stream));
  int flag = 1;
  if (stencil_rows) {
    // the dependent-load chain per row is short and thread-per-row keeps 32x more rows in flight than a warp per row
    B200_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
    // C rows and isolated rows are written by a streaming kernel; the row kernel then runs over the LIST of F rows, so that every
    // lane of its warps has a row to build (a third of the lanes would otherwise retire at once and idle through the warp's work)
    static const bool f_list = [] { const char *e = getenv("B200_EXTPI_F_LIST"); return !(e && e[0] == '0'); }();
    int *frows = nullptr, nf = n;
    if (f_list) {
      extpi_trivial_rows_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, d_cf, f2c, max_elmts, sj, sa, cnt);
      B200_LAUNCH_CHECK();
      B200_TRY(build_row_list(h, n, cnt, PENDING, PENDING, &frows, &nf));
    }
    if (nf > 0) {
      if (avg_row <= 10.0)
        extpi_thread_kernel<32, 16><<<b200_grid(nf, 128), 128, 0, h->stream>>>(nf, A->i, A->j, A->a, S->i, S->j, d_cf, f2c, trunc_factor,
                                                                             max_elmts, sj, sa, cnt, d_flag, frows);
      else
        extpi_thread_kernel<64, 32><<<b200_grid(nf, 128), 128, 0, h->stream>>>(nf, A->i, A->j, A->a, S->i, S->j, d_cf, f2c, trunc_factor,
                                                                             max_elmts, sj, sa, cnt, d_flag, frows);
      B200_LAUNCH_CHECK();
    }
    B200_TRY(b200_dfree(h, frows));
    B200_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  // rows still PENDING (all of them on the coarse levels; the few the private lists could not hold otherwise)
  for (int pass = 0; pass < 2 && flag; pass++) {
    B200_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
    int *rows = nullptr, m = n;
    if (pass == 1 || stencil_rows) B200_TRY(build_row_list(h, n, cnt, PENDING, PENDING, &rows, &m));
#define B200_EXTPI_LAUNCH(CAPV, BPS)                                                                              \
    {                                                                                                             \
      constexpr int CAP = CAPV;                                                                                   \
      const size_t bytes = (size_t)WPB * (sizeof(double) * (CAP / 2 + 32) + sizeof(int) * (2 * CAP + CAP / 2));   \
      B200_TRY(set_smem(extpi_warp_kernel<CAP>, bytes));                                                          \
      extpi_warp_kernel<CAP><<<occ_grid(h, extpi_warp_kernel<CAP>, m, bytes), 32 * WPB, bytes, h->stream>>>(                               \
          m, A->i, A->j, A->a, S->i, S->j, d_cf, f2c, trunc_factor, max_elmts, sj, sa, cnt, d_flag, rows);        \
    }
    if (m > 0) {
      if (pass == 0) B200_EXTPI_LAUNCH(256, 16)
      else B200_EXTPI_LAUNCH(1024, 3)
      B200_LAUNCH_CHECK();
    }
#undef B200_EXTPI_LAUNCH
    B200_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dfree(h, rows));
  }
  if (flag) {                                         // some row outgrew even the large shared table
    B200_TRY(b200_dfree(h, sj)); B200_TRY(b200_dfree(h, sa)); B200_TRY(b200_dfree(h, cnt));
    if (!d_f2c_in) B200_TRY(b200_dfree(h, f2c));
    B200_TRY(b200_dfree(h, d_flag));
    return 0;
  }
  B200_TRY(b200_exclusive_scan_inplace(h, cnt, (size_t)n + 1));
  int nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, cnt + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr P = nullptr;
  B200_TRY(b200_csr_alloc(h, n, ncoarse, nnz, true, &P));
  B200_CUDA(cudaMemcpyAsync(P->i, cnt, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
  strided_to_csr_kernel<<<b200_grid(n, 256), 256, 0, h->stream>>>(n, max_elmts, P->i, sj, sa, P->j, P->a);
  B200_LAUNCH_CHECK();
  B200_TRY(b200_dfree(h, sj)); B200_TRY(b200_dfree(h, sa)); B200_TRY(b200_dfree(h, cnt));
  if (!d_f2c_in) B200_TRY(b200_dfree(h, f2c));
  B200_TRY(b200_dfree(h, d_flag));
  *out = P;
  *done = 1;
  return 0;
}

// One-pass product: every row is formed once, in shared memory, and parked in a scratch slot of its upper-bound
// length; the exact CSR is compacted afterwards.  Saves the whole symbolic pass (a second walk over A and B with the
// same hashing) for one extra copy of C.  *done = 0 -> the caller uses the two-pass path.
template <int GB>
static int spgemm_fused_run(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C, b200_csr *out, int *done) {
  *done = 0;
  const int n = A->nrows;
  int *off = nullptr, *cnt = nullptr, *d_flag = nullptr;
  B200_TRY(b200_dalloc<int>(h, &off, (size_t)n + 1));
  spgemm_row_bound_kernel<<<b200_grid((size_t)n + 1, 256), 256, 0, h->stream>>>(n, A->i, A->j, B->i, allsquare, 1024, off);
  B200_LAUNCH_CHECK();
  long long total = 0;
  B200_TRY(b200_reduce_sum_int(h, off, (size_t)n, &total));
  if (total >= 0x7fffffffLL || total > 12LL * ((long long)A->nnz + B->nnz) + 64LL * n) {   // scratch would dwarf the operands
    B200_TRY(b200_dfree(h, off));
    return 0;
  }
  B200_TRY(b200_exclusive_scan_inplace(h, off, (size_t)n + 1));
  int *sj = nullptr;
  double *sa = nullptr;
  B200_TRY(b200_dalloc<int>(h, &sj, (size_t)total + 1));
  B200_TRY(b200_dalloc<double>(h, &sa, (size_t)total + 1));
  B200_TRY(b200_dalloc<int>(h, &cnt, (size_t)n + 1));
  B200_TRY(b200_dalloc<int>(h, &d_flag, 1));
  fill_int_kernel<<<fill_grid(h, n), 256, 0, h->stream>>>((size_t)n, PENDING, cnt);
  B200_LAUNCH_CHECK();
  B200_CUDA(cudaMemsetAsync(cnt + n, 0, sizeof(int), h->stream));
  int flag = 1;
  // Small products (the coarse end of the hierarchy: a few thousand long rows) start with the 1024-slot table: occupancy is
  // irrelevant there, and every size class that turns rows away costs a row list (scan + host round trip) and a flag round trip.
  const int pass0 = (n <= 32768) ? 2 : 0;
  for (int pass = pass0; pass < 4 && flag; pass++) {
    B200_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
    int *rows = nullptr, m = n;
    if (pass > pass0) B200_TRY(build_row_list(h, n, cnt, PENDING, PENDING, &rows, &m));
#define B200_SPGEMM_FUSED(CAPV, BPS)                                                                              \
    {                                                                                                             \
      constexpr int CAP = CAPV;                                                                                   \
      const size_t bytes = (size_t)WPB * (sizeof(double) * (CAP / 2) + sizeof(int) * (2 * CAP + CAP / 2));        \
      B200_TRY(set_smem(spgemm_warp_kernel<CAP, GB, 2>, bytes));                                                  \
      spgemm_warp_kernel<CAP, GB, 2><<<occ_grid(h, spgemm_warp_kernel<CAP, GB, 2>, m, bytes), 32 * WPB, bytes, h->stream>>>(                       \
          m, A->i, A->j, A->a, B->i, B->j, B->a, allsquare, diag_base, rows, cnt, off, sj, sa, d_flag,            \
          pass == 0 ? 512 : 0);                                                                                   \
    }
    if (m > 0) {
      if (pass == 0) B200_SPGEMM_FUSED(256, 15)           // rows of <= 128 entries
      else if (pass == 1) B200_SPGEMM_FUSED(512, 7)       // <= 256
      else if (pass == 2) B200_SPGEMM_FUSED(1024, 3)      // <= 512
      else B200_SPGEMM_FUSED(2048, 1)                     // <= 1024
      B200_LAUNCH_CHECK();
    }
#undef B200_SPGEMM_FUSED
    B200_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dfree(h, rows));
  }
  if (!flag) {
    B200_TRY(b200_exclusive_scan_inplace(h, cnt, (size_t)n + 1));
    int nnz = 0;
    B200_CUDA(cudaMemcpyAsync(&nnz, cnt + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    b200_csr C = nullptr;
    B200_TRY(b200_csr_alloc(h, n, ncols_C, nnz, true, &C));
    B200_CUDA(cudaMemcpyAsync(C->i, cnt, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
    spgemm_compact_kernel<<<warp_grid(h, n, 16), 32 * WPB, 0, h->stream>>>(n, off, sj, sa, C->i, C->j, C->a);
    B200_LAUNCH_CHECK();
    *out = C;
    *done = 1;
  }
  B200_TRY(b200_dfree(h, off)); B200_TRY(b200_dfree(h, cnt)); B200_TRY(b200_dfree(h, d_flag));
  B200_TRY(b200_dfree(h, sj)); B200_TRY(b200_dfree(h, sa));
  return 0;
}

template <int GB>
static int spgemm_warp_run(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C, b200_csr *out, int *done) {
  static const bool two_pass = [] { const char *e = getenv("B200_SPGEMM_TWO_PASS"); return e && e[0] == '1'; }();
  if (!two_pass && A->a && B->a) {
    B200_TRY(spgemm_fused_run<GB>(h, A, B, allsquare, diag_base, ncols_C, out, done));
    if (*done) return 0;
  }
  const int n = A->nrows;
  int *cnt = nullptr, *d_flag = nullptr;
  B200_TRY(b200_dalloc<int>(h, &cnt, (size_t)n + 1));
  B200_TRY(b200_dalloc<int>(h, &d_flag, 1));
  fill_int_kernel<<<fill_grid(h, n), 256, 0, h->stream>>>((size_t)n, PENDING, cnt);
  B200_LAUNCH_CHECK();
  B200_CUDA(cudaMemsetAsync(cnt + n, 0, sizeof(int), h->stream));
  // symbolic: small tables first, rows that overflow are retried with the large table
  int flag = 1;
  for (int pass = 0; pass < 2 && flag; pass++) {
    B200_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
    if (pass == 0) {
      constexpr int CAP = 256;
      const size_t bytes = (size_t)WPB * sizeof(int) * CAP;
      spgemm_warp_kernel<CAP, GB, 0><<<occ_grid(h, spgemm_warp_kernel<CAP, GB, 0>, n, bytes), 32 * WPB, bytes, h->stream>>>(
          n, A->i, A->j, nullptr, B->i, B->j, nullptr, allsquare, diag_base, nullptr, cnt, nullptr, nullptr, nullptr, d_flag);
      B200_LAUNCH_CHECK();
    } else {
      constexpr int CAP = 2048;
      const size_t bytes = (size_t)WPB * sizeof(int) * CAP;
      int *rows = nullptr, m = 0;
      B200_TRY(build_row_list(h, n, cnt, PENDING, PENDING, &rows, &m));
      if (m > 0) {
        spgemm_warp_kernel<CAP, GB, 0><<<occ_grid(h, spgemm_warp_kernel<CAP, GB, 0>, m, bytes), 32 * WPB, bytes, h->stream>>>(
            m, A->i, A->j, nullptr, B->i, B->j, nullptr, allsquare, diag_base, rows, cnt, nullptr, nullptr, nullptr, d_flag);
        B200_LAUNCH_CHECK();
      }
      B200_TRY(b200_dfree(h, rows));
    }
    B200_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
  }
  if (flag) { B200_TRY(b200_dfree(h, cnt)); B200_TRY(b200_dfree(h, d_flag)); return 0; }
  int *C_i = nullptr;
  B200_TRY(b200_dalloc<int>(h, &C_i, (size_t)n + 1));
  B200_CUDA(cudaMemcpyAsync(C_i, cnt, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
  B200_TRY(b200_exclusive_scan_inplace(h, C_i, (size_t)n + 1));
  int nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, C_i + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr C = nullptr;
  B200_TRY(b200_csr_alloc(h, n, ncols_C, nnz, true, &C));
  B200_CUDA(cudaMemcpyAsync(C->i, C_i, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, h->stream));
  B200_TRY(b200_dfree(h, C_i));
  // numeric: one launch per size class (order-preserving row lists), each with the smallest table that holds the row
#define B200_SPGEMM_NUMERIC(CAPV, LO, BPS)                                                                        \
  {                                                                                                               \
    constexpr int CAP = CAPV;                                                                                     \
    const size_t bytes = (size_t)WPB * (sizeof(double) * (CAP / 2) + sizeof(int) * (2 * CAP + CAP / 2));          \
    int *rows = nullptr, m = 0;                                                                                   \
    B200_TRY(build_row_list(h, n, cnt, LO, CAP / 2, &rows, &m));                                                  \
    if (m > 0) {                                                                                                  \
      B200_TRY(set_smem(spgemm_warp_kernel<CAP, GB, 1>, bytes));                                               \
      spgemm_warp_kernel<CAP, GB, 1><<<occ_grid(h, spgemm_warp_kernel<CAP, GB, 1>, m, bytes), 32 * WPB, bytes, h->stream>>>(                    \
          m, A->i, A->j, A->a, B->i, B->j, B->a, allsquare, diag_base, rows, cnt, C->i, C->j, C->a, d_flag);                 \
      B200_LAUNCH_CHECK();                                                                                        \
    }                                                                                                             \
    B200_TRY(b200_dfree(h, rows));                                                                                \
  }
  B200_SPGEMM_NUMERIC(128, 0, 16)        // rows of 1..64 entries
  B200_SPGEMM_NUMERIC(512, 64, 7)        // 65..256
  B200_SPGEMM_NUMERIC(2048, 256, 1)      // 257..1024 (the symbolic pass guarantees <= 1024)
#undef B200_SPGEMM_NUMERIC
  B200_TRY(b200_dfree(h, cnt)); B200_TRY(b200_dfree(h, d_flag));
  *out = C;
  *done = 1;
  return 0;
}

int b200_csr_multiply_warp(b200_handle h, b200_csr A, b200_csr B, int allsquare, int diag_base, int ncols_C,
                           b200_csr *out, int *done) {
  *done = 0;
  const int n = A->nrows;
  if (n == 0) return 0;
  int *d_max = nullptr;
  B200_TRY(b200_dalloc<int>(h, &d_max, 1));
  B200_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), h->stream));
  if (B->nrows) {
    max_rowlen_kernel<<<b200_grid(B->nrows, 256), 256, 0, h->stream>>>(B->nrows, B->i, d_max);
    B200_LAUNCH_CHECK();
  }
  int maxlen = 0;
  B200_CUDA(cudaMemcpyAsync(&maxlen, d_max, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, d_max));
  if (maxlen <= 4) return spgemm_warp_run<4>(h, A, B, allsquare, diag_base, ncols_C, out, done);
  if (maxlen <= 8) return spgemm_warp_run<8>(h, A, B, allsquare, diag_base, ncols_C, out, done);
  if (maxlen <= 16) return spgemm_warp_run<16>(h, A, B, allsquare, diag_base, ncols_C, out, done);
  return spgemm_warp_run<32>(h, A, B, allsquare, diag_base, ncols_C, out, done);
}
