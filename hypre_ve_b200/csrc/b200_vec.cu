// b200_vec.cu -- BLAS-1 kernels of the Krylov / V-cycle path.
// Reference: hypre_SeqVector{SetConstantValues,Copy,Scale,Axpy,InnerProd}
// (seq_mv/vector.c:238,:321,:394,:451,:511).  Dot products are deterministic: a fixed grid of
// CTAs each reduces a fixed slice with warp shuffles, then one CTA folds the partials in order.
#include "b200_internal.h"

namespace {
constexpr int VT = 256;

__global__ void fill_kernel(size_t n, double v, double *__restrict__ x) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = v;
}
__global__ void scale_kernel(size_t n, double a, double *__restrict__ y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] *= a;
}
__global__ void axpy_kernel(size_t n, double a, const double *__restrict__ x, double *__restrict__ y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] += a * x[i];
}

__device__ __forceinline__ double block_sum(double s) {
  __shared__ double w[VT / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = threadIdx.x < VT / 32 ? w[threadIdx.x] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
  }
  __syncthreads();
  return t;   // valid in thread 0
}

// Fixed grid, fixed slices: deterministic for a given n.  128-bit loads, four of them in flight per thread and operand
// (16N bytes stream at the HBM roof only with >= 100 KB in flight per SM).  TWO = 1 also accumulates <x,x> in the same
// pass (PCG needs <r,s> and <r,r> at the same point: one read of r instead of two).
template <int TWO>
__global__ void __launch_bounds__(VT) dot_partial_kernel(size_t n, const double *__restrict__ x,
                                                          const double *__restrict__ y,
                                                          double *__restrict__ partial, double *__restrict__ partial2) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double s = 0.0, q = 0.0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  size_t done = 0;
  if (aligned) {
    const size_t n2 = n / 2;
    const double2 *x2 = reinterpret_cast<const double2 *>(x), *y2 = reinterpret_cast<const double2 *>(y);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    size_t i = tid;
    for (; i + 3 * stride < n2; i += 4 * stride) {
      const double2 u0 = x2[i], u1 = x2[i + stride], u2 = x2[i + 2 * stride], u3 = x2[i + 3 * stride];
      const double2 v0 = y2[i], v1 = y2[i + stride], v2 = y2[i + 2 * stride], v3 = y2[i + 3 * stride];
      a0 += u0.x * v0.x; a0 += u0.y * v0.y; a1 += u1.x * v1.x; a1 += u1.y * v1.y;
      a2 += u2.x * v2.x; a2 += u2.y * v2.y; a3 += u3.x * v3.x; a3 += u3.y * v3.y;
      if (TWO) {
        b0 += u0.x * u0.x; b0 += u0.y * u0.y; b1 += u1.x * u1.x; b1 += u1.y * u1.y;
        b2 += u2.x * u2.x; b2 += u2.y * u2.y; b3 += u3.x * u3.x; b3 += u3.y * u3.y;
      }
    }
    for (; i < n2; i += stride) {
      const double2 u = x2[i], v = y2[i];
      a0 += u.x * v.x; a0 += u.y * v.y;
      if (TWO) { b0 += u.x * u.x; b0 += u.y * u.y; }
    }
    s = (a0 + a1) + (a2 + a3);
    q = (b0 + b1) + (b2 + b3);
    done = n2 * 2;
  }
  for (size_t i = done + tid; i < n; i += stride) { s += x[i] * y[i]; if (TWO) q += x[i] * x[i]; }
  double t = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
  if (TWO) {
    double t2 = block_sum(q);
    if (threadIdx.x == 0) partial2[blockIdx.x] = t2;
  }
}
__global__ void __launch_bounds__(VT) dot_final_kernel(int np, const double *__restrict__ partial,
                                                        double *__restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < np; i += VT) s += partial[i];
  double t = block_sum(s);
  if (threadIdx.x == 0) out[0] = t;
}
}  // namespace

static inline int vec_grid(b200_handle h, size_t n) {
  size_t g = (n + VT - 1) / VT;
  size_t cap = (size_t)h->num_sm * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

extern "C" int b200_vec_fill(b200_handle h, int n, double v, double *x) {
  if (n <= 0) return 0;
  fill_kernel<<<vec_grid(h, n), VT, 0, h->stream>>>((size_t)n, v, x);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_vec_copy(b200_handle h, int n, const double *x, double *y) {
  if (n <= 0) return 0;
  B200_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}
extern "C" int b200_vec_scale(b200_handle h, int n, double a, double *y) {
  if (n <= 0) return 0;
  scale_kernel<<<vec_grid(h, n), VT, 0, h->stream>>>((size_t)n, a, y);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_vec_axpy(b200_handle h, int n, double a, const double *x, double *y) {
  if (n <= 0) return 0;
  axpy_kernel<<<vec_grid(h, n), VT, 0, h->stream>>>((size_t)n, a, x, y);
  B200_LAUNCH_CHECK();
  return 0;
}

// device-side result (no sync): out is a device pointer
int b200_vec_dot_dev(b200_handle h, int n, const double *x, const double *y, double *d_out) {
  int g = vec_grid(h, n > 0 ? n : 1);
  if (g > h->n_partials) g = h->n_partials;
  dot_partial_kernel<0><<<g, VT, 0, h->stream>>>((size_t)(n > 0 ? n : 0), x, y, h->d_partials, nullptr);
  B200_LAUNCH_CHECK();
  dot_final_kernel<<<1, VT, 0, h->stream>>>(g, h->d_partials, d_out);
  B200_LAUNCH_CHECK();
  return 0;
}

// <x,y> -> d_out[0] and <x,x> -> d_out[stride2] in ONE pass over x (device-side results, no sync)
__global__ void __launch_bounds__(VT) dot2_final_kernel(int np, const double *__restrict__ p1, const double *__restrict__ p2,
                                                         double *__restrict__ out1, double *__restrict__ out2) {
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < np; i += VT) { s += p1[i]; q += p2[i]; }
  const double t = block_sum(s);
  const double t2 = block_sum(q);
  if (threadIdx.x == 0) { out1[0] = t; out2[0] = t2; }
}
int b200_vec_dot2_dev(b200_handle h, int n, const double *x, const double *y, double *d_xy, double *d_xx) {
  int g = vec_grid(h, n > 0 ? n : 1);
  if (g > h->n_partials / 2) g = h->n_partials / 2;
  dot_partial_kernel<1><<<g, VT, 0, h->stream>>>((size_t)(n > 0 ? n : 0), x, y, h->d_partials, h->d_partials + h->n_partials / 2);
  B200_LAUNCH_CHECK();
  dot2_final_kernel<<<1, VT, 0, h->stream>>>(g, h->d_partials, h->d_partials + h->n_partials / 2, d_xy, d_xx);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_vec_dot(b200_handle h, int n, const double *x, const double *y, double *result) {
  double *d_out = h->d_partials + (h->n_partials - 1);   // last slot is never used as a partial (g < n_partials)
  int g = vec_grid(h, n > 0 ? n : 1);
  if (g > h->n_partials - 1) g = h->n_partials - 1;
  dot_partial_kernel<0><<<g, VT, 0, h->stream>>>((size_t)(n > 0 ? n : 0), x, y, h->d_partials, nullptr);
  B200_LAUNCH_CHECK();
  dot_final_kernel<<<1, VT, 0, h->stream>>>(g, h->d_partials, d_out);
  B200_LAUNCH_CHECK();
  B200_CUDA(cudaMemcpyAsync(h->h_pinned, d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  *result = h->h_pinned[0];
  return 0;
}
