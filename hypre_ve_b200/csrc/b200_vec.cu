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

__global__ void __launch_bounds__(VT) dot_partial_kernel(size_t n, const double *__restrict__ x,
                                                          const double *__restrict__ y,
                                                          double *__restrict__ partial) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double s0 = 0.0, s1 = 0.0;
  for (; i + stride < n; i += 2 * stride) {
    s0 += x[i] * y[i];
    s1 += x[i + stride] * y[i + stride];
  }
  if (i < n) s0 += x[i] * y[i];
  double t = block_sum(s0 + s1);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
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
  dot_partial_kernel<<<g, VT, 0, h->stream>>>((size_t)(n > 0 ? n : 0), x, y, h->d_partials);
  B200_LAUNCH_CHECK();
  dot_final_kernel<<<1, VT, 0, h->stream>>>(g, h->d_partials, d_out);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_vec_dot(b200_handle h, int n, const double *x, const double *y, double *result) {
  double *d_out = h->d_partials + (h->n_partials - 1);   // last slot is never used as a partial (g < n_partials)
  int g = vec_grid(h, n > 0 ? n : 1);
  if (g > h->n_partials - 1) g = h->n_partials - 1;
  dot_partial_kernel<<<g, VT, 0, h->stream>>>((size_t)(n > 0 ? n : 0), x, y, h->d_partials);
  B200_LAUNCH_CHECK();
  dot_final_kernel<<<1, VT, 0, h->stream>>>(g, h->d_partials, d_out);
  B200_LAUNCH_CHECK();
  B200_CUDA(cudaMemcpyAsync(h->h_pinned, d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  *result = h->h_pinned[0];
  return 0;
}
