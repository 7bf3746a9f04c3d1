// b200_runtime.cu -- handle, stream, stream-ordered memory pool, scans.
// Plays the role of hypre_handle()/hypre_TAlloc(HYPRE_MEMORY_DEVICE) in the reference
// (utilities/hypre_general.c:128-197, utilities/hypre_memory.c) for the B200 path.
#include "b200_internal.h"
#include <cub/device/device_scan.cuh>

thread_local std::string g_b200_err;
long long g_b200_launches = 0;

int b200_set_error(const char *file, int line, const char *msg) {
  char buf[1024];
  snprintf(buf, sizeof buf, "%s:%d: %s", file, line, msg);
  g_b200_err = buf;
  return 1;
}

extern "C" const char *b200_last_error(void) { return g_b200_err.c_str(); }
extern "C" long long b200_launch_count(void) { return g_b200_launches; }

extern "C" int b200_init(int device, b200_handle *out) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    B200_FAIL("no CUDA device: libhypre_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) B200_FAIL("bad device index");
  B200_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) B200_FAIL("libhypre_b200 is built for sm_100a (B200) only");
  b200_handle h = new b200_handle_s();
  h->device = device;
  h->num_sm = prop.multiProcessorCount;
  B200_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  B200_CUDA(cudaEventCreate(&h->ev0));
  B200_CUDA(cudaEventCreate(&h->ev1));
  // keep freed blocks cached in the pool: setup allocates and frees many temporaries
  cudaMemPool_t pool;
  B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  h->n_partials = 4096;
  B200_CUDA(cudaMalloc(&h->d_partials, sizeof(double) * h->n_partials));
  B200_CUDA(cudaMallocHost(&h->h_pinned, sizeof(double) * 64));
  *out = h;
  return 0;
}

extern "C" int b200_finalize(b200_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->d_partials);
  cudaFreeHost(h->h_pinned);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

extern "C" void *b200_stream(b200_handle h) { return (void *)h->stream; }
extern "C" int b200_sync(b200_handle h) {
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int b200_malloc(b200_handle h, void **p, size_t bytes) {
  char *q = nullptr;
  B200_TRY(b200_dalloc<char>(h, &q, bytes));
  *p = q;
  return 0;
}
extern "C" int b200_free(b200_handle h, void *p) { return b200_dfree(h, p); }
extern "C" int b200_memcpy_h2d(b200_handle h, void *d, const void *s, size_t bytes) {
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}
extern "C" int b200_memcpy_d2h(b200_handle h, void *d, const void *s, size_t bytes) {
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int b200_memcpy_d2d(b200_handle h, void *d, const void *s, size_t bytes) {
  B200_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}
extern "C" int b200_memset(b200_handle h, void *d, int byte, size_t bytes) {
  B200_CUDA(cudaMemsetAsync(d, byte, bytes, h->stream));
  return 0;
}
extern "C" int b200_timer_start(b200_handle h) {
  B200_CUDA(cudaEventRecord(h->ev0, h->stream));
  return 0;
}
extern "C" int b200_timer_stop_ms(b200_handle h, double *ms) {
  B200_CUDA(cudaEventRecord(h->ev1, h->stream));
  B200_CUDA(cudaEventSynchronize(h->ev1));
  float f = 0.f;
  B200_CUDA(cudaEventElapsedTime(&f, h->ev0, h->ev1));
  *ms = f;
  return 0;
}

// in-place exclusive prefix sum over n ints (hypre_prefix_sum* in utilities/hypre_prefix_sum.c)
int b200_exclusive_scan_inplace(b200_handle h, int *d, size_t n) {
  if (n == 0) return 0;
  size_t tmp_bytes = 0;
  B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d, d, (int)n, h->stream));
  char *tmp = nullptr;
  B200_TRY(b200_dalloc<char>(h, &tmp, tmp_bytes));
  B200_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d, d, (int)n, h->stream));
  ++g_b200_launches;
  B200_TRY(b200_dfree(h, tmp));
  return 0;
}

namespace {
__global__ void sum_int_kernel(size_t n, const int *__restrict__ d, unsigned long long *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  long long s = 0;
  for (; i < n; i += stride) s += d[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, (unsigned long long)s);   // integer sum: order-free
}
}  // namespace

int b200_reduce_sum_int(b200_handle h, const int *d, size_t n, long long *out) {
  unsigned long long *d_out = nullptr;
  B200_TRY(b200_dalloc<unsigned long long>(h, &d_out, 1));
  B200_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), h->stream));
  if (n) {
    size_t g = (n + 255) / 256, cap = (size_t)h->num_sm * 8;
    sum_int_kernel<<<(int)(g < cap ? g : cap), 256, 0, h->stream>>>(n, d, d_out);
    B200_LAUNCH_CHECK();
  }
  B200_CUDA(cudaMemcpyAsync(out, d_out, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_TRY(b200_dfree(h, d_out));
  return 0;
}
