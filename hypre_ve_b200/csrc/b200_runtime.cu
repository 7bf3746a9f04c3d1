// b200_runtime.cu -- handle, stream, stream-ordered memory pool, scans.
// Plays the role of hypre_handle()/hypre_TAlloc(HYPRE_MEMORY_DEVICE) in the reference
// (utilities/hypre_general.c:128-197, utilities/hypre_memory.c) for the B200 path.
#include "b200_internal.h"
#include <cub/device/device_scan.cuh>
#include <map>
#include <mutex>

// ------------------------------------------------------------------------------------------
// Slab sub-allocator.  BoomerAMG setup allocates and frees hundreds of temporaries, several of
// them gigabytes; going to the driver for each (cudaMalloc / cudaMallocAsync pool growth and
// remapping) cost more than the kernels (profiles/README.md r1_b).  Slabs are cudaMalloc'd once
// (>= 1 GiB, or the request if larger), carved best-fit with address-ordered coalescing of free
// blocks, and only returned to the driver at b200_finalize.
// ------------------------------------------------------------------------------------------
struct b200_pool_s {
  static constexpr size_t ALIGN = 512;
  static constexpr size_t SLAB = (size_t)1 << 30;
  std::vector<void *> slabs;
  std::vector<size_t> slab_size;
  std::map<char *, size_t> free_by_addr;              // address -> size
  std::multimap<size_t, char *> free_by_size;         // size -> address
  std::map<char *, size_t> used;                      // address -> size
  size_t total = 0, in_use = 0, peak = 0;

  void add_free(char *p, size_t n) {
    // coalesce with the neighbours (never across slabs: slab ends are not adjacent free blocks by construction
    // unless the driver returned adjacent ranges, which is harmless for a single-owner pool)
    auto next = free_by_addr.lower_bound(p);
    if (next != free_by_addr.end() && p + n == next->first && !is_slab_start(next->first)) {
      n += next->second;
      erase_size(next->second, next->first);
      next = free_by_addr.erase(next);
    }
    if (next != free_by_addr.begin()) {
      auto prev = std::prev(next);
      if (prev->first + prev->second == p && !is_slab_start(p)) {
        p = prev->first;
        n += prev->second;
        erase_size(prev->second, prev->first);
        free_by_addr.erase(prev);
      }
    }
    free_by_addr[p] = n;
    free_by_size.emplace(n, p);
  }
  bool is_slab_start(char *p) const {
    for (void *s : slabs) if (s == (void *)p) return true;
    return false;
  }
  void erase_size(size_t n, char *p) {
    auto r = free_by_size.equal_range(n);
    for (auto it = r.first; it != r.second; ++it)
      if (it->second == p) { free_by_size.erase(it); return; }
  }
};

int b200_pool_alloc(b200_handle h, void **out, size_t bytes) {
  b200_pool_s *P = h->pool;
  size_t n = (bytes + b200_pool_s::ALIGN - 1) / b200_pool_s::ALIGN * b200_pool_s::ALIGN;
  auto it = P->free_by_size.lower_bound(n);
  if (it == P->free_by_size.end()) {
    // slabs grow geometrically from 64 MiB to 1 GiB: a handle that only ever holds a small problem (one rank of a
    // threads-as-ranks test, a coarse-level service) does not reserve a gigabyte, a large one reaches 1 GiB slabs after 5
    size_t want = P->total < ((size_t)64 << 20) ? ((size_t)64 << 20) : (P->total > b200_pool_s::SLAB ? b200_pool_s::SLAB : P->total);
    size_t slab = n > want ? n : want;
    void *s = nullptr;
    cudaError_t e = cudaMalloc(&s, slab);
    if (e != cudaSuccess) return b200_set_error(__FILE__, __LINE__, cudaGetErrorString(e));
    P->slabs.push_back(s);
    P->slab_size.push_back(slab);
    P->total += slab;
    P->free_by_addr[(char *)s] = slab;
    it = P->free_by_size.emplace(slab, (char *)s);
  }
  char *p = it->second;
  size_t have = it->first;
  P->free_by_size.erase(it);
  P->free_by_addr.erase(p);
  if (have > n) {
    P->free_by_addr[p + n] = have - n;
    P->free_by_size.emplace(have - n, p + n);
  }
  P->used[p] = n;
  P->in_use += n;
  if (P->in_use > P->peak) P->peak = P->in_use;
  *out = p;
  return 0;
}

int b200_pool_free(b200_handle h, void *ptr) {
  b200_pool_s *P = h->pool;
  auto it = P->used.find((char *)ptr);
  if (it == P->used.end()) return b200_set_error(__FILE__, __LINE__, "free of a pointer this handle did not allocate");
  size_t n = it->second;
  P->used.erase(it);
  P->in_use -= n;
  P->add_free((char *)ptr, n);
  return 0;
}

// Return every slab that is completely free to the driver (cudaFree synchronises the device: call it between
// phases, never inside a timed region).  Long-lived handles (a test session, a service) call this after
// destroying a hierarchy so that the high-water mark of one problem does not stay reserved for good.
extern "C" int b200_pool_trim(b200_handle h, size_t *bytes_released) {
  b200_pool_s *P = h->pool;
  size_t released = 0;
  B200_CUDA(cudaStreamSynchronize(h->stream));
  for (size_t k = 0; k < P->slabs.size();) {
    char *s = (char *)P->slabs[k];
    auto it = P->free_by_addr.find(s);
    // free blocks never coalesce across a slab start (add_free), so the slab is entirely free iff one free block covers it
    size_t n = (it != P->free_by_addr.end()) ? it->second : 0;
    bool whole = n == P->slab_size[k];
    if (whole) {
      P->erase_size(n, s);
      P->free_by_addr.erase(it);
      B200_CUDA(cudaFree(s));
      P->total -= n;
      released += n;
      P->slabs.erase(P->slabs.begin() + k);
      P->slab_size.erase(P->slab_size.begin() + k);
    } else k++;
  }
  if (bytes_released) *bytes_released = released;
  return 0;
}
extern "C" int b200_pool_stats(b200_handle h, size_t *reserved, size_t *in_use, size_t *peak) {
  if (reserved) *reserved = h->pool->total;
  if (in_use) *in_use = h->pool->in_use;
  if (peak) *peak = h->pool->peak;
  return 0;
}

// ---- region profiler (B200_PROF=1) -----------------------------------------------------------------------------------------
struct b200_prof_s {
  struct Rec { std::string label; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<size_t> open;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};
static bool prof_on() {
  static const bool on = [] { const char *e = getenv("B200_PROF"); return e && e[0] == '1'; }();
  return on;
}
void b200_prof_begin(b200_handle h, const char *label, int level) {
  if (!prof_on()) return;
  if (!h->prof) h->prof = new b200_prof_s();
  b200_prof_s::Rec r;
  char buf[96];
  if (level >= 0) { snprintf(buf, sizeof buf, "L%d %s", level, label); r.label = buf; } else r.label = label;
  r.a = h->prof->get(); r.b = h->prof->get();
  cudaEventRecord(r.a, h->stream);
  h->prof->open.push_back(h->prof->recs.size());
  h->prof->recs.push_back(r);
}
void b200_prof_end(b200_handle h) {
  if (!prof_on() || !h->prof || h->prof->open.empty()) return;
  cudaEventRecord(h->prof->recs[h->prof->open.back()].b, h->stream);
  h->prof->open.pop_back();
}
void b200_prof_report(b200_handle h, const char *title) {
  if (!prof_on() || !h->prof) return;
  cudaStreamSynchronize(h->stream);
  std::map<std::string, std::pair<double, int>> acc;
  for (auto &r : h->prof->recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    acc[r.label].first += ms; acc[r.label].second++;
    h->prof->pool.push_back(r.a); h->prof->pool.push_back(r.b);
  }
  h->prof->recs.clear(); h->prof->open.clear();
  fprintf(stderr, "[b200 prof] %s (device %d)\n", title, h->device);
  for (auto &kv : acc)
    fprintf(stderr, "[b200 prof]   %-34s %10.3f ms  %6d calls  %9.2f us/call\n", kv.first.c_str(), kv.second.first, kv.second.second,
            1e3 * kv.second.first / kv.second.second);
}

// ---- CUDA-graph replay ---------------------------------------------------------------------------------------------------------
bool b200_graph_enabled() {
  static const bool on = [] { const char *e = getenv("B200_GRAPH"); return !(e && e[0] == '0'); }();
  return on && !prof_on();
}
int b200_graph_begin(b200_handle h) {
  B200_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
  return 0;
}
int b200_graph_end(b200_handle h, cudaGraphExec_t *exec) {
  *exec = nullptr;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(h->stream, &g);
  if (e != cudaSuccess || !g) { cudaGetLastError(); if (g) cudaGraphDestroy(g); return 0; }
  e = cudaGraphInstantiate(exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { cudaGetLastError(); *exec = nullptr; }
  return 0;
}
int b200_graph_launch(b200_handle h, cudaGraphExec_t exec) {
  B200_CUDA(cudaGraphLaunch(exec, h->stream));
  return 0;
}
void b200_graph_destroy(cudaGraphExec_t exec) { if (exec) cudaGraphExecDestroy(exec); }

thread_local std::string g_b200_err;
std::atomic<long long> g_b200_launches{0};

// pinned, device-visible diagnostic words: a peer-to-peer wait that times out records what it was waiting for before it
// traps (the context is lost after the trap; pinned host memory is not)
unsigned long long *g_b200_p2p_dbg = nullptr;

int b200_set_error(const char *file, int line, const char *msg) {
  char buf[1024];
  if (g_b200_p2p_dbg && g_b200_p2p_dbg[0])
    snprintf(buf, sizeof buf, "%s:%d: %s [peer-to-peer wait timed out: kind %llu (1 halo push/ack, 2 halo pull/arrival, 3 allreduce), "
             "peer slot %llu, wanted sequence %llu, found %llu]", file, line, msg, g_b200_p2p_dbg[0], g_b200_p2p_dbg[1],
             g_b200_p2p_dbg[2], g_b200_p2p_dbg[3]);
  else
    snprintf(buf, sizeof buf, "%s:%d: %s", file, line, msg);
  g_b200_err = buf;
  return 1;
}

extern "C" const char *b200_last_error(void) { return g_b200_err.c_str(); }
extern "C" long long b200_launch_count(void) { return g_b200_launches.load(); }

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
  h->pool = new b200_pool_s();
  h->n_partials = 4096;
  B200_CUDA(cudaMalloc(&h->d_partials, sizeof(double) * h->n_partials));
  B200_CUDA(cudaMallocHost(&h->h_pinned, sizeof(double) * 1024));
  {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!g_b200_p2p_dbg) {
      B200_CUDA(cudaHostAlloc((void **)&g_b200_p2p_dbg, sizeof(unsigned long long) * 8, cudaHostAllocPortable | cudaHostAllocMapped));
      memset(g_b200_p2p_dbg, 0, sizeof(unsigned long long) * 8);
    }
  }
  *out = h;
  return 0;
}

extern "C" int b200_finalize(b200_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->d_partials);
  cudaFreeHost(h->h_pinned);
  for (void *s : h->pool->slabs) cudaFree(s);
  delete h->pool;
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
