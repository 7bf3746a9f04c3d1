// b200_comm.cu -- inter-GPU communication for the row-partitioned (ParCSR) path.
//
// Reference: hypre_ParCSRCommPkg / hypre_ParCSRCommHandleCreate_v2 / Destroy
// (parcsr_mv/par_csr_communication.c:307-631: nonblocking neighbour exchange, jobs 1/2 double
// forward/reverse, 11/12 int forward/reverse), hypre_MatvecCommPkgCreate (:909-947), Allreduce in
// hypre_ParVectorInnerProd (parcsr_mv/par_vector.c:481-501), MPI_Scan/Allgather for coarse
// numbering (parcsr_ls/par_coarse_parms.c:99-122).
//
// B200 design: one process per GPU; the exchange primitive is a grouped ncclSend/ncclRecv over
// NVLink enqueued on the handle's stream (NCCL is dlopen'ed from the library torch already
// loaded; the communicator is created from a unique id the host broadcasts, e.g. with
// torch.distributed).  A second backend runs N ranks as N host threads of ONE process on ONE GPU
// (device-to-device copies through a shared mailbox): it exists so that the multi-rank algorithms
// can be tested on a single-GPU box, as B200_PROFILING.md prescribes.
#include "b200_internal.h"
#include "b200_comm.h"
#include <dlfcn.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <algorithm>

// ---- minimal NCCL ABI (nccl.h 2.27/2.28: stable since 2.x) -------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4, ncclFloat64 = 8 };
enum { ncclSum = 0 };
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
static int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);      // several rank threads may open communicators at once
  if (g_nccl.ok) return 0;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the copy torch has loaded, if any
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) B200_FAIL("cannot dlopen libnccl.so.2 (import torch first, or put NCCL on LD_LIBRARY_PATH)");
#define B200_SYM(name) *(void **)(&g_nccl.name) = dlsym(lib, "nccl" #name); if (!g_nccl.name) B200_FAIL("NCCL symbol nccl" #name " missing");
  B200_SYM(GetUniqueId) B200_SYM(CommInitRank) B200_SYM(CommDestroy) B200_SYM(CommAbort) B200_SYM(Send) B200_SYM(Recv)
  B200_SYM(GroupStart) B200_SYM(GroupEnd) B200_SYM(AllGather) B200_SYM(GetErrorString)
#undef B200_SYM
  g_nccl.ok = true;
  return 0;
}
#define B200_NCCL(call)                                                                      \
  do {                                                                                       \
    ncclResult_t r__ = (call);                                                               \
    if (r__ != 0) return b200_set_error(__FILE__, __LINE__, g_nccl.GetErrorString(r__));     \
  } while (0)

// ---- in-process "threads as ranks" group --------------------------------------------------------
struct b200_comm_group_s {
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long long generation = 0;
  // mailbox: what each rank exposes for the current exchange
  std::vector<std::vector<b200_xfer>> sends;          // [rank] -> list of (peer, ptr, bytes)
  std::vector<const void *> host_ptr;                 // [rank] host pointer for allgather
  bool aborted = false;       // set by a rank that failed (b200_comm_group_abort): its peers return an error instead of waiting
  int timeout_s = 300;        // B200_COMM_TIMEOUT_S: a peer that never arrives turns into an error, not a hang
  // returns 0, or nonzero when the group was aborted / a peer did not arrive in time
  int barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) return 1;
    long long gen = generation;
    if (++arrived == nranks) { arrived = 0; ++generation; cv.notify_all(); return 0; }
    bool ok = cv.wait_for(lk, std::chrono::seconds(timeout_s), [&] { return generation != gen || aborted; });
    if (!ok) { aborted = true; cv.notify_all(); return 2; }
    return (generation != gen) ? 0 : 1;
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
};
#define B200_BARRIER(g)                                                                                          \
  do {                                                                                                           \
    int rb__ = (g)->barrier();                                                                                   \
    if (rb__) B200_FAIL(rb__ == 2 ? "rank group: a peer did not reach the exchange in time (B200_COMM_TIMEOUT_S)" \
                                  : "rank group aborted: another rank failed");                                  \
  } while (0)

struct b200_comm_s {
  int rank = 0, nranks = 1;
  int backend = 0;                 // 0 single, 1 nccl, 2 threads
  ncclComm_t nccl = nullptr;
  b200_comm_group_s *group = nullptr;
  void *d_stage = nullptr;         // device staging for host allgather over NCCL
  size_t stage_bytes = 0;
};

extern "C" int b200_comm_create_single(b200_comm *out) {
  *out = new b200_comm_s();
  return 0;
}
extern "C" int b200_comm_group_create(int nranks, b200_comm_group *out) {
  if (nranks < 1) B200_FAIL("nranks must be >= 1");
  b200_comm_group_s *g = new b200_comm_group_s();
  g->nranks = nranks;
  g->sends.resize(nranks);
  g->host_ptr.resize(nranks, nullptr);
  if (const char *e = getenv("B200_COMM_TIMEOUT_S")) { int t = atoi(e); if (t > 0) g->timeout_s = t; }
  *out = g;
  return 0;
}
extern "C" int b200_comm_group_destroy(b200_comm_group g) { delete g; return 0; }
// a rank that failed calls this so that its peers, blocked in (or arriving at) an exchange, return an error
extern "C" int b200_comm_group_abort(b200_comm_group g) { if (g) g->abort(); return 0; }
extern "C" int b200_comm_create_threads(b200_comm_group g, int rank, b200_comm *out) {
  if (!g || rank < 0 || rank >= g->nranks) B200_FAIL("bad group / rank");
  b200_comm_s *c = new b200_comm_s();
  c->rank = rank; c->nranks = g->nranks; c->backend = 2; c->group = g;
  *out = c;
  return 0;
}
extern "C" int b200_comm_nccl_unique_id(char *id128) {
  B200_TRY(load_nccl());
  ncclUniqueId id;
  B200_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return 0;
}
extern "C" int b200_comm_create_nccl(b200_handle h, int nranks, int rank, const char *id128, b200_comm *out) {
  B200_TRY(load_nccl());
  B200_CUDA(cudaSetDevice(h->device));
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  b200_comm_s *c = new b200_comm_s();
  c->rank = rank; c->nranks = nranks; c->backend = 1;
  B200_NCCL(g_nccl.CommInitRank(&c->nccl, nranks, id, rank));
  *out = c;
  return 0;
}
extern "C" int b200_comm_destroy(b200_handle h, b200_comm c) {
  if (!c) return 0;
  if (c->d_stage) b200_dfree(h, c->d_stage);
  if (c->backend == 1 && c->nccl) g_nccl.CommDestroy(c->nccl);
  delete c;
  return 0;
}
extern "C" int b200_comm_abort(b200_comm c) {
  if (!c) return 0;
  if (c->backend == 1 && c->nccl) { g_nccl.CommAbort(c->nccl); c->nccl = nullptr; }
  if (c->backend == 2 && c->group) c->group->abort();
  return 0;
}
extern "C" int b200_comm_rank(b200_comm c) { return c ? c->rank : 0; }
extern "C" int b200_comm_size(b200_comm c) { return c ? c->nranks : 1; }

// ---- the exchange primitive: every rank passes its sends and recvs (device pointers) -------------
int b200_comm_exchange(b200_handle h, b200_comm c, const std::vector<b200_xfer> &sends, const std::vector<b200_xfer> &recvs) {
  if (c->nranks == 1) {
    // self messages only
    for (const auto &r : recvs)
      for (const auto &s : sends)
        if (s.peer == c->rank && r.peer == c->rank && s.bytes == r.bytes && r.bytes)
          B200_CUDA(cudaMemcpyAsync(r.ptr, s.ptr, r.bytes, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
  }
  if (c->backend == 1) {
    bool any = false;
    for (const auto &s : sends) any = any || s.bytes;
    for (const auto &r : recvs) any = any || r.bytes;
    if (!any) return 0;                                  // nothing to or from this rank: no NCCL call at all
    B200_NCCL(g_nccl.GroupStart());
    for (const auto &s : sends)
      if (s.bytes) B200_NCCL(g_nccl.Send(s.ptr, s.bytes, ncclInt8, s.peer, c->nccl, h->stream));
    for (const auto &r : recvs)
      if (r.bytes) B200_NCCL(g_nccl.Recv(r.ptr, r.bytes, ncclInt8, r.peer, c->nccl, h->stream));
    B200_NCCL(g_nccl.GroupEnd());
    return 0;
  }
  // threads backend: publish sends, sync, pull from the peers' buffers, sync
  b200_comm_group_s *g = c->group;
  B200_CUDA(cudaStreamSynchronize(h->stream));           // my send buffers are complete
  g->sends[c->rank] = sends;
  B200_BARRIER(g);
  for (const auto &r : recvs) {
    if (!r.bytes) continue;
    const void *src = nullptr;
    int seen = 0;
    for (const auto &s : g->sends[r.peer])
      if (s.peer == c->rank) { if (seen == r.tag) { src = s.ptr; if (s.bytes != r.bytes) B200_FAIL("exchange size mismatch"); break; } seen++; }
    if (!src) B200_FAIL("exchange: matching send not found");
    B200_CUDA(cudaMemcpyAsync(r.ptr, src, r.bytes, cudaMemcpyDeviceToDevice, h->stream));
  }
  B200_CUDA(cudaStreamSynchronize(h->stream));
  B200_BARRIER(g);                                       // senders may now reuse their buffers
  return 0;
}

// host allgather of `bytes` bytes per rank
int b200_comm_allgather_host(b200_handle h, b200_comm c, const void *mine, size_t bytes, void *all) {
  if (c->nranks == 1) { memcpy(all, mine, bytes); return 0; }
  if (c->backend == 2) {
    b200_comm_group_s *g = c->group;
    g->host_ptr[c->rank] = mine;
    B200_BARRIER(g);
    for (int r = 0; r < c->nranks; r++) memcpy((char *)all + (size_t)r * bytes, g->host_ptr[r], bytes);
    B200_BARRIER(g);
    return 0;
  }
  const size_t need = bytes * (size_t)(c->nranks + 1);
  if (c->stage_bytes < need) {
    if (c->d_stage) B200_TRY(b200_dfree(h, c->d_stage));
    char *p = nullptr;
    B200_TRY(b200_dalloc<char>(h, &p, need * 2));
    c->d_stage = p; c->stage_bytes = need * 2;
  }
  char *send = (char *)c->d_stage, *recv = send + bytes;
  B200_CUDA(cudaMemcpyAsync(send, mine, bytes, cudaMemcpyHostToDevice, h->stream));
  B200_NCCL(g_nccl.AllGather(send, recv, bytes, ncclInt8, c->nccl, h->stream));
  B200_CUDA(cudaMemcpyAsync(all, recv, bytes * (size_t)c->nranks, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// k device-resident partial sums -> global sums on the host, added in rank order.  Over NCCL the partials go
// device-to-device into the all-gather and come back with ONE copy + synchronisation (no host round trip first).
int b200_comm_allreduce_sum_dev(b200_handle h, b200_comm c, const double *d_vals, int k, double *h_out) {
  if (k < 1 || (size_t)k * c->nranks > 64) B200_FAIL("allreduce_sum_dev: at most 64 values over all ranks");
  if (c->nranks == 1 || c->backend != 1) {
    B200_CUDA(cudaMemcpyAsync(h->h_pinned, d_vals, sizeof(double) * k, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    for (int j = 0; j < k; j++) h_out[j] = h->h_pinned[j];
    return b200_comm_allreduce_sum(h, c, h_out, k);
  }
  const size_t need = sizeof(double) * (size_t)k * c->nranks;
  if (c->stage_bytes < need) {
    if (c->d_stage) B200_TRY(b200_dfree(h, c->d_stage));
    char *p = nullptr;
    B200_TRY(b200_dalloc<char>(h, &p, 4096));
    c->d_stage = p; c->stage_bytes = 4096;
  }
  B200_NCCL(g_nccl.AllGather(d_vals, c->d_stage, sizeof(double) * k, ncclInt8, c->nccl, h->stream));
  B200_CUDA(cudaMemcpyAsync(h->h_pinned, c->d_stage, need, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  for (int j = 0; j < k; j++) {
    double s = 0.0;
    for (int r = 0; r < c->nranks; r++) s += h->h_pinned[(size_t)r * k + j];
    h_out[j] = s;
  }
  return 0;
}

// deterministic global sums: gather the per-rank partials and add them in rank order
int b200_comm_allreduce_sum(b200_handle h, b200_comm c, double *vals, int k) {
  if (c->nranks == 1) return 0;
  std::vector<double> all((size_t)k * c->nranks);
  B200_TRY(b200_comm_allgather_host(h, c, vals, sizeof(double) * k, all.data()));
  for (int j = 0; j < k; j++) {
    double s = 0.0;
    for (int r = 0; r < c->nranks; r++) s += all[(size_t)r * k + j];
    vals[j] = s;
  }
  return 0;
}
int b200_comm_allreduce_sum_ll(b200_handle h, b200_comm c, long long *vals, int k) {
  if (c->nranks == 1) return 0;
  std::vector<long long> all((size_t)k * c->nranks);
  B200_TRY(b200_comm_allgather_host(h, c, vals, sizeof(long long) * k, all.data()));
  for (int j = 0; j < k; j++) {
    long long s = 0;
    for (int r = 0; r < c->nranks; r++) s += all[(size_t)r * k + j];
    vals[j] = s;
  }
  return 0;
}
