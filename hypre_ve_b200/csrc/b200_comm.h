// b200_comm.h -- internal interface of the communication layer (b200_comm.cu) and the
// halo plan (b200_dist.cu).  See include/hypre_b200.h for the exported C-ABI.
#pragma once
#include "b200_internal.h"
#include <vector>

struct b200_xfer {
  int peer;
  void *ptr;       // device pointer
  size_t bytes;
  int tag;         // ordinal among the messages exchanged with `peer` in this call (0 = first)
};

int b200_comm_exchange(b200_handle h, b200_comm c, const std::vector<b200_xfer> &sends, const std::vector<b200_xfer> &recvs);
int b200_comm_allgather_host(b200_handle h, b200_comm c, const void *mine, size_t bytes, void *all);
int b200_comm_allreduce_sum(b200_handle h, b200_comm c, double *vals, int k);
int b200_comm_allreduce_sum_dev(b200_handle h, b200_comm c, const double *d_vals, int k, double *h_out);
int b200_comm_allreduce_sum_ll(b200_handle h, b200_comm c, long long *vals, int k);

// ---- direct peer-to-peer layer (NVLink loads/stores, no NCCL call on the data path) -----------------------------------
// Every rank owns an arena of the same size; all arenas are mapped into every rank's address space once, at communicator
// creation (CUDA IPC between processes, plain pointers between the rank threads of one process).  The arena is a SYMMETRIC
// heap: allocation and release are collective and deterministic, so an object sits at the same offset on every rank and the
// address of a peer's copy is peer_base[r] + offset -- no per-object handshake.  Synchronisation is by sequence numbers
// written with system-scope release stores: one counter per communicator, identical on all ranks because every rank runs
// the same sequence of collective operations.
#define B200_P2P_MAXPEER 16
#define B200_P2P_SLOT 16             // flag slots are 128 bytes apart (unsigned long long [16])
bool   b200_comm_p2p_ok(b200_comm c);
// symmetric allocation; returns the offset or (size_t)-1 when the arena is exhausted (same answer on every rank)
size_t b200_comm_p2p_alloc(b200_handle h, b200_comm c, size_t bytes);
// released regions are reused only after a collective quiescence point (taken lazily by the next allocation)
void   b200_comm_p2p_free(b200_comm c, size_t offset, size_t bytes);
char  *b200_comm_p2p_base(b200_comm c, int rank);          // base of rank's arena in this rank's address space
// k <= 8 device-resident partial sums -> global sums in device memory on every rank, added in rank order (deterministic);
// one kernel, no host synchronisation.  Falls back to ncclAllGather + a summation kernel without the peer layer.
long long b200_comm_host_ops(b200_comm c);      // number of exchanges that did not stay on the device-only path
int b200_comm_allreduce_sum_dev2dev(b200_handle h, b200_comm c, const double *d_vals, int k, double *d_out);

// Halo plan = hypre_ParCSRCommPkg (parcsr_mv/par_csr_communication.h:54-82) for one ghost set:
// ghosts are the sorted global ids this rank reads but does not own; each ghost block is
// contiguous per owner (owners own contiguous id ranges), so receives land in place.
struct b200_halo_s {
  int n_owned = 0;                       // ids owned by this rank: [first, first + n_owned)
  int first = 0;
  int ng = 0;                            // number of ghosts
  bool any_traffic = false;              // some rank sends or receives under this plan: the exchange is collective, so a rank
                                         // without ghosts and without send entries still takes part (rank-group backend)
  int *d_ghost_gid = nullptr;            // device, sorted [ng]
  std::vector<int> recv_cnt, recv_off;   // per peer, into the ghost array
  std::vector<int> send_cnt, send_off;   // per peer, into d_send_idx
  int n_send = 0;
  int *d_send_idx = nullptr;             // device, local indices to pack [n_send]
  void *d_send_buf = nullptr;            // device staging, 8 bytes per send entry
  // direct path (b200_halo_forward_f64): enabled lazily by the first forward exchange of doubles under this plan
  b200_comm comm = nullptr;
  std::vector<int> all_cnt;              // [R x R] all_cnt[r * R + s] = entries rank r receives from rank s
  int p2p_state = 0;                     // 0 not tried, 1 enabled, -1 unavailable (NCCL path)
  size_t p2p_off = 0, p2p_bytes = 0;     // symmetric region: [rbuf parity 0 | rbuf parity 1 | arrival flags | ack flags | local: count, CTA counters]
  int p2p_cap = 0;                       // doubles per parity buffer (max ghosts over the ranks, rounded up)
  void *p2p_args = nullptr;              // the exchange kernel's argument block (fixed for the life of the plan)
};

// halo operations (b200_dist.cu)
int b200_halo_build(b200_handle h, b200_comm c, const std::vector<int> &starts, int *d_ghost_gid_sorted, int ng,
                    b200_halo_s **plan_out);      // takes ownership of d_ghost_gid_sorted
void b200_halo_free(b200_handle h, b200_halo_s *p);
int b200_halo_forward_i32(b200_handle h, b200_comm c, b200_halo_s *p, const int *d_owned, int *d_ghost_out);
int b200_halo_forward_f64(b200_handle h, b200_comm c, b200_halo_s *p, const double *d_owned, double *d_ghost_out);
int b200_halo_reverse_add_i32(b200_handle h, b200_comm c, b200_halo_s *p, const int *d_ghost_in, int *d_owned_inout);
// reference rule of par_coarsen.c:2509-2526: a ghost copy that was cleared clears the owner's tentative mark
int b200_halo_reverse_clear_i32(b200_handle h, b200_comm c, b200_halo_s *p, const int *d_ghost_in, int *d_owned_inout);
