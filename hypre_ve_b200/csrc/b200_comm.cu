// b200_comm.cu -- inter-GPU communication: halo plan + exchange (hypre_ParCSRCommPkg /
// hypre_ParCSRCommHandleCreate, parcsr_mv/par_csr_communication.c:307-631).
#include "b200_internal.h"

struct b200_halo_s { int unused; };

int b200_halo_exchange(b200_handle h, b200_parcsr A, const double *d_x) {
  (void)h; (void)A; (void)d_x;
  B200_FAIL("multi-rank halo exchange not built yet");
}
void b200_halo_destroy(b200_handle h, b200_halo_s *halo) { (void)h; delete halo; }
