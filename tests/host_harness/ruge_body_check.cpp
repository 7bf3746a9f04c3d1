// ruge_body_check.cpp -- TEST INFRASTRUCTURE.  Compiles hypre_ve_b200/csrc/b200_hmis_body.h (the text the device runs on one
// thread for the Ruge-Stueben first pass of HMIS) for the HOST and runs it on a strength pattern read from a file, so that
// the transcription can be compared with the reference on a machine without a GPU (tests/test_oracle.py).
// usage: ruge_body_check IN OUT [agg2]     IN: int32 n, nnz, I[n+1], J[nnz]     OUT: int32 status, cf[n]
#include <cstdio>
#include <cstdlib>
#include <vector>
#define B200_HD
#include "../../hypre_ve_b200/csrc/b200_hmis_body.h"

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  int n = 0, nnz = 0;
  if (fread(&n, 4, 1, f) != 1 || fread(&nnz, 4, 1, f) != 1) return 2;
  std::vector<int> I(n + 1), J(nnz > 0 ? nnz : 1);
  if (fread(I.data(), 4, n + 1, f) != (size_t)n + 1) return 2;
  if (nnz && fread(J.data(), 4, nnz, f) != (size_t)nnz) return 2;
  fclose(f);
  // S^T with rows ordered by source row, as b200_csr_transpose (stable sort by column) leaves it
  std::vector<int> TI(n + 1, 0), TJ(nnz > 0 ? nnz : 1);
  for (int k = 0; k < nnz; k++) TI[J[k] + 1]++;
  for (int i = 0; i < n; i++) TI[i + 1] += TI[i];
  std::vector<int> nxt(TI.begin(), TI.end());
  for (int i = 0; i < n; i++)
    for (int k = I[i]; k < I[i + 1]; k++) TJ[nxt[J[k]]++] = i;
  int hmax = 0;
  for (int i = 0; i < n; i++) hmax = TI[i + 1] - TI[i] > hmax ? TI[i + 1] - TI[i] : hmax;
  const int nb = 2 * hmax + 4;                                   // as b200_ruge_first_pass sizes the buckets
  std::vector<int> cf(n), meas(n), next(n), prev(n), head(nb, -1), tail(nb, -1);
  const int status = b200_ruge_first_pass_body(n, I.data(), J.data(), TI.data(), TJ.data(), argc > 3 ? atoi(argv[3]) : 0, nb,
                                               cf.data(), meas.data(), next.data(), prev.data(), head.data(), tail.data());
  f = fopen(argv[2], "wb");
  fwrite(&status, 4, 1, f);
  fwrite(cf.data(), 4, n, f);
  fclose(f);
  return 0;
}
