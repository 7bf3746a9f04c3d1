// b200_hmis_body.h -- the sequential Ruge-Stueben first pass (par_coarsen.c:1130-1343 with the list order of
// utilities/amg_linklist.c) as one plain function.  b200_hmis.cu runs it on ONE device thread (B200_HD = __device__);
// tests/host_harness/ruge_body_check.cpp compiles the very same text for the host (B200_HD empty) so that the
// transcription can be checked on a machine without a GPU.  The library itself never runs it on the host.
#pragma once
#ifndef B200_HD
#define B200_HD __device__
#endif

struct b200_ruge_lists {                  // one FIFO per measure value: the order hypre_enter_on_lists / hypre_remove_point keep
  int *head, *tail, *next, *prev;
  int nb, maxm, bad;
  B200_HD void enter(int m, int i) {
    if (m < 0 || m >= nb) { bad = 1; return; }
    next[i] = -1;
    prev[i] = tail[m];
    if (tail[m] >= 0) next[tail[m]] = i; else head[m] = i;
    tail[m] = i;
    if (m > maxm) maxm = m;
  }
  B200_HD void remove(int m, int i) {
    if (m < 0 || m >= nb) { bad = 1; return; }
    const int p = prev[i], q = next[i];
    if (p >= 0) next[p] = q; else head[m] = q;
    if (q >= 0) prev[q] = p; else tail[m] = p;
    for (int guard = 0; guard < nb && maxm > 0 && head[maxm] < 0; guard++) maxm--;
  }
};

// markers as in par_coarsen.c:860-865: C_PT 1, F_PT -1, Z_PT -2, SF_PT -3, SC_PT 3, UNDECIDED 0
// returns 0, or 1 (a measure outside the buckets), 2 (empty list while points are left), 3 (points left at the end)
B200_HD inline int b200_ruge_first_pass_body(int n, const int *S_i, const int *S_j, const int *T_i, const int *T_j, int agg2, int nb,
                                             int *cf, int *meas, int *next, int *prev, int *head, int *tail) {
  b200_ruge_lists L{head, tail, next, prev, nb, 0, 0};
  int num_left = 0;
  for (int j = 0; j < n; j++) {                                   // :1130-1158, measures = row sums of S^T (:1056-1059)
    meas[j] = T_i[j + 1] - T_i[j];
    if (S_i[j + 1] - S_i[j] == 0) { cf[j] = agg2 ? 3 : -3; meas[j] = 0; }
    else { cf[j] = 0; num_left++; }
  }
  for (int j = 0; j < n; j++) {                                   // :1179-1222
    const int measure = meas[j];
    if (cf[j] == -3 || cf[j] == 3) continue;
    if (measure > 0) { L.enter(measure, j); continue; }
    cf[j] = -2;                                                   // nothing depends on j: f_pnt = Z_PT
    for (int k = S_i[j]; k < S_i[j + 1]; k++) {
      const int nabor = S_j[k];
      if (cf[nabor] == -3 || cf[nabor] == 3) continue;
      if (nabor < j) {
        int nm = meas[nabor];
        if (nm > 0) L.remove(nm, nabor);
        nm = ++meas[nabor];
        L.enter(nm, nabor);
      } else {
        ++meas[nabor];
      }
    }
    --num_left;
  }
  for (int step = 0; step < n && num_left > 0; step++) {          // :1245-1320, at most one C point per step
    const int index = head[L.maxm];
    if (index < 0 || index >= n) { L.bad = 2; break; }
    const int measure = meas[index];
    cf[index] = 1;
    meas[index] = 0;
    --num_left;
    L.remove(measure, index);
    for (int j = T_i[index]; j < T_i[index + 1]; j++) {           // the points that depend on the new C point become F
      const int nabor = T_j[j];
      if (cf[nabor] != 0) continue;
      cf[nabor] = -1;
      L.remove(meas[nabor], nabor);
      --num_left;
      for (int k = S_i[nabor]; k < S_i[nabor + 1]; k++) {         // ... and what they depend on gains a measure point
        const int n2 = S_j[k];
        if (cf[n2] != 0) continue;
        L.remove(meas[n2], n2);
        ++meas[n2];
        L.enter(meas[n2], n2);
      }
    }
    for (int j = S_i[index]; j < S_i[index + 1]; j++) {           // what the C point depends on loses a measure point
      const int nabor = S_j[j];
      if (cf[nabor] != 0) continue;
      int m2 = meas[nabor];
      L.remove(m2, nabor);
      meas[nabor] = --m2;
      if (m2 > 0) { L.enter(m2, nabor); continue; }
      cf[nabor] = -1;
      --num_left;
      for (int k = S_i[nabor]; k < S_i[nabor + 1]; k++) {
        const int n2 = S_j[k];
        if (cf[n2] != 0) continue;
        L.remove(meas[n2], n2);
        ++meas[n2];
        L.enter(meas[n2], n2);
      }
    }
  }
  for (int i = 0; i < n; i++)
    if (cf[i] == 3) cf[i] = 1;                                    // :1337-1343 SC_PT -> C_PT
  return L.bad ? L.bad : (num_left > 0 ? 3 : 0);
}
