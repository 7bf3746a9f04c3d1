/* mpi.h -- single-process stand-in for the handful of MPI calls the reference's examples and drivers make around the
 * solver (hypre's own sequential build has utilities/mpistubs.h for the library, but its examples include <mpi.h> through
 * HYPRE_utilities.h and expect a real MPI).  One rank, rank 0.  Used with `-include mpi.h` or found as <mpi.h> on the
 * include path; harmless next to the reference's `typedef HYPRE_Int MPI_Comm` (same type). */
#ifndef B200_MPI_STUB_H
#define B200_MPI_STUB_H
#include <stdlib.h>
#include <string.h>
#include <time.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif
#define MPI_SUCCESS 0
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_LONG 8
#define MPI_LONG_LONG_INT 9
#define MPI_FLOAT 5
#define MPI_DOUBLE 16
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
static inline int b200_mpi_size_of(MPI_Datatype t) { return t == MPI_CHAR ? 1 : (t == MPI_INT || t == MPI_FLOAT) ? 4 : 8; }
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return 0; }
static inline double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; if (s != r) memcpy(r, s, (size_t)n * b200_mpi_size_of(t)); return 0; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  (void)root; return MPI_Allreduce(s, r, n, t, op, c); }
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
#endif
