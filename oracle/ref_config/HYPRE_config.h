/* Build configuration for compiling the reference (hypre 2.20.0, SX-Aurora fork)
 * CPU sources IN PLACE from /root/reference/src, without running its configure.
 * Mirrors what `./configure --without-MPI --with-openmp --disable-fortran`
 * generates (reference template: src/config/HYPRE_config.h.in).
 * TEST INFRASTRUCTURE ONLY: used by oracle/build_ref.py, never by the product. */
#ifndef HYPRE_B200_REF_CONFIG_H
#define HYPRE_B200_REF_CONFIG_H
#define HYPRE_RELEASE_NAME "hypre"
#define HYPRE_RELEASE_VERSION "2.20.0"
#define HYPRE_RELEASE_NUMBER 22000
#define HYPRE_RELEASE_DATE "2020/09/24"
#define HYPRE_RELEASE_TIME "00:00:00"
#define HYPRE_RELEASE_BUGS "https://github.com/hypre-space/hypre/issues"
#define HYPRE_MAXDIM 3
#define HYPRE_NO_GLOBAL_PARTITION 1
#define HYPRE_SEQUENTIAL 1
#define HYPRE_USING_HOST_MEMORY 1
#define HYPRE_USING_HYPRE_BLAS 1
#define HYPRE_USING_HYPRE_LAPACK 1
#define HYPRE_USING_OPENMP 1
#define HYPRE_LINUX 1
#define HYPRE_FMANGLE 0
#define HYPRE_FMANGLE_BLAS 0
#define HYPRE_FMANGLE_LAPACK 0
/* oracle/build_ref.py --bigint: the reference's own --enable-bigint configuration (64-bit HYPRE_Int / HYPRE_BigInt).
 * One process holds the whole problem here (no MPI), and at 512^3 unknowns the local nonzero counters pass 2^31: the default
 * 32-bit build of the reference segfaults on `ij -n 512 512 512`.  Same algorithms, same random numbers, same results. */
#ifdef B200_REF_BIGINT
#define HYPRE_BIGINT 1
#endif
#endif
