/* HYPRE_utilities.h -- drop-in for the reference's public header of the same name (hypre 2.20: src/HYPRE.h, krylov/HYPRE_krylov.h,
 * parcsr_ls/HYPRE_parcsr_ls.h, IJ_mv/HYPRE_IJ_mv.h, utilities/HYPRE_utilities.h, parcsr_mv/HYPRE_parcsr_mv.h,
 * seq_mv/HYPRE_seq_mv.h).  A program written against hypre's headers compiles UNCHANGED with -Iinclude/hypre_compat and
 * links against libhypre_b200.so for the BoomerAMG / PCG / GMRES / BiCGSTAB / IJ path; everything is declared in one place. */
#ifndef B200_COMPAT_HYPRE_utilities_H
#define B200_COMPAT_HYPRE_utilities_H
#include "../HYPRE_b200.h"
#endif
