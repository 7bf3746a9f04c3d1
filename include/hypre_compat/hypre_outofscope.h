/* hypre_outofscope.h -- declarations of reference entry points that are OUT of this library's scope (other
 * preconditioners and Krylov drivers) but that unmodified reference programs mention: src/examples/ex5.c selects ParaSails,
 * FlexGMRES and ILU by command-line flag.  They are NOT implemented in libhypre_b200.so; a program that names them links
 * `-lhypre_b200 -lHYPRE` (the reference library after ours: first definition wins for everything on the hot path, the
 * reference serves the rest).  Objects do not cross: a matrix assembled by this library's HYPRE_IJMatrix* must not be
 * handed to these functions ("struct ABI").  Prototypes as in parcsr_ls/HYPRE_parcsr_ls.h:1467-1594,
 * :2385-2480 and krylov/HYPRE_krylov.h:75-77, :560-643. */
#ifndef B200_COMPAT_OUTOFSCOPE_H
#define B200_COMPAT_OUTOFSCOPE_H
#include "../HYPRE_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
#ifndef HYPRE_MODIFYPC
#define HYPRE_MODIFYPC
typedef HYPRE_Int (*HYPRE_PtrToModifyPCFcn)(HYPRE_Solver, HYPRE_Int, HYPRE_Real);
#endif
HYPRE_Int HYPRE_ParaSailsCreate(MPI_Comm comm, HYPRE_Solver *solver);
HYPRE_Int HYPRE_ParaSailsDestroy(HYPRE_Solver solver);
HYPRE_Int HYPRE_ParaSailsSetup(HYPRE_Solver solver, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x);
HYPRE_Int HYPRE_ParaSailsSolve(HYPRE_Solver solver, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x);
HYPRE_Int HYPRE_ParaSailsSetParams(HYPRE_Solver solver, HYPRE_Real thresh, HYPRE_Int nlevels);
HYPRE_Int HYPRE_ParaSailsSetFilter(HYPRE_Solver solver, HYPRE_Real filter);
HYPRE_Int HYPRE_ParaSailsSetSym(HYPRE_Solver solver, HYPRE_Int sym);
HYPRE_Int HYPRE_ParaSailsSetLogging(HYPRE_Solver solver, HYPRE_Int logging);
HYPRE_Int HYPRE_ParCSRFlexGMRESCreate(MPI_Comm comm, HYPRE_Solver *solver);
HYPRE_Int HYPRE_ParCSRFlexGMRESDestroy(HYPRE_Solver solver);
HYPRE_Int HYPRE_ParCSRFlexGMRESSetup(HYPRE_Solver solver, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x);
HYPRE_Int HYPRE_ParCSRFlexGMRESSolve(HYPRE_Solver solver, HYPRE_ParCSRMatrix A, HYPRE_ParVector b, HYPRE_ParVector x);
HYPRE_Int HYPRE_FlexGMRESSetKDim(HYPRE_Solver solver, HYPRE_Int k_dim);
HYPRE_Int HYPRE_FlexGMRESSetTol(HYPRE_Solver solver, HYPRE_Real tol);
HYPRE_Int HYPRE_FlexGMRESSetMaxIter(HYPRE_Solver solver, HYPRE_Int max_iter);
HYPRE_Int HYPRE_FlexGMRESSetPrintLevel(HYPRE_Solver solver, HYPRE_Int print_level);
HYPRE_Int HYPRE_FlexGMRESSetLogging(HYPRE_Solver solver, HYPRE_Int logging);
HYPRE_Int HYPRE_FlexGMRESSetPrecond(HYPRE_Solver solver, HYPRE_PtrToSolverFcn precond, HYPRE_PtrToSolverFcn precond_setup,
                                    HYPRE_Solver precond_solver);
HYPRE_Int HYPRE_FlexGMRESSetModifyPC(HYPRE_Solver solver, HYPRE_PtrToModifyPCFcn modify_pc);
HYPRE_Int HYPRE_FlexGMRESGetNumIterations(HYPRE_Solver solver, HYPRE_Int *num_iterations);
HYPRE_Int HYPRE_FlexGMRESGetFinalRelativeResidualNorm(HYPRE_Solver solver, HYPRE_Real *norm);
HYPRE_Int HYPRE_ILUDestroy(HYPRE_Solver solver);
#ifdef __cplusplus
}
#endif
#endif
