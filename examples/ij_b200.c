/* ij_b200.c -- a plain-C client of the reference's public API, linked against libhypre_b200.so.
 *
 * It walks the same call sequence as the reference driver test/ij.c does for
 *     ij -laplacian [-27pt] | -difconv [-a ax ay az] [-atype t] | -rotate -alpha A -eps E  -n nx ny nz [-c cx cy cz] -solver {0,1,2,3,4,9,10,-1} ...
 * (problem: ij.c:7785-7830 / :9084; rhs b = 1, x0 = 0: :2318-2340; AMG-PCG: :3884-4043, :4270-4330;
 * AMG alone: :3390-3560; AMG-GMRES: :5298-5480; AMG-BiCGSTAB: :6364-6500; matvec loop: :3206-3243) using only HYPRE_* calls, so it shows that host C
 * written against hypre's API runs on the B200 path by relinking.  It prints the result lines in the
 * driver's format ("Iterations = ", "Final Relative Residual Norm = ") so the tests can diff it
 * against the reference's own `ij` run with the same flags.
 *
 * With -ijbuild the operator is first assembled row by row through HYPRE_IJMatrixSetValues (the
 * examples/ex5.c pattern) instead of GenerateLaplacian; both give the same ParCSR matrix.
 *
 * build:  gcc -O2 -Iinclude examples/ij_b200.c -o examples/ij_b200 -Lhypre_ve_b200 -lhypre_b200 \
 *             -Wl,-rpath,'$ORIGIN/../hypre_ve_b200' -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "HYPRE_b200.h"

static HYPRE_ParCSRMatrix build_by_ij(int nx, int ny, int nz, const double *cv, HYPRE_IJMatrix *ij_out) {
  /* 7-point operator, entries inserted in the generator's order: centre, z-, y-, x-, x+, y+, z+ */
  const int n = nx * ny * nz;
  HYPRE_IJMatrix ij;
  HYPRE_IJMatrixCreate(MPI_COMM_WORLD, 0, n - 1, 0, n - 1, &ij);
  HYPRE_IJMatrixSetObjectType(ij, HYPRE_PARCSR);
  HYPRE_IJMatrixInitialize(ij);
  for (int iz = 0; iz < nz; iz++)
    for (int iy = 0; iy < ny; iy++)
      for (int ix = 0; ix < nx; ix++) {
        int row = (iz * ny + iy) * nx + ix, cols[7], k = 0;
        double vals[7];
        cols[k] = row; vals[k++] = cv[0];
        if (iz > 0) { cols[k] = row - nx * ny; vals[k++] = cv[3]; }
        if (iy > 0) { cols[k] = row - nx; vals[k++] = cv[2]; }
        if (ix > 0) { cols[k] = row - 1; vals[k++] = cv[1]; }
        if (ix + 1 < nx) { cols[k] = row + 1; vals[k++] = cv[1]; }
        if (iy + 1 < ny) { cols[k] = row + nx; vals[k++] = cv[2]; }
        if (iz + 1 < nz) { cols[k] = row + nx * ny; vals[k++] = cv[3]; }
        HYPRE_IJMatrixSetValues(ij, 1, &k, &row, cols, vals);
      }
  HYPRE_IJMatrixAssemble(ij);
  void *obj = NULL;
  HYPRE_IJMatrixGetObject(ij, &obj);
  *ij_out = ij;
  return (HYPRE_ParCSRMatrix)obj;
}

/* the seven values of the convection-diffusion stencil -cx Dxx - cy Dyy - cz Dzz + ax Dx + ay Dy + az Dz, as
 * BuildParDifConv computes them (ij.c:8266-8409): centre, x-, y-, z-, x+, y+, z+ */
static int sign_double(double a) { return (0.0 < a) - (0.0 > a); }
static void difconv_values(int nx, int ny, int nz, double cx, double cy, double cz, double ax, double ay, double az, int atype,
                           double *v) {
  const int n[3] = {nx, ny, nz};
  const double c[3] = {cx, cy, cz}, a[3] = {ax, ay, az};
  v[0] = 0.;
  for (int d = 0; d < 3; d++) {
    const double hin = 1. / (double)(n[d] + 1);
    if (atype == 0 || atype == 1 || atype == 3) {
      const int back = atype == 1 || (atype == 3 && sign_double(c[d]) * sign_double(a[d]) == 1);
      if (back) {                                         /* backward differences for the convection term */
        v[1 + d] = -c[d] / (hin * hin) - a[d] / hin;
        v[4 + d] = -c[d] / (hin * hin);
        if (n[d] > 1) v[0] += 2.0 * c[d] / (hin * hin) + 1. * a[d] / hin;
      } else {                                            /* forward differences */
        v[1 + d] = -c[d] / (hin * hin);
        v[4 + d] = -c[d] / (hin * hin) + a[d] / hin;
        if (n[d] > 1) v[0] += 2.0 * c[d] / (hin * hin) - 1. * a[d] / hin;
      }
    } else {                                              /* centred differences */
      v[1 + d] = -c[d] / (hin * hin) - a[d] / (2. * hin);
      v[4 + d] = -c[d] / (hin * hin) + a[d] / (2. * hin);
      if (n[d] > 1) v[0] += 2.0 * c[d] / (hin * hin);
    }
  }
}

int main(int argc, char **argv) {
  int nx = 10, ny = 10, nz = 10, solver_id = 1, stencil27 = 0, ijbuild = 0, difconv = 0, atype = 0, k_dim = 5, rotate = 0, print_system = 0;
  const char *fromfile = NULL;                          /* ij -fromfile NAME: HYPRE_IJMatrixRead(NAME.00000) (ij.c:7504-7540) */
  double cx = 1., cy = 1., cz = 1., ax = 1., ay = 1., az = 1., alpha = 0., eps = 1.;
  /* driver defaults, test/ij.c:203-330 and :1181-1205 */
  int coarsen_type = 10, interp_type = 6, P_max_elmts = 4, relax_type = -1, relax_order = 0, max_levels = 25;
  int agg_num_levels = 0, rap2 = 0, mod_rap2 = 0, keepTranspose = 1, num_sweeps = 1, max_iter = 1000, mg_max_iter = 100;
  int coarse_threshold = 9, min_coarse_size = 0, ioutdat = 3, poutdat = 1, two_norm = 1, gs_blocks = 1;
  int cycle_type = 1, fcycle = 0, ns_coarse = 1, ns_down = -1, ns_up = -1;                        /* ij.c:167, :205-206 */
  double strong_threshold = 0.25, max_row_sum = 1.0, trunc_factor = 0.0, tol = 1.e-8, pc_tol = 0., relax_wt = 1., outer_wt = 1.;
  for (int a = 1; a < argc; a++) if (!strcmp(argv[a], "-rotate")) rotate = 1;      /* decides how many numbers -n takes */
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "-laplacian")) ;
    else if (!strcmp(argv[a], "-27pt")) stencil27 = 1;
    else if (!strcmp(argv[a], "-fromfile") && a + 1 < argc) fromfile = argv[++a];
    else if (!strcmp(argv[a], "-print")) print_system = 1;               /* ij.c:3169-3188: IJ.out.A, IJ.out.b, IJ.out.x0 */
    else if (!strcmp(argv[a], "-difconv")) difconv = 1;
    else if (!strcmp(argv[a], "-rotate")) rotate = 1;                    /* 2-D: -n nx ny (ij.c:9136-9160) */
    else if (!strcmp(argv[a], "-alpha") && a + 1 < argc) alpha = atof(argv[++a]);
    else if (!strcmp(argv[a], "-eps") && a + 1 < argc) eps = atof(argv[++a]);
    else if (!strcmp(argv[a], "-a") && a + 3 < argc) { ax = atof(argv[a + 1]); ay = atof(argv[a + 2]); az = atof(argv[a + 3]); a += 3; }
    else if (!strcmp(argv[a], "-atype") && a + 1 < argc) atype = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-k") && a + 1 < argc) k_dim = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-ijbuild")) ijbuild = 1;
    else if (!strcmp(argv[a], "-n") && rotate && a + 2 < argc) { nx = atoi(argv[a + 1]); ny = atoi(argv[a + 2]); nz = 1; a += 2; }
    else if (!strcmp(argv[a], "-n") && a + 3 < argc) { nx = atoi(argv[a + 1]); ny = atoi(argv[a + 2]); nz = atoi(argv[a + 3]); a += 3; }
    else if (!strcmp(argv[a], "-c") && a + 3 < argc) { cx = atof(argv[a + 1]); cy = atof(argv[a + 2]); cz = atof(argv[a + 3]); a += 3; }
    else if (!strcmp(argv[a], "-solver") && a + 1 < argc) solver_id = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-pmis")) coarsen_type = 8;
    else if (!strcmp(argv[a], "-hmis")) coarsen_type = 10;
    else if (!strcmp(argv[a], "-rlx") && a + 1 < argc) relax_type = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-rlx_order") && a + 1 < argc) relax_order = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-w") && a + 1 < argc) relax_wt = atof(argv[++a]);
    else if (!strcmp(argv[a], "-interptype") && a + 1 < argc) interp_type = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-Pmx") && a + 1 < argc) P_max_elmts = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-tr") && a + 1 < argc) trunc_factor = atof(argv[++a]);
    else if (!strcmp(argv[a], "-th") && a + 1 < argc) strong_threshold = atof(argv[++a]);
    else if (!strcmp(argv[a], "-mxrs") && a + 1 < argc) max_row_sum = atof(argv[++a]);
    else if (!strcmp(argv[a], "-agg_nl") && a + 1 < argc) agg_num_levels = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-rap") && a + 1 < argc) rap2 = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-mod_rap2") && a + 1 < argc) mod_rap2 = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-keepT") && a + 1 < argc) keepTranspose = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-ns") && a + 1 < argc) num_sweeps = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-ns_coarse") && a + 1 < argc) ns_coarse = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-ns_down") && a + 1 < argc) ns_down = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-ns_up") && a + 1 < argc) ns_up = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-mu") && a + 1 < argc) cycle_type = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-fmg")) fcycle = 1;
    else if (!strcmp(argv[a], "-mxl") && a + 1 < argc) max_levels = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-max_iter") && a + 1 < argc) max_iter = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-mg_max_iter") && a + 1 < argc) mg_max_iter = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-tol") && a + 1 < argc) tol = atof(argv[++a]);
    else if (!strcmp(argv[a], "-iout") && a + 1 < argc) ioutdat = atoi(argv[++a]);
    else if (!strcmp(argv[a], "-gs_blocks") && a + 1 < argc) gs_blocks = atoi(argv[++a]);   /* = OMP_NUM_THREADS of the reference run */
    else { fprintf(stderr, "ij_b200: unknown option %s\n", argv[a]); return 2; }
  }
  if (HYPRE_Init()) { fprintf(stderr, "ij_b200: HYPRE_Init failed (no B200 / CUDA device?)\n"); return 1; }

  /* operator */
  double values[7];
  HYPRE_ParCSRMatrix A;
  HYPRE_IJMatrix ij_A = NULL;
  if (fromfile) {
    if (HYPRE_IJMatrixRead(fromfile, MPI_COMM_WORLD, HYPRE_PARCSR, &ij_A)) { fprintf(stderr, "ij_b200: could not read %s.00000\n", fromfile); return 1; }
    void *o = NULL;
    HYPRE_IJMatrixGetObject(ij_A, &o);
    A = (HYPRE_ParCSRMatrix)o;
  } else if (stencil27) {
    values[0] = 26.0; values[1] = -1.0;                                  /* ij.c:9063-9071 */
    if (nx == 1 || ny == 1 || nz == 1) values[0] = 8.0;
    if (nx * ny == 1 || nx * nz == 1 || ny * nz == 1) values[0] = 2.0;
    A = GenerateLaplacian27pt(MPI_COMM_WORLD, nx, ny, nz, 1, 1, 1, 0, 0, 0, values);
  } else if (rotate) {
    A = GenerateRotate7pt(MPI_COMM_WORLD, nx, ny, 1, 1, 0, 0, alpha, eps);  /* ij.c:9198 */
  } else if (difconv) {
    difconv_values(nx, ny, nz, cx, cy, cz, ax, ay, az, atype, values);  /* ij.c:8266-8409 */
    A = GenerateDifConv(MPI_COMM_WORLD, nx, ny, nz, 1, 1, 1, 0, 0, 0, values);
  } else {
    values[1] = -cx; values[2] = -cy; values[3] = -cz;                   /* ij.c:7789-7806 */
    values[0] = 0.;
    if (nx > 1) values[0] += 2.0 * cx;
    if (ny > 1) values[0] += 2.0 * cy;
    if (nz > 1) values[0] += 2.0 * cz;
    A = ijbuild ? build_by_ij(nx, ny, nz, values, &ij_A)
                : GenerateLaplacian(MPI_COMM_WORLD, nx, ny, nz, 1, 1, 1, 0, 0, 0, values);
  }
  if (!A) { fprintf(stderr, "ij_b200: could not build the operator\n"); return 1; }
  HYPRE_BigInt M, N;
  HYPRE_ParCSRMatrixGetDims(A, &M, &N);
  if (fromfile) printf("  FromFile: %s  rows = %d\n", fromfile, M);
  else printf("  %s%s:   (nx, ny, nz) = (%d, %d, %d)  rows = %d\n", rotate ? "Rotate 7pt" : difconv && !stencil27 ? "Convection-Diffusion" : "Laplacian",
         stencil27 ? " 27pt" : "", nx, ny, nz, M);

  /* rhs = 1, x0 = 0 (the driver's default build_rhs_type 2 / build_x0_type) */
  HYPRE_IJVector ij_b, ij_x;
  HYPRE_ParVector b, x;
  void *obj;
  HYPRE_IJVectorCreate(MPI_COMM_WORLD, 0, M - 1, &ij_b);
  HYPRE_IJVectorSetObjectType(ij_b, HYPRE_PARCSR);
  HYPRE_IJVectorInitialize(ij_b);
  double *ones = (double *)malloc(sizeof(double) * (size_t)M);
  for (int i = 0; i < M; i++) ones[i] = 1.0;
  HYPRE_IJVectorSetValues(ij_b, M, NULL, ones);
  HYPRE_IJVectorAssemble(ij_b);
  HYPRE_IJVectorGetObject(ij_b, &obj);
  b = (HYPRE_ParVector)obj;
  HYPRE_IJVectorCreate(MPI_COMM_WORLD, 0, N - 1, &ij_x);
  HYPRE_IJVectorSetObjectType(ij_x, HYPRE_PARCSR);
  HYPRE_IJVectorInitialize(ij_x);
  HYPRE_IJVectorAssemble(ij_x);
  HYPRE_IJVectorGetObject(ij_x, &obj);
  x = (HYPRE_ParVector)obj;

  if (print_system) {
    if (ij_A) HYPRE_IJMatrixPrint(ij_A, "IJ.out.A"); else hypre_ParCSRMatrixPrintIJ(A, 0, 0, "IJ.out.A");
    HYPRE_IJVectorPrint(ij_b, "IJ.out.b");
    HYPRE_IJVectorPrint(ij_x, "IJ.out.x0");
  }
  int num_iterations = 0;
  double final_res_norm = 0.;
  if (solver_id == -1) {                                                   /* ij.c:3206-3243 */
    HYPRE_ParCSRMatrixMatvec(1., A, b, 0., x);
    for (int i = 0; i < 100; i++) HYPRE_ParCSRMatrixMatvec(1., A, b, 0., x);
    double xx = 0;
    HYPRE_ParVectorInnerProd(x, x, &xx);
    printf("Matvec x 100 done, <Ab, Ab> = %.15e\n", xx);
  } else if (solver_id == 0 || solver_id == 1 || solver_id == 3 || solver_id == 9) {
    HYPRE_Solver amg, pcg = NULL;
    HYPRE_BoomerAMGCreate(&amg);
    HYPRE_BoomerAMGSetInterpType(amg, interp_type);
    HYPRE_BoomerAMGSetCoarsenType(amg, coarsen_type);
    HYPRE_BoomerAMGSetStrongThreshold(amg, strong_threshold);
    HYPRE_BoomerAMGSetMaxCoarseSize(amg, coarse_threshold);
    HYPRE_BoomerAMGSetMinCoarseSize(amg, min_coarse_size);
    HYPRE_BoomerAMGSetTruncFactor(amg, trunc_factor);
    HYPRE_BoomerAMGSetPMaxElmts(amg, P_max_elmts);
    HYPRE_BoomerAMGSetPrintLevel(amg, solver_id == 0 ? ioutdat : poutdat);
    HYPRE_BoomerAMGSetCycleType(amg, cycle_type);
    HYPRE_BoomerAMGSetFCycle(amg, fcycle);
    HYPRE_BoomerAMGSetNumSweeps(amg, num_sweeps);
    HYPRE_BoomerAMGSetCycleNumSweeps(amg, ns_coarse, 3);
    if (solver_id == 0) {                                                  /* ij.c:3494-3502: only the AMG solver takes these */
      if (ns_down > -1) HYPRE_BoomerAMGSetCycleNumSweeps(amg, ns_down, 1);
      if (ns_up > -1) HYPRE_BoomerAMGSetCycleNumSweeps(amg, ns_up, 2);
    }
    if (relax_type > -1) HYPRE_BoomerAMGSetRelaxType(amg, relax_type);
    HYPRE_BoomerAMGSetRelaxOrder(amg, relax_order);
    HYPRE_BoomerAMGSetRelaxWt(amg, relax_wt);
    HYPRE_BoomerAMGSetOuterWt(amg, outer_wt);
    HYPRE_BoomerAMGSetMaxLevels(amg, max_levels);
    HYPRE_BoomerAMGSetMaxRowSum(amg, max_row_sum);
    HYPRE_BoomerAMGSetNumFunctions(amg, 1);
    HYPRE_BoomerAMGSetAggNumLevels(amg, agg_num_levels);
    HYPRE_BoomerAMGSetRAP2(amg, rap2);
    HYPRE_BoomerAMGSetModuleRAP2(amg, mod_rap2);
    HYPRE_BoomerAMGSetKeepTranspose(amg, keepTranspose);
    HYPRE_b200_BoomerAMGSetGSBlocks(amg, gs_blocks);
    if (solver_id == 0) {
      HYPRE_BoomerAMGSetTol(amg, tol);
      HYPRE_BoomerAMGSetMaxIter(amg, mg_max_iter);
      printf("Solver:  AMG\n");
      HYPRE_BoomerAMGSetup(amg, A, b, x);
      HYPRE_BoomerAMGSolve(amg, A, b, x);
      HYPRE_BoomerAMGGetNumIterations(amg, &num_iterations);
      HYPRE_BoomerAMGGetFinalRelativeResidualNorm(amg, &final_res_norm);
      printf("\nBoomerAMG Iterations = %d\n", num_iterations);
    } else if (solver_id == 3) {                                            /* ij.c:5298-5480 */
      HYPRE_ParCSRGMRESCreate(MPI_COMM_WORLD, &pcg);
      HYPRE_GMRESSetKDim(pcg, k_dim);
      HYPRE_GMRESSetMaxIter(pcg, max_iter);
      HYPRE_GMRESSetTol(pcg, tol);
      HYPRE_GMRESSetAbsoluteTol(pcg, 0.);
      HYPRE_GMRESSetLogging(pcg, 1);
      HYPRE_GMRESSetPrintLevel(pcg, ioutdat);
      HYPRE_GMRESSetRelChange(pcg, 0);
      printf("Solver: AMG-GMRES\n");
      HYPRE_BoomerAMGSetTol(amg, pc_tol);
      HYPRE_BoomerAMGSetMaxIter(amg, 1);
      HYPRE_GMRESSetMaxIter(pcg, mg_max_iter);
      HYPRE_GMRESSetPrecond(pcg, (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSolve, (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSetup, amg);
      HYPRE_Solver got = NULL;
      HYPRE_GMRESGetPrecond(pcg, &got);
      if (got != amg) { printf("HYPRE_GMRESGetPrecond got bad precond\n"); return -1; }
      HYPRE_GMRESSetup(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
      HYPRE_GMRESSolve(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
      HYPRE_GMRESGetNumIterations(pcg, &num_iterations);
      HYPRE_GMRESGetFinalRelativeResidualNorm(pcg, &final_res_norm);
      double ts = 0, tv = 0;
      HYPRE_b200_KrylovGetTimes(pcg, &ts, &tv);
      printf("GMRES Setup: device time = %f seconds\nGMRES Solve: device time = %f seconds\n", ts, tv);
      printf("\nGMRES Iterations = %d\n", num_iterations);
      HYPRE_ParCSRGMRESDestroy(pcg);
    } else if (solver_id == 9) {                                            /* ij.c:6364-6500 */
      HYPRE_ParCSRBiCGSTABCreate(MPI_COMM_WORLD, &pcg);
      HYPRE_BiCGSTABSetMaxIter(pcg, max_iter);
      HYPRE_BiCGSTABSetTol(pcg, tol);
      HYPRE_BiCGSTABSetAbsoluteTol(pcg, 0.);
      HYPRE_BiCGSTABSetLogging(pcg, ioutdat);
      HYPRE_BiCGSTABSetPrintLevel(pcg, ioutdat);
      printf("Solver: AMG-BiCGSTAB\n");
      HYPRE_BoomerAMGSetTol(amg, pc_tol);
      HYPRE_BoomerAMGSetMaxIter(amg, 1);
      HYPRE_BiCGSTABSetPrecond(pcg, (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSolve, (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSetup, amg);
      HYPRE_BiCGSTABSetup(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
      HYPRE_BiCGSTABSolve(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
      HYPRE_BiCGSTABGetNumIterations(pcg, &num_iterations);
      HYPRE_BiCGSTABGetFinalRelativeResidualNorm(pcg, &final_res_norm);
      printf("\nBiCGSTAB Iterations = %d\n", num_iterations);
      HYPRE_ParCSRBiCGSTABDestroy(pcg);
    } else {
      HYPRE_ParCSRPCGCreate(MPI_COMM_WORLD, &pcg);
      HYPRE_PCGSetMaxIter(pcg, max_iter);
      HYPRE_PCGSetTol(pcg, tol);
      HYPRE_PCGSetTwoNorm(pcg, two_norm);
      HYPRE_PCGSetRelChange(pcg, 0);
      HYPRE_PCGSetPrintLevel(pcg, ioutdat);
      HYPRE_PCGSetAbsoluteTol(pcg, 0.);
      HYPRE_PCGSetRecomputeResidual(pcg, 0);
      printf("Solver: AMG-PCG\n");
      HYPRE_BoomerAMGSetTol(amg, pc_tol);
      HYPRE_BoomerAMGSetMaxIter(amg, 1);
      HYPRE_PCGSetMaxIter(pcg, mg_max_iter);
      HYPRE_PCGSetPrecond(pcg, (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSolve, (HYPRE_PtrToSolverFcn)HYPRE_BoomerAMGSetup, amg);
      HYPRE_Solver got = NULL;
      HYPRE_PCGGetPrecond(pcg, &got);
      if (got != amg) { printf("HYPRE_ParCSRPCGGetPrecond got bad precond\n"); return -1; }
      HYPRE_PCGSetup(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
      HYPRE_PCGSolve(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
      HYPRE_PCGGetNumIterations(pcg, &num_iterations);
      HYPRE_PCGGetFinalRelativeResidualNorm(pcg, &final_res_norm);
      double ts = 0, tv = 0;
      HYPRE_b200_PCGGetTimes(pcg, &ts, &tv);
      printf("PCG Setup: device time = %f seconds\nPCG Solve: device time = %f seconds\n", ts, tv);
      printf("\nIterations = %d\n", num_iterations);
      HYPRE_ParCSRPCGDestroy(pcg);
    }
    int nl = 0;
    HYPRE_b200_BoomerAMGGetNumLevels(amg, &nl);
    for (int l = 0; l < nl; l++) {
      int rows = 0, nnz = 0;
      HYPRE_b200_BoomerAMGGetLevelSize(amg, l, &rows, &nnz);
      printf("level %2d rows %10d nnz %12d\n", l, rows, nnz);
    }
    printf("Final Relative Residual Norm = %e\n", final_res_norm);
    HYPRE_BoomerAMGDestroy(amg);
  } else if (solver_id == 2) {
    HYPRE_Solver pcg;
    HYPRE_ParCSRPCGCreate(MPI_COMM_WORLD, &pcg);
    HYPRE_PCGSetMaxIter(pcg, max_iter);
    HYPRE_PCGSetTol(pcg, tol);
    HYPRE_PCGSetTwoNorm(pcg, two_norm);
    HYPRE_PCGSetPrintLevel(pcg, ioutdat);
    printf("Solver: DS-PCG\n");
    HYPRE_PCGSetPrecond(pcg, (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScale, (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScaleSetup, NULL);
    HYPRE_PCGSetup(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
    HYPRE_PCGSolve(pcg, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
    HYPRE_PCGGetNumIterations(pcg, &num_iterations);
    HYPRE_PCGGetFinalRelativeResidualNorm(pcg, &final_res_norm);
    printf("\nIterations = %d\nFinal Relative Residual Norm = %e\n", num_iterations, final_res_norm);
    HYPRE_ParCSRPCGDestroy(pcg);
  } else if (solver_id == 4) {                                             /* ij.c:5481-5491 */
    HYPRE_Solver gm;
    HYPRE_ParCSRGMRESCreate(MPI_COMM_WORLD, &gm);
    HYPRE_GMRESSetKDim(gm, k_dim);
    HYPRE_GMRESSetMaxIter(gm, max_iter);
    HYPRE_GMRESSetTol(gm, tol);
    HYPRE_GMRESSetAbsoluteTol(gm, 0.);
    HYPRE_GMRESSetLogging(gm, 1);
    HYPRE_GMRESSetPrintLevel(gm, ioutdat);
    HYPRE_GMRESSetRelChange(gm, 0);
    printf("Solver: DS-GMRES\n");
    HYPRE_GMRESSetPrecond(gm, (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScale, (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScaleSetup, NULL);
    HYPRE_GMRESSetup(gm, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
    HYPRE_GMRESSolve(gm, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
    HYPRE_GMRESGetNumIterations(gm, &num_iterations);
    HYPRE_GMRESGetFinalRelativeResidualNorm(gm, &final_res_norm);
    printf("\nGMRES Iterations = %d\nFinal Relative Residual Norm = %e\n", num_iterations, final_res_norm);
    HYPRE_ParCSRGMRESDestroy(gm);
  } else if (solver_id == 10) {                                            /* ij.c:6481-6491 */
    HYPRE_Solver bi;
    HYPRE_ParCSRBiCGSTABCreate(MPI_COMM_WORLD, &bi);
    HYPRE_BiCGSTABSetMaxIter(bi, max_iter);
    HYPRE_BiCGSTABSetTol(bi, tol);
    HYPRE_BiCGSTABSetAbsoluteTol(bi, 0.);
    HYPRE_BiCGSTABSetLogging(bi, ioutdat);
    HYPRE_BiCGSTABSetPrintLevel(bi, ioutdat);
    printf("Solver: DS-BiCGSTAB\n");
    HYPRE_BiCGSTABSetPrecond(bi, (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScale, (HYPRE_PtrToSolverFcn)HYPRE_ParCSRDiagScaleSetup, NULL);
    HYPRE_BiCGSTABSetup(bi, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
    HYPRE_BiCGSTABSolve(bi, (HYPRE_Matrix)A, (HYPRE_Vector)b, (HYPRE_Vector)x);
    HYPRE_BiCGSTABGetNumIterations(bi, &num_iterations);
    HYPRE_BiCGSTABGetFinalRelativeResidualNorm(bi, &final_res_norm);
    printf("\nBiCGSTAB Iterations = %d\nFinal Relative Residual Norm = %e\n", num_iterations, final_res_norm);
    HYPRE_ParCSRBiCGSTABDestroy(bi);
  } else {
    fprintf(stderr, "ij_b200: solver %d is not on the B200 path (0 AMG, 1 AMG-PCG, 2 DS-PCG, 3 AMG-GMRES, 4 DS-GMRES, 9 AMG-BiCGSTAB, 10 DS-BiCGSTAB, -1 matvec)\n", solver_id);
    return 2;
  }
  if (print_system && solver_id >= 0) HYPRE_IJVectorPrint(ij_x, "IJ.out.x");      /* ij.c:7440-7443 */
  /* read a few solution values back through the IJ interface, as examples/ex5.c does */
  if (solver_id >= 0) {
    int idx[2];
    double xv[2];
    idx[0] = 0; idx[1] = M - 1;
    HYPRE_IJVectorGetValues(ij_x, 2, idx, xv);
    printf("x[0] = %.15e  x[last] = %.15e\n", xv[0], xv[1]);
  }
  free(ones);
  HYPRE_IJVectorDestroy(ij_b);
  HYPRE_IJVectorDestroy(ij_x);
  if (ij_A) HYPRE_IJMatrixDestroy(ij_A); else HYPRE_ParCSRMatrixDestroy(A);
  const int e = HYPRE_GetError();
  if (e) {
    char msg[128] = "";
    HYPRE_DescribeError(e, msg);
    printf("hypre error flag = %d %s\n", e, msg);
  }
  HYPRE_Finalize();
  return e ? 1 : 0;
}
