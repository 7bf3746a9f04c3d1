/* ref_dump.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small driver over the UNMODIFIED reference library (oracle/_ref/libHYPRE_ref.so,
 * compiled in place from /root/reference/src by oracle/build_ref.py).  It runs the
 * same call sequence as the reference driver for `ij -laplacian ... -solver 1`
 * (src/test/ij.c:3891-4010 create/set, :4013-4050 setup/solve) through the public
 * HYPRE_* API and then dumps, for parity tests, what the reference built:
 *   per level l:  A_l (diag CSR), CF_marker_l, S_l (strength, recomputed with the
 *                 reference's hypre_BoomerAMGCreateS), P_l (diag CSR), l1 norms,
 *   and the PCG residual history, iteration count and final relative residual.
 *
 * Output: a flat binary stream of records
 *     [u32 namelen][name][u32 dtype 0=i32 1=f64][u64 count][payload]
 * read by tests/refio.py.
 *
 * Flags follow ij.c's spelling: -n nx ny nz, -27pt, -c cx cy cz, -pmis, -rlx T,
 * -Pmx K, -agg_nl L, -mod_rap2 B, -keepT B, -th theta, -tol t, -interptype I,
 * -mxrs r (max_row_sum), -o FILE, -matvec K (time K SpMVs, ij -solver -1 analogue),
 * -rotate -alpha A -eps E (GenerateRotate7pt, 2-D: -n nx ny 1),
 * -ijbuild MODE [-noamg] (operator re-assembled through the reference's HYPRE_IJMatrix interface),
 * -solver 1|3|9 (AMG-PCG, AMG-GMRES with -k K, AMG-BiCGSTAB: ij.c:5298-5330, :6364-6380),
 * -difconv [-a ax ay az] [-atype T] (GenerateDifConv, nonsymmetric), -nodump (timing only), -ns / -ns_coarse / -mu / -fmg (cycle shape), -perturb SEED (non-Laplacian values, see below).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "_hypre_utilities.h"
#include "HYPRE.h"
#include "HYPRE_parcsr_mv.h"
#include "HYPRE_parcsr_ls.h"
#include "_hypre_parcsr_mv.h"
#include "_hypre_parcsr_ls.h"
#include "HYPRE_krylov.h"
#include "HYPRE_IJ_mv.h"

static FILE *g_out;

static void put(const char *name, int dtype, const void *p, size_t n)
{
   unsigned int nl = (unsigned int) strlen(name), dt = (unsigned int) dtype;
   unsigned long long cnt = n;
   if (!g_out) return;
   fwrite(&nl, 4, 1, g_out); fwrite(name, 1, nl, g_out);
   fwrite(&dt, 4, 1, g_out); fwrite(&cnt, 8, 1, g_out);
   if (n) fwrite(p, dtype ? 8 : 4, n, g_out);
}
static void put_csr(const char *pre, int l, hypre_CSRMatrix *M, int with_data)
{
   char nm[64];
   int n = hypre_CSRMatrixNumRows(M);
   int nnz = hypre_CSRMatrixI(M)[n];
   int dims[3] = { n, hypre_CSRMatrixNumCols(M), nnz };
   sprintf(nm, "%s%d.dims", pre, l); put(nm, 0, dims, 3);
   sprintf(nm, "%s%d.i", pre, l);    put(nm, 0, hypre_CSRMatrixI(M), n + 1);
   sprintf(nm, "%s%d.j", pre, l);    put(nm, 0, hypre_CSRMatrixJ(M), nnz);
   if (with_data) { sprintf(nm, "%s%d.a", pre, l); put(nm, 1, hypre_CSRMatrixData(M), nnz); }
}

/* -perturb SEED: a non-Laplacian SPD test operator on the stencil's pattern -- symmetric pseudo-random off-diagonal
 * magnitudes in [0.05, 1.5] (weak and strong connections), one in sixteen with a POSITIVE sign, strictly dominant
 * diagonal.  The same function lives in oracle/ref_dump.c and oracle/amg_oracle.c (test infrastructure). */
static void perturb_operator(int n, const int *I, const int *J, double *a, unsigned seed)
{
   int i, k;
   for (i = 0; i < n; i++)
   {
      double sum = 0.0;
      for (k = I[i] + 1; k < I[i + 1]; k++)
      {
         unsigned lo = (unsigned) (i < J[k] ? i : J[k]), hi = (unsigned) (i < J[k] ? J[k] : i);
         unsigned h = (lo * 73856093u) ^ (hi * 19349663u) ^ (seed * 83492791u);
         h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
         double f = 0.05 + 1.45 * (double) (h & 0xffffu) / 65535.0;
         a[k] = (((h >> 16) & 15u) == 0u) ? 0.25 * f : -f;
         sum += fabs(a[k]);
      }
      a[I[i]] = sum + 0.05;
   }
}

/* values[7] of the convection-diffusion stencil -cx Dxx - cy Dyy - cz Dzz + ax Dx + ay Dy + az Dz, computed as the
 * reference driver does (src/test/ij.c:8266-8409, BuildParDifConv): centre, x-, y-, z-, x+, y+, z+;
 * atype 0 forward, 1 backward, 3 upwind, else centred differences for the convection term. */
static int sign_double(double a) { return (0.0 < a) - (0.0 > a); }
static void difconv_values(int nx, int ny, int nz, const double *c, const double *a, int atype, double *v)
{
   int n[3] = { nx, ny, nz }, d;
   v[0] = 0.;
   for (d = 0; d < 3; d++)
   {
      double hin = 1. / (double) (n[d] + 1);
      int back = atype == 1 || (atype == 3 && sign_double(c[d]) * sign_double(a[d]) == 1);
      if (atype == 0 || atype == 1 || atype == 3)
      {
         if (back)
         {
            v[1 + d] = -c[d] / (hin * hin) - a[d] / hin;
            v[4 + d] = -c[d] / (hin * hin);
            if (n[d] > 1) v[0] += 2.0 * c[d] / (hin * hin) + 1. * a[d] / hin;
         }
         else
         {
            v[1 + d] = -c[d] / (hin * hin);
            v[4 + d] = -c[d] / (hin * hin) + a[d] / hin;
            if (n[d] > 1) v[0] += 2.0 * c[d] / (hin * hin) - 1. * a[d] / hin;
         }
      }
      else
      {
         v[1 + d] = -c[d] / (hin * hin) - a[d] / (2. * hin);
         v[4 + d] = -c[d] / (hin * hin) + a[d] / (2. * hin);
         if (n[d] > 1) v[0] += 2.0 * c[d] / (hin * hin);
      }
   }
}

/* -ijbuild MODE: the operator is re-assembled from the generated one through the reference's own HYPRE_IJMatrix
 * interface by a fixed stream of SetValues / AddToValues calls (mirrored record for record by tests/ijstream.py):
 *   phase 1  SetValues, two rows per call, rows visited from the last to the first, each row's entries rotated by
 *            (row mod length), half of every value;
 *   phase 2  AddToValues, one row per call, first to last, rotated by ((row+1) mod length), the other half;
 *   MODE 2 only, phase 3 (every row with row mod 7 == 3): one SetValues call holding a far column TWICE and an existing
 *            column (in-call duplicates stay duplicates, IJMatrix_parcsr.c:941-962), then one AddToValues call that
 *            lists the same row twice with one more column (the second block finds what the first appended);
 *   Assemble; then, after assembly, AddToValues(+1) on the diagonal of every third row and SetValues of the second
 *   stored entry of every fifth row to itself times 1 (existing entries only, :727-905); Assemble again. */
static HYPRE_ParCSRMatrix ijbuild(HYPRE_ParCSRMatrix G, int mode, HYPRE_IJMatrix *ij_out)
{
   hypre_CSRMatrix *D = hypre_ParCSRMatrixDiag((hypre_ParCSRMatrix *) G);
   int N = hypre_CSRMatrixNumRows(D), *I = hypre_CSRMatrixI(D), *J = hypre_CSRMatrixJ(D), r, k, q;
   double *a = hypre_CSRMatrixData(D);
   HYPRE_IJMatrix ij;
   int rows[2], ncols[2], cols[64];
   double vals[64];
   HYPRE_IJMatrixCreate(hypre_MPI_COMM_WORLD, 0, N - 1, 0, N - 1, &ij);
   HYPRE_IJMatrixSetObjectType(ij, HYPRE_PARCSR);
   HYPRE_IJMatrixInitialize(ij);
   for (r = N - 1; r >= 0; r -= 2)                       /* phase 1 */
   {
      int nr = 0, at = 0;
      for (q = 0; q < 2 && r - q >= 0; q++)
      {
         int row = r - q, len = I[row + 1] - I[row], rot = row % len;
         rows[nr] = row; ncols[nr++] = len;
         for (k = 0; k < len; k++) { int e = I[row] + (k + rot) % len; cols[at] = J[e]; vals[at++] = 0.5 * a[e]; }
      }
      HYPRE_IJMatrixSetValues(ij, nr, ncols, rows, cols, vals);
   }
   for (r = 0; r < N; r++)                               /* phase 2 */
   {
      int len = I[r + 1] - I[r], rot = (r + 1) % len;
      rows[0] = r; ncols[0] = len;
      for (k = 0; k < len; k++) { int e = I[r] + (k + rot) % len; cols[k] = J[e]; vals[k] = 0.5 * a[e]; }
      HYPRE_IJMatrixAddToValues(ij, 1, ncols, rows, cols, vals);
   }
   if (mode == 2)
      for (r = 3; r < N; r += 7)                         /* phase 3 */
      {
         int far = (int) (((long long) r * 31 + 17) % N), c2 = (int) (((long long) r * 13 + 5) % N);
         rows[0] = r; ncols[0] = 3;
         cols[0] = far; cols[1] = far; cols[2] = J[I[r] + (I[r + 1] - I[r]) / 2];
         vals[0] = 0.125; vals[1] = 0.25; vals[2] = 9.0;
         HYPRE_IJMatrixSetValues(ij, 1, ncols, rows, cols, vals);
         rows[0] = r; rows[1] = r; ncols[0] = 1; ncols[1] = 1;
         cols[0] = c2; cols[1] = c2; vals[0] = 1.5; vals[1] = 2.5;
         HYPRE_IJMatrixAddToValues(ij, 2, ncols, rows, cols, vals);
      }
   HYPRE_IJMatrixAssemble(ij);
   {
      void *obj; hypre_CSRMatrix *E; int *EI, *EJ; double *Ea;
      HYPRE_IJMatrixGetObject(ij, &obj);
      E = hypre_ParCSRMatrixDiag((hypre_ParCSRMatrix *) obj);
      EI = hypre_CSRMatrixI(E); EJ = hypre_CSRMatrixJ(E); Ea = hypre_CSRMatrixData(E);
      for (r = 0; r < N; r += 3)
      {
         rows[0] = r; ncols[0] = 1; cols[0] = r; vals[0] = 1.0;
         HYPRE_IJMatrixAddToValues(ij, 1, ncols, rows, cols, vals);
      }
      for (r = 0; r < N; r += 5)
         if (EI[r + 1] - EI[r] > 1)
         {
            rows[0] = r; ncols[0] = 1; cols[0] = EJ[EI[r] + 1]; vals[0] = Ea[EI[r] + 1] * 1.0;
            HYPRE_IJMatrixSetValues(ij, 1, ncols, rows, cols, vals);
         }
      HYPRE_IJMatrixAssemble(ij);
      HYPRE_IJMatrixGetObject(ij, &obj);
      *ij_out = ij;
      return (HYPRE_ParCSRMatrix) obj;
   }
}
static double now(void)
{
   struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
   return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
   int nx = 10, ny = 10, nz = 10, pt27 = 0, pmis = 0, rlx = -1, Pmx = 4, agg_nl = 0;
   int mod_rap2 = 0, keepT = 0, interp_type = 6, nodump = 0, matvec = 0, max_iter = 100;
   int ns = 1, ns_coarse = 1, mu = 1, fmg = 0, perturb = 0;     /* ij.c: -ns, -ns_coarse, -mu, -fmg */
   double cx = 1, cy = 1, cz = 1, th = 0.25, tol = 1e-8, mxrs = 1.0;
   int difconv = 0, atype = 0;                                   /* ij.c: -difconv, -a ax ay az, -atype */
   int rotate = 0; double alpha = 0., eps = 1.;                  /* ij.c: -rotate -alpha A -eps E (2-D: -n nx ny 1) */
   int ij_mode = 0, noamg = 0;                                   /* -ijbuild MODE (see ijbuild above), -noamg: dump A0 only */
   int solver_id = 1, k_dim = 5;                                 /* ij.c: -solver 1 AMG-PCG, 3 AMG-GMRES, 9 AMG-BiCGSTAB; -k */
   double ax = 1, ay = 1, az = 1;
   const char *ofile = NULL;
   int i;
   for (i = 1; i < argc; i++)
   {
      if (!strcmp(argv[i], "-n")) { nx = atoi(argv[++i]); ny = atoi(argv[++i]); nz = atoi(argv[++i]); }
      else if (!strcmp(argv[i], "-27pt")) pt27 = 1;
      else if (!strcmp(argv[i], "-c")) { cx = atof(argv[++i]); cy = atof(argv[++i]); cz = atof(argv[++i]); }
      else if (!strcmp(argv[i], "-difconv")) difconv = 1;
      else if (!strcmp(argv[i], "-a")) { ax = atof(argv[++i]); ay = atof(argv[++i]); az = atof(argv[++i]); }
      else if (!strcmp(argv[i], "-atype")) atype = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-rotate")) rotate = 1;
      else if (!strcmp(argv[i], "-alpha")) alpha = atof(argv[++i]);
      else if (!strcmp(argv[i], "-eps")) eps = atof(argv[++i]);
      else if (!strcmp(argv[i], "-ijbuild")) ij_mode = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-noamg")) noamg = 1;
      else if (!strcmp(argv[i], "-solver")) solver_id = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-k")) k_dim = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-pmis")) pmis = 1;
      else if (!strcmp(argv[i], "-hmis")) pmis = 0;                      /* coarsen_type 10, the default */
      else if (!strcmp(argv[i], "-rlx")) rlx = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-Pmx")) Pmx = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-agg_nl")) agg_nl = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-mod_rap2")) mod_rap2 = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-keepT")) keepT = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-interptype")) interp_type = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-th")) th = atof(argv[++i]);
      else if (!strcmp(argv[i], "-tol")) tol = atof(argv[++i]);
      else if (!strcmp(argv[i], "-mxrs")) mxrs = atof(argv[++i]);
      else if (!strcmp(argv[i], "-max_iter")) max_iter = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-matvec")) matvec = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-nodump")) nodump = 1;
      else if (!strcmp(argv[i], "-ns")) ns = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-ns_coarse")) ns_coarse = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-mu")) mu = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-fmg")) fmg = 1;
      else if (!strcmp(argv[i], "-perturb")) perturb = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-o")) ofile = argv[++i];
      else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
   }
   hypre_MPI_Init(&argc, &argv);
   HYPRE_Init();
   if (ofile && !nodump) g_out = fopen(ofile, "wb");

   /* matrix: same generator calls and stencil values as ij.c:7788-7810 / :9078-9086 */
   HYPRE_ParCSRMatrix A;
   double t0 = now();
   if (pt27)
   {
      HYPRE_Real values[2];
      values[0] = 26.0;
      if (nx == 1 || ny == 1 || nz == 1) values[0] = 8.0;
      if (nx * ny == 1 || nx * nz == 1 || ny * nz == 1) values[0] = 2.0;
      values[1] = -1.;
      A = (HYPRE_ParCSRMatrix) GenerateLaplacian27pt(hypre_MPI_COMM_WORLD, nx, ny, nz, 1, 1, 1, 0, 0, 0, values);
   }
   else if (rotate)
   {
      if (nz != 1) { fprintf(stderr, "-rotate is two-dimensional: -n nx ny 1\n"); return 2; }
      A = (HYPRE_ParCSRMatrix) GenerateRotate7pt(hypre_MPI_COMM_WORLD, nx, ny, 1, 1, 0, 0, alpha, eps);
   }
   else if (difconv)
   {
      HYPRE_Real values[7], c[3] = { cx, cy, cz }, a[3] = { ax, ay, az };
      difconv_values(nx, ny, nz, c, a, atype, values);
      A = (HYPRE_ParCSRMatrix) GenerateDifConv(hypre_MPI_COMM_WORLD, nx, ny, nz, 1, 1, 1, 0, 0, 0, values);
   }
   else
   {
      HYPRE_Real values[4];
      values[1] = -cx; values[2] = -cy; values[3] = -cz; values[0] = 0.;
      if (nx > 1) values[0] += 2.0 * cx;
      if (ny > 1) values[0] += 2.0 * cy;
      if (nz > 1) values[0] += 2.0 * cz;
      A = (HYPRE_ParCSRMatrix) GenerateLaplacian(hypre_MPI_COMM_WORLD, nx, ny, nz, 1, 1, 1, 0, 0, 0, values);
   }
   double t_gen = now() - t0;
   HYPRE_IJMatrix ij_A = NULL;
   if (ij_mode) A = ijbuild(A, ij_mode, &ij_A);      /* (the generated matrix is left to the OS) */
   hypre_ParCSRMatrix *pA = (hypre_ParCSRMatrix *) A;
   int N = hypre_CSRMatrixNumRows(hypre_ParCSRMatrixDiag(pA));
   if (perturb)                                   /* one rank: the whole operator is the diag block, diagonal entry first */
      perturb_operator(N, hypre_CSRMatrixI(hypre_ParCSRMatrixDiag(pA)), hypre_CSRMatrixJ(hypre_ParCSRMatrixDiag(pA)),
                       hypre_CSRMatrixData(hypre_ParCSRMatrixDiag(pA)), (unsigned) perturb);
   HYPRE_BigInt *row_starts = hypre_ParCSRMatrixRowStarts(pA);

   hypre_ParVector *b = hypre_ParVectorCreate(hypre_MPI_COMM_WORLD, N, row_starts);
   hypre_ParVectorSetPartitioningOwner(b, 0);
   hypre_ParVectorInitialize(b);
   hypre_ParVectorSetConstantValues(b, 1.0);       /* ij.c:2714-2748 build_rhs_type 2 */
   hypre_ParVector *x = hypre_ParVectorCreate(hypre_MPI_COMM_WORLD, N, row_starts);
   hypre_ParVectorSetPartitioningOwner(x, 0);
   hypre_ParVectorInitialize(x);
   hypre_ParVectorSetConstantValues(x, 0.0);

   printf("ref_dump: n=%d %d %d rows=%d nnz=%d threads=%d gen=%.3fs\n", nx, ny, nz, N,
          hypre_CSRMatrixI(hypre_ParCSRMatrixDiag(pA))[N], hypre_NumThreads(), t_gen);

   if (noamg)
   {  /* assembly check only: the operator as the reference's IJ interface built it */
      int hdr[8] = { nx, ny, nz, 1, 0, pt27, Pmx, rlx };
      put("hdr", 0, hdr, 8);
      put_csr("A", 0, hypre_ParCSRMatrixDiag(pA), 1);
      if (g_out) fclose(g_out);
      HYPRE_Finalize(); hypre_MPI_Finalize();
      return 0;
   }
   if (matvec > 0)
   {  /* ij -solver -1 analogue (ij.c:3206-3243): y = A x repeated */
      hypre_ParVectorSetConstantValues(x, 1.0);
      hypre_ParCSRMatrixMatvec(1.0, pA, x, 0.0, b);
      t0 = now();
      for (i = 0; i < matvec; i++) hypre_ParCSRMatrixMatvec(1.0, pA, x, 0.0, b);
      double dt = (now() - t0) / matvec;
      printf("ref_dump: matvec_ms=%.6f reps=%d\n", dt * 1e3, matvec);
      HYPRE_Finalize(); hypre_MPI_Finalize();
      return 0;
   }

   if (solver_id != 1 && solver_id != 3 && solver_id != 9) { fprintf(stderr, "-solver must be 1, 3 or 9\n"); return 2; }
   HYPRE_Solver pcg, amg;
   if (solver_id == 3)
   {  /* ij.c:5305-5312; print level 1 because gmres.c:513 records norms[iter] only then (it prints nothing else) */
      HYPRE_ParCSRGMRESCreate(hypre_MPI_COMM_WORLD, &pcg);
      HYPRE_GMRESSetKDim(pcg, k_dim);
      HYPRE_GMRESSetTol(pcg, tol);
      HYPRE_GMRESSetAbsoluteTol(pcg, 0.);
      HYPRE_GMRESSetLogging(pcg, 1);
      HYPRE_GMRESSetPrintLevel(pcg, 1);
      HYPRE_GMRESSetRelChange(pcg, 0);
   }
   else if (solver_id == 9)
   {  /* ij.c:6364-6369 */
      HYPRE_ParCSRBiCGSTABCreate(hypre_MPI_COMM_WORLD, &pcg);
      HYPRE_BiCGSTABSetTol(pcg, tol);
      HYPRE_BiCGSTABSetAbsoluteTol(pcg, 0.);
      HYPRE_BiCGSTABSetLogging(pcg, 1);
      HYPRE_BiCGSTABSetPrintLevel(pcg, 0);
   }
   else
   {
   HYPRE_ParCSRPCGCreate(hypre_MPI_COMM_WORLD, &pcg);
   HYPRE_PCGSetMaxIter(pcg, 1000);
   HYPRE_PCGSetTol(pcg, tol);
   HYPRE_PCGSetTwoNorm(pcg, 1);
   HYPRE_PCGSetRelChange(pcg, 0);
   HYPRE_PCGSetPrintLevel(pcg, 0);
   HYPRE_PCGSetLogging(pcg, 1);
   }

   HYPRE_BoomerAMGCreate(&amg);
   HYPRE_BoomerAMGSetInterpType(amg, interp_type);
   HYPRE_BoomerAMGSetTol(amg, 0.);
   HYPRE_BoomerAMGSetCoarsenType(amg, pmis ? 8 : 10);
   HYPRE_BoomerAMGSetStrongThreshold(amg, th);
   HYPRE_BoomerAMGSetMaxCoarseSize(amg, 9);
   HYPRE_BoomerAMGSetTruncFactor(amg, 0.);
   HYPRE_BoomerAMGSetPMaxElmts(amg, Pmx);
   HYPRE_BoomerAMGSetPrintLevel(amg, 0);
   HYPRE_BoomerAMGSetMaxIter(amg, 1);
   HYPRE_BoomerAMGSetCycleType(amg, mu);
   HYPRE_BoomerAMGSetFCycle(amg, fmg);
   HYPRE_BoomerAMGSetNumSweeps(amg, ns);
   if (rlx > -1) HYPRE_BoomerAMGSetRelaxType(amg, rlx);
   HYPRE_BoomerAMGSetRelaxOrder(amg, 0);
   HYPRE_BoomerAMGSetRelaxWt(amg, 1.0);
   HYPRE_BoomerAMGSetOuterWt(amg, 1.0);
   HYPRE_BoomerAMGSetMaxLevels(amg, 25);
   HYPRE_BoomerAMGSetMaxRowSum(amg, mxrs);
   HYPRE_BoomerAMGSetNumFunctions(amg, 1);
   HYPRE_BoomerAMGSetAggNumLevels(amg, agg_nl);
   HYPRE_BoomerAMGSetAggInterpType(amg, 4);
   HYPRE_BoomerAMGSetCycleNumSweeps(amg, ns_coarse, 3);
   HYPRE_BoomerAMGSetRAP2(amg, 0);
   HYPRE_BoomerAMGSetModuleRAP2(amg, mod_rap2);
   HYPRE_BoomerAMGSetKeepTranspose(amg, keepT);
   double t_setup, t_solve;
   HYPRE_Int its; HYPRE_Real relres;
   const double *norms;
   if (solver_id == 3)
   {
      HYPRE_GMRESSetMaxIter(pcg, max_iter);
      HYPRE_GMRESSetPrecond(pcg, (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSolve,
                            (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSetup, amg);
      t0 = now();
      HYPRE_GMRESSetup(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
      t_setup = now() - t0;
      t0 = now();
      HYPRE_GMRESSolve(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
      t_solve = now() - t0;
      HYPRE_GMRESGetNumIterations(pcg, &its);
      HYPRE_GMRESGetFinalRelativeResidualNorm(pcg, &relres);
      norms = ((hypre_GMRESData *) pcg)->norms;
   }
   else if (solver_id == 9)
   {
      HYPRE_BiCGSTABSetMaxIter(pcg, max_iter);
      HYPRE_BiCGSTABSetPrecond(pcg, (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSolve,
                               (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSetup, amg);
      t0 = now();
      HYPRE_BiCGSTABSetup(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
      t_setup = now() - t0;
      t0 = now();
      HYPRE_BiCGSTABSolve(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
      t_solve = now() - t0;
      HYPRE_BiCGSTABGetNumIterations(pcg, &its);
      HYPRE_BiCGSTABGetFinalRelativeResidualNorm(pcg, &relres);
      norms = ((hypre_BiCGSTABData *) pcg)->norms;
   }
   else
   {
   HYPRE_PCGSetMaxIter(pcg, max_iter);
   HYPRE_PCGSetPrecond(pcg, (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSolve,
                       (HYPRE_PtrToSolverFcn) HYPRE_BoomerAMGSetup, amg);

   t0 = now();
   HYPRE_PCGSetup(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
   t_setup = now() - t0;
   t0 = now();
   HYPRE_PCGSolve(pcg, (HYPRE_Matrix) A, (HYPRE_Vector) b, (HYPRE_Vector) x);
   t_solve = now() - t0;

   HYPRE_PCGGetNumIterations(pcg, &its);
   HYPRE_PCGGetFinalRelativeResidualNorm(pcg, &relres);
   norms = ((hypre_PCGData *) pcg)->norms;      /* residual history (pcg.c:597-600 norms[]) */
   }
   HYPRE_ClearAllErrors();       /* a run that stops at max_iter sets HYPRE_ERROR_CONV; the dump is still wanted */

   hypre_ParAMGData *ad = (hypre_ParAMGData *) amg;
   int nl = hypre_ParAMGDataNumLevels(ad);
   printf("ref_dump: levels=%d iterations=%d relres=%.6e setup_s=%.4f solve_s=%.4f\n",
          nl, (int) its, relres, t_setup, t_solve);
   for (i = 0; i < nl; i++)
   {
      hypre_ParCSRMatrix *Al = hypre_ParAMGDataAArray(ad)[i];
      int n = hypre_CSRMatrixNumRows(hypre_ParCSRMatrixDiag(Al));
      printf("ref_dump: level %d rows=%d nnz=%d\n", i, n, hypre_CSRMatrixI(hypre_ParCSRMatrixDiag(Al))[n]);
   }

   if (g_out)
   {
      int hdr[8] = { nx, ny, nz, nl, (int) its, pt27, Pmx, rlx };
      put("hdr", 0, hdr, 8);
      put("relres", 1, &relres, 1);
      put("norms", 1, norms, its + 1);
      put("x", 1, hypre_VectorData(hypre_ParVectorLocalVector(x)), N);
      for (i = 0; i < nl; i++)
      {
         char nm[64];
         hypre_ParCSRMatrix *Al = hypre_ParAMGDataAArray(ad)[i];
         int n = hypre_CSRMatrixNumRows(hypre_ParCSRMatrixDiag(Al));
         put_csr("A", i, hypre_ParCSRMatrixDiag(Al), 1);
         if (hypre_ParAMGDataL1Norms(ad) && hypre_ParAMGDataL1Norms(ad)[i])
         {
            sprintf(nm, "l1_%d", i);
            put(nm, 1, hypre_VectorData(hypre_ParAMGDataL1Norms(ad)[i]), n);
         }
         if (i < nl - 1)
         {
            hypre_ParCSRMatrix *S = NULL;
            sprintf(nm, "CF%d", i);
            put(nm, 0, hypre_ParAMGDataCFMarkerArray(ad)[i], n);
            put_csr("P", i, hypre_ParCSRMatrixDiag(hypre_ParAMGDataPArray(ad)[i]), 1);
            if (agg_nl == 0 || i >= agg_nl)
            {
               hypre_BoomerAMGCreateS(Al, th, mxrs, 1, NULL, &S);
               put_csr("S", i, hypre_ParCSRMatrixDiag(S), 0);
               hypre_ParCSRMatrixDestroy(S);
            }
         }
      }
      fclose(g_out);
   }
   HYPRE_Finalize();
   hypre_MPI_Finalize();
   return 0;
}
