/* amg_oracle.c -- CPU RESTATEMENT (port) of the reference's hot path.  TEST INFRASTRUCTURE ONLY:
 * nothing under hypre_ve_b200/ includes, links or executes this file.
 *
 * A plain sequential C restatement of the BoomerAMG path on one rank: the generators (7-point, 27-point, convection-diffusion,
 * rotated anisotropy; -P gives the numbering of a process grid), strength, PMIS and HMIS coarsening (-pmis / -hmis), aggressive
 * levels with multipass interpolation, ext+i interpolation with P_max_elmts truncation, both Galerkin-product orders, the
 * smoothers 0 / 7 / 18 (Jacobi family), 3 / 4 / 6 / 8 / 13 / 14 (hybrid Gauss-Seidel with thread or rank blocks) and 16
 * (Chebyshev), V / W / F cycles, Gaussian elimination or the stalled-coarsening sweep on the coarsest grid, and the drivers
 * -solver 0 (BoomerAMG alone), 1 (PCG, 2-norm test), 3 (GMRES(k)), 9 (BiCGSTAB).
 * Every routine cites the reference lines it follows (paths under /root/reference/src).
 *
 * PINNED: tests/test_oracle.py checks this program's output bit for bit (integers AND doubles)
 * against the golden dumps in tests/golden/, which were produced by the reference's own CPU build
 * (oracle/_ref/ref_dump, see tests/golden/make_golden.py), and -- where oracle/_ref exists --
 * against live reference runs, the SURVEY's known answers, and the two single-process jobs of the reference's own regression
 * suite (src/test/TEST_ij/default.saved, coarsening.saved).
 *
 * CLI and output format are those of oracle/ref_dump.c so the two can be diffed record by record.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA, like the reference's x86-64 build)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { int n, m, nnz; int *i, *j; double *a; } csr_t;

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "oom\n"); exit(1); } return p; }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) { fprintf(stderr, "oom\n"); exit(1); } return p; }
static csr_t csr_new(int n, int m, int nnz, int with_data)
{
   csr_t A; A.n = n; A.m = m; A.nnz = nnz;
   A.i = (int *) xcalloc((size_t) n + 1, sizeof(int));
   A.j = (int *) xmalloc(sizeof(int) * (size_t) nnz);
   A.a = with_data ? (double *) xmalloc(sizeof(double) * (size_t) nnz) : NULL;
   return A;
}
static void csr_free(csr_t *A) { free(A->i); free(A->j); free(A->a); A->i = A->j = NULL; A->a = NULL; }

/* ---- generators: parcsr_ls/par_laplace.c:124-300 and par_difconv.c:247-330 (7-pt: centre, z-,y-,x-,x+,y+,z+; v = centre,
 *      x-, y-, z-, x+, y+, z+ -- the Laplacian passes v[4..6] = v[1..3]) and
 *      parcsr_ls/par_laplace_27pt.c fill pass (centre, then (dz,dy,dx) lexicographic), 1 rank ---- */
/* -P P Q R: the operator in the numbering a P x Q x R process grid gives it (hypre_map, par_laplace.c:363-387: boxes in
 * rank order, lexicographic inside a box), i.e. the matrix an np = P*Q*R run holds, gathered.  Entries keep stencil order. */
static int g_P[3] = { 1, 1, 1 };
static void part1d(int length, int nprocs, int id, int *lo, int *hi)      /* hypre_GeneratePartitioning, seq_mv/genpart.c:18-38 */
{
   int size = length / nprocs, rest = length - size * nprocs;
   *lo = id * size + (id < rest ? id : rest);
   *hi = *lo + size + (id < rest ? 1 : 0);
}
static int owner1d(int i, int length, int nprocs) { int id, lo, hi; for (id = 0; id < nprocs; id++) { part1d(length, nprocs, id, &lo, &hi); if (i >= lo && i < hi) return id; } return nprocs - 1; }
static int box_index(int nx, int ny, int nz, int ix, int iy, int iz)
{
   int p = owner1d(ix, nx, g_P[0]), q = owner1d(iy, ny, g_P[1]), r = owner1d(iz, nz, g_P[2]), xa, xb, ya, yb, za, zb;
   part1d(nx, g_P[0], p, &xa, &xb); part1d(ny, g_P[1], q, &ya, &yb); part1d(nz, g_P[2], r, &za, &zb);
   return za * nx * ny + ya * nx * (zb - za) + xa * (yb - ya) * (zb - za) + ((iz - za) * (yb - ya) + (iy - ya)) * (xb - xa) + (ix - xa);
}
static csr_t gen_laplace(int nx, int ny, int nz, int pt27, const double *v)
{
   int n = nx * ny * nz, pass, ix, iy, iz, k;
   csr_t A; memset(&A, 0, sizeof A);
   for (pass = 0; pass < 2; pass++)
   {
      int cnt = 0, row = 0;
      for (iz = 0; iz < nz; iz++) for (iy = 0; iy < ny; iy++) for (ix = 0; ix < nx; ix++, row++)
      {
         if (pass) A.i[row] = cnt;
         int ns = pt27 == 1 ? 27 : 7;
         for (k = 0; k < ns; k++)
         {
            int dx, dy, dz; double val;
            if (pt27 == 2)
            {  /* GenerateRotate7pt (par_rotate_7pt.c:228-350): centre, (-1,-1), (0,-1), (-1,0), (+1,0), (0,+1), (+1,+1) */
               static const int rx[7] = {0, -1, 0, -1, 1, 0, 1}, ry[7] = {0, -1, -1, 0, 0, 1, 1}, rv[7] = {0, 3, 2, 1, 1, 2, 3};
               dx = rx[k]; dy = ry[k]; dz = 0; val = v[rv[k]];
            }
            else if (!pt27)
            {
               static const int ox[7] = {0, 0, 0, -1, 1, 0, 0}, oy[7] = {0, 0, -1, 0, 0, 1, 0}, oz[7] = {0, -1, 0, 0, 0, 0, 1};
               static const int vi[7] = {0, 3, 2, 1, 4, 5, 6};   /* par_difconv.c:247-330: lower and upper coefficients separate */
               dx = ox[k]; dy = oy[k]; dz = oz[k]; val = v[vi[k]];
            }
            else
            {
               if (k == 0) { dx = dy = dz = 0; val = v[0]; }
               else { int m = k - 1; if (m >= 13) m++; dz = m / 9 - 1; dy = (m / 3) % 3 - 1; dx = m % 3 - 1; val = v[1]; }
            }
            int jx = ix + dx, jy = iy + dy, jz = iz + dz;
            if (jx < 0 || jx >= nx || jy < 0 || jy >= ny || jz < 0 || jz >= nz) continue;
            if (pass) { A.j[cnt] = (jz * ny + jy) * nx + jx; A.a[cnt] = val; }
            cnt++;
         }
      }
      if (!pass) A = csr_new(n, n, cnt, 1); else A.i[n] = cnt;
   }
   if (g_P[0] * g_P[1] * g_P[2] > 1)
   {  /* renumber rows and columns into the process grid's order; every row keeps its entry order */
      int *perm = (int *) xmalloc(sizeof(int) * n), *len = (int *) xcalloc(n + 1, sizeof(int)), row = 0, r, p;
      csr_t B = csr_new(n, n, A.nnz, 1);
      for (iz = 0; iz < nz; iz++) for (iy = 0; iy < ny; iy++) for (ix = 0; ix < nx; ix++, row++) perm[row] = box_index(nx, ny, nz, ix, iy, iz);
      for (r = 0; r < n; r++) len[perm[r] + 1] = A.i[r + 1] - A.i[r];
      for (r = 0; r < n; r++) len[r + 1] += len[r];
      memcpy(B.i, len, sizeof(int) * (n + 1));
      for (r = 0; r < n; r++)
         for (p = 0, k = A.i[r]; k < A.i[r + 1]; k++, p++) { B.j[B.i[perm[r]] + p] = perm[A.j[k]]; B.a[B.i[perm[r]] + p] = A.a[k]; }
      free(perm); free(len); csr_free(&A);
      A = B;
   }
   return A;
}

/* ---- y = alpha*A*x + beta*b : seq_mv/csr_matvec.c:195-328 (row sums left to right, seeded
 *      with the scaled b term exactly as the optimized branch does) ---- */
static void matvec(double alpha, const csr_t *A, const double *x, double beta, const double *b, double *y)
{
   int i, jj;
   if (alpha == 0.0) { for (i = 0; i < A->n; i++) y[i] = beta * b[i]; return; }
   double temp = beta / alpha;
   for (i = 0; i < A->n; i++)
   {
      double t;
      if (temp == 0.0) t = 0.0;
      else if (temp == -1.0) t = (alpha == -1.0) ? b[i] : -b[i];
      else if (temp == 1.0) t = (alpha == -1.0) ? -b[i] : b[i];
      else t = (alpha == -1.0) ? -b[i] * temp : b[i] * temp;
      if (alpha == -1.0) for (jj = A->i[i]; jj < A->i[i + 1]; jj++) t -= A->a[jj] * x[A->j[jj]];
      else               for (jj = A->i[i]; jj < A->i[i + 1]; jj++) t += A->a[jj] * x[A->j[jj]];
      y[i] = (alpha == 1.0 || alpha == -1.0) ? t : alpha * t;
   }
}
static double dot(int n, const double *x, const double *y)     /* seq_mv/vector.c:511-545 */
{ double r = 0.0; int i; for (i = 0; i < n; i++) r += y[i] * x[i]; return r; }

/* ---- strength: parcsr_ls/par_strength.c:231-504 (num_functions 1, no offd) ---- */
static csr_t strength(const csr_t *A, double theta, double max_row_sum)
{
   int n = A->n, i, jA, pass;
   csr_t S; memset(&S, 0, sizeof S);
   for (pass = 0; pass < 2; pass++)
   {
      int cnt = 0;
      for (i = 0; i < n; i++)
      {
         if (pass) S.i[i] = cnt;
         int b = A->i[i], e = A->i[i + 1];
         if (b == e) continue;
         double diag = A->a[b], row_scale = 0.0, row_sum = diag;
         for (jA = b + 1; jA < e; jA++)
         {
            double v = A->a[jA];
            if (diag < 0) row_scale = (row_scale < v) ? v : row_scale; else row_scale = (row_scale < v) ? row_scale : v;
            row_sum += v;
         }
         if ((fabs(row_sum) > fabs(diag) * max_row_sum) && (max_row_sum < 1.0)) continue;   /* all weak */
         for (jA = b + 1; jA < e; jA++)
         {
            int weak = diag < 0 ? (A->a[jA] <= theta * row_scale) : (A->a[jA] >= theta * row_scale);
            if (!weak) { if (pass) S.j[cnt] = A->j[jA]; cnt++; }
         }
      }
      if (!pass) S = csr_new(n, n, cnt, 0); else S.i[n] = cnt;
   }
   return S;
}

/* ---- hypre_Rand: utilities/random.c:71-106 (Schrage form of 16807*seed mod 2^31-1) ---- */
static int g_seed = 13579;
static double hrand(void)
{
   int high = g_seed / 127773, low = g_seed % 127773, test = 16807 * low - 2836 * high;
   g_seed = test > 0 ? test : test + 2147483647;
   return (double) g_seed / 2147483647;
}

/* ---- PMIS: parcsr_ls/par_coarsen.c:2159-2700, CF_init 0, one rank; measures from
 *      par_indepset.c:44-59 (seed 2747, one draw per row in row order) ---- */
static int *pmis_init(const csr_t *S, int cf_init, int *cf_in);
static int *pmis(const csr_t *S) { return pmis_init(S, 0, NULL); }
/* cf_init 3 (second PMIS of aggressive coarsening, par_amg_setup.c:1253): isolated rows become C points
 * (:2322-2326) and the first sweep skips the independent-set selection (`if (!CF_init || iter)`, :2420) */
/* cf_init 1 (HMIS, :2795): cf_in holds the markers of the Ruge-Stueben first pass; its C points are the first
 * independent set, every F point becomes undecided again and Z points stay F only if nothing depends on them (:2279-2309) */
static int *pmis_init(const csr_t *S, int cf_init, int *cf_in)
{
   int n = S->n, i, k, jS;
   double *m = (double *) xcalloc(n, sizeof(double));
   int *cf = cf_in ? cf_in : (int *) xcalloc(n, sizeof(int)), *graph = (int *) xmalloc(sizeof(int) * n), gsize = 0;
   for (k = 0; k < S->nnz; k++) m[S->j[k]] += 1.0;
   g_seed = 2747;
   for (i = 0; i < n; i++) m[i] += hrand();
   if (cf_init == 1)
      for (i = 0; i < n; i++)
      {
         if (cf[i] != -3)
         {
            if (cf[i] == -1) cf[i] = 0;
            if (cf[i] == -2)
            {
               if (m[i] >= 1.0 || S->i[i + 1] - S->i[i] > 0) { cf[i] = 0; graph[gsize++] = i; }
               else cf[i] = -1;
            }
            else graph[gsize++] = i;
         }
         else m[i] = 0;
      }
   else
   for (i = 0; i < n; i++)
   {
      cf[i] = 0;
      if (S->i[i + 1] - S->i[i] == 0) { cf[i] = (cf_init == 3) ? 1 : -3; m[i] = 0; } else graph[gsize++] = i;
   }
   int iter = 0;
   while (gsize > 0)
   {
      int ig;
      if (!cf_init || iter)
      {
      for (ig = 0; ig < gsize; ig++) { i = graph[ig]; if (m[i] > 1) cf[i] = 1; }
      for (ig = 0; ig < gsize; ig++)
      {
         i = graph[ig];
         if (m[i] > 1)
            for (jS = S->i[i]; jS < S->i[i + 1]; jS++)
            {
               int j = S->j[jS];
               if (m[j] > 1) { if (m[i] > m[j]) cf[j] = 0; else if (m[j] > m[i]) cf[i] = 0; }
            }
      }
      }
      iter++;
      for (ig = 0; ig < gsize; ig++)
      {
         i = graph[ig];
         if (m[i] < 1) cf[i] = -1;
         if (cf[i] > 0) cf[i] = 1;
         else for (jS = S->i[i]; jS < S->i[i + 1]; jS++) if (cf[S->j[jS]] > 0) cf[i] = -1;
      }
      int g2 = 0;
      for (ig = 0; ig < gsize; ig++) { i = graph[ig]; if (cf[i] != 0) m[i] = 0; else graph[g2++] = i; }
      gsize = g2;
   }
   free(m); free(graph);
   return cf;
}

static csr_t transpose(const csr_t *A);

/* ---- Ruge-Stueben first pass as HMIS uses it: hypre_BoomerAMGCoarsenRuge with coarsen_type 10 -> 11, f_pnt = Z_PT,
 *      measure_type 0, no cut factor, one rank (par_coarsen.c:1046-1330).  The reference keeps the undecided points in a
 *      list of lists ordered by measure (utilities/amg_linklist.c): within one measure the points leave in the order they
 *      entered, and the next C point is the oldest point of the largest measure.  Same behaviour here with one FIFO per
 *      measure value, indexed directly. ---- */
typedef struct { int *head, *tail, *next, *prev, nb, maxm; } buckets_t;
static void bk_grow(buckets_t *b, int m)
{
   if (m < b->nb) return;
   int nb = 2 * m + 8, k;
   b->head = (int *) realloc(b->head, sizeof(int) * nb); b->tail = (int *) realloc(b->tail, sizeof(int) * nb);
   for (k = b->nb; k < nb; k++) b->head[k] = b->tail[k] = -1;
   b->nb = nb;
}
static void bk_enter(buckets_t *b, int m, int i)          /* hypre_enter_on_lists: append to the list of measure m */
{
   bk_grow(b, m);
   b->next[i] = -1; b->prev[i] = b->tail[m];
   if (b->tail[m] >= 0) b->next[b->tail[m]] = i; else b->head[m] = i;
   b->tail[m] = i;
   if (m > b->maxm) b->maxm = m;
}
static void bk_remove(buckets_t *b, int m, int i)         /* hypre_remove_point */
{
   if (b->prev[i] >= 0) b->next[b->prev[i]] = b->next[i]; else b->head[m] = b->next[i];
   if (b->next[i] >= 0) b->prev[b->next[i]] = b->prev[i]; else b->tail[m] = b->prev[i];
   while (b->maxm > 0 && b->head[b->maxm] < 0) b->maxm--;
}
/* agg2: the second coarsening of an aggressive level calls it with measure_type + 3 (par_amg_setup.c:1247-1250): isolated
 * points become special C points (SC_PT) instead of special F points */
static int *ruge_first_pass(const csr_t *S, int agg2)
{
   int n = S->n, i, j, k, num_left = 0;
   csr_t ST = transpose(S);                                /* :1014-1043, rows of ST ordered by source row */
   int *cf = (int *) xcalloc(n, sizeof(int)), *meas = (int *) xmalloc(sizeof(int) * n);
   buckets_t b; memset(&b, 0, sizeof b);
   b.next = (int *) xmalloc(sizeof(int) * n); b.prev = (int *) xmalloc(sizeof(int) * n);
   bk_grow(&b, 16);
   for (i = 0; i < n; i++) meas[i] = ST.i[i + 1] - ST.i[i];                     /* :1056-1059 */
   for (j = 0; j < n; j++)                                                        /* :1130-1158 (CF_marker starts at 0) */
   {
      if (S->i[j + 1] - S->i[j] == 0) { cf[j] = agg2 ? 3 : -3; meas[j] = 0; }
      else { cf[j] = 0; num_left++; }
   }
   for (j = 0; j < n; j++)                                                        /* :1179-1222 */
   {
      int measure = meas[j];
      if (cf[j] == -3 || cf[j] == 3) continue;
      if (measure > 0) bk_enter(&b, measure, j);
      else
      {
         cf[j] = -2;                                                              /* f_pnt = Z_PT */
         for (k = S->i[j]; k < S->i[j + 1]; k++)
         {
            int nabor = S->j[k];
            if (cf[nabor] != -3 && cf[nabor] != 3)
            {
               if (nabor < j)
               {
                  int nm = meas[nabor];
                  if (nm > 0) bk_remove(&b, nm, nabor);
                  nm = ++meas[nabor];
                  bk_enter(&b, nm, nabor);
               }
               else ++meas[nabor];
            }
         }
         --num_left;
      }
   }
   while (num_left > 0)                                                           /* :1245-1320 */
   {
      int index = b.head[b.maxm], measure = meas[index];
      cf[index] = 1;
      meas[index] = 0;
      --num_left;
      bk_remove(&b, measure, index);
      for (j = ST.i[index]; j < ST.i[index + 1]; j++)
      {
         int nabor = ST.j[j];
         if (cf[nabor] == 0)
         {
            cf[nabor] = -1;
            bk_remove(&b, meas[nabor], nabor);
            --num_left;
            for (k = S->i[nabor]; k < S->i[nabor + 1]; k++)
            {
               int n2 = S->j[k];
               if (cf[n2] == 0) { bk_remove(&b, meas[n2], n2); ++meas[n2]; bk_enter(&b, meas[n2], n2); }
            }
         }
      }
      for (j = S->i[index]; j < S->i[index + 1]; j++)
      {
         int nabor = S->j[j];
         if (cf[nabor] == 0)
         {
            int m2 = meas[nabor];
            bk_remove(&b, m2, nabor);
            meas[nabor] = --m2;
            if (m2 > 0) bk_enter(&b, m2, nabor);
            else
            {
               cf[nabor] = -1;
               --num_left;
               for (k = S->i[nabor]; k < S->i[nabor + 1]; k++)
               {
                  int n2 = S->j[k];
                  if (cf[n2] == 0) { bk_remove(&b, meas[n2], n2); ++meas[n2]; bk_enter(&b, meas[n2], n2); }
               }
            }
         }
      }
   }
   for (i = 0; i < n; i++) if (cf[i] == 3) cf[i] = 1;                            /* :1337-1343 SC_PT -> C_PT */
   free(meas); free(b.head); free(b.tail); free(b.next); free(b.prev); csr_free(&ST);
   return cf;
}
/* hypre_BoomerAMGCoarsenHMIS (:2774-2797): the first pass above, then PMIS seeded with its C points */
static int g_coarsen_type = 8;                             /* 8 PMIS (ij -pmis), 10 HMIS (library and driver default) */
static int *hmis(const csr_t *S, int agg2) { return pmis_init(S, 1, ruge_first_pass(S, agg2)); }

/* ---- utilities/hypre_qsort.c:367-387 ---- */
static void swap2(int *v, double *w, int i, int j) { int t = v[i]; v[i] = v[j]; v[j] = t; double s = w[i]; w[i] = w[j]; w[j] = s; }
static void qsort2_abs(int *v, double *w, int left, int right)
{
   int i, last;
   if (left >= right) return;
   swap2(v, w, left, (left + right) / 2);
   last = left;
   for (i = left + 1; i <= right; i++) if (fabs(w[i]) > fabs(w[left])) swap2(v, w, ++last, i);
   swap2(v, w, left, last);
   qsort2_abs(v, w, left, last - 1);
   qsort2_abs(v, w, last + 1, right);
}

/* ---- ext+i interpolation: parcsr_ls/par_lr_interp.c:1301-1416 (C-hat discovery) and
 *      :1523-1803 (weights); truncation parcsr_mv/par_csr_matrix.c:2906-3020 (rescale 1) ---- */
static csr_t extpi(const csr_t *A, const csr_t *S, const int *cf, int max_elmts, int *ncoarse_out)
{
   int n = A->n, i, jj, kk, jj1, nc = 0;
   int *f2c = (int *) xmalloc(sizeof(int) * n);
   /* the reference's P_marker / strong_f_marker bookkeeping (par_lr_interp.c:1547-1605) restated as
    * "owner[c] == i" (c was touched while building row i) + pos_of[c] (>=0: slot in C-hat_i, -2: strong F) */
   int *owner = (int *) xmalloc(sizeof(int) * n), *pos_of = (int *) xmalloc(sizeof(int) * n);
   for (i = 0; i < n; i++) { f2c[i] = cf[i] >= 0 ? nc++ : -1; owner[i] = -1; pos_of[i] = -1; }
   /* rows are built one at a time into a row buffer, truncated, then appended */
   int cap = 16 * n + 64, cnt = 0;
   csr_t P = csr_new(n, nc, cap, 1);
   int rowcap = 1024; int *rj = (int *) xmalloc(sizeof(int) * rowcap); double *ra = (double *) xmalloc(sizeof(double) * rowcap);
   for (i = 0; i < n; i++)
   {
      int len = 0;
      P.i[i] = cnt;
      if (cf[i] >= 0) { rj[0] = f2c[i]; ra[0] = 1.0; len = 1; }
      else if (cf[i] != -3)
      {
         #define IN_CHAT(c) (owner[c] == i && pos_of[c] >= 0)
         #define IS_SF(c) (owner[c] == i && pos_of[c] == -2)
         #define PUSH(c) do { if (!(owner[c] == i)) { owner[c] = i; pos_of[c] = len; \
               if (len >= rowcap) { rowcap *= 2; rj = (int *) realloc(rj, sizeof(int) * rowcap); ra = (double *) realloc(ra, sizeof(double) * rowcap); } \
               rj[len] = f2c[c]; ra[len] = 0.0; len++; } } while (0)
         for (jj = S->i[i]; jj < S->i[i + 1]; jj++)
         {
            int i1 = S->j[jj];
            if (cf[i1] >= 0) PUSH(i1);
            else if (cf[i1] != -3)
            {
               owner[i1] = i; pos_of[i1] = -2;
               for (kk = S->i[i1]; kk < S->i[i1 + 1]; kk++) { int k1 = S->j[kk]; if (cf[k1] >= 0) PUSH(k1); }
            }
         }
         double diagonal = A->a[A->i[i]];
         for (jj = A->i[i] + 1; jj < A->i[i + 1]; jj++)
         {
            int i1 = A->j[jj];
            if (IN_CHAT(i1)) ra[pos_of[i1]] += A->a[jj];
            else if (IS_SF(i1))
            {
               double sum = 0.0; int sgn = 1;
               if (A->a[A->i[i1]] < 0) sgn = -1;
               for (jj1 = A->i[i1] + 1; jj1 < A->i[i1 + 1]; jj1++)
               {
                  int i2 = A->j[jj1];
                  if ((IN_CHAT(i2) || i2 == i) && (sgn * A->a[jj1]) < 0) sum += A->a[jj1];
               }
               if (sum != 0)
               {
                  double distribute = A->a[jj] / sum;
                  for (jj1 = A->i[i1] + 1; jj1 < A->i[i1 + 1]; jj1++)
                  {
                     int i2 = A->j[jj1];
                     if (IN_CHAT(i2) && (sgn * A->a[jj1]) < 0) ra[pos_of[i2]] += distribute * A->a[jj1];
                     if (i2 == i && (sgn * A->a[jj1]) < 0) diagonal += distribute * A->a[jj1];
                  }
               }
               else diagonal += A->a[jj];
            }
            else if (cf[i1] != -3) diagonal += A->a[jj];
         }
         if (diagonal) for (jj = 0; jj < len; jj++) ra[jj] /= -diagonal;
         #undef IN_CHAT
         #undef IS_SF
         #undef PUSH
         /* truncation to max_elmts largest |w|, kept in sorted order, rescaled to the row sum */
         if (max_elmts > 0 && len > max_elmts)
         {
            double row_sum = 0, scale = 0;
            for (jj = 0; jj < len; jj++) row_sum += ra[jj];
            qsort2_abs(rj, ra, 0, len - 1);
            for (jj = 0; jj < max_elmts; jj++) scale += ra[jj];
            len = max_elmts;
            if (scale != 0. && scale != row_sum) { scale = row_sum / scale; for (jj = 0; jj < len; jj++) ra[jj] *= scale; }
         }
      }
      if (cnt + len > cap) { cap = 2 * (cnt + len); P.j = (int *) realloc(P.j, sizeof(int) * cap); P.a = (double *) realloc(P.a, sizeof(double) * cap); }
      memcpy(P.j + cnt, rj, sizeof(int) * len); memcpy(P.a + cnt, ra, sizeof(double) * len);
      cnt += len;
   }
   P.i[n] = cnt; P.nnz = cnt;
   free(f2c); free(owner); free(pos_of); free(rj); free(ra);
   *ncoarse_out = nc;
   return P;
}

/* ---- transpose: seq_mv/csr_matop.c:651-773 (stable counting sort by column) ---- */
static csr_t transpose(const csr_t *A)
{
   csr_t T = csr_new(A->m, A->n, A->nnz, 1);
   int i, k;
   for (k = 0; k < A->nnz; k++) T.i[A->j[k] + 1]++;
   for (i = 0; i < A->m; i++) T.i[i + 1] += T.i[i];
   int *next = (int *) xmalloc(sizeof(int) * (A->m + 1)); memcpy(next, T.i, sizeof(int) * (A->m + 1));
   for (i = 0; i < A->n; i++) for (k = A->i[i]; k < A->i[i + 1]; k++) { int p = next[A->j[k]]++; T.j[p] = i; if (A->a) T.a[p] = A->a[k]; }
   free(next);
   return T;
}

/* ---- C = A*B: seq_mv/csr_matop.c:375-468 (diagonal first when square, first-touch column order) ---- */
static csr_t multiply(const csr_t *A, const csr_t *B)
{
   int allsquare = A->n == B->m, ic, ia, ib, pass;
   int *mark = (int *) xmalloc(sizeof(int) * B->m);
   csr_t C; memset(&C, 0, sizeof C);
   for (pass = 0; pass < 2; pass++)
   {
      int cnt = 0;
      for (ib = 0; ib < B->m; ib++) mark[ib] = -1;
      for (ic = 0; ic < A->n; ic++)
      {
         int row_start = cnt;
         if (pass) C.i[ic] = cnt;
         if (allsquare) { mark[ic] = cnt; if (pass) { C.j[cnt] = ic; C.a[cnt] = 0; } cnt++; }
         for (ia = A->i[ic]; ia < A->i[ic + 1]; ia++)
         {
            int ja = A->j[ia]; double a = A->a[ia];
            for (ib = B->i[ja]; ib < B->i[ja + 1]; ib++)
            {
               int jb = B->j[ib];
               if (mark[jb] < row_start) { mark[jb] = cnt; if (pass) { C.j[cnt] = jb; C.a[cnt] = a * B->a[ib]; } cnt++; }
               else if (pass) C.a[mark[jb]] += a * B->a[ib];
            }
         }
      }
      if (!pass) C = csr_new(A->n, B->m, cnt, 1); else C.i[A->n] = cnt;
   }
   free(mark);
   return C;
}

/* ---- l1 norms option 1: parcsr_ls/ams.c:648-657,:739-769 + csr_matop.c:1326-1352 ---- */
static int g_gs_blocks = 1;   /* the reference's num_threads = hypre_NumThreads() */
/* hypre_ParCSRComputeL1Norms (ams.c:571-760), which hands over to ...L1NormsThreads (ams.c:3398-3650) when
 * num_threads > 1: option 1 = sum |a_ij|; option 4 = |a_ii| + 0.5 * sum of |a_ij| over the entries OUTSIDE
 * the thread block [ns, ne) of row i (plus the offd block: none on one rank), truncated by Remark 6.2 */
/* Gauss-Seidel blocks of a level: the reference's OpenMP threads split the rows evenly (par_relax.c:4400-4412); with -P the
 * blocks are the RANKS of an np = P*Q*R run instead, i.e. each rank's own rows -- its box on the finest level, its C points
 * on every coarser one (g_blk[level], filled by amg_setup) */
static int *g_blk[64];
static const int *g_blk_cur = NULL;
static void block_range(int n, int T, int k, int *ns, int *ne)
{
   if (g_blk_cur) { *ns = g_blk_cur[k]; *ne = g_blk_cur[k + 1]; return; }
   int size = n / T, rest = n - size * T;
   if (k < rest) { *ns = k * size + k; *ne = (k + 1) * size + k + 1; }
   else { *ns = k * size + rest; *ne = (k + 1) * size + rest; }
}
static double *l1_norms(const csr_t *A, int option)
{
   double *l1 = (double *) xmalloc(sizeof(double) * A->n); int i, j, k, T = g_gs_blocks, n = A->n;
   for (k = 0; k < T; k++)
   {
      int ns, ne;
      block_range(n, T, k, &ns, &ne);
      for (i = ns; i < ne; i++)
      {
         double s = 0.0, d = 0.0;
         if (option == 1) { for (j = A->i[i]; j < A->i[i + 1]; j++) s += fabs(A->a[j]); }
         else if (option == 5)
         {  /* relax 7: the diagonal itself, 1 where it is zero; no sign handling (ams.c:704-725) */
            for (j = A->i[i]; j < A->i[i + 1]; j++) if (A->j[j] == i) { s = A->a[j]; break; }
            l1[i] = (s == 0.0) ? 1.0 : s;
            continue;
         }
         else
         {
            for (j = A->i[i]; j < A->i[i + 1]; j++)
            {
               int ii = A->j[j];
               if (ii == i) { d = fabs(A->a[j]); s += fabs(A->a[j]); }
               else if (ii < ns || ii >= ne) s += 0.5 * fabs(A->a[j]);
            }
            if (s <= 4.0 / 3.0 * d) s = d;
         }
         l1[i] = (A->a[A->i[i]] < 0) ? -s : s;
      }
   }
   return l1;
}
/* ---- sstruct_ls/gselim.h ---- */
static void gselim(double *A, double *x, int n)
{
   int j, k, m;
   if (n == 1) { if (A[0] != 0.0) x[0] = x[0] / A[0]; return; }
   for (k = 0; k < n - 1; k++) if (A[k * n + k] != 0.0)
   {
      double divA = 1.0 / A[k * n + k];
      for (j = k + 1; j < n; j++) if (A[j * n + k] != 0.0)
      {
         double factor = A[j * n + k] * divA;
         for (m = k + 1; m < n; m++) A[j * n + m] -= factor * A[k * n + m];
         x[j] -= factor * x[k];
      }
   }
   for (k = n - 1; k > 0; --k) if (A[k * n + k] != 0.0)
   {
      x[k] /= A[k * n + k];
      for (j = 0; j < k; j++) if (A[j * n + k] != 0.0) x[j] -= x[k] * A[j * n + k];
   }
   if (A[0] != 0.0) x[0] /= A[0];
}

#define MAXLEV 25
typedef struct { int nl; csr_t A[MAXLEV], P[MAXLEV], R[MAXLEV], S[MAXLEV]; int *cf[MAXLEV]; double *l1[MAXLEV];
                 double *F[MAXLEV], *U[MAXLEV], *V; double *ge; int ge_n;
                 double *cheby_ds[MAXLEV], cheby_coefs[MAXLEV][5], max_eig[MAXLEV], min_eig[MAXLEV]; } amg_t;

/* ---- setup loop: parcsr_ls/par_amg_setup.c:889-2890 for coarsen 8 / interp 6 / mod_rap2 1 ---- */
/* ---- aggressive coarsening (par_amg_setup.c:1239-1256, :1590-1605), one rank ------------------------------ */
/* hypre_BoomerAMGCreate2ndSHost, num_paths 1 (par_strength.c:2326-2400 count, :2620-2700 fill): row ic of S2
 * (coarse point i1) lists, in first-touch order, the C points among the strong neighbours i2 of i1 and among
 * the strong neighbours of every such i2 (C or F), except ic itself */
static csr_t create2ndS(const csr_t *S, const int *cf, int *nc_out)
{
   int n = S->n, i, ic, jj1, jj2, nc = 0, nnz = 0, pass;
   int *f2c = (int *) xmalloc(sizeof(int) * n), *c2f = (int *) xmalloc(sizeof(int) * n);
   for (i = 0; i < n; i++) { f2c[i] = -1; if (cf[i] > 0) { f2c[i] = nc; c2f[nc++] = i; } }
   int *mark = (int *) xmalloc(sizeof(int) * (nc ? nc : 1));
   csr_t C; memset(&C, 0, sizeof C);
   for (pass = 0; pass < 2; pass++)
   {
      if (pass) { C = csr_new(nc, nc, nnz, 0); }
      for (i = 0; i < nc; i++) mark[i] = -1;
      nnz = 0;
      for (ic = 0; ic < nc; ic++)
      {
         int i1 = c2f[ic], row0 = nnz;
         if (pass) C.i[ic] = nnz;
         for (jj1 = S->i[i1]; jj1 < S->i[i1 + 1]; jj1++)
         {
            int i2 = S->j[jj1];
            if (cf[i2] > 0)
            {
               int idx = f2c[i2];
               if (mark[idx] < row0) { mark[idx] = nnz; if (pass) C.j[nnz] = idx; nnz++; }
            }
            for (jj2 = S->i[i2]; jj2 < S->i[i2 + 1]; jj2++)
            {
               int i3 = S->j[jj2];
               if (cf[i3] > 0)
               {
                  int idx = f2c[i3];
                  if (idx != ic && mark[idx] < row0) { mark[idx] = nnz; if (pass) C.j[nnz] = idx; nnz++; }
               }
            }
         }
      }
      if (pass) C.i[nc] = nnz;
   }
   free(f2c); free(c2f); free(mark);
   *nc_out = nc;
   return C;
}
/* hypre_BoomerAMGBuildMultipass (par_multi_interp.c:16-2061), one rank, weight_option 0, no truncation.
 * pass 0 = C points (identity rows); pass 1 = F points with a strong C neighbour: the strong C entries of the
 * row of A in A's order, scaled by alfa = -sum_N / (sum_C * a_ii) (:1600-1660); pass p >= 2 = points with a strong
 * neighbour of pass p-1: sum over those neighbours j (A's order) of a_ij * (row j of P), columns in first-touch
 * order, sum_C / sum_N accumulated product by product, same scaling (:1770-1860). At most 9 passes (:102). */
static csr_t multipass(const csr_t *A, const csr_t *S, int *cf, int *ncoarse_out)
{
   int n = A->n, i, j, k, nc = 0, pass, npass, remaining = 0;
   int *f2c = (int *) xmalloc(sizeof(int) * n), *assigned = (int *) xmalloc(sizeof(int) * n);
   for (i = 0; i < n; i++)
   {
      f2c[i] = -1; assigned[i] = -1;
      if (cf[i] == 1) { f2c[i] = nc++; assigned[i] = 0; } else if (cf[i] == -1) remaining++;
   }
   /* pass numbers (:404-510) */
   for (i = 0; i < n; i++) if (cf[i] == -1)
      for (j = S->i[i]; j < S->i[i + 1]; j++) if (cf[S->j[j]] == 1) { assigned[i] = 1; }
   for (i = 0; i < n; i++) if (assigned[i] == 1) remaining--;
   pass = 2;
   while (remaining && pass < 10)
   {
      for (i = 0; i < n; i++) if (cf[i] == -1 && assigned[i] == -1)
         for (j = S->i[i]; j < S->i[i + 1]; j++) if (assigned[S->j[j]] == pass - 1) { assigned[i] = pass; break; }
      for (i = 0; i < n; i++) if (assigned[i] == pass) remaining--;
      pass++;
   }
   npass = pass;
   /* rows are built pass by pass into per-row buffers */
   int **rj = (int **) xcalloc(n, sizeof(int *)); double **ra = (double **) xcalloc(n, sizeof(double *));
   int *rl = (int *) xcalloc(n, sizeof(int));
   int *marker = (int *) xmalloc(sizeof(int) * n), *pos = (int *) xmalloc(sizeof(int) * (nc ? nc : 1)), *pm = (int *) xmalloc(sizeof(int) * (nc ? nc : 1));
   for (i = 0; i < n; i++) marker[i] = -1;
   for (i = 0; i < nc; i++) pm[i] = -1;
   double alfa = 1.0;
   for (i = 0; i < n; i++) if (cf[i] == 1)
   { rj[i] = (int *) xmalloc(sizeof(int)); ra[i] = (double *) xmalloc(sizeof(double)); rj[i][0] = f2c[i]; ra[i][0] = 1.0; rl[i] = 1; }
   for (i = 0; i < n; i++) if (assigned[i] == 1)
   {
      int len = 0; double sum_C = 0, sum_N = 0;
      for (j = S->i[i]; j < S->i[i + 1]; j++) if (cf[S->j[j]] == 1) { marker[S->j[j]] = i; len++; }
      rj[i] = (int *) xmalloc(sizeof(int) * (len ? len : 1)); ra[i] = (double *) xmalloc(sizeof(double) * (len ? len : 1));
      len = 0;
      for (j = A->i[i] + 1; j < A->i[i + 1]; j++)
      {
         int j1 = A->j[j];
         if (cf[j1] != -3) sum_N += A->a[j];
         if (marker[j1] == i) { ra[i][len] = A->a[j]; rj[i][len++] = f2c[j1]; sum_C += A->a[j]; }
      }
      rl[i] = len;
      double diagonal = A->a[A->i[i]];
      if (sum_C * diagonal != 0) alfa = -sum_N / (sum_C * diagonal);
      for (j = 0; j < len; j++) ra[i][j] *= alfa;
   }
   for (pass = 2; pass < npass; pass++)
      for (i = 0; i < n; i++) if (assigned[i] == pass)
      {
         int len = 0, cap = 0; double sum_C = 0, sum_N = 0;
         for (j = S->i[i]; j < S->i[i + 1]; j++) if (assigned[S->j[j]] == pass - 1) { marker[S->j[j]] = i; cap += rl[S->j[j]]; }
         rj[i] = (int *) xmalloc(sizeof(int) * (cap ? cap : 1)); ra[i] = (double *) xmalloc(sizeof(double) * (cap ? cap : 1));
         for (j = A->i[i] + 1; j < A->i[i + 1]; j++)
         {
            int j1 = A->j[j];
            if (marker[j1] == i)
               for (k = 0; k < rl[j1]; k++)
               {
                  int k1 = rj[j1][k];
                  if (pm[k1] != i) { pm[k1] = i; pos[k1] = len; rj[i][len] = k1; ra[i][len] = 0; len++; }
                  alfa = A->a[j] * ra[j1][k];
                  ra[i][pos[k1]] += alfa;
                  sum_C += alfa; sum_N += alfa;
               }
            else if (cf[j1] != -3) sum_N += A->a[j];
         }
         rl[i] = len;
         double diagonal = A->a[A->i[i]];
         if (sum_C * diagonal != 0) alfa = -sum_N / (sum_C * diagonal);
         for (j = 0; j < len; j++) ra[i][j] *= alfa;
      }
   int nnz = 0;
   for (i = 0; i < n; i++) nnz += rl[i];
   csr_t P = csr_new(n, nc, nnz, 1);
   nnz = 0;
   for (i = 0; i < n; i++)
   {
      P.i[i] = nnz;
      for (j = 0; j < rl[i]; j++) { P.j[nnz] = rj[i][j]; P.a[nnz++] = ra[i][j]; }
      free(rj[i]); free(ra[i]);
   }
   P.i[n] = nnz;
   for (i = 0; i < n; i++) if (cf[i] == -3) cf[i] = -1;               /* :2030-2036 */
   free(rj); free(ra); free(rl); free(marker); free(pos); free(pm); free(f2c); free(assigned);
   *ncoarse_out = nc;
   return P;
}

static int g_agg_nl = 0, g_mod_rap2 = 0;

/* ---- Chebyshev smoother (relax 16): parcsr_ls/par_cheby.c + hypre_ParCSRMaxEigEstimateCG (par_relax_more.c:115-330) ---- */
/* eigenvalues of the symmetric tridiagonal matrix (d[0..n-1]; e[i] couples i-1 and i), ascending, by the implicit QL
 * iteration.  The reference calls the EISPACK routine tql1 (hypre_LINPACKcgtql1); any backward-stable method returns
 * the same values to a few ulp, which is all the Chebyshev coefficients need. */
static void tridiag_eigenvalues(int n, double *d, double *e)
{
   int l, m, i, iter;
   for (i = 1; i < n; i++) e[i - 1] = e[i];
   if (n > 0) e[n - 1] = 0.0;
   for (l = 0; l < n; l++)
   {
      iter = 0;
      do
      {
         for (m = l; m < n - 1; m++)
         {
            double dd = fabs(d[m]) + fabs(d[m + 1]);
            if (fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
         }
         if (m != l)
         {
            if (iter++ == 60) break;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]), r = sqrt(g * g + 1.0);
            g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
            double sn = 1.0, c = 1.0, p = 0.0;
            for (i = m - 1; i >= l; i--)
            {
               double f = sn * e[i], b = c * e[i];
               r = sqrt(f * f + g * g);
               e[i + 1] = r;
               if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
               sn = f / r; c = g / r;
               g = d[i + 1] - p;
               r = (d[i] - g) * sn + 2.0 * c * b;
               p = sn * r;
               d[i + 1] = g + p;
               g = c * r - b;
            }
            if (r == 0.0 && i >= l) continue;
            d[l] -= p; e[l] = g; e[m] = 0.0;
         }
      } while (m != l);
   }
   for (i = 1; i < n; i++) { double v = d[i]; int k = i - 1; while (k >= 0 && d[k] > v) { d[k + 1] = d[k]; k--; } d[k + 1] = v; }
}
static int g_cheby_order = 2, g_cheby_eig_est = 10, g_cheby_variant = 0, g_cheby_scale = 1;
static double g_cheby_fraction = 0.3;
static void cheby_setup(amg_t *g, int l)
{
   const csr_t *A = &g->A[l];
   int n = A->n, i, j, max_iter = g_cheby_eig_est;
   if (n < max_iter) max_iter = n;
   double *r = (double *) xmalloc(sizeof(double) * n), *p = (double *) xcalloc(n, sizeof(double)), *sv = (double *) xcalloc(n, sizeof(double));
   double *ds = (double *) xmalloc(sizeof(double) * n), *u = (double *) xcalloc(n, sizeof(double));
   double *td = (double *) xcalloc(max_iter + 1, sizeof(double)), *to = (double *) xcalloc(max_iter + 1, sizeof(double));
   g_seed = 1;                                                   /* hypre_ParVectorSetRandomValues(r, 1), vector.c:286-300 */
   for (i = 0; i < n; i++) r[i] = 2.0 * hrand() - 1.0;
   for (i = 0; i < n; i++) ds[i] = g_cheby_scale ? 1 / sqrt(A->a[A->i[i]]) : 1.0;
   double gamma = dot(n, r, p), gamma_old, beta = 1.0;
   i = 0;
   while (i < max_iter)
   {
      memcpy(sv, r, sizeof(double) * n);
      gamma_old = gamma;
      gamma = dot(n, r, sv);
      if (i == 0) { beta = 1.0; memcpy(p, sv, sizeof(double) * n); }
      else { beta = gamma / gamma_old; for (j = 0; j < n; j++) p[j] = sv[j] + beta * p[j]; }
      if (g_cheby_scale)
      {
         for (j = 0; j < n; j++) u[j] = ds[j] * p[j];
         matvec(1.0, A, u, 0.0, u, sv);
         for (j = 0; j < n; j++) sv[j] = ds[j] * sv[j];
      }
      else matvec(1.0, A, p, 0.0, p, sv);
      double sdotp = dot(n, sv, p), alpha = gamma / sdotp, alphainv = 1.0 / alpha;
      td[i + 1] = alphainv; td[i] *= beta; td[i] += alphainv;
      to[i + 1] = alphainv; to[i] *= sqrt(beta);
      for (j = 0; j < n; j++) r[j] += -alpha * sv[j];
      i++;
   }
   tridiag_eigenvalues(i, td, to);
   double max_eig = td[i - 1], min_eig = td[0];
   g->max_eig[l] = max_eig; g->min_eig[l] = min_eig;
   /* hypre_ParCSRRelax_Cheby_Setup (par_cheby.c:41-187) */
   int order = g_cheby_order; if (order > 4) order = 4; if (order < 1) order = 1;
   int co = order - 1;
   double ub = max_eig * 1.1, lb = (ub - min_eig) * g_cheby_fraction + min_eig, theta = (ub + lb) / 2, delta = (ub - lb) / 2, den;
   double *c = g->cheby_coefs[l];
   if (g_cheby_variant == 1)
   {
      switch (co)
      {
         case 0: c[0] = 1.0 / theta; break;
         case 1: den = (theta * theta + delta * theta); c[0] = (delta + 2 * theta) / den; c[1] = -1.0 / den; break;
         case 2: den = 2 * delta * theta * theta - delta * delta * theta - pow(delta, 3) + 2 * pow(theta, 3);
                 c[0] = (4 * delta * theta - pow(delta, 2) + 6 * pow(theta, 2)) / den; c[1] = -(2 * delta + 6 * theta) / den; c[2] = 2 / den; break;
         case 3: den = -(4 * delta * pow(theta, 3) - 3 * pow(delta, 2) * pow(theta, 2) - 3 * pow(delta, 3) * theta + 4 * pow(theta, 4));
                 c[0] = (6 * pow(delta, 2) * theta - 12 * delta * pow(theta, 2) + 3 * pow(delta, 3) - 16 * pow(theta, 3)) / den;
                 c[1] = (12 * delta * theta - 3 * pow(delta, 2) + 24 * pow(theta, 2)) / den; c[2] = -(4 * delta + 16 * theta) / den; c[3] = 4 / den; break;
      }
   }
   else
   {
      switch (co)
      {
         case 0: c[0] = 1.0 / theta; break;
         case 1: den = delta * delta - 2 * theta * theta; c[0] = -4 * theta / den; c[1] = 2 / den; break;
         case 2: den = 3 * (delta * delta) * theta - 4 * (theta * theta * theta);
                 c[0] = (3 * delta * delta - 12 * theta * theta) / den; c[1] = 12 * theta / den; c[2] = -4 / den; break;
         case 3: den = pow(delta, 4) - 8 * delta * delta * theta * theta + 8 * pow(theta, 4);
                 c[0] = (32 * pow(theta, 3) - 16 * delta * delta * theta) / den; c[1] = (8 * delta * delta - 48 * theta * theta) / den;
                 c[2] = 32 * theta / den; c[3] = -8 / den; break;
      }
   }
   g->cheby_ds[l] = ds;
   free(r); free(p); free(sv); free(u); free(td); free(to);
}
/* hypre_ParCSRRelax_Cheby_Solve (par_cheby.c:190-345) */
static void cheby_solve(amg_t *g, int l, const double *f, double *u)
{
   const csr_t *A = &g->A[l];
   int n = A->n, i, j, order = g_cheby_order; if (order > 4) order = 4; if (order < 1) order = 1;
   int co = order - 1;
   const double *c = g->cheby_coefs[l], *ds = g->cheby_ds[l];
   double *orig = (double *) xmalloc(sizeof(double) * n), *r = (double *) xmalloc(sizeof(double) * n);
   double *v = (double *) xmalloc(sizeof(double) * n), *tmp = (double *) xmalloc(sizeof(double) * n);
   if (!g_cheby_scale)
   {
      memcpy(r, f, sizeof(double) * n);
      matvec(-1.0, A, u, 1.0, r, r);
      for (i = 0; i < n; i++) { orig[i] = u[i]; u[i] = r[i] * c[co]; }
      for (i = co - 1; i >= 0; i--)
      {
         matvec(1.0, A, u, 0.0, u, v);
         for (j = 0; j < n; j++) u[j] = c[i] * r[j] + v[j];
      }
      for (i = 0; i < n; i++) u[i] = orig[i] + u[i];
   }
   else
   {
      matvec(-1.0, A, u, 0.0, u, tmp);
      for (j = 0; j < n; j++) r[j] = ds[j] * (f[j] + tmp[j]);
      for (j = 0; j < n; j++) { orig[j] = u[j]; u[j] = r[j] * c[co]; }
      for (i = co - 1; i >= 0; i--)
      {
         for (j = 0; j < n; j++) tmp[j] = ds[j] * u[j];
         matvec(1.0, A, tmp, 0.0, tmp, v);
         for (j = 0; j < n; j++) u[j] = c[i] * r[j] + ds[j] * v[j];
      }
      for (j = 0; j < n; j++) u[j] = orig[j] + ds[j] * u[j];
   }
   free(orig); free(r); free(v); free(tmp);
}
static int g_relax_down = 18, g_relax_up = 18;
static int g_user_relax = -1;        /* hypre_ParAMGDataUserRelaxType: the -rlx value, -1 when only the library defaults are set */
static int g_coarse_relax = 9;       /* grid_relax_type[3]: 9 = Gaussian elimination unless coarsening stalled (see amg_setup) */   /* grid_relax_type[1], [2] (par_amg.c:206-209, :1650-1672) */
static void amg_setup(amg_t *g, csr_t A0, double theta, double mrs, int pmax, int max_coarse)
{
   int l = 0, i;
   memset(g, 0, sizeof *g);
   g->A[0] = A0;
   while (1)
   {
      csr_t S = strength(&g->A[l], theta, mrs);
      int *cf = g_coarsen_type == 10 ? hmis(&S, 0) : pmis(&S), n = g->A[l].n, nc = 0;
      if (l < g_agg_nl)
      {  /* second coarsening on the distance-two graph of the C points, then CorrectCFMarker (par_strength.c:2957-2974) */
         int nc1 = 0, cnt = 0;
         csr_t S2 = create2ndS(&S, cf, &nc1);
         int *cfn = g_coarsen_type == 10 ? hmis(&S2, 1) : pmis_init(&S2, 3, NULL);
         for (i = 0; i < n; i++) if (cf[i] > 0) { if (cf[i] == 1) cf[i] = cfn[cnt++]; else { cf[i] = 1; cnt++; } }
         csr_free(&S2); free(cfn);
      }
      for (i = 0; i < n; i++) if (cf[i] == 1) nc++;
      if (nc == 0 || nc == n)
      {  /* no coarse grid: stop, and the coarsest solve becomes ONE sweep of grid_relax_type[0] -- the -rlx type, or 3 with
            the library defaults (par_amg_setup.c:1484-1497, par_amg.c:2100-2102) */
         g_coarse_relax = g_user_relax > -1 ? g_user_relax : 3;
         csr_free(&S); free(cf); break;
      }
      if (l < g_agg_nl) g->P[l] = multipass(&g->A[l], &S, cf, &nc);
      else g->P[l] = extpi(&g->A[l], &S, cf, pmax, &nc);
      for (i = 0; i < n; i++) if (cf[i] == -3) cf[i] = -1;            /* par_lr_interp.c:1888-1894 */
      g->cf[l] = cf; g->S[l] = S;
      if (g_blk[l])
      {  /* the ranks keep their own C points: block k of the next level = the C points among block k's rows */
         int T = g_gs_blocks, k, c = 0, r = 0;
         g_blk[l + 1] = (int *) xcalloc(T + 1, sizeof(int));
         for (k = 0; k < T; k++) { for (; r < g_blk[l][k + 1]; r++) if (cf[r] > 0) c++; g_blk[l + 1][k + 1] = c; }
      }
      g->R[l] = transpose(&g->P[l]);                                  /* par_csr_triplemat.c:874-876 */
      if (g_mod_rap2)
      {  /* hypre_ParCSRMatrixRAPKT: R (A P), par_csr_triplemat.c:872-888 */
         csr_t Q = multiply(&g->A[l], &g->P[l]);
         g->A[l + 1] = multiply(&g->R[l], &Q);
         csr_free(&Q);
      }
      else
      {  /* hypre_BoomerAMGBuildCoarseOperatorKT (the library default): row ic of R A is formed first
            (par_rap.c:1640-1700), then multiplied by P with the diagonal entry created first (:1546-1553, :1790-1857) */
         csr_t Q = multiply(&g->R[l], &g->A[l]);
         g->A[l + 1] = multiply(&Q, &g->P[l]);
         csr_free(&Q);
      }
      l++;
      if (l == MAXLEV - 1 || nc <= max_coarse) break;
   }
   g->nl = l + 1;
   for (i = 0; i < g->nl; i++)
   {
      int n = g->A[i].n;
      g_blk_cur = g_blk[i];
      g->l1[i] = l1_norms(&g->A[i], g_relax_down == 18 ? 1 : (g_relax_down == 7 ? 5 : 4));      /* par_amg_setup.c:3018-3100 */
      g->F[i] = (double *) xcalloc(n, sizeof(double)); g->U[i] = (double *) xcalloc(n, sizeof(double));
      if (g_relax_down == 16 && (i < g->nl - 1 || g->A[i].n > max_coarse)) cheby_setup(g, i);      /* par_amg_setup.c:3137-3160 */
   }
   g->V = (double *) xcalloc(g->A[0].n, sizeof(double));
   csr_t *Ac = &g->A[g->nl - 1];
   if (Ac->n <= max_coarse && g_coarse_relax == 9 && g->nl > 1)
   {  /* par_gauss_elim.c:100-115 */
      int n = Ac->n, jj; g->ge_n = n; g->ge = (double *) xcalloc((size_t) n * n, sizeof(double));
      for (i = 0; i < n; i++) for (jj = Ac->i[i]; jj < Ac->i[i + 1]; jj++) g->ge[i * n + Ac->j[jj]] = Ac->a[jj];
   }
}

/* one row of the hybrid Gauss-Seidel family (par_relax.c), thread block [ns, ne): in-block neighbours are
 * read from u (new where already swept), everything else from tmp (the iterate before the sweep).
 * l1 variants 8/13/14 (:3492-4091, :4340-5124): res = f_i - sum_j a_ij u_j over the WHOLE row in storage
 * order, u_i += res / l1_i;  classic variants 3/4/6 (:1875-2265, original type 6 kept under `#if 0` at
 * :2685-2753): the diagonal (stored first) is skipped and u_i = res / a_ii. */
static void gs_row(const csr_t *A, const double *l1, const double *f, double *u, const double *tmp, int ns, int ne,
                   int i, int classic)
{
   int jj, ii;
   if (classic)
   {
      if (A->a[A->i[i]] != 0.0)
      {
         double res = f[i];
         for (jj = A->i[i] + 1; jj < A->i[i + 1]; jj++)
         { ii = A->j[jj]; res -= A->a[jj] * ((ii >= ns && ii < ne) ? u[ii] : tmp[ii]); }
         u[i] = res / A->a[A->i[i]];
      }
   }
   else if (l1[i] != 0.0)
   {
      double res = f[i];
      for (jj = A->i[i]; jj < A->i[i + 1]; jj++)
      { ii = A->j[jj]; res -= A->a[jj] * ((ii >= ns && ii < ne) ? u[ii] : tmp[ii]); }
      u[i] += res / l1[i];
   }
}
/* relaxation sweep of type `type`, relax_weight = omega = 1, relax_points = 0, one rank, g_gs_blocks thread
 * blocks (par_relax.c:4400-4412).  18: l1-Jacobi, parcsr_ls/ams.c:72-92 (v=f; v=-A u + v; u += v/l1) */
static void relax(amg_t *g, int l, int type, const double *f, double *u)
{
   int n = g->A[l].n, i, j, T = g_gs_blocks; double *v = g->V;
   const csr_t *A = &g->A[l]; const double *l1 = g->l1[l];
   if (type == 18 || type == 7)
   {  /* 7: Jacobi through the matvec, Vtemp = w f - w A u, u += Vtemp / a_ii (par_relax.c:3463-3490), w = 1 */
      matvec(-1.0, A, u, 1.0, f, v);
      for (i = 0; i < n; i++) u[i] += v[i] / l1[i];
      return;
   }
   if (type == 16) { cheby_solve(g, l, f, u); return; }
   if (type == 0)
   {  /* weighted Jacobi, weight 1 (par_relax.c:139-248): every point from the previous iterate, zero diagonals skipped */
      memcpy(v, u, sizeof(double) * n);
      for (i = 0; i < n; i++)
      {
         if (A->a[A->i[i]] != 0.0)
         {
            double res = f[i];
            for (j = A->i[i] + 1; j < A->i[i + 1]; j++) res -= A->a[j] * v[A->j[j]];
            u[i] *= 0.0;                                  /* one_minus_weight */
            u[i] += 1.0 * res / A->a[A->i[i]];
         }
      }
      return;
   }
   int classic = (type == 3 || type == 4 || type == 6);
   int fwd = (type == 3 || type == 13 || type == 6 || type == 8), bwd = (type == 4 || type == 14 || type == 6 || type == 8);
   if (!fwd && !bwd) { fprintf(stderr, "amg_oracle: relax type %d not restated\n", type); exit(2); }
   double *tmp = (double *) xmalloc(sizeof(double) * n);
   memcpy(tmp, u, sizeof(double) * n);
   g_blk_cur = g_blk[l];
   for (j = 0; j < T; j++)
   {
      int ns, ne;
      block_range(n, T, j, &ns, &ne);
      if (fwd) for (i = ns; i < ne; i++) gs_row(A, l1, f, u, tmp, ns, ne, i, classic);
      if (bwd) for (i = ne - 1; i > ns - 1; i--) gs_row(A, l1, f, u, tmp, ns, ne, i, classic);
   }
   free(tmp);
}
/* One multigrid cycle, parcsr_ls/par_cycle.c:180-622: the level-counter state machine (V: cycle_type 1, W: 2, F-cycle flag),
 * num_grid_sweeps[1..3] sweeps on the way down / up / on the coarsest grid (par_amg.c:1934-1962, :1990-2030). */
static int g_ns[4] = { 1, 1, 1, 1 }, g_cycle_type = 1, g_fcycle = 0;
static double g_cycle_op_count;     /* hypre_ParAMGDataCycleOpCount: nnz(A_level) per relaxation sweep of the last cycle (par_cycle.c:352) */
static void cycle(amg_t *g, const double *f, double *u)
{
   int i, j, k, nl = g->nl, level = 0, cycle_param = 1, not_finished = 1, fcycle_lev = nl - 2;
   int *lev_counter = (int *) xmalloc(sizeof(int) * nl);
   lev_counter[0] = 1;
   for (k = 1; k < nl; k++) lev_counter[k] = g_fcycle ? 1 : g_cycle_type;
   while (not_finished)
   {
      const double *F = level ? g->F[level] : f; double *U = level ? g->U[level] : u;
      int num_sweep = (nl > 1) ? g_ns[cycle_param] : 1;
      int type = (cycle_param == 2) ? g_relax_up : g_relax_down;
      if (nl == 1) type = g_user_relax > -1 ? g_user_relax : 6;        /* no coarsening at all: one sweep of the user's type, else 6 (par_cycle.c:289-300) */
      else if (cycle_param == 3 && g_coarse_relax != 9) { type = g_coarse_relax; num_sweep = 1; }
      for (j = 0; j < num_sweep; j++)
      {
         g_cycle_op_count += (double) g->A[level].i[g->A[level].n];
         if (level == nl - 1 && g->ge)
         {  /* grid_relax_type[3] = 9 (par_gauss_elim.c) */
            int n = g->ge_n; double *T = (double *) xmalloc(sizeof(double) * n * n), *b = (double *) xmalloc(sizeof(double) * n);
            memcpy(T, g->ge, sizeof(double) * n * n); memcpy(b, F, sizeof(double) * n);
            gselim(T, b, n);
            memcpy(U, b, sizeof(double) * n); free(T); free(b);
         }
         else relax(g, level, type, F, U);
      }
      --lev_counter[level];
      if (lev_counter[level] >= 0 && level != nl - 1)
      {  /* :534-591 */
         for (i = 0; i < g->A[level + 1].n; i++) g->U[level + 1][i] = 0.0;
         matvec(-1.0, &g->A[level], U, 1.0, F, g->V);
         matvec(1.0, &g->R[level], g->V, 0.0, g->V, g->F[level + 1]);
         ++level;
         if (lev_counter[level] < g_cycle_type) lev_counter[level] = g_cycle_type;
         cycle_param = (level == nl - 1) ? 3 : 1;
      }
      else if (level != 0)
      {  /* :592-625 */
         double *Uf = (level - 1) ? g->U[level - 1] : u;
         matvec(1.0, &g->P[level - 1], g->U[level], 1.0, Uf, Uf);
         --level;
         cycle_param = 2;
         if (g_fcycle && fcycle_lev == level) { if (lev_counter[level] < 1) lev_counter[level] = 1; fcycle_lev--; }
      }
      else not_finished = 0;
   }
   free(lev_counter);
}



/* ---- GMRES(k) with right preconditioning: krylov/gmres.c:226-800 (rel_change 0, cf_tol 0, min_iter 0,
 *      skip_real_r_check 0); M^{-1} = one cycle from a zero guess.  Returns the iteration count; *rnorm = last r_norm. ---- */
static void axpy(int n, double a, const double *x, double *y) { int i; for (i = 0; i < n; i++) y[i] += a * x[i]; }   /* vector.c:451-490 */
static void scal(int n, double a, double *y) { int i; for (i = 0; i < n; i++) y[i] *= a; }                          /* vector.c:394-430 */
static int gmres(amg_t *g, const csr_t *A, const double *b, double *x, int k_dim, double r_tol, int max_iter, double *norms,
                 double *rel_out)
{
   int n = A->n, i = 0, j, k, iter = 0;
   double **p = (double **) xmalloc(sizeof(double *) * (k_dim + 1)), **hh = (double **) xmalloc(sizeof(double *) * (k_dim + 1));
   double *r = (double *) xcalloc(n, sizeof(double)), *w = (double *) xcalloc(n, sizeof(double));
   double *rs = (double *) xcalloc(k_dim + 1, sizeof(double)), *c = (double *) xcalloc(k_dim, sizeof(double)), *sn = (double *) xcalloc(k_dim, sizeof(double));
   double epsmac = 1.e-16, t, gamma, r_norm, b_norm, den_norm, epsilon, real_r_norm_old, real_r_norm_new;
   for (j = 0; j <= k_dim; j++) { p[j] = (double *) xcalloc(n, sizeof(double)); hh[j] = (double *) xcalloc(k_dim, sizeof(double)); }
   memcpy(p[0], b, sizeof(double) * n);
   matvec(-1.0, A, x, 1.0, p[0], p[0]);                                   /* :316-319 */
   b_norm = sqrt(dot(n, b, b)); real_r_norm_old = b_norm;
   r_norm = sqrt(dot(n, p[0], p[0]));
   norms[0] = r_norm;
   den_norm = b_norm > 0.0 ? b_norm : r_norm;                             /* :388-394 */
   epsilon = r_tol * den_norm;                                            /* :403, a_tol 0 */
   while (iter < max_iter)
   {
      rs[0] = r_norm;
      if (r_norm == 0.0) break;                                           /* :427-439 */
      if (r_norm <= epsilon)                                              /* :443-462 */
      {
         memcpy(r, b, sizeof(double) * n);
         matvec(-1.0, A, x, 1.0, r, r);
         r_norm = sqrt(dot(n, r, r));
         if (r_norm <= epsilon) break;
      }
      t = 1.0 / r_norm;
      scal(n, t, p[0]);
      i = 0;
      while (i < k_dim && iter < max_iter)                                /* :469 restart cycle */
      {
         i++; iter++;
         memset(r, 0, sizeof(double) * n);
         cycle(g, p[i - 1], r);
         matvec(1.0, A, r, 0.0, p[i], p[i]);
         for (j = 0; j < i; j++)                                          /* modified Gram-Schmidt :476-479 */
         {
            hh[j][i - 1] = dot(n, p[j], p[i]);
            axpy(n, -hh[j][i - 1], p[j], p[i]);
         }
         t = sqrt(dot(n, p[i], p[i]));
         hh[i][i - 1] = t;
         if (t != 0.0) { t = 1.0 / t; scal(n, t, p[i]); }
         for (j = 1; j < i; j++)                                          /* :488-492 */
         {
            t = hh[j - 1][i - 1];
            hh[j - 1][i - 1] = sn[j - 1] * hh[j][i - 1] + c[j - 1] * t;
            hh[j][i - 1] = -sn[j - 1] * t + c[j - 1] * hh[j][i - 1];
         }
         t = hh[i][i - 1] * hh[i][i - 1];
         t += hh[i - 1][i - 1] * hh[i - 1][i - 1];
         gamma = sqrt(t);
         if (gamma == 0.0) gamma = epsmac;
         c[i - 1] = hh[i - 1][i - 1] / gamma;
         sn[i - 1] = hh[i][i - 1] / gamma;
         rs[i] = -hh[i][i - 1] * rs[i - 1];
         rs[i] /= gamma;
         rs[i - 1] = c[i - 1] * rs[i - 1];
         hh[i - 1][i - 1] = sn[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];
         r_norm = fabs(rs[i]);
         norms[iter] = r_norm;
         if (r_norm <= epsilon) break;                                    /* :541, no relative change */
      }
      rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];                           /* :641-649 triangular solve */
      for (k = i - 2; k >= 0; k--)
      {
         t = 0.0;
         for (j = k + 1; j < i; j++) t -= hh[k][j] * rs[j];
         t += rs[k];
         rs[k] = t / hh[k][k];
      }
      memcpy(w, p[i - 1], sizeof(double) * n);                            /* :651-654 */
      scal(n, rs[i - 1], w);
      for (j = i - 2; j >= 0; j--) axpy(n, rs[j], p[j], w);
      memset(r, 0, sizeof(double) * n);
      cycle(g, w, r);
      axpy(n, 1.0, r, x);
      if (r_norm <= epsilon)                                              /* :664-752 check the true residual */
      {
         memcpy(r, b, sizeof(double) * n);
         matvec(-1.0, A, x, 1.0, r, r);
         real_r_norm_new = r_norm = sqrt(dot(n, r, r));
         if (r_norm <= epsilon) break;
         if (real_r_norm_new >= real_r_norm_old) break;
         memcpy(p[0], r, sizeof(double) * n);
         i = 0;
         real_r_norm_old = real_r_norm_new;
      }
      for (j = i; j > 0; j--)                                             /* :755-768 residual vector for the restart */
      {
         rs[j - 1] = -sn[j - 1] * rs[j];
         rs[j] = c[j - 1] * rs[j];
      }
      if (i) axpy(n, rs[i] - 1.0, p[i], p[i]);
      for (j = i - 1; j > 0; j--) axpy(n, rs[j], p[j], p[i]);
      if (i) { axpy(n, rs[0] - 1.0, p[0], p[0]); axpy(n, 1.0, p[i], p[0]); }
   }
   *rel_out = b_norm > 0.0 ? r_norm / b_norm : r_norm;                    /* :777-783 */
   return iter;
}

/* ---- BiCGSTAB: krylov/bicgstab.c:207-530 (stop_crit 0, a_tol 0, cf_tol 0, min_iter 0) ---- */
static int bicgstab(amg_t *g, const csr_t *A, const double *b, double *x, double r_tol, int max_iter, double *norms, double *rel_out)
{
   int n = A->n, iter = 0;
   double *r = (double *) xcalloc(n, sizeof(double)), *r0 = (double *) xcalloc(n, sizeof(double)), *s = (double *) xcalloc(n, sizeof(double));
   double *v = (double *) xcalloc(n, sizeof(double)), *p = (double *) xcalloc(n, sizeof(double)), *q = (double *) xcalloc(n, sizeof(double));
   double alpha, beta, gamma, epsilon, temp, res, r_norm, b_norm, den_norm, gamma_numer, gamma_denom, epsmac = 2.2250738585072014e-308;
   memcpy(r0, b, sizeof(double) * n);
   matvec(-1.0, A, x, 1.0, r0, r0);
   memcpy(r, r0, sizeof(double) * n);
   memcpy(p, r0, sizeof(double) * n);
   b_norm = sqrt(dot(n, b, b));
   res = dot(n, r0, r0);
   r_norm = sqrt(res);
   norms[0] = r_norm;
   den_norm = b_norm > 0.0 ? b_norm : r_norm;
   epsilon = r_tol * den_norm;
   *rel_out = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   if (r_norm == 0.0 || r_norm <= epsilon) return 0;                     /* :396-413 */
   while (iter < max_iter)
   {
      iter++;
      memset(v, 0, sizeof(double) * n); cycle(g, p, v);
      matvec(1.0, A, v, 0.0, q, q);
      temp = dot(n, r0, q);
      if (fabs(temp) >= epsmac) alpha = res / temp; else { fprintf(stderr, "BiCGSTAB broke down\n"); break; }
      axpy(n, alpha, v, x);
      axpy(n, -alpha, q, r);
      memset(v, 0, sizeof(double) * n); cycle(g, r, v);
      matvec(1.0, A, v, 0.0, s, s);
      gamma_numer = dot(n, r, s);
      gamma_denom = dot(n, s, s);
      gamma = (gamma_numer == 0.0 && gamma_denom == 0.0) ? 0.0 : gamma_numer / gamma_denom;
      axpy(n, gamma, v, x);
      axpy(n, -gamma, s, r);
      r_norm = sqrt(dot(n, r, r));
      norms[iter] = r_norm;
      if (r_norm <= epsilon)                                              /* :464-481 */
      {
         memcpy(r, b, sizeof(double) * n);
         matvec(-1.0, A, x, 1.0, r, r);
         r_norm = sqrt(dot(n, r, r));
         if (r_norm <= epsilon) break;
      }
      if (fabs(res) >= epsmac) beta = 1.0 / res; else { fprintf(stderr, "BiCGSTAB broke down\n"); break; }
      res = dot(n, r0, r);
      beta *= res;
      axpy(n, -gamma, q, p);
      if (fabs(gamma) >= epsmac) scal(n, beta * alpha / gamma, p); else { fprintf(stderr, "BiCGSTAB broke down\n"); break; }
      axpy(n, 1.0, r, p);
   }
   *rel_out = b_norm > 0.0 ? r_norm / b_norm : r_norm;
   return iter;
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

/* ---- output in ref_dump.c's record format ---- */
static FILE *g_out;
static void put(const char *name, int dtype, const void *p, size_t n)
{
   unsigned int nl = (unsigned int) strlen(name), dt = (unsigned int) dtype; unsigned long long cnt = n;
   if (!g_out) return;
   fwrite(&nl, 4, 1, g_out); fwrite(name, 1, nl, g_out); fwrite(&dt, 4, 1, g_out); fwrite(&cnt, 8, 1, g_out);
   if (n) fwrite(p, dtype ? 8 : 4, n, g_out);
}
static void put_csr(const char *pre, int l, const csr_t *M, int with_data)
{
   char nm[64]; int dims[3] = { M->n, M->m, M->i[M->n] };
   sprintf(nm, "%s%d.dims", pre, l); put(nm, 0, dims, 3);
   sprintf(nm, "%s%d.i", pre, l); put(nm, 0, M->i, M->n + 1);
   sprintf(nm, "%s%d.j", pre, l); put(nm, 0, M->j, M->i[M->n]);
   if (with_data) { sprintf(nm, "%s%d.a", pre, l); put(nm, 1, M->a, M->i[M->n]); }
}

/* values[7] of the convection-diffusion stencil -cx Dxx - cy Dyy - cz Dzz + ax Dx + ay Dy + az Dz, computed as the
 * reference driver does (src/test/ij.c:8266-8409, BuildParDifConv; same helper as oracle/ref_dump.c): centre, x-, y-, z-, x+, y+, z+;
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
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char **argv)
{
   int nx = 10, ny = 10, nz = 10, pt27 = 0, Pmx = 4, max_iter = 100, i, matvec_reps = 0, rlx = -1, perturb = 0;
   double cx = 1, cy = 1, cz = 1, th = 0.25, tol = 1e-8, mxrs = 1.0, ax = 1, ay = 1, az = 1;
   int difconv = 0, atype = 0, solver_id = 1, k_dim = 5, rotate = 0, xisone = 0;
   double rot_alpha = 0., rot_eps = 1.;
   const char *ofile = NULL;
   for (i = 1; i < argc; i++)
   {
      if (!strcmp(argv[i], "-n")) { nx = atoi(argv[++i]); ny = atoi(argv[++i]); nz = atoi(argv[++i]); }
      else if (!strcmp(argv[i], "-27pt")) pt27 = 1;
      else if (!strcmp(argv[i], "-c")) { cx = atof(argv[++i]); cy = atof(argv[++i]); cz = atof(argv[++i]); }
      else if (!strcmp(argv[i], "-P")) { g_P[0] = atoi(argv[++i]); g_P[1] = atoi(argv[++i]); g_P[2] = atoi(argv[++i]); }
      else if (!strcmp(argv[i], "-xisone")) xisone = 1;                       /* ij.c:629: b = A * ones (solution of all ones) */
      else if (!strcmp(argv[i], "-solver")) solver_id = atoi(argv[++i]);     /* 0 AMG, 1 AMG-PCG, 3 AMG-GMRES, 9 AMG-BiCGSTAB */
      else if (!strcmp(argv[i], "-k")) k_dim = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-rotate")) rotate = 1;
      else if (!strcmp(argv[i], "-alpha")) rot_alpha = atof(argv[++i]);
      else if (!strcmp(argv[i], "-eps")) rot_eps = atof(argv[++i]);
      else if (!strcmp(argv[i], "-difconv")) difconv = 1;
      else if (!strcmp(argv[i], "-a")) { ax = atof(argv[++i]); ay = atof(argv[++i]); az = atof(argv[++i]); }
      else if (!strcmp(argv[i], "-atype")) atype = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-Pmx")) Pmx = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-th")) th = atof(argv[++i]);
      else if (!strcmp(argv[i], "-tol")) tol = atof(argv[++i]);
      else if (!strcmp(argv[i], "-mxrs")) mxrs = atof(argv[++i]);
      else if (!strcmp(argv[i], "-max_iter")) max_iter = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-matvec")) matvec_reps = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-o")) ofile = argv[++i];
      else if (!strcmp(argv[i], "-pmis")) g_coarsen_type = 8;
      else if (!strcmp(argv[i], "-hmis")) g_coarsen_type = 10;                           /* the driver default when -pmis is absent */
      else if (!strcmp(argv[i], "-nodump")) { }
      else if (!strcmp(argv[i], "-rlx")) rlx = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-gs_blocks")) g_gs_blocks = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-agg_nl")) g_agg_nl = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-mod_rap2")) g_mod_rap2 = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-keepT")) { ++i; }
      else if (!strcmp(argv[i], "-ns")) { g_ns[1] = g_ns[2] = atoi(argv[++i]); }          /* ij.c:881-885 -> SetNumSweeps */
      else if (!strcmp(argv[i], "-ns_down")) g_ns[1] = atoi(argv[++i]);                  /* ij.c:891-900 -> SetCycleNumSweeps */
      else if (!strcmp(argv[i], "-ns_up")) g_ns[2] = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-ns_coarse")) g_ns[3] = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-mu")) g_cycle_type = atoi(argv[++i]);                  /* ij.c:1490 */
      else if (!strcmp(argv[i], "-perturb")) perturb = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-fmg")) g_fcycle = 1;                                   /* ij.c:1495-1499 */
      else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
   }
   double v[7];
   if (pt27) { v[0] = 26.0; if (nx == 1 || ny == 1 || nz == 1) v[0] = 8.0; if (nx * ny == 1 || nx * nz == 1 || ny * nz == 1) v[0] = 2.0; v[1] = -1.; }
   else { v[1] = -cx; v[2] = -cy; v[3] = -cz; v[0] = 0.; if (nx > 1) v[0] += 2.0 * cx; if (ny > 1) v[0] += 2.0 * cy; if (nz > 1) v[0] += 2.0 * cz; v[4] = v[1]; v[5] = v[2]; v[6] = v[3]; }
   if (difconv && !pt27) { double c[3] = { cx, cy, cz }, a[3] = { ax, ay, az }; difconv_values(nx, ny, nz, c, a, atype, v); }
   if (rotate)
   {  /* par_rotate_7pt.c:62-73 */
      double pi = 4.0 * atan(1.0), xr = pi * rot_alpha / 180.0, sn = sin(xr), cs = cos(xr);
      double ac = -(cs * cs + rot_eps * sn * sn), bc = 2.0 * (1.0 - rot_eps) * sn * cs, cc = -(sn * sn + rot_eps * cs * cs);
      if (nz != 1) { fprintf(stderr, "-rotate is two-dimensional: -n nx ny 1\n"); return 2; }
      v[0] = -2 * (2 * ac + bc + 2 * cc); v[1] = 2 * ac + bc; v[2] = bc + 2 * cc; v[3] = -bc;
   }
   csr_t A = gen_laplace(nx, ny, nz, rotate ? 2 : pt27, v);
   if (g_P[0] * g_P[1] * g_P[2] > 1)
   {  /* -P: one Gauss-Seidel block per rank, in rank order (p fastest): the boxes of the process grid */
      int T = g_P[0] * g_P[1] * g_P[2], k = 0, p, q, r, lo, hi, sx, sy, sz;
      g_gs_blocks = T;
      g_blk[0] = (int *) xcalloc(T + 1, sizeof(int));
      for (r = 0; r < g_P[2]; r++) for (q = 0; q < g_P[1]; q++) for (p = 0; p < g_P[0]; p++, k++)
      {
         part1d(nx, g_P[0], p, &lo, &hi); sx = hi - lo; part1d(ny, g_P[1], q, &lo, &hi); sy = hi - lo; part1d(nz, g_P[2], r, &lo, &hi); sz = hi - lo;
         g_blk[0][k + 1] = g_blk[0][k] + sx * sy * sz;
      }
   }
   int N = A.n;
   if (perturb) perturb_operator(N, A.i, A.j, A.a, (unsigned) perturb);
   double *b = (double *) xmalloc(sizeof(double) * N), *x = (double *) xcalloc(N, sizeof(double));
   for (i = 0; i < N; i++) b[i] = 1.0;
   if (xisone) { double *ones = (double *) xmalloc(sizeof(double) * N); for (i = 0; i < N; i++) ones[i] = 1.0; matvec(1.0, &A, ones, 0.0, b, b); free(ones); }
   if (matvec_reps > 0)
   {
      double *y = (double *) xcalloc(N, sizeof(double)), t0;
      matvec(1.0, &A, b, 0.0, y, y); t0 = now();
      for (i = 0; i < matvec_reps; i++) matvec(1.0, &A, b, 0.0, y, y);
      printf("amg_oracle: matvec_ms=%.6f reps=%d\n", (now() - t0) / matvec_reps * 1e3, matvec_reps);
      return 0;
   }
   if (rlx > -1) g_relax_down = g_relax_up = rlx; else { g_relax_down = 13; g_relax_up = 14; }
   g_user_relax = rlx;
   amg_t g;
   double t0 = now();
   amg_setup(&g, A, th, mxrs, Pmx, 9);
   double t_setup = now() - t0;

   /* PCG: krylov/pcg.c:347-757, two_norm 1 */
   double *p = (double *) xcalloc(N, sizeof(double)), *s = (double *) xcalloc(N, sizeof(double)), *r = (double *) xmalloc(sizeof(double) * N);
   double *norms = (double *) xcalloc(max_iter + 2, sizeof(double));
   t0 = now();
   double bi_prod = dot(N, b, b), eps = tol * tol, i_prod = 0, gamma, gamma_old, krylov_rel = 0;
   int it = 0;
   if (solver_id == 0)
   {  /* BoomerAMG as the solver: hypre_BoomerAMGSolve (par_amg_solve.c:143-311), cycles until ||f - A u|| / ||f|| < tol */
      double rhs_norm = sqrt(dot(N, b, b)), resid, rel = 1.0;
      memcpy(r, b, sizeof(double) * N); matvec(-1.0, &A, x, 1.0, r, r);
      resid = sqrt(dot(N, r, r)); rel = rhs_norm ? resid / rhs_norm : resid;
      norms[0] = resid;
      double resid_init = resid;
      while (rel >= tol && it < max_iter)
      {
         g_cycle_op_count = 0;                                            /* par_amg_solve.c:245 */
         cycle(&g, b, x);
         memcpy(r, b, sizeof(double) * N); matvec(-1.0, &A, x, 1.0, r, r);
         resid = sqrt(dot(N, r, r)); rel = rhs_norm ? resid / rhs_norm : resid;
         norms[++it] = resid;
      }
      krylov_rel = rel;
      {  /* closing statistics of hypre_BoomerAMGSolve (par_amg_solve.c:327-391) */
         double tc = 0, tv = 0;
         for (i = 0; i < g.nl; i++) { tc += (double) g.A[i].i[g.A[i].n]; tv += (double) g.A[i].n; }
         printf("amg_oracle: conv_factor=%f grid=%f operator=%f cycle=%f\n",
                (it > 0 && resid_init) ? pow(resid / resid_init, 1.0 / (double) it) : 1.0, tv / (double) N, tc / (double) A.nnz,
                g_cycle_op_count / (double) A.nnz);
      }
      goto solved;
   }
   if (solver_id == 3) { it = gmres(&g, &A, b, x, k_dim, tol, max_iter, norms, &krylov_rel); goto solved; }
   if (solver_id == 9) { it = bicgstab(&g, &A, b, x, tol, max_iter, norms, &krylov_rel); goto solved; }
   memcpy(r, b, sizeof(double) * N);
   matvec(-1.0, &A, x, 1.0, r, r);
   memset(p, 0, sizeof(double) * N); cycle(&g, r, p);
   gamma = dot(N, r, p);
   norms[0] = sqrt(dot(N, r, r));
   while (it + 1 <= max_iter)
   {
      it++;
      matvec(1.0, &A, p, 0.0, s, s);
      double sdotp = dot(N, s, p);
      if (sdotp == 0.0) break;
      double alpha = gamma / sdotp;
      gamma_old = gamma;
      for (i = 0; i < N; i++) x[i] += alpha * p[i];
      for (i = 0; i < N; i++) r[i] += -alpha * s[i];
      memset(s, 0, sizeof(double) * N); cycle(&g, r, s);
      gamma = dot(N, r, s);
      i_prod = dot(N, r, r);
      norms[it] = sqrt(i_prod);
      if (i_prod / bi_prod < eps) break;
      double beta = gamma / gamma_old;
      for (i = 0; i < N; i++) p[i] *= beta;
      for (i = 0; i < N; i++) p[i] += 1.0 * s[i];
   }
solved: ;
   double t_solve = now() - t0, relres = solver_id == 1 ? sqrt(i_prod / bi_prod) : krylov_rel;
   printf("amg_oracle: n=%d %d %d rows=%d nnz=%d\n", nx, ny, nz, N, A.nnz);
   printf("amg_oracle: levels=%d iterations=%d relres=%.6e setup_s=%.4f solve_s=%.4f\n", g.nl, it, relres, t_setup, t_solve);
   for (i = 0; i < g.nl; i++) printf("amg_oracle: level %d rows=%d nnz=%d\n", i, g.A[i].n, g.A[i].i[g.A[i].n]);
   if (ofile)
   {
      g_out = fopen(ofile, "wb");
      int hdr[8] = { nx, ny, nz, g.nl, it, pt27, Pmx, rlx };
      put("hdr", 0, hdr, 8); put("relres", 1, &relres, 1); put("norms", 1, norms, it + 1); put("x", 1, x, N);
      for (i = 0; i < g.nl; i++)
      {
         char nm[64];
         put_csr("A", i, &g.A[i], 1);
         if ((i < g.nl - 1 || !g.ge || g_relax_down == 7) && (g_relax_down == 18 || g_relax_down == 8 || g_relax_down == 13 || g_relax_down == 14 || g_relax_down == 7))
         { sprintf(nm, "l1_%d", i); put(nm, 1, g.l1[i], g.A[i].n); }   /* par_amg_setup.c:3045-3060 */
         if (i < g.nl - 1)
         {
            sprintf(nm, "CF%d", i); put(nm, 0, g.cf[i], g.A[i].n);
            put_csr("P", i, &g.P[i], 1);
            put_csr("S", i, &g.S[i], 0);
         }
      }
      fclose(g_out);
   }
   return 0;
}
