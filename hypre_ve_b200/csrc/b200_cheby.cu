// b200_cheby.cu -- Chebyshev polynomial smoother (hypre_BoomerAMGRelax type 16).
//
// Reference: hypre_ParCSRMaxEigEstimateCG (parcsr_ls/par_relax_more.c:115-330), hypre_ParCSRRelax_Cheby_Setup /
// _Solve (parcsr_ls/par_cheby.c:41-345), call sites par_amg_setup.c:3137-3160 and par_cycle.c:440-452.
//
// A sweep of order k is k applications of A plus fused vector passes -- pure HBM streaming through the SpMV kernel,
// no dependency chain: the bandwidth-bound "strong" smoother of this path.  The spectrum estimate is the reference's:
// `eig_est` steps of (diagonally scaled) CG from the reference's random vector (hypre_SeedRand(1), one draw per row,
// evaluated per row by LCG jump-ahead), the Lanczos tridiagonal matrix, and its extreme eigenvalues.  The tridiagonal
// matrix is at most 10 x 10: its eigenvalues and the polynomial coefficients are host scalars, like the PCG scalars.
#include "b200_internal.h"
#include <algorithm>
#include <cmath>

int b200_csr_spmv_epi(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha, double beta,
                      const double *b, const double *d);

struct b200_cheby_s {
  int n = 0, order = 2, scale = 1;
  double coefs[5] = {0, 0, 0, 0, 0};
  double max_eig = 0, min_eig = 0;
  double *ds = nullptr;                                   // D^{-1/2}
  double *r = nullptr, *v = nullptr, *tmp = nullptr, *orig = nullptr;   // work vectors of the sweep
};

namespace {

__device__ __forceinline__ unsigned long long mulmod31c(unsigned long long a, unsigned long long b) { return (a * b) % 2147483647ULL; }
// state of hypre_Rand's generator after k calls (utilities/random.c:77-106: x <- 16807 x mod 2^31-1)
__device__ __forceinline__ int lcg_after(int seed, unsigned long long k) {
  unsigned long long base = 16807ULL, acc = (unsigned long long)seed;
  while (k) {
    if (k & 1ULL) acc = mulmod31c(acc, base);
    base = mulmod31c(base, base);
    k >>= 1;
  }
  return (int)acc;
}
inline int cgrid(b200_handle h, size_t n) {
  size_t g = (n + 255) / 256, cap = (size_t)h->num_sm * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
#define GRID_STRIDE(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

__global__ void random_kernel(size_t n, int seed, double *__restrict__ r) {           // seq_mv/vector.c:286-300
  GRID_STRIDE(i, n) r[i] = 2.0 * ((double)lcg_after(seed, i + 1) / 2147483647.0) - 1.0;
}
__global__ void ds_kernel(size_t n, int scale, const int *__restrict__ A_i, const double *__restrict__ A_a, double *__restrict__ ds) {
  GRID_STRIDE(i, n) ds[i] = scale ? 1 / sqrt(A_a[A_i[i]]) : 1.0;
}
__global__ void p_update_kernel(size_t n, int first, double beta, const double *__restrict__ s, double *__restrict__ p,
                                const double *__restrict__ ds, double *__restrict__ u) {
  GRID_STRIDE(i, n) {
    const double pv = first ? s[i] : s[i] + beta * p[i];
    p[i] = pv;
    u[i] = ds[i] * pv;
  }
}
__global__ void scale_kernel(size_t n, const double *__restrict__ ds, double *__restrict__ s) {
  GRID_STRIDE(i, n) s[i] = ds[i] * s[i];
}
__global__ void axpy_kernel(size_t n, double a, const double *__restrict__ x, double *__restrict__ y) {
  GRID_STRIDE(i, n) y[i] += a * x[i];
}
// sweep, scaled form (par_cheby.c:266-335)
__global__ void start_kernel(size_t n, int zero, double c_top, const double *__restrict__ f, const double *tmp,
                             const double *__restrict__ ds, double *__restrict__ r, double *__restrict__ orig, double *__restrict__ u,
                             double *t2) {                       // t2 may alias tmp (element i is read before it is written)
  GRID_STRIDE(i, n) {
    const double rv = ds[i] * (f[i] + (zero ? 0.0 : tmp[i]));       // r = D^{-1/2} (f - A u)
    r[i] = rv;
    orig[i] = zero ? 0.0 : u[i];
    const double uv = rv * c_top;
    u[i] = uv;
    t2[i] = ds[i] * uv;                                             // input of the next application of A
  }
}
__global__ void step_kernel(size_t n, int last, double mult, const double *__restrict__ r, const double *__restrict__ ds,
                            const double *__restrict__ v, const double *__restrict__ orig, double *__restrict__ u,
                            double *__restrict__ t2) {
  GRID_STRIDE(i, n) {
    const double uv = mult * r[i] + ds[i] * v[i];
    if (last) u[i] = orig[i] + ds[i] * uv;
    else { u[i] = uv; t2[i] = ds[i] * uv; }
  }
}
__global__ void finish_kernel(size_t n, const double *__restrict__ orig, const double *__restrict__ ds, double *__restrict__ u) {
  GRID_STRIDE(i, n) u[i] = orig[i] + ds[i] * u[i];
}

// eigenvalues of the symmetric tridiagonal matrix (d[0..n-1]; e[i] couples i-1 and i), ascending, implicit QL.
// The reference calls EISPACK's tql1 (hypre_LINPACKcgtql1); both are backward stable, the extreme eigenvalues agree to
// a few ulp, which is all the Chebyshev interval needs.
void tridiag_eigenvalues(int n, double *d, double *e) {
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  if (n > 0) e[n - 1] = 0.0;
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) break;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]), r = std::sqrt(g * g + 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double sn = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          const double f = sn * e[i], b = c * e[i];
          r = std::sqrt(f * f + g * g);
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
  std::sort(d, d + n);
}

}  // namespace

int b200_cheby_destroy(b200_handle h, b200_cheby_s *c) {
  if (!c) return 0;
  B200_TRY(b200_dfree(h, c->ds)); B200_TRY(b200_dfree(h, c->r)); B200_TRY(b200_dfree(h, c->v));
  B200_TRY(b200_dfree(h, c->tmp)); B200_TRY(b200_dfree(h, c->orig));
  delete c;
  return 0;
}

// hypre_ParCSRMaxEigEstimateCG + hypre_ParCSRRelax_Cheby_Setup for one level
int b200_cheby_setup(b200_handle h, b200_csr A, int eig_est, int order, double fraction, int variant, int scale,
                     b200_cheby_s **out) {
  if (eig_est <= 0) B200_FAIL("Chebyshev: ChebyEigEst must be > 0 (the power-iteration estimate is not implemented)");
  const int n = A->nrows;
  b200_cheby_s *C = new b200_cheby_s();
  C->n = n; C->scale = scale;
  C->order = std::min(4, std::max(1, order));
  *out = C;
  for (double **v : {&C->ds, &C->r, &C->v, &C->tmp, &C->orig}) B200_TRY(b200_dalloc<double>(h, v, (size_t)n + 1));
  double *p = nullptr;
  B200_TRY(b200_dalloc<double>(h, &p, (size_t)n + 1));
  double *r = C->r, *s = C->v, *u = C->tmp;
  const int g = cgrid(h, n);
  random_kernel<<<g, 256, 0, h->stream>>>((size_t)n, 1, r);
  B200_LAUNCH_CHECK();
  ds_kernel<<<g, 256, 0, h->stream>>>((size_t)n, scale, A->i, A->a, C->ds);
  B200_LAUNCH_CHECK();
  const int max_iter = std::min(eig_est, n);
  std::vector<double> td((size_t)max_iter + 1, 0.0), to((size_t)max_iter + 1, 0.0);
  double gamma = 0.0, gamma_old, beta = 1.0;
  int i = 0;
  while (i < max_iter) {
    gamma_old = gamma;
    B200_TRY(b200_vec_dot(h, n, r, r, &gamma));                              // s = C r = r ; gamma = <r, s>
    beta = i == 0 ? 1.0 : gamma / gamma_old;
    p_update_kernel<<<g, 256, 0, h->stream>>>((size_t)n, i == 0, beta, r, p, C->ds, u);   // p = s + beta p ; u = D^{-1/2} p
    B200_LAUNCH_CHECK();
    B200_TRY(b200_csr_spmv_epi(h, A, u, s, 0, 1.0, 0.0, nullptr, nullptr));     // s = A u
    scale_kernel<<<g, 256, 0, h->stream>>>((size_t)n, C->ds, s);                // s = D^{-1/2} A D^{-1/2} p
    B200_LAUNCH_CHECK();
    double sdotp = 0.0;
    B200_TRY(b200_vec_dot(h, n, s, p, &sdotp));
    const double alpha = gamma / sdotp, alphainv = 1.0 / alpha;
    td[i + 1] = alphainv; td[i] *= beta; td[i] += alphainv;
    to[i + 1] = alphainv; to[i] *= std::sqrt(beta);
    axpy_kernel<<<g, 256, 0, h->stream>>>((size_t)n, -alpha, s, r);
    B200_LAUNCH_CHECK();
    i++;
  }
  B200_TRY(b200_dfree(h, p));
  tridiag_eigenvalues(i, td.data(), to.data());
  C->max_eig = i ? td[i - 1] : 0.0;
  C->min_eig = i ? td[0] : 0.0;
  // coefficients of u = u_0 + s(A) r_0, s of degree order-1 (par_cheby.c:74-160)
  const int co = C->order - 1;
  const double ub = C->max_eig * 1.1, lb = (ub - C->min_eig) * fraction + C->min_eig;
  const double theta = (ub + lb) / 2, delta = (ub - lb) / 2;
  double den, *c = C->coefs;
  if (variant == 1) {
    switch (co) {
      case 0: c[0] = 1.0 / theta; break;
      case 1: den = (theta * theta + delta * theta); c[0] = (delta + 2 * theta) / den; c[1] = -1.0 / den; break;
      case 2:
        den = 2 * delta * theta * theta - delta * delta * theta - pow(delta, 3) + 2 * pow(theta, 3);
        c[0] = (4 * delta * theta - pow(delta, 2) + 6 * pow(theta, 2)) / den; c[1] = -(2 * delta + 6 * theta) / den; c[2] = 2 / den;
        break;
      default:
        den = -(4 * delta * pow(theta, 3) - 3 * pow(delta, 2) * pow(theta, 2) - 3 * pow(delta, 3) * theta + 4 * pow(theta, 4));
        c[0] = (6 * pow(delta, 2) * theta - 12 * delta * pow(theta, 2) + 3 * pow(delta, 3) - 16 * pow(theta, 3)) / den;
        c[1] = (12 * delta * theta - 3 * pow(delta, 2) + 24 * pow(theta, 2)) / den; c[2] = -(4 * delta + 16 * theta) / den; c[3] = 4 / den;
    }
  } else {
    switch (co) {
      case 0: c[0] = 1.0 / theta; break;
      case 1: den = delta * delta - 2 * theta * theta; c[0] = -4 * theta / den; c[1] = 2 / den; break;
      case 2:
        den = 3 * (delta * delta) * theta - 4 * (theta * theta * theta);
        c[0] = (3 * delta * delta - 12 * theta * theta) / den; c[1] = 12 * theta / den; c[2] = -4 / den;
        break;
      default:
        den = pow(delta, 4) - 8 * delta * delta * theta * theta + 8 * pow(theta, 4);
        c[0] = (32 * pow(theta, 3) - 16 * delta * delta * theta) / den; c[1] = (8 * delta * delta - 48 * theta * theta) / den;
        c[2] = 32 * theta / den; c[3] = -8 / den;
    }
  }
  return 0;
}

// hypre_ParCSRRelax_Cheby_Solve, in place on u.  As (the operator of the sweep) may be a column-sorted copy of A.
// The unscaled form (ChebyScale 0) runs the same kernels with ds == 1.
int b200_cheby_solve(b200_handle h, b200_cheby_s *C, b200_csr As, bool zero, const double *f, double *u) {
  const int n = C->n;
  if (n == 0) return 0;
  const int g = cgrid(h, n), co = C->order - 1;
  if (!zero) B200_TRY(b200_csr_spmv_epi(h, As, u, C->tmp, 0, -1.0, 0.0, nullptr, nullptr));   // tmp = -A u
  // r = D^{-1/2}(f + tmp) ; orig = u ; u = r * c[co] ; tmp = D^{-1/2} u
  start_kernel<<<g, 256, 0, h->stream>>>((size_t)n, zero ? 1 : 0, C->coefs[co], f, C->tmp, C->ds, C->r, C->orig, u, C->tmp);
  B200_LAUNCH_CHECK();
  if (co == 0) {
    finish_kernel<<<g, 256, 0, h->stream>>>((size_t)n, C->orig, C->ds, u);
    B200_LAUNCH_CHECK();
    return 0;
  }
  for (int i = co - 1; i >= 0; i--) {
    B200_TRY(b200_csr_spmv_epi(h, As, C->tmp, C->v, 0, 1.0, 0.0, nullptr, nullptr));          // v = A D^{-1/2} u
    step_kernel<<<g, 256, 0, h->stream>>>((size_t)n, i == 0, C->coefs[i], C->r, C->ds, C->v, C->orig, u, C->tmp);
    B200_LAUNCH_CHECK();
  }
  return 0;
}
