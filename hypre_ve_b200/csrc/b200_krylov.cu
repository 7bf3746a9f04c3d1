// b200_krylov.cu -- the Krylov drivers for nonsymmetric operators over the same SpMV / BLAS-1 kernels as PCG:
//   restarted GMRES(k) with right preconditioning   hypre_GMRESSolve     (krylov/gmres.c:226-800)
//   BiCGSTAB                                        hypre_BiCGSTABSolve  (krylov/bicgstab.c:207-530)
// (`ij -solver 3` AMG-GMRES, `-solver 9` AMG-BiCGSTAB, ij.c:5298-5330, :6364-6380).
//
// Vectors stay in HBM; the Hessenberg matrix / Givens rotations (k_dim <= 60 numbers) are host arithmetic as in the
// reference.  On one GPU an Arnoldi step enqueues the whole modified Gram-Schmidt sweep -- dot_j -> axpy_j, the
// coefficient read from device memory by the axpy -- and reads the i+1 inner products back with ONE stream
// synchronisation; across ranks every inner product is one reduction (b200_dist.cu supplies the operations).
// BiCGSTAB needs three synchronisations per iteration (alpha, gamma, the norms).  Operation order follows the
// reference line by line so that iteration counts are equal and residual histories agree to 1e-10 (dot products are
// tree sums).  The two loops are written once (b200_gmres_core / b200_bicgstab_core) over b200_krylov_ops.
#include <cmath>
#include <vector>
#include "b200_internal.h"

int b200_vec_dot_dev(b200_handle h, int n, const double *x, const double *y, double *d_out);   // b200_vec.cu
int b200_amg_precond(b200_handle h, b200_amg amg, const double *d_rhs, double *d_out);        // b200_amg.cu

namespace {
constexpr int VT = 256;
// y += (sign * *d_a) * x : Axpy with the coefficient still on the device (gmres.c:477-478)
__global__ void axpy_dev_kernel(size_t n, double sign, const double *__restrict__ d_a, const double *__restrict__ x,
                                double *__restrict__ y) {
  const double a = sign * d_a[0];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] += a * x[i];
}
// y += a * y : hypre_SeqVectorAxpy called with x == y (gmres.c:761,:766)
__global__ void axpy_self_kernel(size_t n, double a, double *y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] += a * y[i];
}
inline int vgrid(b200_handle h, size_t n) {
  size_t g = (n + VT - 1) / VT, cap = (size_t)h->num_sm * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// the vectors and the scalar scratch of one solve; operator, preconditioner and inner products come from b200_krylov_ops
// (single GPU below, row-partitioned in b200_dist.cu) so that both run the same restatement of the reference loops
struct Krylov {
  b200_handle h;
  const b200_krylov_ops *ops;
  int n;
  double *sc = nullptr;          // device scalars
  std::vector<double *> owned;
  int alloc(double **p) {
    B200_TRY(b200_dalloc<double>(h, p, (size_t)ops->cap));
    B200_CUDA(cudaMemsetAsync(*p, 0, sizeof(double) * (size_t)ops->cap, h->stream));
    owned.push_back(*p);
    return 0;
  }
  void release() {
    for (double *p : owned) b200_dfree(h, p);
    owned.clear();
    b200_dfree(h, sc);
    sc = nullptr;
  }
  int precond(const double *rhs, double *out) { return ops->precond(rhs, out); }     // ClearVector(out); precond(A, rhs, out)
  // k inner products, results on the host after ONE synchronisation / reduction over the ranks
  int dots(int k, const double *const *xs, const double *const *ys, double *out) {
    for (int i = 0; i < k; i++) B200_TRY(b200_vec_dot_dev(h, n, xs[i], ys[i], sc + i));
    return ops->reduce(sc, k, out);
  }
  int dot1(const double *x, const double *y, double *out) { return dots(1, &x, &y, out); }
  int dot2(const double *x1, const double *y1, const double *x2, const double *y2, double *out) {
    const double *xs[2] = {x1, x2}, *ys[2] = {y1, y2};
    return dots(2, xs, ys, out);
  }
  int residual(const double *b, const double *x, double *r) {      // CopyVector(b, r); Matvec(-1, A, x, 1, r)
    return ops->matvec(-1.0, x, 1.0, b, r);
  }
  int matvec(const double *x, double *y) { return ops->matvec(1.0, x, 0.0, nullptr, y); }
};

}  // namespace

int b200_gmres_core(b200_handle h, const b200_krylov_ops *ops, const b200_gmres_params *prm, const double *d_b, double *d_x_user,
                    int *iters_out, double *final_rel_res, double *h_norms, int *converged_out) {
  if (!prm || !ops) B200_FAIL("gmres: null argument");
  if (prm->rel_change) B200_FAIL("gmres: rel_change is not implemented");
  if (prm->cf_tol > 0.0) B200_FAIL("gmres: convergence-factor tolerance is not implemented");
  const int k_dim = prm->k_dim, max_iter = prm->max_iter, min_iter = prm->min_iter;
  if (k_dim < 1 || k_dim > 60) B200_FAIL("gmres: k_dim must be in 1..60");
  Krylov K{h, ops, ops->n};
  const int n = K.n;
  std::vector<double *> p(k_dim + 1, nullptr);
  double *r = nullptr, *w = nullptr, *d_x = d_x_user;
  std::vector<double> rs(k_dim + 1, 0.0), c(k_dim, 0.0), s(k_dim, 0.0), hs(k_dim + 2, 0.0);
  std::vector<std::vector<double>> hh(k_dim + 1, std::vector<double>(k_dim, 0.0));
  const double epsmac = 1.e-16;
  int rc = 0, iter = 0, i = 0, converged = 0;
  double r_norm = 0, b_norm = 0, epsilon = 0, real_r_norm_old = 0, real_r_norm_new = 0, t = 0;
  do {
    if ((rc = b200_dalloc<double>(h, &K.sc, 64))) break;
    if ((rc = K.alloc(&r)) || (rc = K.alloc(&w))) break;
    for (int j = 0; j <= k_dim && !rc; j++) rc = K.alloc(&p[j]);
    if (rc) break;
    if (ops->cap > n) {                                     // the operator reads a ghost tail behind its input: iterate on a padded copy
      if ((rc = K.alloc(&d_x))) break;
      if ((rc = b200_vec_copy(h, n, d_x_user, d_x))) break;
    }
    if ((rc = K.residual(d_b, d_x, p[0]))) break;                              // :316-319
    if ((rc = K.dot2(d_b, d_b, p[0], p[0], hs.data()))) break;                 // b_norm :321, r_norm :347
    b_norm = std::sqrt(hs[0]);
    real_r_norm_old = b_norm;
    r_norm = std::sqrt(hs[1]);
    if ((b_norm != 0. && !(b_norm / b_norm == b_norm / b_norm)) || (r_norm != 0. && !(r_norm / r_norm == r_norm / r_norm))) {
      rc = b200_set_error(__FILE__, __LINE__, "hypre_GMRESSolve: INFs and/or NaNs detected in input");   // :326-372
      break;
    }
    if (h_norms) h_norms[0] = r_norm;
    const double den_norm = b_norm > 0.0 ? b_norm : r_norm;                    // :388-394
    epsilon = std::fmax(prm->a_tol, prm->tol * den_norm);                      // :403
    while (iter < max_iter) {                                                  // :423 outer cycle
      rs[0] = r_norm;
      if (r_norm == 0.0) break;                                                // :427-439
      if (r_norm <= epsilon && iter >= min_iter) {                             // :443-462 already converged?
        if ((rc = K.residual(d_b, d_x, r))) break;
        if ((rc = K.dot1(r, r, &t))) break;
        r_norm = std::sqrt(t);
        if (r_norm <= epsilon) break;
      }
      if ((rc = b200_vec_scale(h, n, 1.0 / r_norm, p[0]))) break;              // :464-465
      i = 0;
      while (i < k_dim && iter < max_iter) {                                   // :469 restart cycle
        i++;
        iter++;
        if ((rc = K.precond(p[i - 1], r))) break;                              // :472-473
        if ((rc = K.matvec(r, p[i]))) break;                                   // :474
        if (ops->device_mgs) {                                                 // modified Gram-Schmidt :476-480, one synchronisation
          for (int j = 0; j < i && !rc; j++) {
            if ((rc = b200_vec_dot_dev(h, n, p[j], p[i], K.sc + j))) break;
            axpy_dev_kernel<<<vgrid(h, n), VT, 0, h->stream>>>((size_t)n, -1.0, K.sc + j, p[j], p[i]);
            ++g_b200_launches;
          }
          if (rc) break;
          if ((rc = b200_vec_dot_dev(h, n, p[i], p[i], K.sc + i))) break;
          if ((rc = ops->reduce(K.sc, i + 1, hs.data()))) break;
        } else {                                                               // the same sweep with every product reduced over the ranks
          for (int j = 0; j < i && !rc; j++) {
            if ((rc = K.dot1(p[j], p[i], &hs[j]))) break;
            rc = b200_vec_axpy(h, n, -hs[j], p[j], p[i]);
          }
          if (rc) break;
          if ((rc = K.dot1(p[i], p[i], &hs[i]))) break;
        }
        for (int j = 0; j < i; j++) hh[j][i - 1] = hs[j];
        t = std::sqrt(hs[i]);
        hh[i][i - 1] = t;
        if (t != 0.0) {
          t = 1.0 / t;
          if ((rc = b200_vec_scale(h, n, t, p[i]))) break;                     // :482-485
        }
        for (int j = 1; j < i; j++) {                                          // :488-492 apply the earlier rotations
          t = hh[j - 1][i - 1];
          hh[j - 1][i - 1] = s[j - 1] * hh[j][i - 1] + c[j - 1] * t;
          hh[j][i - 1] = -s[j - 1] * t + c[j - 1] * hh[j][i - 1];
        }
        t = hh[i][i - 1] * hh[i][i - 1];
        t += hh[i - 1][i - 1] * hh[i - 1][i - 1];
        double gamma = std::sqrt(t);
        if (gamma == 0.0) gamma = epsmac;
        c[i - 1] = hh[i - 1][i - 1] / gamma;
        s[i - 1] = hh[i][i - 1] / gamma;
        rs[i] = -hh[i][i - 1] * rs[i - 1];
        rs[i] /= gamma;
        rs[i - 1] = c[i - 1] * rs[i - 1];
        hh[i - 1][i - 1] = s[i - 1] * hh[i][i - 1] + c[i - 1] * hh[i - 1][i - 1];   // :508
        r_norm = std::fabs(rs[i]);
        if (h_norms) h_norms[iter] = r_norm;                                   // :513
        if (r_norm <= epsilon && iter >= min_iter) break;                      // :541, no relative-change test
      }
      if (rc) break;
      rs[i - 1] = rs[i - 1] / hh[i - 1][i - 1];                                // :641-649 triangular solve
      for (int k = i - 2; k >= 0; k--) {
        t = 0.0;
        for (int j = k + 1; j < i; j++) t -= hh[k][j] * rs[j];
        t += rs[k];
        rs[k] = t / hh[k][k];
      }
      if ((rc = b200_vec_copy(h, n, p[i - 1], w))) break;                      // :651-654
      if ((rc = b200_vec_scale(h, n, rs[i - 1], w))) break;
      for (int j = i - 2; j >= 0 && !rc; j--) rc = b200_vec_axpy(h, n, rs[j], p[j], w);
      if (rc) break;
      if ((rc = K.precond(w, r))) break;                                       // :656-658 correction
      if ((rc = b200_vec_axpy(h, n, 1.0, r, d_x))) break;                      // :661
      if (r_norm <= epsilon && iter >= min_iter) {                             // :664-752 check the true residual
        if (prm->skip_real_r_check) { converged = 1; break; }
        if ((rc = K.residual(d_b, d_x, r))) break;
        if ((rc = K.dot1(r, r, &t))) break;
        real_r_norm_new = r_norm = std::sqrt(t);
        if (r_norm <= epsilon) { converged = 1; break; }
        if (real_r_norm_new >= real_r_norm_old) { converged = 1; break; }      // :733-741 no progress in the true residual
        if ((rc = b200_vec_copy(h, n, r, p[0]))) break;                        // :748-750 "false convergence 2": restart from r
        i = 0;
        real_r_norm_old = real_r_norm_new;
      }
      for (int j = i; j > 0; j--) {                                            // :755-758 residual vector of the restart
        rs[j - 1] = -s[j - 1] * rs[j];
        rs[j] = c[j - 1] * rs[j];
      }
      if (i) {
        axpy_self_kernel<<<vgrid(h, n), VT, 0, h->stream>>>((size_t)n, rs[i] - 1.0, p[i]);   // :760-761
        ++g_b200_launches;
      }
      for (int j = i - 1; j > 0 && !rc; j--) rc = b200_vec_axpy(h, n, rs[j], p[j], p[i]);
      if (rc) break;
      if (i) {
        axpy_self_kernel<<<vgrid(h, n), VT, 0, h->stream>>>((size_t)n, rs[0] - 1.0, p[0]);   // :765-768
        ++g_b200_launches;
        if ((rc = b200_vec_axpy(h, n, 1.0, p[i], p[0]))) break;
      }
    }
  } while (0);
  if (!rc && cudaGetLastError() != cudaSuccess) rc = b200_set_error(__FILE__, __LINE__, "gmres: kernel launch failed");
  if (!rc && d_x != d_x_user) rc = b200_vec_copy(h, n, d_x, d_x_user);
  if (!rc) {
    if (iters_out) *iters_out = iter;
    if (final_rel_res) *final_rel_res = b_norm > 0.0 ? r_norm / b_norm : r_norm;   // :777-783
    if (converged_out) *converged_out = converged;
  }
  K.release();
  return rc;
}

int b200_bicgstab_core(b200_handle h, const b200_krylov_ops *ops, const b200_bicgstab_params *prm, const double *d_b,
                       double *d_x_user, int *iters_out, double *final_rel_res, double *h_norms, int *converged_out) {
  if (!prm || !ops) B200_FAIL("bicgstab: null argument");
  if (prm->cf_tol > 0.0) B200_FAIL("bicgstab: convergence-factor tolerance is not implemented");
  const int max_iter = prm->max_iter, min_iter = prm->min_iter;
  Krylov K{h, ops, ops->n};
  const int n = K.n;
  double *r = nullptr, *r0 = nullptr, *s = nullptr, *v = nullptr, *p = nullptr, *q = nullptr, *d_x = d_x_user;
  const double epsmac = 2.2250738585072014e-308;       // HYPRE_REAL_MIN
  int rc = 0, iter = 0, converged = 0;
  double alpha = 0, beta = 0, gamma = 0, epsilon = 0, temp = 0, res = 0, r_norm = 0, b_norm = 0, hs[4] = {0, 0, 0, 0};
  do {
    if ((rc = b200_dalloc<double>(h, &K.sc, 64))) break;
    if ((rc = K.alloc(&r)) || (rc = K.alloc(&r0)) || (rc = K.alloc(&s)) || (rc = K.alloc(&v)) || (rc = K.alloc(&p)) ||
        (rc = K.alloc(&q))) break;
    if (ops->cap > n) {
      if ((rc = K.alloc(&d_x))) break;
      if ((rc = b200_vec_copy(h, n, d_x_user, d_x))) break;
    }
    if ((rc = K.residual(d_b, d_x, r0))) break;                                // :269-273
    if ((rc = b200_vec_copy(h, n, r0, r))) break;
    if ((rc = b200_vec_copy(h, n, r0, p))) break;
    if ((rc = K.dot2(d_b, d_b, r0, r0, hs))) break;                            // :277, :303
    b_norm = std::sqrt(hs[0]);
    res = hs[1];
    r_norm = std::sqrt(res);
    if ((b_norm != 0. && !(b_norm / b_norm == b_norm / b_norm)) || (r_norm != 0. && !(r_norm / r_norm == r_norm / r_norm))) {
      rc = b200_set_error(__FILE__, __LINE__, "hypre_BiCGSTABSolve: INFs and/or NaNs detected in input");
      break;
    }
    if (h_norms) h_norms[0] = r_norm;
    const double den_norm = b_norm > 0.0 ? b_norm : r_norm;                    // :346-356
    if (prm->stop_crit) epsilon = prm->a_tol == 0.0 ? prm->tol : prm->a_tol;   // :359-367
    else epsilon = std::fmax(prm->a_tol, prm->tol * den_norm);                 // :378
    if (r_norm == 0.0) break;                                                  // :400-404
    if (r_norm <= epsilon && iter >= min_iter) { converged = 1; break; }       // :405-416
    while (iter < max_iter) {                                                  // :418
      iter++;
      if ((rc = K.precond(p, v))) break;                                       // :422-423
      if ((rc = K.matvec(v, q))) break;
      if ((rc = K.dot1(r0, q, &temp))) break;                                  // :425
      if (std::fabs(temp) >= epsmac) alpha = res / temp;
      else { rc = b200_set_error(__FILE__, __LINE__, "BiCGSTAB broke down!! divide by near zero"); break; }
      if ((rc = b200_vec_axpy(h, n, alpha, v, d_x))) break;                    // :435-436
      if ((rc = b200_vec_axpy(h, n, -alpha, q, r))) break;
      if ((rc = K.precond(r, v))) break;                                       // :437-439
      if ((rc = K.matvec(v, s))) break;
      if ((rc = K.dot2(r, s, s, s, hs))) break;                                // :441-442
      gamma = (hs[0] == 0.0 && hs[1] == 0.0) ? 0.0 : hs[0] / hs[1];            // :443-446
      if ((rc = b200_vec_axpy(h, n, gamma, v, d_x))) break;                    // :447-448
      if ((rc = b200_vec_axpy(h, n, -gamma, s, r))) break;
      if ((rc = K.dot2(r, r, r0, r, hs))) break;                               // :450; <r0,r> is the res of :507 (r unchanged unless the branch below runs)
      r_norm = std::sqrt(hs[0]);
      double res_new = hs[1];
      if (h_norms) h_norms[iter] = r_norm;
      if (r_norm <= epsilon && iter >= min_iter) {                             // :464-481 evaluate the actual residual
        if ((rc = K.residual(d_b, d_x, r))) break;
        if ((rc = K.dot1(r, r, &temp))) break;
        r_norm = std::sqrt(temp);
        if (r_norm <= epsilon) { converged = 1; break; }
        if ((rc = K.dot1(r0, r, &res_new))) break;                             // r was replaced by the true residual
      }
      if (std::fabs(res) >= epsmac) beta = 1.0 / res;                          // :499-506
      else { rc = b200_set_error(__FILE__, __LINE__, "BiCGSTAB broke down!! res=0"); break; }
      res = res_new;                                                           // :507-508
      beta *= res;
      if ((rc = b200_vec_axpy(h, n, -gamma, q, p))) break;                     // :509
      if (std::fabs(gamma) >= epsmac) { if ((rc = b200_vec_scale(h, n, beta * alpha / gamma, p))) break; }
      else { rc = b200_set_error(__FILE__, __LINE__, "BiCGSTAB broke down!! gamma=0"); break; }
      if ((rc = b200_vec_axpy(h, n, 1.0, r, p))) break;                        // :519
    }
  } while (0);
  if (!rc && d_x != d_x_user) rc = b200_vec_copy(h, n, d_x, d_x_user);
  if (!rc) {
    if (iters_out) *iters_out = iter;
    if (final_rel_res) *final_rel_res = b_norm > 0.0 ? r_norm / b_norm : r_norm;   // :522-526
    if (converged_out) *converged_out = converged;
  }
  K.release();
  return rc;
}

// ---- single GPU: operator = the ParCSR diag block, inner products complete on this device -----------------------------
namespace {
int single_gpu_ops(b200_handle h, b200_parcsr A, b200_amg amg, int precond, b200_krylov_ops *ops) {
  if (!A) B200_FAIL("krylov: null matrix");
  if (A->offd->ncols > 0) B200_FAIL("krylov: this entry point is single-rank; use b200_dist_gmres_solve / b200_dist_bicgstab_solve");
  if (precond < 0 || precond > 2) B200_FAIL("krylov: precond must be 0 (none), 1 (BoomerAMG) or 2 (diagonal scaling)");
  if (precond == 1 && !amg) B200_FAIL("krylov: precond 1 needs a BoomerAMG handle");
  const int n = A->diag->nrows;
  ops->n = ops->cap = n;
  ops->device_mgs = true;
  ops->matvec = [h, A](double alpha, const double *x, double beta, const double *b, double *y) {
    return b200_parcsr_matvec(h, alpha, A, x, beta, b, y);
  };
  ops->precond = [h, A, amg, precond, n](const double *rhs, double *out) {
    if (precond == 1) return b200_amg_precond(h, amg, rhs, out);
    if (precond == 2) return b200_parcsr_diag_scale(h, A, rhs, out);
    return b200_vec_copy(h, n, rhs, out);          // hypre_ParKrylovIdentity
  };
  ops->reduce = [h](const double *d_partials, int k, double *out) {
    B200_CUDA(cudaMemcpyAsync(h->h_pinned, d_partials, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < k; i++) out[i] = h->h_pinned[i];
    return 0;
  };
  return 0;
}
}  // namespace

extern "C" int b200_gmres_solve(b200_handle h, b200_parcsr A, b200_amg amg, const b200_gmres_params *prm, const double *d_b,
                                double *d_x, int *iters_out, double *final_rel_res, double *h_norms, int *converged_out) {
  if (!prm) B200_FAIL("gmres: null argument");
  b200_krylov_ops ops;
  B200_TRY(single_gpu_ops(h, A, amg, prm->precond, &ops));
  return b200_gmres_core(h, &ops, prm, d_b, d_x, iters_out, final_rel_res, h_norms, converged_out);
}

extern "C" int b200_bicgstab_solve(b200_handle h, b200_parcsr A, b200_amg amg, const b200_bicgstab_params *prm,
                                   const double *d_b, double *d_x, int *iters_out, double *final_rel_res, double *h_norms,
                                   int *converged_out) {
  if (!prm) B200_FAIL("bicgstab: null argument");
  b200_krylov_ops ops;
  B200_TRY(single_gpu_ops(h, A, amg, prm->precond, &ops));
  return b200_bicgstab_core(h, &ops, prm, d_b, d_x, iters_out, final_rel_res, h_norms, converged_out);
}
