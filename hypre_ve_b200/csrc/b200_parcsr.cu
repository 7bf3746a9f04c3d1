// b200_parcsr.cu -- row-partitioned ParCSR matrix (diag + offd blocks), the on-device problem
// generators, and the distributed SpMV.
//
// Reference: hypre_ParCSRMatrix (parcsr_mv/par_csr_matrix.h:27-95), GenerateLaplacian
// (parcsr_ls/par_laplace.c:15-357), GenerateLaplacian27pt (parcsr_ls/par_laplace_27pt.c:15),
// hypre_ParCSRMatrixMatvecOutOfPlace (parcsr_mv/par_csr_matvec.c:22-359).
//
// The generators run as one thread per grid point (count -> scan -> fill) and emit exactly the
// reference's entry order; ghost columns are compressed by sort/unique + binary search instead of
// the reference's O(num_cols_offd^2) remap loop (par_laplace.c:316-323).
#include <cmath>
#include "b200_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

int b200_csr_spmv_epi(b200_handle h, b200_csr A, const double *x, double *y, int mode, double alpha,
                      double beta, const double *b, const double *d);
int b200_halo_exchange(b200_handle h, b200_parcsr A, const double *d_x);   // b200_comm.cu
void b200_halo_destroy(b200_handle h, b200_halo_s *halo);

namespace {

struct Box {
  int nx, ny, nz;                 // global grid
  int x0, x1, y0, y1, z0, z1;     // this rank's box [x0,x1) ...
  int xl0, xl1, yl0, yl1, zl0, zl1;   // neighbour boxes: lower neighbour [xl0,x0), upper [x1,xl1) etc.
};

__host__ __device__ inline void part_range(int length, int nprocs, int id, int *lo, int *hi) {
  // hypre_GeneratePartitioning (seq_mv/genpart.c:18-38)
  int size = length / nprocs, rest = length - size * nprocs;
  *lo = id * size + (id < rest ? id : rest);
  *hi = *lo + size + (id < rest ? 1 : 0);
}

struct Grid {
  int nx, ny, nz, P, Q, R, p, q, r;
  int x0, x1, y0, y1, z0, z1;
};

// global index of grid point (ix,iy,iz): hypre_map (par_laplace.c:363-387), owner box found from
// the coordinates (the reference passes the owner explicitly; it is the box containing the point).
__device__ inline int owner_1d(int i, int length, int nprocs) {
  int size = length / nprocs, rest = length - size * nprocs;
  int split = rest * (size + 1);
  return i < split ? i / (size + 1) : rest + (i - split) / (size > 0 ? size : 1);
}
__device__ inline int global_index(const Grid &g, int ix, int iy, int iz) {
  int pp = owner_1d(ix, g.nx, g.P), qq = owner_1d(iy, g.ny, g.Q), rr = owner_1d(iz, g.nz, g.R);
  int xa, xb, ya, yb, za, zb;
  part_range(g.nx, g.P, pp, &xa, &xb);
  part_range(g.ny, g.Q, qq, &ya, &yb);
  part_range(g.nz, g.R, rr, &za, &zb);
  int nxl = xb - xa, nyl = yb - ya, nzl = zb - za;
  long long gi = (long long)za * g.nx * g.ny + (long long)ya * g.nx * nzl + (long long)xa * (nyl * nzl);
  gi += (long long)((iz - za) * nyl + (iy - ya)) * nxl + (ix - xa);
  return (int)gi;
}

// STENCIL 7: centre, z-, y-, x-, x+, y+, z+ (par_laplace.c:206-300); 27: centre then (dz,dy,dx)
// lexicographic (par_laplace_27pt.c fill pass).
struct StencilVals { double v[7]; };     // centre, x-, y-, z-, x+, y+, z+ (27-pt: centre, off-centre)

// STENCIL 72 = the 2-D rotated-anisotropy 7-point stencil of GenerateRotate7pt (par_rotate_7pt.c:228-350, nz = 1):
// centre, (-1,-1), (0,-1), (-1,0), (+1,0), (0,+1), (+1,+1) with values v0, v3, v2, v1, v1, v2, v3.
__host__ __device__ constexpr int stencil_points(int stencil) { return stencil == 72 ? 7 : stencil; }
__host__ __device__ constexpr int stencil_values(int stencil) { return stencil == 7 ? 7 : stencil == 72 ? 4 : 2; }
template <int STENCIL>
__device__ inline void stencil_offset(int k, int *dx, int *dy, int *dz, int *vidx) {
  if (STENCIL == 72) {
    const int ox[7] = {0, -1, 0, -1, 1, 0, 1};
    const int oy[7] = {0, -1, -1, 0, 0, 1, 1};
    const int vi[7] = {0, 3, 2, 1, 1, 2, 3};
    *dx = ox[k]; *dy = oy[k]; *dz = 0; *vidx = vi[k];
  } else if (STENCIL == 7) {
    const int ox[7] = {0, 0, 0, -1, 1, 0, 0};
    const int oy[7] = {0, 0, -1, 0, 0, 1, 0};
    const int oz[7] = {0, -1, 0, 0, 0, 0, 1};
    const int vi[7] = {0, 3, 2, 1, 4, 5, 6};      // GenerateDifConv's seven values (par_difconv.c:247-330); the Laplacian passes v[4..6] = v[1..3]
    *dx = ox[k]; *dy = oy[k]; *dz = oz[k]; *vidx = vi[k];
  } else {
    if (k == 0) { *dx = *dy = *dz = 0; *vidx = 0; return; }
    int m = k - 1;          // 0..25 over the 26 neighbours in lexicographic order, skipping centre (13)
    if (m >= 13) m += 1;
    *dz = m / 9 - 1; *dy = (m / 3) % 3 - 1; *dx = m % 3 - 1; *vidx = 1;
  }
}

template <int STENCIL, bool FILL>
__global__ void gen_kernel(Grid g, StencilVals sv, int *diag_i, int *offd_i,
                           int *diag_j, double *diag_a, int *offd_gj, double *offd_a) {
  const int nxl = g.x1 - g.x0, nyl = g.y1 - g.y0, nzl = g.z1 - g.z0;
  const long long nloc = (long long)nxl * nyl * nzl;
  long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nloc) return;
  const int lx = (int)(row % nxl), ly = (int)((row / nxl) % nyl), lz = (int)(row / ((long long)nxl * nyl));
  const int ix = g.x0 + lx, iy = g.y0 + ly, iz = g.z0 + lz;
  int cd = 0, co = 0;
  int pd = 0, po = 0;
  if (FILL) { pd = diag_i[row]; po = offd_i[row]; }
  const double *vals = sv.v;
  for (int k = 0; k < stencil_points(STENCIL); k++) {
    int dx, dy, dz, vi;
    stencil_offset<STENCIL>(k, &dx, &dy, &dz, &vi);
    const int jx = ix + dx, jy = iy + dy, jz = iz + dz;
    if (jx < 0 || jx >= g.nx || jy < 0 || jy >= g.ny || jz < 0 || jz >= g.nz) continue;
    const bool local = jx >= g.x0 && jx < g.x1 && jy >= g.y0 && jy < g.y1 && jz >= g.z0 && jz < g.z1;
    if (local) {
      if (FILL) {
        diag_j[pd + cd] = (int)(((long long)(jz - g.z0) * nyl + (jy - g.y0)) * nxl + (jx - g.x0));
        diag_a[pd + cd] = vals[vi];
      }
      cd++;
    } else {
      if (FILL) {
        offd_gj[po + co] = global_index(g, jx, jy, jz);
        offd_a[po + co] = vals[vi];
      }
      co++;
    }
  }
  if (!FILL) { diag_i[row] = cd; offd_i[row] = co; }
}

// merged form for the multi-rank path: local rows, GLOBAL column ids, reference entry order
template <int STENCIL, bool FILL>
__global__ void gen_global_kernel(Grid g, StencilVals sv, int *A_i, int *A_j, double *A_a) {
  const int nxl = g.x1 - g.x0, nyl = g.y1 - g.y0, nzl = g.z1 - g.z0;
  const long long nloc = (long long)nxl * nyl * nzl;
  long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nloc) return;
  const int lx = (int)(row % nxl), ly = (int)((row / nxl) % nyl), lz = (int)(row / ((long long)nxl * nyl));
  const int ix = g.x0 + lx, iy = g.y0 + ly, iz = g.z0 + lz;
  int cnt = 0;
  const int pos = FILL ? A_i[row] : 0;
  const double *vals = sv.v;
  for (int k = 0; k < stencil_points(STENCIL); k++) {
    int dx, dy, dz, vi;
    stencil_offset<STENCIL>(k, &dx, &dy, &dz, &vi);
    const int jx = ix + dx, jy = iy + dy, jz = iz + dz;
    if (jx < 0 || jx >= g.nx || jy < 0 || jy >= g.ny || jz < 0 || jz >= g.nz) continue;
    if (FILL) { A_j[pos + cnt] = global_index(g, jx, jy, jz); A_a[pos + cnt] = vals[vi]; }
    cnt++;
  }
  if (!FILL) A_i[row] = cnt;
}

__global__ void lookup_kernel(int n, const int *__restrict__ gj, int ncols, const int *__restrict__ col_map,
                              int *__restrict__ j) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int target = gj[k], lo = 0, hi = ncols - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (col_map[mid] < target) lo = mid + 1; else hi = mid;
  }
  j[k] = lo;
}


}  // namespace

// Compress global ghost column ids to [0,ncols_offd) with a sorted col_map (reference:
// par_laplace.c:303-330, hypre_ParCSRMatrix col_map_offd is sorted ascending).
int b200_compress_offd(b200_handle h, int nnz, const int *d_gj, int *d_j, int *ncols_out, int **col_map_out) {
  *ncols_out = 0; *col_map_out = nullptr;
  if (nnz == 0) return 0;
  int *keys = nullptr, *sorted = nullptr, *uniq = nullptr, *d_num = nullptr;
  B200_TRY(b200_dalloc<int>(h, &keys, nnz));
  B200_TRY(b200_dalloc<int>(h, &sorted, nnz));
  B200_TRY(b200_dalloc<int>(h, &uniq, nnz));
  B200_TRY(b200_dalloc<int>(h, &d_num, 1));
  B200_CUDA(cudaMemcpyAsync(keys, d_gj, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, h->stream));
  size_t tb = 0, tb2 = 0;
  B200_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, sorted, nnz, 0, 32, h->stream));
  B200_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, sorted, uniq, d_num, nnz, h->stream));
  char *tmp = nullptr;
  B200_TRY(b200_dalloc<char>(h, &tmp, tb > tb2 ? tb : tb2));
  B200_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, sorted, nnz, 0, 32, h->stream));
  B200_CUDA(cub::DeviceSelect::Unique(tmp, tb2, sorted, uniq, d_num, nnz, h->stream));
  g_b200_launches += 2;
  int num = 0;
  B200_CUDA(cudaMemcpyAsync(&num, d_num, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  int *col_map = nullptr;
  B200_TRY(b200_dalloc<int>(h, &col_map, num));
  B200_CUDA(cudaMemcpyAsync(col_map, uniq, sizeof(int) * (size_t)num, cudaMemcpyDeviceToDevice, h->stream));
  lookup_kernel<<<b200_grid(nnz, 256), 256, 0, h->stream>>>(nnz, d_gj, num, col_map, d_j);
  B200_LAUNCH_CHECK();
  B200_TRY(b200_dfree(h, keys)); B200_TRY(b200_dfree(h, sorted)); B200_TRY(b200_dfree(h, uniq));
  B200_TRY(b200_dfree(h, d_num)); B200_TRY(b200_dfree(h, tmp));
  *ncols_out = num; *col_map_out = col_map;
  return 0;
}

int b200_box_first_row(int nx, int ny, int nz, int P, int Q, int R, int p, int q, int r) {
  int x0, x1, y0, y1, z0, z1;
  part_range(nx, P, p, &x0, &x1); part_range(ny, Q, q, &y0, &y1); part_range(nz, R, r, &z0, &z1);
  (void)x1; (void)R;
  return (int)((long long)z0 * nx * ny + ((long long)y0 * nx + (long long)x0 * (y1 - y0)) * (z1 - z0));   // par_laplace.c:78
}

template <int STENCIL>
static int generate_global(b200_handle h, Grid g, const double *vals, b200_csr *out) {
  const int nxl = g.x1 - g.x0, nyl = g.y1 - g.y0, nzl = g.z1 - g.z0;
  const int nloc = nxl * nyl * nzl;
  int *ai = nullptr;
  B200_TRY(b200_dalloc<int>(h, &ai, (size_t)nloc + 1));
  B200_CUDA(cudaMemsetAsync(ai + nloc, 0, sizeof(int), h->stream));
  StencilVals sv;
  for (int k = 0; k < 7; k++) sv.v[k] = k < stencil_values(STENCIL) ? vals[k] : 0.0;
  if (nloc) {
    gen_global_kernel<STENCIL, false><<<b200_grid(nloc, 256), 256, 0, h->stream>>>(g, sv, ai, nullptr, nullptr);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_exclusive_scan_inplace(h, ai, (size_t)nloc + 1));
  int nnz = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz, ai + nloc, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  b200_csr A = nullptr;
  B200_TRY(b200_csr_alloc(h, nloc, g.nx * g.ny * g.nz, nnz, true, &A));
  B200_CUDA(cudaMemcpyAsync(A->i, ai, sizeof(int) * ((size_t)nloc + 1), cudaMemcpyDeviceToDevice, h->stream));
  if (nloc) {
    gen_global_kernel<STENCIL, true><<<b200_grid(nloc, 256), 256, 0, h->stream>>>(g, sv, A->i, A->j, A->a);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_dfree(h, ai));
  *out = A;
  return 0;
}

int b200_generate_stencil_global(b200_handle h, int nx, int ny, int nz, int P, int Q, int R, int p, int q, int r,
                                 int stencil, const double *vals, b200_csr *out, int *first_row) {
  if (nx < 1 || ny < 1 || nz < 1 || P < 1 || Q < 1 || R < 1) B200_FAIL("bad grid");
  if ((long long)nx * ny * nz > 2147483647LL) B200_FAIL("global size exceeds int32 (HYPRE_BigInt=int)");
  Grid g{nx, ny, nz, P, Q, R, p, q, r, 0, 0, 0, 0, 0, 0};
  part_range(nx, P, p, &g.x0, &g.x1);
  part_range(ny, Q, q, &g.y0, &g.y1);
  part_range(nz, R, r, &g.z0, &g.z1);
  *first_row = b200_box_first_row(nx, ny, nz, P, Q, R, p, q, r);
  if (stencil == 72) {
    if (nz != 1 || R != 1) B200_FAIL("the rotated 7-point operator is two-dimensional (nz = 1, R = 1)");
    return generate_global<72>(h, g, vals, out);
  }
  return stencil == 7 ? generate_global<7>(h, g, vals, out) : generate_global<27>(h, g, vals, out);
}

template <int STENCIL>
static int generate(b200_handle h, int nx, int ny, int nz, int P, int Q, int R, int p, int q, int r,
                    const double *vals, b200_parcsr *out) {
  if (nx < 1 || ny < 1 || nz < 1 || P < 1 || Q < 1 || R < 1) B200_FAIL("bad grid");
  if (p < 0 || p >= P || q < 0 || q >= Q || r < 0 || r >= R) B200_FAIL("bad process coordinates");
  if ((long long)nx * ny * nz > 2147483647LL) B200_FAIL("global size exceeds int32 (HYPRE_BigInt=int)");
  Grid g{nx, ny, nz, P, Q, R, p, q, r, 0, 0, 0, 0, 0, 0};
  part_range(nx, P, p, &g.x0, &g.x1);
  part_range(ny, Q, q, &g.y0, &g.y1);
  part_range(nz, R, r, &g.z0, &g.z1);
  const int nxl = g.x1 - g.x0, nyl = g.y1 - g.y0, nzl = g.z1 - g.z0;
  const long long nloc_ll = (long long)nxl * nyl * nzl;
  const int nloc = (int)nloc_ll;
  int *di = nullptr, *oi = nullptr;
  B200_TRY(b200_dalloc<int>(h, &di, (size_t)nloc + 1));
  B200_TRY(b200_dalloc<int>(h, &oi, (size_t)nloc + 1));
  B200_CUDA(cudaMemsetAsync(di + nloc, 0, sizeof(int), h->stream));
  B200_CUDA(cudaMemsetAsync(oi + nloc, 0, sizeof(int), h->stream));
  StencilVals sv;
  for (int k = 0; k < 7; k++) sv.v[k] = k < stencil_values(STENCIL) ? vals[k] : 0.0;
  if (nloc) {
    gen_kernel<STENCIL, false><<<b200_grid(nloc, 256), 256, 0, h->stream>>>(g, sv, di, oi, nullptr,
                                                                           nullptr, nullptr, nullptr);
    B200_LAUNCH_CHECK();
  }
  B200_TRY(b200_exclusive_scan_inplace(h, di, (size_t)nloc + 1));
  B200_TRY(b200_exclusive_scan_inplace(h, oi, (size_t)nloc + 1));
  int nnz_d = 0, nnz_o = 0;
  B200_CUDA(cudaMemcpyAsync(&nnz_d, di + nloc, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(&nnz_o, oi + nloc, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));

  b200_parcsr A = new b200_parcsr_s();
  A->global_rows = A->global_cols = nx * ny * nz;
  A->first_row = A->first_col = (int)((long long)g.z0 * nx * ny + ((long long)g.y0 * nx + (long long)g.x0 * nyl) * nzl);
  B200_TRY(b200_csr_alloc(h, nloc, nloc, nnz_d, true, &A->diag));
  B200_TRY(b200_csr_alloc(h, nloc, 0, nnz_o, true, &A->offd));
  B200_CUDA(cudaMemcpyAsync(A->diag->i, di, sizeof(int) * ((size_t)nloc + 1), cudaMemcpyDeviceToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(A->offd->i, oi, sizeof(int) * ((size_t)nloc + 1), cudaMemcpyDeviceToDevice, h->stream));
  int *ogj = nullptr;
  B200_TRY(b200_dalloc<int>(h, &ogj, (size_t)nnz_o));
  if (nloc) {
    gen_kernel<STENCIL, true><<<b200_grid(nloc, 256), 256, 0, h->stream>>>(g, sv, A->diag->i, A->offd->i,
                                                                          A->diag->j, A->diag->a, ogj, A->offd->a);
    B200_LAUNCH_CHECK();
  }
  int ncols_offd = 0;
  B200_TRY(b200_compress_offd(h, nnz_o, ogj, A->offd->j, &ncols_offd, &A->col_map_offd));
  A->offd->ncols = ncols_offd;
  A->h_col_map_offd.resize(ncols_offd);
  if (ncols_offd) {
    B200_CUDA(cudaMemcpyAsync(A->h_col_map_offd.data(), A->col_map_offd, sizeof(int) * (size_t)ncols_offd,
                              cudaMemcpyDeviceToHost, h->stream));
    B200_CUDA(cudaStreamSynchronize(h->stream));
    B200_TRY(b200_dalloc<double>(h, &A->x_ghost, ncols_offd));
  }
  B200_TRY(b200_dfree(h, ogj)); B200_TRY(b200_dfree(h, di)); B200_TRY(b200_dfree(h, oi));
  B200_TRY(b200_csr_build_plan(h, A->diag));
  if (nnz_o) B200_TRY(b200_csr_build_plan(h, A->offd));
  *out = A;
  return 0;
}

extern "C" int b200_generate_laplacian(b200_handle h, int nx, int ny, int nz, int P, int Q, int R, int p, int q,
                                       int r, const double values[4], b200_parcsr *A) {
  const double v7[7] = {values[0], values[1], values[2], values[3], values[1], values[2], values[3]};
  return generate<7>(h, nx, ny, nz, P, Q, R, p, q, r, v7, A);
}
// GenerateDifConv (parcsr_ls/par_difconv.c:15-365): the 7-point pattern and entry order of GenerateLaplacian with
// separate lower (x-,y-,z-) and upper (x+,y+,z+) coefficients -- a nonsymmetric operator.
extern "C" int b200_generate_difconv(b200_handle h, int nx, int ny, int nz, int P, int Q, int R, int p, int q,
                                     int r, const double values[7], b200_parcsr *A) {
  return generate<7>(h, nx, ny, nz, P, Q, R, p, q, r, values, A);
}
// value[4] of GenerateRotate7pt from the rotation angle (degrees) and the anisotropy (par_rotate_7pt.c:62-73)
void b200_rotate7pt_values(double alpha, double eps, double *value) {
  const double pi = 4.0 * atan(1.0);
  const double x = pi * alpha / 180.0;
  const double s = sin(x), c = cos(x);
  const double ac = -(c * c + eps * s * s), bc = 2.0 * (1.0 - eps) * s * c, cc = -(s * s + eps * c * c);
  value[0] = -2 * (2 * ac + bc + 2 * cc);
  value[1] = 2 * ac + bc;
  value[2] = bc + 2 * cc;
  value[3] = -bc;
}
// GenerateRotate7pt (parcsr_ls/par_rotate_7pt.c:15-400): -(c^2 + eps s^2) u_xx - 2(1 - eps) s c u_xy - (s^2 + eps c^2) u_yy
// on an nx x ny grid, P x Q process grid, this rank at (p, q)
extern "C" int b200_generate_rotate7pt(b200_handle h, int nx, int ny, int P, int Q, int p, int q, double alpha, double eps,
                                       b200_parcsr *A) {
  double v[4];
  b200_rotate7pt_values(alpha, eps, v);
  return generate<72>(h, nx, ny, 1, P, Q, 1, p, q, 0, v, A);
}
extern "C" int b200_generate_laplacian27(b200_handle h, int nx, int ny, int nz, int P, int Q, int R, int p, int q,
                                         int r, const double values[2], b200_parcsr *A) {
  return generate<27>(h, nx, ny, nz, P, Q, R, p, q, r, values, A);
}

extern "C" int b200_parcsr_create_from_host(b200_handle h, int nrows, int ncols, int nnz, const int *h_i,
                                            const int *h_j, const double *h_a, b200_parcsr *out) {
  b200_parcsr A = new b200_parcsr_s();
  A->global_rows = nrows; A->global_cols = ncols;
  B200_TRY(b200_csr_create_from_host(h, nrows, ncols, nnz, h_i, h_j, h_a, &A->diag));
  B200_TRY(b200_csr_alloc(h, nrows, 0, 0, true, &A->offd));
  B200_CUDA(cudaMemsetAsync(A->offd->i, 0, sizeof(int) * ((size_t)nrows + 1), h->stream));
  *out = A;
  return 0;
}

extern "C" int b200_parcsr_destroy(b200_handle h, b200_parcsr A) {
  if (!A) return 0;
  B200_TRY(b200_csr_destroy(h, A->diag));
  B200_TRY(b200_csr_destroy(h, A->offd));
  B200_TRY(b200_dfree(h, A->col_map_offd));
  B200_TRY(b200_dfree(h, A->x_ghost));
  if (A->halo) b200_halo_destroy(h, A->halo);
  delete A;
  return 0;
}

extern "C" int b200_parcsr_local_rows(b200_parcsr A, int *nrows, int *nnz_diag, int *nnz_offd, int *ncols_offd) {
  if (!A) B200_FAIL("null matrix");
  if (nrows) *nrows = A->diag->nrows;
  if (nnz_diag) *nnz_diag = A->diag->nnz;
  if (nnz_offd) *nnz_offd = A->offd->nnz;
  if (ncols_offd) *ncols_offd = A->offd->ncols;
  return 0;
}
extern "C" b200_csr b200_parcsr_diag(b200_parcsr A) { return A ? A->diag : nullptr; }
extern "C" b200_csr b200_parcsr_offd(b200_parcsr A) { return A ? A->offd : nullptr; }

// y = alpha*A*x + beta*b: halo exchange of x (job 1), diag SpMV, then y += alpha*offd*x_ghost
// (par_csr_matvec.c:250-315).
extern "C" int b200_parcsr_matvec(b200_handle h, double alpha, b200_parcsr A, const double *d_x, double beta,
                                  const double *d_b, double *d_y) {
  if (!A) B200_FAIL("null matrix");
  if (A->offd->ncols > 0) {
    if (!A->halo) B200_FAIL("matrix has ghost columns but no halo plan (call b200_parcsr_build_halo)");
    B200_TRY(b200_halo_exchange(h, A, d_x));
  }
  B200_TRY(b200_csr_matvec(h, alpha, A->diag, d_x, beta, d_b, d_y));
  if (A->offd->ncols > 0 && A->offd->nnz > 0 && alpha != 0.0)
    B200_TRY(b200_csr_spmv_epi(h, A->offd, A->x_ghost, d_y, 0, alpha, 1.0, d_y, nullptr));
  return 0;
}
