"""GPU parity tests (through the C-ABI) for the generators, the streaming CSR SpMV and BLAS-1.

Checker: the reference's own CPU build (oracle/_ref/ref_dump -> GenerateLaplacian / 27pt output)
for the matrices, and a scipy CSR product for the SpMV values.  Tolerance for the FP64 SpMV:
1e-13 relative to ||A||_inf*||x||_inf per row (sums of <= a few hundred terms), stated below.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import refio

pytestmark = pytest.mark.gpu


def _spmv_check(handle, i, j, a, ncols, alpha, beta, seed=0):
    import hypre_ve_b200 as hb
    rng = np.random.default_rng(seed)
    n = i.size - 1
    A = hb.Csr.from_host(handle, i, j, a)
    x = rng.standard_normal(ncols)
    b = rng.standard_normal(n)
    dx, db, dy = handle.array(x), handle.array(b), handle.zeros(n)
    A.matvec(alpha, dx, beta, db, dy)
    y = dy.numpy()
    M = sp.csr_matrix((a, j, i), shape=(n, ncols))
    ref = alpha * (M @ x) + beta * b
    scale = (abs(M) @ np.abs(x)) * abs(alpha) + abs(beta) * np.abs(b) + 1e-300
    err = np.max(np.abs(y - ref) / scale) if n else 0.0
    assert err < 1e-13, err
    # in-place form b == y (hypre_CSRMatrixMatvec: y = alpha*A*x + beta*y)
    dy2 = handle.array(b)
    A.matvec(alpha, dx, beta, dy2, dy2)
    assert np.max(np.abs(dy2.numpy() - ref) / scale) < 1e-13 if n else True
    A.destroy()


@pytest.mark.parametrize("dims", [(7, 5, 3), (40, 33, 27), (1, 9, 1), (64, 64, 64)])
def test_laplacian7_matches_reference_generator(handle, dims):
    import hypre_ve_b200 as hb
    A = hb.ParCsr.laplacian(handle, *dims)
    i, j, a = A.diag.download()
    d, _ = refio.run_ref(["-n", *dims, "-pmis", "-rlx", 18, "-max_iter", 1])
    ri, rj, ra, _ = refio.csr(d, "A", 0)
    assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra)
    A.destroy()


@pytest.mark.parametrize("dims", [(6, 5, 4), (20, 17, 15), (1, 8, 8)])
def test_laplacian27_matches_reference_generator(handle, dims):
    import hypre_ve_b200 as hb
    A = hb.ParCsr.laplacian27(handle, *dims)
    i, j, a = A.diag.download()
    d, _ = refio.run_ref(["-n", *dims, "-27pt", "-pmis", "-rlx", 18, "-max_iter", 1])
    ri, rj, ra, _ = refio.csr(d, "A", 0)
    assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra)
    A.destroy()


@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-1.0, 1.0), (2.5, -0.75), (0.0, 2.0), (1.0, 1.0)])
def test_spmv_stencil(handle, alpha, beta):
    import hypre_ve_b200 as hb
    A = hb.ParCsr.laplacian(handle, 37, 29, 23)
    i, j, a = A.diag.download()
    _spmv_check(handle, i, j, a, i.size - 1, alpha, beta)
    A.destroy()


@pytest.mark.parametrize("kind", ["empty_rows", "ragged", "long_rows", "dense_coarse", "one_row", "no_rows", "no_nnz"])
def test_spmv_ragged(handle, kind):
    rng = np.random.default_rng(42)
    if kind == "no_rows":
        _spmv_check(handle, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0), 5, 1.0, 0.0)
        return
    if kind == "no_nnz":
        _spmv_check(handle, np.zeros(101, np.int32), np.zeros(0, np.int32), np.zeros(0), 50, 2.0, 3.0)
        return
    n, m = {"empty_rows": (5000, 4000), "ragged": (20000, 20000), "long_rows": (300, 30000),
            "dense_coarse": (3000, 3000), "one_row": (1, 10000)}[kind]
    if kind == "empty_rows":
        lens = rng.integers(0, 4, n) * (rng.random(n) < 0.3)
    elif kind == "ragged":
        lens = rng.integers(0, 60, n)
    elif kind == "long_rows":
        lens = rng.integers(0, 12000, n) * (rng.random(n) < 0.2)   # longer than the 4096-entry shared tile
    elif kind == "dense_coarse":
        lens = rng.integers(40, 170, n)
    else:
        lens = np.array([7777])
    lens = np.minimum(lens, m)
    i = np.zeros(n + 1, np.int32)
    i[1:] = np.cumsum(lens)
    j = np.concatenate([rng.choice(m, size=int(k), replace=False) for k in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    a = rng.standard_normal(j.size)
    for alpha, beta in [(1.0, 0.0), (-1.0, 1.0), (0.3, 0.6)]:
        _spmv_check(handle, i, j, a, m, alpha, beta)


def test_blas1(handle):
    rng = np.random.default_rng(1)
    for n in [1, 31, 1000, 1 << 20, (1 << 22) + 17]:
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        dx, dy = handle.array(x), handle.array(y)
        d = handle.dot(dx, dy)
        ref = float(np.dot(x, y))
        assert abs(d - ref) <= 1e-13 * float(np.dot(np.abs(x), np.abs(y))) + 1e-300
        assert handle.dot(dx, dy) == d          # deterministic reduction
        handle.axpy(-0.37, dx, dy)
        np.testing.assert_allclose(dy.numpy(), y - 0.37 * x, rtol=1e-15, atol=1e-15)
        handle.scale(1.5, dy)
        np.testing.assert_allclose(dy.numpy(), 1.5 * (y - 0.37 * x), rtol=1e-15, atol=1e-15)
        handle.fill(dx, 2.0)
        assert np.all(dx.numpy() == 2.0)
        handle.copy(dy, dx)
        assert np.array_equal(dx.numpy(), dy.numpy())
        dx.free(); dy.free()


def test_spmv_full_size_linearity(handle):
    """BASELINE config-2 size (256^3 7-pt): size-independent properties instead of an oracle run:
    A*1 = row sums (0 in the interior, boundary deficit elsewhere) and linearity A(x+2z)=Ax+2Az."""
    import hypre_ve_b200 as hb
    n1 = 256
    A = hb.ParCsr.laplacian(handle, n1, n1, n1)
    n, nnz, _, _ = A.local
    assert n == n1 ** 3 and nnz == 7 * n1 ** 3 - 6 * n1 ** 2
    ones = handle.zeros(n); handle.fill(ones, 1.0)
    y = handle.zeros(n)
    A.matvec(1.0, ones, 0.0, None, y)
    ys = y.numpy().reshape(n1, n1, n1)
    assert np.all(ys[1:-1, 1:-1, 1:-1] == 0.0)
    assert ys.sum() == 6.0 * n1 * n1      # each boundary face point misses one -1
    rng = np.random.default_rng(3)
    x, z = rng.standard_normal(n), rng.standard_normal(n)
    dx, dz = handle.array(x), handle.array(z)
    dxz = handle.array(x + 2.0 * z)
    y1, y2, y3 = handle.zeros(n), handle.zeros(n), handle.zeros(n)
    A.matvec(1.0, dx, 0.0, None, y1)
    A.matvec(1.0, dz, 0.0, None, y2)
    A.matvec(1.0, dxz, 0.0, None, y3)
    np.testing.assert_allclose(y3.numpy(), y1.numpy() + 2.0 * y2.numpy(), rtol=0, atol=1e-12)
    A.destroy()


def test_matvecT_matches_sequential_scatter(handle):
    """y = alpha A^T x + beta b (hypre_CSRMatrixMatvecT, csr_matvec.c:424-668) against the reference's
    row-by-row scatter loop (1e-14: the SpMV kernel may split a row's sum over several lanes)"""
    import scipy.sparse as sp
    import hypre_ve_b200 as hb
    rng = np.random.default_rng(4)
    M = sp.random(900, 400, density=0.02, random_state=9, format="csr")
    M.sort_indices()
    x, b = rng.standard_normal(900), rng.standard_normal(400)
    A = hb.Csr.from_host(handle, M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data)
    dx, db, dy = handle.array(x), handle.array(b), handle.zeros(400)
    A.matvecT(1.0, dx, 0.0, None, dy)
    want = np.zeros(400)
    for i in range(900):                               # csr_matvec.c:560-575 (one thread)
        for jj in range(M.indptr[i], M.indptr[i + 1]):
            want[M.indices[jj]] += M.data[jj] * x[i]
    np.testing.assert_allclose(dy.numpy(), want, rtol=1e-14, atol=1e-14)
    A.matvecT(-2.0, dx, 0.5, db, dy)
    np.testing.assert_allclose(dy.numpy(), -2.0 * (M.T @ x) + 0.5 * b, rtol=1e-13, atol=1e-13)
    A.destroy()
