"""Helper of tests/test_gpu_spmv_dict.py: runs the SpMV on stencil-structured operators and prints one line per case,
`name sha256(y) max_rel_err`, so that the test can compare the dictionary-compressed path (B200_SPMV_DICT, read once per
process) with the plain one bit for bit.  The scipy product is the independent checker (1e-13 relative per row)."""
import hashlib
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb


def run(name, handle, i, j, a, alpha, beta, seed):
    rng = np.random.default_rng(seed)
    n = i.size - 1
    A = hb.Csr.from_host(handle, i, j, a)
    x = rng.standard_normal(n)
    b = rng.standard_normal(n)
    dx, db, dy = handle.array(x), handle.array(b), handle.zeros(n)
    A.matvec(alpha, dx, beta, db, dy)
    y = dy.numpy()
    M = sp.csr_matrix((a, j, i), shape=(n, n))
    ref = alpha * (M @ x) + beta * b
    scale = (abs(M) @ np.abs(x)) * abs(alpha) + abs(beta) * np.abs(b) + 1e-300
    err = float(np.max(np.abs(y - ref) / scale))
    print(name, hashlib.sha256(y.tobytes()).hexdigest(), "%.3e" % err, flush=True)
    assert err < 1e-13, (name, err)
    A.destroy()


def main():
    h = hb.Handle(0)
    rng = np.random.default_rng(7)
    A = hb.ParCsr.laplacian(h, 37, 29, 23)
    i, j, a = A.diag.download()
    run("lap7", h, i, j, a, 1.0, 0.0, 1)
    run("lap7_axpby", h, i, j, a, -1.0, 1.0, 2)
    run("lap7_random_values", h, i, j, rng.standard_normal(a.size), 2.5, -0.75, 3)       # offsets compress, values do not
    A.destroy()
    A = hb.ParCsr.laplacian27(h, 20, 17, 15)
    i, j, a = A.diag.download()
    run("lap27", h, i, j, a, 1.0, 1.0, 4)
    A.destroy()
    A = hb.ParCsr.difconv(h, 24, 21, 19, a=(3.0, -2.0, 1.0), atype=3)
    i, j, a = A.diag.download()
    run("difconv", h, i, j, a, 1.0, 0.0, 5)
    A.destroy()
    # a banded operator with 300 distinct offsets but two values: only the values compress
    n = 6000
    offs = np.unique(np.concatenate([[0], rng.integers(-1500, 1500, 600)]))[:300]
    rows = []
    for r in range(n):
        pick = offs[(r * 7 + np.arange(5)) % offs.size]
        cols = np.unique(np.clip(r + pick, 0, n - 1))
        cols = np.concatenate([[r], cols[cols != r]])
        rows.append(cols)
    i = np.zeros(n + 1, np.int32)
    i[1:] = np.cumsum([len(c) for c in rows])
    j = np.concatenate(rows).astype(np.int32)
    a = np.where(j == np.repeat(np.arange(n), np.diff(i)), 4.0, -0.5)
    run("many_offsets_two_values", h, i, j, a, 1.0, 0.0, 6)
    h.close()


if __name__ == "__main__":
    main()
