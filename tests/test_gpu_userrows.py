"""More cases of b200_dist_matrix_create_from_host / _from_ij (tests/test_gpu_dist.py holds the first three)."""
import numpy as np
import pytest

from test_gpu_dist import USER_ROWS_MORE, check_callers_rows, run_ranks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("args,nranks,how", USER_ROWS_MORE)
def test_more_operators_from_the_callers_rows(handle, args, nranks, how):
    """a 2-D rotated-anisotropy operator in four uneven row blocks (CSR rows), and a 27-point operator with one aggressive
    level in two blocks assembled from a SetValues / AddToValues stream"""
    check_callers_rows(handle, args, nranks, how)


def test_operator_from_rows_rejects_bad_input():
    """a row whose diagonal entry is not first, or a column outside the global range, fails on EVERY rank (collective verdict)"""
    import hypre_ve_b200 as hb

    def fn(r, h, c):
        I = np.array([0, 2, 4], np.int32)
        J = np.array([2 * r, 2 * r + 1, 2 * r + 1, 2 * r], np.int32)
        a = np.array([2.0, -1.0, 2.0, -1.0])
        if r == 1:
            J[2], J[3] = J[3], J[2]                          # rank 1: second row stores its off-diagonal first
        try:
            hb.DistMatrix.from_rows(h, c, I, J, a)
            return "built"
        except hb.B200Error as e:
            return str(e)
    out = run_ranks(2, fn)
    assert all("diagonal entry first" in o for o in out), out
