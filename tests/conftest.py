import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# GPU suite order: the rows the headline metric is named after run first (SpMV, the full-size config-2 parity and
# properties, the hierarchy / PCG parity), the N-rank cases (threads as ranks on one GPU) after them.
GPU_ORDER = ["test_gpu_spmv", "test_gpu_fullsize", "test_gpu_amg", "test_gpu_hmis", "test_gpu_api", "test_gpu_refsrc",
             "test_gpu_gs", "test_gpu_ij", "test_gpu_krylov", "test_gpu_agg", "test_gpu_userrows", "test_gpu_dist"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(session, config, items):
    def key(it):
        mod = os.path.splitext(os.path.basename(str(it.fspath)))[0]
        return GPU_ORDER.index(mod) if mod in GPU_ORDER else len(GPU_ORDER)
    items.sort(key=key)          # stable: the order inside a module is kept


@pytest.fixture(scope="module")
def handle():
    """One library handle per test module: b200_finalize at the end of the module returns every device slab to the
    driver, so no module inherits the high-water mark (or a leaked object) of an earlier one."""
    import hypre_ve_b200 as hb
    h = hb.Handle(0)
    yield h
    try:
        h.close()
    except Exception as e:          # a context poisoned by a failing case must not turn into a teardown error
        print("handle.close():", e)
