"""HMIS coarsening on the device (b200_hmis: Ruge-Stueben first pass by one device thread + PMIS seeded with its C points,
hypre_ve_b200/csrc/b200_hmis.cu) against the reference CPU build running its default coarsen_type 10.

The CPU restatement of the same algorithm (oracle/amg_oracle.c -hmis) is pinned bit for bit in tests/test_oracle.py."""
import os
import re
import subprocess

import numpy as np
import pytest

import refio

pytestmark = pytest.mark.gpu
REF_IJ = os.path.join(refio.ROOT, "oracle", "_ref", "ij")


def nlev(d):
    return int(d["hdr"][3])


@pytest.mark.parametrize("args", [
    ["-n", 12, 12, 12, "-rlx", 18], ["-n", 17, 13, 9, "-rlx", 18], ["-n", 10, 10, 10, "-27pt", "-rlx", 18],
    ["-n", 14, 13, 12, "-difconv", "-rlx", 18], ["-n", 24, 20, 1, "-rotate", "-alpha", 45, "-eps", 0.001, "-rlx", 18],
    ["-n", 13, 12, 11, "-perturb", 3, "-rlx", 18], ["-n", 1, 1, 9, "-rlx", 18],
])
def test_hmis_cf_splitting_each_level(handle, args):
    """b200_hmis on the reference's own strength pattern of every level: CF markers equal to the reference's"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(args + ["-keepT", 1])                       # no -pmis: coarsen_type 10
    for l in range(nlev(d) - 1):
        i, j, _, _ = refio.csr(d, "S", l)
        S = hb.Csr.from_host(handle, i, j, None, ncols=i.size - 1)       # square: the first pass walks S^T too
        cf = handle.hmis(S).numpy()
        assert np.array_equal(cf, d["CF%d" % l]), l
        S.destroy()


@pytest.mark.parametrize("args,params", [
    (["-n", 12, 12, 12], dict(RelaxType=13, RelaxTypeUp=14)),                                  # the literal library defaults
    (["-n", 20, 17, 13, "-rlx", 18], dict(RelaxType=18)),
    (["-n", 10, 10, 10, "-27pt", "-rlx", 18, "-mod_rap2", 1], dict(RelaxType=18, ModuleRAP2=1)),
    (["-n", 14, 13, 12, "-difconv", "-rlx", 18], dict(RelaxType=18)),
])
def test_hmis_hierarchy_and_pcg(handle, args, params):
    """CoarsenType 10 through the setup driver: every level bit-identical to the reference's default run, PCG history to 1e-10"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(args + ["-keepT", 1])
    i0, j0, a0, _ = refio.csr(d, "A", 0)
    A = hb.ParCsr.from_host(handle, i0, j0, a0)
    kw = dict(ModuleRAP2=0, CoarsenType=10)
    kw.update(params)
    amg = hb.Amg(handle, **kw)
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    for l in range(nlev(d)):
        i, j, a = amg.level_A(l).download()
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra), ("A", l)
        if l < nlev(d) - 1:
            assert np.array_equal(amg.level_CF(l), d["CF%d" % l]), ("CF", l)
    n = i0.size - 1
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4])
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    amg.destroy(); A.destroy()


def test_hmis_committed_golden(handle):
    import hypre_ve_b200 as hb
    d = refio.read_dump(os.path.join(refio.GOLDEN, "lap7_12_hmis_default.bin"))
    A = hb.ParCsr.laplacian(handle, 12, 12, 12)
    amg = hb.Amg(handle, ModuleRAP2=0, CoarsenType=10, RelaxType=13, RelaxTypeUp=14)
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    for l in range(nlev(d) - 1):
        assert np.array_equal(amg.level_CF(l), d["CF%d" % l]), l
    amg.destroy(); A.destroy()


def test_config_1_as_written_through_the_public_api():
    """BASELINE.json configs[0], `ij -laplacian -n 50 50 50 -solver 1` with every default (HMIS, 13/14): the plain-C client
    over libhypre_b200.so prints the reference driver's result lines (8 iterations, 7.138942e-10)"""
    from hypre_ve_b200 import build as b
    exe = b.build_examples()
    flags = ["-laplacian", "-n", "50", "50", "50", "-solver", "1"]
    p = subprocess.run([exe] + flags, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    its = int(re.search(r"^Iterations = (\d+)", p.stdout, re.M).group(1))
    rel = float(re.search(r"Final Relative Residual Norm = (\S+)", p.stdout).group(1))
    assert its == 8 and abs(rel / 7.138942e-10 - 1) < 1e-6
    q = subprocess.run([REF_IJ] + flags, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert int(re.search(r"^Iterations = (\d+)", q.stdout, re.M).group(1)) == its
