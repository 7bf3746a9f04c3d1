"""GPU parity tests for aggressive coarsening: hypre_BoomerAMGCreate2ndS, the second
PMIS (CF_init 3) + CorrectCFMarker, hypre_BoomerAMGBuildMultipass, and whole BoomerAMG-PCG solves with
`-agg_nl L`, against the reference's own CPU build (hierarchy bit-exact, iteration counts exact, residual
history 1e-10)."""
import numpy as np
import pytest

import refio

pytestmark = pytest.mark.gpu

CASES = [
    (["-n", 14, 11, 9], 1), (["-n", 20, 20, 20], 1), (["-n", 20, 20, 20], 2), (["-n", 12, 12, 12, "-27pt"], 1),
    (["-n", 16, 16, 16, "-c", 1, 1, 0.001], 1), (["-n", 16, 16, 16, "-c", 1, 1, 0.001], 2), (["-n", 36, 31, 28], 1),
    (["-n", 50, 50, 50], 1),
]


@pytest.mark.parametrize("args,agg_nl", CASES)
def test_amg_pcg_with_aggressive_coarsening(handle, args, agg_nl):
    import hypre_ve_b200 as hb
    d, out = refio.run_ref(args + ["-pmis", "-rlx", 18, "-mod_rap2", 1, "-agg_nl", agg_nl])
    nx, ny, nz = args[1:4]
    A = hb.ParCsr.laplacian27(handle, nx, ny, nz) if "-27pt" in args else \
        hb.ParCsr.laplacian(handle, nx, ny, nz, c=tuple(args[5:8]) if "-c" in args else (1.0, 1.0, 1.0))
    amg = hb.Amg(handle, AggNumLevels=agg_nl)
    amg.setup(A)
    nl = int(d["hdr"][3])
    assert amg.num_levels == nl
    for l in range(nl):
        i, j, a = amg.level_A(l).download()
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj), ("A structure", l)
        assert np.array_equal(a, ra), ("A values", l, float(np.max(np.abs(a - ra))))
        if l < nl - 1:
            assert np.array_equal(amg.level_CF(l), d["CF%d" % l]), ("CF", l)
            i, j, a = amg.level_P(l).download()
            pi, pj, pa, _ = refio.csr(d, "P", l)
            assert np.array_equal(i, pi) and np.array_equal(j, pj), ("P structure", l)
            assert np.array_equal(a, pa), ("P values", l, float(np.max(np.abs(a - pa))))
    n = A.local[0]
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4]), (its, int(d["hdr"][4]))
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    if args == ["-n", 50, 50, 50]:
        assert its == 27 and amg.level_A(1).dims[0] == 7654          # SURVEY.md 8c known answer for -agg_nl 1
    amg.destroy(); A.destroy()


def test_second_strength_graph_is_first_touch_distance_two(handle):
    """b200_create_2nd_s against a direct Python restatement of par_strength.c:2326-2400 / :2620-2700"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(["-n", 13, 12, 11, "-pmis", "-rlx", 18, "-mod_rap2", 1])
    si, sj, _, _ = refio.csr(d, "S", 0)
    S = hb.Csr.from_host(handle, si, sj, None)
    cf = handle.pmis(S, 2747)                     # first coarsening: {1, -1, -3}
    hcf = cf.numpy()
    S2 = handle.create_2nd_s(S, cf)
    i2, j2, _ = S2.download(with_data=False)
    f2c = np.cumsum(hcf > 0) - 1
    c2f = np.where(hcf > 0)[0]
    wi, wj = [0], []
    for ic, i1 in enumerate(c2f):
        row = []
        for p in sj[si[i1]:si[i1 + 1]]:
            if hcf[p] > 0 and f2c[p] not in row:
                row.append(f2c[p])
            for q in sj[si[p]:si[p + 1]]:
                if hcf[q] > 0 and f2c[q] != ic and f2c[q] not in row:
                    row.append(f2c[q])
        wj += row
        wi.append(len(wj))
    assert np.array_equal(i2, np.array(wi)) and np.array_equal(j2, np.array(wj))
    S.destroy(); S2.destroy()


def test_aggressive_coarsening_through_the_public_api():
    """`ij -agg_nl 1` call sequence (HYPRE_BoomerAMGSetAggNumLevels) against the reference driver"""
    import os, re, subprocess
    from hypre_ve_b200 import build as b
    exe = b.build_examples()
    flags = ["-laplacian", "-n", "24", "22", "20", "-solver", "1", "-pmis", "-rlx", "18", "-mod_rap2", "1", "-agg_nl", "1"]
    out = subprocess.run([exe] + flags, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    its = int(re.search(r"^Iterations = (\d+)", out.stdout, re.M).group(1))
    ref_ij = os.path.join(refio.ROOT, "oracle", "_ref", "ij")
    if os.path.exists(ref_ij):
        r = subprocess.run([ref_ij] + flags, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert its == int(re.search(r"^Iterations = (\d+)", r.stdout, re.M).group(1))
