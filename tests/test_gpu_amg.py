"""GPU parity tests for the BoomerAMG setup stages, the hierarchy, the V-cycle and PCG.

Checker = the reference's own CPU build (oracle/_ref/ref_dump, single thread), run live on the same
problem.  Bars (BASELINE.json north_star): CF splitting, S/P/A_l sparsity and PCG iteration counts
bit-exact; we additionally require P and A_l *values* to be bit-identical (np.array_equal) because
the setup kernels replay the reference's floating-point order.  Residual history: 1e-10 relative.
"""
import numpy as np
import pytest

import refio

pytestmark = pytest.mark.gpu

CASES = {
    "lap7_24x20x18": ["-n", 24, 20, 18],
    "lap7_aniso": ["-n", 20, 20, 20, "-c", 1, 1, 0.001],
    "lap27_16": ["-n", 16, 16, 16, "-27pt"],
}
_cache = {}


def ref(case):
    if case not in _cache:
        d, out = refio.run_ref(CASES[case] + ["-pmis", "-rlx", 18, "-mod_rap2", 1, "-keepT", 1])
        _cache[case] = d
    return _cache[case]


def up(handle, d, pre, l):
    import hypre_ve_b200 as hb
    i, j, a, _ = refio.csr(d, pre, l)
    return hb.Csr.from_host(handle, i, j, a)


def nlev(d):
    return int(d["hdr"][3])


@pytest.mark.parametrize("case", list(CASES))
def test_strength_each_level(handle, case):
    d = ref(case)
    for l in range(nlev(d) - 1):
        A = up(handle, d, "A", l)
        S = handle.strength(A, 0.25, 1.0)
        i, j, _ = S.download(with_data=False)
        ri, rj, _, _ = refio.csr(d, "S", l)
        assert np.array_equal(i, ri), (case, l)
        assert np.array_equal(j, rj), (case, l)
        S.destroy(); A.destroy()


def test_strength_max_row_sum(handle):
    """all-weak branch (par_strength.c:330-344) with max_row_sum < 1"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(["-n", 12, 11, 10, "-pmis", "-rlx", 18, "-mod_rap2", 1, "-mxrs", 0.9])
    for l in range(nlev(d) - 1):
        A = up(handle, d, "A", l)
        S = handle.strength(A, 0.25, 0.9)
        i, j, _ = S.download(with_data=False)
        ri, rj, _, _ = refio.csr(d, "S", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj)
        S.destroy(); A.destroy()


@pytest.mark.parametrize("case", list(CASES))
def test_pmis_each_level(handle, case):
    import hypre_ve_b200 as hb
    d = ref(case)
    for l in range(nlev(d) - 1):
        ri, rj, _, _ = refio.csr(d, "S", l)
        S = hb.Csr.from_host(handle, ri, rj, None)
        cf = handle.pmis(S, 2747).numpy()
        cf[cf == -3] = -1              # the reference folds SF_PT into F after interpolation (par_lr_interp.c:1888)
        assert np.array_equal(cf, d["CF%d" % l]), (case, l, int((cf != d["CF%d" % l]).sum()))
        S.destroy()


@pytest.mark.parametrize("case", list(CASES))
def test_extpi_interp_each_level(handle, case):
    import hypre_ve_b200 as hb
    d = ref(case)
    for l in range(nlev(d) - 1):
        A = up(handle, d, "A", l)
        ri, rj, _, _ = refio.csr(d, "S", l)
        S = hb.Csr.from_host(handle, ri, rj, None)
        # recover SF_PT (-3): rows of S that are empty were isolated points
        cf = d["CF%d" % l].copy()
        cf[(np.diff(ri) == 0) & (cf < 0)] = -3
        dcf = handle.array(cf.astype(np.int32))
        P = handle.extpi_interp(A, S, dcf, 0.0, 4)
        i, j, a = P.download()
        pi, pj, pa, _ = refio.csr(d, "P", l)
        assert np.array_equal(i, pi), (case, l)
        assert np.array_equal(j, pj), (case, l, int((j != pj).sum()))
        assert np.array_equal(a, pa), (case, l, float(np.max(np.abs(a - pa))))
        P.destroy(); S.destroy(); A.destroy(); dcf.free()


@pytest.mark.parametrize("case", list(CASES))
def test_rap_each_level(handle, case):
    d = ref(case)
    for l in range(nlev(d) - 1):
        A = up(handle, d, "A", l)
        P = up(handle, d, "P", l)
        R = P.transpose()
        Q = A.multiply(P)
        C = R.multiply(Q)
        i, j, a = C.download()
        ci, cj, ca, _ = refio.csr(d, "A", l + 1)
        assert np.array_equal(i, ci), (case, l)
        assert np.array_equal(j, cj), (case, l)
        assert np.array_equal(a, ca), (case, l, float(np.max(np.abs(a - ca))))
        for m in (A, P, R, Q, C):
            m.destroy()


def test_transpose_order(handle):
    """AT rows list source rows ascending (stable counting sort, csr_matop.c:740-767)"""
    import scipy.sparse as sp
    import hypre_ve_b200 as hb
    rng = np.random.default_rng(5)
    M = sp.random(300, 170, density=0.05, random_state=7, format="csr")
    M.sort_indices()
    # scramble the column order inside rows: transpose must not depend on it beyond stability
    A = hb.Csr.from_host(handle, M.indptr, M.indices, M.data)
    T = A.transpose()
    i, j, a = T.download()
    MT = M.T.tocsr(); MT.sort_indices()
    assert np.array_equal(i, MT.indptr) and np.array_equal(j, MT.indices) and np.array_equal(a, MT.data)
    A.destroy(); T.destroy()


@pytest.mark.parametrize("case", list(CASES))
def test_l1_norms(handle, case):
    d = ref(case)
    for l in range(nlev(d) - 1):
        A = up(handle, d, "A", l)
        l1 = handle.l1_norms(A, 1).numpy()
        assert np.array_equal(l1, d["l1_%d" % l]), (case, l)
        A.destroy()


FULL = {
    "lap7_50": (["-n", 50, 50, 50], dict(nx=50, ny=50, nz=50), 15, 4.192356e-09),   # SURVEY.md 8c known answer
    "lap7_31x17x40": (["-n", 31, 17, 40], dict(nx=31, ny=17, nz=40), None, None),
    "lap27_24": (["-n", 24, 24, 24, "-27pt"], dict(nx=24, ny=24, nz=24, pt27=True), None, None),
}


@pytest.mark.parametrize("case", list(FULL))
def test_full_hierarchy_and_pcg(handle, case):
    import hypre_ve_b200 as hb
    args, g, known_its, known_rel = FULL[case]
    d, out = refio.run_ref(args + ["-pmis", "-rlx", 18, "-mod_rap2", 1, "-keepT", 1])
    if g.get("pt27"):
        A = hb.ParCsr.laplacian27(handle, g["nx"], g["ny"], g["nz"])
    else:
        A = hb.ParCsr.laplacian(handle, g["nx"], g["ny"], g["nz"])
    amg = hb.Amg(handle, KeepS=1)
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    for l in range(nlev(d)):
        i, j, a = amg.level_A(l).download()
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj), (case, "A", l)
        assert np.array_equal(a, ra), (case, "A values", l)
        if l < nlev(d) - 1:
            assert np.array_equal(amg.level_CF(l), d["CF%d" % l]), (case, "CF", l)
            i, j, a = amg.level_P(l).download()
            pi, pj, pa, _ = refio.csr(d, "P", l)
            assert np.array_equal(i, pi) and np.array_equal(j, pj) and np.array_equal(a, pa), (case, "P", l)
            assert np.array_equal(amg.level_l1(l), d["l1_%d" % l]), (case, "l1", l)
    n = A.local[0]
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    ref_its = int(d["hdr"][4])
    assert its == ref_its, (its, ref_its)
    if known_its is not None:
        assert its == known_its
        assert abs(rel - known_rel) / known_rel < 1e-5      # printed to 7 digits by the reference
    rn = d["norms"]
    assert len(norms) == len(rn)
    # residual history: 1e-10 relative to ||r_0|| (north star), plus a relative check on each entry
    assert np.max(np.abs(norms - rn)) / rn[0] < 1e-10, np.max(np.abs(norms - rn)) / rn[0]
    xs = x.numpy()
    assert np.max(np.abs(xs - d["x"])) / np.max(np.abs(d["x"])) < 1e-9
    amg.destroy(); A.destroy()


def test_vcycle_matches_reference_preconditioner(handle):
    """one application of the preconditioner to b = 1 equals the reference's first search direction:
    checked through the first PCG residual (depends on C*b only)."""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(["-n", 20, 20, 20, "-pmis", "-rlx", 18, "-mod_rap2", 1, "-max_iter", 2])
    A = hb.ParCsr.laplacian(handle, 20, 20, 20)
    amg = hb.Amg(handle)
    amg.setup(A)
    n = A.local[0]
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=2)
    assert its == 2
    assert np.allclose(norms, d["norms"][:3], rtol=1e-12, atol=0)
    # amg_solve with a non-zero initial guess takes the general path; compare with zero-guess result
    u0 = handle.zeros(n)
    amg.solve(b, u0)
    u1 = handle.zeros(n)
    handle.fill(u1, 0.0)
    amg.solve(b, u1)
    assert np.array_equal(u0.numpy(), u1.numpy())
    amg.destroy(); A.destroy()


def test_unsupported_configurations_fail_loudly(handle):
    import hypre_ve_b200 as hb
    A = hb.ParCsr.laplacian(handle, 8, 8, 8)
    for k, v in [("CoarsenType", 6), ("InterpType", 0), ("RelaxType", 5), ("AggNumLevels", -1), ("RAP2", 1)]:
        amg = hb.Amg(handle)
        amg.set(k, v)
        with pytest.raises(hb.B200Error):
            amg.setup(A)
        amg.destroy()
    A.destroy()


def test_general_hbm_scratch_path_is_identical(handle, monkeypatch):
    """b200_setup.cu's thread-per-row kernels (used when a row outgrows the warp kernels' shared memory)
    must give the same bits as the warp-per-row kernels."""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(["-n", 22, 19, 17, "-pmis", "-rlx", 18, "-mod_rap2", 1, "-keepT", 1])
    for force in ("0", "1"):
        monkeypatch.setenv("B200_FORCE_GENERAL_SETUP", force)
        A = hb.ParCsr.laplacian(handle, 22, 19, 17)
        amg = hb.Amg(handle)
        amg.setup(A)
        assert amg.num_levels == nlev(d)
        for l in range(nlev(d)):
            i, j, a = amg.level_A(l).download()
            ri, rj, ra, _ = refio.csr(d, "A", l)
            assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra), (force, l)
            if l < nlev(d) - 1:
                i, j, a = amg.level_P(l).download()
                pi, pj, pa, _ = refio.csr(d, "P", l)
                assert np.array_equal(i, pi) and np.array_equal(j, pj) and np.array_equal(a, pa), (force, "P", l)
        amg.destroy(); A.destroy()


@pytest.mark.parametrize("args", [["-n", 24, 20, 18], ["-n", 16, 16, 16, "-27pt"], ["-n", 20, 20, 20, "-c", 1, 1, 0.001],
                                  ["-n", 30, 30, 30, "-agg_nl", 1]])
def test_default_fused_galerkin_product_order(handle, args):
    """ModuleRAP2 0 = hypre_BoomerAMGBuildCoarseOperatorKT, the reference's default: (R A) P.  Hierarchy bit-exact,
    same iteration count as the reference run WITHOUT -mod_rap2 1"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(args + ["-pmis", "-rlx", 18])
    nx, ny, nz = args[1:4]
    A = hb.ParCsr.laplacian27(handle, nx, ny, nz) if "-27pt" in args else \
        hb.ParCsr.laplacian(handle, nx, ny, nz, c=tuple(args[5:8]) if "-c" in args else (1.0, 1.0, 1.0))
    amg = hb.Amg(handle, ModuleRAP2=0, AggNumLevels=(1 if "-agg_nl" in args else 0))
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    for l in range(nlev(d)):
        i, j, a = amg.level_A(l).download()
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra), l
    n = A.local[0]
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4])
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    amg.destroy(); A.destroy()


@pytest.mark.parametrize("args,w", [(["-n", 24, 20, 18], 1.0), (["-n", 14, 14, 14, "-27pt"], 1.0)])
def test_weighted_jacobi_relax_7(handle, args, w):
    """relax 7: Jacobi through the matvec with the diagonal as "l1 norm" (par_relax.c:3463-3490, option 5 norms)"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(args + ["-pmis", "-rlx", 7, "-mod_rap2", 1])
    nx, ny, nz = args[1:4]
    A = hb.ParCsr.laplacian27(handle, nx, ny, nz) if "-27pt" in args else hb.ParCsr.laplacian(handle, nx, ny, nz)
    amg = hb.Amg(handle, RelaxType=7)
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    for l in range(nlev(d) - 1):
        assert np.array_equal(amg.level_l1(l), d["l1_%d" % l]), l           # option 5: the diagonal
    n = A.local[0]
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4])
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    amg.destroy(); A.destroy()


@pytest.mark.parametrize("args", [["-n", 24, 20, 18], ["-n", 14, 14, 14, "-27pt"], ["-n", 20, 20, 20, "-c", 1, 1, 0.001],
                                  ["-n", 40, 40, 40]])
def test_chebyshev_smoother_relax_16(handle, args):
    """relax 16 (par_cheby.c): CG-Lanczos spectrum estimate from the reference's random vector, order-2 polynomial;
    iteration counts equal the reference's, residual history to 1e-10"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(args + ["-pmis", "-rlx", 16, "-mod_rap2", 1])
    nx, ny, nz = args[1:4]
    A = hb.ParCsr.laplacian27(handle, nx, ny, nz) if "-27pt" in args else \
        hb.ParCsr.laplacian(handle, nx, ny, nz, c=tuple(args[5:8]) if "-c" in args else (1.0, 1.0, 1.0))
    amg = hb.Amg(handle, RelaxType=16)
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    n = A.local[0]
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4]), (its, int(d["hdr"][4]))
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    assert np.max(np.abs(x.numpy() - d["x"])) / np.max(np.abs(d["x"])) < 1e-9
    amg.destroy(); A.destroy()


@pytest.mark.parametrize("extra,params", [
    (["-ns", 2], dict(NumSweeps=2)),
    (["-mu", 2], dict(CycleType=2)),
    (["-mu", 2, "-ns", 2], dict(CycleType=2, NumSweeps=2)),
    (["-fmg"], dict(FCycle=1)),
    (["-ns_coarse", 2, "-ns", 3], dict(NumSweeps=3, NumSweepsCoarse=2)),
])
@pytest.mark.parametrize("rlx", [18, -1, 16])
def test_cycle_shapes(handle, extra, params, rlx):
    """W / F cycles and several sweeps per visit: the reference's level-counter state machine (par_cycle.c:180-622);
    residual history of AMG-PCG against the reference CPU build"""
    import hypre_ve_b200 as hb
    args = ["-n", 21, 19, 17, "-pmis"] + (["-rlx", rlx] if rlx > -1 else []) + extra
    d, _ = refio.run_ref(args)
    A = hb.ParCsr.laplacian(handle, 21, 19, 17)
    kw = dict(RelaxType=rlx, ModuleRAP2=0) if rlx > -1 else dict(RelaxType=13, RelaxTypeUp=14, ModuleRAP2=0)
    amg = hb.Amg(handle, **kw, **params)
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    n = A.local[0]
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4]), (its, int(d["hdr"][4]))
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    assert np.max(np.abs(x.numpy() - d["x"])) / np.max(np.abs(d["x"])) < 1e-9
    amg.destroy(); A.destroy()


@pytest.mark.parametrize("args,params", [
    (["-n", 22, 19, 17, "-perturb", 1, "-rlx", 18], dict(ModuleRAP2=0)),
    (["-n", 22, 19, 17, "-perturb", 2, "-rlx", 18, "-mod_rap2", 1], dict(ModuleRAP2=1)),
    (["-n", 13, 12, 11, "-27pt", "-perturb", 7, "-rlx", 18], dict(ModuleRAP2=0)),
    (["-n", 20, 20, 20, "-perturb", 11, "-rlx", 18, "-th", 0.5, "-Pmx", 6], dict(ModuleRAP2=0, StrongThreshold=0.5, PMaxElmts=6)),
    (["-n", 18, 18, 18, "-perturb", 5, "-rlx", 18, "-agg_nl", 1], dict(ModuleRAP2=0, AggNumLevels=1)),
    (["-n", 18, 16, 14, "-perturb", 3], dict(ModuleRAP2=0, RelaxType=13, RelaxTypeUp=14)),
    (["-n", 12, 12, 12, "-27pt", "-perturb", 9, "-rlx", 16], dict(ModuleRAP2=0, RelaxType=16)),
    (["-n", 40, 36, 30, "-perturb", 4, "-rlx", 18], dict(ModuleRAP2=0)),
])
def test_non_laplacian_operator_hierarchy_and_pcg(handle, args, params):
    """An SPD operator that is NOT a Laplacian (ref_dump -perturb: random symmetric magnitudes spanning weak and
    strong connections, 1/16 positive off-diagonals) uploaded from host CSR: every level bit-identical to the
    reference CPU build, residual history to 1e-10.  Exercises the sign filters and weak-connection branches of
    strength / ext+i / multipass that the stencil operators never reach."""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(args + ["-pmis", "-keepT", 1])
    i0, j0, a0, _ = refio.csr(d, "A", 0)
    assert (a0[i0[:-1]] > 0).all() and (a0 > 0).sum() > i0.size - 1          # positive diagonal AND positive off-diagonals
    A = hb.ParCsr.from_host(handle, i0, j0, a0)
    amg = hb.Amg(handle, **params)
    amg.setup(A)
    assert amg.num_levels == nlev(d)
    for l in range(nlev(d)):
        i, j, a = amg.level_A(l).download()
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra), ("A", l)
        if l < nlev(d) - 1:
            assert np.array_equal(amg.level_CF(l), d["CF%d" % l]), ("CF", l)
            i, j, a = amg.level_P(l).download()
            pi, pj, pa, _ = refio.csr(d, "P", l)
            assert np.array_equal(i, pi) and np.array_equal(j, pj) and np.array_equal(a, pa), ("P", l)
    n = i0.size - 1
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4]), (its, int(d["hdr"][4]))
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    assert np.max(np.abs(x.numpy() - d["x"])) / np.max(np.abs(d["x"])) < 1e-9
    amg.destroy(); A.destroy()
