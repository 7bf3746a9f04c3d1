"""Multi-rank (row-partitioned) path, exercised with N ranks = N host threads on ONE GPU (the
in-process comm backend: B200_PROFILING.md asks for exactly this when fewer GPUs than ranks).

Parity statements checked:
  * the gathered N-rank operator equals the reference generator's matrix for the same -P grid
    numbering (checked through a single-GPU setup on the gathered matrix);
  * the N-rank hierarchy (every A_l, P_l, CF_l) is bit-identical to the single-GPU hierarchy built
    from the gathered matrix -- for every rank count and process grid;
  * for z-slab grids (1,1,N) the numbering is the lexicographic one, so the hierarchy also equals
    the reference CPU build's (oracle/_ref) bit for bit;
  * PCG iteration counts match the single-GPU / reference run, residual histories to 1e-10.
"""
import threading

import numpy as np
import pytest

import refio

pytestmark = pytest.mark.gpu


def run_ranks(nranks, fn):
    """run fn(rank, handle, comm) on nranks threads sharing one GPU; returns the list of results"""
    import hypre_ve_b200 as hb
    group = hb.Comm.group_create(nranks)
    out, err = [None] * nranks, [None] * nranks

    def body(r):
        h = None
        try:
            h = hb.Handle(0)
            c = hb.Comm.threads(h, group, r)
            out[r] = fn(r, h, c)
        except BaseException as e:      # noqa: a failing rank aborts the group, so its peers return an error instead of waiting
            err[r] = e
            hb.Comm.group_abort(group)
        finally:
            if h is not None:
                try:
                    h.close()           # b200_finalize: every slab of this rank goes back to the driver
                except Exception as e:  # noqa
                    err[r] = err[r] or e
    ts = [threading.Thread(target=body, args=(r,)) for r in range(nranks)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    alive = any(t.is_alive() for t in ts)
    if not alive:
        hb.Comm.group_destroy(group)
    # report the root cause, not the "group aborted" echo of the peers
    first = [e for e in err if e is not None and "rank group aborted" not in str(e)] or [e for e in err if e is not None]
    if first:
        raise first[0]
    assert not alive, "a rank thread is still running"
    return out


def gather(parts):
    """stack per-rank (first_row, i, j, a) blocks into one global CSR"""
    parts = sorted(parts, key=lambda p: p[0])
    I = [np.zeros(1, np.int64)]
    off = 0
    for _, i, j, a in parts:
        I.append(i[1:].astype(np.int64) + off)
        off += int(i[-1])
    return (np.concatenate(I).astype(np.int32), np.concatenate([p[2] for p in parts]), np.concatenate([p[3] for p in parts]))


def dist_case(nranks, grid, dims, stencil=7, solve=True, solver=1, **params):
    import hypre_ve_b200 as hb
    nx, ny, nz = dims
    P, Q, R = grid

    def fn(r, h, c):
        if callable(stencil):                                    # the caller's own rows: stencil(rank, handle, comm) -> DistMatrix
            A = stencil(r, h, c)
        elif isinstance(stencil, dict) and "alpha" in stencil:   # GenerateRotate7pt, two-dimensional (nz = R = 1)
            A = hb.DistMatrix.rotate7pt(h, c, nx, ny, P, Q, stencil["alpha"], stencil["eps"])
        elif isinstance(stencil, dict):       # GenerateDifConv with these -c / -a / -atype values (nonsymmetric)
            A = hb.DistMatrix.difconv(h, c, nx, ny, nz, P, Q, R, **stencil)
        else:
            A = hb.DistMatrix.laplacian(h, c, nx, ny, nz, P, Q, R, stencil)
        prm = hb.Amg(h, **params)
        amg = hb.DistAmg(h, c, prm, A)
        lv = []
        for l in range(amg.num_levels):
            MA = amg.level_A(l)
            ent = {"A": (MA.info["first_row"],) + MA.download()}
            if l < amg.num_levels - 1:
                MP = amg.level_P(l)
                ent["P"] = (MP.info["first_row"],) + MP.download()
                ent["cf"] = (MA.info["first_row"], amg.level_cf(l))
            lv.append(ent)
        res = {"levels": lv}
        if solve:
            b = A.vector(1.0)
            x = A.vector(0.0)
            if solver == 3:
                its, rel, norms, _ = hb.dist_gmres(h, c, A, amg, b, x, tol=1e-8, max_iter=100)
            elif solver == 9:
                its, rel, norms, _ = hb.dist_bicgstab(h, c, A, amg, b, x, tol=1e-8, max_iter=100)
            else:
                its, rel, norms = hb.dist_pcg(h, c, A, amg, b, x, tol=1e-8, max_iter=100)
            res.update(its=its, rel=rel, norms=norms, x=(A.info["first_row"], x.numpy()[:A.info["local_rows"]]))
        return res
    return run_ranks(nranks, fn)


def single_gpu_on(handle, i, j, a, solver=1, **params):
    import hypre_ve_b200 as hb
    A = hb.ParCsr.from_host(handle, i, j, a)
    amg = hb.Amg(handle, **params)
    amg.setup(A)
    lv = []
    for l in range(amg.num_levels):
        ent = {"A": amg.level_A(l).download()}
        if l < amg.num_levels - 1:
            ent["P"] = amg.level_P(l).download()
            ent["cf"] = amg.level_CF(l)
        lv.append(ent)
    n = i.size - 1
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    if solver == 3:
        its, rel, norms, _ = handle.gmres(A, amg, b, x, tol=1e-8, max_iter=100)
    elif solver == 9:
        its, rel, norms, _ = handle.bicgstab(A, amg, b, x, tol=1e-8, max_iter=100)
    else:
        its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    out = dict(levels=lv, its=its, rel=rel, norms=norms, x=x.numpy())
    amg.destroy(); A.destroy()
    return out


@pytest.mark.parametrize("nranks,grid,dims,stencil", [
    (2, (1, 1, 2), (12, 11, 10), 7),
    (3, (1, 1, 3), (9, 8, 13), 7),
    (4, (2, 2, 1), (13, 12, 9), 7),
    (8, (2, 2, 2), (14, 13, 12), 7),
    (2, (2, 1, 1), (10, 9, 8), 27),
    (4, (1, 2, 2), (9, 10, 11), 27),
])
def test_nrank_hierarchy_equals_single_gpu(handle, nranks, grid, dims, stencil):
    check_against_single_gpu(handle, nranks, grid, dims, stencil)


@pytest.mark.parametrize("nranks,grid,dims,stencil", [
    (2, (1, 1, 2), (12, 11, 10), 7), (4, (2, 2, 1), (13, 12, 9), 7), (8, (2, 2, 2), (14, 13, 12), 7), (3, (3, 1, 1), (10, 9, 8), 27),
])
def test_nrank_fused_galerkin_order_equals_single_gpu(handle, nranks, grid, dims, stencil):
    """ModuleRAP2 0, the reference default (R A) P, across ranks"""
    check_against_single_gpu(handle, nranks, grid, dims, stencil, ModuleRAP2=0)


@pytest.mark.parametrize("nranks,grid,dims,stencil,seq,params", [
    (2, (1, 1, 2), (12, 11, 10), 7, 500, {}),                 # level 1 (a few hundred rows) and below replicated
    (4, (2, 2, 1), (13, 12, 9), 7, 100000, {}),               # everything below the finest level replicated
    (8, (2, 2, 2), (14, 13, 12), 7, 120, dict(ModuleRAP2=0)),  # only the last levels
    (3, (3, 1, 1), (10, 9, 8), 27, 200, dict(ModuleRAP2=0)),
    (4, (1, 2, 2), (16, 14, 12), 7, 600, dict(AggNumLevels=1)),
])
def test_nrank_replicated_coarse_levels_equal_single_gpu(handle, nranks, grid, dims, stencil, seq, params):
    """SeqThreshold (HYPRE_BoomerAMGSetSeqThreshold): levels with at most that many rows in total are gathered onto every
    rank and built / cycled redundantly as one single-GPU hierarchy (no halo exchange below that level).  Every level --
    distributed or replicated -- stays bit-identical to the single-GPU hierarchy of the gathered operator; iteration count
    equal, residual history to 1e-10."""
    check_against_single_gpu(handle, nranks, grid, dims, stencil, SeqThreshold=seq, **params)


@pytest.mark.parametrize("nranks,grid,dims,stencil,agg", [
    (2, (1, 1, 2), (12, 11, 10), 7, 1), (3, (1, 3, 1), (14, 15, 9), 7, 1), (4, (2, 2, 1), (16, 14, 12), 7, 2),
    (8, (2, 2, 2), (16, 16, 16), 7, 1), (2, (2, 1, 1), (10, 9, 8), 27, 1),
])
def test_nrank_aggressive_coarsening_equals_single_gpu(handle, nranks, grid, dims, stencil, agg):
    """-agg_nl L across ranks: distance-two strength graph from fetched S rows, second PMIS over the halo of S2,
    multipass interpolation with pass numbers and pass rows exchanged over the halo of A"""
    check_against_single_gpu(handle, nranks, grid, dims, stencil, AggNumLevels=agg)
    check_against_single_gpu(handle, nranks, grid, dims, stencil, AggNumLevels=agg, ModuleRAP2=0)


@pytest.mark.parametrize("nranks,grid,dims,gen", [
    (2, (1, 1, 2), (12, 11, 10), dict()),
    (4, (2, 2, 1), (13, 12, 9), dict(a=(3.0, -2.0, 1.0), atype=3)),
    (8, (2, 2, 2), (14, 13, 12), dict(a=(2.0, 1.0, 0.0), atype=1)),
    (3, (1, 3, 1), (9, 12, 8), dict(a=(10.0, 10.0, 10.0), atype=2)),
])
def test_nrank_nonsymmetric_operator_equals_single_gpu(handle, nranks, grid, dims, gen):
    """GenerateDifConv across ranks (par_difconv.c): a_ij != a_ji, so every exchange that would be redundant for a
    symmetric operator (S^T counts of PMIS, a_ki lookups of ext+i in fetched rows, R = P^T) is exercised for real;
    N-rank hierarchy bit-identical to the single-GPU hierarchy of the gathered matrix, both Galerkin orders"""
    check_against_single_gpu(handle, nranks, grid, dims, gen)
    check_against_single_gpu(handle, nranks, grid, dims, gen, ModuleRAP2=0)


@pytest.mark.parametrize("nranks,grid,dims,gen,solver", [
    (2, (1, 1, 2), (12, 11, 10), dict(), 3),
    (4, (2, 2, 1), (13, 12, 9), dict(a=(3.0, -2.0, 1.0), atype=3), 3),
    (8, (2, 2, 2), (14, 13, 12), dict(a=(2.0, 1.0, 0.0), atype=1), 9),
    (3, (3, 1, 1), (12, 9, 8), 7, 3),
    (2, (1, 2, 1), (10, 12, 8), dict(), 9),
])
def test_nrank_gmres_and_bicgstab_equal_single_gpu(handle, nranks, grid, dims, gen, solver):
    """b200_dist_gmres_solve / b200_dist_bicgstab_solve: the loops of b200_krylov.cu over the row-partitioned operator and
    the distributed cycle; iteration count equal to the single-GPU run on the gathered matrix, history to 1e-10"""
    check_against_single_gpu(handle, nranks, grid, dims, gen, solver=solver)


@pytest.mark.parametrize("nranks,grid,dims,gen", [
    (2, (2, 1, 1), (14, 12, 1), dict(alpha=45.0, eps=0.001)),
    (4, (2, 2, 1), (17, 16, 1), dict(alpha=30.0, eps=0.01)),
    (3, (1, 3, 1), (12, 19, 1), dict(alpha=120.0, eps=0.1)),
    (6, (3, 2, 1), (20, 18, 1), dict(alpha=60.0, eps=0.05)),
])
def test_nrank_rotate7pt_equals_single_gpu(handle, nranks, grid, dims, gen):
    """GenerateRotate7pt on a P x Q process grid: diagonal neighbours cross box corners (hypre_map2 with p-1, q-1);
    N-rank hierarchy bit-identical to the single-GPU hierarchy of the gathered matrix"""
    check_against_single_gpu(handle, nranks, grid, dims, gen, ModuleRAP2=0)


def test_yslab_rotate7pt_equals_reference_cpu_build():
    """(1, N) slabs keep the lexicographic numbering of the 2-D grid: gathered operator and hierarchy = the reference's"""
    d, _ = refio.run_ref(["-n", 18, 20, 1, "-rotate", "-alpha", 45, "-eps", 0.001, "-pmis", "-rlx", 18, "-keepT", 1])
    res = dist_case(4, (1, 4, 1), (18, 20, 1), dict(alpha=45.0, eps=0.001), ModuleRAP2=0)
    nl = int(d["hdr"][3])
    assert len(res[0]["levels"]) == nl
    for l in range(nl):
        i, j, a = gather([r["levels"][l]["A"] for r in res])
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra), l
    assert res[0]["its"] == int(d["hdr"][4])


def uneven_cuts(N, nranks):
    """contiguous row blocks of different sizes (nothing like a box decomposition)"""
    w = np.array([1.0 + 0.7 * ((3 * r) % nranks) for r in range(nranks)])
    cuts = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * N).astype(int)])
    cuts[-1] = N
    assert np.all(np.diff(cuts) > 0)
    return cuts


USER_ROWS = [
    (["-n", 14, 13, 12, "-perturb", 3, "-rlx", 18], 3, "rows"),
    (["-n", 14, 13, 12, "-perturb", 3, "-rlx", 18], 3, "ij"),
    (["-n", 16, 15, 11, "-difconv", "-a", 3, -2, 1, "-atype", 3, "-rlx", 18], 5, "ij"),
]
# two more cases and the bad-input test live in tests/test_gpu_userrows.py (they sort last: see its docstring)
USER_ROWS_MORE = [
    (["-n", 26, 23, 1, "-rotate", "-alpha", 45, "-eps", 0.001, "-rlx", 18], 4, "rows"),
    (["-n", 10, 10, 10, "-27pt", "-rlx", 18, "-agg_nl", 1], 2, "ij"),
]


@pytest.mark.parametrize("args,nranks,how", USER_ROWS)
def test_operator_from_the_callers_rows_equals_reference_cpu_build(handle, args, nranks, how):
    check_callers_rows(handle, args, nranks, how)


def check_callers_rows(handle, args, nranks, how):
    """b200_dist_matrix_create_from_host / _from_ij: every rank hands in an arbitrary contiguous block of rows of the
    reference's operator (global column ids) -- as CSR arrays, or as a scrambled SetValues / AddToValues stream merged on
    the device -- and the N-rank hierarchy, iteration count and residual history are the reference np=1 run's"""
    import hypre_ve_b200 as hb
    import ijstream
    d, _ = refio.run_ref(args + ["-pmis", "-keepT", 1])
    I, J, a, _ = refio.csr(d, "A", 0)
    N = I.size - 1
    cuts = uneven_cuts(N, nranks)
    agg = int(args[args.index("-agg_nl") + 1]) if "-agg_nl" in args else 0

    def build(r, h, c):
        r0, r1 = int(cuts[r]), int(cuts[r + 1])
        Il = (I[r0:r1 + 1] - I[r0]).astype(np.int32)
        Jl, al = J[I[r0]:I[r1]], a[I[r0]:I[r1]]
        if how == "rows":
            return hb.DistMatrix.from_rows(h, c, Il, Jl, al)
        asm = hb.IJAssembler(h, r0, r1 - 1, global_cols=N)
        # the -ijbuild mode-1 stream of this block (half the value set, rotated; the other half added, rotated differently),
        # with row numbers shifted back to global ones
        for add, rows, ncols, cols, vals in ijstream.calls_before_assembly(Il, Jl, al, 1):
            assert asm.set_values(ncols, np.asarray(rows) + r0, cols, vals, add=bool(add)) == 0
        M = hb.DistMatrix.from_ij(h, c, asm)
        asm.destroy()
        i2, j2, a2 = M.download()
        assert np.array_equal(i2, Il) and np.array_equal(a2[i2[:-1]], al[Il[:-1]])            # diagonal first, values restored
        assert np.array_equal(np.sort(j2), np.sort(Jl))
        return M

    if how == "ij":      # the assembled entry order differs from the reference's generator order: compare with the single-GPU run
        res = dist_case(nranks, (nranks, 1, 1), (N, 1, 1), build, ModuleRAP2=0, AggNumLevels=agg)
        gi, gj, ga = gather([r["levels"][0]["A"] for r in res])
        ref = single_gpu_on(handle, gi, gj, ga, ModuleRAP2=0, AggNumLevels=agg)
        nl = len(ref["levels"])
        assert all(len(r["levels"]) == nl for r in res)
        for l in range(nl):
            i, j, v = gather([r["levels"][l]["A"] for r in res])
            ri, rj, ra = ref["levels"][l]["A"]
            assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(v, ra), l
        assert res[0]["its"] == ref["its"] and np.max(np.abs(res[0]["norms"] - ref["norms"])) / ref["norms"][0] < 1e-10
        return
    res = dist_case(nranks, (nranks, 1, 1), (N, 1, 1), build, ModuleRAP2=0, AggNumLevels=agg)
    nl = int(d["hdr"][3])
    assert all(len(r["levels"]) == nl for r in res)
    for l in range(nl):
        i, j, v = gather([r["levels"][l]["A"] for r in res])
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(v, ra), l
        if l < nl - 1:
            i, j, v = gather([r["levels"][l]["P"] for r in res])
            pi, pj, pa, _ = refio.csr(d, "P", l)
            assert np.array_equal(i, pi) and np.array_equal(j, pj) and np.array_equal(v, pa), ("P", l)
    assert res[0]["its"] == int(d["hdr"][4])
    assert np.max(np.abs(res[0]["norms"] - d["norms"])) / d["norms"][0] < 1e-10


def test_zslab_difconv_gmres_equals_reference_cpu_build():
    """3 z-slabs, upwind convection-diffusion, AMG-GMRES: hierarchy and iteration count of the reference np=1 run"""
    dims = (13, 12, 14)
    d, _ = refio.run_ref(["-n", *dims, "-difconv", "-a", 3, -2, 1, "-atype", 3, "-pmis", "-rlx", 18, "-mod_rap2", 1, "-keepT", 1,
                          "-solver", 3])
    res = dist_case(3, (1, 1, 3), dims, dict(a=(3.0, -2.0, 1.0), atype=3), solver=3)
    assert res[0]["its"] == int(d["hdr"][4])
    assert np.max(np.abs(res[0]["norms"] - d["norms"])) / d["norms"][0] < 1e-10
    x = np.concatenate([v for _, v in sorted((r["x"] for r in res), key=lambda t: t[0])])
    assert np.max(np.abs(x - d["x"])) / np.max(np.abs(d["x"])) < 1e-9


def test_zslab_difconv_equals_reference_cpu_build():
    """z-slabs keep the lexicographic numbering: the gathered 3-rank convection-diffusion hierarchy is the reference's"""
    dims = (13, 12, 14)
    d, _ = refio.run_ref(["-n", *dims, "-difconv", "-a", 3, -2, 1, "-atype", 3, "-pmis", "-rlx", 18, "-mod_rap2", 1, "-keepT", 1])
    res = dist_case(3, (1, 1, 3), dims, dict(a=(3.0, -2.0, 1.0), atype=3))
    nl = int(d["hdr"][3])
    assert len(res[0]["levels"]) == nl
    for l in range(nl):
        i, j, a = gather([r["levels"][l]["A"] for r in res])
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra), l
    assert res[0]["its"] == int(d["hdr"][4])
    assert np.max(np.abs(res[0]["norms"] - d["norms"])) / d["norms"][0] < 1e-10


def check_against_single_gpu(handle, nranks, grid, dims, stencil, solver=1, **params):
    res = dist_case(nranks, grid, dims, stencil, solver=solver, **params)
    nl = len(res[0]["levels"])
    assert all(len(r["levels"]) == nl for r in res)
    gi, gj, ga = gather([r["levels"][0]["A"] for r in res])
    ref = single_gpu_on(handle, gi, gj, ga, solver=solver, **params)
    assert len(ref["levels"]) == nl
    for l in range(nl):
        i, j, a = gather([r["levels"][l]["A"] for r in res])
        ri, rj, ra = ref["levels"][l]["A"]
        assert np.array_equal(i, ri) and np.array_equal(j, rj), ("A structure", l)
        assert np.array_equal(a, ra), ("A values", l)
        if l < nl - 1:
            i, j, a = gather([r["levels"][l]["P"] for r in res])
            pi, pj, pa = ref["levels"][l]["P"]
            assert np.array_equal(i, pi) and np.array_equal(j, pj) and np.array_equal(a, pa), ("P", l)
            cf = np.concatenate([c for _, c in sorted((r["levels"][l]["cf"] for r in res), key=lambda t: t[0])])
            assert np.array_equal(cf, ref["levels"][l]["cf"]), ("CF", l)
    assert all(r["its"] == ref["its"] for r in res)
    norms = res[0]["norms"]
    assert np.max(np.abs(norms - ref["norms"])) / ref["norms"][0] < 1e-10
    x = np.concatenate([v for _, v in sorted((r["x"] for r in res), key=lambda t: t[0])])
    assert np.max(np.abs(x - ref["x"])) / np.max(np.abs(ref["x"])) < 1e-9


@pytest.mark.parametrize("nranks", [2, 4])
def test_zslab_partition_equals_reference_cpu_build(nranks):
    """(1,1,N) slabs keep the lexicographic numbering => same hierarchy as the reference np=1 run"""
    dims = (16, 15, 14)
    d, _ = refio.run_ref(["-n", *dims, "-pmis", "-rlx", 18, "-mod_rap2", 1, "-keepT", 1])
    res = dist_case(nranks, (1, 1, nranks), dims, 7)
    nl = int(d["hdr"][3])
    assert len(res[0]["levels"]) == nl
    for l in range(nl):
        i, j, a = gather([r["levels"][l]["A"] for r in res])
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra), l
    assert res[0]["its"] == int(d["hdr"][4])
    assert np.max(np.abs(res[0]["norms"] - d["norms"])) / d["norms"][0] < 1e-10


def test_dist_matvec_matches_gathered(handle):
    import scipy.sparse as sp
    import hypre_ve_b200 as hb
    dims, grid, nranks = (11, 10, 9), (2, 2, 1), 4
    rng = np.random.default_rng(0)
    xg = rng.standard_normal(dims[0] * dims[1] * dims[2])

    def fn(r, h, c):
        A = hb.DistMatrix.laplacian(h, c, *dims, *grid, 7)
        inf = A.info
        x = A.vector(0.0)
        xl = np.zeros(x.n)
        xl[:inf["local_rows"]] = xg[inf["first_row"]:inf["first_row"] + inf["local_rows"]]
        x.upload(xl)
        y = h.zeros(inf["local_rows"])
        A.matvec(1.0, x, 0.0, None, y)
        return (inf["first_row"], A.download(), y.numpy())
    res = run_ranks(nranks, fn)
    gi, gj, ga = gather([(f,) + m for f, m, _ in res])
    M = sp.csr_matrix((ga, gj, gi), shape=(xg.size, xg.size))
    y = np.concatenate([v for _, _, v in sorted(res, key=lambda t: t[0])])
    np.testing.assert_allclose(y, M @ xg, rtol=0, atol=1e-12)


@pytest.mark.parametrize("nranks,T,relax_type", [(2, 1, 13), (3, 1, 14), (2, 5, 8), (4, 3, 13), (4, 700, 8)])
def test_hybrid_gauss_seidel_sweep_across_ranks(nranks, T, relax_type):
    """one hypre_BoomerAMGRelax call of the l1 hybrid Gauss-Seidel family on N ranks x T blocks equals the
    sequential restatement on the gathered matrix with the same block list: Gauss-Seidel inside a block,
    pre-sweep values (and half-weighted l1 contributions) for everything outside -- other blocks and other ranks"""
    import hypre_ve_b200 as hb
    from test_gpu_gs import seq_relax, l1_opt4, blocks_of
    dims = (9, 8, 4 * nranks)
    N = dims[0] * dims[1] * dims[2]
    rng = np.random.default_rng(3)
    fg, ug = rng.standard_normal(N), rng.standard_normal(N)

    def fn(r, h, c):
        A = hb.DistMatrix.laplacian(h, c, *dims, 1, 1, nranks, 7)
        inf = A.info
        lo, n = inf["first_row"], inf["local_rows"]
        f = h.array(fg[lo:lo + n].copy())
        u = A.vector(0.0)
        ul = np.zeros(u.n); ul[:n] = ug[lo:lo + n]
        u.upload(ul)
        A.relax_gs(relax_type, T, f, u)
        return (lo, A.download(), u.numpy()[:n])
    res = run_ranks(nranks, fn)
    gi, gj, ga = gather([(lo,) + m for lo, m, _ in res])
    blocks = []
    for lo, m, _ in sorted(res, key=lambda t: t[0]):
        blocks += [(lo + ns, lo + ne) for ns, ne in blocks_of(len(m[0]) - 1, T)]
    l1 = l1_opt4(gi, gj, ga, 0, blocks)
    want = seq_relax(gi, gj, ga, fg, l1, ug, relax_type in (13, 8), relax_type in (14, 8), False, 0, blocks)
    got = np.concatenate([v for _, _, v in sorted(res, key=lambda t: t[0])])
    assert np.array_equal(got, want), float(np.max(np.abs(got - want)))


@pytest.mark.parametrize("nranks,grid,dims,T", [(1, (1, 1, 1), (12, 11, 10), 1), (1, (1, 1, 1), (12, 11, 10), 6),
                                               (2, (1, 1, 2), (12, 11, 10), 1), (4, (2, 2, 1), (14, 12, 9), 4)])
def test_dist_amg_pcg_with_hybrid_gauss_seidel(handle, nranks, grid, dims, T):
    """BoomerAMG-PCG with the default 13-down / 14-up smoother across ranks: one rank reproduces the
    single-GPU path exactly; N ranks converge with a comparable iteration count (the decomposition --
    N ranks x T blocks -- is part of the smoother's definition, par_relax.c:4352-4412)"""
    import hypre_ve_b200 as hb

    def fn(r, h, c):
        A = hb.DistMatrix.laplacian(h, c, *dims, *grid, 7)
        prm = hb.Amg(h, RelaxType=13, RelaxTypeUp=14, GSBlocks=T)
        amg = hb.DistAmg(h, c, prm, A)
        b = A.vector(1.0)
        x = A.vector(0.0)
        its, rel, norms = hb.dist_pcg(h, c, A, amg, b, x, tol=1e-8, max_iter=100)
        return dict(its=its, rel=rel, norms=norms, A=(A.info["first_row"],) + A.download())
    res = run_ranks(nranks, fn)
    gi, gj, ga = gather([r["A"] for r in res])
    A = hb.ParCsr.from_host(handle, gi, gj, ga)
    amg = hb.Amg(handle, RelaxType=13, RelaxTypeUp=14, GSBlocks=T, CoarsenType=8)
    amg.setup(A)
    n = gi.size - 1
    b = handle.zeros(n); handle.fill(b, 1.0)
    x = handle.zeros(n)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert all(r["rel"] < 1e-8 for r in res)
    if nranks == 1:
        assert res[0]["its"] == its
        assert np.max(np.abs(res[0]["norms"] - norms)) / norms[0] < 1e-12
    else:
        assert abs(res[0]["its"] - its) <= 3, (res[0]["its"], its)
    amg.destroy(); A.destroy()


@pytest.mark.parametrize("nranks,dims,rlx", [(2, (12, 11, 10), 8), (3, (9, 8, 13), 8), (2, (12, 11, 10), -1), (4, (10, 9, 13), 13)])
def test_nrank_hybrid_gs_pcg_history_equals_the_restatement(nranks, dims, rlx):
    """BoomerAMG-PCG across z-slab ranks with ONE Gauss-Seidel block per rank (the reference's mpirun -np N, one thread):
    symmetric l1-GS 8 / forward l1-GS 13 / the default 13-14.  The first sweep of every cycle starts from a zero iterate, so both
    halves of a symmetric sweep must see zero ghosts (par_relax.c builds Vext once per call) -- the ghost tail still holds
    the previous cycle's halo.  Oracle: oracle/amg_oracle.c -P 1 1 N (rank-shaped blocks on every level, pinned against the
    reference's own np = 8 record in tests/test_oracle.py); iteration count equal, residual history to 1e-10."""
    import subprocess
    import tempfile
    import os
    import hypre_ve_b200 as hb
    oracle = os.path.join(refio.ROOT, "oracle", "_build", "amg_oracle")
    if not os.path.exists(oracle):
        subprocess.run(["make", "-s", "-C", os.path.join(refio.ROOT, "oracle")], check=True)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.bin")
        args = ["-n", *dims, "-P", 1, 1, nranks, "-pmis", "-solver", 1, "-o", path] + (["-rlx", rlx] if rlx >= 0 else [])
        subprocess.run([oracle] + [str(a) for a in args], check=True, capture_output=True)
        d = refio.read_dump(path)
    prm = dict(RelaxType=rlx) if rlx >= 0 else dict(RelaxType=13, RelaxTypeUp=14)

    def fn(r, h, c):
        A = hb.DistMatrix.laplacian(h, c, *dims, 1, 1, nranks, 7)
        amg = hb.DistAmg(h, c, hb.Amg(h, GSBlocks=1, ModuleRAP2=0, **prm), A)
        b = A.vector(1.0)
        x = A.vector(0.0)
        its, rel, norms = hb.dist_pcg(h, c, A, amg, b, x, tol=1e-8, max_iter=100)
        return dict(its=its, norms=norms, levels=amg.num_levels)
    res = run_ranks(nranks, fn)
    assert res[0]["levels"] == int(d["hdr"][3])
    assert res[0]["its"] == int(d["hdr"][4]), (res[0]["its"], int(d["hdr"][4]))
    assert np.max(np.abs(res[0]["norms"] - d["norms"])) / d["norms"][0] < 1e-10


def test_a_failing_rank_fails_the_job_instead_of_hanging():
    """a rank that raises before an exchange aborts the group (b200_comm_group_abort): its peers return an error from the
    barrier instead of waiting for ever, and every rank's handle is closed"""
    import hypre_ve_b200 as hb

    def fn(r, h, c):
        if r == 1:
            raise RuntimeError("rank 1 gives up")
        A = hb.DistMatrix.laplacian(h, c, 8, 8, 8, 1, 1, 3, 7)      # collective: blocks until every rank arrives
        return A.info["local_rows"]
    with pytest.raises(RuntimeError, match="rank 1 gives up"):
        run_ranks(3, fn)


def test_direct_peer_to_peer_halos_and_reductions_between_rank_threads():
    """The NVLink data path of the multi-GPU solve (halo pack-and-push into the neighbour's receive buffer + flags,
    device-to-device rank-ordered reductions: b200_halo_forward_f64, b200_comm_allreduce_sum_dev2dev) run between the rank
    threads of ONE GPU: a subset of this file is repeated in a child process with B200_P2P_THREADS=1.  The kernels of one rank
    spin on flags raised by kernels on other streams of the same device, hence eager module loading and one hardware queue
    per stream in the child's environment."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, B200_P2P_THREADS="1", B200_P2P_TIMEOUT_S="20", CUDA_MODULE_LOADING="EAGER", CUDA_DEVICE_MAX_CONNECTIONS="32")
    sel = ("test_zslab_partition_equals_reference_cpu_build or test_dist_matvec_matches_gathered or "
           "test_nrank_replicated_coarse_levels_equal_single_gpu or "
           "test_nrank_hierarchy_equals_single_gpu or test_nrank_gmres_and_bicgstab_equal_single_gpu or "
           "test_nrank_hybrid_gs_pcg_history_equals_the_restatement or test_dist_amg_pcg_with_hybrid_gauss_seidel")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider", "-k", sel],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=280)
    assert p.returncode == 0, p.stdout[-3000:]
    assert " passed" in p.stdout and "failed" not in p.stdout, p.stdout[-1000:]


def test_launch_helper_agrees_with_the_library_partition():
    """hypre_ve_b200.launch (host arithmetic, covered on CPU by tests/test_launch_gloo.py) describes the same
    row partition the library generates"""
    import hypre_ve_b200 as hb
    from hypre_ve_b200 import launch
    dims, grid, nranks = (13, 12, 9), (2, 2, 1), 4

    def fn(r, h, c):
        A = hb.DistMatrix.laplacian(h, c, *dims, *grid, 7)
        inf = A.info
        return inf["local_rows"], inf["first_row"]
    res = run_ranks(nranks, fn)
    for r, (rows, first) in enumerate(res):
        box, f0 = launch.local_box(r, dims, grid)
        assert rows == box[0] * box[1] * box[2] and first == f0
