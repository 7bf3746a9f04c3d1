"""GPU parity tests of the device-side IJ assembly (b200_ij_*): the call stream of
`ref_dump -ijbuild MODE` -- SetValues / AddToValues in scrambled row order, rotated entry order, rows split over calls,
in-call duplicates, a row listed twice in one call, updates after assembly -- is fed to the device assembler and the
result compared bit for bit with the matrix the reference's own HYPRE_IJMatrix interface assembled from the same calls
(oracle/_ref/ref_dump), then with the BoomerAMG hierarchy and PCG history built on it."""
import numpy as np
import pytest

import ijstream
import refio

pytestmark = pytest.mark.gpu


def feed(asm, calls):
    rejected = 0
    for call in calls:
        add, rows, ncols, cols, vals = ijstream.flat(call)
        rejected += asm.set_values(ncols, rows, cols, vals, add=bool(add))
    return rejected


@pytest.mark.parametrize("args,mode", [
    (["-n", 6, 5, 4], 1), (["-n", 6, 5, 4], 2), (["-n", 5, 4, 3, "-27pt"], 2), (["-n", 9, 1, 7], 1), (["-n", 1, 1, 13], 2),
    (["-n", 20, 18, 16], 2), (["-n", 12, 11, 10, "-27pt"], 1), (["-n", 14, 13, 12, "-difconv", "-a", 3, -2, 1, "-atype", 3], 2),
])
def test_device_assembly_equals_the_reference_ij_interface(handle, args, mode):
    import hypre_ve_b200 as hb
    g, _ = refio.run_ref(args + ["-noamg"])
    I, J, a, _ = refio.csr(g, "A", 0)
    d, _ = refio.run_ref(args + ["-ijbuild", mode, "-noamg"])
    ri, rj, ra, _ = refio.csr(d, "A", 0)
    N = I.size - 1
    asm = hb.IJAssembler(handle, 0, N - 1)
    assert feed(asm, ijstream.calls_before_assembly(I, J, a, mode)) == 0
    A, miss = asm.assemble()
    assert miss == 0
    ei, ej, ea = A.diag.download()
    pi, pj, pa = ijstream.assemble(ijstream.replay(ijstream.calls_before_assembly(I, J, a, mode), N))
    assert np.array_equal(ei, pi) and np.array_equal(ej, pj) and np.array_equal(ea, pa)      # first assembly
    assert feed(asm, ijstream.calls_after_assembly(ei, ej, ea)) == 0
    A2, miss = asm.assemble()
    assert miss == 0 and A2.p.value == A.p.value                                              # updated in place
    ei, ej, ea = A.diag.download()
    assert np.array_equal(ei, ri) and np.array_equal(ej, rj) and np.array_equal(ea, ra)      # = the reference's matrix
    asm.destroy(); A.destroy()


@pytest.mark.parametrize("args,params", [
    (["-n", 16, 14, 12, "-rlx", 18], dict(RelaxType=18)),
    (["-n", 10, 10, 10, "-27pt"], dict(RelaxType=13, RelaxTypeUp=14)),
    (["-n", 14, 13, 12, "-difconv", "-rlx", 18, "-agg_nl", 1], dict(RelaxType=18, AggNumLevels=1)),
])
def test_hierarchy_on_the_device_assembled_operator(handle, args, params):
    """the entry order the assembly leaves (diagonal first, the rest in insertion order -- NOT the generator's order)
    feeds strength / ext+i / truncation: every level and the PCG history equal the reference run on ITS assembled matrix"""
    import hypre_ve_b200 as hb
    prob = args[:args.index("-rlx")] if "-rlx" in args else args
    g, _ = refio.run_ref(prob + ["-noamg"])
    I, J, a, _ = refio.csr(g, "A", 0)
    d, _ = refio.run_ref(args + ["-ijbuild", 1, "-pmis", "-keepT", 1])
    N = I.size - 1
    asm = hb.IJAssembler(handle, 0, N - 1)
    feed(asm, ijstream.calls_before_assembly(I, J, a, 1))
    A, _ = asm.assemble()
    ei, ej, ea = A.diag.download()
    feed(asm, ijstream.calls_after_assembly(ei, ej, ea))
    asm.assemble()
    assert not np.array_equal(A.diag.download()[1], J)                # a different entry order than the generator's
    amg = hb.Amg(handle, ModuleRAP2=0, **params)
    amg.setup(A)
    nl = int(d["hdr"][3])
    assert amg.num_levels == nl
    for l in range(nl):
        i, j, v = amg.level_A(l).download()
        ri, rj, ra, _ = refio.csr(d, "A", l)
        assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(v, ra), ("A", l)
        if l < nl - 1:
            assert np.array_equal(amg.level_CF(l), d["CF%d" % l]), ("CF", l)
            i, j, v = amg.level_P(l).download()
            pi, pj, pa, _ = refio.csr(d, "P", l)
            assert np.array_equal(i, pi) and np.array_equal(j, pj) and np.array_equal(v, pa), ("P", l)
    b = handle.zeros(N); handle.fill(b, 1.0)
    x = handle.zeros(N)
    its, rel, norms = handle.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    assert its == int(d["hdr"][4])
    assert np.max(np.abs(norms - d["norms"])) / d["norms"][0] < 1e-10
    amg.destroy(); asm.destroy(); A.destroy()


def test_rejected_and_missing_records(handle):
    """rows / columns outside the declared ranges are dropped and counted at SetValues; after assembly a value set on
    an element that does not exist is counted (the reference's ' Error, element %b %b does not exist')"""
    import hypre_ve_b200 as hb
    asm = hb.IJAssembler(handle, 10, 19, 10, 19)                      # a 10 x 10 block of a larger numbering
    assert asm.set_values([2, 2], [12, 25], [12, 13, 25, 26], [4.0, -1.0, 4.0, -1.0]) == 2        # row 25 is not local
    assert asm.set_values([3], [13], [13, 9, 20], [4.0, -1.0, -1.0]) == 2                         # columns 9 and 20 do not exist
    assert asm.set_values([0, 1], [14, 15], [15], [2.0]) == 0                                     # an empty row block
    A, miss = asm.assemble()
    i, j, a = A.diag.download()
    assert np.array_equal(np.diff(i), [0, 0, 2, 1, 0, 1, 0, 0, 0, 0])                             # rows never set stay empty
    assert np.array_equal(j, [2, 3, 3, 5]) and np.array_equal(a, [4.0, -1.0, 4.0, 2.0])           # local column numbers
    assert asm.set_values([2], [12], [12, 17], [1.0, 1.0], add=True) == 0
    _, miss = asm.assemble()
    assert miss == 1                                                                              # (12, 17) was never inserted
    assert np.array_equal(A.diag.download()[2], [5.0, -1.0, 4.0, 2.0])
    asm.destroy(); A.destroy()
    asm = hb.IJAssembler(handle, 0, 4)                                                            # nothing set at all
    A, miss = asm.assemble()
    i, j, a = A.diag.download()
    assert miss == 0 and i.tolist() == [0] * 6 and j.size == 0
    asm.destroy(); A.destroy()


def test_large_stream_through_the_pinned_chunks(handle):
    """14.6 M records (128^3, 7-point) in a few large calls: several pinned chunks stream to the device log, which grows
    geometrically; the assembled operator equals the generator's bit for bit and so does a residual computed with it"""
    import hypre_ve_b200 as hb
    n1 = 128
    G = hb.ParCsr.laplacian(handle, n1, n1, n1)
    I, J, a = G.diag.download()
    N = I.size - 1
    asm = hb.IJAssembler(handle, 0, N - 1)
    step = 300000
    order = list(range(0, N, step))[::-1]                            # row blocks from the last to the first
    rows_all = np.arange(N, dtype=np.int32)
    for r0 in order:
        r1 = min(N, r0 + step)
        s, e = I[r0], I[r1]
        half = np.float64(0.5) * a[s:e]
        assert asm.set_values(np.diff(I[r0:r1 + 1]), rows_all[r0:r1], J[s:e], half) == 0
        assert asm.set_values(np.diff(I[r0:r1 + 1]), rows_all[r0:r1], J[s:e], half, add=True) == 0
    A, miss = asm.assemble()
    i, j, v = A.diag.download()
    assert miss == 0 and np.array_equal(i, I) and np.array_equal(j, J) and np.array_equal(v, a)
    x = handle.array(np.linspace(0.0, 1.0, N))
    y1, y2 = handle.zeros(N), handle.zeros(N)
    A.matvec(1.0, x, 0.0, None, y1)
    G.matvec(1.0, x, 0.0, None, y2)
    assert np.array_equal(y1.numpy(), y2.numpy())
    asm.destroy(); A.destroy(); G.destroy()
