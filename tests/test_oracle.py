"""CPU tests (-m "not gpu"): pin the restatement oracle (oracle/amg_oracle.c) against the golden dumps
produced by the reference's own CPU build, check the reference build (when present) against the
survey's known answers, and check that the C-ABI library loads and exports every declared symbol."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import refio

ROOT = refio.ROOT
ORACLE = os.path.join(ROOT, "oracle", "_build", "amg_oracle")

GOLDEN_ARGS = {
    "lap7_20_pmis_rlx18_modrap.bin": ["-n", "20", "20", "20"],
    "lap7_13x9x11_pmis_rlx18_modrap.bin": ["-n", "13", "9", "11"],
    "lap27_10_pmis_rlx18_modrap.bin": ["-n", "10", "10", "10", "-27pt"],
    "aniso_12_pmis_rlx18_modrap.bin": ["-n", "12", "12", "12", "-c", "1", "1", "0.001"],
}


@pytest.fixture(scope="session")
def oracle_bin():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    assert os.path.exists(ORACLE)
    return ORACLE


def run_oracle(binary, args):
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.bin")
        out = subprocess.run([binary] + [str(a) for a in args] + ["-pmis", "-rlx", "18", "-mod_rap2", "1", "-o", path],
                             check=True, capture_output=True, text=True).stdout
        return refio.read_dump(path), out


@pytest.mark.parametrize("name", sorted(GOLDEN_ARGS))
def test_restatement_matches_reference_golden_bit_for_bit(oracle_bin, name):
    gold = refio.read_dump(os.path.join(refio.GOLDEN, name))
    mine, _ = run_oracle(oracle_bin, GOLDEN_ARGS[name])
    assert set(gold) == set(mine)
    for k in gold:
        assert np.array_equal(gold[k], mine[k]), k      # ints and doubles, hierarchy and residual history


# fixtures of the later features: oracle flags (full list) and whether the run used OpenMP threads (dot products then
# differ in the last bits: hierarchy exact, residual history to 1e-12)
GOLDEN_MORE = {
    "lap7_11_pmis_rlx18_defaultrap.bin": (["-n", 11, 11, 11, "-pmis", "-rlx", 18], True),
    "lap7_12x11x9_agg1_modrap.bin": (["-n", 12, 11, 9, "-pmis", "-rlx", 18, "-mod_rap2", 1, "-agg_nl", 1], True),
    "aniso_11_agg2_defaultrap.bin": (["-n", 11, 11, 11, "-c", 1, 1, 0.001, "-pmis", "-rlx", 18, "-agg_nl", 2], True),
    "lap7_11_gs1314_modrap.bin": (["-n", 11, 11, 11, "-pmis", "-mod_rap2", 1], True),
    "lap7_11_gs8_blocks4_modrap.bin": (["-n", 11, 11, 11, "-pmis", "-rlx", 8, "-mod_rap2", 1, "-gs_blocks", 4], False),
    "lap27_8_rlx7_modrap.bin": (["-n", 8, 8, 8, "-27pt", "-pmis", "-rlx", 7, "-mod_rap2", 1], True),
    # Chebyshev: the extreme Lanczos eigenvalues come from a different (equally stable) tridiagonal solver than the
    # reference's EISPACK tql1 -> last-bit differences in the coefficients, hierarchy exact
    "lap7_11_cheby16_modrap.bin": (["-n", 11, 11, 11, "-pmis", "-rlx", 16, "-mod_rap2", 1], False),
    # other cycle shapes (par_cycle.c level counters): W(2,2) with l1-Jacobi, F-cycle with 13/14 and two coarse sweeps
    "lap7_11_w22_rlx18.bin": (["-n", 11, 11, 11, "-pmis", "-rlx", 18, "-mu", 2, "-ns", 2], True),
    "lap7_11_fmg_gs1314_coarse2.bin": (["-n", 11, 11, 11, "-pmis", "-fmg", "-ns_coarse", 2], True),
    # non-Laplacian SPD operators (ref_dump -perturb): mixed-sign, weak and strong off-diagonals
    "perturbed7_11_rlx18.bin": (["-n", 11, 11, 11, "-perturb", 1, "-pmis", "-rlx", 18], True),
    "perturbed27_8_agg1_gs.bin": (["-n", 8, 8, 8, "-27pt", "-perturb", 7, "-pmis", "-agg_nl", 1], True),
    # nonsymmetric convection-diffusion (GenerateDifConv, par_difconv.c) under AMG-PCG / AMG-GMRES / AMG-BiCGSTAB
    # (krylov/gmres.c, bicgstab.c restated in the oracle)
    "difconv_11_pcg_rlx18.bin": (["-n", 11, 11, 11, "-difconv", "-pmis", "-rlx", 18], True),
    "difconv_11_gmres_rlx18.bin": (["-n", 11, 11, 11, "-difconv", "-pmis", "-rlx", 18, "-solver", 3], True),
    "difconv_13x11x9_upwind_gmres3_agg1.bin": (["-n", 13, 11, 9, "-difconv", "-a", 3, -2, 1, "-atype", 3, "-pmis", "-rlx", 18,
                                                 "-solver", 3, "-k", 3, "-agg_nl", 1], True),
    "difconv_11_bicgstab_gs.bin": (["-n", 11, 11, 11, "-difconv", "-a", 2, 1, 0, "-atype", 1, "-pmis", "-solver", 9], True),
    "lap7_11_gmres_gs1314.bin": (["-n", 11, 11, 11, "-pmis", "-solver", 3], True),
    # GenerateRotate7pt (par_rotate_7pt.c): 2-D rotated anisotropy
    "rotate_24x20_a45_e001_rlx18.bin": (["-n", 24, 20, 1, "-rotate", "-alpha", 45, "-eps", 0.001, "-pmis", "-rlx", 18], True),
    # HMIS (par_coarsen.c:874-1330 first pass with the amg_linklist.c ordering, :2774 HMIS, :2279 PMIS CF_init 1)
    "lap7_12_hmis_default.bin": (["-n", 12, 12, 12, "-hmis"], True),
    "aniso_11_hmis_agg1_rlx18.bin": (["-n", 11, 11, 11, "-c", 1, 1, 0.001, "-rlx", 18, "-agg_nl", 1, "-hmis"], True),
    "rotate_20x20_a30_e01_agg1_gs.bin": (["-n", 20, 20, 1, "-rotate", "-alpha", 30, "-eps", 0.01, "-pmis", "-agg_nl", 1], True),
}


@pytest.mark.parametrize("name", sorted(GOLDEN_MORE))
def test_restatement_matches_golden_of_later_features(oracle_bin, name):
    """default fused Galerkin order, aggressive coarsening + multipass, hybrid Gauss-Seidel (1 and 4 blocks), relax 7"""
    args, exact = GOLDEN_MORE[name]
    gold = refio.read_dump(os.path.join(refio.GOLDEN, name))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.bin")
        subprocess.run([oracle_bin] + [str(a) for a in args] + ["-o", path], check=True, capture_output=True)
        mine = refio.read_dump(path)
    assert set(gold) <= set(mine)          # (the dumper leaves S out on levels whose strength graph it cannot recompute)
    for k in gold:
        if exact or k not in ("norms", "relres", "x"):
            assert np.array_equal(gold[k], mine[k]), k
        else:
            assert gold[k].shape == mine[k].shape and np.max(np.abs(gold[k] - mine[k])) <= 1e-12 * np.max(np.abs(gold[k])), k


def test_golden_known_answers():
    """iteration counts recorded when the fixtures were generated (make_golden.py output)"""
    want = {"lap7_20_pmis_rlx18_modrap.bin": (6, 13), "lap7_13x9x11_pmis_rlx18_modrap.bin": (5, 12),
            "lap27_10_pmis_rlx18_modrap.bin": (4, 11), "aniso_12_pmis_rlx18_modrap.bin": (6, 12)}
    for name, (levels, its) in want.items():
        d = refio.read_dump(os.path.join(refio.GOLDEN, name))
        assert int(d["hdr"][3]) == levels and int(d["hdr"][4]) == its
        assert len(d["norms"]) == its + 1
        # structural invariants of the reference output the GPU path relies on
        for l in range(levels - 1):
            i, j, a, (n, m) = refio.csr(d, "A", l)
            assert np.all(j[i[:-1]] == np.arange(n))             # diagonal stored first in every row
            cf = d["CF%d" % l]
            assert set(np.unique(cf)) <= {-1, 1}
            pi, pj, pa, (pn, pm) = refio.csr(d, "P", l)
            assert pm == int((cf == 1).sum()) and np.all(np.diff(pi) <= 4)   # P_max_elmts = 4


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_reference_build_reproduces_survey_known_answer():
    """ij -n 50 50 50 -solver 1 -pmis -rlx 18 -> 15 its, 4.192356e-09, level table"""
    _, out = refio.run_ref(["-n", 50, 50, 50, "-pmis", "-rlx", 18, "-mod_rap2", 1], dump=False)
    assert "iterations=15 relres=4.192356e-09" in out
    rows = [int(x) for x in re.findall(r"level \d+ rows=(\d+)", out)]
    nnz = [int(x) for x in re.findall(r"level \d+ rows=\d+ nnz=(\d+)", out)]
    assert rows == [125000, 39654, 5595, 674, 117, 22, 3]
    assert nnz == [860000, 1091990, 342947, 43682, 5005, 344, 9]


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args", [["-n", 17, 23, 9], ["-n", 14, 14, 14, "-27pt"], ["-n", 1, 40, 40], ["-n", 30, 30, 30]])
def test_restatement_matches_live_reference(oracle_bin, args):
    gold, _ = refio.run_ref(args + ["-pmis", "-rlx", 18, "-mod_rap2", 1])
    mine, _ = run_oracle(oracle_bin, args)
    for k in gold:
        assert np.array_equal(gold[k], mine[k]), k


def test_library_exports_every_declared_symbol():
    """include/hypre_b200.h <-> libhypre_b200.so <-> ctypes table agree (no compute calls: no GPU here)"""
    import hypre_ve_b200 as hb
    lib = hb.load_library()
    hdr = open(os.path.join(ROOT, "include", "hypre_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
        assert name in hb.SIGNATURES, "symbol %s has no ctypes signature" % name
    assert set(hb.SIGNATURES) <= declared


def test_public_api_header_symbols_are_exported_and_client_links():
    """include/HYPRE_b200.h (boundary B1): every declared HYPRE_* / Generate* function is exported by the
    library, and the plain-C client examples/ij_b200.c compiles and links against it with gcc"""
    import ctypes
    import hypre_ve_b200 as hb
    from hypre_ve_b200 import build as b
    lib = ctypes.CDLL(hb.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "HYPRE_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:HYPRE_|Generate)[A-Za-z0-9_]+)\s*\(", hdr)) - {"HYPRE_Int", "HYPRE_ParCSRMatrix"}
    assert len(declared) > 150
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol %s declared in HYPRE_b200.h but not exported" % name
    exe = b.build_examples()
    assert os.access(exe, os.X_OK)


def test_public_api_fails_loudly_without_device():
    import torch
    from hypre_ve_b200 import build as b
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([b.build_examples(), "-n", "4", "4", "4"], capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


def test_no_cpu_fallback_without_device():
    import hypre_ve_b200 as hb
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hb.B200Error, match="no CUDA device"):
        hb.Handle(0)


def test_product_does_not_reference_the_oracle():
    """the product tree must never import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "hypre_ve_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".c")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt and "amg_oracle" not in txt and "ref_dump" not in txt, os.path.join(dirpath, f)


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,mode", [(["-n", 6, 5, 4], 1), (["-n", 6, 5, 4], 2), (["-n", 5, 4, 3, "-27pt"], 2), (["-n", 9, 1, 7], 1)])
def test_ij_call_stream_mirror_matches_the_reference_ij_interface(args, mode):
    """tests/ijstream.py (the call stream of `ref_dump -ijbuild` and a restatement of the reference's SetValues /
    AddToValues / Assemble rules) reproduces the matrix the reference's own HYPRE_IJMatrix interface assembles, bit
    for bit -- so the GPU test may feed the same stream to b200_ij_* and compare with the reference's output"""
    import ijstream
    g, _ = refio.run_ref(args + ["-noamg"])
    I, J, a, _ = refio.csr(g, "A", 0)
    d, _ = refio.run_ref(args + ["-ijbuild", mode, "-noamg"])
    ri, rj, ra, _ = refio.csr(d, "A", 0)
    rows = ijstream.replay(ijstream.calls_before_assembly(I, J, a, mode), I.size - 1)
    EI, EJ, Ea = ijstream.assemble(rows)
    assert ijstream.update(EI, EJ, Ea, ijstream.calls_after_assembly(EI, EJ, Ea)) == 0
    assert np.array_equal(EI, ri) and np.array_equal(EJ, rj) and np.array_equal(Ea, ra)
    if mode == 2:
        assert ri[-1] > I[-1] and any(len(set(rj[ri[r]:ri[r + 1]])) < ri[r + 1] - ri[r] for r in range(3, I.size - 1, 7))   # duplicates survive


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args", [["-n", 17, 23, 9], ["-n", 12, 12, 12, "-27pt"], ["-n", 15, 14, 13, "-difconv", "-rlx", 18],
                                  ["-n", 30, 26, 1, "-rotate", "-alpha", 60, "-eps", 0.01, "-rlx", 18],
                                  ["-n", 14, 13, 12, "-perturb", 5, "-rlx", 18, "-agg_nl", 1], ["-n", 1, 30, 30], ["-n", 3, 3, 3]])
def test_hmis_restatement_matches_live_reference(oracle_bin, args):
    """the driver-default coarsening (HMIS): Ruge-Stueben first pass (FIFO per measure, largest measure first) followed by
    PMIS seeded with its C points, aggressive levels included -- every dumped array equal to the reference build's"""
    gold, _ = refio.run_ref(args)                       # no -pmis: coarsen_type 10
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.bin")
        subprocess.run([oracle_bin] + [str(a) for a in args] + ["-hmis", "-o", path], check=True, capture_output=True)
        mine = refio.read_dump(path)
    for k in gold:
        assert np.array_equal(gold[k], mine[k]), k


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built")
def test_config_1_literal_driver_default_known_answer(oracle_bin):
    """BASELINE.json configs[0] as written, `ij -laplacian -n 50 50 50 -solver 1` (HMIS, 13/14): SURVEY.md 8c gives 8
    iterations, 7.138942e-10, 8 levels -- reproduced by the reference build and, bit for bit, by the restatement"""
    gold, out = refio.run_ref(["-n", 50, 50, 50])
    assert "levels=8 iterations=8 relres=7.138942e-10" in out
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "o.bin")
        subprocess.run([oracle_bin, "-n", "50", "50", "50", "-hmis", "-o", path], check=True, capture_output=True)
        mine = refio.read_dump(path)
    for k in gold:
        assert np.array_equal(gold[k], mine[k]), k


HMIS_CASES = [["-n", 12, 12, 12, "-rlx", 18], ["-n", 9, 9, 9, "-27pt", "-rlx", 18], ["-n", 14, 13, 12, "-difconv", "-rlx", 18],
              ["-n", 24, 20, 1, "-rotate", "-alpha", 45, "-eps", 0.001, "-rlx", 18], ["-n", 11, 10, 9, "-perturb", 8, "-rlx", 18, "-th", 0.5],
              ["-n", 1, 1, 9, "-rlx", 18]]


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args", HMIS_CASES)
def test_device_hmis_algorithm_replayed_on_the_host(args, tmp_path):
    """No GPU here, so the two halves of the device HMIS are checked as far as a host can: (1) the text the device runs for
    the sequential first pass (csrc/b200_hmis_body.h) is compiled for the host by tests/host_harness and must give the same
    markers as the Python replay; (2) that replay followed by the replay of the device's PMIS sweeps (tests/hmis_emul.py)
    must give the reference's CF markers on every level."""
    import struct
    import hmis_emul
    exe = os.path.join(str(tmp_path), "ruge_body_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host_harness", "ruge_body_check.cpp")], check=True)
    d, _ = refio.run_ref(args + ["-keepT", 1])                     # no -pmis: the reference coarsens with HMIS
    for l in range(int(d["hdr"][3]) - 1):
        I, J, _, _ = refio.csr(d, "S", l)
        n = I.size - 1
        rs = hmis_emul.ruge(I, J, n)
        fin, fout = os.path.join(str(tmp_path), "s.bin"), os.path.join(str(tmp_path), "cf.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<ii", n, J.size)); f.write(I.astype(np.int32).tobytes()); f.write(J.astype(np.int32).tobytes())
        subprocess.run([exe, fin, fout], check=True)
        out = np.fromfile(fout, dtype=np.int32)
        assert out[0] == 0 and np.array_equal(out[1:], rs), ("first pass", l)
        cf, _ = hmis_emul.pmis_device_style(I, J, n, rs)
        assert np.array_equal(cf, d["CF%d" % l]), ("CF", l)


def test_reference_saved_known_answers_of_its_own_test_suite(oracle_bin):
    """The two single-process jobs of the reference's own regression suite that lie on this path (src/test/TEST_ij):
      default.jobs:11     `ij -pmis -Pmx 0 -rlx 0 -xisone`          -> default.saved:  average convergence factor 0.678738,
                                                                        complexities grid 1.407000 / operator 3.252344 / cycle 6.499062
      coarsening.jobs:62  `ij -n 2 2 2 -agg_nl 1 -mxrs 0.1`         -> coarsening.saved (out.14): 10 iterations, 7.834527e-09
    (every other job needs mpirun).  The first is reproduced by the reference build (oracle/_ref/ij); the second cannot be,
    because no coarsening happens and the single level is smoothed with relax type 6, which this fork only has for the Vector
    Engine (par_relax.c:2266-3461; stubbed in oracle/build_ref.py) -- the restatement, whose type 6 follows the `#if 0`
    original kept in the file (:2685-2753), reproduces the recorded answer digit for digit."""
    out = subprocess.run([oracle_bin, "-n", "2", "2", "2", "-agg_nl", "1", "-mxrs", "0.1", "-hmis", "-solver", "0"], check=True,
                         capture_output=True, text=True).stdout
    assert "levels=1 iterations=10 relres=7.834527e-09" in out
    # default.saved lists the SAME four numbers for np = 1 (-pmis), np = 2 and np = 3 (-P 1 1 N -pmis1): the reference's own record
    # that measures drawn from the global row index make the hierarchy independent of the partition -- the rule the
    # row-partitioned device path follows.  The restatement (no truncation, relax 0, AMG as the solver, b = A*1)
    # prints them digit for digit.
    out = subprocess.run([oracle_bin, "-pmis", "-Pmx", "0", "-rlx", "0", "-xisone", "-solver", "0"], check=True, capture_output=True,
                         text=True).stdout
    assert "conv_factor=0.678738 grid=1.407000 operator=3.252344 cycle=6.499062" in out
    assert "iterations=48 relres=8.350438e-09" in out
    for grid in (["1", "1", "2"], ["1", "1", "3"]):           # default.out.1 / .2: z-slabs keep the lexicographic numbering
        out = subprocess.run([oracle_bin, "-P"] + grid + ["-pmis", "-Pmx", "0", "-rlx", "0", "-xisone", "-solver", "0"], check=True,
                             capture_output=True, text=True).stdout
        assert "conv_factor=0.678738 grid=1.407000 operator=3.252344 cycle=6.499062" in out
    # coarsening.jobs:60 `mpirun -np 8 ./ij -P 2 2 2 -pmis1` -> coarsening.saved (out.13): 14 iterations, 3.301634e-09.
    # The restatement run on the gathered operator in the process grid's numbering (-P 2 2 2: measures by global row, one
    # Gauss-Seidel block and one set of l1 norms per RANK on every level) takes the same 14 iterations; its residual (2.66e-09)
    # is not the recorded one: with eight ranks the reference truncates interpolation rows stored as [own columns | ghost
    # columns], and hypre_qsort2_abs breaks the many ties of a Laplacian by storage order -- that one is NOT pinned.
    out = subprocess.run([oracle_bin, "-P", "2", "2", "2", "-pmis", "-solver", "0"], check=True, capture_output=True, text=True).stdout
    assert "levels=5 iterations=14 " in out
    if refio.have_ref():
        ij = os.path.join(ROOT, "oracle", "_ref", "ij")
        out = subprocess.run([ij, "-pmis", "-Pmx", "0", "-rlx", "0", "-xisone"], check=True, capture_output=True, text=True,
                             env=dict(os.environ, OMP_NUM_THREADS="1")).stdout
        assert "Average Convergence Factor = 0.678738" in out
        assert re.search(r"grid = 1\.407000\s+operator = 3\.252344\s+cycle = 6\.499062", out)


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flags", [["-n", 12, 12, 12, "-pmis", "-rlx", 18], ["-n", 10, 9, 8, "-pmis"], ["-n", 9, 9, 9, "-27pt", "-rlx", 18],
                                   ["-n", 14, 14, 14, "-rlx", 18, "-agg_nl", 1]])
def test_boomeramg_as_solver_restatement_matches_the_reference_driver(oracle_bin, flags):
    """`ij -solver 0` (hypre_BoomerAMGSolve, par_amg_solve.c): iteration count and final relative residual of the restatement
    equal the reference driver's printed lines, PMIS and HMIS hierarchies"""
    extra = [] if "-pmis" in flags else ["-hmis"]
    out = subprocess.run([oracle_bin] + [str(f) for f in flags] + extra + ["-solver", "0"], check=True, capture_output=True, text=True).stdout
    its, rel = re.search(r"iterations=(\d+) relres=(\S+)", out).groups()
    ij = os.path.join(ROOT, "oracle", "_ref", "ij")
    ref = subprocess.run([ij, "-laplacian"] + [str(f) for f in flags] + ["-solver", "0"], check=True, capture_output=True, text=True,
                         env=dict(os.environ, OMP_NUM_THREADS="1")).stdout
    assert int(re.search(r"BoomerAMG Iterations = (\d+)", ref).group(1)) == int(its)
    assert re.search(r"Final Relative Residual Norm = (\S+)", ref).group(1) == rel


@pytest.mark.parametrize("flags,its,levels,opc", [
    (["-pmis"], 10, 7, 2.725555),                                               # default smoother 13/14 on one thread
    (["-pmis", "-rlx", 18, "-agg_nl", 1], 27, 6, 1.306381),
    (["-27pt", "-pmis"], 9, 6, 1.208921),
    (["-27pt", "-pmis", "-rlx", 18], 14, 6, 1.208921),
    (["-c", 1, 1, 0.001, "-pmis", "-agg_nl", 1, "-rlx", 18], 31, 9, 1.426037),
    (["-difconv", "-pmis", "-agg_nl", 1, "-rlx", 18], 32, 6, 1.305040),
    (["-hmis"], 8, 8, 3.153179),                                               # configs[0] as written
])
def test_survey_known_answers_at_50_cubed(oracle_bin, flags, its, levels, opc):
    """SURVEY.md 8c, "known answers generated here" for `ij -n 50 50 50 -solver 1 ...` (np = 1): iteration counts, level counts
    and operator complexities (sum of nnz(A_l) / nnz(A_0), the driver's "operator = " line) reproduced by the restatement"""
    out = subprocess.run([oracle_bin, "-n", "50", "50", "50"] + [str(f) for f in flags], check=True, capture_output=True, text=True).stdout
    assert "levels=%d iterations=%d " % (levels, its) in out, out
    nnz = [int(x) for x in re.findall(r"level \d+ rows=\d+ nnz=(\d+)", out)]
    assert abs(sum(nnz) / nnz[0] - opc) < 5e-7, sum(nnz) / nnz[0]


@pytest.mark.skipif(not refio.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args,kw", [
    (["-n", 7, 6, 5, "-difconv"], dict()),
    (["-n", 7, 6, 5, "-difconv", "-a", 3, -2, 1, "-atype", 3], dict(a=(3, -2, 1), atype=3)),
    (["-n", 7, 6, 5, "-difconv", "-a", 2, 1, 0, "-atype", 1, "-c", 1, 2, 0.5], dict(a=(2, 1, 0), atype=1, c=(1, 2, 0.5))),
    (["-n", 7, 4, 5, "-difconv", "-atype", 2], dict(atype=2)),
])
def test_host_side_difconv_coefficients_equal_the_reference_drivers(args, kw):
    """hypre_ve_b200.difconv_values (what ParCsr.difconv / DistMatrix.difconv hand to the generator kernel) against the seven
    values BuildParDifConv computes (ij.c:8266-8409), read back from the operator the reference generated: bit-equal"""
    import hypre_ve_b200 as hb
    d, _ = refio.run_ref(args + ["-noamg"])
    I, J, a, _ = refio.csr(d, "A", 0)
    nx, ny, nz = args[1:4]
    r = ((nz // 2) * ny + ny // 2) * nx + nx // 2                   # an interior row: centre, z-, y-, x-, x+, y+, z+
    assert I[r + 1] - I[r] == 7
    v = hb.difconv_values(nx, ny, nz, **kw)
    assert list(a[I[r]:I[r + 1]]) == [v[0], v[3], v[2], v[1], v[4], v[5], v[6]]
