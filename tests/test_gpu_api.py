"""GPU tests of the reference-facing public API (include/HYPRE_b200.h, boundary B1): the plain-C client
examples/ij_b200.c -- the reference driver's call sequence for `ij -laplacian ... -solver {0,1,2}` --
is run against libhypre_b200.so and its result lines are compared with the reference's own `ij`
(oracle/_ref/ij, the unmodified driver over the reference CPU library) on the same flags; where the
reference build is absent the survey's known answers are used."""
import os
import re
import subprocess

import pytest

import refio

ROOT = refio.ROOT
EXE = os.path.join(ROOT, "examples", "ij_b200")
REF_IJ = os.path.join(ROOT, "oracle", "_ref", "ij")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exe():
    from hypre_ve_b200 import build as b
    return b.build_examples()


def run(cmd, env=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    return p.returncode, p.stdout


def result(out, key="Iterations"):
    its = int(re.search(r"^%s = (\d+)" % key, out, re.M).group(1))
    rel = float(re.search(r"Final Relative Residual Norm = (\S+)", out).group(1))
    return its, rel


def ref_result(flags, key="Iterations", threads=1):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    rc, out = run([REF_IJ, "-laplacian"] + flags, env)
    assert rc == 0, out
    return result(out, key)


AMG_PCG = ["-solver", "1", "-pmis", "-rlx", "18", "-mod_rap2", "1"]
CASES = [
    (["-n", "20", "20", "20"], (13, None)),
    (["-n", "50", "50", "50"], (15, 4.192356e-09)),                 # SURVEY.md 8c known answer
    (["-n", "13", "9", "11"], (12, None)),
    (["-27pt", "-n", "14", "14", "14"], None),
    (["-n", "16", "16", "16", "-c", "1", "1", "0.001"], None),
    (["-n", "24", "24", "24", "-Pmx", "6", "-th", "0.5"], None),
    (["-n", "24", "24", "24", "-tr", "0.1"], None),
]


@pytest.mark.parametrize("flags,known", CASES)
def test_amg_pcg_through_the_public_api_matches_reference_driver(exe, flags, known):
    rc, out = run([exe, "-laplacian"] + flags + AMG_PCG)
    assert rc == 0, out
    its, rel = result(out)
    if os.path.exists(REF_IJ):
        rits, rrel = ref_result(flags + AMG_PCG)
        assert its == rits, (its, rits)
        assert abs(rel - rrel) <= 1e-10 * max(rrel, 1e-300) + 2e-15 or abs(rel / rrel - 1) < 1e-6   # printed with 7 digits
    if known:
        assert its == known[0]
        if known[1]:
            assert abs(rel / known[1] - 1) < 1e-6


@pytest.mark.parametrize("flags,known", [
    (["-n", "50", "50", "50", "-solver", "1", "-pmis", "-mod_rap2", "1"], 10),           # library default 13 down / 14 up
    (["-n", "20", "18", "16", "-solver", "1", "-pmis", "-mod_rap2", "1", "-rlx", "8"], None),
    (["-27pt", "-n", "16", "16", "16", "-solver", "1", "-pmis", "-mod_rap2", "1"], None),
    (["-n", "14", "14", "14", "-solver", "0", "-pmis", "-mod_rap2", "1", "-rlx", "14"], None),
])
def test_hybrid_gauss_seidel_smoothers_through_the_public_api(exe, flags, known):
    rc, out = run([exe, "-laplacian"] + flags)
    assert rc == 0, out
    key = "BoomerAMG Iterations" if "0" == flags[flags.index("-solver") + 1] else "Iterations"
    its, rel = result(out, key)
    if os.path.exists(REF_IJ):
        rits, rrel = ref_result(flags, key)
        assert its == rits and abs(rel / rrel - 1) < 1e-6, (its, rits, rel, rrel)
    if known:
        assert its == known


@pytest.mark.parametrize("flags", [
    ["-n", "20", "20", "20", "-solver", "1", "-pmis", "-rlx", "18", "-ns", "2"],                     # V(2,2)
    ["-n", "20", "20", "20", "-solver", "1", "-pmis", "-rlx", "18", "-mu", "2"],                     # W(1,1)
    ["-n", "22", "18", "20", "-solver", "1", "-pmis", "-rlx", "18", "-mu", "2", "-ns", "2", "-mod_rap2", "1"],
    ["-n", "20", "20", "20", "-solver", "1", "-pmis", "-rlx", "18", "-fmg"],                         # F-cycle
    ["-n", "20", "20", "20", "-solver", "1", "-pmis", "-mu", "2", "-ns", "2"],                       # W(2,2), 13 down / 14 up
    ["-n", "18", "18", "18", "-solver", "1", "-pmis", "-rlx", "16", "-mu", "2", "-ns", "2"],         # Chebyshev
    ["-n", "20", "20", "20", "-solver", "1", "-pmis", "-rlx", "18", "-ns_coarse", "3", "-ns", "2"],
    ["-27pt", "-n", "12", "12", "12", "-solver", "1", "-pmis", "-rlx", "7", "-mu", "2", "-agg_nl", "1"],
    ["-n", "16", "16", "16", "-solver", "0", "-pmis", "-rlx", "18", "-ns_down", "2", "-ns_up", "1"],  # V(2,1), AMG as the solver
    ["-n", "16", "16", "16", "-solver", "0", "-pmis", "-rlx", "18", "-ns_down", "0", "-ns_up", "3"],  # V(0,3)
    ["-n", "16", "16", "16", "-solver", "0", "-pmis", "-mu", "2", "-ns_down", "1", "-ns_up", "2"],
])
def test_cycle_shapes_match_reference_driver(exe, flags):
    """-ns / -ns_down / -ns_up / -ns_coarse (num_grid_sweeps), -mu (cycle type), -fmg (F-cycle): par_cycle.c:180-622"""
    rc, out = run([exe, "-laplacian"] + flags)
    assert rc == 0, out
    key = "BoomerAMG Iterations" if "0" == flags[flags.index("-solver") + 1] else "Iterations"
    its, rel = result(out, key)
    assert os.path.exists(REF_IJ)
    rits, rrel = ref_result(flags, key)
    assert its == rits and abs(rel / rrel - 1) < 1e-6, (its, rits, rel, rrel)


def test_gauss_seidel_blocks_equal_reference_thread_count(exe):
    """-gs_blocks T on our side == OMP_NUM_THREADS=T on the reference's"""
    flags = ["-n", "30", "30", "30", "-solver", "1", "-pmis", "-mod_rap2", "1"]
    rc, out = run([exe, "-laplacian"] + flags + ["-gs_blocks", "24"])
    assert rc == 0, out
    its, rel = result(out)
    if os.path.exists(REF_IJ):
        rits, rrel = ref_result(flags, threads=24)
        assert its == rits and abs(rel / rrel - 1) < 1e-6, (its, rits, rel, rrel)


def test_ij_assembled_operator_equals_generated_operator(exe):
    """HYPRE_IJMatrixSetValues/Assemble (diagonal first, insertion order) gives the generator's matrix:
    identical iteration count and residual through the whole setup + solve."""
    flags = ["-laplacian", "-n", "18", "14", "11"] + AMG_PCG
    rc1, o1 = run([exe] + flags)
    rc2, o2 = run([exe, "-ijbuild"] + flags)
    assert rc1 == 0 and rc2 == 0, o1 + o2
    assert result(o1) == result(o2)
    assert re.findall(r"level .*", o1) == re.findall(r"level .*", o2)
    assert re.search(r"x\[0\] = .*", o1).group(0) == re.search(r"x\[0\] = .*", o2).group(0)


def test_boomeramg_as_solver_matches_reference_driver(exe):
    flags = ["-n", "12", "12", "12", "-solver", "0", "-pmis", "-rlx", "18", "-mod_rap2", "1"]
    rc, out = run([exe, "-laplacian"] + flags)
    assert rc == 0, out
    its, rel = result(out, "BoomerAMG Iterations")
    if os.path.exists(REF_IJ):
        rits, rrel = ref_result(flags, "BoomerAMG Iterations")
    else:
        rits, rrel = 25, 7.395815e-09
    assert its == rits and abs(rel / rrel - 1) < 1e-6


def test_diagonal_scaled_pcg_matches_reference_driver(exe):
    flags = ["-n", "12", "12", "12", "-solver", "2"]
    rc, out = run([exe, "-laplacian"] + flags)
    assert rc == 0, out
    its, rel = result(out)
    if os.path.exists(REF_IJ):
        rits, rrel = ref_result(flags)
    else:
        rits, rrel = 28, 9.892025e-09
    assert its == rits and abs(rel / rrel - 1) < 1e-6


def _stats_block(out):
    """the setup statistics the reference prints at print_level 1: both tables and the complexity lines"""
    a = out.index("Operator Matrix Information:")
    b = out.index("memory = ", a)
    b = out.index("\n", b)
    return [l.rstrip() for l in out[a:b].splitlines()]


@pytest.mark.parametrize("flags", [["-n", "12", "12", "12"], ["-n", "30", "28", "26"], ["-27pt", "-n", "14", "14", "14"],
                                   ["-n", "20", "20", "20", "-agg_nl", "1"], ["-n", "16", "16", "16", "-c", "1", "1", "0.001"]])
def test_setup_statistics_tables_equal_the_reference_drivers(exe, flags):
    """hypre_BoomerAMGSetupStats (par_stats.c): "Operator Matrix Information", "Interpolation Matrix Information" and
    the complexities printed by our Setup are the reference driver's lines, character for character"""
    if not os.path.exists(REF_IJ):
        pytest.skip("reference driver not built")
    f = flags + ["-solver", "1", "-pmis", "-rlx", "18"]
    rc, out = run([exe, "-laplacian"] + f)
    assert rc == 0, out
    rc2, ref = run([REF_IJ, "-laplacian"] + f, dict(os.environ, OMP_NUM_THREADS="1"))
    assert rc2 == 0
    mine, theirs = _stats_block(out), _stats_block(ref)
    assert mine == theirs, "\n".join(mine) + "\n-----\n" + "\n".join(theirs)


def test_driver_default_galerkin_product(exe):
    """no -mod_rap2 flag: the driver default (fused hypre_BoomerAMGBuildCoarseOperatorKT order)"""
    flags = ["-n", "50", "50", "50", "-solver", "1", "-pmis", "-rlx", "18"]
    rc, out = run([exe, "-laplacian"] + flags)
    assert rc == 0, out
    its, rel = result(out)
    assert its == 15 and abs(rel / 4.192356e-09 - 1) < 1e-6          # SURVEY.md 8c known answer
    if os.path.exists(REF_IJ):
        assert (its, ) == ref_result(flags)[:1]


def test_out_of_scope_configuration_is_rejected_loudly(exe):
    """classical modified interpolation (-interptype 0) is not on the B200 path: Setup must fail, not fall back"""
    rc, out = run([exe, "-laplacian", "-n", "8", "8", "8", "-solver", "1", "-pmis", "-interptype", "0", "-rlx", "18", "-mod_rap2", "1"])
    assert rc != 0
    assert "InterpType" in out and "hypre error flag" in out


def test_matvec_loop(exe):
    rc, out = run([exe, "-laplacian", "-n", "9", "9", "9", "-solver", "-1"])
    assert rc == 0, out
    # b = 1: (A b)_i = 6 - (number of neighbours); <Ab, Ab> = sum over the grid
    n = 9
    tot = 0.0
    for z in range(n):
        for y in range(n):
            for x in range(n):
                nb = sum(1 for v in (x, y, z) for d in (-1, 1) if 0 <= v + d < n)
                tot += (6 - nb) ** 2
    got = float(re.search(r"<Ab, Ab> = (\S+)", out).group(1))
    assert got == tot


@pytest.mark.parametrize("flags", [
    ["-difconv", "-n", "7", "6", "5", "-solver", "1", "-pmis", "-rlx", "18"],
    ["-laplacian", "-27pt", "-n", "6", "6", "6", "-solver", "1", "-pmis"],
    ["-rotate", "-n", "12", "11", "-alpha", "30", "-eps", "0.01", "-solver", "1", "-pmis", "-rlx", "18"],
])
def test_ij_file_format_print_and_read(exe, tmp_path, flags):
    """ij -print / -fromfile: HYPRE_IJMatrixPrint / hypre_ParCSRMatrixPrintIJ / HYPRE_IJVectorPrint write the files the
    reference driver writes (byte-identical for the operator, the right-hand side and the initial guess; the solution
    to 1e-9), and HYPRE_IJMatrixRead of the REFERENCE's file gives the same solve as the reference reading it"""
    import numpy as np
    assert os.path.exists(REF_IJ)
    mine, ref = tmp_path / "mine", tmp_path / "ref"
    mine.mkdir(); ref.mkdir()
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p1 = subprocess.run([exe] + flags + ["-print"], cwd=mine, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    p2 = subprocess.run([REF_IJ] + flags + ["-print"], cwd=ref, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p1.returncode == 0 and p2.returncode == 0, p1.stdout + p2.stdout
    for name in ("IJ.out.A.00000", "IJ.out.b.00000", "IJ.out.x0.00000"):
        assert (mine / name).read_bytes() == (ref / name).read_bytes(), name
    xm = np.loadtxt(mine / "IJ.out.x.00000", skiprows=1)
    xr = np.loadtxt(ref / "IJ.out.x.00000", skiprows=1)
    assert np.array_equal(xm[:, 0], xr[:, 0]) and np.max(np.abs(xm[:, 1] - xr[:, 1])) < 1e-9 * np.max(np.abs(xr[:, 1]))
    solver = [f for f in flags if f not in ("-difconv", "-laplacian", "-27pt", "-rotate")]
    solver = solver[solver.index("-solver"):]
    q1 = subprocess.run([exe, "-fromfile", str(ref / "IJ.out.A")] + solver, cwd=mine, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    q2 = subprocess.run([REF_IJ, "-fromfile", str(ref / "IJ.out.A")] + solver, cwd=ref, env=env, stdout=subprocess.PIPE,
                        stderr=subprocess.STDOUT, text=True)
    assert q1.returncode == 0 and q2.returncode == 0, q1.stdout + q2.stdout
    (its, rel), (rits, rrel) = result(q1.stdout), result(q2.stdout)
    assert its == rits and abs(rel / rrel - 1) < 1e-6, (its, rits, rel, rrel)
    assert result(q1.stdout)[0] == result(p1.stdout)[0]            # 14 printed digits are enough to keep the iteration count
