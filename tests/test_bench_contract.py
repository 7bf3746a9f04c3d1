"""CPU-side checks of bench.py's contract: the reference arm (the unmodified reference driver on the host cores) prints ONE JSON
line with the keys the driver reads, ranks other than 0 stay silent under a multi-rank launch, and the algorithmic-byte formulas
of the roofline lines add up on a hand-made hierarchy."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_IJ = os.path.join(ROOT, "oracle", "_ref", "ij")


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.skipif(not os.path.exists(REF_IJ), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--edge", "20", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "boomeramg_pcg_setup_plus_solve_seconds" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["value"] - (d["setup_s"] + d["solve_s"])) < 1e-9
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and "20x20x20" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["iterations"] > 0


def test_reference_arm_is_silent_on_the_other_ranks():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert p.returncode == 0 and p.stdout.strip() == "", (p.stdout, p.stderr[-1000:])


def test_algorithmic_byte_formulas():
    b = load_bench()
    # two levels: N0 = 1000 rows, 7000 entries, P with 2300 entries; coarse level 300 rows, 8700 entries
    levels = [(1000, 7000, 2300), (300, 8700, 0)]
    n, za, zp, nc = 1000.0, 7000.0, 2300.0, 300.0
    cycle = 24 * n + (12 * za + 28 * n) + (12 * zp + 12 * nc + 8 * n) + (12 * zp + 20 * n + 8 * nc) + (12 * za + 36 * n)
    pcg = 12 * za + 20 * n + (16 + 24 + 24 + 24 + 24) * n
    assert b.solve_bytes_per_iteration(levels) == cycle + pcg
    s = b.setup_bytes(levels)
    zs, zc = za - n, 8700.0
    assert s["strength"] == 12 * za + 4 * n + 4 * zs + 4 * n
    assert s["pmis"] == 4 * zs + 4 * n + 20 * n
    assert s["interp"] == 12 * za + 4 * n + 4 * zs + 4 * n + 4 * n + 12 * zp + 4 * n
    assert s["transpose"] == 12 * zp + 4 * n + 12 * zp + 4 * nc
    assert s["rap"] == 12 * (2 * zp + za) + 4 * (nc + 2 * n) + 12 * zc + 4 * nc


def test_clock_sampler_summary_without_samples():
    b = load_bench()
    s = b.ClockSampler(0)            # no device here: NVML is absent or fails, nothing is sampled
    out = s.summary()
    assert out["samples"] == 0 and out["sm_mhz"] is None and out["reasons"] == []


def test_reference_iteration_table_holds_the_bench_grids():
    b = load_bench()
    assert b.reference_iterations((256, 256, 256)) == 22
    for dims in [(256, 256, 512), (256, 512, 512), (512, 512, 512)]:
        assert isinstance(b.reference_iterations(dims), int)
