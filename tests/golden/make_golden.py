#!/usr/bin/env python3
"""Regenerate the golden fixtures from the reference's own CPU build (oracle/_ref/ref_dump,
OMP_NUM_THREADS=1).  Run in the container that has /root/reference after
`python oracle/build_ref.py`.  Each .bin is a ref_dump stream (see tests/refio.py)."""
import os
import sys
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref", "ref_dump")
COMMON = ["-pmis", "-rlx", "18", "-mod_rap2", "1", "-keepT", "1"]
CASES = {
    "lap7_20_pmis_rlx18_modrap.bin": ["-n", "20", "20", "20"],
    "lap7_13x9x11_pmis_rlx18_modrap.bin": ["-n", "13", "9", "11"],
    "lap27_10_pmis_rlx18_modrap.bin": ["-n", "10", "10", "10", "-27pt"],
    "aniso_12_pmis_rlx18_modrap.bin": ["-n", "12", "12", "12", "-c", "1", "1", "0.001"],
}
# features added after the first fixtures: full flag lists, (threads = Gauss-Seidel blocks)
MORE = {
    "lap7_11_pmis_rlx18_defaultrap.bin": (["-n", "11", "11", "11", "-pmis", "-rlx", "18"], 1),          # fused (R A) P, the driver default
    "lap7_12x11x9_agg1_modrap.bin": (["-n", "12", "11", "9", "-pmis", "-rlx", "18", "-mod_rap2", "1", "-agg_nl", "1"], 1),
    "aniso_11_agg2_defaultrap.bin": (["-n", "11", "11", "11", "-c", "1", "1", "0.001", "-pmis", "-rlx", "18", "-agg_nl", "2"], 1),
    "lap7_11_gs1314_modrap.bin": (["-n", "11", "11", "11", "-pmis", "-mod_rap2", "1"], 1),               # library default 13 down / 14 up
    "lap7_11_gs8_blocks4_modrap.bin": (["-n", "11", "11", "11", "-pmis", "-rlx", "8", "-mod_rap2", "1"], 4),
    "lap7_11_cheby16_modrap.bin": (["-n", "11", "11", "11", "-pmis", "-rlx", "16", "-mod_rap2", "1"], 1),
    "lap27_8_rlx7_modrap.bin": (["-n", "8", "8", "8", "-27pt", "-pmis", "-rlx", "7", "-mod_rap2", "1"], 1),
    "lap7_11_w22_rlx18.bin": (["-n", "11", "11", "11", "-pmis", "-rlx", "18", "-mu", "2", "-ns", "2"], 1),            # W(2,2) cycle
    "lap7_11_fmg_gs1314_coarse2.bin": (["-n", "11", "11", "11", "-pmis", "-fmg", "-ns_coarse", "2"], 1),              # F-cycle, GS
    "perturbed7_11_rlx18.bin": (["-n", "11", "11", "11", "-perturb", "1", "-pmis", "-rlx", "18"], 1),                 # non-Laplacian SPD
    "perturbed27_8_agg1_gs.bin": (["-n", "8", "8", "8", "-27pt", "-perturb", "7", "-pmis", "-agg_nl", "1"], 1),
    # nonsymmetric convection-diffusion (GenerateDifConv) and the Krylov drivers for it: AMG-PCG, AMG-GMRES(5), GMRES(3)
    # restarted on upwind convection, AMG-BiCGSTAB
    "difconv_11_pcg_rlx18.bin": (["-n", "11", "11", "11", "-difconv", "-pmis", "-rlx", "18"], 1),
    "difconv_11_gmres_rlx18.bin": (["-n", "11", "11", "11", "-difconv", "-pmis", "-rlx", "18", "-solver", "3"], 1),
    "difconv_13x11x9_upwind_gmres3_agg1.bin": (["-n", "13", "11", "9", "-difconv", "-a", "3", "-2", "1", "-atype", "3", "-pmis",
                                                 "-rlx", "18", "-solver", "3", "-k", "3", "-agg_nl", "1"], 1),
    "difconv_11_bicgstab_gs.bin": (["-n", "11", "11", "11", "-difconv", "-a", "2", "1", "0", "-atype", "1", "-pmis", "-solver", "9"], 1),
    "lap7_11_gmres_gs1314.bin": (["-n", "11", "11", "11", "-pmis", "-solver", "3"], 1),
    # 2-D rotated anisotropic diffusion (GenerateRotate7pt): positive off-diagonals, strong diagonal coupling
    "rotate_24x20_a45_e001_rlx18.bin": (["-n", "24", "20", "1", "-rotate", "-alpha", "45", "-eps", "0.001", "-pmis", "-rlx", "18"], 1),
    # the literal driver default: HMIS coarsening (coarsen_type 10 = Ruge-Stueben first pass + PMIS), 13/14 smoothing
    "lap7_12_hmis_default.bin": (["-n", "12", "12", "12"], 1),
    "aniso_11_hmis_agg1_rlx18.bin": (["-n", "11", "11", "11", "-c", "1", "1", "0.001", "-rlx", "18", "-agg_nl", "1"], 1),
    "rotate_20x20_a30_e01_agg1_gs.bin": (["-n", "20", "20", "1", "-rotate", "-alpha", "30", "-eps", "0.01", "-pmis", "-agg_nl", "1"], 1),
}
if __name__ == "__main__":
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for name, args in CASES.items():
        if sys.argv[1:] and name not in sys.argv[1:]:
            continue
        out = subprocess.run([REF] + args + COMMON + ["-o", os.path.join(HERE, name)], env=env, check=True,
                             capture_output=True, text=True).stdout
        print(name, out.splitlines()[1])
    only = sys.argv[1:]                  # optional: regenerate just the named fixtures
    for name, (args, threads) in MORE.items():
        if only and name not in only:
            continue
        out = subprocess.run([REF] + args + ["-keepT", "1", "-o", os.path.join(HERE, name)], check=True, capture_output=True,
                             text=True, env=dict(os.environ, OMP_NUM_THREADS=str(threads))).stdout
        print(name, out.splitlines()[1])
