#!/usr/bin/env python3
"""Record the PCG iteration counts of the reference's own CPU build (oracle/_ref/ij, np = 1) on the grids bench.py runs:
`ij -n nx ny nz -solver 1 -pmis -rlx 18 -keepT 1`.  bench.py prints them as `reference_iterations` beside its own count and
tests/test_gpu_fullsize.py asserts equality.  Usage (where oracle/_ref exists; the big grids need up to 57 GB of host memory):

    python tests/golden/make_reference_iterations.py 256 256 256  256 256 512  256 512 512  512 512 512
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "reference_iterations.json")


def main():
    tab = json.load(open(OUT)) if os.path.exists(OUT) else {"flags": "-solver 1 -pmis -rlx 18 -keepT 1", "counts": {}, "final_rel_res": {}}
    v = sys.argv[1:]
    for k in range(0, len(v), 3):
        dims = v[k:k + 3]
        big = int(dims[0]) * int(dims[1]) * int(dims[2]) > 100_000_000        # 32-bit counters overflow: the --enable-bigint build
        exe = os.path.join(ROOT, "oracle", "_ref", "big", "ij_big") if big else os.path.join(ROOT, "oracle", "_ref", "ij")
        out = subprocess.run([exe, "-n"] + dims + tab["flags"].split(), capture_output=True,
                             text=True, check=True, env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()))).stdout
        tab["counts"][" ".join(dims)] = int(re.search(r"^Iterations = (\d+)", out, re.M).group(1))
        tab["final_rel_res"][" ".join(dims)] = float(re.search(r"Final Relative Residual Norm = (\S+)", out).group(1))
        json.dump(tab, open(OUT, "w"), indent=1, sort_keys=True)
        print(dims, tab["counts"][" ".join(dims)], tab["final_rel_res"][" ".join(dims)])


if __name__ == "__main__":
    main()
