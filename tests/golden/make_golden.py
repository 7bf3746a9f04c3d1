#!/usr/bin/env python3
"""Regenerate the golden fixtures from the reference's own CPU build (oracle/_ref/ref_dump,
OMP_NUM_THREADS=1).  Run in the container that has /root/reference after
`python oracle/build_ref.py`.  Each .bin is a ref_dump stream (see tests/refio.py)."""
import os
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
if __name__ == "__main__":
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for name, args in CASES.items():
        out = subprocess.run([REF] + args + COMMON + ["-o", os.path.join(HERE, name)], env=env, check=True,
                             capture_output=True, text=True).stdout
        print(name, out.splitlines()[1])
