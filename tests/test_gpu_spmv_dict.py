"""The dictionary-compressed solve copy (csrc/b200_spmv_dict.cu: one byte per entry instead of the column and / or the value for
stencil-structured operators) must give the SAME BITS as the plain streaming kernel: the products and their order are unchanged.
The mode is read once per process (B200_SPMV_DICT: bit 0 columns, bit 1 values), so every mode runs tests/dict_case.py in a
process of its own; B200_SPMV_DICT_MIN_NNZ=0 lets the small test operators take the path that only large ones take by default."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_mode(mode):
    env = dict(os.environ, B200_SPMV_DICT=str(mode), B200_SPMV_DICT_MIN_NNZ="0", B200_DEBUG_PLAN="1")
    p = subprocess.run([sys.executable, os.path.join(HERE, "dict_case.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       env=env, timeout=280)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = {l.split()[0]: l.split()[1] for l in p.stdout.splitlines() if len(l.split()) == 3}
    built = [l for l in p.stderr.splitlines() if l.startswith("[b200 dict]")]
    return res, built


def test_dictionary_compressed_spmv_is_bit_identical_to_the_plain_kernel():
    plain, built0 = run_mode(0)
    assert not built0 and len(plain) == 6
    for mode in (1, 2, 3):
        got, built = run_mode(mode)
        assert built, "no dictionary was built in mode %d" % mode
        assert got == plain, (mode, got, plain)
    # what compresses: the Laplacians both ways (2 bytes per entry), random values on a stencil pattern columns only (9);
    # the banded operator with 300 offsets is not a one-lane-per-row operator and keeps the plain path
    _, built = run_mode(3)
    per_entry = sorted(int(l.split("->")[1].split()[0]) for l in built)
    assert 2 in per_entry and 9 in per_entry, built
