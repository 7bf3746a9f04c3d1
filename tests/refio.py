"""Reader for oracle/ref_dump.c streams and runner for the reference dumper (test infrastructure)."""
import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    o = 0
    while o < len(buf):
        (nl,) = struct.unpack_from("<I", buf, o); o += 4
        name = buf[o:o + nl].decode(); o += nl
        dt, cnt = struct.unpack_from("<IQ", buf, o); o += 12
        dtype = np.float64 if dt else np.int32
        nbytes = cnt * (8 if dt else 4)
        out[name] = np.frombuffer(buf, dtype=dtype, count=cnt, offset=o).copy()
        o += nbytes
    return out


def have_ref():
    return os.path.exists(REF_DUMP)


def run_ref(args, threads=1, dump=True):
    """Run the reference dumper; returns (dict of arrays or None, stdout)."""
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "d.bin")
        cmd = [REF_DUMP] + [str(a) for a in args] + (["-o", path] if dump else ["-nodump"])
        p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
        return (read_dump(path) if dump else None), p.stdout


def csr(d, pre, l):
    n, m, nnz = d["%s%d.dims" % (pre, l)]
    return d["%s%d.i" % (pre, l)], d["%s%d.j" % (pre, l)], d.get("%s%d.a" % (pre, l)), (int(n), int(m))
