#!/usr/bin/env python3
"""Build libhypre_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Everything is compiled with -fmad=false: the reference CPU build (gcc -O2, x86-64 baseline)
emits no fused multiply-adds, and bit-exact interpolation sparsity needs bit-exact weights
(SURVEY.md 7.3-1).  The solve kernels are HBM-bound, so separate mul/add costs nothing there and
keeps the residual history closer to the reference's.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libhypre_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
          "--expt-relaxed-constexpr", "-I" + os.path.join(HERE, "..", "include")]


def newer(src, obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(SRC, "*.cu")) + glob.glob(os.path.join(SRC, "*.cpp")))
    deps = glob.glob(os.path.join(SRC, "*.h")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s).rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or newer(s, o, deps):
            extra = ["-fmad=false"]
            if verbose:
                extra += ["-Xptxas", "-v"]
            jobs.append([NVCC] + ARCH + COMMON + extra + ["-x", "cu", "-c", s, "-o", o])

    def run(cmd):
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return p.returncode, p.stdout, cmd
    with cf.ThreadPoolExecutor(max(1, os.cpu_count() or 4)) as ex:
        for rc, out, cmd in ex.map(run, jobs):
            if out.strip() and (rc or verbose):
                print(out)
            if rc:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode:
            print(p.stdout)
            raise RuntimeError("link failed")
    return LIB


def build_examples():
    """gcc-compile the plain-C client of the HYPRE_* API (examples/ij_b200.c) against the library."""
    root = os.path.abspath(os.path.join(HERE, ".."))
    src = os.path.join(root, "examples", "ij_b200.c")
    exe = os.path.join(root, "examples", "ij_b200")
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return exe
    cmd = ["gcc", "-O2", "-Wall", "-I" + os.path.join(root, "include"), src, "-o", exe, "-L" + HERE, "-lhypre_b200",
           "-Wl,-rpath,$ORIGIN/../hypre_ve_b200", "-lm"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode:
        print(p.stdout)
        raise RuntimeError("gcc failed on examples/ij_b200.c")
    return exe


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    print(build_examples())
