#!/usr/bin/env python3
"""Compile the reference's own CPU implementation (hypre 2.20.0, SX-Aurora fork)
from the sources where they lie under /root/reference into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (hypre_ve_b200/) imports, links
or executes anything produced here; tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs use it as the checker and the CPU baseline.

What it does (no reference build system is run, no reference source is copied
into the repository):
  * reads the FILES / CUFILES lists of each reference sub-directory Makefile to know
    which .c files make up the library (reference: src/<dir>/Makefile);
  * compiles them with gcc -O2 -fopenmp (sequential MPI stubs: HYPRE_SEQUENTIAL,
    see oracle/ref_config/HYPRE_config.h) into oracle/_ref/obj/, in parallel;
  * three reference files do not compile off the NEC Vector Engine; a patched
    *temporary* copy of each is written to oracle/_ref/patched/ (git-ignored) and
    compiled from there:
      - src/parcsr_ls/par_relax.c: cases 3 (:354-1582) and 6 (:2266-3461) use
        VE-only types/fields (sblas_int_t, asl_sort_t, ms_* ...) -> replaced by
        an error stub (those relax types are not used by any oracle run);
      - src/distributed_ls/ParaSails/Matrix.c:73-74,:169-170: VE-only struct fields;
      - src/parcsr_mv/par_csr_matrix.c:933: `I` -> `II` typo (off-path, ReadIJ);
  * links oracle/_ref/libHYPRE_ref.so, the reference driver oracle/_ref/ij
    (src/test/ij.c, unmodified) and oracle/_ref/ref_dump (oracle/ref_dump.c, our
    hierarchy dumper that calls the reference's public API).

Usage: python oracle/build_ref.py [--ref /root/reference] [-j N]
"""
import argparse
import concurrent.futures as cf
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

LIB_DIRS = [
    "utilities", "blas", "lapack", "seq_mv", "parcsr_mv", "parcsr_block_mv",
    "distributed_matrix", "matrix_matrix", "IJ_mv", "multivector", "krylov",
    "distributed_ls/Euclid", "distributed_ls/ParaSails", "distributed_ls/pilut",
    "parcsr_ls",
]


def makefile_list(path, var):
    """Return the file names assigned to `var = \\ ...` in a reference Makefile."""
    txt = open(path).read()
    m = re.search(r"^%s\s*=\s*\\?\n((?:.*\\\n)*.*\n)" % re.escape(var), txt, re.M)
    if not m:
        return []
    body = m.group(1)
    names = []
    for line in body.splitlines():
        line = line.strip()
        if not line:
            break
        cont = line.endswith("\\")
        line = line.rstrip("\\").strip()
        names += [w for w in line.split() if w.endswith(".c") or w.endswith(".cxx")]
        if not cont:
            break
    return names


def patched_copy(ref_src, rel):
    """Write a patched temporary copy of reference file `rel`; return its path."""
    src = os.path.join(ref_src, rel)
    lines = open(src).read().split("\n")
    if rel == "parcsr_ls/par_relax.c":
        def find_case(tag, start):
            for i in range(start, len(lines)):
                if re.match(r"\s*case %s:" % tag, lines[i]):
                    return i
            raise RuntimeError("case %s not found" % tag)

        def end_of_case(i):
            # the case body ends at the first "  } break;" at indentation 2
            for k in range(i + 1, len(lines)):
                if lines[k].rstrip() == "  } break;":
                    return k
            raise RuntimeError("end of case not found")
        c3 = find_case("3", 300)
        e3 = end_of_case(c3)
        stub3 = ['  case 3: { hypre_error_w_msg(HYPRE_ERROR_GENERIC,'
                 '"relax 3 is VE-only in this fork"); relax_error = 1; } break;']
        lines[c3:e3 + 1] = stub3
        c6 = find_case("6", c3 + 1)
        e6 = end_of_case(c6)
        stub6 = ['  case 6: { hypre_error_w_msg(HYPRE_ERROR_GENERIC,'
                 '"relax 6 is VE-only in this fork"); relax_error = 1; } break;']
        lines[c6:e6 + 1] = stub6
    elif rel == "distributed_ls/ParaSails/Matrix.c":
        out = []
        for ln in lines:
            if re.match(r"\s*mat->(flag|t_flag) = 0;", ln):
                out.append("#ifdef __ve__")
                out.append(ln)
                out.append("#endif")
            elif ln.strip() == "#ifndef _FTRACE":
                out.append("#if !defined(_FTRACE) && defined(__ve__)")
            else:
                out.append(ln)
        lines = out
    elif rel == "parcsr_mv/par_csr_matrix.c":
        lines = [ln.replace("(HYPRE_Int)(I-big_base_i-first_row_index)",
                            "(HYPRE_Int)(II-big_base_i-first_row_index)") for ln in lines]
    dst = os.path.join(OUT, "patched", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        f.write("\n".join(lines))
    return dst


def run(cmd):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return p.returncode, p.stdout, cmd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("-j", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--bigint", action="store_true",
                    help="build oracle/_ref/big/{libHYPRE_ref_big.so, ij_big}: the reference's --enable-bigint configuration, needed by "
                         "`ij -n 512 512 512` in one process (local nonzero counts pass 2^31)")
    a = ap.parse_args()
    if a.bigint:
        global OUT
        OUT = os.path.join(OUT, "big")
    ref_src = os.path.join(a.ref, "src")
    if not os.path.isdir(ref_src):
        print("reference not present at %s; keeping prebuilt oracle/_ref" % a.ref)
        return 0
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)

    inc = ["-I" + os.path.join(HERE, "ref_config"), "-I" + ref_src]
    for d in LIB_DIRS + ["distributed_ls", "test", "struct_mv"]:
        inc.append("-I" + os.path.join(ref_src, d))
    cflags = ["-O2", "-fopenmp", "-fPIC", "-w", "-DHAVE_CONFIG_H", "-DHYPRE_VE"] + (["-DB200_REF_BIGINT"] if a.bigint else []) + inc

    patched = {"parcsr_ls/par_relax.c", "distributed_ls/ParaSails/Matrix.c",
               "parcsr_mv/par_csr_matrix.c"}
    jobs = []
    objs = []
    for d in LIB_DIRS:
        mk = os.path.join(ref_src, d, "Makefile")
        files = (makefile_list(mk, "FILES") + makefile_list(mk, "CUFILES") +
                 makefile_list(mk, "BLAS_FILES") + makefile_list(mk, "LAPACK_FILES"))
        if d == "lapack":
            files.append("dlamch.c")   # separate rule in the reference Makefile
        for f in files:
            if f.startswith("F90_") or not f.endswith(".c"):
                continue            # Fortran interface shims: --disable-fortran
            rel = os.path.join(d, f)
            src = os.path.join(ref_src, rel)
            if not os.path.exists(src):
                continue
            obj = os.path.join(OUT, "obj", rel.replace("/", "__")[:-2] + ".o")
            objs.append(obj)
            if os.path.exists(obj) and not a.force and os.path.getmtime(obj) > os.path.getmtime(src):
                continue
            extra = []
            if rel in patched:
                src = patched_copy(ref_src, rel)
                extra = ["-I" + os.path.join(ref_src, d)]
            jobs.append(["gcc"] + cflags + extra + ["-c", src, "-o", obj])
    print("compiling %d reference files (%d up to date)" % (len(jobs), len(objs) - len(jobs)))
    fails = 0
    with cf.ThreadPoolExecutor(a.j) as ex:
        for rc, out, cmd in ex.map(run, jobs):
            if rc:
                fails += 1
                print("FAILED:", cmd[-3], "\n", out[-2000:])
    if fails:
        return 1
    lib = os.path.join(OUT, "libHYPRE_ref_big.so" if a.bigint else "libHYPRE_ref.so")
    rc, out, _ = run(["gcc", "-shared", "-fopenmp", "-o", lib] + objs + ["-lm"])
    if rc:
        print(out[-4000:])
        return 1
    link = ["-L" + OUT, "-lHYPRE_ref_big" if a.bigint else "-lHYPRE_ref", "-Wl,-rpath,$ORIGIN", "-fopenmp", "-lm"]
    if a.bigint:
        rc, out, _ = run(["gcc"] + cflags + ["-DHYPRE_TIMING", os.path.join(ref_src, "test", "ij.c"), "-o", os.path.join(OUT, "ij_big")] + link)
        if rc:
            print(out[-4000:])
            return 1
        print("built", lib, "ij_big")
        return 0
    # the reference's own driver, unmodified
    rc, out, _ = run(["gcc"] + cflags + ["-DHYPRE_TIMING", os.path.join(ref_src, "test", "ij.c"),
                                          "-o", os.path.join(OUT, "ij")] + link)
    if rc:
        print(out[-4000:])
        return 1
    dump = os.path.join(HERE, "ref_dump.c")
    if os.path.exists(dump):
        rc, out, _ = run(["gcc"] + cflags + ["-DHYPRE_TIMING", dump, "-o", os.path.join(OUT, "ref_dump")] + link)
        if rc:
            print(out[-4000:])
            return 1
    # ---- drop-in boundary evidence (tests/test_gpu_refsrc.py) ----------------------------------------------------------------
    # The reference's own programs, UNMODIFIED, compiled against the reference's own headers and linked against
    # libhypre_b200.so FIRST, the reference library after it for everything outside the hot path (first
    # definition in link order wins):   ij_on_b200  = src/test/ij.c,   ex5_on_b200 = src/examples/ex5.c.
    # ex5 expects a real <mpi.h>; include/hypre_compat/mpi.h is a one-rank stand-in, force-included (no source change).
    # ex5_on_b200 is compiled against the SHIM headers include/hypre_compat/HYPRE*.h instead of the reference's.
    root = os.path.dirname(HERE)
    b200_lib_dir = os.path.join(root, "hypre_ve_b200")
    compat = os.path.join(root, "include", "hypre_compat")
    if os.path.exists(os.path.join(b200_lib_dir, "libhypre_b200.so")):
        both = ["-L" + b200_lib_dir, "-lhypre_b200", "-L" + OUT, "-lHYPRE_ref", "-Wl,-rpath,$ORIGIN/../../hypre_ve_b200",
                "-Wl,-rpath,$ORIGIN", "-fopenmp", "-lm"]
        ex5 = os.path.join(ref_src, "examples", "ex5.c")
        stub = ["-include", os.path.join(compat, "mpi.h")]
        for cmd in (
            ["gcc"] + cflags + ["-DHYPRE_TIMING", os.path.join(ref_src, "test", "ij.c"), "-o", os.path.join(OUT, "ij_on_b200")] + both,
            ["gcc"] + cflags + stub + [ex5, "-o", os.path.join(OUT, "ex5")] + link,
            ["gcc", "-O2", "-w"] + stub + ["-I" + compat, ex5, "-o", os.path.join(OUT, "ex5_on_b200")] + both,
        ):
            rc, out, _ = run(cmd)
            if rc:
                print(out[-4000:])
                return 1
    print("built", lib, "ij", "ref_dump" if os.path.exists(dump) else "", "ij_on_b200 ex5 ex5_on_b200")
    return 0


if __name__ == "__main__":
    sys.exit(main())
