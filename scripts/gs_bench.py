"""Time the Gauss-Seidel path: level-schedule build, one forward / backward sweep, and BoomerAMG-PCG with
the library-default 13/14 smoother.  usage: gs_bench.py N [27]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import hypre_ve_b200 as hb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
st = int(sys.argv[2]) if len(sys.argv) > 2 else 7
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n, n, n) if st == 7 else hb.ParCsr.laplacian27(h, n, n, n)
N, nnz = A.local[0], A.local[1]
D = A.diag
f = h.zeros(N); h.fill(f, 1.0)
u = h.zeros(N)
l1 = h.l1_norms(D, 4)
h.sync(); t0 = time.time()
h.relax_gs(D, 13, f, l1, u)          # builds the plan
h.sync(); t_plan = time.time() - t0
for typ in (13, 14):
    h.relax_gs(D, typ, f, l1, u)
    h.timer_start()
    for _ in range(10):
        h.relax_gs(D, typ, f, l1, u)
    ms = h.timer_stop_ms() / 10
    byt = 12.0 * nnz + 36.0 * N
    print("relax %d: %.3f ms/sweep  %.1f GB/s algorithmic (12 nnz + 36 N = %.3f GB)" % (typ, ms, byt / ms / 1e6, byt / 1e9))
print("plan build + first sweep: %.1f ms" % (t_plan * 1e3))
for rl in ((13, 14), (18, 18), (8, 8)):
    amg = hb.Amg(h)
    amg.set("RelaxType", rl[0]); amg.set("RelaxTypeUp", rl[1])
    amg.setup(A); amg.destroy()
    amg = hb.Amg(h)
    amg.set("RelaxType", rl[0]); amg.set("RelaxTypeUp", rl[1])
    h.timer_start(); amg.setup(A); set_ms = h.timer_stop_ms()
    x = h.zeros(N)
    h.timer_start()
    its, rel, norms = h.pcg(A, amg, f, x, tol=1e-8, max_iter=200)
    sol_ms = h.timer_stop_ms()
    print("relax %s: levels %d setup %.1f ms solve %.1f ms iterations %d rel %.3e phases %s" %
          (rl, amg.num_levels, set_ms, sol_ms, its, rel, [round(t, 1) for t in amg.setup_times()]))
    amg.destroy()
