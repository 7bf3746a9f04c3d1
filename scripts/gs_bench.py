"""Time the Gauss-Seidel path: level-schedule build, forward / backward sweeps, and BoomerAMG-PCG with the
library-default 13/14 smoother for several Gauss-Seidel block counts.  usage: gs_bench.py N [7|27] [T ...]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import hypre_ve_b200 as hb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
st = int(sys.argv[2]) if len(sys.argv) > 2 else 7
Ts = [int(t) for t in sys.argv[3:]] or [1184, 4096]
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n, n, n) if st == 7 else hb.ParCsr.laplacian27(h, n, n, n)
N, nnz = A.local[0], A.local[1]
D = A.diag
f = h.zeros(N); h.fill(f, 1.0)
byt = 12.0 * nnz + 36.0 * N
for T in Ts:
    u = h.zeros(N)
    l1 = h.l1_norms(D, 4, T)
    h.sync(); t0 = time.time()
    h.relax_gs(D, 13, f, l1, u, T)          # builds the plan
    h.sync(); t_plan = time.time() - t0
    out = []
    for typ in (13, 14):
        h.relax_gs(D, typ, f, l1, u, T)
        h.timer_start()
        for _ in range(10):
            h.relax_gs(D, typ, f, l1, u, T)
        ms = h.timer_stop_ms() / 10
        out.append("relax %d %.3f ms %.0f GB/s" % (typ, ms, byt / ms / 1e6))
    print("T=%d: %s | plan+first sweep %.1f ms | algorithmic bytes %.3f GB" % (T, " ; ".join(out), t_plan * 1e3, byt / 1e9))
    u.free(); l1.free()
    amg = hb.Amg(h, GSBlocks=T, RelaxType=13, RelaxTypeUp=14)
    amg.setup(A); amg.destroy()
    amg = hb.Amg(h, GSBlocks=T, RelaxType=13, RelaxTypeUp=14)
    h.timer_start(); amg.setup(A); set_ms = h.timer_stop_ms()
    x = h.zeros(N)
    h.timer_start()
    its, rel, norms = h.pcg(A, amg, f, x, tol=1e-8, max_iter=200)
    sol_ms = h.timer_stop_ms()
    print("   AMG-PCG 13/14: levels %d setup %.1f ms solve %.1f ms iterations %d rel %.3e phases %s" %
          (amg.num_levels, set_ms, sol_ms, its, rel, [round(t, 1) for t in amg.setup_times()]))
    amg.destroy(); x.free()
