"""Smoother comparison at N^3 (config 2 operator, driver-default Galerkin order): l1-Jacobi 18, Chebyshev 16, and both with
one aggressive level.  usage: cheby_bench.py [N] [7|27]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
st = int(sys.argv[2]) if len(sys.argv) > 2 else 7
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n, n, n) if st == 7 else hb.ParCsr.laplacian27(h, n, n, n)
N = A.local[0]
b = h.zeros(N); h.fill(b, 1.0)
for rlx, agg in ((18, 0), (16, 0), (18, 1), (16, 1)):
    for rep in range(2):
        amg = hb.Amg(h, RelaxType=rlx, AggNumLevels=agg, ModuleRAP2=0)
        h.timer_start(); amg.setup(A); set_ms = h.timer_stop_ms()
        x = h.zeros(N)
        h.timer_start()
        its, rel, norms = h.pcg(A, amg, b, x, tol=1e-8, max_iter=200)
        sol_ms = h.timer_stop_ms()
        if rep:
            print("%d-pt %d^3 relax %d agg_nl %d: setup %.1f ms solve %.1f ms total %.1f ms its %d rel %.2e" %
                  (st, n, rlx, agg, set_ms, sol_ms, set_ms + sol_ms, its, rel))
        amg.destroy(); x.free()
