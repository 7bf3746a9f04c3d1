"""Config-4-style run: 3-D anisotropic diffusion (-c 1 1 0.001) with one aggressive level, and the isotropic
config-2 operator with -agg_nl 1, at N^3 on one GPU.  usage: agg_bench.py [N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = hb.Handle(0)
for name, c, agg in (("7-pt isotropic", (1.0, 1.0, 1.0), 0), ("7-pt isotropic", (1.0, 1.0, 1.0), 1),
                     ("anisotropic c=(1,1,0.001)", (1.0, 1.0, 0.001), 0), ("anisotropic c=(1,1,0.001)", (1.0, 1.0, 0.001), 1)):
    A = hb.ParCsr.laplacian(h, n, n, n, c=c)
    N = A.local[0]
    b = h.zeros(N); h.fill(b, 1.0)
    for rep in range(2):
        amg = hb.Amg(h, AggNumLevels=agg)
        h.timer_start(); amg.setup(A); set_ms = h.timer_stop_ms()
        x = h.zeros(N)
        h.timer_start()
        its, rel, norms = h.pcg(A, amg, b, x, tol=1e-8, max_iter=200)
        sol_ms = h.timer_stop_ms()
        if rep:
            sizes = [amg.level_A(l).dims for l in range(amg.num_levels)]
            opc = sum(s[2] for s in sizes) / sizes[0][2]
            print("%s %d^3 agg_nl %d: levels %d setup %.1f ms solve %.1f ms its %d rel %.2e op.complexity %.3f rows %s phases %s" %
                  (name, n, agg, amg.num_levels, set_ms, sol_ms, its, rel, opc, [s[0] for s in sizes], [round(t, 1) for t in amg.setup_times()]))
        amg.destroy(); x.free()
    A.destroy()
