import sys
sys.path.insert(0, ".")
import hypre_ve_b200 as hb
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, 256, 256, 256)
N = A.local[0]
b = h.zeros(N); h.fill(b, 1.0)
for mod in (1, 0, 1, 0):
    amg = hb.Amg(h, ModuleRAP2=mod)
    h.timer_start(); amg.setup(A); ms = h.timer_stop_ms()
    x = h.zeros(N)
    h.timer_start(); its, rel, _ = h.pcg(A, amg, b, x, tol=1e-8, max_iter=100); sol = h.timer_stop_ms()
    print("ModuleRAP2", mod, "setup %.1f ms solve %.1f ms its %d rel %.6e" % (ms, sol, its, rel), [round(t, 1) for t in amg.setup_times()],
          [amg.level_A(l).dims[0] for l in range(amg.num_levels)])
    amg.destroy(); x.free()
