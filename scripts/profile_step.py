"""One warm-up step + one step of config 2 (256^3 setup + PCG solve) -- the command profiled under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n1, n1, n1)
n = A.local[0]
b = h.zeros(n); h.fill(b, 1.0)
for rep in range(steps):
    amg = hb.Amg(h, ModuleRAP2=0)          # the driver default, as bench.py runs it
    amg.setup(A)
    x = h.zeros(n)
    its, rel, norms = h.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    print("step", rep, "its", its, "rel", rel, "launches", h.launch_count())
    amg.destroy(); x.free()
