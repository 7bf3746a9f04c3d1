"""y = A0 x on the 256^3 7-pt operator, a few repetitions (command profiled with ncu --set full)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
kind = sys.argv[1] if len(sys.argv) > 1 else "7"
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, 256, 256, 256) if kind == "7" else hb.ParCsr.laplacian27(h, 160, 160, 160)
n = A.local[0]
x = h.zeros(n); h.fill(x, 1.0)
y = h.zeros(n)
for _ in range(10):
    A.matvec(1.0, x, 0.0, None, y)
h.sync()
print("ok")
