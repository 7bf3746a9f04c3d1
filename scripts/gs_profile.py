"""One sweep of every Gauss-Seidel scheduler at N^3 (for ncu): global soft-barrier (T=1), CTA per block (T=4096),
thread per block over CSR (T=65536), thread per block over the sliced copy (T=262144)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n, n, n)
N = A.local[0]
D = A.diag
f = h.zeros(N); h.fill(f, 1.0)
for T in (1, 4096, 65536, 262144):
    u = h.zeros(N)
    l1 = h.l1_norms(D, 4, T)
    h.relax_gs(D, 13, f, l1, u, T)
    h.relax_gs(D, 14, f, l1, u, T)
    h.sync()
    print("T", T, "ok")
    u.free(); l1.free()
