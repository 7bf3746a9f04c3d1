"""Config 2 at full size: 256^3 7-pt, PMIS + ext+i(Pmx 4) + l1-Jacobi PCG on one GPU (dev aid)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hypre_ve_b200 as hb

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n1, n1, n1)
n = A.local[0]
b = h.zeros(n); h.fill(b, 1.0)
for rep in range(3):
    amg = hb.Amg(h)
    h.sync(); t0 = time.time()
    h.timer_start()
    amg.setup(A)
    setup_ms = h.timer_stop_ms()
    x = h.zeros(n)
    h.timer_start()
    its, rel, norms = h.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    solve_ms = h.timer_stop_ms()
    rows = [amg.level_A(l).dims[0] for l in range(amg.num_levels)]
    nnz = [amg.level_A(l).dims[2] for l in range(amg.num_levels)]
    print(json.dumps({"rep": rep, "setup_ms": setup_ms, "solve_ms": solve_ms, "its": its, "rel": rel,
                      "rows": rows, "nnz": nnz, "phases_ms": amg.setup_times(), "norms3": list(norms[:4])}))
    amg.destroy()
