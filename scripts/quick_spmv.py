"""Quick device timing of the streaming SpMV (development aid; bench.py is the judged harness)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hypre_ve_b200 as hb

h = hb.Handle(0)
peak = 6454.9
for name, mk in [("7pt-256", lambda: hb.ParCsr.laplacian(h, 256, 256, 256)),
                 ("27pt-160", lambda: hb.ParCsr.laplacian27(h, 160, 160, 160))]:
    A = mk()
    n, nnz, _, _ = A.local
    x = h.zeros(n); h.fill(x, 1.0)
    y = h.zeros(n)
    for _ in range(5):
        A.matvec(1.0, x, 0.0, None, y)
    reps = 50
    h.timer_start()
    for _ in range(reps):
        A.matvec(1.0, x, 0.0, None, y)
    ms = h.timer_stop_ms() / reps
    bytes_ = 12.0 * nnz + 4.0 * (n + 1) + 8.0 * n + 8.0 * n
    print(json.dumps({"case": name, "stages": os.environ.get("B200_SPMV_STAGES"), "ms": round(ms, 4), "GBs": round(bytes_ / ms * 1e-6, 1),
                      "frac": round(bytes_ / ms * 1e-6 / peak, 3)}))
    A.destroy()
