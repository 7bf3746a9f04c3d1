"""Per-level SpMV efficiency on the config-2 hierarchy (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n1, n1, n1)
amg = hb.Amg(h)
amg.setup(A)
peak = 6454.9
tot_ms = 0
for l in range(amg.num_levels):
    if only >= 0 and l != only:
        continue
    for name, M in (("A", amg.level_A(l)), ("Asorted", amg.level_A(l).sorted_copy() if l > 0 else None),
                    ("P", amg.level_P(l) if l < amg.num_levels - 1 else None)):
        if M is None or not M.p:
            continue
        n, m, nnz = M.dims
        x = h.zeros(m); h.fill(x, 1.0)
        y = h.zeros(n)
        for _ in range(3):
            M.matvec(1.0, x, 0.0, None, y)
        reps = 20
        h.timer_start()
        for _ in range(reps):
            M.matvec(1.0, x, 0.0, None, y)
        ms = h.timer_stop_ms() / reps
        b = 12.0 * nnz + 4.0 * (n + 1) + 8.0 * n + 8.0 * m
        print(json.dumps({"level": l, "mat": name, "rows": n, "nnz": nnz, "avg": round(nnz / max(n, 1), 1), "us": round(ms * 1e3, 1),
                          "GBs": round(b / ms * 1e-6, 1), "frac": round(b / ms * 1e-6 / peak, 3)}))
        x.free(); y.free()
