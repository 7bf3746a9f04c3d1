"""Per-kernel achieved bandwidth of the solve phase at config 2 (256^3 7-pt), CUDA-event timed, against the
algorithmic bytes of SURVEY.md 8(d).  usage: kernel_roofline.py [N] [7|27]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
st = int(sys.argv[2]) if len(sys.argv) > 2 else 7
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n1, n1, n1) if st == 7 else hb.ParCsr.laplacian27(h, n1, n1, n1)
amg = hb.Amg(h)
amg.setup(A)
rows = []


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    h.timer_start()
    for _ in range(reps):
        fn()
    return h.timer_stop_ms() / reps


def rec(name, ms, byts):
    gbs = byts / ms / 1e6
    rows.append((name, ms, byts / 1e9, gbs, gbs / peak))
    print("%-46s %8.3f ms  %7.3f GB  %8.1f GB/s  %5.2f of measured peak" % (name, ms, byts / 1e9, gbs, gbs / peak))


nl = amg.num_levels
tot_cycle = 0.0
for l in range(min(nl - 1, 4)):
    M = amg.level_A(l)
    n, m, nnz = M.dims
    x = h.zeros(m); h.fill(x, 1.0)
    y = h.zeros(n); b = h.zeros(n); h.fill(b, 1.0)
    rec("SpMV y=A%d x          (12nnz+4N+16N)" % l, timeit(lambda: M.matvec(1.0, x, 0.0, None, y)), 12.0 * nnz + 20.0 * n)
    rec("residual r=b-A%d x    (12nnz+4N+24N)" % l, timeit(lambda: M.matvec(-1.0, x, 1.0, b, y)), 12.0 * nnz + 28.0 * n)
    Pm = amg.level_P(l)
    pn, pm, pnnz = Pm.dims
    xc = h.zeros(pm); h.fill(xc, 1.0)
    rec("prolong u+=P%d e      (12nnz+4Nf+16Nf+8Nc)" % l, timeit(lambda: Pm.matvec(1.0, xc, 1.0, y, y)), 12.0 * pnnz + 20.0 * pn + 8.0 * pm)
    R = Pm.transpose()
    yc = h.zeros(pm)
    rec("restrict f=R%d r      (12nnz+4Nc+8Nc+8Nf)" % l, timeit(lambda: R.matvec(1.0, y, 0.0, None, yc)), 12.0 * pnnz + 12.0 * pm + 8.0 * pn)
    tot_cycle += 3 * (12.0 * nnz) + 2 * 36.0 * n + 28.0 * n + 2 * 12.0 * pnnz + 28.0 * pn + 20.0 * pm
    R.destroy()
    for v in (x, y, b, xc, yc):
        v.free()
N = A.local[0]
x = h.zeros(N); h.fill(x, 1.0); y = h.zeros(N); h.fill(y, 2.0)
rec("dot <x,y>            (16N)", timeit(lambda: h.dot(x, y)), 16.0 * N)
rec("axpy y+=a x          (24N)", timeit(lambda: h.axpy(0.5, x, y)), 24.0 * N)
# whole V(1,1) cycle: sum over ALL levels of 3 A-passes (2 Jacobi sweeps, the pre-sweep from zero skips its SpMV,
# + residual) + restriction + prolongation, SURVEY 8(d) formula with the pre-sweep shortcut
byt = 0.0
for l in range(nl - 1):
    n, m, nnz = amg.level_A(l).dims
    pn, pm, pnnz = amg.level_P(l).dims
    byt += (12.0 * nnz + 28.0 * n) + (12.0 * nnz + 36.0 * n) + 24.0 * n + 2 * 12.0 * pnnz + 28.0 * pn + 20.0 * pm
f = h.zeros(N); h.fill(f, 1.0); u = h.zeros(N)
rec("V(1,1) cycle, all %d levels (l1-Jacobi)" % nl, timeit(lambda: amg.solve(f, u), reps=20), byt)
its, rel, norms = h.pcg(A, amg, f, u, tol=1e-8, max_iter=100)
h.fill(u, 0.0)
h.timer_start()
its, rel, norms = h.pcg(A, amg, f, u, tol=1e-8, max_iter=100)
ms = h.timer_stop_ms()
n0, _, nnz0 = amg.level_A(0).dims
rec("PCG iteration (cycle + SpMV + 5 vector passes)", ms / its, byt + 12.0 * nnz0 + 20.0 * n0 + (16 * 3 + 24 * 2 + 24) * n0)
print("PCG: %d iterations, %.1f ms" % (its, ms))
