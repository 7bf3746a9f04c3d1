"""Small end-to-end exercise of every kernel family, meant to be run under compute-sanitizer
(memcheck / racecheck / synccheck):  compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import sys, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
h = hb.Handle(0)
which = sys.argv[1:] or ["jacobi", "gs", "cheby", "agg", "wcycle", "27pt", "dist"]


def solve(A, **kw):
    amg = hb.Amg(h, **kw)
    amg.setup(A)
    n = A.local[0]
    b = h.zeros(n); h.fill(b, 1.0)
    x = h.zeros(n)
    its, rel, _ = h.pcg(A, amg, b, x, tol=1e-8, max_iter=100)
    amg.destroy(); b.free(); x.free()
    return its, rel


A = hb.ParCsr.laplacian(h, 24, 22, 20)
if "jacobi" in which:
    print("l1-Jacobi, both Galerkin orders", solve(A, ModuleRAP2=0), solve(A, ModuleRAP2=1), flush=True)
if "gs" in which:
    for T in (1, 7, 300):
        print("hybrid GS 13/14 blocks", T, solve(A, RelaxType=13, RelaxTypeUp=14, GSBlocks=T, ModuleRAP2=0), flush=True)
    print("GS 6", solve(A, RelaxType=6, GSBlocks=3), flush=True)
if "cheby" in which:
    print("Chebyshev", solve(A, RelaxType=16), flush=True)
if "agg" in which:
    print("aggressive", solve(A, AggNumLevels=1, ModuleRAP2=0), flush=True)
if "wcycle" in which:
    print("W(2,2)", solve(A, CycleType=2, NumSweeps=2), flush=True)
A.destroy()
if "27pt" in which:
    A = hb.ParCsr.laplacian27(h, 14, 14, 14)
    print("27-pt", solve(A, ModuleRAP2=0), flush=True)
    A.destroy()
if "dist" in which:
    group = hb.Comm.group_create(2)
    out = [None, None]

    def body(r):
        hh = hb.Handle(0)
        c = hb.Comm.threads(hh, group, r)
        D = hb.DistMatrix.laplacian(hh, c, 20, 18, 16, 2, 1, 1, 7)
        for kw in (dict(ModuleRAP2=0), dict(ModuleRAP2=1, AggNumLevels=1), dict(RelaxType=13, RelaxTypeUp=14, GSBlocks=5, ModuleRAP2=0)):
            prm = hb.Amg(hh, **kw)
            amg = hb.DistAmg(hh, c, prm, D)
            b, x = D.vector(1.0), D.vector(0.0)
            its, rel, _ = hb.dist_pcg(hh, c, D, amg, b, x, tol=1e-8, max_iter=100)
            amg.destroy()
            out[r] = (its, rel)
    ts = [threading.Thread(target=body, args=(r,)) for r in range(2)]
    [t.start() for t in ts]; [t.join() for t in ts]
    print("2 ranks", out, flush=True)
print("done")
