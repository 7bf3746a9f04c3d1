"""Diagnostic: AMG-GMRES / AMG-PCG on a large convection-diffusion operator through the single-GPU entry points and
through the row-partitioned ones with one rank; prints level tables and the first residual norms (one JSON line per
variant) to compare with `oracle/_ref/ref_dump -difconv ... -nodump` on the same size.
usage: diag_large.py EDGE variant[,variant...]   variant = path:agg:solver  (path s|d, solver 1|3)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb

n = int(sys.argv[1])
h = hb.Handle(0)
for var in sys.argv[2].split(","):
    path, agg, solver = var.split(":")
    agg, solver = int(agg), int(solver)
    out = dict(edge=n, path=path, agg=agg, solver=solver)
    if path == "s":
        A = hb.ParCsr.difconv(h, n, n, n)
        amg = hb.Amg(h, RelaxType=18, ModuleRAP2=0, AggNumLevels=agg)
        amg.setup(A)
        N = n ** 3
        b = h.zeros(N); h.fill(b, 1.0)
        x = h.zeros(N)
        if solver == 3:
            its, rel, norms, _ = h.gmres(A, amg, b, x, tol=1e-8, max_iter=80)
        else:
            its, rel, norms = h.pcg(A, amg, b, x, tol=1e-8, max_iter=80)
        out["levels"] = [list(amg.level_A(l).dims)[:3] for l in range(amg.num_levels)]
        amg.destroy(); A.destroy(); b.free(); x.free()
    else:
        c = hb.Comm.single(h)
        A = hb.DistMatrix.difconv(h, c, n, n, n, 1, 1, 1)
        prm = hb.Amg(h, RelaxType=18, ModuleRAP2=0, AggNumLevels=agg)
        amg = hb.DistAmg(h, c, prm, A)
        b, x = A.vector(1.0), A.vector(0.0)
        if solver == 3:
            its, rel, norms, _ = hb.dist_gmres(h, c, A, amg, b, x, tol=1e-8, max_iter=80)
        else:
            its, rel, norms = hb.dist_pcg(h, c, A, amg, b, x, tol=1e-8, max_iter=80)
        out["levels"] = [[amg.level_A(l).info["global_rows"], amg.level_A(l).info["local_nnz"]] for l in range(amg.num_levels)]
        amg.destroy(); A.destroy(); b.free(); x.free()
    out.update(its=its, rel=rel, norms=[float(v) for v in norms[:8]])
    print(json.dumps(out), flush=True)
