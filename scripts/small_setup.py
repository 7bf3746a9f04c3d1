"""Setup time of a small operator (the size of a replicated tail level), repeated: where do the milliseconds go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 53
h = hb.Handle(0)
A = hb.ParCsr.laplacian27(h, n1, n1, n1)
for rep in range(4):
    amg = hb.Amg(h, ModuleRAP2=0, CoarsenType=8)
    h.sync(); t0 = time.perf_counter()
    amg.setup(A)
    h.sync(); t1 = time.perf_counter()
    print("rep", rep, "wall ms %.2f" % ((t1 - t0) * 1e3), "levels", amg.num_levels, "phases", [round(x, 2) for x in amg.setup_times()], flush=True)
    amg.destroy()
