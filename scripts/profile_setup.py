"""Setup only (config 2), profiled under ncu for the per-kernel launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n1, n1, n1)
amg = hb.Amg(h)
amg.setup(A)
print("levels", amg.num_levels, "phases", [round(x, 1) for x in amg.setup_times()])
