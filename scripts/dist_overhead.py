"""What the row-partitioned code path costs beyond the single-GPU one, on ONE GPU: the same 256^3 operator through
b200_amg_setup (phase times) and through b200_dist_amg_setup with a one-rank communicator (B200_TRACE=1 phases on stderr)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, n, n, n)
for rep in range(2):
    amg = hb.Amg(h, ModuleRAP2=0)
    h.timer_start(); amg.setup(A); ms = h.timer_stop_ms()
    t = amg.setup_times()
    amg.destroy()
print("single: setup %.1f ms  strength %.1f pmis %.1f interp %.1f trunc %.1f transpose %.1f rap %.1f other %.1f" % ((ms,) + tuple(t[:7])))
A.destroy()
comm = hb.Comm.single(h)
D = hb.DistMatrix.laplacian(h, comm, n, n, n, 1, 1, 1, 7)
prm = hb.Amg(h, ModuleRAP2=0)
b, x = D.vector(1.0), D.vector(0.0)
for rep in range(2):
    if rep == 1:
        os.environ["B200_TRACE"] = "1"
    amg = hb.DistAmg(h, comm, prm, D)
    os.environ.pop("B200_TRACE", None)
    s_ms = amg.setup_ms
    h.fill(x, 0.0)
    h.timer_start()
    its, rel, _ = hb.dist_pcg(h, comm, D, amg, b, x, tol=1e-8, max_iter=200)
    v_ms = h.timer_stop_ms()
    amg.destroy()
print("dist path, 1 rank: setup %.1f ms (traced, synchronised) solve %.1f ms its %d" % (s_ms, v_ms, its))
amg = hb.DistAmg(h, comm, prm, D)
print("dist path, 1 rank: setup %.1f ms (untraced)" % amg.setup_ms)
