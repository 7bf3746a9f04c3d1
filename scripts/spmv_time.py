"""y = A0 x timing (CUDA events, 100 repetitions) on the 256^3 7-point operator."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hypre_ve_b200 as hb
h = hb.Handle(0)
A = hb.ParCsr.laplacian(h, 256, 256, 256)
n = A.local[0]
x = h.zeros(n); h.fill(x, 1.0)
y = h.zeros(n)
for _ in range(10):
    A.matvec(1.0, x, 0.0, None, y)
h.timer_start()
for _ in range(100):
    A.matvec(1.0, x, 0.0, None, y)
print("stages", os.environ.get("B200_DICT_STAGES", "2"), "bytes/entry", A.diag.stream_bytes_per_entry, "spmv ms %.4f" % (h.timer_stop_ms() / 100))
