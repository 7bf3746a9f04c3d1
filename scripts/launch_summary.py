"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`):
total serialised kernel time, the SpMV family's share, and the kernels above 0.4 % of the step.
usage: launch_summary.py launches.csv "header line" > summary.txt"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"\(.*$", "", name).replace("<unnamed>::", "")
    tot[name] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
    cnt[name] += 1
total = sum(tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("total serialised kernel time %.1f ms over %d launches" % (total, sum(cnt.values())))
spmv = sum(v for k, v in tot.items() if k.startswith("spmv_pipe_kernel"))
print("  SpMV family (spmv_pipe_kernel<*>, all levels): %.1f ms = %.1f%% of the step" % (spmv, 100 * spmv / total))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v / total < 0.004:
        break
    print("  %10.3f ms  %5.1f%% n=%5d  %s" % (v, 100 * v / total, cnt[k], k[:150]))
