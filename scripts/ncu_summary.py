"""Summarise an .ncu-rep (read on the CPU box with `ncu -i ... --page raw --csv`) into the handful of metrics the
roofline discussion uses.  usage: ncu_summary.py file.ncu-rep [algorithmic_bytes_per_launch ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum"]
units = rows[1]
for r in rows[2:]:
    d = dict(zip(h, r))
    u = dict(zip(h, units))
    name = d["Kernel Name"]
    print("kernel:", name[:140])
    def val(k):
        try:
            return float(d[k].replace(",", ""))
        except Exception:
            return None
    t, rd, wr = val("gpu__time_duration.sum"), val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    for k in keep:
        if k in d:
            print("  %-70s %s %s" % (k, d[k], u.get(k, "")))
    def to_bytes(v, unit):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    def to_s(v, unit):
        return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(unit, 1e-9)
    if t and rd is not None and wr is not None:
        tb = to_bytes(rd, u["dram__bytes_read.sum"]) + to_bytes(wr, u["dram__bytes_write.sum"])
        ts = to_s(t, u["gpu__time_duration.sum"])
        print("  => DRAM traffic %.3f GB in %.3f ms = %.0f GB/s (under ncu: cold caches, serialised)" % (tb / 1e9, ts * 1e3, tb / ts / 1e9))
    print()
