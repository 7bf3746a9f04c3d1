#!/bin/bash
# usage: scripts/r2_final.sh TAG -- GPU suite, smoke, bench (both arms) on one B200 (no profiler pass)
TAG=$1
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/${TAG}_pytest.log
nvidia-smi --query-gpu=memory.used --format=csv,noheader
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1]); r=json.loads(open("gpurun_out/${TAG}_bench_reference.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","setup_s","solve_s","iterations","spmv_ms","gpu_launches")}, d["e2e"]["value"], d["roofline_solve"]["frac"], d["setup_phases_ms"], "reference", r["value"], d.get("cpu_baseline",{}).get("value"))
P
