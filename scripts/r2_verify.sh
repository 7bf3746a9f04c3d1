#!/bin/bash
# usage: scripts/r2_verify.sh TAG  -- the round-end sequence on one B200: GPU suite, smoke, bench (both arms), ncu launch list
TAG=$1
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/${TAG}_pytest.log
nvidia-smi --query-gpu=memory.used --format=csv,noheader
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_reference.json
# launch list of ONE step (setup + solve: about 1850 launches; ncu manages about 10 launches per second on these boxes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
gzip -f gpurun_out/${TAG}_launches.csv
