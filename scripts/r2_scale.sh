#!/bin/bash
# usage: scripts/r2_scale.sh N tag  -- bench at N GPUs (direct path), one line
N=$1; TAG=$2; shift 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
echo "rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("n_gpus","value","setup_s","solve_s","iterations","reference_iterations","spmv_gbs","gpu_launches")}, d["roofline_solve"]["ms_per_iteration"], d["roofline_solve"]["frac"], d["e2e"]["value"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/${TAG}.err").read()[-3000:])
P
