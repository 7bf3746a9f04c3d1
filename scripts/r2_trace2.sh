#!/bin/bash
# usage: scripts/r2_trace2.sh N tag -- bench at N GPUs, then one traced step (B200_TRACE=1 B200_TRACE2=1) for the setup diagnosis
N=$1; TAG=$2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
echo "rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("n_gpus","value","setup_s","solve_s","iterations","reference_iterations","spmv_gbs","gpu_launches")}, d["roofline_solve"]["ms_per_iteration"], d["roofline"].get("stream_bytes_per_entry"), d["e2e"]["value"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/${TAG}.err").read()[-3000:])
P
B200_TRACE=1 B200_TRACE2=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_trace.json 2> gpurun_out/${TAG}_trace.err
echo "trace rc=$?"
