#!/bin/bash
# usage: scripts/r2_multi.sh N tag [extra bench args]  -- bench at N GPUs with the direct path and with NCCL send/recv
N=$1; TAG=$2; shift 2
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline "${EXTRA[@]}" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "$name rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","setup_s","solve_s","iterations","reference_iterations","spmv_gbs","gpu_launches")}, d["roofline_solve"]["ms_per_iteration"], d["roofline_solve"]["frac"], d["e2e"]["value"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-2000:])
P
}
EXTRA=("$@")
run p2p B200_P2P=1
run nccl B200_P2P=0
