#!/bin/bash
# What still has to run on a B200, cheapest first; meant for ONE gpurun call of about
# 4 minutes:  /usr/local/graft/bin/gpurun --timeout 420 -- 'bash scripts/gpu_verify_pending.sh'
# Logs go to gpurun_out/ (copy the ones worth keeping into profiles/).
set -u
mkdir -p gpurun_out
# 1. the cases written after round 1's GPU minutes ran out (xfail-guarded: look for XPASS / xfailed in the log)
python -m pytest tests/test_gpu_zz_hmis.py tests/test_gpu_userrows.py -q -rxX --tb=short -p no:cacheprovider > gpurun_out/pending.log 2>&1
tail -n 25 gpurun_out/pending.log
# 2. the files of the suite that the last round-1 run did not reach on the final tree
python -m pytest tests/test_gpu_dist.py tests/test_gpu_fullsize.py tests/test_gpu_gs.py tests/test_gpu_ij.py tests/test_gpu_krylov.py \
       tests/test_gpu_spmv.py -q --tb=short -p no:cacheprovider > gpurun_out/rest_of_suite.log 2>&1
tail -n 5 gpurun_out/rest_of_suite.log
# 3. headline line (config 2) once more on the final tree
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err
cat gpurun_out/bench_next.json
