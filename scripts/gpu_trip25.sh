#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_bake.py -q -m gpu --timeout 300 > gpurun_out/bake_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/bake_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:|Error" gpurun_out/bake_tests.log | head -12
for m in 1 0 1 0; do
echo "persist $m:"; UTX_NN_PERSIST=$m timeout 300 python scripts/bake_ab.py 2>&1 | tail -1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bake_launches_v7.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; grep "nn_query" gpurun_out/r01_bake_launches_v7.csv | tail -1 | awk -F'","' '{print $5, $NF}'
