#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_bake.py tests/test_gpu_fullsize.py tests/test_gpu_e2e.py -q -m gpu --timeout 300 > gpurun_out/bake_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/bake_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:|Error" gpurun_out/bake_tests.log | head -12
for i in 1 2 3; do timeout 300 python scripts/bake_ab.py 2>&1 | tail -1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bake_launches_v8.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ray_kernel|nn_query_kernel|texel_prep" -s 3 -c 3 -o gpurun_out/r01_bake_ray_nn_final -f python scripts/profile_bake.py > gpurun_out/bake_ncu_full.log 2>&1; echo "ncu full exit $?"
