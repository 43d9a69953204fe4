#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_bake.py -q -m gpu --timeout 300 -k knn > gpurun_out/bake_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/bake_tests.log | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ray_kernel|nn_query_kernel" -s 2 -c 2 -o gpurun_out/r01_bake_ray_nn -f python scripts/profile_bake.py > gpurun_out/bake_ncu_full.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/bake_ncu_full.log
