#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_bake.py tests/test_gpu_fullsize.py -q -m gpu --timeout 300 -k "bake or knn or lbvh or raster" > gpurun_out/bake_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/bake_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:|Error" gpurun_out/bake_tests.log | head -12
echo "new:"; timeout 300 python scripts/bake_ab.py 2>&1 | tail -1
echo "fused:"; UTX_BAKE_FUSED_TEXEL=1 timeout 300 python scripts/bake_ab.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bake_launches_v6.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; echo "ncu exit $?"
