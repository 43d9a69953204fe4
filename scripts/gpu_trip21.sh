#!/bin/bash
# bake: packet ray traversal + clustered point tree: parity, A/B timing, launch list
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_bake.py tests/test_gpu_fullsize.py tests/test_gpu_e2e.py tests/test_gpu_pipeline.py -q -m gpu --timeout 300 > gpurun_out/bake_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/bake_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:|Error" gpurun_out/bake_tests.log | head -12
echo "new:"; timeout 300 python scripts/bake_ab.py 2>&1 | tail -1; timeout 120 python scripts/knn_diag.py 2>&1 | tail -8
echo "fused:"; UTX_BAKE_FUSED_TEXEL=1 timeout 300 python scripts/bake_ab.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bake_launches_v5.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; echo "ncu exit $?"
