#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_bake.py tests/test_gpu_fullsize.py tests/test_gpu_e2e.py -q -m gpu --timeout 300 > gpurun_out/bake_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/bake_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:" gpurun_out/bake_tests.log | head -8
timeout 300 python scripts/profile_bake.py 2>&1 | tail -1
timeout 300 python scripts/profile_bake.py 2>&1 | tail -1
