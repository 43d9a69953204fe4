#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?"
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -1; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/gpu_tests.log | head -30
for i in 1 2; do timeout 300 python scripts/bake_ab.py 2>&1 | tail -1; done
