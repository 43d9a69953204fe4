#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -x -k "attention" > gpurun_out/attn_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/attn_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:|rel err" gpurun_out/attn_tests.log | head -8
timeout -k 10 120 python scripts/bench_attn.py 2>&1 | tail -2
