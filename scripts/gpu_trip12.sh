#!/bin/bash
mkdir -p gpurun_out
for p in 0 1 2 3; do echo "poly $p"; UTX_ATTN_POLY=$p timeout -k 10 120 python scripts/bench_attn.py 2>&1 | tail -1; done
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 250 -k "attention" > gpurun_out/attn_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/attn_tests.log | tail -1
