#!/bin/bash
mkdir -p gpurun_out
timeout 60 scripts/_build/attn_trace 9728 > gpurun_out/attn_trace.txt 2>&1; echo "trace exit $?"; cat gpurun_out/attn_trace.txt | head -40
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 250 -k "attention" > gpurun_out/attn_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/attn_tests.log | tail -1; grep -E "^(FAILED|ERROR)" gpurun_out/attn_tests.log | head -8
timeout -k 10 120 python scripts/bench_attn.py > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err; echo "bench_attn exit $?"; cat gpurun_out/bench_attn.json
