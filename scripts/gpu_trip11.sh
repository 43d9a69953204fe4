#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 250 -k "attention" > gpurun_out/attn_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/attn_tests.log | tail -1
timeout -k 10 120 python scripts/bench_attn.py > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err; echo "bench_attn exit $?"; cat gpurun_out/bench_attn.json
UTX_ATTN_IMPL=2 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 2 -c 1 -f -o gpurun_out/r01_attention2b python scripts/bench_attn.py > gpurun_out/ncu_attn2b.log 2>&1; echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
