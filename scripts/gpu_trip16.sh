#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 250 -k "gemm" > gpurun_out/gemm_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/gemm_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:" gpurun_out/gemm_tests.log | head -8
timeout -k 10 200 python scripts/bench_gemm.py 2> gpurun_out/bench_gemm.err | tail -3
run() { timeout -k 10 600 python bench.py --no-bake --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(sys.argv[1], round(d['value'],3), round(d['ms_per_step'],1), 'gemm', round(r['ms_per_step'],1), round(r['achieved']), 'attn', round(r['attention']['ms_per_step'],1), round(r['attention']['achieved']), d['clocks']['sm_mhz'])" "$1"; }
run default
run default
UTX_GEMM_IMPL=1 run gemm1cta
