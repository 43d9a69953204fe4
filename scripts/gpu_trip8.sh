#!/bin/bash
mkdir -p gpurun_out
run() { timeout -k 10 ${4:-300} python -m pytest "$2" -q -m gpu --timeout 250 -k "$3" > gpurun_out/$1.log 2>&1; echo "$1 exit $?"; grep -E "passed|failed" gpurun_out/$1.log | tail -1; grep -E "^(FAILED|ERROR)" gpurun_out/$1.log | head -8; }
run attn tests/test_gpu_kernels.py "attention"
timeout -k 10 120 python scripts/bench_attn.py > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err; echo "bench_attn exit $?"; cat gpurun_out/bench_attn.json
run gemm1 tests/test_gpu_kernels.py "gemm and gemm_impl0 or gemm and 1-"
run gemm tests/test_gpu_kernels.py "gemm"
grep -E "utx:|Error|error" gpurun_out/gemm.log | head -10
timeout -k 10 200 python scripts/bench_gemm.py > gpurun_out/bench_gemm.json 2> gpurun_out/bench_gemm.err; echo "bench_gemm exit $?"; cat gpurun_out/bench_gemm.json; tail -3 gpurun_out/bench_gemm.err
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --no-bake --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'gemm', r['ms_per_step'], r['achieved'], 'attn', r['attention']['ms_per_step'], r['attention']['achieved'], 'elem', r['elementwise_ms_per_step'], d['clocks'])"
UTX_GEMM_IMPL=2 timeout -k 10 900 python bench.py --steps 5 --warmup 3 --no-bake --no-cpu-baseline > gpurun_out/bench_gemm2.json 2> gpurun_out/bench_gemm2.err; echo "bench(gemm2) exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_gemm2.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'gemm', r['ms_per_step'], r['achieved'], 'attn', r['attention']['ms_per_step'], r['attention']['achieved'])"; tail -3 gpurun_out/bench_gemm2.err
