#!/bin/bash
mkdir -p gpurun_out
run() { timeout -k 10 ${4:-300} python -m pytest "$2" -q -m gpu --timeout 250 -k "$3" > gpurun_out/$1.log 2>&1; echo "$1 exit $?"; grep -E "passed|failed" gpurun_out/$1.log | tail -1; grep -E "^(FAILED|ERROR)" gpurun_out/$1.log | head; }
run kern tests/test_gpu_kernels.py "attention or rope or gemm"
run flux tests/test_gpu_flux.py ""
timeout -k 10 120 python scripts/bench_attn.py > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err; echo "bench_attn exit $?"; cat gpurun_out/bench_attn.json
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --no-bake --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'gemm', r['ms_per_step'], r['achieved'], 'attn', r['attention']['ms_per_step'], r['attention']['achieved'], 'elem', r['elementwise_ms_per_step'], d['clocks'])"
${EXTRA_CMD}
