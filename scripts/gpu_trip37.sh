#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py tests/test_gpu_reference_golden.py -q -m gpu --timeout 300 2>&1 | tail -2
timeout -k 10 600 python bench.py --no-bake --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/bench_ln.json 2> gpurun_out/bench_ln.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_ln.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'gemm', r['ms_per_step'], 'attn', r['attention']['ms_per_step'], 'elem', r['elementwise_ms_per_step'], d['clocks'])"
