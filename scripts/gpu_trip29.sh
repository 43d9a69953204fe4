#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?"
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -1; grep -E "^(FAILED|ERROR)|Error" gpurun_out/gpu_tests.log | head -20
timeout -k 10 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'gemm', r['ms_per_step'], r['achieved'], 'attn', r['attention']['ms_per_step'], r['attention']['achieved'], 'elem', r['elementwise_ms_per_step'], d['clocks']); print('e2e', d['e2e']); print('delight', d.get('delight')); print('call', d.get('pipeline_call')); print('eager', d.get('gpu_eager_baseline',{}).get('value')); print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores')); print('bake', {k: d['uv_bake'][k] for k in ('value','gpu_ms_per_bake','bvh_build_ms')}, d['uv_bake']['roofline']['frac']); print('vae', d['vae_decode']['ms'])"
