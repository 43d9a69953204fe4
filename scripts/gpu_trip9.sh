#!/bin/bash
# full GPU suite + default bench + current bake launch list
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -1; grep -E "^(FAILED|ERROR)" gpurun_out/gpu_tests.log | head -8
timeout -k 10 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'gemm', r['ms_per_step'], r['achieved'], 'attn', r['attention']['ms_per_step'], r['attention']['achieved'], 'elem', r['elementwise_ms_per_step'], d['clocks']); print('e2e', d['e2e']); print('delight', d.get('delight')); print('eager', d.get('gpu_eager_baseline')); print('cpu', d.get('cpu_baseline')); print('bake', {k: d['uv_bake'][k] for k in ('value','gpu_ms_per_bake','bvh_build_ms')}); print('vae', d['vae_decode'])"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bake_launches.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; echo "ncu exit $?"
