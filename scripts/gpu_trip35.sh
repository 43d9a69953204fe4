#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['uv_bake']['value'], d['uv_bake'].get('cpu_baseline'), d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
timeout -k 10 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-250
