#!/bin/bash
mkdir -p gpurun_out
run() { timeout -k 10 600 python bench.py --no-bake --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(sys.argv[1], round(d['value'],3), round(d['ms_per_step'],1), 'gemm', round(r['ms_per_step'],1), 'attn', round(r['attention']['ms_per_step'],1), round(r['attention']['achieved']), d['clocks']['sm_mhz'])" "$1"; }
UTX_ATTN_POLY=0 run poly0
UTX_ATTN_POLY=1 run poly1
UTX_ATTN_POLY=2 run poly2
UTX_ATTN_POLY=3 run poly3
UTX_ATTN_IMPL=3 run impl3
UTX_GEMM_IMPL=1 run gemm1cta
