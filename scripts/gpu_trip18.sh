#!/bin/bash
mkdir -p gpurun_out
for g in 2 4 8 16 38; do
  for shape in "9728 9216 3072" "9728 3072 15360"; do
    UTX_GEMM_GROUP_M=$g timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:gemm2 -s 2 -c 1 --csv python scripts/gemm_one.py $shape 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' -v g=$g -v s="$shape" '{printf "g=%s [%s] %s %s %s\n", g, s, $(NF-2), $(NF-1), $NF}'
  done
done
run() { timeout -k 10 600 python bench.py --no-bake --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print(sys.argv[1], round(d['value'],3), round(d['ms_per_step'],1), 'gemm', round(r['ms_per_step'],1), round(r['achieved']), 'attn', round(r['attention']['ms_per_step'],1), d['clocks']['sm_mhz'])" "$1"; }
UTX_GEMM_GROUP_M=4 run g4
UTX_GEMM_GROUP_M=8 run g8
UTX_GEMM_GROUP_M=16 run g16
UTX_GEMM_GROUP_M=38 run g38
