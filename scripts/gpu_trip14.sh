#!/bin/bash
mkdir -p gpurun_out
UTX_ATTN_IMPL=3 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attention3 -s 2 -c 1 -f -o gpurun_out/r01_attention3 python scripts/bench_attn.py > gpurun_out/ncu_attn3.log 2>&1; echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
