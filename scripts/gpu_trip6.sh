#!/bin/bash
mkdir -p gpurun_out
run() { timeout -k 10 ${4:-300} python -m pytest "$2" -q -m gpu --timeout 250 -k "$3" > gpurun_out/$1.log 2>&1; echo "$1 exit $?"; grep -E "passed|failed" gpurun_out/$1.log | tail -1; grep -E "^(FAILED|ERROR)" gpurun_out/$1.log | head; }
run bake tests/test_gpu_bake.py "bake or raster or lbvh or raytracing"
run e2e tests/test_gpu_e2e.py "pipeline" 600
tail -15 gpurun_out/e2e.log
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 1 -c 1 -f -o gpurun_out/r01_attention2_kernel \
  python scripts/profile_kernels.py 2 > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attn2 exit $?"
