#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -k "attention" > gpurun_out/attn2.log 2>&1
echo "attn(v1+v2) exit $?"; grep -E "passed|failed" gpurun_out/attn2.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/attn2.log | head -20
timeout -k 10 120 python scripts/bench_attn.py > gpurun_out/bench_attn.json 2> gpurun_out/bench_attn.err
echo "bench_attn exit $?"; cat gpurun_out/bench_attn.json; tail -3 gpurun_out/bench_attn.err
timeout -k 10 300 python -m pytest tests/test_gpu_flux.py -q -m gpu --timeout 200 > gpurun_out/flux2.log 2>&1
echo "flux(v2 default) exit $?"; grep -E "passed|failed" gpurun_out/flux2.log | tail -2
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_bake_launches.csv \
  python scripts/profile_bake.py > gpurun_out/bake_under_ncu.log 2>&1
echo "bake launch list exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:texel_kernel -s 1 -c 1 -f -o gpurun_out/r01_texel_kernel \
  python scripts/profile_bake.py > gpurun_out/ncu_texel.log 2>&1
echo "ncu texel exit $?"
