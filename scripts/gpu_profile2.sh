#!/bin/bash
# Round-end ncu evidence for profiles/: (1) every launch of ~1.2 bench steps with its device time, (2) one --set full
# capture of the default GEMM (2-CTA) and of the default attention kernel at the bench shapes, (3) the bake launch list.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1800} -c ${COUNT:-420} --csv \
  --log-file gpurun_out/${R}_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bake > gpurun_out/${R}_bench_under_ncu.log 2>&1
echo "launch list exit $?"
for k in gemm2_bf16_tn attention2; do
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${R}_${k}_final \
    python scripts/profile_kernels.py 2 > gpurun_out/${R}_ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_bake_launches_final.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; echo "bake ncu exit $?"
ls -la gpurun_out/ | head -30
