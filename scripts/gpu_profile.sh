#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one bench run with its device time, (2) full captures of the hot kernels.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1800} -c ${COUNT:-420} --csv \
  --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
echo "launch list exit $?"
for k in gemm_bf16_tn attention_kernel ${EXTRA_KERNELS}; do
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${R}_$k \
    python scripts/profile_kernels.py 2 > gpurun_out/${R}_ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
ls -la gpurun_out/
