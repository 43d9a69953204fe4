#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2; do
echo "order $m:"; UTX_NN_ORDER=$m timeout 300 python scripts/bake_ab.py 2>&1 | tail -1
UTX_NN_ORDER=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/nn_order_$m.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; grep nn_query_kernel gpurun_out/nn_order_$m.csv | tail -1 | awk -F'","' '{print $NF}'
done
