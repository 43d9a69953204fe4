#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_bake.py tests/test_gpu_reference_golden.py tests/test_gpu_fullsize.py -q -m gpu --timeout 300 2>&1 | tail -3
for impl in 0 1 1; do echo "UTX_NN_IMPL=$impl"; UTX_NN_IMPL=$impl timeout 300 python scripts/bake_ab.py 2>&1 | tail -1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nn_query|ray_kernel" -c 12 --csv --log-file gpurun_out/r01_nn_refill.csv python scripts/profile_bake.py > gpurun_out/bake_ncu.log 2>&1; echo "ncu exit $?"
grep -E "nn_query|ray_kernel" gpurun_out/r01_nn_refill.csv | awk -F'","' '{print substr($5,1,50), $(NF)}' | head -12
