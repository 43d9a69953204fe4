#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nn_query_kernel" -s 2 -c 1 -o gpurun_out/r01_nn_query_persist -f python scripts/profile_bake.py > gpurun_out/bake_ncu_full.log 2>&1; echo "ncu full exit $?"
