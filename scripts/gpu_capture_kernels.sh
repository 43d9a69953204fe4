#!/bin/bash
# ncu --set full captures of the hot kernels at the bench shapes (one launch each) -> gpurun_out/${R}_<name>.ncu-rep.
# The text summaries judged are made from these with scripts/ncu_summarise.sh (profiles/${R}_<name>.metrics.csv).
mkdir -p gpurun_out
R=${ROUND:-r02}
for k in gemm2_bf16_tn attention2; do
  timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${R}_${k}_final \
    python scripts/profile_kernels.py 2 > gpurun_out/${R}_ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
# the implicit-GEMM convolution of the VAE (256 -> 256 channels at 1024^2): profile_kernels.py launches it last; with 1 iteration
# the script issues 2 plain gemm2 launches before it
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_tn -s 2 -c 1 -f -o gpurun_out/${R}_vae_conv_final \
  python scripts/profile_kernels.py 1 > gpurun_out/${R}_ncu_vae.log 2>&1
echo "ncu vae conv exit $?"
