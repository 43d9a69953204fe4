#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_vae.py -q -m gpu --timeout 250 > gpurun_out/vae_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed" gpurun_out/vae_tests.log | tail -1; grep -E "^(FAILED|ERROR)|utx:|Error" gpurun_out/vae_tests.log | head -8
python - <<'PY'
import torch, sys, os
sys.path.insert(0, ".")
import bench
print("implicit", bench.bench_vae_decode(torch.device("cuda", 0)))
os.environ["UTX_VAE_IM2COL"] = "1"
print("im2col  ", bench.bench_vae_decode(torch.device("cuda", 0)))
PY
