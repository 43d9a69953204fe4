#!/bin/bash
# 2-GPU run of the bench exactly as the driver launches it (torchrun, NCCL), both arms; + the Euler unit test
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "euler" --timeout 200 2>&1 | tail -2
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
tail -3 gpurun_out/bench_n2.err; tail -1 gpurun_out/bench_n2.json | cut -c1-600
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-500
