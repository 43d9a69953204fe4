#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
tail -2 gpurun_out/bench_n$N.err | cut -c1-300; tail -1 gpurun_out/bench_n$N.json | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
