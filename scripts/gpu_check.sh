#!/bin/bash
# One gpurun trip: GPU parity tests group by group (each group in its own process under its own timeout, so a trapped
# or hung kernel cannot poison the others), smoke, then a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, file, -k expression
  timeout -k 10 ${TEST_TIMEOUT:-300} python -m pytest "$2" -q -m gpu --timeout 200 -k "$3" > gpurun_out/$1.log 2>&1
  echo "$1 exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|error" gpurun_out/$1.log | tail -2
  grep -E "^(FAILED|ERROR)" gpurun_out/$1.log | head -12
}
run elem tests/test_gpu_kernels.py "ln_modulate or rope or gemv"
run gemm tests/test_gpu_kernels.py "gemm"
run attn tests/test_gpu_kernels.py "attention"
run flux tests/test_gpu_flux.py "flux or oracle or denoise or forward"
run pipe tests/test_gpu_pipeline.py "pipeline"
run vae tests/test_gpu_vae.py "vae or decode or encode or pipeline"
run bake tests/test_gpu_bake.py "bake or raster or lbvh or raytracing"
for f in ${EXTRA_TESTS}; do run $(basename $f .py) $f ""; done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log
if [ "${RUN_BENCH:-1}" = "1" ]; then
  timeout -k 10 ${BENCH_TIMEOUT:-900} python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" | tee -a gpurun_out/summary.txt
  tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
