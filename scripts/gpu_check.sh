#!/bin/bash
# One gpurun trip: GPU parity tests file by file (each under its own timeout so a hung kernel cannot eat the box),
# smoke, then a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_kernels.py tests/test_gpu_flux.py ${EXTRA_TESTS}; do
  n=$(basename $f .py)
  timeout -k 10 ${TEST_TIMEOUT:-300} python -m pytest $f -q -m gpu -x --timeout 120 ${PYTEST_ARGS} > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$n.log
done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log
if [ "${RUN_BENCH:-1}" = "1" ]; then
  timeout -k 10 ${BENCH_TIMEOUT:-900} python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" | tee -a gpurun_out/summary.txt
  tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
