#!/bin/bash
# full GPU suite after the reference-golden work (Euler increment rounding, fp32 schedule) + smoke
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu --timeout 400 -s > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?"
grep -E "passed|failed" gpurun_out/gpu_tests.log | tail -1; grep -E "^(FAILED|ERROR)|latent .* dB|Error" gpurun_out/gpu_tests.log | head -20
timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
