#!/bin/bash
# regression: all GPU tests + phase times at 100 MB.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/gpu_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
timeout 600 python tools/phase_times.py 100000000 > gpurun_out/phase100.log 2>&1
cat gpurun_out/phase100.log
