#!/bin/bash
# first GPU bring-up: primitives, parity, phase times.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -q --tb=short > gpurun_out/prim.log 2>&1
echo "prim exit $?" >> gpurun_out/prim.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "not full_size" > gpurun_out/parity.log 2>&1
echo "parity exit $?" >> gpurun_out/parity.log
timeout 600 python tools/phase_times.py 1000000 10000000 > gpurun_out/phase.log 2>&1
echo "phase exit $?" >> gpurun_out/phase.log
tail -5 gpurun_out/prim.log; tail -30 gpurun_out/parity.log; tail -8 gpurun_out/phase.log
