mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests_q.log 2>&1; tail -2 gpurun_out/gpu_tests_q.log
python tools/phase_times.py --kinds=dna 100000000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('FLUSH', d['phases_ms'])"
python tools/phase_times.py --kinds=dna 100000000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('FLUSH', d['phases_ms'])"
bash tools/sanitize.sh
