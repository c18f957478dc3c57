mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_sharded.py tests/test_gpu_next.py -m "gpu and not slow" -q --tb=short -x -k "lcp or bad_table or sharded or medium" > gpurun_out/gpu_tests_q.log 2>&1; tail -2 gpurun_out/gpu_tests_q.log
python tools/phase_times.py --kinds=dna,bytes 100000000 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(d['input'], d['lcp_phases_ms'])"
