mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -q --tb=short -x -k "lcp or medium or adversarial or long_runs" > gpurun_out/gpu_tests_q.log 2>&1; tail -2 gpurun_out/gpu_tests_q.log
for k in 1 2 4 1 2 4; do
B200SA_LCP_K=$k python tools/phase_times.py --kinds=dna,dna_nl 100000000 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('K$k', d['input'], d['lcp_phases_ms']['lcp_direct'])"
done
