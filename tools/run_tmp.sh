mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py tests/test_sharded.py tests/test_gpu_next.py -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests_q.log 2>&1; tail -2 gpurun_out/gpu_tests_q.log
for v in wide narrow; do
if [ $v = narrow ]; then export B200SA_SORT_NARROW=1; fi
python tools/phase_times.py --kinds=dna,english,tiled 100000000 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$v', d['input'], round(sum(d['phases_ms'].values()),2), {k:v for k,v in d['phases_ms'].items() if 'induce' not in k})"
done
