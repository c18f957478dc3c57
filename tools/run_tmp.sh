mkdir -p gpurun_out
timeout 600 python -m pytest tests -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -5 gpurun_out/gpu_tests.log
B200SA_TRACE=1 timeout 300 python tools/phase_times.py 100000000 --kinds=english,tiled,dna_nl 2>&1 | grep -v "doubling round" | cut -c1-700
