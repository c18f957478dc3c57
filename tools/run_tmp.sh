mkdir -p gpurun_out
timeout 600 python -m pytest tests -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests.log
timeout 200 python tools/phase_times.py 100000000 --kinds=dna,dna_nl 2>&1 | grep -o '"lcp_phases_ms[^}]*}'
