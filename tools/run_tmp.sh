mkdir -p gpurun_out
timeout 600 python -m pytest tests -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -8 gpurun_out/gpu_tests.log
echo "== default (v5/16-bit)"; timeout 200 python tools/phase_times.py 100000000 --kinds=dna 2>&1 | cut -c1-600
timeout 120 python tools/steplog.py 100000000 > gpurun_out/steplog_dna_v5.txt 2>&1; head -30 gpurun_out/steplog_dna_v5.txt
