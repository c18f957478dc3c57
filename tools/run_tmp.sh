mkdir -p gpurun_out
timeout 600 python -m pytest tests -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -12 gpurun_out/gpu_tests.log
echo "== default (fused classify, v3)"; timeout 200 python tools/phase_times.py 100000000 --kinds=dna,bytes 2>&1 | tee gpurun_out/phase100.log
echo "== CLASSIFY_V1"; B200SA_CLASSIFY_V1=1 timeout 200 python tools/phase_times.py 100000000 --kinds=dna 2>&1 | grep -o '"phases_ms[^}]*'
timeout 120 python tools/steplog.py 100000000 > gpurun_out/steplog_dna.txt 2>&1; head -80 gpurun_out/steplog_dna.txt
