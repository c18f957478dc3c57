mkdir -p gpurun_out
timeout 600 python -m pytest tests -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -12 gpurun_out/gpu_tests.log
echo "== default"; timeout 200 python tools/phase_times.py 100000000 --kinds=dna,dna_nl 2>&1 | tee gpurun_out/phase100.log
echo "== L2PERSIST=48"; B200SA_L2PERSIST=48 timeout 200 python tools/phase_times.py 100000000 --kinds=dna,dna_nl 2>&1 | tee gpurun_out/phase100_l2.log
echo "== INDUCE_V1"; B200SA_INDUCE_V1=1 timeout 200 python tools/phase_times.py 100000000 --kinds=dna 2>&1 | tee gpurun_out/phase100_v1.log
tools/gpu_ncu_full.sh r02c scan 'k_scan_lb.*InLmsActive1' 1 1
tools/gpu_ncu_full.sh r02c ospass 'k_os_pass|k_os_hist' 5 3
tools/gpu_ncu_full.sh r02c induce2 'k_induce2' 2 2
