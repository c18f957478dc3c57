mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_parity.py -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests_tma.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_tma.log; tail -4 gpurun_out/gpu_tests_tma.log
for i in 1 2; do
python tools/phase_times.py --kinds=dna 100000000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('TMA', d['phases_ms'])"
B200SA_CLASSIFY_NO_TMA=1 python tools/phase_times.py --kinds=dna 100000000 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('NOTMA', d['phases_ms'])"
done
tools/gpu_ncu_full.sh r02i induce6 "k_induce6" 2 2 | tail -3
tools/gpu_ncu_full.sh r02i classify "k_classify_fused" 1 1 | tail -3
python tools/steplog.py 100000000 > gpurun_out/steplog_v6.txt 2>&1; tail -5 gpurun_out/steplog_v6.txt
