mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m "gpu" -q --tb=short -x --durations=5 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -14 gpurun_out/gpu_tests.log
for v in 1 2 4; do echo "== INDUCE=$v tests"; B200SA_INDUCE=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -q --tb=short -x -k "adversarial or medium or long_runs or kat" 2>&1 | tail -2; done
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -c 2500 gpurun_out/bench_r02b.json
