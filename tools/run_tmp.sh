mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -q -x -k "variants or long_runs or adversarial" > gpurun_out/gpu_tests_q.log 2>&1; tail -2 gpurun_out/gpu_tests_q.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_r02e.json
tools/gpu_ncu_full.sh r02j induce6 "k_induce6" 2 2 | tail -3
tools/gpu_launches.sh r02j 100000000 dna | tail -25
python tools/steplog.py 100000000 > gpurun_out/steplog_v6c.txt 2>&1; grep -E "tiles=   (191|382|143|144) " gpurun_out/steplog_v6c.txt | head -4
