mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m "gpu" -q --tb=short -x --durations=4 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -9 gpurun_out/gpu_tests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_r02d.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_r02d.json 2>> gpurun_out/bench_r02d.err; tail -c 600 gpurun_out/ref_r02d.json
tools/gpu_launches.sh r02h 100000000 dna | tail -22
