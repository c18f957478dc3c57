mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q --tb=short -x > gpurun_out/gpu_tests_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/gpu_tests_2gpu.log; tail -6 gpurun_out/gpu_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench exit $?"; tail -c 1800 gpurun_out/bench_n2.json
