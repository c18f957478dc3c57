mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_sharded.py -m "gpu" -q --tb=short -x 2>&1 | tail -15
timeout 300 python tools/positions_bench.py 100000000 1000000 2>&1 | tail -2 | tee gpurun_out/positions_bench.json
