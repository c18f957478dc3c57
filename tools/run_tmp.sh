mkdir -p gpurun_out
timeout 280 python -m pytest tests -m "gpu and not slow" -q --tb=short -x > gpurun_out/gpu_tests_final.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests_final.log; tail -4 gpurun_out/gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()"
