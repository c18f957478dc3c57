mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m "gpu" -q --tb=short -x --durations=3 > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -8 gpurun_out/gpu_tests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02g.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['sa_only'], d['phase_ms'], d['roofline']['frac'], d['cpu_baseline']['gpu_matches_oracle'], d['gpu_launches'])"
tools/gpu_launches.sh r02l 100000000 dna | tail -24
