mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02h.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['sa_only'], d['phase_ms'], d['roofline']['frac'], d['cpu_baseline']['gpu_matches_oracle'], d['cpu_baseline']['gpu_matches_oracle_e2e'], d['gpu_launches'], d['clocks'])"
