mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02f.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['sa_only'], d['phase_ms'], d['roofline']['frac'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_r02f.json 2>> gpurun_out/bench_r02f.err; tail -c 700 gpurun_out/ref_r02f.json
tools/gpu_launches.sh r02k 100000000 dna | tail -24
tools/gpu_ncu_full.sh r02k classify "k_classify_fused" 1 1 | tail -2
