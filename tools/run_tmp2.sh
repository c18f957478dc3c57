mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_sharded.py -m "gpu" -q --tb=short -x -k "nccl" 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print(json.dumps(d.get('sharded'),indent=0)[:1600]); print(d['value'], d['e2e'])
PY
