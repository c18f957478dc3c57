python -c "import __graft_entry__ as g; g.smoke()"
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
