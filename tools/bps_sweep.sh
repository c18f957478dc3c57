#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/bps.log
for b in 1 2 3 4; do
  echo "== BPS=$b" >> gpurun_out/bps.log
  B200SA_INDUCE_BPS=$b timeout 300 python tools/phase_times.py 100000000 2>&1 | grep -E '"input": "(dna|bytes)"' | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); p = d['phases_ms']
    print(d['input'], 'blocks', d['stats']['induce_blocks'], {k: v for k, v in p.items() if k.startswith('induce')}, 'sa_MBps', d['sa_MBps'])
" >> gpurun_out/bps.log
done
cat gpurun_out/bps.log
