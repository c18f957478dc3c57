#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp4.log
run() { echo "== $1" >> gpurun_out/exp4.log; env $2 timeout 300 python tools/phase_times.py 100000000 2>&1 | grep -E '"input": "(dna)"' | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); p = d['phases_ms']
    print(d['input'], 'blocks', d['stats']['induce_blocks'], {k: v for k, v in p.items() if k.startswith('induce')}, 'sa_MBps', d['sa_MBps'])
" >> gpurun_out/exp4.log; }
run base "X=1"
run minb3 "B200SA_LIB=$PWD/suffix_b200/libb200sa_minb3.so"
cat gpurun_out/exp4.log
