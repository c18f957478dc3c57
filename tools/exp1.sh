#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp1.log
run() { echo "== $1" >> gpurun_out/exp1.log; env $2 timeout 300 python tools/phase_times.py 100000000 2>&1 | grep -E '"input": "dna"' | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); p = d['phases_ms']; l = d['lcp_phases_ms']
    print({k: v for k, v in p.items() if k.startswith('rsa_sort') or k in ('name','compact_lms')}, l, 'sa_MBps', d['sa_MBps'])
" >> gpurun_out/exp1.log; }
run base "X=1"
run l2fetch32 "B200SA_L2FETCH=32"
run l2fetch128 "B200SA_L2FETCH=128"
run minb4 "B200SA_LIB=$PWD/suffix_b200/libb200sa_minb4.so"
run minb5 "B200SA_LIB=$PWD/suffix_b200/libb200sa_minb5.so"
cat gpurun_out/exp1.log
