#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp2.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log; tail -3 gpurun_out/gpu_tests.log
run() { echo "== $1" >> gpurun_out/exp2.log; env $2 timeout 300 python tools/phase_times.py 100000000 2>&1 | grep -E '"input": "(dna|bytes)"' | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); p = d['phases_ms']; l = d['lcp_phases_ms']
    print(d['input'], {k: v for k, v in p.items() if k.startswith('induce')}, l, 'sa_MBps', d['sa_MBps'])
" >> gpurun_out/exp2.log; }
run base "X=1"
run g1_ldcg "B200SA_LIB=$PWD/suffix_b200/libb200sa_g1.so"
run g2_ldcs "B200SA_LIB=$PWD/suffix_b200/libb200sa_g2.so"
run g3_ldlu "B200SA_LIB=$PWD/suffix_b200/libb200sa_g3.so"
run bps2 "B200SA_INDUCE_BPS=2"
run bps4 "B200SA_INDUCE_BPS=4"
cat gpurun_out/exp2.log
