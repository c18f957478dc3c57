#!/bin/bash
mkdir -p gpurun_out
for k in 4 3 2; do
  echo "== KGRAM=$k" >> gpurun_out/kgram.log
  B200SA_TRACE=1 B200SA_KGRAM=$k timeout 300 python tools/phase_times.py 100000000 >> gpurun_out/kgram.log 2>&1
done
grep -E "KGRAM|doubling round|\"input\": \"dna\"|english" gpurun_out/kgram.log | cut -c1-900
