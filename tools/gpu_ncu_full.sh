#!/bin/bash
# ncu --set full of selected kernels of one build (reports stay on the box; summaries come back).
# usage: tools/gpu_ncu_full.sh TAG NAME REGEX SKIP COUNT [n] [kind]
mkdir -p gpurun_out /tmp/prof
TAG=$1; NAME=$2; RE=$3; SKIP=$4; CNT=$5; N=${6:-100000000}; KIND=${7:-dna}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT \
    -o /tmp/prof/$NAME -f python tools/one_build.py $N $KIND > gpurun_out/ncu_${NAME}_${TAG}.log 2>&1
echo "ncu $NAME exit $?"
python tools/ncu_summarize.py full /tmp/prof/$NAME.ncu-rep gpurun_out/ncu_${NAME}_${TAG}.json
(ncu -i /tmp/prof/$NAME.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_${NAME}_${TAG}.csv 2>/dev/null || ncu -i /tmp/prof/$NAME.ncu-rep --page source --csv > gpurun_out/src_${NAME}_${TAG}.csv 2>/dev/null)
ncu -i /tmp/prof/$NAME.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_stalls.py > gpurun_out/stalls_${NAME}_${TAG}.txt
ls -la /tmp/prof/$NAME.ncu-rep gpurun_out/src_${NAME}_${TAG}.csv
