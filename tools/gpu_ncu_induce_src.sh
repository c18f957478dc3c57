#!/bin/bash
# source-level stall profile of one k_induce launch (L pass of the final induce), summarised on the box
mkdir -p gpurun_out /tmp/prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_induce' -s 6 -c 1 \
    -o /tmp/prof/ind1 -f python tools/one_build.py > gpurun_out/ncu_ind1.log 2>&1
ncu -i /tmp/prof/ind1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/src_induce_final.csv 2>/dev/null || \
ncu -i /tmp/prof/ind1.ncu-rep --page source --csv > gpurun_out/src_induce_final.csv 2>/dev/null
ncu -i /tmp/prof/ind1.ncu-rep --page raw --csv > gpurun_out/raw_induce_final.csv 2>/dev/null
wc -l gpurun_out/src_induce_final.csv; head -3 gpurun_out/src_induce_final.csv | cut -c1-400
