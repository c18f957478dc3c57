#!/bin/bash
# --set full captures summarised ON THE BOX (raw + source pages as csv) so that only small files travel back.
mkdir -p gpurun_out /tmp/prof
TAG=${1:-r01b}
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 \
      -o /tmp/prof/$1 -f python tools/one_build.py > gpurun_out/ncu_$1_$TAG.log 2>&1
  ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/raw_$1_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof/$1.ncu-rep --page source --csv > gpurun_out/src_$1_$TAG.csv 2>/dev/null
  ls -la /tmp/prof/$1.ncu-rep
}
cap ospass 'k_os_pass' 14 2
cap lcp 'k_phi|k_plcp|k_lcp_gather' 3 3
cap induce 'k_induce' 4 2
cap misc 'k_name_flags|k_scan_apply|k_os_hist' 12 6
du -sh gpurun_out
