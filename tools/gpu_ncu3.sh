#!/bin/bash
mkdir -p gpurun_out /tmp/prof
TAG=${1:-r01d}
cap() {
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 \
      -o /tmp/prof/$1 -f python tools/one_build.py > gpurun_out/ncu_$1_$TAG.log 2>&1
  ncu -i /tmp/prof/$1.ncu-rep --page raw --csv > gpurun_out/raw_$1_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof/$1.ncu-rep --page source --csv > gpurun_out/src_$1_$TAG.csv 2>/dev/null
}
cap oshist 'k_os_hist' 0 2
cap ospass1 'k_os_pass' 0 2
cap grpscan 'k_scan_apply.*OutGroupRank' 0 1
python tools/ncu_raw.py gpurun_out/raw_oshist_$TAG.csv gpurun_out/raw_ospass1_$TAG.csv gpurun_out/raw_grpscan_$TAG.csv | grep -E "^---|time_duration|dram__bytes|issue_active|long_scoreboard|barrier|short_score|warps_active|registers"
