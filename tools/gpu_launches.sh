#!/bin/bash
# ncu launch list (device time per kernel launch; cold-cache, serialised) of one build+lcp.
# usage: tools/gpu_launches.sh TAG [n] [kind]
mkdir -p gpurun_out
TAG=${1:-r02}; N=${2:-100000000}; KIND=${3:-dna}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${TAG}_${KIND}.csv python tools/one_build.py $N $KIND > gpurun_out/ncu_launch_${TAG}_${KIND}.log 2>&1
echo "ncu exit $?"
python tools/ncu_summarize.py launches gpurun_out/launches_${TAG}_${KIND}.csv gpurun_out/launches_${TAG}_${KIND}.txt
head -40 gpurun_out/launches_${TAG}_${KIND}.txt
