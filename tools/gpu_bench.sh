#!/bin/bash
# bench + ncu evidence, one GPU.  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python bench.py --gpus 1 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/ref_$TAG.json
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches_$TAG.csv
# full captures: the induce kernels of the second build, then a few streaming kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_induce -s 4 -c 4 \
    -o gpurun_out/prof_induce_$TAG -f python tools/one_build.py > gpurun_out/ncu_induce_$TAG.log 2>&1
echo "ncu induce exit $?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_cls_types|k_cls_block_state|k_radix_scatter|k_lcp_kasai|k_isa|k_name_flags' -s 40 -c 24 \
    -o gpurun_out/prof_stream_$TAG -f python tools/one_build.py > gpurun_out/ncu_stream_$TAG.log 2>&1
echo "ncu stream exit $?"
ls -la gpurun_out/
