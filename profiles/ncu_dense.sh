#!/bin/bash
# ncu --set full capture of the dense kernel on one config: profiles/ncu_dense.sh <cfg> <tag>
cfg=$1; tag=$2
ncu --set full --clock-control none --import-source on -k regex:'dense_kernel|count_kernel' -s 1 -c 1 -f -o gpurun_out/prof_$tag \
    python profiles/run_configs.py $cfg > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
