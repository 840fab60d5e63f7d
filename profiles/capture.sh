#!/bin/bash
# Run on the GPU box (under gpurun): bench without ncu, then the launch list and one --set full capture
# of pair_kernel.  usage: profiles/capture.sh <tag> [bench args]
tag=$1; shift
set -x
python bench.py --steps 6 --warmup 3 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu1_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu2_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
