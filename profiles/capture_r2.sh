#!/bin/bash
# Round-2 capture on the GPU box (under gpurun): bench without ncu, the other configurations, then - only after each command
# has run clean - the ncu launch list of the bench and one `--set full` capture per pair kernel.
#   profiles/capture_r2.sh          -> gpurun_out/r2/
out=gpurun_out/r2; mkdir -p $out
set -x
timeout 300 python bench.py > $out/bench.json 2> $out/bench.err || exit 1
timeout 600 python profiles/run_configs.py cfg1 cfg2 cfg1_all cfg3 cfg3_g000 cfg3_dense cfg4a cfg4b cfg4c cfg5 cfg5_g000 cfg5_cells cfg5_cells_ring cfg5_g000_cells cfg4b_cells cfg3_cells cfg3_cells_ring cfg3_g000_cells > $out/configs.jsonl 2> $out/configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -f -o $out/prof_pair \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $out/ncu_pair.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:count_kernel -s 1 -c 1 -f -o $out/prof_count \
    python profiles/run_configs.py cfg1 > $out/ncu_count.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 1 -c 1 -f -o $out/prof_dense \
    python profiles/run_configs.py cfg2 > $out/ncu_dense.log 2>&1
ls -la $out
