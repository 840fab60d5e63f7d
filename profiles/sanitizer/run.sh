#!/bin/bash
# compute-sanitizer over the small golden cases in all histogram / filter variants.
#   gpurun -- bash profiles/sanitizer/run.sh        -> gpurun_out/sanitizer/<tool>.log
# racecheck: the survivor rings are single-producer / single-consumer queues on volatile shared words ordered by
# __threadfence_block, which racecheck cannot see through; its report is kept and explained in profiles/sanitizer/README.md.
cd "$(dirname "$0")/../.."
out=gpurun_out/sanitizer; mkdir -p $out
python profiles/sanitizer/run_cases.py > $out/plain.log 2>&1 || { echo "cases fail without the sanitizer"; tail -5 $out/plain.log; exit 1; }
for tool in memcheck synccheck initcheck racecheck; do
    arg=""; [ $tool = racecheck ] && arg="quick"
    extra=""; [ $tool = memcheck ] && extra="--leak-check no"
    [ $tool = initcheck ] && extra="--track-unused-memory no"
    timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool $extra --print-limit 40 python profiles/sanitizer/run_cases.py $arg > $out/$tool.log 2>&1
    echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' $out/$tool.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY\|sanitizer cases done' $out/$tool.log | tr '\n' ' ')"
done
