#!/bin/bash
# On the GPU box: cfg3 throughput of each experiment library built by exp_build.sh.
#   profiles/exp_run.sh name1 name2 ...     (env settings may be attached as name:VAR=1)
cd "$(dirname "$0")/.."
N=$PWD/mdy-newanalysis-package_b200/newanalysis
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=${spec#*:}
  echo "== $spec"
  env $envs LD_LIBRARY_PATH=$N/exp/$name timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs $BENCH_ARGS > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  echo "rc=$?"; tail -c 300 gpurun_out/exp_$name.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/exp_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['parity']['g000_exact'])" || echo "no result (hang or crash)"
done
