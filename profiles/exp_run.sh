#!/bin/bash
# On the GPU box: cfg3 throughput of each experiment library built by exp_build.sh.
#   profiles/exp_run.sh name1 name2 ...     (env settings may be attached as name:VAR=1)
cd "$(dirname "$0")/.."
N=mdy-newanalysis-package_b200/newanalysis
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=${spec#*:}
  cp $N/exp/libgfn_$name.so $N/libgfn.so
  echo "== $spec"
  env $envs timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $BENCH_ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
