#!/bin/bash
# quick device-resident throughput of cfg3 for a list of env settings: profiles/quick.sh "A=1" "A=2 B=3" ...
for e in "$@"; do
  echo "== $e"
  env $e python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
