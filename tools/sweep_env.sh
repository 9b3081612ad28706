#!/bin/bash
# usage: tools/sweep_env.sh "VAR=a VAR=b ..." [bench args]  -- run bench.py once per setting, print value / e2e / kernel ms
settings="$1"; shift
for v in $settings; do
  echo "== $v"
  env $v timeout 600 python bench.py --no-cpu-baseline "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['launch_ms'],4), round(d['e2e']['ms_per_step'],3))"
done
