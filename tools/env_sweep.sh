#!/bin/bash
# Same-box sweep of an environment switch over the step timeline + brief bench:
#   gpurun --timeout 1500 -- 'bash tools/env_sweep.sh <tag> VAR v1 v2 v3 ...'
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    export $VAR=$v
    echo "=== $VAR=$v rep $rep" >> gpurun_out/${TAG}_sweep.txt
    timeout 200 python tools/step_timeline.py 2>&1 | grep -E "ms per step|n= *(2|16) " >> gpurun_out/${TAG}_sweep.txt
    timeout 200 python bench.py --brief 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench value %.1f sustained %.1f e2e %.1f' % (d['value'], d['sustained']['value'], d['e2e']['value']))" >> gpurun_out/${TAG}_sweep.txt
  done
done
cat gpurun_out/${TAG}_sweep.txt
