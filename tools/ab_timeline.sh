#!/bin/bash
# Same-box A/B of two library builds: regennet_b200/csrc/libregen_ab_old.so (A) vs libregen_sm100.so (B), alternating.
#   gpurun --timeout 900 -- 'bash tools/ab_timeline.sh <tag>'
TAG=${1:-ab}
mkdir -p gpurun_out
OLD=$PWD/regennet_b200/csrc/libregen_ab_old.so
for rep in 1 2; do
  for v in A B; do
    if [ $v = A ]; then export REGEN_LIB_PATH=$OLD; else unset REGEN_LIB_PATH; fi
    echo "=== $v rep $rep" >> gpurun_out/${TAG}_ab.txt
    timeout 200 python tools/step_timeline.py 2>&1 | grep -E "ms per step|n= *(2|16) " >> gpurun_out/${TAG}_ab.txt
    timeout 200 python bench.py --brief 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench value %.1f sustained %.1f e2e %.1f' % (d['value'], d['sustained']['value'], d['e2e']['value']))" >> gpurun_out/${TAG}_ab.txt
  done
done
cat gpurun_out/${TAG}_ab.txt
