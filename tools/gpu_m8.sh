#!/bin/bash
# mixed8 bring-up on a GPU box: parity tests, then a same-box A/B of the two operand schemes (step timeline + brief bench).
#   gpurun --timeout 900 -- 'bash tools/gpu_m8.sh <tag>'
TAG=${1:-m8}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mixed8.py -x -q -s 2>&1 | grep -vE "^$|warnings|Warning" | tail -25 | tee gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
  for prec in bf16x3 mixed8; do
    echo "=== $prec rep $rep" >> gpurun_out/${TAG}_ab.txt
    REGEN_PRECISION=$prec timeout 200 python tools/step_timeline.py 2>&1 | grep -E "ms per step|n= *(2|16) " >> gpurun_out/${TAG}_ab.txt
    REGEN_PRECISION=$prec timeout 300 python bench.py --brief 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench value %.1f sustained %.1f e2e %.1f' % (d['value'], d['sustained']['value'], d['e2e']['value']))" >> gpurun_out/${TAG}_ab.txt
  done
done
cat gpurun_out/${TAG}_ab.txt
