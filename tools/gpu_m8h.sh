#!/bin/bash
# mixed8h bring-up on a GPU box: parity tests of both mixed8 schemes, then a same-box A/B (step timeline + brief bench).
#   gpurun --timeout 420 -- 'bash tools/gpu_m8h.sh <tag>'
TAG=${1:-m8h}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_mixed8.py -x -q -s 2>&1 | grep -vE "^$|warnings|Warning" | tail -40 | tee gpurun_out/${TAG}_pytest.log
: > gpurun_out/${TAG}_ab.txt
for prec in mixed8 mixed8h mixed8 mixed8h; do
  echo "=== $prec" >> gpurun_out/${TAG}_ab.txt
  REGEN_PRECISION=$prec timeout 100 python tools/step_timeline.py 2>&1 | grep -E "ms per step|n= *(2|16) " >> gpurun_out/${TAG}_ab.txt
done
for prec in mixed8 mixed8h; do
  echo "=== bench $prec" >> gpurun_out/${TAG}_ab.txt
  REGEN_PRECISION=$prec timeout 150 python bench.py --brief 2>gpurun_out/${TAG}_bench_$prec.err | tee gpurun_out/${TAG}_bench_$prec.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench value %.1f sustained %.1f e2e %.1f' % (d['value'], d['sustained']['value'], d['e2e']['value']))
print({k: round(v['span_us'], 2) for k, v in d['kernels'].items()})" >> gpurun_out/${TAG}_ab.txt
done
cat gpurun_out/${TAG}_ab.txt
