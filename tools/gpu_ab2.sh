#!/bin/bash
# quick same-box check of one build: GPU parity tests of the kernels touched + step timeline + brief bench, for both operand schemes
#   gpurun --timeout 900 -- 'bash tools/gpu_ab2.sh <tag>'
TAG=${1:-ab2}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_attention_mc.py tests/test_gpu_denoiser.py tests/test_gpu_mixed8.py -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
for prec in bf16x3 mixed8; do
  echo "=== $prec" >> gpurun_out/${TAG}_ab.txt
  REGEN_PRECISION=$prec timeout 200 python tools/step_timeline.py 2>&1 | grep -E "ms per step|n= *(2|16) " >> gpurun_out/${TAG}_ab.txt
  REGEN_PRECISION=$prec timeout 300 python bench.py --brief 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench value %.1f sustained %.1f e2e %.1f' % (d['value'], d['sustained']['value'], d['e2e']['value']))" >> gpurun_out/${TAG}_ab.txt
done
cat gpurun_out/${TAG}_ab.txt
