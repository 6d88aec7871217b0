timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02el2_pytest.log
for prec in bf16x3 mixed8; do
  echo "=== $prec" >> gpurun_out/r02el2_ab.txt
  REGEN_PRECISION=$prec timeout 200 python tools/step_timeline.py 2>&1 | grep -E "ms per step|n= *(2|16) " >> gpurun_out/r02el2_ab.txt
done
timeout 300 python tools/config_bench.py 2>&1 | tail -8 >> gpurun_out/r02el2_ab.txt
cat gpurun_out/r02el2_ab.txt
