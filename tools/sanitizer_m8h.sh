#!/bin/bash
# compute-sanitizer memcheck over the precision='mixed8h' kernels (GPU box, bounded to ~2 minutes):
#   gpurun --timeout 200 -- 'bash tools/sanitizer_m8h.sh'
# pair-GEMM mixed8 main loop unit tests + whole forwards on the fused route (H8 residual-stream pack in the input projection
# and both fused GEMM+LayerNorm kernels, mixed8 main loop in QKV / FFN1 / output projection)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
LOG=gpurun_out/r02j_sanitizer_memcheck_m8h.log
echo "=== memcheck: tests/test_gpu_gemm.py -k mixed8" > $LOG
timeout 60 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_gemm.py -k mixed8 -x -q -m gpu >> $LOG 2>&1
echo "exit $?" >> $LOG
echo "=== memcheck: tests/test_gpu_mixed8.py forwards with precision='mixed8h' (B=256 T=60, B=48 T=196)" >> $LOG
timeout 110 $CS --tool memcheck --print-limit 20 python -m pytest \
  "tests/test_gpu_mixed8.py::test_forward_matches_oracle_and_bf16x3[256-60-mixed8h]" \
  "tests/test_gpu_mixed8.py::test_forward_matches_oracle_and_bf16x3[48-196-mixed8h]" -x -q -m gpu >> $LOG 2>&1
echo "exit $?" >> $LOG
grep -E "^=== |ERROR SUMMARY|passed|failed|exit " $LOG
