#!/bin/bash
# compute-sanitizer memcheck + racecheck over the tcgen05 / TMA / mbarrier kernels (GPU box):
#   gpurun --timeout 1500 -- 'bash tools/sanitizer_run.sh'
# -> gpurun_out/r02b_sanitizer_{memcheck,racecheck}.log (summaries are copied to profiles/ by hand)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
TESTS="tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_attention_mc.py tests/test_gpu_elementwise.py"
FUSED='tests/test_gpu_denoiser.py::test_forward_matches_oracle_various_sizes tests/test_gpu_mixed8.py::test_forward_matches_oracle_and_bf16x3 tests/test_gpu_mixed8.py::test_offline_arch_stays_close_to_bf16x3'
for tool in memcheck racecheck; do
  echo "=== $tool: kernel unit tests" > gpurun_out/r02b_sanitizer_$tool.log
  timeout 600 $CS --tool $tool --print-limit 20 python -m pytest $TESTS -x -q -m gpu >> gpurun_out/r02b_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02b_sanitizer_$tool.log
  echo "=== $tool: fused-route forward (B=256, T=60) + small batches + precision='mixed8' (all three attention kernels, offline arch)" >> gpurun_out/r02b_sanitizer_$tool.log
  timeout 600 $CS --tool $tool --print-limit 20 python -m pytest $FUSED -x -q -m gpu >> gpurun_out/r02b_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02b_sanitizer_$tool.log
  grep -E "^=== |ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " gpurun_out/r02b_sanitizer_$tool.log
done
