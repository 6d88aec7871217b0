#!/bin/bash
mkdir -p gpurun_out
tools/fp32x2_probe.bin > gpurun_out/r02e_fp32x2.txt 2>&1; cat gpurun_out/r02e_fp32x2.txt
timeout 600 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_gemm.py tests/test_gpu_sampler.py tests/test_gpu_postprocess.py tests/test_rot2xyz.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --brief > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["sustained"]["value"], d["e2e"]["value"])
for k in ("roofline_hbm", "roofline_rot6d"):
    print(k, d[k]["achieved"], d[k]["frac"], d[k]["ms"])
PY
timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/r02e_racecheck_gemm.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02e_racecheck_gemm.log
