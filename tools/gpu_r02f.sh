#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stgcn.py tests/test_gpu_elementwise.py -x -q -m gpu -s 2>&1 | tail -12
timeout 600 python tools/stgcn_bench.py 1000 2>&1 | tail -3 | tee gpurun_out/r02f_stgcn_bench.txt
timeout 600 python bench.py --brief > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["sustained"]["value"], d["e2e"]["value"])
for k in ("roofline_hbm", "roofline_rot6d"):
    print(k, d[k]["achieved"], d[k]["frac"], d[k]["ms"])
PY
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report all --print-limit 6 python -m pytest "tests/test_gpu_gemm.py" -x -q -m gpu -k "test_gemm" > gpurun_out/r02f_racecheck_detail.log 2>&1; grep -c "hazard" gpurun_out/r02f_racecheck_detail.log; grep -B2 -A14 "hazard detected" gpurun_out/r02f_racecheck_detail.log | head -60
