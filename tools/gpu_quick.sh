#!/bin/bash
# quick GPU check: full GPU suite + brief bench line.   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --brief > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.1f sustained %.1f e2e %.1f gemm-frac %.4f" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"]))
print({k: round(v["span_us"], 2) for k, v in d["kernels"].items()})
for k in ("roofline_hbm", "roofline_rot6d"):
    print(k, round(d[k]["achieved"]), round(d[k]["frac"], 3), round(d[k]["ms"] * 1e3, 2), "us")
PY
