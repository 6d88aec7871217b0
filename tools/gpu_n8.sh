#!/bin/bash
# Eight-GPU bench line (brief: no CPU / library baselines).   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_n8.sh <tag>'
TAG=${1:-n8}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --brief \
    > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
echo "bench n8 exit $?"; tail -c 300 gpurun_out/${TAG}_bench_n8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value %.1f sustained %.1f e2e %.1f collective %.3f / alone %.3f ms; gather %s" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["e2e"]["collective_ms"], d["e2e"]["collective_alone_ms"], d["e2e"]["gather_check"]))
PY
