#!/bin/bash
# Two-GPU run: bench line through dist.sharded_sample + the NCCL parity test.   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_n2.sh <tag>'
TAG=${1:-n2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
echo "bench n2 exit $?"; tail -c 400 gpurun_out/${TAG}_bench_n2.err
timeout 300 python -m pytest tests/test_gpu_dist_nccl.py -x -q 2>&1 | tail -2
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.1f sustained %.1f e2e %.1f collective %.3f / alone %.3f ms; gather %s" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["e2e"]["collective_ms"], d["e2e"]["collective_alone_ms"], d["e2e"]["gather_check"]))
PY
