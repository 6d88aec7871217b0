#!/bin/bash
# First gpurun call of the next round: what this round could not run on a GPU any more.
#   gpurun --timeout 600 -- 'bash tools/first_call_next_round.sh'
# 1. the whole GPU suite with the xfail / xpass summary (tests/test_gpu_stgcn.py is non-strict xfail: XPASS = the ST-GCN
#    launch code works -> remove the marker);  2. the bench line (checks the roofline_attention object added last);
# 3. the other configs at HEAD.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -rxX > gpurun_out/next_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/next_pytest.log
timeout 300 python bench.py --gpus 1 > gpurun_out/next_bench_n1.json 2> gpurun_out/next_bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/next_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline_attention"])
PY
timeout 200 python tools/config_bench.py > gpurun_out/next_configs.txt 2>&1; cat gpurun_out/next_configs.txt
