#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -rs > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r02b_pytest.log
grep -E "max abs err|stress|1000-step" gpurun_out/r02b_pytest.log | head
timeout 900 python -m pytest tests/test_gpu_parity_stress.py tests/test_rot2xyz.py -q -m gpu -s 2>&1 | grep -E "err|passed|failed" | head -20
timeout 900 python bench.py --gpus 1 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench exit $?"; tail -5 gpurun_out/r02b_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b_bench_n1.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "sustained", "e2e", "roofline", "kernels", "roofline_hbm", "roofline_rot6d", "roofline_attention", "breakdown_ms", "other_configs", "library_baseline", "cpu_baseline", "clocks"):
    print(k, json.dumps(d.get(k))[:1500])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err; echo "ref exit $?"; cat gpurun_out/r02b_bench_reference.json | cut -c1-1500
