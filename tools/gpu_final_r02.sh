#!/bin/bash
# Round-2 closing run on one GPU box: full GPU suite, smoke, the bench line at HEAD, the reference arm.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final_r02.sh <tag>'
TAG=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "reference arm exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print("value %.1f sustained %.1f e2e %.1f gemm-frac %.4f whole-step frac %.4f" % (d["value"], d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"], d["tensor_frac_whole_step"]))
print({k: round(v["span_us"], 2) for k, v in d["kernels"].items()})
print("precision_ab", d.get("precision_ab"))
print("library", {k: d["library_baseline"][k] for k in ("ours_vs_fp32_eager", "ours_vs_tf32_eager", "ours_max_abs_err_vs_fp32_forward")})
print("other", {k: (round(v["ms_per_step"], 3), v.get("frac_of_bf16_sustained")) for k, v in d["other_configs"].items()})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
r = json.loads(open("gpurun_out/${TAG}_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["kind"], r.get("config1_cpu", {}).get("steps_per_s"))
PY
