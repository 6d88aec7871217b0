#!/bin/bash
# One GPU-box round: parity tests, smoke, bench (ours), ncu launch list and one ncu --set full capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [full]'
# Everything worth keeping goes to gpurun_out/ (merged back into the working tree).
TAG=${1:-r01}
FULL=${2:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --gpus 1 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench_n1.json
if [ "$FULL" = "full" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
      > gpurun_out/${TAG}_ncu_launch.log 2>&1
  echo "ncu launches exit $?"
  # skip setup + warm-up launches, then capture ~two layers of one step
  timeout 900 ncu --set full --clock-control none --import-source on -s 200 -c 14 -f \
      -o gpurun_out/${TAG}_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
      > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit $?"
  ls -la gpurun_out/
fi
