#!/bin/bash
# Round-2 profiling pass (one GPU): ncu launch list + one --set full capture of steady-state launches + summaries.
# The .ncu-rep goes to /tmp (it can exceed gpurun_out's 64 MiB merge limit); only the text / json summaries come back.
TAG=${2:-r02i}
mkdir -p gpurun_out
if [ "$1" = "launches" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python tools/ncu_target.py 4 > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches exit $?"
  python profiles/summarize.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.txt
fi
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'gemm|attention|p_sample_update|split_rows|rot6d' \
    -s 144 -c 16 -f -o /tmp/${TAG}_full python tools/ncu_target.py 4 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la /tmp/${TAG}_full.ncu-rep
python profiles/summarize.py full /tmp/${TAG}_full.ncu-rep gpurun_out/${TAG}_ncu_full.txt
python profiles/summarize.py traffic /tmp/${TAG}_full.ncu-rep gpurun_out/${TAG}_traffic.json
grep -E "^void|^layers|^regen|time_duration|tensor_cycles_active|dram__bytes" gpurun_out/${TAG}_ncu_full.txt | head -80
