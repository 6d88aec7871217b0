#!/bin/bash
# Round-2 profiling pass (one GPU): ncu launch list + one --set full capture of a steady-state step + summaries.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python tools/ncu_target.py 4 > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -s 330 -c 60 -f -o gpurun_out/r02_full \
    python tools/ncu_target.py 4 > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/r02_full.ncu-rep
python profiles/summarize.py launches gpurun_out/r02_launches.csv gpurun_out/r02_launches.txt; head -30 gpurun_out/r02_launches.txt
