#!/bin/bash
# warp-stall sampling of k_rollout_ws at config 2 per SASS instruction (ncu source page), exported as CSV
mkdir -p gpurun_out
python tools/rollout_time.py > gpurun_out/rollout_plain.log 2>&1 || { tail -3 gpurun_out/rollout_plain.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:k_rollout_ws -c 1 -o /tmp/ro -f python tools/rollout_time.py > gpurun_out/ro_ncu.log 2>&1
ncu -i /tmp/ro.ncu-rep --page source --csv > gpurun_out/ro_source.csv 2> gpurun_out/ro_source.err
ncu -i /tmp/ro.ncu-rep --page raw --csv > gpurun_out/ro_raw.csv 2>> gpurun_out/ro_source.err
ls -la gpurun_out/ro_source.csv gpurun_out/ro_raw.csv
