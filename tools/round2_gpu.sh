#!/bin/bash
# One gpurun call: the ncu captures the profiles/ summaries of round 2 are made from (each ncu command right after the same
# command ran clean without ncu).  Outputs under gpurun_out/.
mkdir -p gpurun_out
python tools/profile_kernels.py > gpurun_out/r02_profile_plain.log 2>&1 || { tail -5 gpurun_out/r02_profile_plain.log; exit 1; }
# the report itself (~50 MB) stays on the box: gpurun brings back at most 64 MiB; the raw-metrics page is exported as CSV
PROFILE_ONCE=1 ncu --set full --clock-control none --import-source on -k regex:'k_step|k_rollout|k_qnet|k_gram|k_sample_grads|k_center' \
    -o /tmp/r02_kernels -f python tools/profile_kernels.py > gpurun_out/r02_ncu.log 2>&1
tail -3 gpurun_out/r02_ncu.log
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw.csv 2> gpurun_out/r02_ncu_export.err
ls -la gpurun_out/r02_ncu_raw.csv
SHORT="--steps 20 --warmup 3 --skip-gram --skip-config2 --skip-config4 --skip-variants --skip-cpu --e2e-steps 2"
python bench.py $SHORT > gpurun_out/r02_bench_short.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_config3.csv python bench.py $SHORT > gpurun_out/r02_ncu_launch.log 2>&1
tail -2 gpurun_out/r02_ncu_launch.log
