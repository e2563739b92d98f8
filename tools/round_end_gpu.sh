#!/bin/bash
# One gpurun call at the end of a round: full GPU test suite, the bench line, the reference arm, and the ncu captures the
# profiles/ summaries are made from (each ncu command right after the same command ran clean without ncu).
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/f_pytest.log 2>&1
python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python tools/qnet_ncu.py > gpurun_out/f_qnet_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_qnet -s 2 -c 2 -o gpurun_out/r01_qnet17 python tools/qnet_ncu.py > gpurun_out/f_ncu_qnet.log 2>&1
SHORT="--steps 20 --warmup 3 --skip-gram --skip-config2 --skip-config4 --skip-variants --e2e-steps 2"
python bench.py $SHORT > gpurun_out/f_bench_short.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py $SHORT > gpurun_out/f_ncu_launch.log 2>&1
python bench.py $SHORT > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 3 -o gpurun_out/r01_k_step python bench.py $SHORT > gpurun_out/f_ncu_kstep.log 2>&1
tail -3 gpurun_out/f_pytest.log; cut -c1-300 gpurun_out/f_bench_n1.json; tail -n 2 gpurun_out/f_ncu_qnet.log; tail -n 2 gpurun_out/f_ncu_kstep.log
