mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/f_pytest.log 2>&1
python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python tools/qnet_ncu.py > gpurun_out/f_qnet_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_qnet -s 2 -c 2 -o gpurun_out/r01_qnet17 python tools/qnet_ncu.py > gpurun_out/f_ncu_qnet.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench_short.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/f_ncu_launch.log 2>&1
tail -3 gpurun_out/f_pytest.log; cut -c1-400 gpurun_out/f_bench_n1.json; tail -2 gpurun_out/f_ncu_qnet.log
