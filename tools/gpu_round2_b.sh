set -x
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r02_gputests_b.log 2>&1; tail -25 gpurun_out/r02_gputests_b.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; head -c 2500 gpurun_out/r02_bench_n1_b.json; tail -3 gpurun_out/r02_bench_n1_b.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_512_b.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_launch_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-skip 8 --launch-count 2 -f -o gpurun_out/r02_sweep_b python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full_b.log 2>&1; tail -3 gpurun_out/r02_ncu_full_b.log
ls -la gpurun_out/
