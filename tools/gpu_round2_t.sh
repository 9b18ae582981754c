(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_gputests_t.log 2>&1; tail -4 gpurun_out/r02_gputests_t.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|plane_stage" --launch-skip 8 --launch-count 4 -f -o gpurun_out/r02_sweep_t python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full_t.log 2>&1; tail -1 gpurun_out/r02_ncu_full_t.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_512_t.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench_n1_t.json 2> gpurun_out/r02_bench_n1_t.err; head -c 900 gpurun_out/r02_bench_n1_t.json; echo
python -c "import __graft_entry__ as g; g.smoke()"
