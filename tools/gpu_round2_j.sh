set -x
PLB_B200_EAGER_PLANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"plane_stage_kernel|fd_stage_kernel" --launch-skip 8 --launch-count 4 -f -o gpurun_out/r02_planes_j python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full_j.log 2>&1; tail -3 gpurun_out/r02_ncu_full_j.log
PLB_B200_EAGER_PLANES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench_512_j_eager.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
PLB_B200_EAGER_PLANES=1 timeout 300 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | tail -1 | cut -c1-600
