# 1 GPU: parity subset on the drain-launch kernel, bench, launch list, ncu --set full of one even + one odd sweep and their drain launches
set -x
(time timeout 1200 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_gputests_d.log 2>&1; tail -8 gpurun_out/r02_gputests_d.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; head -c 1800 gpurun_out/r02_bench_n1_d.json; tail -3 gpurun_out/r02_bench_n1_d.err
for w in channel_trt obstacle les_q27_f64 les_q27_f32; do timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_n1_d_$w.json 2>&1; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_d_$w.json') if l.startswith('{')][-1]); print('$w', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_512_d.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_launch_d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-skip 8 --launch-count 4 -f -o gpurun_out/r02_sweep_d python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full_d.log 2>&1; tail -3 gpurun_out/r02_ncu_full_d.log
ls -la gpurun_out/
