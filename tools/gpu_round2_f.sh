set -x
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r02_gputests_f.log 2>&1; tail -8 gpurun_out/r02_gputests_f.log
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('   ms/step %.3f  p50 %.3f min %.3f max %.3f  MLUPS %.0f  sweep %.3f ms  frac %.3f' % (d['ms_per_step'], d['ms_per_step_p50'], d['ms_per_step_min'], d['ms_per_step_max'], d['value'], r['kernel_ms'], r['frac']))
    elif 'rror' in ln: print(ln.strip()[:200])
"; }
{ for ge in 16 32 64; do for go in 32 64 128; do run PLB_B200_GRID_EVEN=$ge PLB_B200_GRID_ODD=$go; done; done; run PLB_B200_EAGER_PLANES=1; } > gpurun_out/r02_gridfactor_mailbox.log 2>&1; cat gpurun_out/r02_gridfactor_mailbox.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err; head -c 1500 gpurun_out/r02_bench_n1_f.json; tail -3 gpurun_out/r02_bench_n1_f.err
for w in channel_trt obstacle les_q27_f64 les_q27_f32 ext_cavity; do timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_n1_f_$w.json 2>&1; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_f_$w.json') if l.startswith('{')][-1]); print('$w', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_512_f.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_launch_f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel --launch-skip 8 --launch-count 4 -f -o gpurun_out/r02_sweep_f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full_f.log 2>&1; tail -3 gpurun_out/r02_ncu_full_f.log
ls -la gpurun_out/
