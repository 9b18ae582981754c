(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_gputests_p.log 2>&1; tail -4 gpurun_out/r02_gputests_p.log
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('   ms/step %.3f  p50 %.3f min %.3f max %.3f  MLUPS %.0f  sweep %.3f ms  frac %.3f launches %s' % (d['ms_per_step'], d['ms_per_step_p50'], d['ms_per_step_min'], d['ms_per_step_max'], d['value'], r['kernel_ms'], r['frac'], d.get('gpu_launches')))
    elif 'rror' in ln: print(ln.strip()[:200])
"; }
run A=1
run PLB_B200_SERIAL_STAGES=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_512_p.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_launch_p.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|plane_stage" --launch-skip 8 --launch-count 4 -f -o gpurun_out/r02_sweep_p python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full_p.log 2>&1; tail -2 gpurun_out/r02_ncu_full_p.log
