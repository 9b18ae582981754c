(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_gputests_q.log 2>&1; tail -4 gpurun_out/r02_gputests_q.log
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('   ms/step %.3f  p50 %.3f min %.3f max %.3f  MLUPS %.0f  sweep %.3f ms  frac %.3f launches %s' % (d['ms_per_step'], d['ms_per_step_p50'], d['ms_per_step_min'], d['ms_per_step_max'], d['value'], r['kernel_ms'], r['frac'], d.get('gpu_launches')))
    elif 'rror' in ln: print(ln.strip()[:200])
"; }
run A=1
run PLB_B200_NO_HANDOFF=1
python tools/face_cost.py 512 none,z,xyz
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench_512_q.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
