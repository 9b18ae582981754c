run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('   ms/step %.3f  p50 %.3f min %.3f max %.3f  MLUPS %.0f  sweep %.3f ms  frac %.3f' % (d['ms_per_step'], d['ms_per_step_p50'], d['ms_per_step_min'], d['ms_per_step_max'], d['value'], r['kernel_ms'], r['frac']))
    elif 'rror' in ln: print(ln.strip()[:200])
"; }
run PLB_B200_EAGER_PLANES=1
run PLB_B200_EAGER_PLANES=1 PLB_B200_DUMMY_SMEM=81920
run PLB_B200_EAGER_PLANES=1 PLB_B200_DUMMY_SMEM=40000
for e in 0 1; do PLB_B200_EAGER_PLANES=$e timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warps_issue_stalled_barrier_per_warp_active.pct --clock-control none -k regex:sweep_kernel --launch-skip 4 --launch-count 4 --csv --log-file gpurun_out/r02_ncu_metrics_eager$e.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1; done
