# grid-factor sweep of the stash kernel (cavity, planes fused) + the plane-free workloads with the new loop structure
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('   ms/step %.3f  p50 %.3f min %.3f max %.3f  MLUPS %.0f  sweep %.3f ms  frac %.3f' % (d['ms_per_step'], d['ms_per_step_p50'], d['ms_per_step_min'], d['ms_per_step_max'], d['value'], r['kernel_ms'], r['frac']))
    elif 'rror' in ln: print(ln.strip()[:200])
"; }
for ge in 1 2 4 8 16 64; do run PLB_B200_GRID_EVEN=$ge PLB_B200_GRID_ODD=$ge; done
run PLB_B200_GRID_EVEN=64 PLB_B200_GRID_ODD=128
run PLB_B200_GRID_EVEN=4 PLB_B200_GRID_ODD=8
run PLB_B200_EAGER_PLANES=1
run PLB_B200_EAGER_PLANES=1 PLB_B200_GRID_EVEN=4 PLB_B200_GRID_ODD=4
for w in channel_trt obstacle; do echo "== $w"; timeout 300 python bench.py --workload $w --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('   ms/step %.3f  MLUPS %.0f  sweep %.3f ms  frac %.3f' % (d['ms_per_step'], d['value'], r['kernel_ms'], r['frac']))
"; done
