# dynamic piece assignment: grid-factor sweep (cavity, planes fused + drain launch), and the eager variant
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('   ms/step %.3f  p50 %.3f min %.3f max %.3f  MLUPS %.0f  sweep %.3f ms  frac %.3f' % (d['ms_per_step'], d['ms_per_step_p50'], d['ms_per_step_min'], d['ms_per_step_max'], d['value'], r['kernel_ms'], r['frac']))
    elif 'rror' in ln: print(ln.strip()[:200])
"; }
python -m pytest tests/test_gpu_parity.py -q -x -k "cavity_bgk or slab or bench_scale_matches and cavity" 2>&1 | tail -3
for ge in 1 2 4 8 16 32 64 128; do run PLB_B200_GRID_EVEN=$ge PLB_B200_GRID_ODD=$ge; done
run PLB_B200_GRID_EVEN=64 PLB_B200_GRID_ODD=128
run PLB_B200_EAGER_PLANES=1
run PLB_B200_EAGER_PLANES=1 PLB_B200_GRID_EVEN=4 PLB_B200_GRID_ODD=4
run PLB_B200_EAGER_PLANES=1 PLB_B200_GRID_EVEN=16 PLB_B200_GRID_ODD=16
