(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_gputests_l.log 2>&1; tail -6 gpurun_out/r02_gputests_l.log
python tools/face_cost.py 512
FC_BLOCK=1 python tools/face_cost.py 512 xyz
timeout 300 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('bench ms/step %.3f p50 %.3f  MLUPS %.0f  sweep %.3f frac %.3f launches %s' % (d['ms_per_step'], d['ms_per_step_p50'], d['value'], r['kernel_ms'], r['frac'], d.get('gpu_launches')))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench_512_l.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
