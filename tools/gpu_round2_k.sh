python bench.py --n 128 --steps 20 --warmup 6 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
FC_BLOCK=1 python tools/face_cost.py 128 xyz
FC_BLOCK=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_fc_launches.csv python tools/face_cost.py 128 xyz
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_fc_launches_bench.csv python bench.py --n 128 --steps 8 --warmup 6 --no-e2e --no-cpu > /dev/null 2>&1
