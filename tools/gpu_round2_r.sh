(time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "handoff or survive or slab") > gpurun_out/r02_gputests_r.log 2>&1; tail -3 gpurun_out/r02_gputests_r.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_kernel|plane_stage" --launch-skip 8 --launch-count 4 -f -o gpurun_out/r02_sweep_r python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_ncu_full_r.log 2>&1; tail -1 gpurun_out/r02_ncu_full_r.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_512_r.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench_n1_r.json 2> gpurun_out/r02_bench_n1_r.err; head -c 1800 gpurun_out/r02_bench_n1_r.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_r.json 2> gpurun_out/r02_bench_ref_r.err; head -c 700 gpurun_out/r02_bench_ref_r.json; echo
for w in obstacle ext_cavity; do timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_n1_r_$w.json 2>&1; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_r_$w.json') if l.startswith('{')][-1]); print('$w', round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms/step  sweep', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))"; done
python -c "import __graft_entry__ as g; g.smoke()"
