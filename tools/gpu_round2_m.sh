(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r02_gputests_m.log 2>&1; tail -4 gpurun_out/r02_gputests_m.log
for w in cavity channel_trt obstacle les_q27_f64 les_q27_f32 ext_cavity; do timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_n1_m_$w.json 2>&1; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_m_$w.json') if l.startswith('{')][-1]); print('$w', round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms/step  sweep', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))"; done
