# usage: bash tools/gpu_round2_mgpu.sh N   (multi-GPU lines of round 2; run under gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
if [ -z "$SKIP_TESTS" ]; then (time timeout 900 python -m pytest tests/test_multi_gpu.py -q -x) > gpurun_out/r02_mgpu_tests_n$N.log 2>&1; tail -4 gpurun_out/r02_mgpu_tests_n$N.log; fi
line() { python -c "
import json,sys
d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); r=d['roofline']; e=d.get('e2e') or {}
print('$1', round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms/step p50', round(d.get('ms_per_step_p50',0),3), 'max', round(d.get('ms_per_step_max',0),3), 'sweep', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'parity', d.get('parity_check'), 'e2e', e.get('value'))"; }
timeout 900 $TR bench.py --gpus $N --steps 200 --warmup 10 --no-cpu > gpurun_out/r02_bench_n${N}_cavity.json 2> gpurun_out/r02_bench_n${N}_cavity.err; line gpurun_out/r02_bench_n${N}_cavity.json
for w in ${WORKLOADS:-channel_trt les_q27_f32 les_q27_f64}; do timeout 900 $TR bench.py --gpus $N --workload $w --steps 30 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_n${N}_$w.json 2> gpurun_out/r02_bench_n${N}_$w.err; line gpurun_out/r02_bench_n${N}_$w.json; done
which mpicxx mpirun mpiexec 2>&1 | head -3
