#!/bin/bash
# runs bench.py (value + roofline only) of every variant under tools/bin/ given on the command line
for v in "$@"; do
  ( cd tools/bin/$v && python bench.py --no-cpu --no-e2e --steps 60 --warmup 6 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v', 'MLUPS %.0f' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'sweep ms %.3f' % r['kernel_ms'], 'frac %.3f' % r['frac'])" )
done
