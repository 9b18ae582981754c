# compute-sanitizer over small scenarios of the hot path: memcheck of the sweep / plane pass / handoff / slabs,
# racecheck of the shared-memory piece counter and statistics reduction
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import palabos_b200 as pb
from cases import CASES
from scenario import replay, run_oracle
for name, slabs in (("cavity_bgk_10x9x8", 1), ("cavity_bgk_10x9x8", 2), ("block_local_q27_12x12x11", 1), ("channel_smago_q27_f32", 3), ("obstacle_zouhe_24x12x10", 1), ("cavity_eraser_10x9x8", 2), ("ext_cavity_incbgk", 1), ("inlet_outlet_mixed_models", 1)):
    sc = CASES[name][0]()
    eng = pb.Lattice(sc.q, sc.prec, sc.nx, sc.ny, sc.nz, sc.dyn, sc.p1, sc.p2, slabs_per_rank=slabs,
                     enveloped_periodic=sc.periodicity if sc.external else None)
    got = replay(sc, eng)
    eng.close()
    want = run_oracle(sc)
    assert np.array_equal(got["f_end"], want["f_end"]), name
    print("ok", name, "slabs", slabs, flush=True)
PY
for tool in memcheck racecheck; do echo "== compute-sanitizer --tool $tool"; timeout 800 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py 2>&1 | grep -v "^$" | tail -12; done
