"""Worker of tests/test_multi_gpu.py: run under torch.distributed.run with one rank per GPU.
Replays parity cases on the x-slab-decomposed engine (NCCL envelope exchange) and compares the
gathered populations with the CPU oracle on rank 0: they must be bit-identical to the one-block run
(the reference's own guarantee across decompositions, jlabos/src/testing/testsuite:3-7)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from cases import CASES  # noqa: E402
from dist_engine import DistEngine  # noqa: E402
from scenario import replay, run_oracle  # noqa: E402

NAMES = ["cavity_bgk_12", "cavity_zouhe_10x9x8", "channel_trt_16x12x10", "channel_smago_q27_f32",
         "obstacle_zouhe_24x12x10", "mixed_omega_11x8x9", "periodic_rlb_q27", "periodic_trt_f32", "inlet_outlet_local",
         "cavity_incbgk_10x9x8",
         # the external-moment sweep on an enveloped lattice (the envelope layer is part of the decomposition)
         "ext_cavity_incbgk", "ext_channel_trt", "ext_open_bgk", "ext_interleaved"]


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import palabos_b200 as pb
    rank, world = dist.get_rank(), dist.get_world_size()
    spr = int(os.environ.get("PLB_TEST_SLABS_PER_RANK", "1"))
    bad = []
    # both transports of the envelope exchange: direct stores into the neighbouring rank's slab (CUDA IPC), and the fallback
    for halo in ("direct", "nccl"):
        os.environ["PLB_B200_HALO"] = halo
        for name in NAMES:
            sc = CASES[name][0]()
            if sc.nx < world * spr:
                continue
            eng = DistEngine(pb, sc.q, sc.prec, sc.nx, sc.ny, sc.nz, sc.dyn, sc.p1, sc.p2, slabs_per_rank=spr, device=dev,
                             enveloped_periodic=sc.periodicity if sc.external else None)
            transport = eng.halo_transport()
            got = replay(sc, eng)
            eng.close()
            if rank == 0:
                want = run_oracle(sc)
                ok = all(np.array_equal(got[k], want[k]) for k in want if k not in ("stats", "time"))
                rel = 1e-12 if sc.prec == "f64" else 1e-6
                ok = ok and np.allclose(got["stats"], want["stats"], rtol=rel, atol=1e-300, equal_nan=True)
                ok = ok and (("direct stores" in transport) == (halo == "direct"))
                print("%-28s %d ranks x %d slabs [%s]: %s" % (name, world, spr, transport[:40], "identical" if ok else "DIFFERENT"), flush=True)
                if not ok:
                    bad.append(name + "/" + halo)
    flag = torch.tensor([len(bad)], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if int(flag.item()):
        raise SystemExit("multi-GPU parity failed: %s" % bad)
    if rank == 0:
        print("MGPU_OK")


if __name__ == "__main__":
    main()
