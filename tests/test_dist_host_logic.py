"""Host-side logic of the N > 1 path on CPU: two gloo ranks agree on the slab decomposition
(plb_b200_slab_range through the C-ABI) and the scatter / gather helpers used around the engine
round-trip a whole-lattice array.  No GPU work."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, nx, spr, ret):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import palabos_b200 as pb
    from dist_engine import gather_x, scatter_x, slab_bounds
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    bounds = slab_bounds(nx, world, spr, pb.slab_range)
    mine = [None] * world
    dist.all_gather_object(mine, bounds[rank])
    assert mine == bounds                                    # every rank computes the same decomposition
    assert bounds[0][0] == 0 and bounds[-1][1] == nx - 1
    for (a, b), (c, d) in zip(bounds[:-1], bounds[1:]):
        assert b + 1 == c and a <= b and c <= d              # contiguous, non-empty, no overlap
    sizes = [b - a + 1 for a, b in bounds]
    assert max(sizes) - min(sizes) <= spr                    # balanced
    rng = np.random.default_rng(10)                          # same seed on all ranks (plbInit.cpp:41-43 uses 10)
    whole = rng.random((nx, 3, 2, 19))
    back = gather_x(scatter_x(whole, bounds, rank), bounds)
    assert np.array_equal(back, whole)
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = 1


def _spawn(world, nx, spr, port):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, nx, spr, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == list(range(world))


def test_two_gloo_ranks_agree_on_slabs_and_roundtrip():
    _spawn(2, 11, 1, 29611)


def test_two_gloo_ranks_two_slabs_each():
    _spawn(2, 37, 2, 29612)


def test_slab_range_properties():
    sys.path.insert(0, os.path.dirname(HERE))
    import palabos_b200 as pb
    for nx, g in ((512, 8), (2048, 8), (11, 3), (7, 7), (4096, 8), (100, 6)):
        prev = -1
        for s in range(g):
            x0, x1 = pb.slab_range(nx, g, s)
            assert x0 == prev + 1 and x1 >= x0
            prev = x1
        assert prev == nx - 1
    try:
        pb.slab_range(4, 5, 0)
        assert False, "more slabs than planes must fail"
    except pb.PlbError:
        pass
