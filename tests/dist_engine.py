"""Multi-rank view of palabos_b200.Lattice for the tests: one process per GPU (torch.distributed for
the plumbing), the lattice decomposed into x-slabs, rank r owning slabs [r*spr, (r+1)*spr).  The
wrapper scatters / gathers whole-lattice arrays so that tests.scenario.replay() can drive it like a
single-block engine.  Test infrastructure only."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(nx, n_ranks, slabs_per_rank, slab_range):
    """x range [x0, x1] owned by each rank (its slabs are contiguous)."""
    g = n_ranks * slabs_per_rank
    out = []
    for r in range(n_ranks):
        lo = slab_range(nx, g, r * slabs_per_rank)[0]
        hi = slab_range(nx, g, (r + 1) * slabs_per_rank - 1)[1]
        out.append((lo, hi))
    return out


def scatter_x(arr, bounds, rank):
    lo, hi = bounds[rank]
    return np.ascontiguousarray(arr[lo:hi + 1])


def gather_x(local, bounds, device=None):
    """all-gather the per-rank x-slices of an [nx, ...] array (works with gloo and nccl)."""
    world = dist.get_world_size()
    rest = local.shape[1:]
    nmax = max(hi - lo + 1 for lo, hi in bounds)
    pad = np.zeros((nmax,) + rest, dtype=local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy()[:hi - lo + 1] for o, (lo, hi) in zip(outs, bounds)]
    return np.concatenate(parts, axis=0)


class DistEngine:
    """Same method names as palabos_b200.Lattice / OracleEngine, whole-lattice arrays in and out."""

    def __init__(self, pb, q, prec, nx, ny, nz, dyn, p1, p2=0.0, slabs_per_rank=1, device=None, enveloped_periodic=None):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device
        self.lat = pb.Lattice(q, prec, nx, ny, nz, dyn, p1, p2, rank=self.rank, n_ranks=self.world, slabs_per_rank=slabs_per_rank,
                              enveloped_periodic=enveloped_periodic)
        uid = [pb.comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=device)
        self.lat.comm_init(uid[0])
        # an enveloped lattice (external-moment sweep) is decomposed with its envelope layer: engine plane = user plane + off
        off = 1 if (enveloped_periodic is not None and not enveloped_periodic[0]) else 0
        eng = slab_bounds(nx + 2 * off, self.world, slabs_per_rank, pb.slab_range)
        self.bounds = [(max(lo - off, 0), min(hi - off, nx - 1)) for lo, hi in eng]
        self.nx, self.ny, self.nz, self.q = nx, ny, nz, q

    def _box_bounds(self, box):
        box = tuple(box or (0, self.nx - 1, 0, self.ny - 1, 0, self.nz - 1))
        # per rank: the x-range of the box it owns, relative to the box (empty: hi < lo)
        return box, [(max(lo, box[0]) - box[0], min(hi, box[1]) - box[0]) for lo, hi in self.bounds]

    def upload_rhobar_j(self, fld, box=None):
        box, bb = self._box_bounds(box)
        lo, hi = bb[self.rank]
        self.lat.upload_rhobar_j(np.ascontiguousarray(np.asarray(fld)[lo:hi + 1]) if hi >= lo else np.zeros((0, 4), dtype=self.lat.dtype), box)

    def download_rhobar_j(self, box=None):
        box, bb = self._box_bounds(box)
        lo, hi = bb[self.rank]
        local = self.lat.download_rhobar_j(box).reshape(max(hi - lo + 1, 0), box[3] - box[2] + 1, box[5] - box[4] + 1, 4)
        return gather_x(local, [(l, max(h, l - 1)) for l, h in bb], self.device)

    def __getattr__(self, name):          # every set-up call is replicated on all ranks unchanged
        return getattr(self.lat, name)

    def upload(self, pops):
        self.lat.upload(scatter_x(np.asarray(pops), self.bounds, self.rank))

    def download(self):
        lo, hi = self.bounds[self.rank]
        local = self.lat.download().reshape(hi - lo + 1, self.ny, self.nz, self.q)
        return gather_x(local, self.bounds, self.device)

    def compute_density_velocity(self):
        lo, hi = self.bounds[self.rank]
        local = self.lat.compute_density_velocity().reshape(hi - lo + 1, self.ny, self.nz, 4)
        return gather_x(local, self.bounds, self.device)

    def close(self):
        self.lat.close()
