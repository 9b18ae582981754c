"""STL -> voxel flags on the device (plb_b200_voxelize) against the voxel matrices of the unmodified reference
(integer flags: bit-exact, BASELINE.json north_star) and, at sizes beyond the fixtures, against the CPU oracle."""
import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "voxel_*.npz")))
NAMES = [os.path.basename(p)[:-4] for p in GOLDEN]


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_device_voxelizer_reproduces_the_reference(path):
    import palabos_b200 as pb
    g = np.load(path)
    bw = int(g["params"][4])
    got = pb.voxelize(g["tri"], g["flags"].shape, bw)
    assert got.dtype == np.int32 and np.array_equal(got, g["flags"])


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_stl_soup_to_flags_through_the_c_abi(path):
    import palabos_b200 as pb
    g = np.load(path)
    res, rdir, margin, extra, bw = [int(x) for x in g["params"]]
    tri, dims, _, _ = pb.mesh_to_lattice(g["soup"], res, rdir, margin, extra)
    assert np.array_equal(pb.voxelize(tri, dims, bw), g["flags"])


def test_larger_resolutions_against_the_cpu_oracle_and_invariants():
    import palabos_b200 as pb
    import voxel_oracle as vo
    g = np.load(os.path.join(HERE, "golden", "voxel_ball_40.npz"))
    for res, rdir, margin, bw in ((97, 1, 3, 1), (160, 0, 4, 2)):
        tri, dims, _, _ = pb.mesh_to_lattice(g["soup"], res, rdir, margin)
        got = pb.voxelize(tri, dims, bw)
        assert np.array_equal(got, vo.voxelize(tri, dims, bw))
        inside = (got == vo.INSIDE) | (got == vo.INNER_BORDER)
        # a ball of diameter res: volume within 2 % of the sphere's, outermost layers outside
        assert abs(inside.sum() / (np.pi / 6 * res ** 3) - 1) < 0.02
        assert not inside[:margin - 1].any() and not inside[:, :, -(margin - 1):].any()
    # voxelising the mirrored mesh gives the mirrored matrix (the tie rule must not leak into the result)
    tri, dims, _, _ = pb.mesh_to_lattice(g["soup"], 61, 0, 4)
    a = pb.voxelize(tri, dims, 1)
    mir = tri.copy()
    mir[..., 2] = (dims[2] - 1) - mir[..., 2]
    b = pb.voxelize(mir[:, ::-1], dims, 1)
    assert np.array_equal(a[:, :, ::-1], b)


def test_voxel_matrix_drives_define_dynamics_mask():
    """the flags are what defineDynamics(lattice, mask, flag, BounceBack) takes: same run as with the oracle's mask"""
    import palabos_b200 as pb
    from scenario import OracleEngine
    g = np.load(os.path.join(HERE, "golden", "voxel_ball_23_z.npz"))
    tri, dims, _, _ = pb.mesh_to_lattice(g["soup"], 12, 0, 2)
    flags = pb.voxelize(tri, dims, 1)
    nx, ny, nz = dims[0] + 6, dims[1], dims[2]
    mask = np.ones((nx, ny, nz), dtype=np.int32)
    mask[3:3 + dims[0]] = flags
    engines = [pb.Lattice(19, "f64", nx, ny, nz, "BGK", 1.6), OracleEngine(19, "f64", nx, ny, nz, "BGK", 1.6)]
    for eng in engines:
        eng.set_periodicity(True, True, True)
        eng.define_dynamics_mask(mask, 3, "BB", 1.0)
        eng.define_dynamics_mask(mask, 4, "BB", 1.0)
        eng.initialize_at_equilibrium((0, nx - 1, 0, ny - 1, 0, nz - 1), 1.0, (0.04, 0.0, 0.01))
        eng.initialize()
        eng.collide_and_stream(12)
    assert np.array_equal(engines[0].download(), engines[1].download())
    for eng in engines:
        eng.close()
