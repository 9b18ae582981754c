"""STL -> voxel flags, CPU side: the numpy oracle against the voxel matrices dumped by
the UNMODIFIED reference (tests/golden/voxel_*.npz, oracle/gen_voxel_golden.py), the host-side mesh
preparation of the C-ABI (no device needed), and the STL reader."""
import glob
import os
import struct
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import voxel_oracle as vo  # noqa: E402

GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "voxel_*.npz")))
NAMES = [os.path.basename(p)[:-4] for p in GOLDEN]


def test_fixtures_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_oracle_reproduces_the_reference_voxel_matrix(path):
    g = np.load(path)
    bw = int(g["params"][4])
    got = vo.voxelize(g["tri"], g["flags"].shape, bw)
    assert got.dtype == g["flags"].dtype and np.array_equal(got, g["flags"])
    # known answers of SURVEY.md 8c (resolution 40, margin 4, borderWidth 1)
    if path.endswith("voxel_ball_40.npz"):
        assert np.bincount(got.ravel(), minlength=5).tolist() == [0, 76371, 7954, 26313, 7011]
    if path.endswith("voxel_cylinder_40.npz"):
        assert np.bincount(got.ravel(), minlength=5).tolist() == [0, 85035, 14034, 68408, 12598]


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_stl_to_flags_end_to_end_on_cpu(path):
    """own mesh preparation (scale, translate, inflate) + oracle voxelizer = the reference's matrix"""
    g = np.load(path)
    res, rdir, margin, extra, bw = [int(x) for x in g["params"]]
    tri, dims = vo.mesh_to_lattice(g["soup"], res, rdir, margin, extra)
    assert dims == g["flags"].shape
    # vertices: equal to the reference's up to the inflate amount (see the oracle's header: one-ring quirk)
    assert np.abs(tri - g["tri"]).max() < 1.0e-3
    assert np.array_equal(vo.voxelize(tri, dims, bw), g["flags"])


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_c_abi_mesh_preparation_matches_oracle_and_reference(path):
    import palabos_b200 as pb
    g = np.load(path)
    res, rdir, margin, extra, bw = [int(x) for x in g["params"]]
    tri, dims, dx, loc = pb.mesh_to_lattice(g["soup"], res, rdir, margin, extra)
    want, wdims = vo.mesh_to_lattice(g["soup"], res, rdir, margin, extra)
    assert dims == wdims == g["flags"].shape
    assert np.abs(tri - want).max() < 1e-12
    tri0, _, _, _ = pb.mesh_to_lattice(g["soup"], res, rdir, margin, extra, inflate=0.0)
    # the reference moved every vertex by the inflate amount (1e-3) from where we put it before inflating
    moved = np.linalg.norm((g["tri"] - tri0).reshape(-1, 3), axis=1)
    assert moved.min() > 0.99e-3 and moved.max() < 1.01e-3
    lo = g["soup"].astype(np.float64).reshape(-1, 3).min(axis=0)
    assert dx > 0 and np.allclose(np.array(loc) + (margin + extra) * dx, lo)


def test_stl_reader_binary_and_ascii(tmp_path):
    import palabos_b200 as pb
    tris = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0.5, 0.25, 1], [1, 2, 3], [-4, 5.5, 6]]], dtype=np.float32)
    b = tmp_path / "b.stl"
    with open(b, "wb") as f:
        f.write(b"binary stl written by the test".ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for t in tris:
            f.write(struct.pack("<3f", 0, 0, 0))
            f.write(t.astype("<f4").tobytes())
            f.write(b"\0\0")
    assert np.array_equal(pb.stl_read(b), tris)
    a = tmp_path / "a.stl"
    with open(a, "w") as f:
        f.write("solid s\n")
        for t in tris:
            f.write(" facet normal 0 0 0\n  outer loop\n")
            for v in t:
                f.write("   vertex %r %r %r\n" % tuple(float(c) for c in v))
            f.write("  endloop\n endfacet\n")
        f.write("endsolid s\n")
    assert np.array_equal(pb.stl_read(a), tris)
    with pytest.raises(pb.PlbError):
        pb.stl_read(tmp_path / "missing.stl")
