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
    # the triangles the reference hands to voxelize(): reproduced to 1e-10 lattice units (coordinates are ~10..70)
    assert np.abs(tri - g["tri"]).max() < 1.0e-10
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
    assert np.abs(tri - g["tri"]).max() < 1.0e-10          # = the unmodified reference's triangles
    tri0, _, _, _ = pb.mesh_to_lattice(g["soup"], res, rdir, margin, extra, inflate=0.0)
    moved = np.linalg.norm((tri - tri0).reshape(-1, 3), axis=1)
    assert np.abs(moved - 1e-3).max() < 1e-12              # inflate moves every vertex by exactly its amount
    lo = g["soup"].astype(np.float64).reshape(-1, 3).min(axis=0)
    assert dx > 0 and np.allclose(np.array(loc) + (margin + extra) * dx, lo)


def _write_binary_stl(path, tris, normals):
    with open(path, "wb") as f:
        f.write(b"binary stl written by the test".ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for t, n in zip(tris, normals):
            f.write(np.asarray(n, dtype="<f4").tobytes())
            f.write(np.asarray(t, dtype="<f4").tobytes())
            f.write(b"\0\0")


def _write_ascii_stl(path, tris, normals):
    with open(path, "w") as f:
        f.write("solid s\n")
        for t, n in zip(tris, normals):
            f.write(" facet normal %r %r %r\n  outer loop\n" % tuple(float(c) for c in n))
            for v in t:
                f.write("   vertex %r %r %r\n" % tuple(float(c) for c in v))
            f.write("  endloop\n endfacet\n")
        f.write("endsolid s\n")


def test_stl_reader_binary_and_ascii(tmp_path):
    """TriangleSet::readSTL: binary floats / ASCII doubles, facets re-oriented after their stated normal, degenerate ones dropped"""
    import palabos_b200 as pb
    tris = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]],            # normal +z, stated +z: kept as is
                     [[0, 0, 0], [1, 0, 0], [0, 1, 0]],            # stated -z: vertices 1, 2 swapped
                     [[0.5, 0.25, 1], [1, 2, 3], [-4, 5.5, 6]],
                     [[1, 1, 1], [1, 1, 1], [2, 3, 4]],            # vanishing edge: dropped
                     [[0, 0, 0], [1, 1, 1], [2, 2, 2]]],           # vanishing area: dropped
                    dtype=np.float64)
    tris[2] += 0.1                                                 # not representable in float: ASCII keeps the doubles
    normals = np.array([[0, 0, 1], [0, 0, -1], [0, 0, 0], [0, 0, 1], [0, 0, 1]], dtype=np.float64)
    want64 = np.stack([tris[0], tris[1][[0, 2, 1]], tris[2]])
    a = tmp_path / "a.stl"
    _write_ascii_stl(a, tris, normals)
    assert np.array_equal(pb.stl_read(a), want64)
    assert np.array_equal(vo.read_stl(a), want64)
    b = tmp_path / "b.stl"
    _write_binary_stl(b, tris, normals)
    t32 = tris.astype(np.float32).astype(np.float64)
    want32 = np.stack([t32[0], t32[1][[0, 2, 1]], t32[2]])
    assert np.array_equal(pb.stl_read(b), want32)
    assert np.array_equal(vo.read_stl(b), want32)
    with pytest.raises(pb.PlbError):
        pb.stl_read(tmp_path / "missing.stl")
