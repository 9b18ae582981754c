"""saveBinaryBlock / loadBinaryBlock (io/serializerIO_3D.cpp:37-66): checkpoint files are
interchangeable with the reference's.  Fixtures: files written by the UNMODIFIED reference
(oracle/gen_checkpoint_golden.py) with the populations at saving time and some steps later."""
import base64
import glob
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "checkpoint_*.npz")))
NAMES = [os.path.basename(p)[:-4] for p in GOLDEN]


def _setup(pb, g):
    nx, ny, nz, q, size, before, after = [int(v) for v in g["params"]]
    dyn, omega, cs = str(g["setup"][0]), float(g["setup"][1]), float(g["setup"][2])
    lat = pb.Lattice(q, "f64" if size == 8 else "f32", nx, ny, nz, dyn, omega, cs)
    box = (0, nx - 1, 0, ny - 1, 0, nz - 1)
    lat.set_velocity_condition_on_block_boundaries("local", box, box)
    lat.set_boundary_velocity(box, (0, 0, 0))
    lat.set_boundary_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (0.02, 0.01, 0.0))
    lat.initialize_at_equilibrium(box, 1.0, (0, 0, 0))
    lat.initialize()
    return lat, before, after


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_reference_file_format_is_size_header_plus_wire_format(path):
    """CPU: the reference's file = base64(uint64 size) + base64(populations [x][y][z][q] as stored)"""
    g = np.load(path)
    blob = g["blob"].tobytes()
    size = struct.unpack("<Q", base64.b64decode(blob[:12]))[0]
    assert size == g["at_save"].nbytes
    assert base64.b64decode(blob[12:]) == g["at_save"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_restart_from_a_reference_checkpoint(path, tmp_path):
    import palabos_b200 as pb
    g = np.load(path)
    lat, before, after = _setup(pb, g)
    f = tmp_path / "ckpt.dat"
    f.write_bytes(g["blob"].tobytes())
    lat.load_binary_block(f)
    assert np.array_equal(lat.download(), g["at_save"])
    lat.collide_and_stream(after)
    assert np.array_equal(lat.download(), g["after"])
    lat.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_our_checkpoint_is_byte_identical_to_the_reference_file(path, tmp_path):
    import palabos_b200 as pb
    g = np.load(path)
    lat, before, after = _setup(pb, g)
    lat.collide_and_stream(before)
    f = tmp_path / "ours.dat"
    lat.save_binary_block(f)
    assert f.read_bytes() == g["blob"].tobytes()
    # and it reads back (odd number of steps: the in-place AA layout is swapped at this point)
    lat.collide_and_stream(1)
    lat.load_binary_block(f)
    assert np.array_equal(lat.download(), g["at_save"])
    with pytest.raises(pb.PlbError):
        lat.load_binary_block(tmp_path / "missing.dat")
    other = pb.Lattice(19, "f64", 3, 3, 3, "BGK", 1.0)
    with pytest.raises(pb.PlbError, match="bytes"):
        other.load_binary_block(f)                          # size header does not match this lattice
    other.close()
    lat.close()


@pytest.mark.gpu
def test_large_checkpoint_streams_in_chunks(tmp_path):
    """more than one 64 MB chunk, payload length not a multiple of 3: round trip"""
    import palabos_b200 as pb
    nx, ny, nz = 70, 103, 101            # 70*103*101*19*8 = 110.7 MB, planes of 1.58 MB
    lat = pb.Lattice(19, "f64", nx, ny, nz, "BGK", 1.5)
    lat.set_periodicity(True, True, True)
    lat.initialize_at_equilibrium((0, nx - 1, 0, ny - 1, 0, nz - 1), 1.0, (0.03, -0.01, 0.02))
    lat.define_dynamics((10, 20, 5, 50, 7, 60), "BB", 1.0)
    lat.initialize()
    lat.collide_and_stream(3)
    want = lat.download()
    f = tmp_path / "big.dat"
    lat.save_binary_block(f)
    assert os.path.getsize(f) == 12 + 4 * ((want.nbytes + 2) // 3)
    lat.collide_and_stream(2)
    lat.load_binary_block(f)
    assert np.array_equal(lat.download(), want)
    lat.close()
