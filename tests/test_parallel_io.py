"""parallelIO::save / load and saveState / loadState (io/multiBlockWriter3D.cpp:226-238, io/multiBlockReader3D.cpp:251-259,
io/utilIO_3D.cpp:45-76): the files are interchangeable with the reference's.  Fixtures: files written by
the UNMODIFIED reference (oracle/gen_parallel_io_golden.py) for several block structures, with the populations at saving time
and some steps later."""
import glob
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "parallel_io_*.npz")))
NAMES = [os.path.basename(p)[:-4] for p in GOLDEN]


def _components(plb):
    """(box, begin, end) of every block of a .plb, as the reference's readXmlSpec takes them"""
    offsets = [int(v) for v in re.search(r"<Offsets>([^<]*)</Offsets>", plb).group(1).split()]
    boxes = [[int(v) for v in m.split()] for m in re.findall(r"<Component[^>]*>([^<]*)</Component>", plb)]
    assert len(boxes) == len(offsets)
    return [(b, ([0] + offsets)[k], offsets[k]) for k, b in enumerate(boxes)]


def _setup(pb, g, slabs=1):
    nx, ny, nz, q, size, before, after = [int(v) for v in g["params"][:7]]
    dyn, omega, cs = str(g["setup"][0]), float(g["setup"][1]), float(g["setup"][2])
    lat = pb.Lattice(q, "f64" if size == 8 else "f32", nx, ny, nz, dyn, omega, cs, slabs_per_rank=slabs)
    box = (0, nx - 1, 0, ny - 1, 0, nz - 1)
    lat.set_velocity_condition_on_block_boundaries("local", box, box)
    lat.set_boundary_velocity(box, (0, 0, 0))
    lat.set_boundary_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (0.02, 0.01, 0.0))
    lat.initialize_at_equilibrium(box, 1.0, (0, 0, 0))
    lat.initialize()
    return lat, before, after


@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_reference_dat_is_the_blocks_back_to_back(path):
    """CPU: <name>.dat = the populations of every block's bulk, [x][y][z][q] as stored, at the offsets of the .plb"""
    g = np.load(path)
    plb, dat = g["plb"].tobytes().decode(), g["dat"].tobytes()
    for box, begin, end in _components(plb):
        want = g["at_save"][box[0]:box[1] + 1, box[2]:box[3] + 1, box[4]:box[5] + 1]
        assert dat[begin:end] == np.ascontiguousarray(want).tobytes()
    assert "<dynamicContent> False </dynamicContent>" in plb


@pytest.mark.gpu
@pytest.mark.parametrize("slabs", [1, 2])
@pytest.mark.parametrize("path", GOLDEN, ids=NAMES)
def test_restart_from_reference_files(path, slabs, tmp_path):
    """files of the reference, any block structure, into a lattice of 1 or 2 slabs: same populations, same continuation"""
    import palabos_b200 as pb
    g = np.load(path)
    lat, before, after = _setup(pb, g, slabs)
    (tmp_path / "ckpt.plb").write_bytes(g["plb"].tobytes().replace(b"ckpt%08d_0.dat" % before, b"ckpt.dat"))
    (tmp_path / "ckpt.dat").write_bytes(g["dat"].tobytes())
    lat.parallel_load(tmp_path / "ckpt")
    assert np.array_equal(lat.download(), g["at_save"])
    lat.collide_and_stream(after)
    assert np.array_equal(lat.download(), g["after"])
    lat.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", [p for p in GOLDEN if "2x2x1" not in p], ids=[n for n in NAMES if "2x2x1" not in n])
def test_our_files_match_the_reference_files(path, tmp_path):
    """same block structure (x-slabs): the .dat byte for byte, and the .plb line for line up to the sections that describe the
    set-up (dynamics dictionary, processors), which the reader treats as optional and this engine does not write"""
    import palabos_b200 as pb
    g = np.load(path)
    blocks = int(g["params"][7])
    lat, before, after = _setup(pb, g, blocks)
    lat.collide_and_stream(before)
    name = "ckpt%08d_0" % before
    lat.parallel_save(tmp_path / name, data_file_entry=name + ".dat")
    assert (tmp_path / (name + ".dat")).read_bytes() == g["dat"].tobytes()
    ours = (tmp_path / (name + ".plb")).read_text()
    ref = g["plb"].tobytes().decode()
    head = ref[:ref.index("        <DynamicsDict>")]
    assert ours == head + "    </Data>\n</Block3D>\n"
    # odd number of steps (the in-place layout is swapped now), then back
    lat.collide_and_stream(1)
    lat.parallel_load(tmp_path / name)
    assert np.array_equal(lat.download(), g["at_save"])
    # error behaviour: a missing file, a file of another lattice, a file with dynamic content
    with pytest.raises(pb.PlbError):
        lat.parallel_load(tmp_path / "missing")
    other = pb.Lattice(19, "f64", 3, 3, 3, "BGK", 1.0)
    with pytest.raises(pb.PlbError, match="match"):
        other.parallel_load(tmp_path / name)
    other.close()
    (tmp_path / "dyn.plb").write_text(ours.replace("<dynamicContent> False", "<dynamicContent> True"))
    with pytest.raises(pb.PlbError, match="dynamic content"):
        lat.parallel_load(tmp_path / "dyn")
    lat.close()


@pytest.mark.gpu
def test_the_reference_loads_our_files(tmp_path):
    """the unmodified reference (oracle/_ref/ref_driver) restarts from files this engine wrote with 2 slabs, into 2 blocks along y"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_driver not built")
    import palabos_b200 as pb
    g = np.load(os.path.join(HERE, "golden", "parallel_io_q19_f64_2x1x1.npz"))
    nx, ny, nz, q = [int(v) for v in g["params"][:4]]
    lat, before, after = _setup(pb, g, 2)
    lat.collide_and_stream(before)
    lat.parallel_save(tmp_path / "ours", data_file_entry="ours.dat")
    lat.close()
    box = "0 %d 0 %d 0 %d" % (nx - 1, ny - 1, nz - 1)
    script = ["blocks 1 2 1", "lattice D3Q19 f64 %d %d %d BGK 1.3 0.0" % (nx, ny, nz), "bc_box local %s %s" % (box, box),
              "set_velocity %s 0.0 0.0 0.0" % box, "set_velocity 1 %d 1 %d %d %d 0.02 0.01 0.0" % (nx - 2, ny - 2, nz - 1, nz - 1),
              "init_eq %s 1.0 0.0 0.0 0.0" % box, "initialize", "pload ours 0", "dump loaded.bin", "run %d" % after, "dump after.bin"]
    (tmp_path / "s.txt").write_text("\n".join(script) + "\n")
    subprocess.run([exe, "s.txt"], cwd=tmp_path, check=True, capture_output=True)
    rd = lambda f: np.fromfile(tmp_path / f, dtype=np.float64).reshape(nx, ny, nz, q)
    assert np.array_equal(rd("loaded.bin"), g["at_save"])
    assert np.array_equal(rd("after.bin"), g["after"])


@pytest.mark.gpu
@pytest.mark.parametrize("slabs", [1, 2])
def test_save_state_load_state_cpp(slabs, tmp_path):
    """examples/restart3d.cpp: saveState after 5 steps, 4 more; a second lattice restarted with loadState runs the same 4"""
    ex = os.path.join(ROOT, "examples")
    subprocess.run(["make", "-C", ex], check=True, capture_output=True)
    out = subprocess.run([os.path.join(ex, "restart3d"), str(tmp_path), str(slabs)], check=True, capture_output=True, text=True).stdout
    assert "restarted at iteration 5" in out
    g = np.load(os.path.join(HERE, "golden", "parallel_io_q19_f64_2x1x1.npz"))
    a = np.fromfile(tmp_path / "continued.bin", dtype=np.float64).reshape(g["after"].shape)
    b = np.fromfile(tmp_path / "restarted.bin", dtype=np.float64).reshape(g["after"].shape)
    assert np.array_equal(a, g["after"]) and np.array_equal(b, g["after"])
    # the files: the reference's, but for the directory in the names
    assert (tmp_path / "ckpt00000005_0.dat").read_bytes() == g["dat"].tobytes()
    want = g["restart_xml"].tobytes().decode().replace("./ckpt00000005", str(tmp_path) + "/ckpt00000005")
    assert (tmp_path / "restart.xml").read_text() == want
    ref = g["plb"].tobytes().decode()
    assert (tmp_path / "ckpt00000005_0.plb").read_text() == ref[:ref.index("        <DynamicsDict>")] + "    </Data>\n</Block3D>\n"


def test_file_name_helpers_of_the_mirror_match_the_reference(tmp_path):
    """CPU: plb::FileName / createFileName of include/palabos_b200/palabos3D.h against the output of the same probe compiled with the
    reference's io/plbFiles.h + io/imageWriter.h (tests/golden/filename_probe.txt, made in the build container with
    g++ -DUSE_REF -DPLB_USE_POSIX -I/root/reference/src ... tests/cpp/filename_probe.cpp oracle/_ref/libplb.a)"""
    exe = tmp_path / "probe"
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "cpp", "filename_probe.cpp"),
                    "-L" + os.path.join(ROOT, "palabos_b200"), "-lplb_b200", "-Wl,-rpath," + os.path.join(ROOT, "palabos_b200"), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out == open(os.path.join(HERE, "golden", "filename_probe.txt")).read()
