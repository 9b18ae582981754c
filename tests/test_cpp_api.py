"""The C++ mirror of the Palabos API (include/palabos_b200/palabos3D.h): examples written like
reference user code, compiled with g++ against the C-ABI, run on the GPU and compared with the
oracle and with the reference's own golden strings (jlabos/src/testing/test_cylinder.py:85-97)."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from cases import cavity, channel
from scenario import run_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


def _exe(name):
    path = os.path.join(EX, name)
    subprocess.run(["make", "-C", EX], check=True, capture_output=True)      # no-op when up to date; never run a stale example
    return path


def test_examples_build_against_the_header():
    """CPU-side: the API mirror compiles as C++11 and links against the library."""
    subprocess.run(["make", "-C", EX, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", EX], check=True, capture_output=True)
    assert os.path.exists(os.path.join(EX, "cavity3d")) and os.path.exists(os.path.join(EX, "channel3d"))
    assert os.path.exists(os.path.join(EX, "voxelize_stl")) and os.path.exists(os.path.join(EX, "external_rhoj3d"))
    assert os.path.exists(os.path.join(EX, "coprocessor3d"))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,q", [((20, 20, 20), 19), ((20, 31, 42), 19), ((9, 8, 10), 27)])
def test_cavity3d_cpp_matches_oracle(shape, q):
    nx, ny, nz = shape
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "f.bin")
        out = subprocess.run([_exe("cavity3d"), str(nx), str(ny), str(nz), "30", dump] + (["D3Q27"] if q == 27 else []),
                             check=True, capture_output=True, text=True).stdout
        got = np.fromfile(dump, dtype=np.float64).reshape(nx, ny, nz, q)
    want = run_oracle(cavity(nx, ny, nz, 30, q=q))
    assert np.array_equal(got, want["f_end"])
    energy = float(re.search(r"Av\. energy (\S+)", out).group(1))
    density = re.search(r"Av\. density (\S+)", out).group(1)
    assert density == "%e" % want["stats"][-1][5]
    assert abs(energy - want["stats"][-1][4]) <= 1e-6 * want["stats"][-1][4]
    if shape == (20, 20, 20):
        assert density == "9.999741e-01"            # the reference's golden string for n = 20
        assert abs(energy - 8.4917e-06) / 8.4917e-06 < 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize("q", [19, 27])
def test_cavity3d_cpp_with_a_block_matches_oracle(q):
    """CAVITY3D_BLOCK=1: a solid block in the cavity through the mirror's addVelocityBoundary* / addInternalVelocityEdge* /
    addInternalVelocityCorner* (the concave edges and corners)"""
    from cases import cavity_block
    nx, ny, nz = 13, 12, 11
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "f.bin")
        subprocess.run([_exe("cavity3d"), str(nx), str(ny), str(nz), "30", dump] + (["D3Q27"] if q == 27 else []),
                       check=True, capture_output=True, text=True, env=dict(os.environ, CAVITY3D_BLOCK="1"))
        got = np.fromfile(dump, dtype=np.float64).reshape(nx, ny, nz, q)
    want = run_oracle(cavity_block(nx, ny, nz, 30, q=q))
    assert np.array_equal(got, want["f_end"])


@pytest.mark.gpu
@pytest.mark.parametrize("model,prec", [("trt", "f64"), ("smago27", "f32"), ("incbgk", "f64")])
def test_channel3d_cpp_matches_oracle(model, prec):
    out = subprocess.run([_exe("channel3d"), "16", "12", "10", "40", model, prec], check=True, capture_output=True, text=True).stdout
    if model == "trt":
        sc = channel(16, 12, 10, 40)
    elif model == "incbgk":
        sc = channel(16, 12, 10, 40, dyn="INCBGK")
    else:
        sc = channel(16, 12, 10, 40, dyn="SMAGO", omega=1.9, q=27, cs=0.14, prec="f32")
    want = run_oracle(sc)["stats"][-1]
    stored = re.search(r"stored: energy (\S+) density (\S+)", out)
    rel = 1e-12 if prec == "f64" else 1e-6
    assert abs(float(stored.group(1)) - want[2]) <= rel * abs(want[2])
    assert abs(float(stored.group(2)) - want[1]) <= rel * abs(want[1])
    assert re.search(r"Av\. density (\S+)", out).group(1) == "%e" % want[5]


@pytest.mark.gpu
def test_voxelize_stl_cpp_matches_reference_matrix_and_oracle_flow():
    """examples/voxelize_stl.cpp (the reference's utility/stl/revoxelize.cpp up to the voxel matrix, then flow around it)"""
    from scenario import OracleEngine
    g = np.load(os.path.join(ROOT, "tests", "golden", "voxel_cylinder_40.npz"))
    with tempfile.TemporaryDirectory() as tmp:
        stl = os.path.join(tmp, "c.stl")
        with open(stl, "w") as f:             # ASCII: keeps the doubles of the fixture; facet normals as the triangles have them
            f.write("solid from_the_golden_fixture\n")
            for t in g["soup"]:
                n = np.cross(t[1] - t[0], t[2] - t[0])
                f.write(" facet normal %r %r %r\n  outer loop\n" % tuple(float(c) for c in n))
                for v in t:
                    f.write("   vertex %r %r %r\n" % tuple(float(c) for c in v))
                f.write("  endloop\n endfacet\n")
            f.write("endsolid\n")
        dump = os.path.join(tmp, "flags.bin")
        out = subprocess.run([_exe("voxelize_stl"), stl, "40", "0", dump], check=True, capture_output=True, text=True).stdout
        flags = np.fromfile(dump, dtype=np.int32).reshape(g["flags"].shape)
    assert np.array_equal(flags, g["flags"])
    assert "outside 85035 outerBorder 14034 inside 68408 innerBorder 12598" in out      # SURVEY.md 8c known answer
    nx, ny, nz = 2 * flags.shape[0], flags.shape[1], flags.shape[2]
    mask = np.zeros((nx, ny, nz), dtype=np.int32)
    mask[flags.shape[0] // 2: flags.shape[0] // 2 + flags.shape[0]] = flags
    o = OracleEngine(19, "f64", nx, ny, nz, "BGK", 1.7)
    o.set_periodicity(True, True, True)
    o.define_dynamics_mask(mask, 3, "BB", 1.0)
    o.define_dynamics_mask(mask, 4, "BB", 1.0)
    o.initialize_at_equilibrium((0, nx - 1, 0, ny - 1, 0, nz - 1), 1.0, (0.03, 0.0, 0.0))
    o.initialize()
    o.collide_and_stream(20)
    m = re.search(r"Av\. energy (\S+) Av\. density (\S+)", out)
    assert m.group(1) == "%e" % o.compute_average_energy() and m.group(2) == "%e" % o.compute_average_density()
    o.close()


@pytest.mark.gpu
def test_cavity3d_cpp_field_extraction_and_checkpoint():
    """computeDensity / computeVelocity / computeVelocityNorm / computeVelocityComponent and saveBinaryBlock of the C++ mirror"""
    import base64
    nx, ny, nz = 9, 8, 10
    with tempfile.TemporaryDirectory() as tmp:
        dump, fields = os.path.join(tmp, "f.bin"), os.path.join(tmp, "fields.bin")
        subprocess.run([_exe("cavity3d"), str(nx), str(ny), str(nz), "25", dump], check=True, capture_output=True, text=True,
                       env=dict(os.environ, CAVITY3D_FIELDS=fields))
        rec = np.fromfile(fields, dtype=np.float64).reshape(nx, ny, nz, 6)
        ckpt = open(fields + ".ckpt", "rb").read()
        vti = open(fields + ".vtk.vti", "rb").read()
    # VtkImageOutput3D: byte-identical to the file the unmodified reference writes for this cavity
    assert vti == np.load(os.path.join(ROOT, "tests", "golden", "vtk_cavity_9x8x10.npz"))["vti"].tobytes()
    want = run_oracle(cavity(nx, ny, nz, 25))
    assert np.array_equal(rec[..., :4], want["rho_u_end"])
    u = want["rho_u_end"][..., 1:]
    assert np.array_equal(rec[..., 4], np.sqrt(u[..., 0] * u[..., 0] + u[..., 1] * u[..., 1] + u[..., 2] * u[..., 2]))
    assert np.array_equal(rec[..., 5], u[..., 1])
    assert base64.b64decode(ckpt[12:]) == want["f_end"].tobytes()


@pytest.mark.gpu
def test_external_rhoj3d_cpp_matches_the_unmodified_reference():
    """examples/external_rhoj3d.cpp: the fork's own time loop (ExternalRhoJcollideAndStream3D + BoxRhoBarJfunctional3D,
    executeInternalProcessors + incrementTime) through the C++ mirror -- populations and rhoBar / j fields identical to the
    dump of the unmodified reference for the same program (tests/golden/ext_cavity_incbgk.npz)"""
    nx, ny, nz, steps = 10, 9, 8, 40
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ext_cavity_incbgk.npz"))
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "f.bin")
        out = subprocess.run([_exe("external_rhoj3d"), str(nx), str(ny), str(nz), str(steps), dump], check=True, capture_output=True, text=True).stdout
        raw = np.fromfile(dump, dtype=np.float64)
    cells = nx * ny * nz
    pops, fld = raw[:cells * 19].reshape(nx, ny, nz, 19), raw[cells * 19:].reshape(nx, ny, nz, 4)
    assert np.array_equal(pops, gold["f_end_full"])
    assert np.array_equal(fld, gold["rj_end_full"])
    assert re.search(r"Av\. energy (\S+)", out).group(1) == "%e" % gold["stats"][-1][4]
    assert re.search(r"Av\. density (\S+)", out).group(1) == "%e" % gold["stats"][-1][5]
    assert "iterations 40" in out


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["dynlocal", "equilibrium", "interp"])
def test_cavity3d_cpp_dynamics_based_boundary_managers(model):
    """createDynamicsBasedLocalBoundaryCondition3D / createEquilibriumBoundaryCondition3D of the C++ mirror against the dump of
    the unmodified reference (in this fork the two managers are the same for planes, outer edges and corners)"""
    nx, ny, nz = 10, 9, 8
    gold = np.load(os.path.join(ROOT, "tests", "golden", ("cavity_interp_10x9x8" if model == "interp" else "cavity_dynlocal_10x9x8") + ".npz"))
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "f.bin")
        subprocess.run([_exe("cavity3d"), str(nx), str(ny), str(nz), "40", dump], check=True, capture_output=True, text=True,
                       env=dict(os.environ, CAVITY3D_BC=model))
        got = np.fromfile(dump, dtype=np.float64).reshape(nx, ny, nz, 19)
    assert np.array_equal(got, gold["f_end_full"])


# ---- the reference's accelerator plug-in interface plb::CoProcessor3D<T> ------------------------------------


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_coprocessor_adapter_cpp_matches_oracle(prec):
    """plb::B200CoProcessor3D<T> (include/palabos_b200/coprocessor_adapter.h) through global::defaultCoProcessor3D<T>(): addDomain,
    send (whole block and a sub-domain), N x collideAndStream, receive -- against the oracle stepped from the same populations.
    The interface leaves the outer layer of a domain undefined (coProcessor3D.h:105-108); the engine is defined there too
    (on-site rule of the reference's own sweep) and agrees with the oracle on every cell."""
    from scenario import Scenario
    nx, ny, nz, steps, omega = 14, 11, 9, 15, 1.3
    dt = np.float64 if prec == "f64" else np.float32
    k = np.arange(nx * ny * nz * 19, dtype=np.float64)
    pops = (1e-3 * np.sin(0.37 * k) / 19.0).astype(dt).reshape(nx, ny, nz, 19)
    sub = (2e-3 * np.cos(0.11 * np.arange(2 * 3 * nz * 19, dtype=np.float64)) / 19.0).astype(dt).reshape(2, 3, nz, 19)
    pops[1:3, 1:4, :, :] = sub
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "f.bin")
        subprocess.run([_exe("coprocessor3d"), str(nx), str(ny), str(nz), repr(omega), str(steps), dump] + (["f32"] if prec == "f32" else []),
                       check=True, capture_output=True, text=True)
        got = np.fromfile(dump, dtype=dt).reshape(nx, ny, nz, 19)
    sc = Scenario(19, prec, nx, ny, nz, "BGK", float(dt(omega)))
    sc.upload(pops)
    sc.run(steps).stats().dump("f_end")
    want = run_oracle(sc)["f_end"]
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_b200_coprocessor_inside_the_unmodified_reference():
    """examples/ref_coprocessor.cpp: the REAL Palabos (oracle/_ref/libplb.a, built from /root/reference where that exists) with
    global::defaultCoProcessor3D<T>() answering with B200CoProcessor3D<T>: initiateCoProcessors -> addDomain per pure block,
    transferToCoProcessors / collideAndStream() / transferFromCoProcessors per step, compared inside the program bit for bit
    with the reference stepping the same lattice alone (periodic box: 8 of 8 blocks on the device; cavity: the centre block of
    27).  The binary travels to the GPU box prebuilt, like the reference driver."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_coprocessor")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_coprocessor not built (needs the reference sources)")
    r = subprocess.run([exe, "20"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "REF_COPROCESSOR_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "8 of 8 blocks on the co-processor" in r.stdout and "1 of 27 blocks on the co-processor" in r.stdout
