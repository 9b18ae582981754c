#!/usr/bin/env python
"""Golden checkpoint files: saveBinaryBlock(lattice, file) written by the UNMODIFIED reference
(oracle/_ref/ref_driver{,_f32}, io/serializerIO_3D.cpp:37-49), with the populations at the moment of saving and after
some more steps.  Run in the build container; writes tests/golden/checkpoint_*.npz:
  blob     uint8   the bytes of the reference's checkpoint file
  at_save  T [nx,ny,nz,q]  populations when it was written
  after    T [nx,ny,nz,q]  populations `more` steps later (what a restart from the file must reproduce)
  params   nx, ny, nz, q, bytes per value, steps before saving, steps after
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CASES = {   # name: (q, prec, nx, ny, nz, dynamics, omega, cSmago, steps before, steps after)
    "checkpoint_cavity_q19_f64": (19, "f64", 6, 5, 4, "BGK", 1.3, 0.0, 5, 4),
    "checkpoint_cavity_q27_f32": (27, "f32", 5, 4, 7, "SMAGO", 1.8, 0.14, 3, 5),
}


def script(q, prec, nx, ny, nz, dyn, omega, cs, before, after, restart):
    box = "0 %d 0 %d 0 %d" % (nx - 1, ny - 1, nz - 1)
    lines = ["blocks 1 1 1", "lattice D3Q%d %s %d %d %d %s %r %r" % (q, prec, nx, ny, nz, dyn, omega, cs),
             "bc_box local %s %s" % (box, box), "set_velocity %s 0.0 0.0 0.0" % box,
             "set_velocity 1 %d 1 %d %d %d 0.02 0.01 0.0" % (nx - 2, ny - 2, nz - 1, nz - 1),
             "init_eq %s 1.0 0.0 0.0 0.0" % box, "initialize"]
    if restart:
        lines += ["load_block ckpt.dat", "run %d" % after, "dump restarted.bin"]
    else:
        lines += ["run %d" % before, "save_block ckpt.dat", "dump at_save.bin", "run %d" % after, "dump after.bin"]
    return "\n".join(lines) + "\n"


def main():
    for name, (q, prec, nx, ny, nz, dyn, omega, cs, before, after) in CASES.items():
        exe = os.path.join(HERE, "_ref", "ref_driver" if prec == "f64" else "ref_driver_f32")
        dt = np.float64 if prec == "f64" else np.float32
        with tempfile.TemporaryDirectory() as tmp:
            for restart in (False, True):
                open(os.path.join(tmp, "s.txt"), "w").write(script(q, prec, nx, ny, nz, dyn, omega, cs, before, after, restart))
                subprocess.run([exe, "s.txt"], cwd=tmp, check=True, capture_output=True)
            rd = lambda f: np.fromfile(os.path.join(tmp, f), dtype=dt).reshape(nx, ny, nz, q)
            at_save, aft, restarted = rd("at_save.bin"), rd("after.bin"), rd("restarted.bin")
            assert np.array_equal(aft, restarted), "the reference's own restart does not reproduce its run"
            blob = np.frombuffer(open(os.path.join(tmp, "ckpt.dat"), "rb").read(), dtype=np.uint8)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), blob=blob, at_save=at_save, after=aft,
                            params=np.array([nx, ny, nz, q, dt().itemsize, before, after], dtype=np.int64),
                            setup=np.array([dyn, repr(omega), repr(cs)]))
        print(name, "file bytes", blob.size, "head", bytes(blob[:24]))


def vtk_golden():
    """tests/golden/vtk_cavity_9x8x10.npz: the .vti file the reference writes for the cavity of examples/cavity3d.cpp
    (velocity norm, velocity, density as Float32; VtkImageOutput3D io/vtkDataOutput.hh)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import cavity
    from scenario import run_reference
    sc = cavity(9, 8, 10, 25, dumps=False)
    sc.ops.append(("ref_file", "vtk out 0.125 0.002", "out.vti", "vti"))
    ref = run_reference(sc)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vtk_cavity_9x8x10.npz"), vti=ref["vti"],
                        params=np.array([9, 8, 10, 25], dtype=np.int64), dx_dt=np.array([0.125, 0.002]))
    print("vtk_cavity_9x8x10", ref["vti"].size, "bytes;", bytes(ref["vti"][:120]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vtk":
        sys.exit(vtk_golden())
    sys.exit(main())
