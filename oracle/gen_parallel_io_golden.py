#!/usr/bin/env python
"""Golden files of parallelIO::save / saveState (io/multiBlockWriter3D.cpp:226-238, io/utilIO_3D.cpp:45-60)
written by the UNMODIFIED reference (oracle/_ref/ref_driver{,_f32}) for the checkpoint cavity, with several block structures.
Run in the build container; writes tests/golden/parallel_io_*.npz:
  plb, dat      uint8   the bytes of the reference's <name>.plb (XML) and <name>.dat (raw populations of the blocks)
  restart_xml   uint8   the file saveState writes next to them
  at_save, after  T [nx,ny,nz,q]  populations when saved / `after` steps later (the reference's own restart reproduces `after`)
  params        nx, ny, nz, q, bytes per value, steps before saving, steps after, blocks in x, y, z
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CASES = {   # name: (q, prec, nx, ny, nz, dynamics, omega, cSmago, steps before, steps after, blocks)
    "parallel_io_q19_f64_2x1x1": (19, "f64", 6, 5, 4, "BGK", 1.3, 0.0, 5, 4, (2, 1, 1)),
    "parallel_io_q19_f64_2x2x1": (19, "f64", 6, 5, 4, "BGK", 1.3, 0.0, 5, 4, (2, 2, 1)),
    "parallel_io_q27_f32_1x1x1": (27, "f32", 5, 4, 7, "SMAGO", 1.8, 0.14, 3, 5, (1, 1, 1)),
}


def script(q, prec, nx, ny, nz, dyn, omega, cs, before, after, blocks, restart):
    box = "0 %d 0 %d 0 %d" % (nx - 1, ny - 1, nz - 1)
    lines = ["blocks %d %d %d" % (blocks if not restart else (1, 1, 1)), "lattice D3Q%d %s %d %d %d %s %r %r" % (q, prec, nx, ny, nz, dyn, omega, cs),
             "bc_box local %s %s" % (box, box), "set_velocity %s 0.0 0.0 0.0" % box,
             "set_velocity 1 %d 1 %d %d %d 0.02 0.01 0.0" % (nx - 2, ny - 2, nz - 1, nz - 1),
             "init_eq %s 1.0 0.0 0.0 0.0" % box, "initialize"]
    if restart:     # into another block structure than the one that saved
        lines += ["load_state 0 restart.xml", "run %d" % after, "dump restarted.bin"]
    else:
        lines += ["run %d" % before, "save_state %d 0 restart.xml ckpt" % before, "dump at_save.bin", "run %d" % after, "dump after.bin"]
    return "\n".join(lines) + "\n"


def main():
    for name, (q, prec, nx, ny, nz, dyn, omega, cs, before, after, blocks) in CASES.items():
        exe = os.path.join(HERE, "_ref", "ref_driver" if prec == "f64" else "ref_driver_f32")
        dt = np.float64 if prec == "f64" else np.float32
        with tempfile.TemporaryDirectory() as tmp:
            for restart in (False, True):
                open(os.path.join(tmp, "s.txt"), "w").write(script(q, prec, nx, ny, nz, dyn, omega, cs, before, after, blocks, restart))
                subprocess.run([exe, "s.txt"], cwd=tmp, check=True, capture_output=True)
            rd = lambda f: np.fromfile(os.path.join(tmp, f), dtype=dt).reshape(nx, ny, nz, q)
            at_save, aft, restarted = rd("at_save.bin"), rd("after.bin"), rd("restarted.bin")
            assert np.array_equal(aft, restarted), "the reference's own restart does not reproduce its run"
            raw = lambda f: np.frombuffer(open(os.path.join(tmp, f), "rb").read(), dtype=np.uint8)
            base = "ckpt%08d_0" % before
            plb, dat, xml = raw(base + ".plb"), raw(base + ".dat"), raw("restart.xml")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), plb=plb, dat=dat, restart_xml=xml, at_save=at_save, after=aft,
                            params=np.array([nx, ny, nz, q, dt().itemsize, before, after] + list(blocks), dtype=np.int64),
                            setup=np.array([dyn, repr(omega), repr(cs)]))
        print(name, "plb", plb.size, "dat", dat.size, "xml", bytes(xml))


if __name__ == "__main__":
    sys.exit(main())
