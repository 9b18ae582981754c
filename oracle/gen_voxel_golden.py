#!/usr/bin/env python
"""Golden vectors of the STL -> voxel-flag pipeline, produced by the UNMODIFIED
reference through oracle/_ref/ref_voxel (utility/stl/revoxelize.cpp:40-80 restated as a driver).
Run in the build container (needs /root/reference); writes tests/golden/voxel_*.npz:
  soup    float32 [nt,3,3]  the triangles of the STL file as stored in it (input of our own mesh pipeline)
  tri     float64 [nt',3,3] the triangles the reference hands to voxelize(): lattice units, after
                            DEFscaledMesh scaling/translation and TriangularSurfaceMesh::inflate()
  flags   int8 [nx,ny,nz]   voxelFlag values (offLattice/voxelizer.h:38-48) after detectBorderLine
  params  int64 [5]         resolution, referenceDirection, margin, extraLayer, borderWidth
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STL_DIR = "/root/reference/viscosityTest/stl"
CASES = {   # name: (file, resolution, referenceDirection, margin, extraLayer, borderWidth)
    "voxel_ball_40": ("ball.stl", 40, 0, 4, 0, 1),
    "voxel_ball_23_z": ("ball.stl", 23, 2, 3, 1, 2),
    "voxel_cylinder_40": ("cylinderSimple.stl", 40, 0, 4, 0, 1),
    "voxel_cylinder_31_z": ("cylinderSimple.stl", 31, 2, 5, 0, 1),
    "voxel_cylinder_17_y": ("cylinderSimple.stl", 17, 1, 2, 2, 1),
}


def read_stl_soup(path):
    """Triangles of a binary or ASCII STL as float32 [nt,3,3] (the numbers in the file, nothing else)."""
    raw = open(path, "rb").read()
    if b"endfacet" in raw[:4096]:
        vals = [ln.split()[1:4] for ln in raw.decode("ascii", "replace").splitlines() if ln.strip().startswith("vertex")]
        return np.array(vals, dtype=np.float32).reshape(-1, 3, 3)
    nt = struct.unpack("<I", raw[80:84])[0]
    rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=nt, offset=84)
    return np.ascontiguousarray(rec["v"])


def main():
    exe = os.path.join(HERE, "_ref", "ref_voxel")
    for name, (fn, res, rdir, margin, extra, bw) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "o.bin")
            subprocess.run([exe, os.path.join(STL_DIR, fn), out, str(res), str(rdir), str(margin), str(extra), str(bw)],
                           check=True, cwd=tmp, stdout=subprocess.DEVNULL)
            b = open(out, "rb").read()
        nt = int(np.frombuffer(b[:8], dtype=np.int64)[0])
        dims = np.frombuffer(b[8:32], dtype=np.int64)
        tri = np.frombuffer(b[32:32 + nt * 72], dtype=np.float64).reshape(nt, 3, 3)
        flags = np.frombuffer(b[32 + nt * 72:], dtype=np.int32).reshape(dims).astype(np.int8)
        soup = read_stl_soup(os.path.join(STL_DIR, fn))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), soup=soup, tri=tri, flags=flags,
                            params=np.array([res, rdir, margin, extra, bw], dtype=np.int64))
        print(name, "triangles", nt, "soup", soup.shape[0], "matrix", tuple(dims), "counts", np.bincount(flags.ravel(), minlength=5))


if __name__ == "__main__":
    sys.exit(main())
