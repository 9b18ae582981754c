#!/usr/bin/env python
"""Golden vectors of the STL -> voxel-flag pipeline, produced by the UNMODIFIED
reference through oracle/_ref/ref_voxel (utility/stl/revoxelize.cpp:40-80 restated as a driver).
Run in the build container (needs /root/reference); writes tests/golden/voxel_*.npz:
  soup    float64 [nt,3,3]  the triangles of the STL file as TriangleSet holds them: ASCII numbers read as double,
                            degenerate facets dropped, vertices 1, 2 swapped where the facet normal says so
                            (read by oracle/voxel_oracle.py:read_stl, the restatement of TriangleSet::readSTL)
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

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import voxel_oracle  # noqa: E402

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
        soup = voxel_oracle.read_stl(os.path.join(STL_DIR, fn))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), soup=soup, tri=tri, flags=flags,
                            params=np.array([res, rdir, margin, extra, bw], dtype=np.int64))
        print(name, "triangles", nt, "soup", soup.shape[0], "matrix", tuple(dims), "counts", np.bincount(flags.ravel(), minlength=5))


if __name__ == "__main__":
    sys.exit(main())
