// oracle/ref_voxel.cpp -- TEST INFRASTRUCTURE: drives the UNMODIFIED reference's STL -> voxel-flag pipeline
// (utility/stl/revoxelize.cpp:40-80: TriangleSet, DEFscaledMesh, TriangleBoundary3D::inflate, VoxelizedDomain3D)
// and dumps (a) the triangles handed to voxelize(), in lattice units, (b) the voxel matrix.
typedef double T;
#include "palabos3D.h"
#include "core/headers3D.hh"
#include "atomicBlock/headers3D.hh"
#include "multiBlock/headers3D.hh"
#include "dataProcessors/headers3D.hh"
#include "offLattice/headers3D.hh"
#include "particles/headers3D.hh"
#include "io/headers3D.hh"
#include "algorithm/headers3D.hh"
#include <cstdio>
using namespace plb;

int main(int argc, char* argv[]) {
    plbInit(&argc, &argv);
    if (argc < 5) { printf("usage: ref_voxel in.stl out.bin resolution referenceDirection [margin extraLayer borderWidth]\n"); return 1; }
    std::string stl = argv[1], out = argv[2];
    plint resolution = atoi(argv[3]), refDir = atoi(argv[4]);
    plint margin = argc > 5 ? atoi(argv[5]) : 4, extraLayer = argc > 6 ? atoi(argv[6]) : 0, borderWidth = argc > 7 ? atoi(argv[7]) : 1;
    TriangleSet<T> triangleSet(stl, FLT);
    DEFscaledMesh<T> mesh(triangleSet, resolution, refDir, margin, extraLayer);
    TriangleBoundary3D<T> boundary(mesh);
    boundary.getMesh().inflate();
    VoxelizedDomain3D<T> vd(boundary, voxelFlag::inside, extraLayer, borderWidth, 1, 0);
    MultiScalarField3D<int>& m = vd.getVoxelMatrix();
    boundary.pushSelect(1, 0);
    TriangularSurfaceMesh<T> const& tm = boundary.getMesh();
    FILE* f = fopen(out.c_str(), "wb");
    long long nt = tm.getNumTriangles(), dims[3] = {m.getNx(), m.getNy(), m.getNz()};
    fwrite(&nt, 8, 1, f); fwrite(dims, 8, 3, f);
    for (plint t = 0; t < nt; ++t) for (int v = 0; v < 3; ++v) { Array<T,3> p = tm.getVertex(t, v); fwrite(&p[0], 8, 3, f); }
    for (plint x = 0; x < dims[0]; ++x) for (plint y = 0; y < dims[1]; ++y) for (plint z = 0; z < dims[2]; ++z) { int v = m.get(x, y, z); fwrite(&v, 4, 1, f); }
    fclose(f);
    printf("triangles %lld, matrix %lld x %lld x %lld\n", nt, dims[0], dims[1], dims[2]);
    return 0;
}
