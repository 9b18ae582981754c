// examples/voxelize_stl.cpp -- the reference's utility/stl/revoxelize.cpp:40-80 up to the voxel matrix, written
// against include/palabos_b200/palabos3D.h: read an STL, scale it to lattice units, voxelize it on the device,
// print the number of voxels of each kind, then run a few steps of flow around it with BounceBack on the
// inside voxels (defineDynamics(lattice, voxelMatrix, BounceBack, flag)).
//   voxelize_stl in.stl resolution referenceDirection [flags.bin]
#include <cstdio>
#include <cstdlib>
#include "palabos_b200/palabos3D.h"

using namespace plb;
typedef double T;
#define DESCRIPTOR descriptors::D3Q19Descriptor

const plint extraLayer = 0;
const plint borderWidth = 1;
const plint margin = 4;

int main(int argc, char* argv[]) {
    if (argc < 4) { std::printf("usage: %s in.stl resolution referenceDirection [flags.bin]\n", argv[0]); return 1; }
    try {
        TriangleSet<T> triangleSet(argv[1], FLT);
        DEFscaledMesh<T> mesh(triangleSet, std::atoi(argv[2]), std::atoi(argv[3]), margin, extraLayer);
        mesh.inflate();
        std::unique_ptr<MultiScalarField3D<int> > voxels = voxelize(mesh, mesh.getMargin() + extraLayer, borderWidth);
        plint count[5] = {0, 0, 0, 0, 0};
        for (plint x = 0; x < voxels->getNx(); ++x) for (plint y = 0; y < voxels->getNy(); ++y) for (plint z = 0; z < voxels->getNz(); ++z)
            ++count[voxels->get(x, y, z)];
        std::printf("matrix %ld x %ld x %ld, %ld triangles\n", (long)voxels->getNx(), (long)voxels->getNy(), (long)voxels->getNz(), (long)mesh.getNumTriangles());
        std::printf("outside %ld outerBorder %ld inside %ld innerBorder %ld\n", (long)count[1], (long)count[2], (long)count[3], (long)count[4]);
        std::printf("Number of inside cells: %ld\n", (long)(count[3] + count[4]));
        if (argc > 4) {
            FILE* f = std::fopen(argv[4], "wb");
            std::fwrite(voxels->b200Data().data(), sizeof(int), voxels->b200Data().size(), f);
            std::fclose(f);
        }
        // flow around the obstacle: periodic box twice as long as the matrix
        const plint nx = 2 * voxels->getNx(), ny = voxels->getNy(), nz = voxels->getNz();
        MultiBlockLattice3D<T, DESCRIPTOR> lattice(nx, ny, nz, new BGKdynamics<T, DESCRIPTOR>(1.7));
        lattice.periodicity().toggleAll(true);
        defineDynamics(lattice, *voxels, Array<plint, 3>(voxels->getNx() / 2, 0, 0), new BounceBack<T, DESCRIPTOR>(1.), voxelFlag::inside);
        defineDynamics(lattice, *voxels, Array<plint, 3>(voxels->getNx() / 2, 0, 0), new BounceBack<T, DESCRIPTOR>(1.), voxelFlag::innerBorder);
        initializeAtEquilibrium(lattice, lattice.getBoundingBox(), (T)1., Array<T, 3>((T)0.03, (T)0., (T)0.));
        lattice.initialize();
        for (int it = 0; it < 20; ++it) lattice.collideAndStream();
        std::printf("after 20 steps: Av. energy %e Av. density %e\n", (double)computeAverageEnergy(lattice), (double)computeAverageDensity(lattice));
    } catch (std::exception const& e) {
        std::printf("error: %s\n", e.what());
        return 2;
    }
    return 0;
}
