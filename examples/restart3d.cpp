// examples/restart3d.cpp -- checkpoint / restart of the lid-driven cavity the way the reference's applications do it
// (saveState / loadState, io/utilIO_3D.cpp:45-76; parallelIO::save / load, io/multiBlockWriter3D.cpp:226-238), written
// against include/palabos_b200/palabos3D.h.
//
//   restart3d <dir> [slabs]
// runs 5 steps, saves the state into <dir>, runs 4 more and dumps the populations to <dir>/continued.bin; then builds a
// second lattice (with `slabs` x-slabs), restarts it from the files and runs the same 4 steps: <dir>/restarted.bin.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "palabos_b200/palabos3D.h"

using namespace plb;
typedef double T;
#define DESCRIPTOR descriptors::D3Q19Descriptor

static void setup(MultiBlockLattice3D<T, DESCRIPTOR>& lattice, plint nx, plint ny, plint nz) {
    OnLatticeBoundaryCondition3D<T, DESCRIPTOR>* bc = createLocalBoundaryCondition3D<T, DESCRIPTOR>();
    bc->setVelocityConditionOnBlockBoundaries(lattice);
    setBoundaryVelocity(lattice, lattice.getBoundingBox(), Array<T, 3>(0., 0., 0.));
    setBoundaryVelocity(lattice, Box3D(1, nx - 2, 1, ny - 2, nz - 1, nz - 1), Array<T, 3>(0.02, 0.01, 0.));
    initializeAtEquilibrium(lattice, lattice.getBoundingBox(), (T)1., Array<T, 3>(0., 0., 0.));
    lattice.initialize();
    delete bc;
}

static int dump(MultiBlockLattice3D<T, DESCRIPTOR>& lattice, const std::string& file) {
    std::vector<T> pops;
    lattice.getPopulations(lattice.getBoundingBox(), pops);
    FILE* fp = std::fopen(file.c_str(), "wb");
    if (!fp) { std::fprintf(stderr, "cannot write %s\n", file.c_str()); return 2; }
    std::fwrite(pops.data(), sizeof(T), pops.size(), fp);
    std::fclose(fp);
    return 0;
}

int main(int argc, char* argv[]) {
    if (argc < 2) { std::fprintf(stderr, "usage: restart3d <dir> [slabs]\n"); return 1; }
    const std::string dir = argv[1];
    const int slabs = argc > 2 ? std::atoi(argv[2]) : 1;
    const plint nx = 6, ny = 5, nz = 4;
    try {
        global::directories().setOutputDir(dir + "/");
        global::directories().setInputDir(dir + "/");
        {
            MultiBlockLattice3D<T, DESCRIPTOR> lattice(nx, ny, nz, new BGKdynamics<T, DESCRIPTOR>(1.3), 0, 1, 2);
            setup(lattice, nx, ny, nz);
            plint iT = 0;
            for (; iT < 5; ++iT) lattice.collideAndStream();
            std::vector<MultiBlockLattice3D<T, DESCRIPTOR>*> blocks;
            blocks.push_back(&lattice);
            const bool saveDynamicContent = false;
            saveState(blocks, iT, saveDynamicContent, "restart.xml", "ckpt", 8);
            for (; iT < 9; ++iT) lattice.collideAndStream();
            if (dump(lattice, dir + "/continued.bin")) return 2;
        }
        {
            MultiBlockLattice3D<T, DESCRIPTOR> lattice(nx, ny, nz, new BGKdynamics<T, DESCRIPTOR>(1.3), 0, 1, slabs);
            setup(lattice, nx, ny, nz);
            std::vector<MultiBlockLattice3D<T, DESCRIPTOR>*> blocks;
            blocks.push_back(&lattice);
            plint iT = -1;
            loadState(blocks, iT, false, dir + "/restart.xml");
            std::printf("restarted at iteration %ld\n", (long)iT);
            for (; iT < 9; ++iT) lattice.collideAndStream();
            if (dump(lattice, dir + "/restarted.bin")) return 2;
        }
    } catch (PlbException const& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 3;
    }
    return 0;
}
