// examples/cavity3d.cpp -- the reference's 3D lid-driven cavity regression case
// (jlabos/src/testing/test_cylinder.py:100-119, jlabos/examples/Cavity3d/Cavity3d.java) written against
// the Palabos API exactly as one would write it for the reference, but compiled against
// include/palabos_b200/palabos3D.h: host code in C++, the hot path on the B200 through libplb_b200.so.
//
//   cavity3d [nx ny nz steps [dump.bin [D3Q27]]]
// prints "Av. energy" / "Av. density" in the format of the reference's golden strings
// (test_cylinder.py:85-97) and optionally dumps the final populations ([x][y][z][q] doubles).
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <string>
#include <vector>
#include "palabos_b200/palabos3D.h"

using namespace plb;
typedef double T;

template <template <typename U> class DESCRIPTOR>
int run(plint nx, plint ny, plint nz, plint steps, const char* dump) {
    const T uMax = 0.02, Re = 100.;
    const T nu = uMax * (T)ny / Re;
    const T tau = 3. * nu + 0.5;

    MultiBlockLattice3D<T, DESCRIPTOR> lattice(nx, ny, nz, new BGKdynamics<T, DESCRIPTOR>(1. / tau));
    // CAVITY3D_BC=dynlocal | equilibrium: the dynamics-based managers (populations completed inside collide()); interp: the
    // finite-difference manager of the reference's cavity3d showcase
    const std::string model = std::getenv("CAVITY3D_BC") ? std::getenv("CAVITY3D_BC") : "local";
    OnLatticeBoundaryCondition3D<T, DESCRIPTOR>* bc =
        model == "dynlocal" ? createDynamicsBasedLocalBoundaryCondition3D<T, DESCRIPTOR>()
                            : (model == "equilibrium" ? createEquilibriumBoundaryCondition3D<T, DESCRIPTOR>()
                               : (model == "interp" ? createInterpBoundaryCondition3D<T, DESCRIPTOR>() : createLocalBoundaryCondition3D<T, DESCRIPTOR>()));
    bc->setVelocityConditionOnBlockBoundaries(lattice);
    if (std::getenv("CAVITY3D_BLOCK")) {
        // a solid block inside the cavity: its faces are velocity planes, its edges and corners the CONCAVE ones of the boundary
        // condition (addInternalVelocityEdge* / addInternalVelocityCorner*); normals point from the fluid into the solid
        const plint x0 = nx / 3, x1 = nx / 3 + 3, y0 = ny / 3, y1 = ny / 3 + 3, z0 = nz / 3, z1 = nz / 3 + 2;
        defineDynamics(lattice, Box3D(x0 + 1, x1 - 1, y0 + 1, y1 - 1, z0 + 1, z1 - 1), new BounceBack<T, DESCRIPTOR>((T)1.));
        bc->addVelocityBoundary0P(Box3D(x0, x0, y0 + 1, y1 - 1, z0 + 1, z1 - 1), lattice);
        bc->addVelocityBoundary0N(Box3D(x1, x1, y0 + 1, y1 - 1, z0 + 1, z1 - 1), lattice);
        bc->addVelocityBoundary1P(Box3D(x0 + 1, x1 - 1, y0, y0, z0 + 1, z1 - 1), lattice);
        bc->addVelocityBoundary1N(Box3D(x0 + 1, x1 - 1, y1, y1, z0 + 1, z1 - 1), lattice);
        bc->addVelocityBoundary2P(Box3D(x0 + 1, x1 - 1, y0 + 1, y1 - 1, z0, z0), lattice);
        bc->addVelocityBoundary2N(Box3D(x0 + 1, x1 - 1, y0 + 1, y1 - 1, z1, z1), lattice);
        bc->addInternalVelocityEdge0PP(Box3D(x0 + 1, x1 - 1, y0, y0, z0, z0), lattice);
        bc->addInternalVelocityEdge0PN(Box3D(x0 + 1, x1 - 1, y0, y0, z1, z1), lattice);
        bc->addInternalVelocityEdge0NP(Box3D(x0 + 1, x1 - 1, y1, y1, z0, z0), lattice);
        bc->addInternalVelocityEdge0NN(Box3D(x0 + 1, x1 - 1, y1, y1, z1, z1), lattice);
        bc->addInternalVelocityEdge1PP(Box3D(x0, x0, y0 + 1, y1 - 1, z0, z0), lattice);
        bc->addInternalVelocityEdge1PN(Box3D(x1, x1, y0 + 1, y1 - 1, z0, z0), lattice);
        bc->addInternalVelocityEdge1NP(Box3D(x0, x0, y0 + 1, y1 - 1, z1, z1), lattice);
        bc->addInternalVelocityEdge1NN(Box3D(x1, x1, y0 + 1, y1 - 1, z1, z1), lattice);
        bc->addInternalVelocityEdge2PP(Box3D(x0, x0, y0, y0, z0 + 1, z1 - 1), lattice);
        bc->addInternalVelocityEdge2PN(Box3D(x0, x0, y1, y1, z0 + 1, z1 - 1), lattice);
        bc->addInternalVelocityEdge2NP(Box3D(x1, x1, y0, y0, z0 + 1, z1 - 1), lattice);
        bc->addInternalVelocityEdge2NN(Box3D(x1, x1, y1, y1, z0 + 1, z1 - 1), lattice);
        bc->addInternalVelocityCornerPPP(x0, y0, z0, lattice); bc->addInternalVelocityCornerPPN(x0, y0, z1, lattice);
        bc->addInternalVelocityCornerPNP(x0, y1, z0, lattice); bc->addInternalVelocityCornerPNN(x0, y1, z1, lattice);
        bc->addInternalVelocityCornerNPP(x1, y0, z0, lattice); bc->addInternalVelocityCornerNPN(x1, y0, z1, lattice);
        bc->addInternalVelocityCornerNNP(x1, y1, z0, lattice); bc->addInternalVelocityCornerNNN(x1, y1, z1, lattice);
    }
    setBoundaryVelocity(lattice, lattice.getBoundingBox(), Array<T, 3>(0., 0., 0.));
    setBoundaryVelocity(lattice, Box3D(1, nx - 2, 1, ny - 2, nz - 1, nz - 1), Array<T, 3>(uMax, 0., 0.));
    initializeAtEquilibrium(lattice, lattice.getBoundingBox(), (T)1., Array<T, 3>(0., 0., 0.));
    lattice.initialize();
    delete bc;

    lattice.synchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (plint it = 0; it < steps; ++it) lattice.collideAndStream();
    lattice.synchronize();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::printf("Av. energy %e\n", (double)computeAverageEnergy(lattice));
    std::printf("Av. density %e\n", (double)computeAverageDensity(lattice));
    std::printf("stored: energy %.17g density %.17g max velocity %.17g\n", getStoredAverageEnergy(lattice),
                getStoredAverageDensity(lattice), getStoredMaxVelocity(lattice));
    std::printf("Mega site updates per second: %.1f\n", (double)nx * ny * nz * steps / sec / 1e6);
    if (dump && std::getenv("CAVITY3D_FIELDS")) {
        // what the reference's cavity3d writes as images / VTK: density, velocity, velocity norm (dataAnalysisWrapper3D.h:50-70);
        // and a checkpoint in the reference's format (saveBinaryBlock)
        std::unique_ptr<MultiScalarField3D<T> > rho = computeDensity(lattice);
        std::unique_ptr<MultiTensorField3D<T, 3> > u = computeVelocity(lattice);
        std::unique_ptr<MultiScalarField3D<T> > un = computeVelocityNorm(lattice);
        std::unique_ptr<MultiScalarField3D<T> > uy = computeVelocityComponent(lattice, lattice.getBoundingBox(), 1);
        FILE* fp = std::fopen(std::getenv("CAVITY3D_FIELDS"), "wb");
        if (!fp) { std::fprintf(stderr, "cannot write fields\n"); return 2; }
        for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
            const T rec[6] = {rho->get(x, y, z), u->get(x, y, z)[0], u->get(x, y, z)[1], u->get(x, y, z)[2], un->get(x, y, z), uy->get(x, y, z)};
            std::fwrite(rec, sizeof(T), 6, fp);
        }
        std::fclose(fp);
        saveBinaryBlock(lattice, std::string(std::getenv("CAVITY3D_FIELDS")) + ".ckpt");
        {   // the writeVTK of the reference's examples (showCases/cavity3d): physical units through dx, dt
            const T dx = 0.125, dt = 0.002;
            VtkImageOutput3D<T, float, 3> vtkOut(std::string(std::getenv("CAVITY3D_FIELDS")) + ".vtk", dx);
            vtkOut.writeData(*computeVelocityNorm(lattice), "velocityNorm", (float)(dx / dt));
            vtkOut.writeData(*computeVelocity(lattice), "velocity", (float)(dx / dt));
            vtkOut.writeData(*computeDensity(lattice), "density", (float)1, (float)-1);
        }
    }
    if (dump) {
        std::vector<T> pops;
        lattice.getPopulations(lattice.getBoundingBox(), pops);
        FILE* fp = std::fopen(dump, "wb");
        if (!fp) { std::fprintf(stderr, "cannot write %s\n", dump); return 2; }
        std::fwrite(pops.data(), sizeof(T), pops.size(), fp);
        std::fclose(fp);
    }
    return 0;
}

int main(int argc, char** argv) {
    const plint nx = argc > 1 ? std::atol(argv[1]) : 20;
    const plint ny = argc > 2 ? std::atol(argv[2]) : nx;
    const plint nz = argc > 3 ? std::atol(argv[3]) : nx;
    const plint steps = argc > 4 ? std::atol(argv[4]) : 30;
    const char* dump = argc > 5 && std::string(argv[5]) != "-" ? argv[5] : 0;
    const bool q27 = argc > 6 && std::string(argv[6]) == "D3Q27";
    try {
        return q27 ? run<descriptors::D3Q27Descriptor>(nx, ny, nz, steps, dump)
                   : run<descriptors::D3Q19Descriptor>(nx, ny, nz, steps, dump);
    } catch (PlbException const& e) {
        std::fprintf(stderr, "Palabos error: %s\n", e.what());
        return 1;
    }
}
