// examples/external_rhoj3d.cpp -- the time loop of the fork's own application (viscosityTest) written against the Palabos
// API as src/variables.hh:461-486, :650 and viscosityTest/src/test.cpp:64-65 write it, compiled against
// include/palabos_b200/palabos3D.h: IncBGK collided with the moments of the rhoBar / j fields
// (ExternalRhoJcollideAndStream3D at level 0), the fields recomputed after every step (BoxRhoBarJfunctional3D at level 3),
// and -- standing for the application's momentum-forcing processors -- a constant added to j on a sub-box half way through.
//
//   external_rhoj3d [nx ny nz steps [dump.bin]]        dump = final populations then final fields ([x][y][z][q], [x][y][z][4])
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>
#include "palabos_b200/palabos3D.h"

using namespace plb;
typedef double T;
#define DESCRIPTOR descriptors::D3Q19Descriptor

int main(int argc, char** argv) {
    const plint nx = argc > 1 ? std::atol(argv[1]) : 10, ny = argc > 2 ? std::atol(argv[2]) : 9, nz = argc > 3 ? std::atol(argv[3]) : 8;
    const plint steps = argc > 4 ? std::atol(argv[4]) : 40;
    const char* dump = argc > 5 ? argv[5] : 0;
    try {
        const T uMax = 0.02, Re = 100.;
        const T omega = 1. / (3. * uMax * (T)ny / Re + 0.5);

        // the one line that differs from the reference: the lattice is told at construction that the sweep will need its
        // envelope layer (plb::WithEnvelope, non-periodic here)
        MultiBlockLattice3D<T, DESCRIPTOR> lattice(nx, ny, nz, new IncBGKdynamics<T, DESCRIPTOR>(omega), WithEnvelope());
        lattice.toggleInternalStatistics(false);

        std::unique_ptr<MultiScalarField3D<T> > rhoBar = generateMultiScalarField<T>((MultiBlock3D&)lattice, 1);   // as src/variables.hh:475
        std::unique_ptr<MultiTensorField3D<T, 3> > j = generateMultiTensorField<T, 3>(lattice, 1);
        std::vector<MultiBlock3D*> rhoBarJarg;
        rhoBarJarg.push_back(&lattice);
        rhoBarJarg.push_back(rhoBar.get());
        rhoBarJarg.push_back(j.get());
        integrateProcessingFunctional(new ExternalRhoJcollideAndStream3D<T, DESCRIPTOR>(), lattice.getBoundingBox(), rhoBarJarg, 0);
        integrateProcessingFunctional(new BoxRhoBarJfunctional3D<T, DESCRIPTOR>(), lattice.getBoundingBox(), rhoBarJarg, 3);

        OnLatticeBoundaryCondition3D<T, DESCRIPTOR>* bc = createLocalBoundaryCondition3D<T, DESCRIPTOR>();
        bc->setVelocityConditionOnBlockBoundaries(lattice);
        setBoundaryVelocity(lattice, lattice.getBoundingBox(), Array<T, 3>(0., 0., 0.));
        setBoundaryVelocity(lattice, Box3D(1, nx - 2, 1, ny - 2, nz - 1, nz - 1), Array<T, 3>(uMax, 0., 0.));
        initializeAtEquilibrium(lattice, lattice.getBoundingBox(), (T)1., Array<T, 3>(0., 0., 0.));
        delete bc;

        applyProcessingFunctional(new BoxRhoBarJfunctional3D<T, DESCRIPTOR>(), lattice.getBoundingBox(), rhoBarJarg);

        lattice.synchronize();
        auto t0 = std::chrono::steady_clock::now();
        for (plint it = 0; it < steps; ++it) {
            if (it == steps / 2) {
                // a momentum source acting on the j field between two steps
                pullRhoBarJ(lattice, *rhoBar, *j);
                const Box3D box(1, nx / 2 > 1 ? nx / 2 : 1, 1, ny - 2, nz / 3 > 1 ? nz / 3 : 1, nz - 2);
                for (plint x = box.x0; x <= box.x1; ++x) for (plint y = box.y0; y <= box.y1; ++y) for (plint z = box.z0; z <= box.z1; ++z) {
                    Array<T, 3>& v = j->get(x, y, z);
                    v[0] += 1e-3; v[1] += -5e-4; v[2] += 2.5e-4;
                }
                pushRhoBarJ(lattice, *rhoBar, *j);
            }
            lattice.executeInternalProcessors();
            lattice.incrementTime();
        }
        lattice.synchronize();
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        std::printf("Av. energy %e\n", (double)computeAverageEnergy(lattice));
        std::printf("Av. density %e\n", (double)computeAverageDensity(lattice));
        std::printf("iterations %ld\n", (long)lattice.getTimeCounter());
        std::printf("Mega site updates per second: %.1f\n", (double)nx * ny * nz * steps / sec / 1e6);
        if (dump) {
            std::vector<T> pops;
            lattice.getPopulations(lattice.getBoundingBox(), pops);
            pullRhoBarJ(lattice, *rhoBar, *j);
            FILE* fp = std::fopen(dump, "wb");
            if (!fp) { std::fprintf(stderr, "cannot write %s\n", dump); return 2; }
            std::fwrite(pops.data(), sizeof(T), pops.size(), fp);
            for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
                const T rec[4] = {rhoBar->get(x, y, z), j->get(x, y, z)[0], j->get(x, y, z)[1], j->get(x, y, z)[2]};
                std::fwrite(rec, sizeof(T), 4, fp);
            }
            std::fclose(fp);
        }
    } catch (PlbException const& e) {
        std::fprintf(stderr, "external_rhoj3d: %s\n", e.what());
        return 1;
    }
    return 0;
}
