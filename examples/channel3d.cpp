// examples/channel3d.cpp -- periodic channel (BASELINE.json configs[2] / [4] in small): periodic in x
// and z, BounceBack walls at y = 0 and y = ny-1, parabolic initial profile; TRT or D3Q27 Smagorinsky
// dynamics, double or float.  Written against the Palabos API (include/palabos_b200/palabos3D.h).
//   channel3d [nx ny nz steps [trt|incbgk|smago27] [f64|f32]]
#include <cstdio>
#include <cstdlib>
#include <string>
#include "palabos_b200/palabos3D.h"

using namespace plb;

template <typename T>
struct Poiseuille {
    plint ny; double uMax;
    void operator()(plint, plint iY, plint, T& rho, Array<T, 3>& u) const {
        const double y = ((double)iY - 0.5) / ((double)ny - 2.0);
        rho = (T)1;
        u[0] = (iY == 0 || iY == ny - 1) ? (T)0 : (T)(4.0 * uMax * y * (1.0 - y));
        u[1] = (T)0; u[2] = (T)0;
    }
};

template <typename T, template <typename U> class DESCRIPTOR>
int run(plint nx, plint ny, plint nz, plint steps, Dynamics<T, DESCRIPTOR>* dyn) {
    MultiBlockLattice3D<T, DESCRIPTOR> lattice(nx, ny, nz, dyn);
    lattice.periodicity().toggle(0, true);
    lattice.periodicity().toggle(2, true);
    defineDynamics(lattice, Box3D(0, nx - 1, 0, 0, 0, nz - 1), new BounceBack<T, DESCRIPTOR>((T)1));
    defineDynamics(lattice, Box3D(0, nx - 1, ny - 1, ny - 1, 0, nz - 1), new BounceBack<T, DESCRIPTOR>((T)1));
    Poiseuille<T> profile = {ny, 0.05};
    initializeAtEquilibrium(lattice, lattice.getBoundingBox(), profile);
    lattice.initialize();
    lattice.collideAndStream(steps);
    std::printf("Av. energy %e\nAv. density %e\n", (double)computeAverageEnergy(lattice), (double)computeAverageDensity(lattice));
    std::printf("stored: energy %.17g density %.17g\n", getStoredAverageEnergy(lattice), getStoredAverageDensity(lattice));
    return 0;
}

int main(int argc, char** argv) {
    const plint nx = argc > 1 ? std::atol(argv[1]) : 32, ny = argc > 2 ? std::atol(argv[2]) : 16, nz = argc > 3 ? std::atol(argv[3]) : 16;
    const plint steps = argc > 4 ? std::atol(argv[4]) : 50;
    const std::string model = argc > 5 ? argv[5] : "trt", prec = argc > 6 ? argv[6] : "f64";
    try {
        using namespace descriptors;
        if (model == "trt" && prec == "f64") return run<double, D3Q19Descriptor>(nx, ny, nz, steps, new TRTdynamics<double, D3Q19Descriptor>(1.2));
        if (model == "incbgk") return run<double, D3Q19Descriptor>(nx, ny, nz, steps, new IncBGKdynamics<double, D3Q19Descriptor>(1.2));
        if (model == "trt") return run<float, D3Q19Descriptor>(nx, ny, nz, steps, new TRTdynamics<float, D3Q19Descriptor>(1.2f));
        if (prec == "f64") return run<double, D3Q27Descriptor>(nx, ny, nz, steps, new SmagorinskyBGKdynamics<double, D3Q27Descriptor>(1.9, 0.14));
        return run<float, D3Q27Descriptor>(nx, ny, nz, steps, new SmagorinskyBGKdynamics<float, D3Q27Descriptor>(1.9f, 0.14f));
    } catch (PlbException const& e) {
        std::fprintf(stderr, "Palabos error: %s\n", e.what());
        return 1;
    }
}
