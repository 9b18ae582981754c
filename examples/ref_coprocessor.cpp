// examples/ref_coprocessor.cpp -- the B200 engine plugged into the REAL, unmodified Palabos through the reference's own
// accelerator interface plb::CoProcessor3D<T>.
//
// This program is compiled against /root/reference/src (headers) + oracle/_ref/libplb.a (the reference's .cpp files) and linked
// with palabos_b200/libplb_b200.so.  The only edit a maintainer makes is the one the reference's own comment asks for
// (coProcessors/coProcessor3D.h:179-186: "simply replace the ... singleton"): global::defaultCoProcessor3D<T>() returns
// plb::B200CoProcessor3D<T> (include/palabos_b200/coprocessor_adapter.h).  Here that one function is patched in a COPY of the
// header made outside the repository by oracle/Makefile, which shadows the original through the include path -- nothing of
// the reference is copied into this tree.
//
// What runs (all of it the reference's code except the four virtual calls into libplb_b200.so):
//   MultiBlockLattice3D with an ExplicitThreadAttribution (threadAttribution.h:95-125: the only attribution that keeps handles)
//   initiateCoProcessors(lattice, BGKdynamics id)         coProcessorCommunication3D.hh:77-124  -> addDomain per pure block
//   per step:  transferToCoProcessors(lattice)            :36-54                               -> send (whole blocks, envelope included)
//              lattice.collideAndStream()                 multiBlockLattice3D.hh:318-360       -> collideAndStream(handle) for pure
//                                                          blocks, the reference's own sweep for the others, then processors
//              transferFromCoProcessors(lattice)          :56-75                               -> receive
//              lattice.duplicateOverlaps(staticVariables)                                      (the device leaves the outer layer
//                                                          of a domain undefined, coProcessor3D.h:105-108: the envelopes are
//                                                          refreshed from the neighbours' bulks, as after any step)
// and is compared, cell by cell and bit for bit, with the same lattice stepped by the reference alone:
//   case A  fully periodic 24x20x16 box in 2x2x2 blocks, all eight pure BGK: everything runs on the device
//   case B  30^3 lid-driven cavity (regularized velocity BC) in 3x3x3 blocks: the 26 blocks that touch the boundary are not
//           pure and stay on the CPU, the centre block runs on the device
// Exit code 0 and "REF_COPROCESSOR_OK" when every population of every bulk cell is identical in both cases.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <map>
#include <vector>

typedef double T;   // must precede the Palabos headers (src/debug.h needs a global T)

#include "palabos3D.h"
#include "latticeBoltzmann/headers3D.hh"
#include "core/headers3D.hh"
#include "basicDynamics/headers3D.hh"
#include "boundaryCondition/headers3D.hh"
#include "complexDynamics/headers3D.hh"
#include "io/headers3D.hh"
#include "atomicBlock/headers3D.hh"
#include "multiBlock/headers3D.hh"
#include "algorithm/headers3D.hh"
#include "dataProcessors/headers3D.hh"
#include "finiteDifference/headers3D.hh"
#include "coProcessors/headers3D.hh"
#include "core/headers2D.hh"
#include "atomicBlock/headers2D.hh"
#include "multiBlock/headers2D.hh"

using namespace plb;
#define D descriptors::D3Q19Descriptor

/// deterministic, non-uniform initial state (evaluated in the periodic envelope too: periodic in the lattice size)
struct WavyRhoU {
    plint nx, ny, nz;
    void operator()(plint iX, plint iY, plint iZ, T& rho, Array<T, 3>& u) const {
        const double pi2 = 6.283185307179586;
        const double x = pi2 * (double)((iX % nx + nx) % nx) / (double)nx, y = pi2 * (double)((iY % ny + ny) % ny) / (double)ny,
                     z = pi2 * (double)((iZ % nz + nz) % nz) / (double)nz;
        rho = (T)(1.0 + 0.01 * std::cos(x) * std::sin(y + 0.3));
        u[0] = (T)(0.03 + 0.01 * std::sin(y) * std::cos(z));
        u[1] = (T)(0.01 * std::sin(x + 0.5) * std::sin(z));
        u[2] = (T)(-0.01 * std::cos(x) * std::sin(y));
    }
};

static MultiBlockLattice3D<T, D>* makeLattice(plint nx, plint ny, plint nz, plint bx, plint by, plint bz, T omega) {
    SparseBlockStructure3D blocks = createRegularDistribution3D(nx, ny, nz, bx, by, bz);
    ExplicitThreadAttribution* attribution = new ExplicitThreadAttribution;
    std::map<plint, Box3D> const& bulks = blocks.getBulks();
    for (std::map<plint, Box3D>::const_iterator it = bulks.begin(); it != bulks.end(); ++it) attribution->addBlock(it->first, 0);
    return new MultiBlockLattice3D<T, D>(MultiBlockManagement3D(blocks, attribution, 1),
                                         defaultMultiBlockPolicy3D().getBlockCommunicator(),
                                         defaultMultiBlockPolicy3D().getCombinedStatistics(),
                                         defaultMultiBlockPolicy3D().getMultiCellAccess<T, D>(), new BGKdynamics<T, D>(omega));
}

static void setUp(MultiBlockLattice3D<T, D>& lattice, bool cavity) {
    const plint nx = lattice.getNx(), ny = lattice.getNy(), nz = lattice.getNz();
    if (cavity) {
        OnLatticeBoundaryCondition3D<T, D>* bc = createLocalBoundaryCondition3D<T, D>();
        bc->setVelocityConditionOnBlockBoundaries(lattice, lattice.getBoundingBox());
        setBoundaryVelocity(lattice, lattice.getBoundingBox(), Array<T, 3>((T)0, (T)0, (T)0));
        setBoundaryVelocity(lattice, Box3D(1, nx - 2, 1, ny - 2, nz - 1, nz - 1), Array<T, 3>((T)0.02, (T)0, (T)0));
        initializeAtEquilibrium(lattice, lattice.getBoundingBox(), (T)1, Array<T, 3>((T)0, (T)0, (T)0));
        WavyRhoU f = {nx, ny, nz};
        initializeAtEquilibrium(lattice, Box3D(3, nx - 4, 3, ny - 4, 3, nz - 4), f);     // something to transport through the centre block
        delete bc;
    } else {
        lattice.periodicity().toggleAll(true);
        WavyRhoU f = {nx, ny, nz};
        initializeAtEquilibrium(lattice, lattice.getBoundingBox(), f);
    }
    lattice.initialize();
}

/// number of populations that differ (as bit patterns of their values) over all bulk cells
static long compare(MultiBlockLattice3D<T, D>& a, MultiBlockLattice3D<T, D>& b, double& maxAbs) {
    long bad = 0;
    maxAbs = 0.;
    Array<T, D<T>::q> fa, fb;
    for (plint x = 0; x < a.getNx(); ++x) for (plint y = 0; y < a.getNy(); ++y) for (plint z = 0; z < a.getNz(); ++z) {
        a.get(x, y, z).getPopulations(fa);
        b.get(x, y, z).getPopulations(fb);
        for (plint i = 0; i < D<T>::q; ++i) {
            if (!(fa[i] == fb[i])) { ++bad; maxAbs = std::max(maxAbs, std::fabs((double)(fa[i] - fb[i]))); }
        }
    }
    return bad;
}

static int runCase(const char* name, plint nx, plint ny, plint nz, plint bx, plint by, plint bz, bool cavity, int steps, int expectPure) {
    const T omega = (T)1.3;
    MultiBlockLattice3D<T, D>* cpu = makeLattice(nx, ny, nz, bx, by, bz, omega);
    MultiBlockLattice3D<T, D>* dev = makeLattice(nx, ny, nz, bx, by, bz, omega);
    setUp(*cpu, cavity);
    setUp(*dev, cavity);

    initiateCoProcessors(*dev, BGKdynamics<T, D>(omega).getId(), false);
    ThreadAttribution const& attribution = dev->getMultiBlockManagement().getThreadAttribution();
    std::vector<plint> const& ids = dev->getMultiBlockManagement().getLocalInfo().getBlocks();
    int nPure = 0;
    for (pluint k = 0; k < ids.size(); ++k) if (attribution.getCoProcessorHandle(ids[k]) >= 0) ++nPure;
    std::printf("%s: %ldx%ldx%ld in %ldx%ldx%ld blocks, %d of %d blocks on the co-processor\n", name, (long)nx, (long)ny, (long)nz,
                (long)bx, (long)by, (long)bz, nPure, (int)ids.size());
    if (nPure != expectPure) { std::printf("  expected %d pure blocks\n", expectPure); return 1; }

    for (int s = 0; s < steps; ++s) {
        cpu->collideAndStream();
        transferToCoProcessors(*dev);
        dev->collideAndStream();
        transferFromCoProcessors(*dev);
        dev->duplicateOverlaps(modif::staticVariables);
    }
    double maxAbs = 0.;
    const long bad = compare(*cpu, *dev, maxAbs);
    std::printf("  after %d steps: %ld differing populations (max abs diff %.3e), average energy %.12e / %.12e\n", steps, bad, maxAbs,
                (double)computeAverageEnergy(*cpu), (double)computeAverageEnergy(*dev));
    delete cpu;
    delete dev;
    return bad == 0 ? 0 : 1;
}

int main(int argc, char* argv[]) {
    plbInit(&argc, &argv);
    const int steps = argc > 1 ? std::atoi(argv[1]) : 20;
    int rc = 0;
    rc |= runCase("A periodic box", 24, 20, 16, 2, 2, 2, false, steps, 8);
    rc |= runCase("B cavity", 30, 30, 30, 3, 3, 3, true, steps, 1);
    if (rc == 0) std::printf("REF_COPROCESSOR_OK\n");
    return rc;
}
