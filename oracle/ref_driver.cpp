// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A scenario interpreter that drives the UNMODIFIED reference (Palabos, compiled from
// /root/reference/src by oracle/Makefile into oracle/_ref/) through its own public API:
// MultiBlockLattice3D, createLocalBoundaryCondition3D / createZouHeBoundaryCondition3D,
// defineDynamics, setBoundaryVelocity/Density, initializeAtEquilibrium, collideAndStream,
// getStored*/computeAverage*.  The same scenario text is fed to oracle/liboracle.so (the C
// restatement) and to libplb_b200.so (the CUDA engine) by tests/scenario.py, so the three
// can be compared cell by cell.  It is also the `cpu_baseline.kind == "reference"` timer of
// bench.py (command `time_run`).  Nothing here is copied from the reference; it only calls it.
//
// Build: make -C oracle ref   (-DREF_REAL=double|float)
// Usage: ref_driver <scenario.txt>      (paths inside the scenario are relative to cwd)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#ifndef REF_REAL
#define REF_REAL double
#endif
typedef REF_REAL T;   // must precede the Palabos headers (src/debug.h needs a global T)

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

static void fail(std::string const& msg) {
    std::fprintf(stderr, "ref_driver: %s\n", msg.c_str());
    std::exit(2);
}

template<typename U>
static std::vector<U> readBlob(std::string const& fname, size_t n) {
    std::vector<U> v(n);
    FILE* fp = std::fopen(fname.c_str(), "rb");
    if (!fp) fail("cannot open " + fname);
    size_t got = std::fread(v.data(), sizeof(U), n, fp);
    std::fclose(fp);
    if (got != n) fail("short read on " + fname);
    return v;
}

static Box3D readBox(std::istringstream& in) {
    plint a[6];
    for (int i = 0; i < 6; ++i) if (!(in >> a[i])) fail("bad box");
    return Box3D(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/// Functor for the generic initializeAtEquilibrium(lattice, box, f) of the reference
/// (dataProcessors/dataInitializerWrapper3D.h, in-place overload): f(x,y,z,rho,u).
struct FieldRhoU {
    FieldRhoU(std::vector<double> const& d, Box3D b) : data(d), box(b) {}
    void operator()(plint iX, plint iY, plint iZ, T& rho, Array<T,3>& u) const {
        // the reference evaluates the functor in the periodic envelope as well (coordinates -1 and n): a tabulated
        // field answers there with its periodic image
        plint nx = box.getNx(), ny = box.getNy(), nz = box.getNz();
        plint x = ((iX - box.x0) % nx + nx) % nx, y = ((iY - box.y0) % ny + ny) % ny, z = ((iZ - box.z0) % nz + nz) % nz;
        size_t idx = ((size_t)x * ny + y) * nz + z;
        rho = (T)data[4*idx];
        u[0] = (T)data[4*idx+1]; u[1] = (T)data[4*idx+2]; u[2] = (T)data[4*idx+3];
    }
    std::vector<double> const& data;
    Box3D box;
};

template<template<typename U> class D>
static Dynamics<T,D>* makeDynamics(std::string const& kind, double p1, double p2) {
    if (kind == "BGK")   return new BGKdynamics<T,D>((T)p1);
    if (kind == "RLB")   return new RegularizedBGKdynamics<T,D>((T)p1);
    if (kind == "TRT")   return new TRTdynamics<T,D>((T)p1);
    if (kind == "SMAGO") return new SmagorinskyBGKdynamics<T,D>((T)p1, (T)p2);
    if (kind == "INCBGK") return new IncBGKdynamics<T,D>((T)p1);
    if (kind == "BB")    return new BounceBack<T,D>((T)p1);
    if (kind == "NODYN") return new NoDynamics<T,D>((T)p1);
    fail("unknown dynamics " + kind);
    return 0;
}

template<template<typename U> class D>
static OnLatticeBoundaryCondition3D<T,D>* makeBC(std::string const& kind) {
    if (kind == "local") return createLocalBoundaryCondition3D<T,D>();
    if (kind == "zouhe") return createZouHeBoundaryCondition3D<T,D>();
    if (kind == "interp") return createInterpBoundaryCondition3D<T,D>();
    if (kind == "dynlocal") return createDynamicsBasedLocalBoundaryCondition3D<T,D>();
    if (kind == "equilibrium") return createEquilibriumBoundaryCondition3D<T,D>();
    fail("unknown boundary condition " + kind);
    return 0;
}

template<template<typename U> class D>
static void addPlane(OnLatticeBoundaryCondition3D<T,D>& bc, MultiBlockLattice3D<T,D>& lat,
                     bool velocity, int dir, int ori, Box3D box)
{
    int code = dir * 2 + (ori > 0 ? 1 : 0);
    if (velocity) {
        switch (code) {
            case 0: bc.addVelocityBoundary0N(box, lat); break;
            case 1: bc.addVelocityBoundary0P(box, lat); break;
            case 2: bc.addVelocityBoundary1N(box, lat); break;
            case 3: bc.addVelocityBoundary1P(box, lat); break;
            case 4: bc.addVelocityBoundary2N(box, lat); break;
            case 5: bc.addVelocityBoundary2P(box, lat); break;
        }
    } else {
        switch (code) {
            case 0: bc.addPressureBoundary0N(box, lat); break;
            case 1: bc.addPressureBoundary0P(box, lat); break;
            case 2: bc.addPressureBoundary1N(box, lat); break;
            case 3: bc.addPressureBoundary1P(box, lat); break;
            case 4: bc.addPressureBoundary2N(box, lat); break;
            case 5: bc.addPressureBoundary2P(box, lat); break;
        }
    }
}

template<template<typename U> class D>
static void addEdge(OnLatticeBoundaryCondition3D<T,D>& bc, MultiBlockLattice3D<T,D>& lat,
                    int plane, int n1, int n2, Box3D box)
{
    int code = plane * 4 + (n1 > 0 ? 2 : 0) + (n2 > 0 ? 1 : 0);
    switch (code) {
        case 0:  bc.addExternalVelocityEdge0NN(box, lat); break;
        case 1:  bc.addExternalVelocityEdge0NP(box, lat); break;
        case 2:  bc.addExternalVelocityEdge0PN(box, lat); break;
        case 3:  bc.addExternalVelocityEdge0PP(box, lat); break;
        case 4:  bc.addExternalVelocityEdge1NN(box, lat); break;
        case 5:  bc.addExternalVelocityEdge1NP(box, lat); break;
        case 6:  bc.addExternalVelocityEdge1PN(box, lat); break;
        case 7:  bc.addExternalVelocityEdge1PP(box, lat); break;
        case 8:  bc.addExternalVelocityEdge2NN(box, lat); break;
        case 9:  bc.addExternalVelocityEdge2NP(box, lat); break;
        case 10: bc.addExternalVelocityEdge2PN(box, lat); break;
        case 11: bc.addExternalVelocityEdge2PP(box, lat); break;
    }
}

template<template<typename U> class D>
static void addCorner(OnLatticeBoundaryCondition3D<T,D>& bc, MultiBlockLattice3D<T,D>& lat,
                      int nx, int ny, int nz, plint x, plint y, plint z)
{
    int code = (nx > 0 ? 4 : 0) + (ny > 0 ? 2 : 0) + (nz > 0 ? 1 : 0);
    switch (code) {
        case 0: bc.addExternalVelocityCornerNNN(x, y, z, lat); break;
        case 1: bc.addExternalVelocityCornerNNP(x, y, z, lat); break;
        case 2: bc.addExternalVelocityCornerNPN(x, y, z, lat); break;
        case 3: bc.addExternalVelocityCornerNPP(x, y, z, lat); break;
        case 4: bc.addExternalVelocityCornerPNN(x, y, z, lat); break;
        case 5: bc.addExternalVelocityCornerPNP(x, y, z, lat); break;
        case 6: bc.addExternalVelocityCornerPPN(x, y, z, lat); break;
        case 7: bc.addExternalVelocityCornerPPP(x, y, z, lat); break;
    }
}

template<template<typename U> class D>
static void addInnerEdge(OnLatticeBoundaryCondition3D<T,D>& bc, MultiBlockLattice3D<T,D>& lat,
                         int plane, int n1, int n2, Box3D box)
{
    int code = plane * 4 + (n1 > 0 ? 2 : 0) + (n2 > 0 ? 1 : 0);
    switch (code) {
        case 0:  bc.addInternalVelocityEdge0NN(box, lat); break;
        case 1:  bc.addInternalVelocityEdge0NP(box, lat); break;
        case 2:  bc.addInternalVelocityEdge0PN(box, lat); break;
        case 3:  bc.addInternalVelocityEdge0PP(box, lat); break;
        case 4:  bc.addInternalVelocityEdge1NN(box, lat); break;
        case 5:  bc.addInternalVelocityEdge1NP(box, lat); break;
        case 6:  bc.addInternalVelocityEdge1PN(box, lat); break;
        case 7:  bc.addInternalVelocityEdge1PP(box, lat); break;
        case 8:  bc.addInternalVelocityEdge2NN(box, lat); break;
        case 9:  bc.addInternalVelocityEdge2NP(box, lat); break;
        case 10: bc.addInternalVelocityEdge2PN(box, lat); break;
        case 11: bc.addInternalVelocityEdge2PP(box, lat); break;
    }
}

template<template<typename U> class D>
static void addInnerCorner(OnLatticeBoundaryCondition3D<T,D>& bc, MultiBlockLattice3D<T,D>& lat,
                           int nx, int ny, int nz, plint x, plint y, plint z)
{
    int code = (nx > 0 ? 4 : 0) + (ny > 0 ? 2 : 0) + (nz > 0 ? 1 : 0);
    switch (code) {
        case 0: bc.addInternalVelocityCornerNNN(x, y, z, lat); break;
        case 1: bc.addInternalVelocityCornerNNP(x, y, z, lat); break;
        case 2: bc.addInternalVelocityCornerNPN(x, y, z, lat); break;
        case 3: bc.addInternalVelocityCornerNPP(x, y, z, lat); break;
        case 4: bc.addInternalVelocityCornerPNN(x, y, z, lat); break;
        case 5: bc.addInternalVelocityCornerPNP(x, y, z, lat); break;
        case 6: bc.addInternalVelocityCornerPPN(x, y, z, lat); break;
        case 7: bc.addInternalVelocityCornerPPP(x, y, z, lat); break;
    }
}

template<template<typename U> class D>
static int runScenario(std::vector<std::string> const& lines, size_t first,
                       plint nx, plint ny, plint nz, std::string dyn, double p1, double p2,
                       plint bx, plint by, plint bz)
{
    const int q = D<T>::q;
    MultiBlockLattice3D<T,D>* latPtr;
    if (bx * by * bz > 1) {
        // Several blocks inside this one process: exercises the reference's envelope logic
        // (multiBlock/serialBlockCommunicator3D.cpp) -- SURVEY.md 8c.
        latPtr = new MultiBlockLattice3D<T,D>(
            MultiBlockManagement3D(createRegularDistribution3D(nx, ny, nz, bx, by, bz),
                                   defaultMultiBlockPolicy3D().getThreadAttribution(), 1),
            defaultMultiBlockPolicy3D().getBlockCommunicator(),
            defaultMultiBlockPolicy3D().getCombinedStatistics(),
            defaultMultiBlockPolicy3D().getMultiCellAccess<T,D>(),
            makeDynamics<D>(dyn, p1, p2));
    } else {
        latPtr = new MultiBlockLattice3D<T,D>(nx, ny, nz, makeDynamics<D>(dyn, p1, p2));
    }
    MultiBlockLattice3D<T,D>& lattice = *latPtr;
    std::vector<OnLatticeBoundaryCondition3D<T,D>*> bcs;
    plint iter = 0;
    // the three blocks ExternalRhoJcollideAndStream3D works on (src/variables.hh:475-486 of the fork's application)
    MultiScalarField3D<T>* rhoBarField = 0;
    MultiTensorField3D<T,3>* jField = 0;
    std::vector<MultiBlock3D*> rhoBarJarg;

    for (size_t iL = first; iL < lines.size(); ++iL) {
        std::istringstream in(lines[iL]);
        std::string cmd;
        if (!(in >> cmd) || cmd[0] == '#') continue;
        if (cmd == "periodic") {
            int px, py, pz; in >> px >> py >> pz;
            lattice.periodicity().toggle(0, px != 0);
            lattice.periodicity().toggle(1, py != 0);
            lattice.periodicity().toggle(2, pz != 0);
        } else if (cmd == "ext_integrate") {
            // src/variables.hh:475-486: rhoBar / j fields generated from the lattice, the external-moment sweep at
            // level 0 and the moment computation at level 3 (order of the calls matters: what is integrated later at
            // level 0 -- the boundary condition's processors -- runs after the sweep)
            rhoBarField = generateMultiScalarField<T>((MultiBlock3D&)lattice, 1).release();
            rhoBarField->toggleInternalStatistics(false);
            jField = generateMultiTensorField<T,3>((MultiBlock3D&)lattice, 1).release();
            jField->toggleInternalStatistics(false);
            for (int d = 0; d < 3; ++d) {
                rhoBarField->periodicity().toggle(d, lattice.periodicity().get(d));
                jField->periodicity().toggle(d, lattice.periodicity().get(d));
            }
            rhoBarJarg.clear();
            rhoBarJarg.push_back(&lattice); rhoBarJarg.push_back(rhoBarField); rhoBarJarg.push_back(jField);
            integrateProcessingFunctional(new ExternalRhoJcollideAndStream3D<T,D>(), lattice.getBoundingBox(), rhoBarJarg, 0);
            integrateProcessingFunctional(new BoxRhoBarJfunctional3D<T,D>(), lattice.getBoundingBox(), rhoBarJarg, 3);
        } else if (cmd == "ext_apply") {
            // applyProcessingFunctional(new BoxRhoBarJfunctional3D, box, rhoBarJarg) src/variables.hh:650
            Box3D box = readBox(in);
            if (!rhoBarField) fail("ext_apply before ext_integrate");
            applyProcessingFunctional(new BoxRhoBarJfunctional3D<T,D>(), box, rhoBarJarg);
        } else if (cmd == "ext_exec") {
            // viscosityTest/src/test.cpp:64-65: the time step of the fork's application
            plint n; in >> n;
            if (!rhoBarField) fail("ext_exec before ext_integrate");
            for (plint i = 0; i < n; ++i) { lattice.executeInternalProcessors(); lattice.incrementTime(); }
            iter += n;
        } else if (cmd == "ext_add_j") {
            // what a momentum-forcing processor does to the j field between two steps
            Box3D box = readBox(in); double dj[3]; in >> dj[0] >> dj[1] >> dj[2];
            if (!jField) fail("ext_add_j before ext_integrate");
            for (plint x = box.x0; x <= box.x1; ++x) for (plint y = box.y0; y <= box.y1; ++y) for (plint z = box.z0; z <= box.z1; ++z) {
                Array<T,3>& j = jField->get(x, y, z);
                j[0] += (T)dj[0]; j[1] += (T)dj[1]; j[2] += (T)dj[2];
            }
            jField->duplicateOverlaps(modif::staticVariables);
        } else if (cmd == "dump_rhobar_j") {
            std::string file; in >> file;
            if (!rhoBarField) fail("dump_rhobar_j before ext_integrate");
            std::vector<T> d((size_t)nx * ny * nz * 4);
            for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
                Array<T,3> const& j = jField->get(x, y, z);
                size_t o = (((size_t)x * ny + y) * nz + z) * 4;
                d[o] = rhoBarField->get(x, y, z); d[o+1] = j[0]; d[o+2] = j[1]; d[o+3] = j[2];
            }
            FILE* fp = std::fopen(file.c_str(), "wb");
            if (!fp) fail("cannot write " + file);
            std::fwrite(d.data(), sizeof(T), d.size(), fp);
            std::fclose(fp);
        } else if (cmd == "dynamics") {
            Box3D box = readBox(in);
            std::string kind; double a = 0, b = 0; in >> kind >> a >> b;
            defineDynamics(lattice, box, makeDynamics<D>(kind, a, b));
        } else if (cmd == "dynamics_mask") {
            std::string file, kind; int flag; double a = 0, b = 0;
            in >> file >> flag >> kind >> a >> b;
            std::vector<int> m = readBlob<int>(file, (size_t)nx * ny * nz);
            MultiScalarField3D<int> mask(lattice);
            for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z)
                mask.get(x, y, z) = m[((size_t)x * ny + y) * nz + z];
            defineDynamics(lattice, mask, makeDynamics<D>(kind, a, b), flag);
        } else if (cmd == "bc_box") {
            std::string kind; in >> kind;
            Box3D block = readBox(in), app = readBox(in);
            bcs.push_back(makeBC<D>(kind));
            bcs.back()->setVelocityConditionOnBlockBoundaries(lattice, block, app);
        } else if (cmd == "bc_vel" || cmd == "bc_pres") {
            std::string kind; int dir, ori; in >> kind >> dir >> ori;
            Box3D box = readBox(in);
            bcs.push_back(makeBC<D>(kind));
            addPlane<D>(*bcs.back(), lattice, cmd == "bc_vel", dir, ori, box);
        } else if (cmd == "bc_edge") {
            std::string kind; int plane, n1, n2; in >> kind >> plane >> n1 >> n2;
            Box3D box = readBox(in);
            bcs.push_back(makeBC<D>(kind));
            addEdge<D>(*bcs.back(), lattice, plane, n1, n2, box);
        } else if (cmd == "bc_corner") {
            std::string kind; int a, b, c; plint x, y, z; in >> kind >> a >> b >> c >> x >> y >> z;
            bcs.push_back(makeBC<D>(kind));
            addCorner<D>(*bcs.back(), lattice, a, b, c, x, y, z);
        } else if (cmd == "bc_iedge") {
            // inner (concave) edges and corners: RegularizedVelocityInnerEdge/CornerDynamics3D, regularizedBoundaryDynamics3D.hh:40-298
            std::string kind; int plane, n1, n2; in >> kind >> plane >> n1 >> n2;
            Box3D box = readBox(in);
            bcs.push_back(makeBC<D>(kind));
            addInnerEdge<D>(*bcs.back(), lattice, plane, n1, n2, box);
        } else if (cmd == "bc_icorner") {
            std::string kind; int a, b, c; plint x, y, z; in >> kind >> a >> b >> c >> x >> y >> z;
            bcs.push_back(makeBC<D>(kind));
            addInnerCorner<D>(*bcs.back(), lattice, a, b, c, x, y, z);
        } else if (cmd == "set_velocity") {
            Box3D box = readBox(in); double u[3]; in >> u[0] >> u[1] >> u[2];
            setBoundaryVelocity(lattice, box, Array<T,3>((T)u[0], (T)u[1], (T)u[2]));
        } else if (cmd == "set_density") {
            Box3D box = readBox(in); double rho; in >> rho;
            setBoundaryDensity(lattice, box, (T)rho);
        } else if (cmd == "init_eq") {
            Box3D box = readBox(in); double rho, u[3]; in >> rho >> u[0] >> u[1] >> u[2];
            initializeAtEquilibrium(lattice, box, (T)rho, Array<T,3>((T)u[0], (T)u[1], (T)u[2]));
        } else if (cmd == "init_eq_field") {
            Box3D box = readBox(in); std::string file; in >> file;
            std::vector<double> d = readBlob<double>(file, (size_t)box.nCells() * 4);
            initializeAtEquilibrium(lattice, box, FieldRhoU(d, box));
        } else if (cmd == "upload") {
            std::string file; in >> file;
            std::vector<T> d = readBlob<T>(file, (size_t)nx * ny * nz * q);
            for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
                Cell<T,D>& cell = lattice.get(x, y, z);
                for (int i = 0; i < q; ++i) cell[i] = d[(((size_t)x * ny + y) * nz + z) * q + i];
            }
            // lattice.get() hands out bulk cells only; refresh the envelopes like any
            // modification of static variables does in the reference.
            lattice.duplicateOverlaps(modif::staticVariables);
        } else if (cmd == "initialize") {
            lattice.initialize();
        } else if (cmd == "run") {
            plint n; in >> n;
            for (plint i = 0; i < n; ++i) lattice.collideAndStream();
            iter += n;
        } else if (cmd == "time_run") {
            plint n; in >> n;
            auto t0 = std::chrono::steady_clock::now();   // not global::timer (SURVEY.md sec. 5 bug)
            for (plint i = 0; i < n; ++i) lattice.collideAndStream();
            auto t1 = std::chrono::steady_clock::now();
            double sec = std::chrono::duration<double>(t1 - t0).count();
            iter += n;
            std::printf("TIME %ld %.9g %.6f\n", (long)n, sec, (double)nx * ny * nz * n / sec / 1e6);
        } else if (cmd == "stats") {
            std::printf("STATS %ld %.17g %.17g %.17g %.17g %.17g\n", (long)iter,
                        (double)getStoredAverageDensity(lattice), (double)getStoredAverageEnergy(lattice),
                        (double)getStoredMaxVelocity(lattice),
                        (double)computeAverageEnergy(lattice), (double)computeAverageDensity(lattice));
        } else if (cmd == "dump") {
            std::string file; in >> file;
            std::vector<T> d((size_t)nx * ny * nz * q);
            for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
                Cell<T,D> const& cell = lattice.get(x, y, z);
                for (int i = 0; i < q; ++i) d[(((size_t)x * ny + y) * nz + z) * q + i] = cell[i];
            }
            FILE* fp = std::fopen(file.c_str(), "wb");
            if (!fp) fail("cannot write " + file);
            std::fwrite(d.data(), sizeof(T), d.size(), fp);
            std::fclose(fp);
        } else if (cmd == "vtk") {
            // the writeVTK of the reference's examples: velocity norm, velocity and density as float, scaled
            std::string file; double dx, dt; in >> file >> dx >> dt;
            global::directories().setOutputDir("./");
            VtkImageOutput3D<T, float, 3> vtkOut(file, dx);
            vtkOut.writeData(*computeVelocityNorm(lattice), "velocityNorm", (float)(dx / dt));
            vtkOut.writeData(*computeVelocity(lattice), "velocity", (float)(dx / dt));
            vtkOut.writeData(*computeDensity(lattice), "density", (float)1, (float)-1);
        } else if (cmd == "save_block") {
            // saveBinaryBlock io/serializerIO_3D.cpp:37-49 (base64 stream of the block serializer, forward ordering)
            std::string file; in >> file;
            saveBinaryBlock(lattice, file);
        } else if (cmd == "load_block") {
            std::string file; in >> file;
            loadBinaryBlock(lattice, file);
        } else if (cmd == "psave") {
            // parallelIO::save io/multiBlockWriter3D.cpp:226-238: <file>.plb (XML) + <file>.dat (raw serialized blocks)
            std::string file; int dyn; in >> file >> dyn;
            global::directories().setOutputDir("./");
            parallelIO::save(lattice, file, dyn != 0);
        } else if (cmd == "pload") {
            std::string file; int dyn; in >> file >> dyn;
            global::directories().setOutputDir("./");
            parallelIO::load(file, lattice, dyn != 0);
        } else if (cmd == "save_state") {
            // saveState io/utilIO_3D.cpp:45-60
            plint it; int dyn; std::string xml, base; in >> it >> dyn >> xml >> base;
            global::directories().setOutputDir("./");
            std::vector<MultiBlock3D*> blocks; blocks.push_back(&lattice);
            saveState(blocks, it, dyn != 0, xml, base, 8);
        } else if (cmd == "load_state") {
            int dyn; std::string xml; in >> dyn >> xml;
            global::directories().setOutputDir("./");
            std::vector<MultiBlock3D*> blocks; blocks.push_back(&lattice);
            plint it = -1;
            loadState(blocks, it, dyn != 0, xml);
            std::printf("LOADED_ITERATION %ld\n", (long)it);
        } else if (cmd == "dump_rho_u") {
            // rho and u through Cell::computeDensity / computeVelocity (dynamics-aware).
            std::string file; in >> file;
            std::vector<T> d((size_t)nx * ny * nz * 4);
            for (plint x = 0; x < nx; ++x) for (plint y = 0; y < ny; ++y) for (plint z = 0; z < nz; ++z) {
                Cell<T,D> const& cell = lattice.get(x, y, z);
                Array<T,3> u; cell.computeVelocity(u);
                size_t o = (((size_t)x * ny + y) * nz + z) * 4;
                d[o] = cell.computeDensity(); d[o+1] = u[0]; d[o+2] = u[1]; d[o+3] = u[2];
            }
            FILE* fp = std::fopen(file.c_str(), "wb");
            if (!fp) fail("cannot write " + file);
            std::fwrite(d.data(), sizeof(T), d.size(), fp);
            std::fclose(fp);
        } else {
            fail("unknown command " + cmd);
        }
    }
    std::fflush(stdout);
    return 0;
}

int main(int argc, char* argv[]) {
    plbInit(&argc, &argv);
    if (argc < 2) fail("usage: ref_driver scenario.txt");
    std::ifstream f(argv[1]);
    if (!f) fail(std::string("cannot open ") + argv[1]);
    std::vector<std::string> lines;
    std::string line;
    while (std::getline(f, line)) lines.push_back(line);

    plint bx = 1, by = 1, bz = 1;
    for (size_t iL = 0; iL < lines.size(); ++iL) {
        std::istringstream in(lines[iL]);
        std::string cmd;
        if (!(in >> cmd) || cmd[0] == '#') continue;
        if (cmd == "blocks") { in >> bx >> by >> bz; continue; }
        if (cmd == "lattice") {
            std::string desc, prec, dyn; plint nx, ny, nz; double p1 = 0, p2 = 0;
            in >> desc >> prec >> nx >> ny >> nz >> dyn >> p1 >> p2;
            if ((prec == "f64") != (sizeof(T) == 8)) fail("precision mismatch: use ref_driver / ref_driver_f32");
            if (desc == "D3Q19")
                return runScenario<descriptors::D3Q19Descriptor>(lines, iL + 1, nx, ny, nz, dyn, p1, p2, bx, by, bz);
            if (desc == "D3Q27")
                return runScenario<descriptors::D3Q27Descriptor>(lines, iL + 1, nx, ny, nz, dyn, p1, p2, bx, by, bz);
            fail("unknown descriptor " + desc);
        }
        fail("scenario must start with [blocks] lattice");
    }
    return 0;
}
