// include/palabos_b200/palabos3D.h -- header-only C++ mirror of the Palabos classes that a user
// touches around ONE hot path, MultiBlockLattice3D::collideAndStream(), implemented on top of the
// C-ABI of libplb_b200.so (include/plb_b200.h).  A simulation written against the reference
//
//     MultiBlockLattice3D<T,DESCRIPTOR> lattice(nx, ny, nz, new BGKdynamics<T,DESCRIPTOR>(omega));
//     OnLatticeBoundaryCondition3D<T,DESCRIPTOR>* bc = createLocalBoundaryCondition3D<T,DESCRIPTOR>();
//     bc->setVelocityConditionOnBlockBoundaries(lattice);
//     setBoundaryVelocity(lattice, lattice.getBoundingBox(), Array<T,3>(0.,0.,0.));
//     initializeAtEquilibrium(lattice, lattice.getBoundingBox(), (T)1., Array<T,3>(0.,0.,0.));
//     lattice.initialize();
//     for (...) { lattice.collideAndStream(); }  getStoredAverageEnergy(lattice);  computeAverageEnergy(lattice);
//
// compiles unchanged with `#include "palabos_b200/palabos3D.h"` and `namespace plb`, and runs the
// sweep, the boundary processors, the envelope exchange and the statistics on the B200.  Same names,
// argument meaning and error behaviour (a failed call throws PlbLogicException, as the reference's
// plbLogicError does, core/runTimeDiagnostics.h:58-130).  Citations: paths under /root/reference/src.
//
// What is NOT here, on purpose: per-cell host access (lattice.get(x,y,z) returns a Cell& in the
// reference; populations live in HBM, use getPopulations / setPopulations on a Box3D instead), data
// processors written by the user, multi-grid, multi-physics, particles, I/O.
#ifndef PALABOS_B200_PALABOS3D_H
#define PALABOS_B200_PALABOS3D_H

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <limits>
#include <stdexcept>
#include <string>
#include <memory>
#include <vector>
#include "../plb_b200.h"

namespace plb {

typedef long plint;             // core/globalDefs.h
typedef std::size_t pluint;

/// core/runTimeDiagnostics.h:80-130
class PlbException : public std::exception {};
class PlbLogicException : public PlbException {
public:
    explicit PlbLogicException(std::string const& m) : msg(m) {}
    virtual ~PlbLogicException() throw() {}
    virtual const char* what() const throw() { return msg.c_str(); }
private:
    std::string msg;
};

namespace detail {
inline void check(int ok) { if (!ok) throw PlbLogicException(plb_b200_last_error()); }
}

/// core/array.h
template <typename T, int n>
class Array {
public:
    Array() { for (int i = 0; i < n; ++i) data[i] = T(); }
    Array(T a, T b, T c) { static_assert(n == 3, "3-component constructor"); data[0] = a; data[1] = b; data[2] = c; }
    T& operator[](int i) { return data[i]; }
    T const& operator[](int i) const { return data[i]; }
private:
    T data[n];
};

/// core/geometry3D.h
struct Box3D {
    Box3D() : x0(0), x1(0), y0(0), y1(0), z0(0), z1(0) {}
    Box3D(plint x0_, plint x1_, plint y0_, plint y1_, plint z0_, plint z1_) : x0(x0_), x1(x1_), y0(y0_), y1(y1_), z0(z0_), z1(z1_) {}
    plint getNx() const { return x1 - x0 + 1; }
    plint getNy() const { return y1 - y0 + 1; }
    plint getNz() const { return z1 - z0 + 1; }
    plint nCells() const { return getNx() * getNy() * getNz(); }
    Box3D enlarge(plint w) const { return Box3D(x0 - w, x1 + w, y0 - w, y1 + w, z0 - w, z1 + w); }
    bool operator==(Box3D const& r) const { return x0 == r.x0 && x1 == r.x1 && y0 == r.y0 && y1 == r.y1 && z0 == r.z0 && z1 == r.z1; }
    plint x0, x1, y0, y1, z0, z1;
};
namespace detail {
struct CBox { int v[6]; explicit CBox(Box3D const& b) { v[0] = (int)b.x0; v[1] = (int)b.x1; v[2] = (int)b.y0; v[3] = (int)b.y1; v[4] = (int)b.z0; v[5] = (int)b.z1; } };
}

/// latticeBoltzmann/nearestNeighborLattices3D.h
namespace descriptors {
template <typename T> struct D3Q19Descriptor { enum { d = 3, q = 19 }; static T fullRho(T rhoBar) { return rhoBar + (T)1; } };
template <typename T> struct D3Q27Descriptor { enum { d = 3, q = 27 }; static T fullRho(T rhoBar) { return rhoBar + (T)1; } };
}

namespace boundary { enum BcType { dirichlet }; }   // boundaryCondition/boundaryCondition.h (the others are out of scope)

// ----------------------------------------------------------------------------------- dynamics
// On the device a dynamics object is an entry (kind, parameters) of the lattice's dynamics table and
// a compiled-in functor; these classes only carry the description (core/dynamics.h:64-266).

template <typename T, template <typename U> class Descriptor>
class Dynamics {
public:
    virtual ~Dynamics() {}
    virtual Dynamics<T, Descriptor>* clone() const = 0;
    virtual int b200Kind() const = 0;
    virtual double b200P1() const = 0;
    virtual double b200P2() const { return 0.; }
    virtual T getOmega() const { return T(); }
};

#define PLB_B200_OMEGA_DYNAMICS(NAME, KIND)                                                          \
    template <typename T, template <typename U> class Descriptor>                                      \
    class NAME : public Dynamics<T, Descriptor> {                                                     \
    public:                                                                                            \
        explicit NAME(T omega_) : omega(omega_) {}                                                     \
        virtual NAME<T, Descriptor>* clone() const { return new NAME<T, Descriptor>(*this); }          \
        virtual int b200Kind() const { return KIND; }                                                  \
        virtual double b200P1() const { return (double)omega; }                                        \
        virtual T getOmega() const { return omega; }                                                   \
    private:                                                                                           \
        T omega;                                                                                       \
    };
PLB_B200_OMEGA_DYNAMICS(BGKdynamics, PLB_B200_DYN_BGK)                  // basicDynamics/isoThermalDynamics.h
PLB_B200_OMEGA_DYNAMICS(RegularizedBGKdynamics, PLB_B200_DYN_RLB)       // basicDynamics/isoThermalDynamics.h
PLB_B200_OMEGA_DYNAMICS(TRTdynamics, PLB_B200_DYN_TRT)                  // complexDynamics/trtDynamics.h
PLB_B200_OMEGA_DYNAMICS(IncBGKdynamics, PLB_B200_DYN_INCBGK)            // basicDynamics/isoThermalDynamics.h (rho0 = 1)
#undef PLB_B200_OMEGA_DYNAMICS

/// complexDynamics/smagorinskyDynamics.h: SmagorinskyBGKdynamics(omega0, cSmago)
template <typename T, template <typename U> class Descriptor>
class SmagorinskyBGKdynamics : public Dynamics<T, Descriptor> {
public:
    SmagorinskyBGKdynamics(T omega0_, T cSmago_) : omega0(omega0_), cSmago(cSmago_) {}
    virtual SmagorinskyBGKdynamics<T, Descriptor>* clone() const { return new SmagorinskyBGKdynamics<T, Descriptor>(*this); }
    virtual int b200Kind() const { return PLB_B200_DYN_SMAGO; }
    virtual double b200P1() const { return (double)omega0; }
    virtual double b200P2() const { return (double)cSmago; }
    virtual T getOmega() const { return omega0; }
private:
    T omega0, cSmago;
};

/// core/dynamics.h: BounceBack(rho = T()), NoDynamics(rho = 1)
template <typename T, template <typename U> class Descriptor>
class BounceBack : public Dynamics<T, Descriptor> {
public:
    explicit BounceBack(T rho_ = T()) : rho(rho_) {}
    virtual BounceBack<T, Descriptor>* clone() const { return new BounceBack<T, Descriptor>(*this); }
    virtual int b200Kind() const { return PLB_B200_DYN_BOUNCEBACK; }
    virtual double b200P1() const { return (double)rho; }
private:
    T rho;
};
template <typename T, template <typename U> class Descriptor>
class NoDynamics : public Dynamics<T, Descriptor> {
public:
    explicit NoDynamics(T rho_ = (T)1) : rho(rho_) {}
    virtual NoDynamics<T, Descriptor>* clone() const { return new NoDynamics<T, Descriptor>(*this); }
    virtual int b200Kind() const { return PLB_B200_DYN_NODYNAMICS; }
    virtual double b200P1() const { return (double)rho; }
private:
    T rho;
};

// ----------------------------------------------------------------------------------- lattice

template <typename T, template <typename U> class Descriptor> class MultiBlockLattice3D;

/// multiBlock/multiBlock3D.h: the common base of lattices and fields -- here only what a std::vector<MultiBlock3D*> argument
/// list of integrateProcessingFunctional / applyProcessingFunctional needs
class MultiBlock3D {
public:
    MultiBlock3D() : id(nextId()++) {}
    MultiBlock3D(MultiBlock3D const&) : id(nextId()++) {}
    virtual ~MultiBlock3D() {}
    virtual plint getNx() const = 0;
    virtual plint getNy() const = 0;
    virtual plint getNz() const = 0;
    /// multiBlock/multiBlock3D.cpp: the multi-blocks of a program are numbered in the order of their construction
    plint getId() const { return id; }
private:
    static plint& nextId() { static plint next = 0; return next; }
    plint id;
};

/// Tag of the lattice constructor for the time loop of the fork's own application (src/variables.hh:461-486): the lattice
/// will carry ExternalRhoJcollideAndStream3D, whose sweep collides and streams the reference's ENVELOPE layer as well
/// (basicDynamics/dynamicsProcessor3D.hh:118-119).  The engine has to allocate that layer, so it must know at construction;
/// periodicity is part of it (the envelope of a periodic direction is the periodic image).
struct WithEnvelope {
    WithEnvelope(bool periodicX = false, bool periodicY = false, bool periodicZ = false) { per[0] = periodicX; per[1] = periodicY; per[2] = periodicZ; }
    bool per[3];
};

/// multiBlock/multiBlock3D.h:69-90
template <typename T, template <typename U> class Descriptor>
class PeriodicitySwitch3D {
public:
    explicit PeriodicitySwitch3D(MultiBlockLattice3D<T, Descriptor>& l) : lattice(l) { per[0] = per[1] = per[2] = false; }
    void toggle(int direction, bool on) { per[direction] = on; push(); }
    void toggleAll(bool on) { per[0] = per[1] = per[2] = on; push(); }
    bool get(int direction) const { return per[direction]; }
    void b200Preset(const bool p[3]) { per[0] = p[0]; per[1] = p[1]; per[2] = p[2]; }
private:
    void push();
    MultiBlockLattice3D<T, Descriptor>& lattice;
    bool per[3];
};

/// multiBlock/multiBlockLattice3D.h:72-137.  The lattice takes ownership of backgroundDynamics, as
/// the reference does.  rank / numRanks / slabsPerRank describe the x-slab decomposition over the
/// GPUs of one box (one process per GPU); the defaults are one GPU.
template <typename T, template <typename U> class Descriptor>
class MultiBlockLattice3D : public MultiBlock3D {
public:
    /// the lattice of generateMultiBlockLattice(.., envelopeWidth, dynamics) + ExternalRhoJcollideAndStream3D (see WithEnvelope)
    MultiBlockLattice3D(plint nx_, plint ny_, plint nz_, Dynamics<T, Descriptor>* backgroundDynamics, WithEnvelope env,
                        int rank = 0, int numRanks = 1, int slabsPerRank = 1)
        : nx(nx_), ny(ny_), nz(nz_), background(backgroundDynamics), periodicitySwitch(*this), statisticsOn(true),
          enveloped(true), externalSweep(false), iterations(0) {
        handle = plb_b200_create_enveloped(Descriptor<T>::q, (int)sizeof(T), (int)nx, (int)ny, (int)nz,
                                           background->b200Kind(), background->b200P1(), background->b200P2(),
                                           env.per[0], env.per[1], env.per[2], rank, numRanks, slabsPerRank);
        if (!handle) { std::string m = plb_b200_last_error(); delete background; throw PlbLogicException(m); }
        periodicitySwitch.b200Preset(env.per);
    }
    MultiBlockLattice3D(plint nx_, plint ny_, plint nz_, Dynamics<T, Descriptor>* backgroundDynamics,
                        int rank = 0, int numRanks = 1, int slabsPerRank = 1)
        : nx(nx_), ny(ny_), nz(nz_), background(backgroundDynamics), periodicitySwitch(*this), statisticsOn(true),
          enveloped(false), externalSweep(false), iterations(0) {
        handle = plb_b200_create_distributed(Descriptor<T>::q, (int)sizeof(T), (int)nx, (int)ny, (int)nz,
                                             background->b200Kind(), background->b200P1(), background->b200P2(),
                                             rank, numRanks, slabsPerRank);
        if (!handle) { std::string m = plb_b200_last_error(); delete background; throw PlbLogicException(m); }
    }
    ~MultiBlockLattice3D() { plb_b200_destroy(handle); delete background; }

    plint getNx() const { return nx; }
    plint getNy() const { return ny; }
    plint getNz() const { return nz; }
    Box3D getBoundingBox() const { return Box3D(0, nx - 1, 0, ny - 1, 0, nz - 1); }
    Dynamics<T, Descriptor> const& getBackgroundDynamics() const { return *background; }
    PeriodicitySwitch3D<T, Descriptor>& periodicity() { return periodicitySwitch; }
    void toggleInternalStatistics(bool on) { statisticsOn = on; detail::check(plb_b200_toggle_internal_statistics(handle, on ? 1 : 0)); }
    bool isInternalStatisticsOn() const { return statisticsOn; }

    /// MultiBlock3D::initialize: run the internal processors once
    void initialize() { detail::check(plb_b200_initialize(handle)); }
    /// multiBlock/multiBlockLattice3D.hh:318-360
    void collideAndStream() { detail::check(plb_b200_collide_and_stream(handle, 1)); }
    /// n steps queued at once (the host does not wait)
    void collideAndStream(plint n) { detail::check(plb_b200_collide_and_stream(handle, (int)n)); }
    /// MultiBlock3D::executeInternalProcessors multiBlock/multiBlock3D.cpp:486-499 of a lattice that carries the external-moment
    /// sweep at level 0 and BoxRhoBarJfunctional3D at level 3: the time step of viscosityTest/src/test.cpp:64-65
    void executeInternalProcessors() {
        if (!externalSweep) throw PlbLogicException("executeInternalProcessors: integrate ExternalRhoJcollideAndStream3D and BoxRhoBarJfunctional3D first");
        detail::check(plb_b200_execute_internal_processors(handle, 1));
    }
    void incrementTime() { ++iterations; }
    plint getTimeCounter() const { return iterations; }
    bool b200Enveloped() const { return enveloped; }
    void b200ExternalSweepIntegrated() { externalSweep = true; }
    void synchronize() { detail::check(plb_b200_synchronize(handle)); }

    /// NCCL communicator for numRanks > 1: rank 0 calls getCommId(), the application broadcasts the
    /// 128 bytes (MPI_Bcast in Palabos), everybody calls initComm().
    static void getCommId(char id128[128]) { detail::check(plb_b200_comm_unique_id(id128)); }
    void initComm(const char id128[128]) { detail::check(plb_b200_comm_init(handle, id128)); }

    /// Populations of a box in the reference's serialization order ([x][y][z][q], f_i - t_i; what
    /// Cell::serialize / CoProcessor3D::receive produce, core/cell.hh:103-111).
    void getPopulations(Box3D domain, std::vector<T>& data) const {
        detail::CBox b(domain);
        data.resize(plb_b200_local_bytes(handle, b.v) / sizeof(T));
        detail::check(plb_b200_download(handle, b.v, data.data(), data.size() * sizeof(T)));
    }
    void setPopulations(Box3D domain, std::vector<T> const& data) {
        detail::CBox b(domain);
        detail::check(plb_b200_upload(handle, b.v, data.data(), data.size() * sizeof(T)));
    }
    /// density and velocity of every cell of a box ([cells][rho,ux,uy,uz]) through Cell::computeDensity / computeVelocity
    void getDensityVelocity(Box3D domain, std::vector<T>& rhoU) const {
        detail::CBox b(domain);
        rhoU.resize(plb_b200_local_bytes(handle, b.v) / sizeof(T) / Descriptor<T>::q * 4);
        detail::check(plb_b200_compute_density_velocity(handle, b.v, rhoU.data(), rhoU.size() * sizeof(T)));
    }

    plb_b200_lattice* b200Handle() const { return handle; }

private:
    MultiBlockLattice3D(MultiBlockLattice3D const&);
    MultiBlockLattice3D& operator=(MultiBlockLattice3D const&);
    plint nx, ny, nz;
    Dynamics<T, Descriptor>* background;
    plb_b200_lattice* handle;
    PeriodicitySwitch3D<T, Descriptor> periodicitySwitch;
    bool statisticsOn;
    bool enveloped, externalSweep;
    plint iterations;
};

template <typename T, template <typename U> class Descriptor>
void PeriodicitySwitch3D<T, Descriptor>::push() {
    detail::check(plb_b200_set_periodicity(lattice.b200Handle(), per[0], per[1], per[2]));
}

/// multiBlock/multiBlockLattice3D.hh:531-548
template <typename T, template <typename U> class Descriptor>
double getStoredAverageDensity(MultiBlockLattice3D<T, Descriptor> const& lattice) {
    double s[3];
    detail::check(plb_b200_get_stored_statistics(lattice.b200Handle(), s, 0));
    return (double)Descriptor<T>::fullRho((T)s[0]);
}
template <typename T, template <typename U> class Descriptor>
double getStoredAverageEnergy(MultiBlockLattice3D<T, Descriptor> const& lattice) {
    double s[3];
    detail::check(plb_b200_get_stored_statistics(lattice.b200Handle(), s, 0));
    return 0.5 * s[1];
}
template <typename T, template <typename U> class Descriptor>
double getStoredMaxVelocity(MultiBlockLattice3D<T, Descriptor> const& lattice) {
    double s[3];
    detail::check(plb_b200_get_stored_statistics(lattice.b200Handle(), s, 0));
    return std::sqrt(s[2]);
}

// ------------------------------------------------------- wrappers (dataProcessors/*Wrapper3D.h)

/// defineDynamics(lattice, domain, dynamics): the function takes ownership of `dynamics` (dataInitializerWrapper3D.h:163-165)
template <typename T, template <typename U> class Descriptor>
void defineDynamics(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain, Dynamics<T, Descriptor>* dynamics) {
    detail::CBox b(domain);
    int ok = plb_b200_define_dynamics_box(lattice.b200Handle(), b.v, dynamics->b200Kind(), dynamics->b200P1(), dynamics->b200P2());
    delete dynamics;
    detail::check(ok);
}
/// defineDynamics(lattice, intMask, dynamics, flag) (dataInitializerWrapper3D.h:194-200); the mask is a
/// plain int array [nx][ny][nz] here instead of a MultiScalarField3D<int>
template <typename T, template <typename U> class Descriptor>
void defineDynamics(MultiBlockLattice3D<T, Descriptor>& lattice, std::vector<int> const& intMask, Dynamics<T, Descriptor>* dynamics, int whichFlag) {
    int ok = 0;
    if ((plint)intMask.size() == lattice.getNx() * lattice.getNy() * lattice.getNz())
        ok = plb_b200_define_dynamics_mask(lattice.b200Handle(), intMask.data(), whichFlag, dynamics->b200Kind(), dynamics->b200P1(), dynamics->b200P2());
    delete dynamics;
    if (!ok && (plint)intMask.size() != lattice.getNx() * lattice.getNy() * lattice.getNz()) throw PlbLogicException("mask size does not match the lattice");
    detail::check(ok);
}
template <typename T, template <typename U> class Descriptor>
void setBoundaryVelocity(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain, Array<T, 3> velocity) {
    detail::CBox b(domain);
    const double u[3] = {(double)velocity[0], (double)velocity[1], (double)velocity[2]};
    detail::check(plb_b200_set_boundary_velocity(lattice.b200Handle(), b.v, u));
}
template <typename T, template <typename U> class Descriptor>
void setBoundaryDensity(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain, T rho) {
    detail::CBox b(domain);
    detail::check(plb_b200_set_boundary_density(lattice.b200Handle(), b.v, (double)rho));
}
template <typename T, template <typename U> class Descriptor>
void initializeAtEquilibrium(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain, T density, Array<T, 3> velocity) {
    detail::CBox b(domain);
    const double u[3] = {(double)velocity[0], (double)velocity[1], (double)velocity[2]};
    detail::check(plb_b200_initialize_at_equilibrium(lattice.b200Handle(), b.v, (double)density, u));
}
/// the functional overload: f(iX, iY, iZ, rho, u) evaluated on the host, equilibrium built on the device
template <typename T, template <typename U> class Descriptor, class RhoUFunction>
void initializeAtEquilibrium(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain, RhoUFunction f) {
    std::vector<double> ru((std::size_t)domain.nCells() * 4);
    std::size_t k = 0;
    for (plint x = domain.x0; x <= domain.x1; ++x)
        for (plint y = domain.y0; y <= domain.y1; ++y)
            for (plint z = domain.z0; z <= domain.z1; ++z, ++k) {
                T rho; Array<T, 3> u;
                f(x, y, z, rho, u);
                ru[4 * k] = (double)rho; ru[4 * k + 1] = (double)u[0]; ru[4 * k + 2] = (double)u[1]; ru[4 * k + 3] = (double)u[2];
            }
    detail::CBox b(domain);
    detail::check(plb_b200_initialize_at_equilibrium_field(lattice.b200Handle(), b.v, ru.data()));
}
/// dataProcessors/dataAnalysisWrapper3D.hh:1252-1262, :1226-1230
template <typename T, template <typename U> class Descriptor>
T computeAverageEnergy(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain) {
    detail::CBox b(domain);
    double out = 0.;
    detail::check(plb_b200_compute_average_energy(lattice.b200Handle(), b.v, &out));
    return (T)out;
}
template <typename T, template <typename U> class Descriptor>
T computeAverageEnergy(MultiBlockLattice3D<T, Descriptor>& lattice) { return computeAverageEnergy(lattice, lattice.getBoundingBox()); }
template <typename T, template <typename U> class Descriptor>
T computeAverageDensity(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain) {
    detail::CBox b(domain);
    double out = 0.;
    detail::check(plb_b200_compute_average_density(lattice.b200Handle(), b.v, &out));
    return (T)out;
}
template <typename T, template <typename U> class Descriptor>
T computeAverageDensity(MultiBlockLattice3D<T, Descriptor>& lattice) { return computeAverageDensity(lattice, lattice.getBoundingBox()); }

/// saveBinaryBlock / loadBinaryBlock io/serializerIO_3D.cpp:37-66 (files interchangeable with the reference's)
template <typename T, template <typename U> class Descriptor>
void saveBinaryBlock(MultiBlockLattice3D<T, Descriptor>& lattice, std::string fName, bool /*enforceUint*/ = false) {
    detail::check(plb_b200_save_binary_block(lattice.b200Handle(), fName.c_str()));
}
template <typename T, template <typename U> class Descriptor>
void loadBinaryBlock(MultiBlockLattice3D<T, Descriptor>& lattice, std::string fName, bool /*enforceUint*/ = false) {
    detail::check(plb_b200_load_binary_block(lattice.b200Handle(), fName.c_str()));
}

// ------------------------------------------- checkpoint / restart (io/plbFiles.h, io/multiBlockWriter3D.h, io/utilIO_3D.h)

/// io/plbFiles.cpp:35-120: path / name / extension of a file name
class FileName {
public:
    FileName() {}
    FileName(std::string file) { initialize(file); }
    FileName(const char* file) { initialize(file); }
    std::string get() const {
        std::string composite;
        if (path != "") composite = path + "/";
        composite += name;
        if (ext != "") composite += "." + ext;
        return composite;
    }
    operator std::string() const { return get(); }
    std::string getPath() const { return path; }
    std::string getName() const { return name; }
    std::string getExt() const { return ext; }
    FileName& setPath(std::string path_) {
        path = path_;
        if (!path.empty() && (path[path.size() - 1] == '/' || path[path.size() - 1] == '\\')) path.erase(path.end() - 1);
        return *this;
    }
    FileName& setName(std::string name_) { name = name_; return *this; }
    FileName& setExt(std::string ext_) {
        ext = ext_;
        if (!ext.empty() && ext[0] == '.') ext.erase(ext.begin());
        return *this;
    }
    FileName& defaultPath(std::string path_) { return path == "" ? setPath(path_) : *this; }
    FileName& defaultExt(std::string ext_) { return ext == "" ? setExt(ext_) : *this; }
private:
    void initialize(std::string file) {
        std::string fullName;
        const size_t sep = file.find_last_of("\\/");
        if (sep != std::string::npos) { fullName = file.substr(sep + 1); path = file.substr(0, sep); }
        else fullName = file;
        const size_t dot = fullName.find_last_of(".");
        if (dot != std::string::npos) { name = fullName.substr(0, dot); ext = fullName.substr(dot + 1); }
        else { name = fullName; ext = ""; }
    }
    std::string path, name, ext;
};

namespace global {
/// core/globalDefs.cpp:90-140
class Directories {
public:
    void setOutputDir(std::string d) { outputDir = d; }
    void setInputDir(std::string d) { inputDir = d; }
    std::string getOutputDir() const { return outputDir; }
    std::string getInputDir() const { return inputDir; }
private:
    std::string outputDir, inputDir;
};
inline Directories& directories() { static Directories instance; return instance; }
}  // namespace global

/// io/imageWriter.h:131-135
inline std::string createFileName(std::string name, plint number, plint width) {
    std::ostringstream s;
    s << name;
    s.fill('0'); s.width(width);
    s << number;
    return s.str();
}

namespace parallelIO {
/// parallelIO::save io/multiBlockWriter3D.cpp:226-238.  dynamicContent = true (the reference's default: the dynamics object of
/// every cell serialized next to its populations) is not available: the set-up of a lattice lives in the program.
template <typename T, template <typename U> class Descriptor>
void save(MultiBlockLattice3D<T, Descriptor>& lattice, FileName fName, bool dynamicContent = true) {
    if (dynamicContent) throw PlbLogicException("parallelIO::save: dynamicContent = true is not supported, pass false (populations only)");
    fName.defaultExt("plb");
    const std::string entry = FileName(fName).setExt("dat").get();
    FileName target(fName);
    target.defaultPath(global::directories().getOutputDir());
    detail::check(plb_b200_parallel_save(lattice.b200Handle(), FileName(target).setExt("").get().c_str(), entry.c_str(), (int)lattice.getId()));
}
/// parallelIO::load io/multiBlockReader3D.cpp:251-259
template <typename T, template <typename U> class Descriptor>
void load(FileName fName, MultiBlockLattice3D<T, Descriptor>& intoBlock, bool dynamicContent = true) {
    if (dynamicContent) throw PlbLogicException("parallelIO::load: dynamicContent = true is not supported, pass false (populations only)");
    fName.defaultPath(global::directories().getInputDir());
    detail::check(plb_b200_parallel_load(intoBlock.b200Handle(), FileName(fName).setExt("").get().c_str()));
}
}  // namespace parallelIO

/// saveState io/utilIO_3D.cpp:45-60: every block through parallelIO::save, and the XML file a restart reads
template <typename T, template <typename U> class Descriptor>
void saveState(std::vector<MultiBlockLattice3D<T, Descriptor>*> blocks, plint iteration, bool saveDynamicContent,
               FileName xmlFileName, FileName baseFileName, plint fileNamePadding = 8) {
    const std::string fname_base = createFileName(baseFileName.get(), iteration, fileNamePadding);
    for (size_t i = 0; i < blocks.size(); ++i) {
        std::ostringstream fname;
        fname << fname_base << "_" << i;
        parallelIO::save(*blocks[i], fname.str(), saveDynamicContent);
    }
    xmlFileName.defaultPath(global::directories().getOutputDir());   // (XMLwriter::print takes the name as it is; the usual call passes a plain name)
    std::ofstream xml(xmlFileName.get().c_str());
    if (!xml) throw PlbLogicException("saveState: cannot write " + xmlFileName.get());
    xml << "<?xml version=\"1.0\" ?>\n<continue>\n    <name> " << FileName(fname_base).defaultPath(global::directories().getOutputDir()).get() << " </name>\n"
        << "    <num_blocks> " << blocks.size() << " </num_blocks>\n    <iteration> " << iteration << " </iteration>\n</continue>\n";
}
/// loadState io/utilIO_3D.cpp:62-76
template <typename T, template <typename U> class Descriptor>
void loadState(std::vector<MultiBlockLattice3D<T, Descriptor>*> blocks, plint& iteration, bool saveDynamicContent, FileName xmlFileName) {
    std::ifstream in(xmlFileName.get().c_str());
    if (!in) throw PlbLogicException("loadState: cannot read " + xmlFileName.get());
    std::stringstream all;
    all << in.rdbuf();
    const std::string xml = all.str();
    auto field = [&](const std::string& tag) {
        const size_t a = xml.find("<" + tag + ">"), b = xml.find("</" + tag + ">");
        if (a == std::string::npos || b == std::string::npos) throw PlbLogicException("loadState: no <" + tag + "> in " + xmlFileName.get());
        std::string t = xml.substr(a + tag.size() + 2, b - a - tag.size() - 2);
        const size_t p = t.find_first_not_of(" \t\r\n"), q = t.find_last_not_of(" \t\r\n");
        return p == std::string::npos ? std::string() : t.substr(p, q - p + 1);
    };
    const std::string fname_base = field("name");
    const size_t numBlocks = (size_t)std::atol(field("num_blocks").c_str());
    iteration = (plint)std::atol(field("iteration").c_str());
    if (blocks.size() != numBlocks) throw PlbLogicException("loadState: the file holds another number of blocks");
    for (size_t i = 0; i < blocks.size(); ++i) {
        std::ostringstream fname;
        fname << fname_base << "_" << i;
        parallelIO::load(fname.str(), *blocks[i], saveDynamicContent);
    }
}

// ------------------------------------------- boundary conditions (boundaryCondition/boundaryCondition3D.h)

template <typename T, template <typename U> class Descriptor>
class OnLatticeBoundaryCondition3D {
public:
    explicit OnLatticeBoundaryCondition3D(int model_) : model(model_) {}
    virtual ~OnLatticeBoundaryCondition3D() {}
    virtual OnLatticeBoundaryCondition3D<T, Descriptor>* clone() const { return new OnLatticeBoundaryCondition3D<T, Descriptor>(model); }
    typedef MultiBlockLattice3D<T, Descriptor> Lattice;

#define PLB_B200_PLANE(DIR, ORI, SUFFIX)                                                                        \
    void addVelocityBoundary##SUFFIX(Box3D domain, Lattice& lattice, boundary::BcType = boundary::dirichlet) {   \
        detail::CBox b(domain); detail::check(plb_b200_add_velocity_boundary(lattice.b200Handle(), model, DIR, ORI, b.v)); } \
    void addPressureBoundary##SUFFIX(Box3D domain, Lattice& lattice, boundary::BcType = boundary::dirichlet) {   \
        detail::CBox b(domain); detail::check(plb_b200_add_pressure_boundary(lattice.b200Handle(), model, DIR, ORI, b.v)); }
    PLB_B200_PLANE(0, -1, 0N) PLB_B200_PLANE(0, +1, 0P)
    PLB_B200_PLANE(1, -1, 1N) PLB_B200_PLANE(1, +1, 1P)
    PLB_B200_PLANE(2, -1, 2N) PLB_B200_PLANE(2, +1, 2P)
#undef PLB_B200_PLANE

#define PLB_B200_EDGE(PLANE, N1, N2, SUFFIX)                                                                          \
    void addExternalVelocityEdge##SUFFIX(Box3D domain, Lattice& lattice, boundary::BcType = boundary::dirichlet) {      \
        detail::CBox b(domain); detail::check(plb_b200_add_external_velocity_edge(lattice.b200Handle(), model, PLANE, N1, N2, b.v)); }
    PLB_B200_EDGE(0, -1, -1, 0NN) PLB_B200_EDGE(0, -1, +1, 0NP) PLB_B200_EDGE(0, +1, -1, 0PN) PLB_B200_EDGE(0, +1, +1, 0PP)
    PLB_B200_EDGE(1, -1, -1, 1NN) PLB_B200_EDGE(1, -1, +1, 1NP) PLB_B200_EDGE(1, +1, -1, 1PN) PLB_B200_EDGE(1, +1, +1, 1PP)
    PLB_B200_EDGE(2, -1, -1, 2NN) PLB_B200_EDGE(2, -1, +1, 2NP) PLB_B200_EDGE(2, +1, -1, 2PN) PLB_B200_EDGE(2, +1, +1, 2PP)
#undef PLB_B200_EDGE

#define PLB_B200_CORNER(XN, YN, ZN, SUFFIX)                                                                                     \
    void addExternalVelocityCorner##SUFFIX(plint x, plint y, plint z, Lattice& lattice, boundary::BcType = boundary::dirichlet) {  \
        detail::check(plb_b200_add_external_velocity_corner(lattice.b200Handle(), model, XN, YN, ZN, (int)x, (int)y, (int)z)); }
    PLB_B200_CORNER(-1, -1, -1, NNN) PLB_B200_CORNER(-1, -1, +1, NNP) PLB_B200_CORNER(-1, +1, -1, NPN) PLB_B200_CORNER(-1, +1, +1, NPP)
    PLB_B200_CORNER(+1, -1, -1, PNN) PLB_B200_CORNER(+1, -1, +1, PNP) PLB_B200_CORNER(+1, +1, -1, PPN) PLB_B200_CORNER(+1, +1, +1, PPP)
#undef PLB_B200_CORNER

    // the concave edges and corners (boundaryCondition3D.h:123-206)
#define PLB_B200_IEDGE(PLANE, N1, N2, SUFFIX)                                                                         \
    void addInternalVelocityEdge##SUFFIX(Box3D domain, Lattice& lattice, boundary::BcType = boundary::dirichlet) {      \
        detail::CBox b(domain); detail::check(plb_b200_add_internal_velocity_edge(lattice.b200Handle(), model, PLANE, N1, N2, b.v)); }
    PLB_B200_IEDGE(0, -1, -1, 0NN) PLB_B200_IEDGE(0, -1, +1, 0NP) PLB_B200_IEDGE(0, +1, -1, 0PN) PLB_B200_IEDGE(0, +1, +1, 0PP)
    PLB_B200_IEDGE(1, -1, -1, 1NN) PLB_B200_IEDGE(1, -1, +1, 1NP) PLB_B200_IEDGE(1, +1, -1, 1PN) PLB_B200_IEDGE(1, +1, +1, 1PP)
    PLB_B200_IEDGE(2, -1, -1, 2NN) PLB_B200_IEDGE(2, -1, +1, 2NP) PLB_B200_IEDGE(2, +1, -1, 2PN) PLB_B200_IEDGE(2, +1, +1, 2PP)
#undef PLB_B200_IEDGE
#define PLB_B200_ICORNER(XN, YN, ZN, SUFFIX)                                                                                    \
    void addInternalVelocityCorner##SUFFIX(plint x, plint y, plint z, Lattice& lattice, boundary::BcType = boundary::dirichlet) {  \
        detail::check(plb_b200_add_internal_velocity_corner(lattice.b200Handle(), model, XN, YN, ZN, (int)x, (int)y, (int)z)); }
    PLB_B200_ICORNER(-1, -1, -1, NNN) PLB_B200_ICORNER(-1, -1, +1, NNP) PLB_B200_ICORNER(-1, +1, -1, NPN) PLB_B200_ICORNER(-1, +1, +1, NPP)
    PLB_B200_ICORNER(+1, -1, -1, PNN) PLB_B200_ICORNER(+1, -1, +1, PNP) PLB_B200_ICORNER(+1, +1, -1, PPN) PLB_B200_ICORNER(+1, +1, +1, PPP)
#undef PLB_B200_ICORNER

    /// boundaryCondition3D.h:416-434 (three overloads), implementation boundaryCondition3D.hh:64-162
    void setVelocityConditionOnBlockBoundaries(Lattice& lattice, boundary::BcType bcType = boundary::dirichlet) {
        setVelocityConditionOnBlockBoundaries(lattice, lattice.getBoundingBox(), lattice.getBoundingBox(), bcType);
    }
    void setVelocityConditionOnBlockBoundaries(Lattice& lattice, Box3D applicationDomain, boundary::BcType bcType = boundary::dirichlet) {
        setVelocityConditionOnBlockBoundaries(lattice, lattice.getBoundingBox(), applicationDomain, bcType);
    }
    void setVelocityConditionOnBlockBoundaries(Lattice& lattice, Box3D block, Box3D applicationDomain, boundary::BcType = boundary::dirichlet) {
        detail::CBox b(block), a(applicationDomain);
        detail::check(plb_b200_set_velocity_condition_on_block_boundaries(lattice.b200Handle(), model, b.v, a.v));
    }
private:
    int model;
};

/// boundaryCondition3D.hh:1162-1167 ("Wrapped regularized") and zouHeBoundary3D.hh:338-350
template <typename T, template <typename U> class Descriptor>
OnLatticeBoundaryCondition3D<T, Descriptor>* createLocalBoundaryCondition3D() { return new OnLatticeBoundaryCondition3D<T, Descriptor>(PLB_B200_BC_LOCAL); }
template <typename T, template <typename U> class Descriptor>
OnLatticeBoundaryCondition3D<T, Descriptor>* createZouHeBoundaryCondition3D() { return new OnLatticeBoundaryCondition3D<T, Descriptor>(PLB_B200_BC_ZOUHE); }
/// boundaryCondition3D.hh:1170-1181: the dynamics-based managers -- in this fork both complete the plane populations inside
/// collide() (regularized dynamics, no data processor); outer edges and corners as createLocalBoundaryCondition3D
template <typename T, template <typename U> class Descriptor>
OnLatticeBoundaryCondition3D<T, Descriptor>* createDynamicsBasedLocalBoundaryCondition3D() { return new OnLatticeBoundaryCondition3D<T, Descriptor>(PLB_B200_BC_DYNAMICS_LOCAL); }
/// boundaryCondition3D.hh:1184-1188: Dirichlet dynamics + the finite-difference plane processor (Skordos)
template <typename T, template <typename U> class Descriptor>
OnLatticeBoundaryCondition3D<T, Descriptor>* createInterpBoundaryCondition3D() { return new OnLatticeBoundaryCondition3D<T, Descriptor>(PLB_B200_BC_INTERP); }
template <typename T, template <typename U> class Descriptor>
OnLatticeBoundaryCondition3D<T, Descriptor>* createEquilibriumBoundaryCondition3D() { return new OnLatticeBoundaryCondition3D<T, Descriptor>(PLB_B200_BC_EQUILIBRIUM); }

// ------------------------------------------------ the reference's plug-in (coProcessors/coProcessor3D.h)

/// coProcessor3D.h:59-116, verbatim interface
template <typename T>
class CoProcessor3D {
public:
    virtual ~CoProcessor3D() {}
    virtual int addDomain(plint nx, plint ny, plint nz, T omega, int& domainHandle) = 0;
    virtual int send(int domainHandle, Box3D const& subDomain, std::vector<char> const& data) = 0;
    virtual int receive(int domainHandle, Box3D const& subDomain, std::vector<char>& data) const = 0;
    virtual int collideAndStream(int domainHandle) = 0;
};

}  // namespace plb
// B200CoProcessor3D<T>: the adapter itself lives in a header of its own, because it is what a maintainer of the REFERENCE includes
// (it needs nothing but the reference's CoProcessor3D<T>, Box3D and plint; examples/ref_coprocessor.cpp builds it against the
// unmodified Palabos sources)
#include "coprocessor_adapter.h"
namespace plb {

// ------------------------------------------- STL -> voxel flags (offLattice/)

enum Precision { FLT, DBL, LDBL, INF };                     // core/globalDefs.h:88-93

namespace voxelFlag {                                       // offLattice/voxelizer.h:38-48
static const int undetermined = 0, outside = 1, outerBorder = 2, inside = 3, innerBorder = 4;
inline bool insideFlag(int arg) { return arg == inside || arg == innerBorder; }
inline bool outsideFlag(int arg) { return arg == outside || arg == outerBorder; }
}

/// TriangleSet<T>(fname, precision) offLattice/triangleSet.h: the triangles of an ASCII or binary STL file
template <typename T>
class TriangleSet {
public:
    explicit TriangleSet(std::string fname, Precision precision_ = DBL) : precision(precision_) {
        double* p = 0; int64_t n = 0;
        detail::check(plb_b200_stl_read(fname.c_str(), &p, &n));
        soup.assign(p, p + 9 * n);
        plb_b200_stl_free(p);
    }
    plint getNumTriangles() const { return (plint)(soup.size() / 9); }
    std::vector<double> const& b200Soup() const { return soup; }
private:
    std::vector<double> soup;
    Precision precision;
};

/// DEFscaledMesh<T>(triangleSet, resolution, referenceDirection, margin, extraLayer) offLattice/triangleBoundary3D.h:44-82,
/// with TriangularSurfaceMesh::inflate() (triangularSurfaceMesh.hh:1419-1425) folded in as inflate().
template <typename T>
class DEFscaledMesh {
public:
    DEFscaledMesh(TriangleSet<T> const& set, plint resolution_, plint referenceDirection_, plint margin_, plint extraLayer_)
        : soup(set.b200Soup()), resolution(resolution_), referenceDirection(referenceDirection_), margin(margin_), extraLayer(extraLayer_) { build(0.); }
    void inflate(T amount = (T)1.e-3) { build((double)amount); }
    plint getMargin() const { return margin; }
    T getDx() const { return (T)dx; }
    Array<T, 3> getPhysicalLocation() const { return Array<T, 3>((T)location[0], (T)location[1], (T)location[2]); }
    plint getNumTriangles() const { return (plint)(soup.size() / 9); }
    std::vector<double> const& b200Triangles() const { return triangles; }
    const int* b200Dims() const { return dims; }
    plint b200Layer() const { return margin + extraLayer; }
private:
    void build(double amount) {
        triangles.resize(soup.size());
        detail::check(plb_b200_mesh_to_lattice(soup.data(), (int64_t)(soup.size() / 9), (int)resolution, (int)referenceDirection, (int)margin,
                                               (int)extraLayer, amount, triangles.data(), dims, &dx, location));
    }
    std::vector<double> soup;
    std::vector<double> triangles;
    plint resolution, referenceDirection, margin, extraLayer;
    int dims[3];
    double dx, location[3];
};

/// the slice of MultiScalarField3D<int> (multiBlock/multiDataField3D.h) a voxel matrix needs
template <typename T>
class MultiScalarField3D : public MultiBlock3D {
public:
    MultiScalarField3D(plint nx_, plint ny_, plint nz_, T ini = T()) : nx(nx_), ny(ny_), nz(nz_), data((size_t)(nx_ * ny_ * nz_), ini) {}
    plint getNx() const { return nx; }
    plint getNy() const { return ny; }
    plint getNz() const { return nz; }
    Box3D getBoundingBox() const { return Box3D(0, nx - 1, 0, ny - 1, 0, nz - 1); }
    T& get(plint x, plint y, plint z) { return data[(size_t)((x * ny + y) * nz + z)]; }
    T const& get(plint x, plint y, plint z) const { return data[(size_t)((x * ny + y) * nz + z)]; }
    std::vector<T>& b200Data() { return data; }
    std::vector<T> const& b200Data() const { return data; }
private:
    plint nx, ny, nz;
    std::vector<T> data;
};

/// the slice of MultiTensorField3D<T,nDim> (multiBlock/multiDataField3D.h) field extraction needs
template <typename T, int nDim>
class MultiTensorField3D : public MultiBlock3D {
public:
    MultiTensorField3D(plint nx_, plint ny_, plint nz_) : nx(nx_), ny(ny_), nz(nz_), data((size_t)(nx_ * ny_ * nz_)) {}
    plint getNx() const { return nx; }
    plint getNy() const { return ny; }
    plint getNz() const { return nz; }
    Array<T, nDim>& get(plint x, plint y, plint z) { return data[(size_t)((x * ny + y) * nz + z)]; }
    Array<T, nDim> const& get(plint x, plint y, plint z) const { return data[(size_t)((x * ny + y) * nz + z)]; }
private:
    plint nx, ny, nz;
    std::vector<Array<T, nDim> > data;
};

namespace detail {
/// rho and u of every cell of `domain` through Cell::computeDensity / computeVelocity (dynamics-aware), computed on the device
template <typename T, template <typename U> class Descriptor>
std::vector<T> densityVelocity(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain) {
    CBox b(domain);
    std::vector<T> ru((size_t)domain.nCells() * 4);
    check(plb_b200_compute_density_velocity(lattice.b200Handle(), b.v, ru.data(), ru.size() * sizeof(T)));
    return ru;
}
}

/// computeDensity / computeVelocity / computeVelocityNorm / computeVelocityComponent (dataProcessors/dataAnalysisWrapper3D.h:50-70,
/// dataAnalysisFunctional3D.hh: BoxDensityFunctional3D, BoxVelocityFunctional3D, BoxVelocityNormFunctional3D, BoxVelocityComponentFunctional3D);
/// the returned field covers `domain`, index (0,0,0) = domain's lower corner
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiScalarField3D<T> > computeDensity(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain) {
    std::vector<T> ru = detail::densityVelocity(lattice, domain);
    std::unique_ptr<MultiScalarField3D<T> > f(new MultiScalarField3D<T>(domain.getNx(), domain.getNy(), domain.getNz()));
    for (size_t k = 0; k < f->b200Data().size(); ++k) f->b200Data()[k] = ru[4 * k];
    return f;
}
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiScalarField3D<T> > computeDensity(MultiBlockLattice3D<T, Descriptor>& lattice) { return computeDensity(lattice, lattice.getBoundingBox()); }
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiTensorField3D<T, 3> > computeVelocity(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain) {
    std::vector<T> ru = detail::densityVelocity(lattice, domain);
    std::unique_ptr<MultiTensorField3D<T, 3> > f(new MultiTensorField3D<T, 3>(domain.getNx(), domain.getNy(), domain.getNz()));
    size_t k = 0;
    for (plint x = 0; x < domain.getNx(); ++x) for (plint y = 0; y < domain.getNy(); ++y) for (plint z = 0; z < domain.getNz(); ++z, ++k)
        f->get(x, y, z) = Array<T, 3>(ru[4 * k + 1], ru[4 * k + 2], ru[4 * k + 3]);
    return f;
}
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiTensorField3D<T, 3> > computeVelocity(MultiBlockLattice3D<T, Descriptor>& lattice) { return computeVelocity(lattice, lattice.getBoundingBox()); }
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiScalarField3D<T> > computeVelocityNorm(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain) {
    std::vector<T> ru = detail::densityVelocity(lattice, domain);
    std::unique_ptr<MultiScalarField3D<T> > f(new MultiScalarField3D<T>(domain.getNx(), domain.getNy(), domain.getNz()));
    for (size_t k = 0; k < f->b200Data().size(); ++k)      // std::sqrt(VectorTemplate<T,Descriptor>::normSqr(velocity))
        f->b200Data()[k] = std::sqrt(ru[4 * k + 1] * ru[4 * k + 1] + ru[4 * k + 2] * ru[4 * k + 2] + ru[4 * k + 3] * ru[4 * k + 3]);
    return f;
}
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiScalarField3D<T> > computeVelocityNorm(MultiBlockLattice3D<T, Descriptor>& lattice) { return computeVelocityNorm(lattice, lattice.getBoundingBox()); }
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiScalarField3D<T> > computeVelocityComponent(MultiBlockLattice3D<T, Descriptor>& lattice, Box3D domain, plint iComponent) {
    std::vector<T> ru = detail::densityVelocity(lattice, domain);
    std::unique_ptr<MultiScalarField3D<T> > f(new MultiScalarField3D<T>(domain.getNx(), domain.getNy(), domain.getNz()));
    for (size_t k = 0; k < f->b200Data().size(); ++k) f->b200Data()[k] = ru[4 * k + 1 + iComponent];
    return f;
}

// ---- the external-moment sweep of the fork's own application ---------------------------------------
// src/variables.hh:475-486:
//     rhoBar.reset(generateMultiScalarField<T>((MultiBlock3D&)*lattice, envelopeWidth).release());
//     j.reset(generateMultiTensorField<T,3>((MultiBlock3D&)*lattice, envelopeWidth).release());
//     rhoBarJarg = { lattice, rhoBar, j };
//     integrateProcessingFunctional(new ExternalRhoJcollideAndStream3D<T,Descriptor>(), lattice->getBoundingBox(), rhoBarJarg, 0);
//     integrateProcessingFunctional(new BoxRhoBarJfunctional3D<T,Descriptor>(),        lattice->getBoundingBox(), rhoBarJarg, 3);
// src/variables.hh:650:  applyProcessingFunctional(new BoxRhoBarJfunctional3D<T,Descriptor>(), lattice->getBoundingBox(), rhoBarJarg);
// viscosityTest/src/test.cpp:64-65:  lattice->executeInternalProcessors(); lattice->incrementTime();
// The fields live on the device next to the populations.  The host objects made by generateMulti*Field are views that are
// filled / sent on request (pullRhoBarJ / pushRhoBarJ) -- a per-step PCIe round trip would defeat the purpose.

/// basicDynamics/dynamicsProcessor3D.h: ExternalRhoJcollideAndStream3D (tag: the sweep itself runs on the device)
template <typename T, template <typename U> class Descriptor> class ExternalRhoJcollideAndStream3D {};
/// dataProcessors/dataAnalysisFunctional3D.h: BoxRhoBarJfunctional3D
template <typename T, template <typename U> class Descriptor> class BoxRhoBarJfunctional3D {};

/// multiBlock/multiBlockGenerator3D.hh:116-135, :668-688
template <typename T, template <typename U> class Descriptor>
std::unique_ptr<MultiScalarField3D<T> > generateMultiScalarField(MultiBlockLattice3D<T, Descriptor>& lattice, plint /*envelopeWidth*/ = 1) {
    return std::unique_ptr<MultiScalarField3D<T> >(new MultiScalarField3D<T>(lattice.getNx(), lattice.getNy(), lattice.getNz()));
}
template <typename T, int nDim, template <typename U> class Descriptor>
std::unique_ptr<MultiTensorField3D<T, nDim> > generateMultiTensorField(MultiBlockLattice3D<T, Descriptor>& lattice, plint /*envelopeWidth*/ = 1) {
    return std::unique_ptr<MultiTensorField3D<T, nDim> >(new MultiTensorField3D<T, nDim>(lattice.getNx(), lattice.getNy(), lattice.getNz()));
}

/// the fork writes generateMultiScalarField<T>((MultiBlock3D&) *lattice, envelopeWidth) (src/variables.hh:475-479): same through the base class
template <typename T>
std::unique_ptr<MultiScalarField3D<T> > generateMultiScalarField(MultiBlock3D& block, plint /*envelopeWidth*/ = 1) {
    return std::unique_ptr<MultiScalarField3D<T> >(new MultiScalarField3D<T>(block.getNx(), block.getNy(), block.getNz()));
}
template <typename T, int nDim>
std::unique_ptr<MultiTensorField3D<T, nDim> > generateMultiTensorField(MultiBlock3D& block, plint /*envelopeWidth*/ = 1) {
    return std::unique_ptr<MultiTensorField3D<T, nDim> >(new MultiTensorField3D<T, nDim>(block.getNx(), block.getNy(), block.getNz()));
}

namespace detail {
template <typename T, template <typename U> class Descriptor>
MultiBlockLattice3D<T, Descriptor>& rhoBarJlattice(std::vector<MultiBlock3D*> const& args, Box3D domain) {
    if (args.size() != 3) throw PlbLogicException("the argument list must be { lattice, rhoBar, j }");
    MultiBlockLattice3D<T, Descriptor>* lattice = dynamic_cast<MultiBlockLattice3D<T, Descriptor>*>(args[0]);
    if (!lattice || !dynamic_cast<MultiScalarField3D<T>*>(args[1]) || !dynamic_cast<MultiTensorField3D<T, 3>*>(args[2]))
        throw PlbLogicException("the argument list must be { lattice, rhoBar, j }");
    if (!lattice->b200Enveloped())
        throw PlbLogicException("construct the lattice with plb::WithEnvelope: the external-moment sweep works on the envelope layer too");
    (void)domain;
    return *lattice;
}
}

/// integrateProcessingFunctional(functional, domain, multiBlockArg, level) dataProcessors/dataProcessorWrapper3D.h, for the two
/// functionals of this path; takes ownership of the functional like the reference.  The sweep must be integrated on the
/// whole lattice and before the boundary condition (whose level-0 processors then run after it).
template <typename T, template <typename U> class Descriptor>
void integrateProcessingFunctional(ExternalRhoJcollideAndStream3D<T, Descriptor>* functional, Box3D domain,
                                   std::vector<MultiBlock3D*> multiBlockArg, plint level = 0) {
    delete functional;
    MultiBlockLattice3D<T, Descriptor>& lattice = detail::rhoBarJlattice<T, Descriptor>(multiBlockArg, domain);
    if (!(domain == lattice.getBoundingBox()) || level != 0)
        throw PlbLogicException("ExternalRhoJcollideAndStream3D: integrate on the bounding box at level 0");
}
template <typename T, template <typename U> class Descriptor>
void integrateProcessingFunctional(BoxRhoBarJfunctional3D<T, Descriptor>* functional, Box3D domain,
                                   std::vector<MultiBlock3D*> multiBlockArg, plint level = 0) {
    delete functional;
    MultiBlockLattice3D<T, Descriptor>& lattice = detail::rhoBarJlattice<T, Descriptor>(multiBlockArg, domain);
    if (!(domain == lattice.getBoundingBox()) || level < 1)
        throw PlbLogicException("BoxRhoBarJfunctional3D: integrate on the bounding box at a level above the sweep's");
    lattice.b200ExternalSweepIntegrated();
}
/// applyProcessingFunctional(new BoxRhoBarJfunctional3D, domain, rhoBarJarg): the moments of the current populations into the fields
template <typename T, template <typename U> class Descriptor>
void applyProcessingFunctional(BoxRhoBarJfunctional3D<T, Descriptor>* functional, Box3D domain, std::vector<MultiBlock3D*> multiBlockArg) {
    delete functional;
    MultiBlockLattice3D<T, Descriptor>& lattice = detail::rhoBarJlattice<T, Descriptor>(multiBlockArg, domain);
    detail::CBox b(domain);
    detail::check(plb_b200_apply_rhobar_j(lattice.b200Handle(), b.v));
}
/// device fields -> host views (whole lattice)
template <typename T, template <typename U> class Descriptor>
void pullRhoBarJ(MultiBlockLattice3D<T, Descriptor>& lattice, MultiScalarField3D<T>& rhoBar, MultiTensorField3D<T, 3>& j) {
    detail::CBox b(lattice.getBoundingBox());
    std::vector<T> w((size_t)lattice.getBoundingBox().nCells() * 4);
    detail::check(plb_b200_download_rhobar_j(lattice.b200Handle(), b.v, w.data(), w.size() * sizeof(T)));
    size_t k = 0;
    for (plint x = 0; x < lattice.getNx(); ++x) for (plint y = 0; y < lattice.getNy(); ++y) for (plint z = 0; z < lattice.getNz(); ++z, ++k) {
        rhoBar.get(x, y, z) = w[4 * k];
        j.get(x, y, z) = Array<T, 3>(w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
}
/// host views -> device fields (whole lattice): what a processor that forces the momentum between two steps ends with
template <typename T, template <typename U> class Descriptor>
void pushRhoBarJ(MultiBlockLattice3D<T, Descriptor>& lattice, MultiScalarField3D<T> const& rhoBar, MultiTensorField3D<T, 3> const& j) {
    detail::CBox b(lattice.getBoundingBox());
    std::vector<T> w((size_t)lattice.getBoundingBox().nCells() * 4);
    size_t k = 0;
    for (plint x = 0; x < lattice.getNx(); ++x) for (plint y = 0; y < lattice.getNy(); ++y) for (plint z = 0; z < lattice.getNz(); ++z, ++k) {
        w[4 * k] = rhoBar.get(x, y, z);
        for (int d = 0; d < 3; ++d) w[4 * k + 1 + d] = j.get(x, y, z)[d];
    }
    detail::check(plb_b200_upload_rhobar_j(lattice.b200Handle(), b.v, w.data(), w.size() * sizeof(T)));
}

/// VtkImageOutput3D<T,TConv,n> io/vtkDataOutput.h:100-136, io/vtkDataOutput.hh:57-392, io/vtkDataOutput.cpp:39-100 (this fork
/// templates the class, not the methods): VTK ImageData XML, every field converted to TConv, optionally scaled / shifted in
/// TConv arithmetic, written x fastest (IndexOrdering::backward) as base64 of a 32-bit byte count followed by base64 of the data.
/// Files are byte-identical to the reference's.
template <typename T, typename TConv = float, plint n = 3>
class VtkImageOutput3D {
public:
    explicit VtkImageOutput3D(std::string fName, double deltaX_ = 1.)
        : ostr((fName + ".vti").c_str()), deltaX(deltaX_), headerWritten(false) { offset[0] = offset[1] = offset[2] = 0.; }
    VtkImageOutput3D(std::string fName, double deltaX_, Array<double, 3> offset_)
        : ostr((fName + ".vti").c_str()), deltaX(deltaX_), headerWritten(false) { for (int d = 0; d < 3; ++d) offset[d] = offset_[d]; }
    ~VtkImageOutput3D() { if (headerWritten) ostr << "</PointData>\n</Piece>\n</ImageData>\n</VTKFile>\n"; }
    void writeData(MultiScalarField3D<T>& f, std::string name, TConv scalingFactor = (TConv)1, TConv additiveOffset = (TConv)0) {
        std::vector<TConv> v((size_t)(f.getNx() * f.getNy() * f.getNz()));
        size_t k = 0;
        for (plint z = 0; z < f.getNz(); ++z) for (plint y = 0; y < f.getNy(); ++y) for (plint x = 0; x < f.getNx(); ++x, ++k) {
            TConv c = (TConv)f.get(x, y, z);
            if (!isOne(scalingFactor)) c = c * scalingFactor;
            if (!isZero(additiveOffset)) c = c + additiveOffset;
            v[k] = c;
        }
        field(f.getNx(), f.getNy(), f.getNz(), 1, v, name);
    }
    void writeData(MultiTensorField3D<T, n>& f, std::string name, TConv scalingFactor = (TConv)1) {
        std::vector<TConv> v((size_t)(f.getNx() * f.getNy() * f.getNz() * n));
        size_t k = 0;
        for (plint z = 0; z < f.getNz(); ++z) for (plint y = 0; y < f.getNy(); ++y) for (plint x = 0; x < f.getNx(); ++x)
            for (int d = 0; d < n; ++d, ++k) {
                TConv c = (TConv)f.get(x, y, z)[d];
                if (!isOne(scalingFactor)) c = c * scalingFactor;
                v[k] = c;
            }
        field(f.getNx(), f.getNy(), f.getNz(), n, v, name);
    }
private:
    // util::isZero / isOne core/util.h:583-596 with getEpsilon<T>() = 10 * machine epsilon of the type (core/globalDefs.h:104-145)
    static TConv eps() { return (TConv)10.0 * std::numeric_limits<TConv>::epsilon(); }
    static bool isZero(TConv x) { return std::fabs(x) <= eps(); }
    static bool isOne(TConv x) { return std::fabs(x - (TConv)1) <= eps(); }
    void field(plint nx, plint ny, plint nz, plint nDim, std::vector<TConv> const& v, std::string const& name) {
        if (!headerWritten) {
            ostr << "<?xml version=\"1.0\"?>\n<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
            ostr << "<ImageData WholeExtent=\"0 " << nx - 1 << " 0 " << ny - 1 << " 0 " << nz - 1 << "\" Origin=\""
                 << offset[0] << " " << offset[1] << " " << offset[2] << "\" Spacing=\"" << deltaX << " " << deltaX << " " << deltaX << "\">\n";
            ostr << "<Piece Extent=\"0 " << nx - 1 << " 0 " << ny - 1 << " 0 " << nz - 1 << "\">\n<PointData>\n";
            headerWritten = true;
        }
        ostr << "<DataArray type=\"" << (sizeof(TConv) == 4 ? "Float32" : "Float64") << "\" Name=\"" << name << "\" format=\"binary\" encoding=\"base64";
        if (nDim > 1) ostr << "\" NumberOfComponents=\"" << nDim;
        ostr << "\">\n";
        const unsigned int bytes = (unsigned int)(v.size() * sizeof(TConv));
        base64(reinterpret_cast<const unsigned char*>(&bytes), sizeof bytes);
        base64(reinterpret_cast<const unsigned char*>(v.data()), bytes);
        ostr << "\n</DataArray>\n";
    }
    void base64(const unsigned char* d, size_t len) {
        static const char* e = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        std::string out;
        out.reserve(len / 3 * 4 + 4);
        size_t i = 0;
        for (; i + 3 <= len; i += 3) {
            out += e[d[i] >> 2]; out += e[((d[i] & 0x03) << 4) | ((d[i + 1] & 0xf0) >> 4)];
            out += e[((d[i + 1] & 0x0f) << 2) | ((d[i + 2] & 0xc0) >> 6)]; out += e[d[i + 2] & 0x3f];
        }
        if (i < len) {
            const unsigned char a = d[i], b = i + 1 < len ? d[i + 1] : 0;
            out += e[a >> 2]; out += e[((a & 0x03) << 4) | ((b & 0xf0) >> 4)];
            out += (i + 1 < len) ? e[(b & 0x0f) << 2] : '=';
            out += '=';
        }
        ostr << out;
    }
    std::ofstream ostr;
    double deltaX, offset[3];
    bool headerWritten;
};

/// voxelize(mesh, symmetricLayer, borderWidth) offLattice/voxelizer.hh:84-104 (+ detectBorderLine :640-690), on the device
template <typename T>
std::unique_ptr<MultiScalarField3D<int> > voxelize(DEFscaledMesh<T> const& mesh, plint /*symmetricLayer*/, plint borderWidth) {
    const int* d = mesh.b200Dims();
    std::unique_ptr<MultiScalarField3D<int> > m(new MultiScalarField3D<int>(d[0], d[1], d[2], voxelFlag::undetermined));
    detail::check(plb_b200_voxelize(mesh.b200Triangles().data(), (int64_t)mesh.getNumTriangles(), d[0], d[1], d[2], (int)borderWidth,
                                    m->b200Data().data(), 0));
    return m;
}

/// defineDynamics(lattice, voxelMatrix, dynamics, flag) with the voxel matrix placed at `offset` in the lattice
/// (the reference's matrix and lattice share a domain; cells outside the matrix keep their dynamics)
template <typename T, template <typename U> class Descriptor>
void defineDynamics(MultiBlockLattice3D<T, Descriptor>& lattice, MultiScalarField3D<int> const& voxels, Array<plint, 3> offset,
                    Dynamics<T, Descriptor>* dynamics, int whichFlag) {
    std::vector<int> mask((size_t)(lattice.getNx() * lattice.getNy() * lattice.getNz()), whichFlag == 0 ? 1 : 0);
    for (plint x = 0; x < voxels.getNx(); ++x) for (plint y = 0; y < voxels.getNy(); ++y) for (plint z = 0; z < voxels.getNz(); ++z) {
        const plint X = x + offset[0], Y = y + offset[1], Z = z + offset[2];
        if (X < 0 || Y < 0 || Z < 0 || X >= lattice.getNx() || Y >= lattice.getNy() || Z >= lattice.getNz()) continue;
        mask[(size_t)((X * lattice.getNy() + Y) * lattice.getNz() + Z)] = voxels.get(x, y, z);
    }
    defineDynamics(lattice, mask, dynamics, whichFlag);
}

namespace global {
template <typename T>
inline CoProcessor3D<T>& defaultCoProcessor3D() {
    static B200CoProcessor3D<T> singleton;
    return singleton;
}
}

}  // namespace plb
#endif
