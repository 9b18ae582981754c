// palabos_b200/csrc/engine.h -- host-side interface between the C-ABI (capi.cu) and the four
// instantiations LatticeImpl<T,Q> of the engine (engine_impl.cuh).  Failures are C++ exceptions
// here; capi.cu turns them into the 1/0 return convention of include/plb_b200.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>

namespace plb200 {

struct CreateDesc {
    int q, prec, nx, ny, nz;
    int dynamics;
    double p1, p2;
    int rank, nRanks, slabsPerRank;
};

struct PlbError : std::runtime_error {
    explicit PlbError(const std::string& m) : std::runtime_error(m) {}
};

#define PLB_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t plb_e_ = (expr);                                                           \
        if (plb_e_ != cudaSuccess)                                                             \
            throw ::plb200::PlbError(std::string(#expr) + ": " + cudaGetErrorString(plb_e_)); \
    } while (0)

class LatticeBase {
public:
    virtual ~LatticeBase() {}
    virtual void comm_init(const char id[128]) = 0;
    virtual void set_stream(cudaStream_t s) = 0;
    virtual void set_periodicity(int px, int py, int pz) = 0;
    virtual void toggle_statistics(bool on) = 0;
    virtual void define_dynamics_box(const int box[6], int dynamics, double p1, double p2) = 0;
    virtual void define_dynamics_mask(const int32_t* mask, int flag, int dynamics, double p1, double p2) = 0;
    virtual void add_velocity_boundary(int model, int dir, int ori, const int box[6]) = 0;
    virtual void add_pressure_boundary(int model, int dir, int ori, const int box[6]) = 0;
    virtual void add_external_velocity_edge(int model, int plane, int n1, int n2, const int box[6]) = 0;
    virtual void add_external_velocity_corner(int model, int xn, int yn, int zn, int x, int y, int z) = 0;
    virtual void add_internal_velocity_edge(int model, int plane, int n1, int n2, const int box[6]) = 0;
    virtual void add_internal_velocity_corner(int model, int xn, int yn, int zn, int x, int y, int z) = 0;
    virtual void set_velocity_condition_on_block_boundaries(int model, const int block[6], const int app[6]) = 0;
    virtual void set_boundary_velocity(const int box[6], const double u[3]) = 0;
    virtual void set_boundary_density(const int box[6], double rho) = 0;
    virtual void initialize_at_equilibrium(const int box[6], double rho, const double u[3]) = 0;
    virtual void initialize_at_equilibrium_field(const int box[6], const double* rho_u) = 0;
    virtual void initialize() = 0;
    virtual size_t local_bytes(const int box[6]) = 0;
    virtual void upload(const int box[6], const void* data, size_t bytes) = 0;
    virtual void download(const int box[6], void* data, size_t bytes) = 0;
    virtual void collide_and_stream(int n) = 0;
    virtual void enable_external_rhoj(const int bulk[6]) = 0;
    virtual void execute_internal_processors(int n) = 0;
    virtual void apply_rhobar_j(const int box[6]) = 0;
    virtual void transfer_rhobar_j(const int box[6], void* data, size_t bytes, bool up) = 0;
    virtual void synchronize() = 0;
    virtual void stored_statistics(double out3[3], uint64_t* n) = 0;
    virtual double average_energy(const int box[6]) = 0;
    virtual double average_density(const int box[6]) = 0;
    virtual void density_velocity(const int box[6], void* out, size_t bytes) = 0;
    virtual uint64_t launch_count() const = 0;
    virtual size_t device_bytes() const = 0;
    virtual void boundary_stages(int* nStages, int64_t* nItems) = 0;
    virtual void profile_sweeps(bool on) = 0;
    virtual void profile_read(double* totalMs, int* nLaunches) = 0;
    virtual const char* halo_transport() const = 0;
};

LatticeBase* make_lattice_f64_q19(const CreateDesc&);
LatticeBase* make_lattice_f64_q27(const CreateDesc&);
LatticeBase* make_lattice_f32_q19(const CreateDesc&);
LatticeBase* make_lattice_f32_q27(const CreateDesc&);

/// Record the message plb_b200_last_error() returns on this thread; returns 0 (the failure code of the ABI).
int set_last_error(const std::string& m);

/// Balanced x-slabs: slab s of n owns [s*nx/n, (s+1)*nx/n).
inline void slab_range(int nx, int nSlabs, int slab, int& x0, int& x1) {
    x0 = (int)((long long)slab * nx / nSlabs);
    x1 = (int)((long long)(slab + 1) * nx / nSlabs) - 1;
}

}  // namespace plb200
