// palabos_b200/csrc/capi.cu -- the extern "C" surface declared in include/plb_b200.h.
// Every entry point catches C++ exceptions and returns 1/0 like the reference's CoProcessor3D
// methods (coProcessors/coProcessor3D.h:77,93,103,115).
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include "../../include/plb_b200.h"
#include "engine.h"
#include "nccl_dyn.h"

using namespace plb200;

struct plb_b200_lattice { LatticeBase* impl; };

static thread_local std::string g_err;

static int fail(const std::string& m) { g_err = m; return 0; }
namespace plb200 { int set_last_error(const std::string& m) { return fail(m); } }

#define PLB_TRY(L, body)                                           \
    if (!(L) || !(L)->impl) return fail("null lattice handle");    \
    try { body; return 1; }                                        \
    catch (const std::exception& e) { return fail(e.what()); }     \
    catch (...) { return fail("unknown error"); }

extern "C" {

const char* plb_b200_last_error(void) { return g_err.c_str(); }
int plb_b200_abi_version(void) { return PLB_B200_ABI_VERSION; }

plb_b200_lattice* plb_b200_create_distributed(int q, int prec, int nx, int ny, int nz, int dynamics, double p1, double p2,
                                              int rank, int nRanks, int slabsPerRank) {
    try {
        if ((q != 19 && q != 27) || (prec != 4 && prec != 8) || nx < 1 || ny < 1 || nz < 1)
            throw PlbError("bad lattice descriptor");
        CreateDesc d;
        d.q = q; d.prec = prec; d.nx = nx; d.ny = ny; d.nz = nz; d.dynamics = dynamics; d.p1 = p1; d.p2 = p2;
        d.rank = rank; d.nRanks = nRanks; d.slabsPerRank = slabsPerRank;
        LatticeBase* impl = prec == 8 ? (q == 19 ? make_lattice_f64_q19(d) : make_lattice_f64_q27(d))
                                      : (q == 19 ? make_lattice_f32_q19(d) : make_lattice_f32_q27(d));
        plb_b200_lattice* L = new plb_b200_lattice;
        L->impl = impl;
        return L;
    } catch (const std::exception& e) { fail(e.what()); return nullptr; }
    catch (...) { fail("unknown error"); return nullptr; }
}

plb_b200_lattice* plb_b200_create(int q, int prec, int nx, int ny, int nz, int dynamics, double p1, double p2) {
    return plb_b200_create_distributed(q, prec, nx, ny, nz, dynamics, p1, p2, 0, 1, 1);
}

void plb_b200_destroy(plb_b200_lattice* L) {
    if (!L) return;
    try { delete L->impl; } catch (...) {}
    delete L;
}

int plb_b200_comm_unique_id(char id128[128]) {
    try {
        NcclApi& n = NcclApi::get();
        ncclUniqueId uid;
        n.check(n.GetUniqueId(&uid), "ncclGetUniqueId");
        std::memcpy(id128, &uid, 128);
        return 1;
    } catch (const std::exception& e) { return fail(e.what()); }
}
int plb_b200_comm_init(plb_b200_lattice* L, const char id128[128]) { PLB_TRY(L, L->impl->comm_init(id128)) }

int plb_b200_slab_range(int nx, int nSlabs, int slab, int* x0, int* x1) {
    if (nx < 1 || nSlabs < 1 || nSlabs > nx || slab < 0 || slab >= nSlabs || !x0 || !x1) return fail("bad slab query");
    slab_range(nx, nSlabs, slab, *x0, *x1);
    return 1;
}

int plb_b200_set_stream(plb_b200_lattice* L, void* s) { PLB_TRY(L, L->impl->set_stream((cudaStream_t)s)) }
int plb_b200_set_periodicity(plb_b200_lattice* L, int px, int py, int pz) { PLB_TRY(L, L->impl->set_periodicity(px, py, pz)) }
int plb_b200_toggle_internal_statistics(plb_b200_lattice* L, int on) { PLB_TRY(L, L->impl->toggle_statistics(on != 0)) }
int plb_b200_define_dynamics_box(plb_b200_lattice* L, const int box[6], int dyn, double p1, double p2) {
    PLB_TRY(L, L->impl->define_dynamics_box(box, dyn, p1, p2))
}
int plb_b200_define_dynamics_mask(plb_b200_lattice* L, const int32_t* mask, int flag, int dyn, double p1, double p2) {
    PLB_TRY(L, L->impl->define_dynamics_mask(mask, flag, dyn, p1, p2))
}
int plb_b200_add_velocity_boundary(plb_b200_lattice* L, int model, int dir, int ori, const int box[6]) {
    PLB_TRY(L, L->impl->add_velocity_boundary(model, dir, ori, box))
}
int plb_b200_add_pressure_boundary(plb_b200_lattice* L, int model, int dir, int ori, const int box[6]) {
    PLB_TRY(L, L->impl->add_pressure_boundary(model, dir, ori, box))
}
int plb_b200_add_external_velocity_edge(plb_b200_lattice* L, int model, int plane, int n1, int n2, const int box[6]) {
    PLB_TRY(L, L->impl->add_external_velocity_edge(model, plane, n1, n2, box))
}
int plb_b200_add_external_velocity_corner(plb_b200_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z) {
    PLB_TRY(L, L->impl->add_external_velocity_corner(model, xn, yn, zn, x, y, z))
}
int plb_b200_set_velocity_condition_on_block_boundaries(plb_b200_lattice* L, int model, const int block[6], const int app[6]) {
    PLB_TRY(L, L->impl->set_velocity_condition_on_block_boundaries(model, block, app))
}
int plb_b200_set_boundary_velocity(plb_b200_lattice* L, const int box[6], const double u[3]) {
    PLB_TRY(L, L->impl->set_boundary_velocity(box, u))
}
int plb_b200_set_boundary_density(plb_b200_lattice* L, const int box[6], double rho) {
    PLB_TRY(L, L->impl->set_boundary_density(box, rho))
}
int plb_b200_initialize_at_equilibrium(plb_b200_lattice* L, const int box[6], double rho, const double u[3]) {
    PLB_TRY(L, L->impl->initialize_at_equilibrium(box, rho, u))
}
int plb_b200_initialize_at_equilibrium_field(plb_b200_lattice* L, const int box[6], const double* ru) {
    PLB_TRY(L, L->impl->initialize_at_equilibrium_field(box, ru))
}
int plb_b200_initialize(plb_b200_lattice* L) { PLB_TRY(L, L->impl->initialize()) }

int plb_b200_upload(plb_b200_lattice* L, const int box[6], const void* data, size_t bytes) {
    PLB_TRY(L, L->impl->upload(box, data, bytes))
}
int plb_b200_download(plb_b200_lattice* L, const int box[6], void* data, size_t bytes) {
    PLB_TRY(L, L->impl->download(box, data, bytes))
}
size_t plb_b200_local_bytes(plb_b200_lattice* L, const int box[6]) {
    if (!L || !L->impl) { fail("null lattice handle"); return 0; }
    try { return L->impl->local_bytes(box); } catch (const std::exception& e) { fail(e.what()); return 0; }
}

int plb_b200_collide_and_stream(plb_b200_lattice* L, int n) { PLB_TRY(L, L->impl->collide_and_stream(n)) }
int plb_b200_synchronize(plb_b200_lattice* L) { PLB_TRY(L, L->impl->synchronize()) }

int plb_b200_get_stored_statistics(plb_b200_lattice* L, double out3[3], uint64_t* n) {
    PLB_TRY(L, L->impl->stored_statistics(out3, n))
}
int plb_b200_compute_average_energy(plb_b200_lattice* L, const int box[6], double* out) {
    PLB_TRY(L, *out = L->impl->average_energy(box))
}
int plb_b200_compute_average_density(plb_b200_lattice* L, const int box[6], double* out) {
    PLB_TRY(L, *out = L->impl->average_density(box))
}
int plb_b200_compute_density_velocity(plb_b200_lattice* L, const int box[6], void* out, size_t bytes) {
    PLB_TRY(L, L->impl->density_velocity(box, out, bytes))
}

uint64_t plb_b200_launch_count(plb_b200_lattice* L) { return (L && L->impl) ? L->impl->launch_count() : 0; }
size_t plb_b200_device_bytes(plb_b200_lattice* L) { return (L && L->impl) ? L->impl->device_bytes() : 0; }
int plb_b200_boundary_stages(plb_b200_lattice* L, int* nStages, int64_t* nItems) {
    PLB_TRY(L, L->impl->boundary_stages(nStages, nItems))
}

int plb_b200_profile_sweeps(plb_b200_lattice* L, int on) { PLB_TRY(L, L->impl->profile_sweeps(on != 0)) }
int plb_b200_profile_read(plb_b200_lattice* L, double* totalMs, int* nLaunches) { PLB_TRY(L, L->impl->profile_read(totalMs, nLaunches)) }

// ---- plb::CoProcessor3D<T> surface (coProcessors/coProcessor3D.h:59-116) -------------------------

static std::mutex g_coprocMutex;
static std::map<int, plb_b200_lattice*> g_coproc;
static int g_nextHandle = 0;

static plb_b200_lattice* coproc(int handle) {
    std::lock_guard<std::mutex> lock(g_coprocMutex);
    auto it = g_coproc.find(handle);
    return it == g_coproc.end() ? nullptr : it->second;
}

int plb_b200_coproc_add_domain(int prec, int nx, int ny, int nz, double omega, int* handle) {
    if (!handle) return fail("null handle pointer");
    plb_b200_lattice* L = plb_b200_create(19, prec, nx, ny, nz, PLB_B200_DYN_BGK, omega, 0.);
    if (!L) return 0;
    // the co-processor contract has no statistics (the reference re-evaluates them on the host copy)
    L->impl->toggle_statistics(false);
    std::lock_guard<std::mutex> lock(g_coprocMutex);
    *handle = g_nextHandle++;
    g_coproc[*handle] = L;
    return 1;
}
int plb_b200_coproc_send(int handle, const int sub[6], const char* data, size_t bytes) {
    plb_b200_lattice* L = coproc(handle);
    if (!L) return fail("unknown domain handle");
    return plb_b200_upload(L, sub, data, bytes);
}
int plb_b200_coproc_receive(int handle, const int sub[6], char* data, size_t bytes) {
    plb_b200_lattice* L = coproc(handle);
    if (!L) return fail("unknown domain handle");
    return plb_b200_download(L, sub, data, bytes);
}
int plb_b200_coproc_collide_and_stream(int handle) {
    plb_b200_lattice* L = coproc(handle);
    if (!L) return fail("unknown domain handle");
    // "the collideAndStream() operation is blocking" (coProcessor3D.h:110-113)
    return plb_b200_collide_and_stream(L, 1) && plb_b200_synchronize(L);
}
int plb_b200_coproc_remove_domain(int handle) {
    plb_b200_lattice* L = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_coprocMutex);
        auto it = g_coproc.find(handle);
        if (it == g_coproc.end()) return fail("unknown domain handle");
        L = it->second;
        g_coproc.erase(it);
    }
    plb_b200_destroy(L);
    return 1;
}

}  // extern "C"
