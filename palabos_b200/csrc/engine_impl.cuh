// palabos_b200/csrc/engine_impl.cuh -- LatticeImpl<T,Q>: the host side of the engine.
//
// Mirrors MultiBlockLattice3D (multiBlock/multiBlockLattice3D.hh) for the hot path only: it owns
// the x-slabs of this process (the reference's per-block BlockLattice3D, allocated
// multiBlockLattice3D.hh:383-400), keeps the set-up state the reference keeps inside per-cell
// Dynamics objects (a 1-byte dynamics index per cell + sparse boundary records), orders the level-0
// data processors into concurrent stages, and sequences one time step as the reference does
// (multiBlockLattice3D.hh:318-360): sweep -> envelope exchange -> processors -> statistics.
// Citations are relative to /root/reference/src.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include <cuda.h>
#include "engine.h"
#include "kernels.cuh"
#include "sweep_launch.cuh"
#include "nccl_dyn.h"

namespace plb200 {

struct Proc {           ///< one data processor added by the boundary-condition set-up
    int cls;
    int box[6];
    int a, b, c;
};

template <class T, int Q>
class LatticeImpl : public LatticeBase {
    using L = LB<T, Q>;

    struct BcRec { int cell; Bc<T> b; bool alive; };
    struct Group { int box[6]; int first, count; };   ///< records appended by one add_*_boundary call

    struct Slab {
        int g = 0;                 ///< global slab index
        int gx0 = 0, nxl = 0;      ///< owned global planes [gx0, gx0+nxl)
        long long cellsStore = 0;  ///< (nxl+2)*P
        T* F = nullptr;
        uint8_t* dKind = nullptr;
        int* dBcIdx = nullptr;
        Bc<T>* dBcs = nullptr; size_t dBcsCap = 0;
        Item* dItems = nullptr; size_t dItemsCap = 0;
        double* dPartials = nullptr; int partialsCap = 0;
        T* dExt = nullptr;                 ///< external-moment sweep: rhoBar [cellsStore] | j_x | j_y | j_z, zero where nothing writes
        int* dBcCells = nullptr; size_t dBcCellsCap = 0;   ///< ... cells that carry a boundary record (-1: dead record)
        int nBcCells = 0;
        double* dStats = nullptr;          ///< [4] sum rhoBar, sum uSqr, max uSqr, cell count of the last sweep
        int statSlots = 0;                 ///< 1 once a sweep has reported
        unsigned int* dTicket = nullptr;
        int nAuto = 0;                     ///< items [0, nAuto): plane cells that complete at collision time (completed right before the sweep, never as a stage)
        std::vector<uint8_t> hKind;
        std::vector<BcRec> recs;
        std::vector<Group> groups;
        std::unordered_map<int, int> cellToRec;
        std::vector<int> stageOffsets;     ///< items of stage s are [stageOffsets[s], stageOffsets[s+1]) ...
        std::vector<int> stageSplit;       ///< ... plane items first, finite-difference items from stageSplit[s]
        size_t nItems = 0;
        unsigned int* dRimTicket = nullptr; ///< counts the rim tiles of a launch that signals a neighbouring rank
        int nReadPlanes0 = 0;              ///< the first nReadPlanes0 plane items of stage 0 are read by items of later stages
        int nHandoff0 = 0;                 ///< the last nHandoff0 plane items of stage 0 are handoff cells (Item::c = 1)
        T* dHbuf = nullptr; size_t hbufCap = 0;   ///< their completed populations between the plane pass and the next sweep, [Q][dBcsCap]
        bool bcAllInPasses = false;        ///< every owned cell with a boundary record is completed by a plane or finite-difference item
        bool valuesOnDevice = false;       ///< the records on the device hold newer u / rhoBar than `recs` (set_values_on_device)
        bool dirtyKind = true, dirtyBcIndex = true, dirtySched = true;
        bool kindUnflagged = true;         ///< dKind was uploaded from hKind and the flag bits it needs (KIND_PLANE, KIND_HANDOFF) are not set yet
        bool kindHasFlags = false;         ///< dKind carries flag bits (of an earlier set-up, if kindUnflagged)
        int bcDirtyLo = 0, bcDirtyHi = 0;  ///< record range whose values must be re-uploaded
    };

    CreateDesc desc_;
    int nx_, ny_, nz_, P_;
    int per_[3] = {0, 0, 0};
    int G_;                                  ///< total number of slabs over all ranks
    std::vector<Slab> slabs_;
    std::vector<Dyn<T>> dyns_;
    std::vector<T> dynCs_;                   ///< cSmago of each entry (identity of a Smagorinsky entry)
    Dyn<T>* dDyns_ = nullptr;
    bool dirtyDyns_ = true;
    std::vector<Proc> procs_;
    bool swapped_ = false;
    bool autoPlanes_ = false;                ///< some plane boundaries complete at collision time (dynamics-based models)
    bool statsOn_ = true;
    bool ext_ = false;                       ///< the lattice is stepped by ExternalRhoJcollideAndStream3D + BoxRhoBarJfunctional3D
    int bulk_[6] = {0, 0, 0, 0, 0, 0};       ///< ... whose level-3 domain is this box (the user's lattice; the rest is envelope)
    bool extImplied_ = false;                ///< the fields are what BoxRhoBarJfunctional3D makes of the CURRENT populations: the fused sweep may recompute them
    bool extComplete_ = true;                ///< the fields are stored for every cell (false: implied by the populations and the boundary records)
    bool extBcStored_ = true;                ///< ... at least for the cells with boundary dynamics (stored under the records they were computed with)
    bool storeBcMoments_ = false;            ///< the boundary passes running now also store the moments of their cells in the fields
    bool extFuse_ = true;                    ///< PLB_B200_EXT_EXPLICIT=1 in the environment: always store and re-read the fields (A/B tests)
    bool anySetup_ = false;
    bool stepped_ = false;
    bool failed_ = false;                    ///< a step died half-way (exchange or launch error): the populations are in no defined layout
    cudaStream_t stream_ = 0;
    uint64_t launches_ = 0;
    size_t devBytes_ = 0;
    int numSMs_ = 0;
    void* dStage_ = nullptr; size_t stageCap_ = 0;
    double* dReduce_ = nullptr;              ///< small device scratch for cross-rank reductions
    ncclComm_t comm_ = nullptr;
    bool profiling_ = false;                 ///< CUDA events around every sweep launch (bench.py's roofline leg)
    std::vector<cudaEvent_t> profEvents_;    ///< pairs (before, after)
    size_t profUsed_ = 0;

    static constexpr size_t STAGE_BYTES = size_t(256) << 20;
    static constexpr size_t XFER_CHUNK_BYTES = size_t(64) << 20;   ///< one of the two upload/download staging halves
    // envelope exchange (section 5 of DESIGN.md).  direct_: the rim cells of the sweep store the face-crossing populations straight
    // into the neighbouring slab -- a slab of this process, or (remote*_) the slab of the neighbouring rank mapped through
    // CUDA IPC, with one step counter per face direction written by the last rim tile and awaited by a stream memory
    // operation.  !direct_ (PLB_B200_HALO=nccl, or peer mapping impossible): ghost planes, cudaMemcpy / NCCL send/recv, unpack.
    bool direct_ = true;
    struct Remote { T* F = nullptr; long long S = 0; int nxl = 0; unsigned int* flag = nullptr; };
    Remote remoteLo_, remoteHi_;             ///< last slab of rank-1 / first slab of rank+1 (ring), and the flag of theirs this rank writes
    unsigned int* dFlags_ = nullptr;         ///< [0]: step count of the neighbour below, [32]: of the neighbour above (128 bytes apart)
    unsigned int stepCount_ = 0;
    std::map<std::string, void*> ipcOpened_; ///< CUDA IPC mappings by handle (an allocation can be opened once per process)
    cudaStream_t sideStream_ = nullptr;      ///< the finite-difference stages beside the plane pass (run_processors)
    cudaEvent_t forkEvent_ = nullptr, joinEvent_ = nullptr;
    bool handoffAllowed_ = true;             ///< PLB_B200_NO_HANDOFF=1 in the environment: the plane pass writes every cell back into the lattice (A/B tests)
    bool handoff_ = false;                   ///< the handoff buffers hold the completed populations of their cells; the lattice holds the raw ones
    bool overlapStages_ = true;              ///< PLB_B200_SERIAL_STAGES=1 in the environment: one stream (A/B tests)
    cudaStream_t copyStream_ = nullptr;
    cudaEvent_t copied_[2] = {nullptr, nullptr}, transposed_[2] = {nullptr, nullptr};

public:
    explicit LatticeImpl(const CreateDesc& d) : desc_(d), nx_(d.nx), ny_(d.ny), nz_(d.nz), P_(d.ny * d.nz) {
        int dev = 0, nDev = 0;
        if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
            throw PlbError("no CUDA device: libplb_b200 has no CPU fallback");
        PLB_CUDA(cudaGetDevice(&dev));
        PLB_CUDA(cudaDeviceGetAttribute(&numSMs_, cudaDevAttrMultiProcessorCount, dev));
        if (d.nRanks < 1 || d.slabsPerRank < 1 || d.rank < 0 || d.rank >= d.nRanks) throw PlbError("bad rank / slab counts");
        G_ = d.nRanks * d.slabsPerRank;
        if (G_ > nx_) throw PlbError("more slabs than x-planes");
        if ((long long)ny_ * nz_ >= (1ll << 30)) throw PlbError("y-z plane too large");
        if (const char* e = std::getenv("PLB_B200_HALO")) direct_ = !(std::strcmp(e, "nccl") == 0 || std::strcmp(e, "copy") == 0);
        if (const char* e = std::getenv("PLB_B200_EXT_EXPLICIT")) extFuse_ = !(e[0] == '1');
        if (const char* e = std::getenv("PLB_B200_SERIAL_STAGES")) overlapStages_ = !(e[0] == '1');
        if (const char* e = std::getenv("PLB_B200_NO_HANDOFF")) handoffAllowed_ = !(e[0] == '1');
        add_dyn(d.dynamics, d.p1, d.p2);     // entry 0: the background dynamics
        for (int s = 0; s < d.slabsPerRank; ++s) {
            Slab sl;
            sl.g = d.rank * d.slabsPerRank + s;
            int x0, x1;
            slab_range(nx_, G_, sl.g, x0, x1);
            sl.gx0 = x0; sl.nxl = x1 - x0 + 1;
            sl.cellsStore = (long long)(sl.nxl + 2) * P_;
            if (sl.cellsStore >= (1ll << 31)) throw PlbError("slab exceeds 2^31 cells: use more slabs");
            slabs_.push_back(std::move(sl));
        }
        for (Slab& sl : slabs_) {
            const size_t nF = (size_t)Q * sl.cellsStore;
            sl.F = (T*)dev_alloc(nF * sizeof(T));
            PLB_CUDA(cudaMemsetAsync(sl.F, 0, nF * sizeof(T), stream_));
            sl.dKind = (uint8_t*)dev_alloc(sl.cellsStore);
            sl.dBcIdx = (int*)dev_alloc(sl.cellsStore * sizeof(int));
            sl.dStats = (double*)dev_alloc(4 * sizeof(double));
            sl.dRimTicket = (unsigned int*)dev_alloc(sizeof(unsigned int));
            PLB_CUDA(cudaMemsetAsync(sl.dRimTicket, 0, sizeof(unsigned int), stream_));
            sl.dTicket = (unsigned int*)dev_alloc(sizeof(unsigned int));
            PLB_CUDA(cudaMemsetAsync(sl.dTicket, 0, sizeof(unsigned int), stream_));
            sl.hKind.assign((size_t)sl.cellsStore, 0);
        }
        dDyns_ = (Dyn<T>*)dev_alloc(256 * sizeof(Dyn<T>));
        dReduce_ = (double*)dev_alloc(8 * sizeof(double));
    }

    ~LatticeImpl() override {
        cudaStreamSynchronize(stream_);
        if (!ipcOpened_.empty()) {
            // nobody frees storage a neighbour still has mapped: unmap, meet, free
            for (auto& kv : ipcOpened_) cudaIpcCloseMemHandle(kv.second);
            try { double one = 1.; all_reduce(&one, 1, false); } catch (...) {}
        }
        cudaFree(dFlags_);
        for (Slab& sl : slabs_) {
            cudaFree(sl.F); cudaFree(sl.dExt); cudaFree(sl.dBcCells); cudaFree(sl.dKind); cudaFree(sl.dBcIdx); cudaFree(sl.dBcs); cudaFree(sl.dItems);
            cudaFree(sl.dPartials); cudaFree(sl.dStats); cudaFree(sl.dTicket); cudaFree(sl.dRimTicket); cudaFree(sl.dHbuf);
        }
        cudaFree(dDyns_); cudaFree(dStage_); cudaFree(dReduce_);
        if (comm_) NcclApi::get().CommDestroy(comm_);
        for (cudaEvent_t e : profEvents_) cudaEventDestroy(e);
        if (sideStream_) { cudaStreamDestroy(sideStream_); cudaEventDestroy(forkEvent_); cudaEventDestroy(joinEvent_); }
        if (copyStream_) {
            cudaStreamDestroy(copyStream_);
            for (int k = 0; k < 2; ++k) { cudaEventDestroy(copied_[k]); cudaEventDestroy(transposed_[k]); }
        }
    }

    // ------------------------------------------------------------------------------ plumbing

    void comm_init(const char id[128]) override {
        if (desc_.nRanks == 1) return;
        NcclApi& n = NcclApi::get();
        ncclUniqueId uid;
        static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
        std::memcpy(&uid, id, 128);
        n.check(n.CommInitRank(&comm_, desc_.nRanks, uid, desc_.rank), "ncclCommInitRank");
        map_neighbours();
    }
    const char* halo_transport() const override {
        if (G_ == 1) return "none (one slab)";
        if (!direct_) return desc_.nRanks > 1 ? "ghost planes: ncclSend/ncclRecv + unpack" : "ghost planes: cudaMemcpy + unpack";
        return desc_.nRanks > 1 ? "direct stores of the rim cells into the neighbour's slab (CUDA IPC over NVLink), one flag per face"
                                : "direct stores of the rim cells into the neighbouring slab";
    }
    void set_stream(cudaStream_t s) override { PLB_CUDA(cudaStreamSynchronize(stream_)); stream_ = s; }
    void synchronize() override { PLB_CUDA(cudaStreamSynchronize(stream_)); }
    uint64_t launch_count() const override { return launches_; }
    size_t device_bytes() const override { return devBytes_; }
    void toggle_statistics(bool on) override { statsOn_ = on; }

    void profile_sweeps(bool on) override { profiling_ = on; profUsed_ = 0; }
    void profile_read(double* totalMs, int* nLaunches) override {
        PLB_CUDA(cudaStreamSynchronize(stream_));
        double ms = 0.;
        for (size_t k = 0; k + 1 < profUsed_; k += 2) {
            float t = 0.f;
            PLB_CUDA(cudaEventElapsedTime(&t, profEvents_[k], profEvents_[k + 1]));
            ms += t;
        }
        if (totalMs) *totalMs = ms;
        if (nLaunches) *nLaunches = (int)(profUsed_ / 2);
    }

    void set_periodicity(int px, int py, int pz) override {
        if (G_ > 1 && (px != 0) != (per_[0] != 0) && anySetup_)
            throw PlbError("decomposed lattice: toggle x-periodicity before defining dynamics or boundaries");
        per_[0] = px != 0; per_[1] = py != 0; per_[2] = pz != 0;
        for (Slab& sl : slabs_) sl.dirtySched = true;
    }

    // -------------------------------------------------------------------------------- set-up

    void define_dynamics_box(const int box[6], int dynamics, double p1, double p2) override {
        check_box(box);
        ext_detach();
        const int idx = add_dyn(dynamics, p1, p2);
        anySetup_ = true;
        for (Slab& sl : slabs_) {
            bool touched = false;
            for (int X = 0; X <= sl.nxl + 1; ++X) {
                const int gx = global_x(sl, X);
                if (gx < box[0] || gx > box[1]) continue;
                touched = true;
                for (int y = box[2]; y <= box[3]; ++y)
                    std::memset(&sl.hKind[((size_t)X * ny_ + y) * nz_ + box[4]], idx, (size_t)(box[5] - box[4] + 1));
            }
            if (!touched) continue;
            sl.dirtyKind = sl.dirtySched = true;
            // defineDynamics replaces the whole dynamics object: a boundary composite on the cell goes away
            for (size_t r = 0; r < sl.recs.size(); ++r) {
                BcRec& rec = sl.recs[r];
                if (!rec.alive) continue;
                int gx, y, z;
                decode(sl, rec.cell, gx, y, z);
                if (in_box(box, gx, y, z)) { rec.alive = false; sl.cellToRec.erase(rec.cell); sl.dirtyBcIndex = true; }
            }
        }
    }

    void define_dynamics_mask(const int32_t* mask, int flag, int dynamics, double p1, double p2) override {
        if (!mask) throw PlbError("null mask");
        ext_detach();
        const int idx = add_dyn(dynamics, p1, p2);
        anySetup_ = true;
        for (Slab& sl : slabs_) {
            for (int X = 0; X <= sl.nxl + 1; ++X) {
                const int gx = global_x(sl, X);
                if (gx < 0) continue;
                const int32_t* m = mask + (size_t)gx * P_;
                uint8_t* k = &sl.hKind[(size_t)X * P_];
                for (int r = 0; r < P_; ++r) {
                    if (m[r] != flag) continue;
                    k[r] = (uint8_t)idx;
                    if (!sl.cellToRec.empty()) {
                        auto it = sl.cellToRec.find(X * P_ + r);
                        if (it != sl.cellToRec.end()) { sl.recs[it->second].alive = false; sl.cellToRec.erase(it); sl.dirtyBcIndex = true; }
                    }
                }
            }
            sl.dirtyKind = sl.dirtySched = true;
        }
    }

    /// BoundaryConditionInstantiator3D::addVelocityBoundary boundaryInstantiator3D.h:1016-1070 with the
    /// Wrapped{Regularized,ZouHe}BoundaryManager3D (boundaryCondition3D.hh:825-840, zouHeBoundary3D.hh):
    /// composite dynamics on the cells, then WrappedLocalBoundaryFunctional3D at level 0 (dirichlet only).
    void add_velocity_boundary(int model, int dir, int ori, const int box[6]) override {
        check_box(box);
        check_dir_ori(dir, ori);
        const int autoPrep = auto_model(model);
        add_bc(model == 1 ? BC_VEL_ZH : BC_VEL_REG, dir, ori, box);
        if (model == 4) add_proc(PROC_PLANE_FD, box, dir, ori, 0);
        else add_proc(PROC_PLANE, box, autoPrep, 0, 0);
    }
    void add_pressure_boundary(int model, int dir, int ori, const int box[6]) override {
        check_box(box);
        check_dir_ori(dir, ori);
        const int autoPrep = auto_model(model);
        add_bc(model == 1 ? BC_RHO_ZH : BC_RHO_REG, dir, ori, box);
        if (model == 4) add_proc(PROC_PLANE_FD, box, dir, ori, 0);
        else add_proc(PROC_PLANE, box, autoPrep, 0, 0);
    }
    /// addExternalVelocityEdge / Corner: StoreVelocityDynamics + Outer{Edge,Corner}Functional3D
    /// (boundaryCondition3D.hh:861-877, :899-915; the Zou-He manager does the same)
    void add_external_velocity_edge(int, int plane, int n1, int n2, const int box[6]) override {
        check_box(box);
        if (plane < 0 || plane > 2 || (n1 != 1 && n1 != -1) || (n2 != 1 && n2 != -1)) throw PlbError("bad edge plane/normals");
        add_bc(BC_STORE_VEL, 0, 0, box);
        add_proc(PROC_EDGE, box, plane, n1, n2);
    }
    void add_external_velocity_corner(int, int xn, int yn, int zn, int x, int y, int z) override {
        const int box[6] = {x, x, y, y, z, z};
        check_box(box);
        if ((xn != 1 && xn != -1) || (yn != 1 && yn != -1) || (zn != 1 && zn != -1)) throw PlbError("bad corner normals");
        add_bc(BC_STORE_VEL, 0, 0, box);
        add_proc(PROC_CORNER, box, xn, yn, zn);
    }

    /// addInternalVelocityEdge / Corner (the concave ones): RegularizedVelocityInnerEdge/CornerDynamics3D on the cells
    /// (regularizedBoundaryDynamics3D.hh:40-298); the wrapped managers (local, Zou-He) add WrappedLocalBoundaryFunctional3D -- a
    /// plane item here --, all others leave automaticPrepareCollision on (boundaryCondition3D.hh:770-820, 883-935, 998-1046, 1109-1157)
    int inner_auto(int model) {
        if (model < 0 || model > 4) throw PlbError("unknown boundary condition model");
        if (model >= 2 && ext_) throw PlbError("dynamics-based boundary conditions are not available with the external-moment sweep");
        if (model >= 2) autoPlanes_ = true;
        return model >= 2 ? 1 : 0;
    }
    void add_internal_velocity_edge(int model, int plane, int n1, int n2, const int box[6]) override {
        check_box(box);
        if (plane < 0 || plane > 2 || (n1 != 1 && n1 != -1) || (n2 != 1 && n2 != -1)) throw PlbError("bad edge plane/normals");
        const int autoPrep = inner_auto(model);
        add_bc(BC_INNER_EDGE, plane, (n1 > 0 ? 2 : 0) + (n2 > 0 ? 1 : 0), box);
        add_proc(PROC_PLANE, box, autoPrep, 0, 0);
    }
    void add_internal_velocity_corner(int model, int xn, int yn, int zn, int x, int y, int z) override {
        const int box[6] = {x, x, y, y, z, z};
        check_box(box);
        if ((xn != 1 && xn != -1) || (yn != 1 && yn != -1) || (zn != 1 && zn != -1)) throw PlbError("bad corner normals");
        const int autoPrep = inner_auto(model);
        add_bc(BC_INNER_CORNER, 0, (xn > 0 ? 4 : 0) + (yn > 0 ? 2 : 0) + (zn > 0 ? 1 : 0), box);
        add_proc(PROC_PLANE, box, autoPrep, 0, 0);
    }

    /// OnLatticeBoundaryCondition3D::setVelocityConditionOnBlockBoundaries boundaryCondition3D.hh:64-162
    /// with BlockSurface3D(block, 1) (core/blockSurface3D.cpp:38-200): 6 faces, 12 edges, 8 corners, in that order.
    void set_velocity_condition_on_block_boundaries(int model, const int d[6], const int app[6]) override {
        check_box(d);
        const int x0 = d[0], x1 = d[1], y0 = d[2], y1 = d[3], z0 = d[4], z1 = d[5];
        const int faces[6][8] = {
            {0, -1, x0, x0, y0 + 1, y1 - 1, z0 + 1, z1 - 1}, {0, +1, x1, x1, y0 + 1, y1 - 1, z0 + 1, z1 - 1},
            {1, -1, x0 + 1, x1 - 1, y0, y0, z0 + 1, z1 - 1}, {1, +1, x0 + 1, x1 - 1, y1, y1, z0 + 1, z1 - 1},
            {2, -1, x0 + 1, x1 - 1, y0 + 1, y1 - 1, z0, z0}, {2, +1, x0 + 1, x1 - 1, y0 + 1, y1 - 1, z1, z1}};
        const int edges[12][9] = {
            {0, -1, -1, x0 + 1, x1 - 1, y0, y0, z0, z0}, {0, -1, +1, x0 + 1, x1 - 1, y0, y0, z1, z1},
            {0, +1, -1, x0 + 1, x1 - 1, y1, y1, z0, z0}, {0, +1, +1, x0 + 1, x1 - 1, y1, y1, z1, z1},
            {1, -1, -1, x0, x0, y0 + 1, y1 - 1, z0, z0}, {1, -1, +1, x1, x1, y0 + 1, y1 - 1, z0, z0},
            {1, +1, -1, x0, x0, y0 + 1, y1 - 1, z1, z1}, {1, +1, +1, x1, x1, y0 + 1, y1 - 1, z1, z1},
            {2, -1, -1, x0, x0, y0, y0, z0 + 1, z1 - 1}, {2, -1, +1, x0, x0, y1, y1, z0 + 1, z1 - 1},
            {2, +1, -1, x1, x1, y0, y0, z0 + 1, z1 - 1}, {2, +1, +1, x1, x1, y1, y1, z0 + 1, z1 - 1}};
        int in[6];
        for (int k = 0; k < 6; ++k)
            if (intersect(&faces[k][2], app, in)) add_velocity_boundary(model, faces[k][0], faces[k][1], in);
        for (int k = 0; k < 12; ++k)
            if (intersect(&edges[k][3], app, in)) add_external_velocity_edge(model, edges[k][0], edges[k][1], edges[k][2], in);
        for (int k = 0; k < 8; ++k) {
            const int xn = (k & 4) ? 1 : -1, yn = (k & 2) ? 1 : -1, zn = (k & 1) ? 1 : -1;
            const int cb[6] = {xn < 0 ? x0 : x1, xn < 0 ? x0 : x1, yn < 0 ? y0 : y1, yn < 0 ? y0 : y1,
                               zn < 0 ? z0 : z1, zn < 0 ? z0 : z1};
            if (intersect(cb, app, in)) add_external_velocity_corner(model, xn, yn, zn, in[0], in[2], in[4]);
        }
    }

    /// setBoundaryVelocity dataInitializerFunctional3D.hh:830-845 -> Cell::defineVelocity: only the
    /// StoreVelocity-type dynamics keep it (boundaryDynamics.hh:333-339); no-op elsewhere.
    void set_boundary_velocity(const int box[6], const double u[3]) override {
        check_box(box);
        const T v[3] = {(T)u[0], (T)u[1], (T)u[2]};
        if (set_values_on_device(box, 0, v, (T)0)) return;
        for_records_in(box, [&](Bc<T>& b) {
            if (L::is_rho_plane(b.kind)) return false;
            if (std::memcmp(b.u, v, sizeof v) == 0) return false;      // unchanged: nothing to re-upload
            b.u[0] = v[0]; b.u[1] = v[1]; b.u[2] = v[2];
            return true;
        });
    }
    /// setBoundaryDensity -> StoreDensityDynamics::defineDensity boundaryDynamics.hh:239-243
    void set_boundary_density(const int box[6], double rho) override {
        check_box(box);
        const T rb = (T)rho - (T)1;
        const T none[3] = {(T)0, (T)0, (T)0};
        if (set_values_on_device(box, 1, none, rb)) return;
        for_records_in(box, [&](Bc<T>& b) {
            if (!L::is_rho_plane(b.kind)) return false;
            if (std::memcmp(&b.rhoBar, &rb, sizeof rb) == 0) return false;
            b.rhoBar = rb;
            return true;
        });
    }

    void initialize_at_equilibrium(const int box[6], double rho, const double u[3]) override {
        check_box(box);
        ext_detach();
        sync_setup(); materialize();
        for (Slab& sl : slabs_) {
            BoxL b;
            if (!local_box(sl, box, b)) continue;
            const long long n = b.cells();
            init_equilibrium_kernel<T, Q><<<(unsigned)((n + 255) / 256), 256, 0, stream_>>>(
                ctx(sl), sl.F, b, (T)rho, (T)u[0], (T)u[1], (T)u[2], nullptr);
            launched();
        }
    }

    void initialize_at_equilibrium_field(const int box[6], const double* ru) override {
        check_box(box);
        if (!ru) throw PlbError("null field");
        ext_detach();
        sync_setup(); materialize();
        const long long rowCells = (long long)(box[3] - box[2] + 1) * (box[5] - box[4] + 1);
        for (Slab& sl : slabs_) {
            BoxL b;
            if (!local_box(sl, box, b)) continue;
            const size_t planeBytes = (size_t)rowCells * 4 * sizeof(double);
            const int planesPerChunk = (int)std::max<size_t>(1, STAGE_BYTES / planeBytes);
            for (int X = b.X0; X <= b.X1; X += planesPerChunk) {
                BoxL cb = b;
                cb.X0 = X; cb.X1 = std::min(b.X1, X + planesPerChunk - 1);
                const size_t bytes = (size_t)cb.cells() * 4 * sizeof(double);
                double* st = (double*)stage(bytes);
                const long long firstCell = (long long)(global_x(sl, cb.X0) - box[0]) * rowCells;
                PLB_CUDA(cudaMemcpyAsync(st, ru + 4 * firstCell, bytes, cudaMemcpyHostToDevice, stream_));
                init_equilibrium_kernel<T, Q><<<(unsigned)((cb.cells() + 255) / 256), 256, 0, stream_>>>(
                    ctx(sl), sl.F, cb, (T)0, (T)0, (T)0, (T)0, st);
                launched();
                PLB_CUDA(cudaStreamSynchronize(stream_));   // the staging buffer is reused
            }
        }
    }

    /// lattice.initialize(): MultiBlock3D::initialize -> executeInternalProcessors (multiBlock/multiBlock3D.cpp:486-510)
    void initialize() override {
        // MultiBlock3D::initialize = executeInternalProcessors (multiBlock/multiBlock3D.cpp:330-332): with the sweep integrated
        // as a level-0 processor this is a whole step
        if (ext_) { advance(1, true); return; }
        sync_setup();
        run_processors();
    }

    // ----------------------------------------------------- the external-moment sweep

    /// generateMultiScalarField / generateMultiTensorField<3> + integrateProcessingFunctional(ExternalRhoJcollideAndStream3D, .., 0)
    /// + integrateProcessingFunctional(BoxRhoBarJfunctional3D, .., 3) of the fork's application (src/variables.hh:475-486).
    /// `bulk`: the box the two functionals were integrated on -- the user's lattice; what lies outside is the envelope
    /// layer which the sweep collides (with the zero moments found there) and streams, dynamicsProcessor3D.hh:118-119.
    void enable_external_rhoj(const int bulk[6]) override {
        check_box(bulk);
        if (!procs_.empty())
            throw PlbError("integrate the external-moment sweep before the boundary condition (its level-0 processors then run after the sweep, as this engine does)");
        for (Slab& sl : slabs_) {
            if (sl.dExt) continue;
            sl.dExt = (T*)dev_alloc(4 * (size_t)sl.cellsStore * sizeof(T));
            PLB_CUDA(cudaMemsetAsync(sl.dExt, 0, 4 * (size_t)sl.cellsStore * sizeof(T), stream_));
        }
        std::memcpy(bulk_, bulk, sizeof bulk_);
        ext_ = true;
        extImplied_ = false; extComplete_ = true; extBcStored_ = true;   // zero fields, as generateMultiScalarField / generateMultiTensorField leave them
        for (Slab& sl : slabs_) sl.dirtySched = true;
    }

    /// n x { lattice.executeInternalProcessors(); lattice.incrementTime(); } (viscosityTest/src/test.cpp:64-65)
    void execute_internal_processors(int n) override {
        if (!ext_) throw PlbError("no external-moment sweep is integrated in this lattice");
        advance(n, true);
    }

    /// applyProcessingFunctional(new BoxRhoBarJfunctional3D, box, rhoBarJarg) (src/variables.hh:650)
    void apply_rhobar_j(const int box[6]) override {
        if (!ext_) throw PlbError("no external-moment sweep is integrated in this lattice");
        check_box(box);
        int b[6];
        if (!intersect(box, bulk_, b)) return;
        sync_setup(); materialize();
        const bool whole = std::memcmp(b, bulk_, sizeof b) == 0;
        if (!whole) ext_complete();            // the rest of the fields keeps its values
        rhobar_j_box(b);
        if (whole) extImplied_ = true;         // (a part: implied stays what it was -- the new values ARE the implied ones)
        extComplete_ = true; extBcStored_ = true;
    }

    /// BoxRhoBarJ of the cells of a box, stored
    void rhobar_j_box(const int b[6]) {
        for (Slab& sl : slabs_) {
            BoxL lb;
            if (!local_box(sl, b, lb)) continue;
            rhobar_j_kernel<T, Q><<<(unsigned)((lb.cells() + 127) / 128), 128, 0, stream_>>>(ctx(sl), lb, sl.dExt, sl.cellsStore);
            launched();
        }
    }
    /// store the implied part of the fields (only ever needed when somebody wants to see or change them)
    void ext_complete() {
        if (!ext_ || extComplete_) return;
        sync_setup(); materialize();
        if (extComplete_) return;              // sync_setup had to store them itself (set-up changes were outstanding)
        rhobar_j_box(bulk_);
        extComplete_ = true; extBcStored_ = true;
    }
    /// the populations, the dynamics or the boundary set-up are about to change under the fields: from now on the fields are
    /// data of their own (the reference's fields keep their values until the next level-3 pass)
    void ext_detach() {
        if (!ext_) return;
        ext_complete();
        extImplied_ = false;
    }

    /// the rhoBar / j fields across the ABI: T [cells of the box owned by this rank][rhoBar, jx, jy, jz]
    void transfer_rhobar_j(const int box[6], void* data, size_t bytes, bool up) override {
        if (!ext_) throw PlbError("no external-moment sweep is integrated in this lattice");
        check_box(box);
        if (bytes != local_bytes(box) / Q * 4) throw PlbError("buffer size does not match the box");
        ext_complete();
        if (up) extImplied_ = false;           // the caller's values now: the next sweep reads them
        char* host = (char*)data;
        for (Slab& sl : slabs_) {
            BoxL b;
            if (!local_box(sl, box, b)) continue;
            const size_t planeBytes = (size_t)b.ny() * b.nz() * 4 * sizeof(T);
            const int planesPerChunk = (int)std::max<size_t>(1, STAGE_BYTES / planeBytes);
            for (int X = b.X0; X <= b.X1; X += planesPerChunk) {
                BoxL cb = b;
                cb.X0 = X; cb.X1 = std::min(b.X1, X + planesPerChunk - 1);
                const size_t nb = (size_t)cb.cells() * 4 * sizeof(T);
                T* st = (T*)stage(nb);
                const unsigned grid = (unsigned)((cb.cells() + 127) / 128);
                if (up) {
                    PLB_CUDA(cudaMemcpyAsync(st, host, nb, cudaMemcpyHostToDevice, stream_));
                    ext_transfer_kernel<T, true><<<grid, 128, 0, stream_>>>(geo(sl), cb, sl.dExt, sl.cellsStore, st);
                } else {
                    ext_transfer_kernel<T, false><<<grid, 128, 0, stream_>>>(geo(sl), cb, sl.dExt, sl.cellsStore, st);
                    PLB_CUDA(cudaMemcpyAsync(host, st, nb, cudaMemcpyDeviceToHost, stream_));
                }
                launched();
                PLB_CUDA(cudaStreamSynchronize(stream_));
                host += nb;
            }
        }
    }

    // ---------------------------------------------------------------------------- data in/out

    size_t local_bytes(const int box[6]) override {
        check_box(box);
        size_t n = 0;
        for (Slab& sl : slabs_) { BoxL b; if (local_box(sl, box, b)) n += (size_t)b.cells(); }
        return n * Q * sizeof(T);
    }

    void upload(const int box[6], const void* data, size_t bytes) override { transfer(box, const_cast<void*>(data), bytes, true); }
    void download(const int box[6], void* data, size_t bytes) override { transfer(box, data, bytes, false); }

    void density_velocity(const int box[6], void* out, size_t bytes) override {
        check_box(box);
        if (bytes != local_bytes(box) / Q * 4) throw PlbError("buffer size does not match the box");
        sync_setup(); materialize();
        char* host = (char*)out;
        for (Slab& sl : slabs_) {
            BoxL b;
            if (!local_box(sl, box, b)) continue;
            const size_t planeBytes = (size_t)b.ny() * b.nz() * 4 * sizeof(T);
            const int planesPerChunk = (int)std::max<size_t>(1, STAGE_BYTES / planeBytes);
            for (int X = b.X0; X <= b.X1; X += planesPerChunk) {
                BoxL cb = b;
                cb.X0 = X; cb.X1 = std::min(b.X1, X + planesPerChunk - 1);
                const size_t nb = (size_t)cb.cells() * 4 * sizeof(T);
                T* st = (T*)stage(nb);
                density_velocity_kernel<T, Q><<<(unsigned)((cb.cells() + 127) / 128), 128, 0, stream_>>>(ctx(sl), cb, st);
                launched();
                PLB_CUDA(cudaMemcpyAsync(host, st, nb, cudaMemcpyDeviceToHost, stream_));
                PLB_CUDA(cudaStreamSynchronize(stream_));
                host += nb;
            }
        }
    }

    // --------------------------------------------------------------------------- the hot path

    /// n x MultiBlockLattice3D::collideAndStream() (multiBlock/multiBlockLattice3D.hh:318-360)
    void collide_and_stream(int n) override {
        if (ext_) throw PlbError("a lattice that carries the external-moment sweep is stepped by execute_internal_processors");
        advance(n, false);
    }

    /// the step; ext: the sweep takes rhoBar and j from the fields and the fields are recomputed at the end (level 3)
    void advance(int n, bool ext) {
        if (n < 0) throw PlbError("negative step count");
        if (desc_.nRanks > 1 && !comm_) throw PlbError("plb_b200_comm_init must be called before stepping a multi-rank lattice");
        if (failed_) throw PlbError("an earlier step of this lattice failed half-way: its state is undefined, create a new lattice");
        sync_setup();
        try { advance_steps(n, ext); }
        catch (...) { failed_ = true; throw; }   // launches of a step were issued and the layout flags toggled only in part
    }
    void advance_steps(int n, bool ext) {
        for (int s = 0; s < n; ++s) {
            // One launch per slab: rim planes first (their stores go to the neighbouring slab), then the interior.  With the
            // direct transport nothing else is needed -- the neighbour reads what was stored into it one step later; the wait
            // below only orders this rank's NEXT accesses to its rim / ghost planes after the neighbours' stores of THIS step.
            ++stepCount_;
            // plane cells of the dynamics-based boundary managers complete at collision time: right before the sweep
            for (Slab& sl : slabs_) if (sl.nAuto > 0) launch_plane_stage(sl, 0, sl.nAuto, stream_);
            for (Slab& sl : slabs_) launch_sweep(sl);
            handoff_ = false;                  // the sweep has taken the handoff cells' populations over
            swapped_ = !swapped_;
            if (direct_) wait_neighbours();
            else exchange_copies(stream_);
            exchange_unpack();
            storeBcMoments_ = ext && extFuse_;     // (level 3 of this step, for the cells the boundary passes complete)
            run_processors();
            storeBcMoments_ = false;
            if (ext) {
                // level 3: BoxRhoBarJfunctional3D on the whole lattice.  Stored only for the cells with boundary dynamics; for the
                // others the next sweep recomputes it from the populations it gathers (launch_sweep), or ext_complete() stores it
                if (extFuse_) {
                    // all of it implied, nothing to run -- except that the boundary passes have left the moments of the cells
                    // they completed in the fields (plane_stage_kernel), which is all cells with boundary dynamics if every
                    // such cell belongs to a processor
                    extComplete_ = false; extBcStored_ = true;
                    for (const Slab& sl : slabs_) extBcStored_ = extBcStored_ && sl.bcAllInPasses;
                } else {
                    rhobar_j_box(bulk_);
                    extComplete_ = true; extBcStored_ = true;
                }
                extImplied_ = true;
            } else {
                stepped_ = true;
            }
        }
    }

    /// BlockStatistics::evaluate core/blockStatistics.cpp:77-120 + CombinedStatistics::combine
    /// multiBlock/combinedStatistics.cpp:108-133 (+ ParallelCombinedStatistics over ranks,
    /// parallelism/parallelStatistics.cpp:41-83).
    void stored_statistics(double out3[3], uint64_t* nCells) override {
        double sum[3] = {0., 0., 0.}, mx = -DBL_MAX;
        if (stepped_ && statsOn_) {
            PLB_CUDA(cudaStreamSynchronize(stream_));
            for (Slab& sl : slabs_) {
                double h[4];
                PLB_CUDA(cudaMemcpy(h, sl.dStats, sizeof h, cudaMemcpyDeviceToHost));
                if (sl.statSlots > 0) {
                    sum[0] += h[0]; sum[1] += h[1]; sum[2] += h[3];
                    mx = std::max(mx, h[2]);
                }
            }
        }
        if (desc_.nRanks > 1) {
            all_reduce(sum, 3, false);
            all_reduce(&mx, 1, true);
        }
        if (sum[2] == 0.) { out3[0] = 0.; out3[1] = 0.; }
        else { out3[0] = sum[0] / sum[2]; out3[1] = sum[1] / sum[2]; }
        out3[2] = mx;
        if (nCells) *nCells = (uint64_t)sum[2];
    }

    /// computeAverageEnergy dataAnalysisWrapper3D.hh:1252-1257 (+ BoxSumEnergyFunctional3D)
    double average_energy(const int box[6]) override {
        const double sum = box_sum(box, 0);
        const long long n = box_cells(box);
        const T sumEnergy = (T)(sum / (T)2);
        return (double)(sumEnergy / (T)n);
    }
    /// computeAverageDensity dataAnalysisWrapper3D.hh:1226-1230 (+ BoxSumRhoBarFunctional3D)
    double average_density(const int box[6]) override {
        const double sum = box_sum(box, 1);
        const long long n = box_cells(box);
        const T s = (T)sum;
        return (double)L::fullRho(s / (T)n);
    }

    void boundary_stages(int* nStages, int64_t* nItems) override {
        sync_setup();
        int st = 0; int64_t it = 0;
        for (Slab& sl : slabs_) { st = std::max(st, (int)sl.stageOffsets.size() - 1); it += (int64_t)sl.nItems; }
        if (nStages) *nStages = st;
        if (nItems) *nItems = it;
    }

private:
    // -------------------------------------------------------------------------------- helpers

    void* dev_alloc(size_t bytes) {
        void* p = nullptr;
        PLB_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        devBytes_ += bytes;
        return p;
    }
    void launched() {
        ++launches_;
        PLB_CUDA(cudaGetLastError());
    }
    void* stage(size_t bytes) {
        if (bytes > stageCap_) {
            PLB_CUDA(cudaStreamSynchronize(stream_));
            if (dStage_) { cudaFree(dStage_); devBytes_ -= stageCap_; }
            dStage_ = dev_alloc(bytes);
            stageCap_ = bytes;
        }
        return dStage_;
    }

    void check_box(const int b[6]) const {
        if (!(b[0] >= 0 && b[1] < nx_ && b[2] >= 0 && b[3] < ny_ && b[4] >= 0 && b[5] < nz_ &&
              b[0] <= b[1] && b[2] <= b[3] && b[4] <= b[5]))
            throw PlbError("box outside the lattice");
    }
    static void check_dir_ori(int dir, int ori) {
        if (dir < 0 || dir > 2 || (ori != 1 && ori != -1)) throw PlbError("bad direction/orientation");
    }
    static bool in_box(const int b[6], int x, int y, int z) {
        return x >= b[0] && x <= b[1] && y >= b[2] && y <= b[3] && z >= b[4] && z <= b[5];
    }
    static long long box_cells(const int b[6]) {
        return (long long)(b[1] - b[0] + 1) * (b[3] - b[2] + 1) * (b[5] - b[4] + 1);
    }
    static bool intersect(const int a[6], const int b[6], int out[6]) {
        for (int k = 0; k < 3; ++k) {
            out[2 * k] = std::max(a[2 * k], b[2 * k]);
            out[2 * k + 1] = std::min(a[2 * k + 1], b[2 * k + 1]);
            if (out[2 * k] > out[2 * k + 1]) return false;
        }
        return true;
    }

    /// global x of storage plane X of a slab (ghost planes wrap in a periodic lattice), or -1
    int global_x(const Slab& sl, int X) const {
        int gx = sl.gx0 - 1 + X;
        if (gx < 0 || gx >= nx_) {
            if (!per_[0] || G_ == 1) return -1;     // one slab wraps in the kernel: its ghost planes are unused
            gx = (gx + nx_) % nx_;
        }
        return gx;
    }
    void decode(const Slab& sl, int cell, int& gx, int& y, int& z) const {
        const int X = cell / P_;
        const int r = cell - X * P_;
        y = r / nz_; z = r - y * nz_;
        gx = global_x(sl, X);
    }
    /// owned part of a global box in storage coordinates
    bool local_box(const Slab& sl, const int box[6], BoxL& b) const {
        const int x0 = std::max(box[0], sl.gx0), x1 = std::min(box[1], sl.gx0 + sl.nxl - 1);
        if (x0 > x1) return false;
        b.X0 = x0 - sl.gx0 + 1; b.X1 = x1 - sl.gx0 + 1;
        b.y0 = box[2]; b.y1 = box[3]; b.z0 = box[4]; b.z1 = box[5];
        return true;
    }

    Geo geo(const Slab& sl) const {
        Geo g;
        g.nxl = sl.nxl; g.ny = ny_; g.nz = nz_; g.P = P_; g.S = sl.cellsStore;
        g.perY = per_[1]; g.perZ = per_[2];
        g.lowMode = sl.g > 0 ? XMODE_GHOST : (per_[0] ? (G_ == 1 ? XMODE_WRAP : XMODE_GHOST) : XMODE_EDGE);
        g.highMode = sl.g < G_ - 1 ? XMODE_GHOST : (per_[0] ? (G_ == 1 ? XMODE_WRAP : XMODE_GHOST) : XMODE_EDGE);
        return g;
    }
    CellCtx<T> ctx(const Slab& sl) const {
        CellCtx<T> c;
        c.F = sl.F; c.kind = sl.dKind; c.bcIdx = sl.dBcIdx; c.dyns = dDyns_; c.bcs = sl.dBcs;
        c.g = geo(sl); c.swapped = swapped_;
        c.ext = storeBcMoments_ ? sl.dExt : nullptr; c.extStride = sl.cellsStore;
        c.hbuf = sl.dHbuf; c.hStride = (long long)sl.hbufCap;
        return c;
    }

    int add_dyn(int type, double p1, double p2) {
        if (type < DYN_BGK || type > DYN_INCBGK) throw PlbError("unknown dynamics id");
        Dyn<T> d;
        std::memset(&d, 0, sizeof d);
        d.type = type;
        T cs = (T)0;
        if (type == DYN_BOUNCEBACK || type == DYN_NODYNAMICS) d.rho = (T)p1;
        else d.omega = (T)p1;
        if (type == DYN_SMAGO) { cs = (T)p2; d.preFactor = L::smago_prefactor(d.omega, cs); }
        for (size_t k = 0; k < dyns_.size(); ++k)
            if (std::memcmp(&dyns_[k], &d, sizeof d) == 0 && dynCs_[k] == cs) return (int)k;
        if (dyns_.size() > (size_t)KIND_MASK) throw PlbError("too many distinct dynamics (64)");
        dyns_.push_back(d);
        dynCs_.push_back(cs);
        dirtyDyns_ = true;
        for (Slab& sl : slabs_) sl.dirtySched = true;
        return (int)dyns_.size() - 1;
    }

    void add_bc(int kind, int dir, int ori, const int box[6]) {
        ext_detach();
        anySetup_ = true;
        for (Slab& sl : slabs_) {
            Group grp;
            std::memcpy(grp.box, box, sizeof grp.box);
            grp.first = (int)sl.recs.size();
            for (int X = 0; X <= sl.nxl + 1; ++X) {
                const int gx = global_x(sl, X);
                if (gx < box[0] || gx > box[1]) continue;
                for (int y = box[2]; y <= box[3]; ++y)
                    for (int z = box[4]; z <= box[5]; ++z) {
                        const int cell = (X * ny_ + y) * nz_ + z;
                        if (sl.cellToRec.count(cell)) throw PlbError("cell already carries a boundary dynamics");
                        BcRec rec;
                        rec.cell = cell; rec.alive = true;
                        std::memset(&rec.b, 0, sizeof rec.b);
                        rec.b.kind = kind; rec.b.dir = dir; rec.b.ori = ori;
                        rec.b.rhoBar = L::rhoBarOf((T)1);          // StoreDensityDynamics ctor boundaryDynamics.hh:200-205
                        sl.cellToRec.emplace(cell, (int)sl.recs.size());
                        sl.recs.push_back(rec);
                    }
            }
            grp.count = (int)sl.recs.size() - grp.first;
            if (grp.count) { sl.groups.push_back(grp); sl.dirtyBcIndex = sl.dirtySched = true; }
        }
    }
    /// createDynamicsBasedLocalBoundaryCondition3D / createEquilibriumBoundaryCondition3D (boundaryCondition3D.hh:1170-1181;
    /// in this fork both managers, :712-747 and :940-974, hand out the regularized dynamics WITHOUT a data processor): the
    /// plane dynamics keep automaticPrepareCollision = true and complete the populations at the beginning of collide()
    /// (CompositeDynamics::collide core/dynamics.hh:470-475) -- nothing runs after streaming, the cell holds raw
    /// populations between two steps.  Here: a plane "processor" entry with a = 1, never executed as a stage after the step:
    /// plane_stage_kernel completes such cells right before the sweep that collides them (advance_steps).
    int auto_model(int model) {
        if (model < 0 || model > 4) throw PlbError("unknown boundary condition model");
        // model 4 = createInterpBoundaryCondition3D (boundaryCondition3D.hh:1049-1085): plain Velocity / DensityDirichletBoundaryDynamics
        // (the records of the regularized model: same computeDensity / computeVelocity / computeRhoBarJ, empty completePopulations
        // is never asked for) + PlaneFdBoundaryFunctional3D, a finite-difference item like the edges and corners
        if (model == 4) return 0;
        if (model >= 2 && ext_) throw PlbError("dynamics-based boundary conditions are not available with the external-moment sweep");
        if (model >= 2) autoPlanes_ = true;
        return model >= 2 ? 1 : 0;
    }

    void add_proc(int cls, const int box[6], int a, int b, int c) {
        Proc p;
        p.cls = cls; std::memcpy(p.box, box, sizeof p.box); p.a = a; p.b = b; p.c = c;
        procs_.push_back(p);
        for (Slab& sl : slabs_) sl.dirtySched = true;
    }

    /// The time loop's form of setBoundaryVelocity / setBoundaryDensity: when the set-up on the device is current (no dynamics,
    /// boundary or schedule change outstanding) the values are changed THERE, by one small kernel over the box -- no walk over
    /// the host records, no upload (the plane completion of the step before has run with the previous values: nothing of it
    /// is outstanding).  The host records are refreshed from the device before they are needed again (pull_record_values).
    bool set_values_on_device(const int box[6], int what, const T u[3], T rhoBar) {
        if (!stepped_ && !ext_) return false;                 // set-up phase: the host path batches everything into one upload
        if (dirtyDyns_) return false;
        for (Slab& sl : slabs_)
            if (sl.dirtyKind || sl.dirtyBcIndex || sl.dirtySched || sl.bcDirtyHi > sl.bcDirtyLo || sl.kindUnflagged) return false;
        if (ext_ && !extComplete_ && !extBcStored_) {
            // the implied moments of the boundary cells are functions of the records about to change: store them first
            materialize();
            for (Slab& sl : slabs_) {
                if (sl.nBcCells == 0) continue;
                rhobar_j_list_kernel<T, Q><<<(sl.nBcCells + 127) / 128, 128, 0, stream_>>>(ctx(sl), sl.dBcCells, sl.nBcCells, sl.dExt, sl.cellsStore);
                launched();
            }
            extBcStored_ = true;
        }
        for (Slab& sl : slabs_) {
            if (sl.nBcCells == 0) continue;
            // owned planes, and the ghost planes (their records answer the neighbour reads of the edge / corner processors)
            BoxL parts[3];
            int nParts = 0;
            if (local_box(sl, box, parts[nParts])) ++nParts;
            for (int X : {0, sl.nxl + 1}) {
                const int gx = global_x(sl, X);
                if (gx < box[0] || gx > box[1]) continue;
                BoxL& b = parts[nParts++];
                b.X0 = b.X1 = X; b.y0 = box[2]; b.y1 = box[3]; b.z0 = box[4]; b.z1 = box[5];
            }
            for (int k = 0; k < nParts; ++k) {
                const long long n = parts[k].cells();
                set_bc_values_kernel<T, Q><<<(unsigned)((n + 127) / 128), 128, 0, stream_>>>(
                    sl.dBcIdx, sl.dBcs, geo(sl), parts[k], what, u[0], u[1], u[2], rhoBar);
                launched();
            }
            sl.valuesOnDevice = true;
        }
        return true;
    }

    /// the device holds newer record values than the host (set_values_on_device): fetch them before the host records are changed
    /// or uploaded again
    void pull_record_values() {
        for (Slab& sl : slabs_) {
            if (!sl.valuesOnDevice) continue;
            const size_t n = std::min<size_t>(sl.recs.size(), (size_t)sl.nBcCells);
            std::vector<Bc<T>> vals(n);
            if (n) {
                PLB_CUDA(cudaMemcpyAsync(vals.data(), sl.dBcs, n * sizeof(Bc<T>), cudaMemcpyDeviceToHost, stream_));
                PLB_CUDA(cudaStreamSynchronize(stream_));
            }
            for (size_t r = 0; r < n; ++r) {
                Bc<T>& b = sl.recs[r].b;
                b.u[0] = vals[r].u[0]; b.u[1] = vals[r].u[1]; b.u[2] = vals[r].u[2]; b.rhoBar = vals[r].rhoBar;
            }
            sl.valuesOnDevice = false;
        }
    }

    template <class Fn>
    void for_records_in(const int box[6], Fn fn) {
        pull_record_values();
        for (Slab& sl : slabs_)
            for (const Group& grp : sl.groups) {
                int tmp[6];
                if (!intersect(grp.box, box, tmp)) continue;
                for (int r = grp.first; r < grp.first + grp.count; ++r) {
                    BcRec& rec = sl.recs[r];
                    if (!rec.alive) continue;
                    int gx, y, z;
                    decode(sl, rec.cell, gx, y, z);
                    if (!in_box(box, gx, y, z)) continue;
                    if (fn(rec.b)) {
                        if (sl.bcDirtyLo == sl.bcDirtyHi) { sl.bcDirtyLo = r; sl.bcDirtyHi = r + 1; }
                        else { sl.bcDirtyLo = std::min(sl.bcDirtyLo, r); sl.bcDirtyHi = std::max(sl.bcDirtyHi, r + 1); }
                    }
                }
            }
    }

    // ------------------------------------------------------------ set-up state -> device

    struct View { int dynType; bool hasBc; int bcKind; };
    View view(const Slab& sl, int cell) const {
        View v;
        v.dynType = dyns_[sl.hKind[cell]].type;
        auto it = sl.cellToRec.find(cell);
        v.hasBc = it != sl.cellToRec.end();
        v.bcKind = v.hasBc ? sl.recs[it->second].b.kind : -1;
        return v;
    }

    /// Order the level-0 processors of a slab into stages.  The reference runs them sequentially, in
    /// insertion order, each over its box in x,y,z order (atomicBlock/atomicBlock3D.cpp:195-222); an item
    /// goes to the earliest stage that keeps every read-after-write, write-after-read and
    /// write-after-write relation of that sequential order, so items of one stage are independent.
    void build_schedule(Slab& sl) {
        std::vector<Item> items;
        std::vector<int> stageOf;
        std::unordered_map<int, int> lastWrite, lastRead;
        const int strideOf[3] = {P_, nz_, 1};
        auto wr = [&](int cell) { auto it = lastWrite.find(cell); return it == lastWrite.end() ? -1 : it->second; };
        auto rd = [&](int cell) { auto it = lastRead.find(cell); return it == lastRead.end() ? -1 : it->second; };
        for (const Proc& p : procs_) {
            BoxL b;
            if (!local_box(sl, p.box, b)) continue;
            for (int X = b.X0; X <= b.X1; ++X) for (int y = b.y0; y <= b.y1; ++y) for (int z = b.z0; z <= b.z1; ++z) {
                const int cell = (X * ny_ + y) * nz_ + z;
                int reads[8], nReads = 0;
                const int pos[3] = {X, y, z};
                auto neighbour = [&](int dir, int step) -> int {
                    // neighbours along y / z must exist in the lattice; along x they may be ghost cells
                    const int v = pos[dir] + step;
                    if (dir == 0) { if (global_x(sl, v) < 0) throw PlbError("boundary processor reaches outside the lattice"); }
                    else if (v < 0 || v >= (dir == 1 ? ny_ : nz_)) throw PlbError("boundary processor reaches outside the lattice");
                    return cell + step * strideOf[dir];
                };
                auto need_pops = [&](int c2, bool density, bool velocity) {
                    const View v = view(sl, c2);
                    const bool need = (density && L::density_reads_pops(v.dynType, v.hasBc, v.bcKind))
                                   || (velocity && L::velocity_reads_pops(v.dynType, v.hasBc, v.bcKind));
                    if (!need) return;
                    const int X2 = c2 / P_;
                    if (X2 < 1 || X2 > sl.nxl)
                        throw PlbError("a boundary edge/corner processor needs populations of a cell owned by another slab (unsupported decomposition)");
                    reads[nReads++] = c2;
                };
                if (p.cls == PROC_PLANE) {
                    if (!sl.cellToRec.count(cell)) continue;
                    if (p.a == 1) {            // completes at collision time: no processor, no dependencies
                        Item it;
                        it.cell = cell; it.bc = sl.cellToRec.at(cell); it.cls = (signed char)PROC_PLANE; it.a = 1; it.b = 0; it.c = 0;   // (c: handoff flag, never for these)
                        items.push_back(it);
                        stageOf.push_back(-1);
                        continue;
                    }
                } else if (p.cls == PROC_PLANE_FD) {
                    // (no isComposite() test in the reference: the processor also runs where defineDynamics has replaced
                    // the Dirichlet dynamics, finiteDifferenceBoundaryProcessor3D.hh:68-117)
                    // PlaneFdBoundaryFunctional3D: own density and velocity, velocity of the neighbour behind the face and of
                    // the four neighbours in the face
                    need_pops(cell, true, true);
                    need_pops(neighbour(p.a, -p.b), false, true);
                    for (int d = 0; d < 3; ++d) {
                        if (d == p.a) continue;
                        need_pops(neighbour(d, +1), false, true);
                        need_pops(neighbour(d, -1), false, true);
                    }
                } else if (p.cls == PROC_EDGE) {
                    const int d1 = (p.a + 1) % 3, d2 = (p.a + 2) % 3;
                    need_pops(neighbour(d1, -p.b), true, false);
                    need_pops(neighbour(d2, -p.c), true, false);
                    need_pops(neighbour(p.a, +1), false, true);
                    need_pops(neighbour(p.a, -1), false, true);
                    need_pops(neighbour(d1, -p.b), false, true);
                    need_pops(neighbour(d2, -p.c), false, true);
                    need_pops(cell, false, true);
                } else {
                    const int nrm[3] = {p.a, p.b, p.c};
                    for (int d = 0; d < 3; ++d) need_pops(neighbour(d, -nrm[d]), true, true);
                    need_pops(cell, false, true);
                }
                int st = std::max(wr(cell), rd(cell)) + 1;
                for (int k = 0; k < nReads; ++k) st = std::max(st, wr(reads[k]) + 1);
                for (int k = 0; k < nReads; ++k) { int& r = lastRead[reads[k]]; r = std::max(rd(reads[k]), st); }
                lastWrite[cell] = st;
                Item it;
                it.cell = cell; it.cls = (signed char)p.cls; it.a = (signed char)p.a; it.b = (signed char)p.b; it.c = (signed char)p.c;
                it.bc = -1;
                if (p.cls == PROC_PLANE) { it.bc = sl.cellToRec.at(cell); it.c = 0; }     // (c: handoff flag, set when the items are sorted)
                items.push_back(it);
                stageOf.push_back(st);
            }
        }
        int nStages = 0;
        for (int s : stageOf) nStages = std::max(nStages, s + 1);
        // bucket 0 = completes at collision time; then 1 + 2*stage + (0 plane | 1 edge/corner)
        const int nBuckets = 1 + 2 * nStages;
        std::vector<int> count0(nBuckets + 1, 0);
        auto bucket = [&](size_t k) { return stageOf[k] < 0 ? 0 : 1 + 2 * stageOf[k] + (items[k].cls == PROC_PLANE ? 0 : 1); };
        for (size_t k = 0; k < items.size(); ++k) ++count0[bucket(k) + 1];
        for (int b = 0; b < nBuckets; ++b) count0[b + 1] += count0[b];
        std::vector<Item> sorted(items.size());
        std::vector<int> cursor(count0.begin(), count0.begin() + nBuckets);
        // (the plane items of stage 0 whose cells a later item reads go first: run_processors can then let the rest of the plane
        // pass run beside the finite-difference stages)
        // ... and the handoff cells last (KIND_HANDOFF, cell_io.cuh): plane cells of stage 0 that nothing else touches and that sit
        // among at most 8 of their kind in the 32 cells a warp of the sweep covers (a z-face: one per row)
        sl.nReadPlanes0 = 0; sl.nHandoff0 = 0;
        auto plane0 = [&](size_t k) { return stageOf[k] == 0 && items[k].cls == PROC_PLANE; };
        auto untouched = [&](size_t k) { return rd(items[k].cell) <= 0 && wr(items[k].cell) <= 0; };
        std::unordered_map<int, int> perPiece;
        if (handoffAllowed_ && !(ext_ && !extFuse_))
            for (size_t k = 0; k < items.size(); ++k) if (plane0(k) && untouched(k)) ++perPiece[(items[k].cell - P_) / 32];
        for (int pass = 0; pass < 3; ++pass)
            for (size_t k = 0; k < items.size(); ++k) {
                int cls = 0;                                   // 0: everything else and the plane items a later stage reads; 1: plain; 2: handoff
                if (plane0(k)) {
                    const bool readLater = rd(items[k].cell) > 0;
                    auto pc = perPiece.find((items[k].cell - P_) / 32);
                    const bool hand = !readLater && untouched(k) && pc != perPiece.end() && pc->second <= 8;
                    cls = readLater ? 0 : (hand ? 2 : 1);
                    if (pass == 0 && readLater) ++sl.nReadPlanes0;
                    if (pass == 2 && hand) ++sl.nHandoff0;
                    if (items[k].cls == PROC_PLANE) items[k].c = (signed char)(hand ? 1 : 0);
                }
                if (cls != pass) continue;
                sorted[cursor[bucket(k)]++] = items[k];
            }
        sl.nAuto = count0[1];
        // count[2*stage + (0|1)] as before
        const int* count = count0.data() + 1;
        {
            std::unordered_map<int, int> seen;
            for (const Item& it : sorted) seen[it.cell] = 1;
            sl.bcAllInPasses = sl.nAuto == 0;
            for (const BcRec& rec : sl.recs) {
                if (!rec.alive || rec.cell < P_ || rec.cell >= (sl.nxl + 1) * P_) continue;
                if (!seen.count(rec.cell)) { sl.bcAllInPasses = false; break; }
            }
        }
        sl.stageOffsets.assign(nStages + 1, 0);
        sl.stageSplit.assign(nStages, 0);
        for (int st = 0; st < nStages; ++st) { sl.stageOffsets[st] = count[2 * st]; sl.stageSplit[st] = count[2 * st + 1]; }
        sl.stageOffsets[nStages] = (int)items.size();
        sl.nItems = sorted.size();
        if (sl.nItems > sl.dItemsCap) {
            if (sl.dItems) { PLB_CUDA(cudaStreamSynchronize(stream_)); cudaFree(sl.dItems); devBytes_ -= sl.dItemsCap * sizeof(Item); }
            sl.dItems = (Item*)dev_alloc(sl.nItems * sizeof(Item));
            sl.dItemsCap = sl.nItems;
        }
        if (sl.nItems) {
            PLB_CUDA(cudaMemcpyAsync(sl.dItems, sorted.data(), sl.nItems * sizeof(Item), cudaMemcpyHostToDevice, stream_));
            PLB_CUDA(cudaStreamSynchronize(stream_));
        }
    }

    /// completePopulations of the plane items [first, first + n) of a slab, in the current layout
    void launch_plane_stage(Slab& sl, int first, int n, cudaStream_t st) {
        if (n <= 0) return;
        if (swapped_) plane_stage_kernel<T, Q, true><<<(n + 127) / 128, 128, 0, st>>>(ctx(sl), sl.F, sl.dItems + first, n);
        else plane_stage_kernel<T, Q, false><<<(n + 127) / 128, 128, 0, st>>>(ctx(sl), sl.F, sl.dItems + first, n);
        launched();
    }

    /// The handoff buffers back into the lattice (handoff_scatter_kernel): before anything but the sweep reads or writes the
    /// populations of those cells.
    void materialize() {
        if (!handoff_) return;
        for (Slab& sl : slabs_) {
            if (sl.nHandoff0 <= 0 || sl.stageOffsets.size() < 2) continue;
            const int first = sl.stageSplit[0] - sl.nHandoff0, n = sl.nHandoff0;
            if (swapped_) handoff_scatter_kernel<T, Q, true><<<(n + 127) / 128, 128, 0, stream_>>>(ctx(sl), sl.F, sl.dItems + first, n);
            else handoff_scatter_kernel<T, Q, false><<<(n + 127) / 128, 128, 0, stream_>>>(ctx(sl), sl.F, sl.dItems + first, n);
            launched();
        }
        handoff_ = false;
    }

    void sync_setup() {
        bool anyDirty = dirtyDyns_;
        for (Slab& sl : slabs_) anyDirty = anyDirty || sl.dirtyKind || sl.dirtyBcIndex || sl.dirtySched || sl.bcDirtyHi > sl.bcDirtyLo;
        if (!anyDirty) return;
        materialize();                         // with the items and records the buffers were filled under
        pull_record_values();
        bool reflag = false;
        bool structural = false;               // (external-moment sweep) which cells read their moments from the fields may have changed
        for (Slab& sl : slabs_) structural = structural || sl.dirtyKind || sl.dirtyBcIndex || sl.dirtySched;
        if (ext_ && !extComplete_) {
            // the implied moments are functions of the device state about to be replaced: store them first -- everything when
            // dynamics / boundary cells change (the fields are then data of their own), the boundary cells when only
            // values of their records do
            if (structural) {
                rhobar_j_box(bulk_);
                extComplete_ = true; extBcStored_ = true; extImplied_ = false;
            } else if (!extBcStored_) {
                for (Slab& sl : slabs_) {
                    if (sl.nBcCells == 0) continue;
                    rhobar_j_list_kernel<T, Q><<<(sl.nBcCells + 127) / 128, 128, 0, stream_>>>(ctx(sl), sl.dBcCells, sl.nBcCells, sl.dExt, sl.cellsStore);
                    launched();
                }
                extBcStored_ = true;
            }
        }
        if (ext_ && structural) reflag = true;
        if (dirtyDyns_) {
            PLB_CUDA(cudaMemcpyAsync(dDyns_, dyns_.data(), dyns_.size() * sizeof(Dyn<T>), cudaMemcpyHostToDevice, stream_));
            PLB_CUDA(cudaStreamSynchronize(stream_));
            dirtyDyns_ = false;
        }
        for (Slab& sl : slabs_) {
            if (sl.dirtyKind) {
                PLB_CUDA(cudaMemcpyAsync(sl.dKind, sl.hKind.data(), (size_t)sl.cellsStore, cudaMemcpyHostToDevice, stream_));
                PLB_CUDA(cudaStreamSynchronize(stream_));
                sl.dirtyKind = false;
                sl.kindUnflagged = true; sl.kindHasFlags = false;
            }
            if (sl.dirtyBcIndex) {
                // records keep their position (dead ones are simply not indexed)
                const size_t n = sl.recs.size();
                if (n > sl.dBcsCap) {
                    if (sl.dBcs) { PLB_CUDA(cudaStreamSynchronize(stream_)); cudaFree(sl.dBcs); devBytes_ -= sl.dBcsCap * sizeof(Bc<T>); }
                    sl.dBcsCap = n + n / 2 + 16;
                    sl.dBcs = (Bc<T>*)dev_alloc(sl.dBcsCap * sizeof(Bc<T>));
                }
                PLB_CUDA(cudaMemsetAsync(sl.dBcIdx, 0xFF, (size_t)sl.cellsStore * sizeof(int), stream_));
                if (n) {
                    std::vector<int> cells(n);
                    std::vector<Bc<T>> vals(n);
                    for (size_t r = 0; r < n; ++r) { cells[r] = sl.recs[r].alive ? sl.recs[r].cell : -1; vals[r] = sl.recs[r].b; }
                    int* dCells = (int*)stage(n * sizeof(int));
                    PLB_CUDA(cudaMemcpyAsync(dCells, cells.data(), n * sizeof(int), cudaMemcpyHostToDevice, stream_));
                    PLB_CUDA(cudaMemcpyAsync(sl.dBcs, vals.data(), n * sizeof(Bc<T>), cudaMemcpyHostToDevice, stream_));
                    scatter_bc_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream_>>>(dCells, (int)n, sl.dBcIdx);
                    launched();
                    if (ext_) {
                        if (n > sl.dBcCellsCap) {
                            if (sl.dBcCells) { PLB_CUDA(cudaStreamSynchronize(stream_)); cudaFree(sl.dBcCells); devBytes_ -= sl.dBcCellsCap * sizeof(int); }
                            sl.dBcCellsCap = n + n / 2 + 16;
                            sl.dBcCells = (int*)dev_alloc(sl.dBcCellsCap * sizeof(int));
                        }
                        // owned planes only: the records on the ghost planes exist for the neighbour reads of the edge / corner processors
                        std::vector<int> owned(cells);
                        for (size_t r = 0; r < n; ++r)
                            if (owned[r] >= 0 && (owned[r] < P_ || owned[r] >= (sl.nxl + 1) * P_)) owned[r] = -1;
                        PLB_CUDA(cudaMemcpyAsync(sl.dBcCells, owned.data(), n * sizeof(int), cudaMemcpyHostToDevice, stream_));
                        PLB_CUDA(cudaStreamSynchronize(stream_));     // `owned` goes out of scope
                    }
                    PLB_CUDA(cudaStreamSynchronize(stream_));
                }
                sl.nBcCells = (int)n;
                sl.dirtyBcIndex = false;
                sl.bcDirtyLo = sl.bcDirtyHi = 0;
            }
            if (sl.bcDirtyHi > sl.bcDirtyLo) {
                const int n = sl.bcDirtyHi - sl.bcDirtyLo;
                std::vector<Bc<T>> vals(n);
                for (int r = 0; r < n; ++r) vals[r] = sl.recs[sl.bcDirtyLo + r].b;
                PLB_CUDA(cudaMemcpyAsync(sl.dBcs + sl.bcDirtyLo, vals.data(), n * sizeof(Bc<T>), cudaMemcpyHostToDevice, stream_));
                PLB_CUDA(cudaStreamSynchronize(stream_));
                sl.bcDirtyLo = sl.bcDirtyHi = 0;
            }
            if (sl.dirtySched) { build_schedule(sl); sl.dirtySched = false; reflag = true; }
            if (sl.kindUnflagged) reflag = true;
        }
        if (structural || reflag) {
            for (Slab& sl : slabs_) {
                if (sl.kindHasFlags) {     // drop the flags of an earlier set-up
                    PLB_CUDA(cudaMemcpyAsync(sl.dKind, sl.hKind.data(), (size_t)sl.cellsStore, cudaMemcpyHostToDevice, stream_));
                    PLB_CUDA(cudaStreamSynchronize(stream_));
                }
                if (ext_) {
                    // KIND_PLANE (external-moment sweep): "boundary dynamics on top, or envelope layer" -- the cells whose moments
                    // are not those of a bulk cell
                    BoxL lb;
                    if (!local_box(sl, bulk_, lb)) { lb.X0 = 1; lb.X1 = 0; lb.y0 = lb.z0 = 0; lb.y1 = lb.z1 = -1; }
                    flag_ext_cells_kernel<<<(unsigned)((sl.cellsStore + 255) / 256), 256, 0, stream_>>>(sl.dKind, sl.dBcIdx, geo(sl), lb, (int)sl.cellsStore);
                    launched();
                }
                if (sl.nHandoff0 > 0) {
                    // KIND_HANDOFF, and the buffer that carries those cells' populations from the plane pass to the next sweep
                    if (sl.hbufCap != sl.dBcsCap) {
                        if (sl.dHbuf) { PLB_CUDA(cudaStreamSynchronize(stream_)); cudaFree(sl.dHbuf); devBytes_ -= sl.hbufCap * Q * sizeof(T); }
                        sl.hbufCap = sl.dBcsCap;
                        sl.dHbuf = (T*)dev_alloc(sl.hbufCap * Q * sizeof(T));
                    }
                    const int first = sl.stageSplit[0] - sl.nHandoff0;
                    flag_handoff_cells_kernel<<<(sl.nHandoff0 + 255) / 256, 256, 0, stream_>>>(sl.dItems + first, sl.nHandoff0, sl.dKind);
                    launched();
                }
                sl.kindHasFlags = sl.nHandoff0 > 0 || ext_;
            }
        }
        for (Slab& sl : slabs_) sl.kindUnflagged = false;
    }

    // ------------------------------------------------------------------------------ the step

    template <int MAIN, bool GEN, bool ODD>
    void launch_sweep_as(Slab& sl, SweepArgs<T>& a, int mainIdx) {
        constexpr int threads = sweep_threads(ODD);
        (void)mainIdx;
        // tile sequence: the planes next to a slab that takes direct stores first, then the rest
        const int lowEnd = a.lo.F ? 2 * P_ : P_, end = (sl.nxl + 1) * P_;
        const int highBegin = a.hi.F ? std::max(sl.nxl * P_, lowEnd) : end;
        a.nRanges = 0; a.nTiles = 0;
        auto add = [&](int b, int e) {
            if (e <= b) return;
            a.range[a.nRanges].begin = b; a.range[a.nRanges].end = e; a.range[a.nRanges].tile0 = a.nTiles;
            a.nTiles += (e - b + threads - 1) / threads;
            ++a.nRanges;
        };
        for (int k = 0; k < 3; ++k) { a.range[k].begin = a.range[k].end = a.range[k].tile0 = 0; }
        add(P_, lowEnd);
        add(highBegin, end);
        a.nRimTiles = (a.sigLo || a.sigHi) ? a.nTiles : 0;
        add(lowEnd, highBegin);
        // blocks per SM in the grid: the measured optimum (kernels.cuh), or PLB_B200_GRID_EVEN / PLB_B200_GRID_ODD (tuning runs)
        static const int gfEnv[2] = {std::getenv("PLB_B200_GRID_EVEN") ? std::atoi(std::getenv("PLB_B200_GRID_EVEN")) : 0,
                                     std::getenv("PLB_B200_GRID_ODD") ? std::atoi(std::getenv("PLB_B200_GRID_ODD")) : 0};
        const int gf = gfEnv[ODD ? 1 : 0] > 0 ? gfEnv[ODD ? 1 : 0] : sweep_grid_factor<T, Q>(ODD);
        const int grid = std::max(1, std::min(a.nTiles, numSMs_ * gf));
        if (grid > sl.partialsCap) {
            if (sl.dPartials) { PLB_CUDA(cudaStreamSynchronize(stream_)); cudaFree(sl.dPartials); devBytes_ -= (size_t)sl.partialsCap * 4 * sizeof(double); }
            sl.partialsCap = std::max(grid, numSMs_ * 16);
            sl.dPartials = (double*)dev_alloc((size_t)sl.partialsCap * 4 * sizeof(double));
        }
        a.partials = sl.dPartials;
        const bool prof = profiling_ && profUsed_ < 2 * 8192;
        if (prof) {
            while (profEvents_.size() < profUsed_ + 2) { cudaEvent_t e; PLB_CUDA(cudaEventCreate(&e)); profEvents_.push_back(e); }
            PLB_CUDA(cudaEventRecord(profEvents_[profUsed_], stream_));
        }
        launch_sweep_kernel<T, Q, MAIN, GEN, ODD>(grid, stream_, a);
        launched();
        if (prof) { PLB_CUDA(cudaEventRecord(profEvents_[profUsed_ + 1], stream_)); profUsed_ += 2; }
        sl.statSlots = 1;
    }

    /// the slab on the other side of a face of `sl`, if the rim cells store into it directly
    void peers_of(const Slab& sl, const Geo& g, SweepArgs<T>& a) {
        a.lo.F = a.hi.F = nullptr; a.lo.S = a.hi.S = 0; a.lo.nxl = a.hi.nxl = 0;
        a.sigLo = a.sigHi = nullptr;
        if (!direct_ || G_ == 1) return;
        auto local = [&](int gIdx, PeerSlab<T>& p) {
            Slab* o = find_slab(gIdx);
            if (!o) return false;
            p.F = o->F; p.S = o->cellsStore; p.nxl = o->nxl;
            return true;
        };
        if (g.lowMode == XMODE_GHOST && !local((sl.g - 1 + G_) % G_, a.lo)) {
            a.lo.F = remoteLo_.F; a.lo.S = remoteLo_.S; a.lo.nxl = remoteLo_.nxl; a.sigLo = remoteLo_.flag;
        }
        if (g.highMode == XMODE_GHOST && !local((sl.g + 1) % G_, a.hi)) {
            a.hi.F = remoteHi_.F; a.hi.S = remoteHi_.S; a.hi.nxl = remoteHi_.nxl; a.sigHi = remoteHi_.flag;
        }
        if ((g.lowMode == XMODE_GHOST && !a.lo.F) || (g.highMode == XMODE_GHOST && !a.hi.F))
            throw PlbError("the neighbouring rank's slab is not mapped (plb_b200_comm_init)");
    }

    /// sweep a slab
    void launch_sweep(Slab& sl) {
        SweepArgs<T> a;
        a.F = sl.F; a.kind = sl.dKind; a.dyns = dDyns_; a.g = geo(sl);
        a.bcIdx = sl.dBcIdx; a.bcs = sl.dBcs;
        peers_of(sl, a.g, a);
        a.rimTicket = sl.dRimTicket; a.sigValue = stepCount_; a.nRimTiles = 0;
        for (int i = 0; i < 27; ++i) a.ldBase[i] = a.stBase[i] = nullptr;
        for (int i = 0; i < Q; ++i) {
            const int o = L::opp(i);
            if (!swapped_) {            // even half: the cell itself
                a.ldBase[i] = sl.F + (long long)i * sl.cellsStore;
                a.stBase[i] = sl.F + (long long)o * sl.cellsStore;
            } else {                    // odd half, every neighbour at its plain offset
                const long long off = (long long)c_of(i, 0) * P_ + (long long)c_of(i, 1) * nz_ + c_of(i, 2);
                a.ldBase[i] = sl.F + (long long)o * sl.cellsStore - off;
                a.stBase[i] = sl.F + (long long)i * sl.cellsStore + off;
            }
        }
        a.omega = dyns_[0].omega; a.preFactor = dyns_[0].preFactor;
        a.gatherStats = statsOn_ ? 1 : 0;
        a.partials = nullptr; a.ticket = sl.dTicket; a.statsOut = sl.dStats;
        a.ext = sl.dExt; a.extStride = sl.cellsStore; a.extBcFromFields = 1;
        a.hbuf = (handoff_ && sl.nHandoff0 > 0) ? sl.dHbuf : nullptr; a.hStride = (long long)sl.hbufCap;
        if (ext_) {
            // the statistics of this mode are never evaluated by the reference's loop (executeInternalProcessors + incrementTime)
            a.gatherStats = 0;
            a.extBcFromFields = extBcStored_ ? 1 : 0;
            if (extImplied_ && extFuse_) {
                if (swapped_) launch_sweep_as<DYN_EXTERNAL_FUSED, true, true>(sl, a, 7);
                else launch_sweep_as<DYN_EXTERNAL_FUSED, true, false>(sl, a, 7);
            } else {
                if (swapped_) launch_sweep_as<DYN_EXTERNAL, true, true>(sl, a, 6);
                else launch_sweep_as<DYN_EXTERNAL, true, false>(sl, a, 6);
            }
            return;
        }
        bool generic = L::is_inert(dyns_[0].type);
        for (size_t k = 1; k < dyns_.size(); ++k) if (!L::is_inert(dyns_[k].type)) generic = true;
        const bool odd = swapped_;
        if (generic) {
            if (odd) launch_sweep_as<DYN_BGK, true, true>(sl, a, 4); else launch_sweep_as<DYN_BGK, true, false>(sl, a, 4);
            return;
        }
        switch (dyns_[0].type) {
        case DYN_BGK:   if (odd) launch_sweep_as<DYN_BGK, false, true>(sl, a, 0);   else launch_sweep_as<DYN_BGK, false, false>(sl, a, 0); break;
        case DYN_RLB:   if (odd) launch_sweep_as<DYN_RLB, false, true>(sl, a, 1);   else launch_sweep_as<DYN_RLB, false, false>(sl, a, 1); break;
        case DYN_TRT:   if (odd) launch_sweep_as<DYN_TRT, false, true>(sl, a, 2);   else launch_sweep_as<DYN_TRT, false, false>(sl, a, 2); break;
        case DYN_INCBGK: if (odd) launch_sweep_as<DYN_INCBGK, false, true>(sl, a, 5); else launch_sweep_as<DYN_INCBGK, false, false>(sl, a, 5); break;
        default:        if (odd) launch_sweep_as<DYN_SMAGO, false, true>(sl, a, 3); else launch_sweep_as<DYN_SMAGO, false, false>(sl, a, 3); break;
        }
    }

    void run_processors() {
        materialize();                         // processors run twice in a row (initialize() called again)
        for (Slab& sl : slabs_) if (sl.nHandoff0 > 0) handoff_ = true;      // (the plane pass below fills the buffers)
        for (Slab& sl : slabs_) {
            const int nStages = (int)sl.stageOffsets.size() - 1;
            // The usual shape -- all plane items in stage 0, edges and corners behind them -- runs as a fork: the few plane cells
            // that the later stages read first, then the finite-difference stages (short dependent launches, ~45 us each) on a
            // side stream BESIDE the rest of the plane pass.  Safe because every storage location belongs to one (population,
            // cell) pair and an item writes the locations of its own cell only.
            bool fork = overlapStages_ && nStages >= 2 && sl.stageSplit[0] - sl.stageOffsets[0] - sl.nReadPlanes0 >= 4096;
            for (int s = 1; s < nStages && fork; ++s) if (sl.stageSplit[s] > sl.stageOffsets[s]) fork = false;    // a plane item in a later stage
            if (fork && sl.stageOffsets[1] > sl.stageSplit[0]) fork = false;                                       // finite-difference items in stage 0
            if (fork) {
                if (!sideStream_) {
                    PLB_CUDA(cudaStreamCreateWithFlags(&sideStream_, cudaStreamNonBlocking));
                    PLB_CUDA(cudaEventCreateWithFlags(&forkEvent_, cudaEventDisableTiming));
                    PLB_CUDA(cudaEventCreateWithFlags(&joinEvent_, cudaEventDisableTiming));
                }
                const int first = sl.stageOffsets[0], split = sl.stageSplit[0];
                launch_plane_stage(sl, first, sl.nReadPlanes0, stream_);
                PLB_CUDA(cudaEventRecord(forkEvent_, stream_));
                PLB_CUDA(cudaStreamWaitEvent(sideStream_, forkEvent_, 0));
                launch_plane_stage(sl, first + sl.nReadPlanes0, split - first - sl.nReadPlanes0, stream_);
                for (int s = 1; s < nStages; ++s) {
                    const int n = sl.stageOffsets[s + 1] - sl.stageSplit[s];
                    if (n <= 0) continue;
                    fd_stage_kernel<T, Q><<<(n + 63) / 64, 64, 0, sideStream_>>>(ctx(sl), sl.F, sl.dItems + sl.stageSplit[s], n);
                    launched();
                }
                PLB_CUDA(cudaEventRecord(joinEvent_, sideStream_));
                PLB_CUDA(cudaStreamWaitEvent(stream_, joinEvent_, 0));
                continue;
            }
            for (int s = 0; s < nStages; ++s) {
                const int first = sl.stageOffsets[s], split = sl.stageSplit[s], last = sl.stageOffsets[s + 1];
                launch_plane_stage(sl, first, split - first, stream_);
                if (last > split) {
                    const int n = last - split;
                    fd_stage_kernel<T, Q><<<(n + 63) / 64, 64, 0, stream_>>>(ctx(sl), sl.F, sl.dItems + split, n);
                    launched();
                }
            }
        }
    }

    // ------------------------------------------------------------------ direct envelope exchange

    /// driver entry points resolved through the runtime (no link-time dependency on libcuda)
    struct DriverApi {
        CUresult (*MemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr) = nullptr;
        CUresult (*StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
        static DriverApi& get() {
            static DriverApi api = load();
            return api;
        }
        static DriverApi load() {
            DriverApi a;
            cudaDriverEntryPointQueryResult st;
            void* fn = nullptr;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
                a.MemGetAddressRange = reinterpret_cast<decltype(a.MemGetAddressRange)>(fn);
            fn = nullptr;
            if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
                a.StreamWaitValue32 = reinterpret_cast<decltype(a.StreamWaitValue32)>(fn);
            cudaGetLastError();
            return a;
        }
    };

    /// what a rank tells its ring neighbours about the two slabs they store into
    struct FaceExport {
        cudaIpcMemHandle_t hFirst, hLast, hFlags;
        unsigned long long offFirst, offLast, offFlags;
        long long SFirst, SLast;
        int nxlFirst, nxlLast;
        int ok;
    };

    static void ipc_export(void* ptr, cudaIpcMemHandle_t& h, unsigned long long& off) {
        DriverApi& d = DriverApi::get();
        if (!d.MemGetAddressRange) throw PlbError("cuMemGetAddressRange is not available");
        CUdeviceptr base = 0; size_t size = 0;
        if (d.MemGetAddressRange(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) throw PlbError("cuMemGetAddressRange failed");
        // (a small cudaMalloc may be carved out of a larger block: the handle names the block, the mapping starts at its base)
        PLB_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
        off = (unsigned long long)((CUdeviceptr)ptr - base);
    }
    void* ipc_open(const cudaIpcMemHandle_t& h, unsigned long long off) {
        const std::string key((const char*)&h, sizeof h);
        auto it = ipcOpened_.find(key);
        void* base = nullptr;
        if (it != ipcOpened_.end()) base = it->second;
        else {
            PLB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            ipcOpened_[key] = base;
        }
        return (char*)base + off;
    }

    /// After ncclCommInitRank: map the slabs of the two ring neighbours (rank - 1: its last slab; rank + 1: its first) and their
    /// flag words into this process (CUDA IPC; across GPUs the mapping is peer memory over NVLink / NVSwitch), so that the rim
    /// cells of the sweep store straight into them.  If any rank cannot, all ranks fall back to the copy-based exchange.
    void map_neighbours() {
        NcclApi& nc = NcclApi::get();
        const int R = desc_.nRanks;
        FaceExport mine;
        std::memset(&mine, 0, sizeof mine);
        if (direct_) {
            try {
                if (!DriverApi::get().StreamWaitValue32) throw PlbError("cuStreamWaitValue32 is not available");
                dFlags_ = (unsigned int*)dev_alloc(64 * sizeof(unsigned int));
                PLB_CUDA(cudaMemsetAsync(dFlags_, 0, 64 * sizeof(unsigned int), stream_));
                Slab& first = slabs_.front();
                Slab& last = slabs_.back();
                ipc_export(first.F, mine.hFirst, mine.offFirst);
                ipc_export(last.F, mine.hLast, mine.offLast);
                ipc_export(dFlags_, mine.hFlags, mine.offFlags);
                mine.SFirst = first.cellsStore; mine.SLast = last.cellsStore; mine.nxlFirst = first.nxl; mine.nxlLast = last.nxl;
                mine.ok = 1;
            } catch (const std::exception&) { mine.ok = 0; }
        }
        FaceExport* dAll = (FaceExport*)dev_alloc((size_t)R * sizeof(FaceExport));
        std::vector<FaceExport> all((size_t)R);
        PLB_CUDA(cudaMemcpyAsync(dAll + desc_.rank, &mine, sizeof mine, cudaMemcpyHostToDevice, stream_));
        nc.check(nc.AllGather(dAll + desc_.rank, dAll, sizeof(FaceExport), ncclInt8, comm_, stream_), "ncclAllGather");
        PLB_CUDA(cudaMemcpyAsync(all.data(), dAll, (size_t)R * sizeof(FaceExport), cudaMemcpyDeviceToHost, stream_));
        PLB_CUDA(cudaStreamSynchronize(stream_));
        cudaFree(dAll); devBytes_ -= (size_t)R * sizeof(FaceExport);
        double ok = 1.;
        for (const FaceExport& e : all) if (!e.ok) ok = 0.;
        if (ok != 0.) {
            try {
                const FaceExport& lo = all[(size_t)((desc_.rank - 1 + R) % R)];
                const FaceExport& hi = all[(size_t)((desc_.rank + 1) % R)];
                remoteLo_.F = (T*)ipc_open(lo.hLast, lo.offLast); remoteLo_.S = lo.SLast; remoteLo_.nxl = lo.nxlLast;
                remoteLo_.flag = (unsigned int*)ipc_open(lo.hFlags, lo.offFlags) + 32;     // I am its neighbour ABOVE
                remoteHi_.F = (T*)ipc_open(hi.hFirst, hi.offFirst); remoteHi_.S = hi.SFirst; remoteHi_.nxl = hi.nxlFirst;
                remoteHi_.flag = (unsigned int*)ipc_open(hi.hFlags, hi.offFlags);          // I am its neighbour BELOW
            } catch (const std::exception&) { ok = 0.; cudaGetLastError(); }
        }
        all_reduce(&ok, 1, false);
        if (ok != (double)R) {          // somebody could not: everybody copies
            for (auto& kv : ipcOpened_) cudaIpcCloseMemHandle(kv.second);
            ipcOpened_.clear();
            remoteLo_ = Remote(); remoteHi_ = Remote();
            direct_ = false;
        }
    }

    /// The neighbouring ranks' stores of the step just launched have landed in this rank's rim / ghost planes: everything
    /// this rank enqueues from here on (boundary processors, the next sweep, downloads) comes after them.  Normally the
    /// flags are long there -- the neighbours store within the first ~50 us of their sweep, this point lies after the whole of ours.
    void wait_neighbours() {
        if (desc_.nRanks == 1 || G_ == 1) return;
        DriverApi& d = DriverApi::get();
        const Geo gFirst = geo(slabs_.front()), gLast = geo(slabs_.back());
        if (gFirst.lowMode == XMODE_GHOST && !find_slab((slabs_.front().g - 1 + G_) % G_)) {
            if (d.StreamWaitValue32((CUstream)stream_, (CUdeviceptr)(dFlags_ + 0), stepCount_, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                throw PlbError("cuStreamWaitValue32 failed");
        }
        if (gLast.highMode == XMODE_GHOST && !find_slab((slabs_.back().g + 1) % G_)) {
            if (d.StreamWaitValue32((CUstream)stream_, (CUdeviceptr)(dFlags_ + 32), stepCount_, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                throw PlbError("cuStreamWaitValue32 failed");
        }
    }

    /// Copy-based envelope exchange, the fallback (replaces BlockCommunicator3D::duplicateOverlaps, multiBlock/serialBlockCommunicator3D.cpp:44-95
    /// and parallelism/parallelBlockCommunicator3D.cpp:251-297).  Only the populations that cross the slab
    /// face travel -- those with c_x = -1 or c_x = +1 (5 of 19, 9 of 27) -- and each is one contiguous
    /// y-z plane of one slot of the SoA storage, so nothing is packed.  Across the face between slab g
    /// (below) and slab g+1 (above):
    ///   after an even step (layout SWAPPED):  g's last plane, slots c_x=-1   -> low ghost plane of g+1
    ///                                         g+1's first plane, slots c_x=+1 -> high ghost plane of g
    ///   after an odd step (layout NORMAL):    high ghost of g, slots c_x=+1   -> low ghost of g+1 -> (unpack) first plane of g+1
    ///                                         low ghost of g+1, slots c_x=-1  -> high ghost of g  -> (unpack) last plane of g
    void exchange_copies(cudaStream_t stream) {
        if (G_ == 1) return;
        const int nFaces = per_[0] ? G_ : G_ - 1;
        const size_t planeBytes = (size_t)P_ * sizeof(T);
        const bool remote = desc_.nRanks > 1;
        NcclApi* n = remote ? &NcclApi::get() : nullptr;
        if (remote) n->check(n->GroupStart(), "ncclGroupStart");
        try {
            for (int face = 0; face < nFaces; ++face) {
                const int gLo = face, gHi = (face + 1) % G_;
                for (int part = 0; part < 2; ++part) {
                    // part 0 carries the slots with c_x = -1, part 1 those with c_x = +1
                    const int cxSel = part == 0 ? -1 : +1;
                    int srcG, dstG, srcPlane, dstPlane;     // plane: 0 low ghost, 1 first, 2 last, 3 high ghost
                    if (swapped_) {
                        if (part == 0) { srcG = gLo; srcPlane = 2; dstG = gHi; dstPlane = 0; }
                        else           { srcG = gHi; srcPlane = 1; dstG = gLo; dstPlane = 3; }
                    } else {
                        // ghost -> ghost; unpack_halo_kernel then moves the valid entries onto the boundary plane
                        if (part == 1) { srcG = gLo; srcPlane = 3; dstG = gHi; dstPlane = 0; }
                        else           { srcG = gHi; srcPlane = 0; dstG = gLo; dstPlane = 3; }
                    }
                    Slab* src = find_slab(srcG);
                    Slab* dst = find_slab(dstG);
                    if (!src && !dst) continue;
                    for (int i = 1; i < Q; ++i) {
                        if (c_of(i, 0) != cxSel) continue;
                        T* sp = src ? plane_ptr(*src, i, srcPlane) : nullptr;
                        T* dp = dst ? plane_ptr(*dst, i, dstPlane) : nullptr;
                        if (src && dst) {
                            PLB_CUDA(cudaMemcpyAsync(dp, sp, planeBytes, cudaMemcpyDeviceToDevice, stream));
                        } else if (src) {
                            n->check(n->Send(sp, planeBytes, ncclInt8, dstG / desc_.slabsPerRank, comm_, stream), "ncclSend");
                        } else {
                            n->check(n->Recv(dp, planeBytes, ncclInt8, srcG / desc_.slabsPerRank, comm_, stream), "ncclRecv");
                        }
                    }
                }
            }
        } catch (...) {
            // never leave the NCCL group open: every later call on this thread would be swallowed by it
            if (remote) n->GroupEnd();
            failed_ = true;
            throw;
        }
        if (remote) n->check(n->GroupEnd(), "ncclGroupEnd");
    }

    /// second half of the exchange after an odd step (on the lattice's stream, once the copies have arrived)
    void exchange_unpack() {
        if (G_ == 1 || direct_) return;
        if (!swapped_) {
            for (Slab& sl : slabs_) {
                const Geo g = geo(sl);
                const unsigned grid = (unsigned)((P_ + 127) / 128);
                if (g.lowMode == XMODE_GHOST) { unpack_halo_kernel<T, Q><<<grid, 128, 0, stream_>>>(sl.F, g, 0); launched(); }
                if (g.highMode == XMODE_GHOST) { unpack_halo_kernel<T, Q><<<grid, 128, 0, stream_>>>(sl.F, g, 1); launched(); }
            }
        }
    }

    static int c_of(int i, int d) { return Q == 19 ? c19(i, d) : c27(i, d); }
    Slab* find_slab(int g) {
        for (Slab& sl : slabs_) if (sl.g == g) return &sl;
        return nullptr;
    }
    T* plane_ptr(Slab& sl, int slot, int which) {
        const int X = which == 0 ? 0 : (which == 1 ? 1 : (which == 2 ? sl.nxl : sl.nxl + 1));
        return sl.F + (size_t)slot * sl.cellsStore + (size_t)X * P_;
    }

    void all_reduce(double* v, int n, bool isMax) {
        NcclApi& nc = NcclApi::get();
        if (!comm_) throw PlbError("plb_b200_comm_init has not been called");
        PLB_CUDA(cudaMemcpyAsync(dReduce_, v, n * sizeof(double), cudaMemcpyHostToDevice, stream_));
        nc.check(nc.AllReduce(dReduce_, dReduce_, n, ncclDouble, isMax ? ncclMax : ncclSum, comm_, stream_), "ncclAllReduce");
        PLB_CUDA(cudaMemcpyAsync(v, dReduce_, n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        PLB_CUDA(cudaStreamSynchronize(stream_));
    }

    // ----------------------------------------------------------------------- transfer, analysis

    /// upload / download in the CoProcessor3D wire format.  The AoS <-> SoA transposition runs on the
    /// device; chunks alternate between two staging buffers so that the PCIe copy of one chunk (copy
    /// stream) overlaps the transposition kernel of the other (lattice stream).
    void transfer(const int box[6], void* data, size_t bytes, bool up) {
        check_box(box);
        if (bytes != local_bytes(box)) throw PlbError("buffer size does not match the box");
        if (!data && bytes) throw PlbError("null buffer");
        if (up) ext_detach();
        sync_setup(); materialize();
        if (!copyStream_) {
            PLB_CUDA(cudaStreamCreateWithFlags(&copyStream_, cudaStreamNonBlocking));
            for (int k = 0; k < 2; ++k) {
                PLB_CUDA(cudaEventCreateWithFlags(&copied_[k], cudaEventDisableTiming));
                PLB_CUDA(cudaEventCreateWithFlags(&transposed_[k], cudaEventDisableTiming));
            }
        }
        const size_t half = std::max(XFER_CHUNK_BYTES, (size_t)(box[3] - box[2] + 1) * (box[5] - box[4] + 1) * Q * sizeof(T));
        char* st = (char*)stage(2 * half);
        char* host = (char*)data;
        int chunk = 0;
        for (Slab& sl : slabs_) {
            BoxL b;
            if (!local_box(sl, box, b)) continue;
            const size_t planeBytes = (size_t)b.ny() * b.nz() * Q * sizeof(T);
            const int planesPerChunk = (int)std::max<size_t>(1, half / planeBytes);
            for (int X = b.X0; X <= b.X1; X += planesPerChunk, ++chunk) {
                BoxL cb = b;
                cb.X0 = X; cb.X1 = std::min(b.X1, X + planesPerChunk - 1);
                const size_t nb = (size_t)cb.cells() * Q * sizeof(T);
                const int k = chunk & 1;
                T* buf = (T*)(st + k * half);
                const unsigned grid = (unsigned)((cb.cells() + 127) / 128);
                if (up) {
                    if (chunk >= 2) PLB_CUDA(cudaStreamWaitEvent(copyStream_, transposed_[k], 0));   // buffer free again
                    PLB_CUDA(cudaMemcpyAsync(buf, host, nb, cudaMemcpyHostToDevice, copyStream_));
                    PLB_CUDA(cudaEventRecord(copied_[k], copyStream_));
                    PLB_CUDA(cudaStreamWaitEvent(stream_, copied_[k], 0));
                    transfer_kernel<T, Q, true><<<grid, 128, 0, stream_>>>(ctx(sl), sl.F, cb, buf);
                    launched();
                    PLB_CUDA(cudaEventRecord(transposed_[k], stream_));
                } else {
                    if (chunk >= 2) PLB_CUDA(cudaStreamWaitEvent(stream_, copied_[k], 0));
                    transfer_kernel<T, Q, false><<<grid, 128, 0, stream_>>>(ctx(sl), sl.F, cb, buf);
                    launched();
                    PLB_CUDA(cudaEventRecord(transposed_[k], stream_));
                    PLB_CUDA(cudaStreamWaitEvent(copyStream_, transposed_[k], 0));
                    PLB_CUDA(cudaMemcpyAsync(host, buf, nb, cudaMemcpyDeviceToHost, copyStream_));
                    PLB_CUDA(cudaEventRecord(copied_[k], copyStream_));
                }
                host += nb;
            }
        }
        PLB_CUDA(cudaStreamSynchronize(copyStream_));
        PLB_CUDA(cudaStreamSynchronize(stream_));
    }

    double box_sum(const int box[6], int what) {
        check_box(box);
        sync_setup(); materialize();
        double total = 0.;
        for (Slab& sl : slabs_) {
            BoxL b;
            if (!local_box(sl, box, b)) continue;
            const int grid = (int)std::min<long long>((b.cells() + 255) / 256, 1024);
            double* part = (double*)stage(grid * sizeof(double));
            box_sum_kernel<T, Q><<<grid, 256, 0, stream_>>>(ctx(sl), b, what, part);
            launched();
            std::vector<double> h(grid);
            PLB_CUDA(cudaMemcpyAsync(h.data(), part, grid * sizeof(double), cudaMemcpyDeviceToHost, stream_));
            PLB_CUDA(cudaStreamSynchronize(stream_));
            for (double v : h) total += v;
        }
        if (desc_.nRanks > 1) all_reduce(&total, 1, false);
        return total;
    }
};

}  // namespace plb200
