// palabos_b200/csrc/kernels.cuh -- device kernels around the sweep: envelope exchange, boundary processors, set-up, transfer,
// analysis (storage layout and addressing: cell_io.cuh; the sweep itself: sweep_kernel.cuh).
#pragma once
#include "cell_io.cuh"

namespace plb200 {

// ------------------------------------------------------------------- envelope exchange

/// Second half of the envelope exchange after an odd step.  The populations that left a slab
/// through its face were written into its ghost plane and have been copied (cudaMemcpy / NCCL) into
/// the ghost plane of the receiving slab; this kernel moves them onto the receiver's boundary plane
/// -- except where the sending cell does not exist (y or z rim of a non-periodic lattice): there the
/// receiver's own cell kept the population under the on-site rule and must not be overwritten.
///   side 0: low face,  slots with c_x = +1, ghost plane X = 0      -> plane X = 1
///   side 1: high face, slots with c_x = -1, ghost plane X = nxl+1  -> plane X = nxl
template <class T, int Q, int I = 1>
struct Unpack {
    __device__ __forceinline__ static void run(T* __restrict__ F, const Geo& g, const Nbr& n, int side, int src, int dst) {
        using L = LB<T, Q>;
        constexpr int cx = L::c(I, 0), cy = L::c(I, 1), cz = L::c(I, 2);
        if ((cx == 1 && side == 0) || (cx == -1 && side == 1)) {
            if (nbr_valid<0, -cy, -cz>(n)) F[(long long)I * g.S + dst] = F[(long long)I * g.S + src];
        }
        Unpack<T, Q, I + 1>::run(F, g, n, side, src, dst);
    }
};
template <class T, int Q>
struct Unpack<T, Q, Q> {
    __device__ __forceinline__ static void run(T*, const Geo&, const Nbr&, int, int, int) {}
};

template <class T, int Q>
__global__ void unpack_halo_kernel(T* F, Geo g, int side) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= g.P) return;
    const int y = r / g.nz, z = r - y * g.nz;
    const int X = side == 0 ? 1 : g.nxl;
    const Nbr n = make_nbr(g, X, y, z);
    const int dst = X * g.P + r;
    const int src = (side == 0 ? 0 : g.nxl + 1) * g.P + r;
    Unpack<T, Q>::run(F, g, n, side, src, dst);
}

// ------------------------------------------------------------------- boundary processors

template <class T>
struct CellCtx {
    const T* F;
    const uint8_t* kind;
    const int* bcIdx;
    const Dyn<T>* dyns;
    const Bc<T>* bcs;
    Geo g;
    bool swapped;
    T* hbuf;             ///< handoff buffer [Q][hStride] by boundary record (see KIND_HANDOFF)
    long long hStride;
    T* ext;              ///< external-moment sweep: the rhoBar / j fields (rhoBar [cells] | j_x | j_y | j_z, extStride apart), else nullptr:
    long long extStride; ///< the boundary passes leave Dynamics::computeRhoBarJ of the cells they complete there (see plane_stage_kernel)
};

template <class T, int Q>
struct CellView {
    Dyn<T> d;
    Bc<T> b;
    bool hasBc;
    T f[Q];
};

/// Dynamics and boundary record of a cell; populations only if `needPops`.
template <class T, int Q>
__device__ __forceinline__ void fetch_cell(const CellCtx<T>& cx, int c, bool needDensityPops, bool needVelocityPops, CellView<T, Q>& v) {
    using L = LB<T, Q>;
    const int kb = cx.kind[c];
    v.d = cx.dyns[kb & KIND_MASK];
    const int bi = cx.bcIdx[c];
    v.hasBc = bi >= 0;
    if (v.hasBc) v.b = cx.bcs[bi];
    else { v.b.kind = -1; v.b.dir = 0; v.b.ori = 1; v.b.u[0] = v.b.u[1] = v.b.u[2] = (T)0; v.b.rhoBar = (T)0; }
    const bool need = (needDensityPops && L::density_reads_pops(v.d.type, v.hasBc, v.b.kind))
                   || (needVelocityPops && L::velocity_reads_pops(v.d.type, v.hasBc, v.b.kind));
    if (need) {
        int X, y, z;
        decode_cell(cx.g, c, X, y, z);
        const Nbr n = make_nbr(cx.g, X, y, z);
        CellIO<T, Q>::load(cx.F, cx.g, n, c, cx.swapped, v.f);
    } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) v.f[i] = (T)0;
    }
}

template <class T, int Q>
__device__ __forceinline__ T density_at(const CellCtx<T>& cx, int c) {
    CellView<T, Q> v;
    fetch_cell<T, Q>(cx, c, true, false, v);
    return LB<T, Q>::cell_density(v.d, v.hasBc, v.b, v.f);
}
template <class T, int Q>
__device__ __forceinline__ void velocity_at(const CellCtx<T>& cx, int c, T u[3]) {
    CellView<T, Q> v;
    fetch_cell<T, Q>(cx, c, false, true, v);
    LB<T, Q>::cell_velocity(v.d, v.hasBc, v.b, v.f, u);
}

__device__ __forceinline__ int axis_stride(const Geo& g, int dir) { return dir == 0 ? g.P : (dir == 1 ? g.nz : 1); }

/// fd::DirectedGradients3D<..,orthogonal=true>::o1_velocityDerivative finiteDifference/finiteDifference3D.h:64-82
template <class T, int Q>
__device__ __forceinline__ void grad_normal(const CellCtx<T>& cx, int c, int dir, int ori, T out[3]) {
    T u0[3], u1[3];
    velocity_at<T, Q>(cx, c, u0);
    velocity_at<T, Q>(cx, c - ori * axis_stride(cx.g, dir), u1);
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = -ori * (-u0[k] + u1[k]);
}
/// ..orthogonal=false: central difference finiteDifference3D.h:158-178
template <class T, int Q>
__device__ __forceinline__ void grad_tangent(const CellCtx<T>& cx, int c, int dir, T out[3]) {
    T up[3], um[3];
    velocity_at<T, Q>(cx, c + axis_stride(cx.g, dir), up);
    velocity_at<T, Q>(cx, c - axis_stride(cx.g, dir), um);
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = (up[k] - um[k]) / (T)2;
}

/// common tail of Outer{Edge,Corner}Functional3D::processCell
/// (boundaryCondition/finiteDifferenceBoundaryProcessor3D.hh:189-221, :287-313)
template <class T, int Q>
__device__ __forceinline__ void fd_regularize(const CellCtx<T>& cx, T* F, int c, T rho, T (*dA_uB)[3]) {
    using L = LB<T, Q>;
    CellView<T, Q> v;
    fetch_cell<T, Q>(cx, c, false, true, v);
    const T omega = L::is_inert(v.d.type) ? (T)0 : v.d.omega;
    const T sToPi = -rho / L::invCs2() / omega;
    T pi[6], vel[3], j[3];
    pi[0] = (T)2 * dA_uB[0][0] * sToPi;
    pi[3] = (T)2 * dA_uB[1][1] * sToPi;
    pi[5] = (T)2 * dA_uB[2][2] * sToPi;
    pi[1] = (dA_uB[0][1] + dA_uB[1][0]) * sToPi;
    pi[2] = (dA_uB[0][2] + dA_uB[2][0]) * sToPi;
    pi[4] = (dA_uB[1][2] + dA_uB[2][1]) * sToPi;
    L::cell_velocity(v.d, v.hasBc, v.b, v.f, vel);
    if (L::vel_is_j(v.d.type)) { j[0] = vel[0]; j[1] = vel[1]; j[2] = vel[2]; }     // finiteDifferenceBoundaryProcessor3D.hh:200-205
    else { j[0] = rho * vel[0]; j[1] = rho * vel[1]; j[2] = rho * vel[2]; }
    const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
    if (L::is_inert(v.d.type)) return;      // regularize is a no-op there: the populations stay
    L::regularize(v.d.type, v.f, L::rhoBarOf(rho), j, jSqr, pi);
    int X, y, z;
    decode_cell(cx.g, c, X, y, z);
    const Nbr n = make_nbr(cx.g, X, y, z);
    CellIO<T, Q>::store(F, cx.g, n, c, cx.swapped, v.f);
    if (cx.ext && v.hasBc) {    // (see plane_stage_kernel)
        T rb, jj[3];
        L::cell_rhoBar_j(v.d, true, v.b, v.f, rb, jj);
        cx.ext[c] = rb; cx.ext[cx.extStride + c] = jj[0]; cx.ext[2 * cx.extStride + c] = jj[1]; cx.ext[3 * cx.extStride + c] = jj[2];
    }
}

/// One stage of the level-0 data processors (AtomicBlock3D::executeInternalProcessors,
/// atomicBlock/atomicBlock3D.cpp:195-222).  The host has ordered the (processor, cell) pairs into
/// stages such that running a stage's items concurrently equals the reference's sequential order.
/// Plane items (local completion of the populations) and finite-difference items (outer edges and
/// corners, which read neighbours) run in separate kernels.
///
/// WrappedLocalBoundaryFunctional3D::process boundaryCondition/wrappedLocalBoundaryProcessor3D.hh:42-62: completePopulations of
/// the plane boundary cells, one thread per cell.  The switch on (direction, orientation) in complete_by_face lets the compiler
/// specialise completePopulations for each of the six faces, as the reference's templates <direction, orientation> do; the
/// items of a processor are consecutive, so a warp works on one face.
/// The same kernel completes the cells of the dynamics-based boundary managers (completion at collision time,
/// boundaryDynamics.hh:66-75 prepareCollision) right BEFORE the sweep that collides them: between two steps their
/// populations stay raw, as in the reference.
/// Cost at 512^3 (profiles/r02_plane_pass.md): the x- and y-faces are rows of consecutive cells (35 us per face of 262 k cells);
/// a z-face is one cell per row -- 19 gathers and 19 scatters of 8 bytes, each in a DRAM line of its own -- and is bound by
/// the number of outstanding misses an SM can track, not by bandwidth.
template <class T, int Q, bool SWAPPED>
__global__ void __launch_bounds__(128, 6) plane_stage_kernel(CellCtx<T> cx, T* F, const Item* __restrict__ items, int nItems) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nItems) return;
    const int c = items[t].cell, bi = items[t].bc;
    const bool handoff = items[t].c != 0;           // completed populations go to the handoff buffer, not back into the lattice
    if (bi < 0) return;
    const int kb = cx.kind[c];
    const Bc<T> b = cx.bcs[bi];
    int X, y, z;
    decode_cell(cx.g, c, X, y, z);
    const Nbr n = make_nbr(cx.g, X, y, z);
    T f[Q];
    CellIO<T, Q>::load(F, cx.g, n, c, SWAPPED, f);
    const Dyn<T> d = cx.dyns[kb & KIND_MASK];
    complete_by_face<T, Q>(d.type, b, f);
    if (handoff) {
#pragma unroll
        for (int i = 0; i < Q; ++i) cx.hbuf[(long long)i * cx.hStride + bi] = f[i];
    } else {
        CellIO<T, Q>::store(F, cx.g, n, c, SWAPPED, f);
    }
    if (cx.ext) {
        // BoxRhoBarJfunctional3D (level 3) of this cell, while its completed populations are at hand: the next external-moment
        // sweep reads the four values instead of recomputing them with one lane of a warp (a z-face has one such cell per row)
        T rhoBar, j[3];
        LB<T, Q>::cell_rhoBar_j(d, true, b, f, rhoBar, j);
        cx.ext[c] = rhoBar; cx.ext[cx.extStride + c] = j[0]; cx.ext[2 * cx.extStride + c] = j[1]; cx.ext[3 * cx.extStride + c] = j[2];
    }
}

/// The handoff buffer back into the lattice: before anything but the sweep reads or writes the populations of those cells
/// (downloads, reductions, re-initialisation, set-up changes, a second initialize()).
template <class T, int Q, bool SWAPPED>
__global__ void __launch_bounds__(128) handoff_scatter_kernel(CellCtx<T> cx, T* F, const Item* __restrict__ items, int nItems) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nItems) return;
    const int c = items[t].cell, bi = items[t].bc;
    if (bi < 0 || items[t].c == 0) return;
    int X, y, z;
    decode_cell(cx.g, c, X, y, z);
    const Nbr n = make_nbr(cx.g, X, y, z);
    T f[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) f[i] = cx.hbuf[(long long)i * cx.hStride + bi];
    CellIO<T, Q>::store(F, cx.g, n, c, SWAPPED, f);
}

/// Mark the handoff cells in the kind bytes.
static __global__ void flag_handoff_cells_kernel(const Item* items, int n, uint8_t* kind) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && items[t].c != 0) kind[items[t].cell] |= (uint8_t)KIND_HANDOFF;
}

template <class T, int Q>
__global__ void __launch_bounds__(64) fd_stage_kernel(CellCtx<T> cx, T* F, const Item* items, int nItems) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nItems) return;
    const Item it = items[t];
    const int c = it.cell;
    if (it.cls == PROC_PLANE_FD) {
        // PlaneFdBoundaryFunctional3D::processCell finiteDifferenceBoundaryProcessor3D.hh:68-117 (createInterpBoundaryCondition3D):
        // density and velocity of the cell through its Dirichlet dynamics, velocity gradients one-sided along the normal
        // and central along the face, then regularize
        const int dir = it.a, ori = it.b;
        const T rho = density_at<T, Q>(cx, c);
        T dA_uB[3][3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (d == dir) grad_normal<T, Q>(cx, c, d, ori, dA_uB[d]);
            else grad_tangent<T, Q>(cx, c, d, dA_uB[d]);
        }
        fd_regularize<T, Q>(cx, F, c, rho, dA_uB);
    } else if (it.cls == PROC_EDGE) {
        // OuterVelocityEdgeFunctional3D::processCell finiteDifferenceBoundaryProcessor3D.hh:164-226
        const int plane = it.a, n1 = it.b, n2 = it.c;
        const int d1 = (plane + 1) % 3, d2 = (plane + 2) % 3;
        const T rho10 = density_at<T, Q>(cx, c - n1 * axis_stride(cx.g, d1));
        const T rho01 = density_at<T, Q>(cx, c - n2 * axis_stride(cx.g, d2));
        const T rho = (T)0.5 * (rho01 + rho10);
        T g0[3], g1[3], g2[3], dA_uB[3][3];
        grad_tangent<T, Q>(cx, c, plane, g0);
        grad_normal<T, Q>(cx, c, d1, n1, g1);
        grad_normal<T, Q>(cx, c, d2, n2, g2);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dA_uB[0][k] = plane == 0 ? g0[k] : (d1 == 0 ? g1[k] : g2[k]);
            dA_uB[1][k] = plane == 1 ? g0[k] : (d1 == 1 ? g1[k] : g2[k]);
            dA_uB[2][k] = plane == 2 ? g0[k] : (d1 == 2 ? g1[k] : g2[k]);
        }
        fd_regularize<T, Q>(cx, F, c, rho, dA_uB);
    } else {
        // OuterVelocityCornerFunctional3D::processCell finiteDifferenceBoundaryProcessor3D.hh:264-318
        const int xn = it.a, yn = it.b, zn = it.c;
        const T rho100 = density_at<T, Q>(cx, c - xn * cx.g.P);
        const T rho010 = density_at<T, Q>(cx, c - yn * cx.g.nz);
        const T rho001 = density_at<T, Q>(cx, c - zn);
        const T rho = (T)1 / (T)3 * (rho001 + rho010 + rho100);
        T dA_uB[3][3];
        grad_normal<T, Q>(cx, c, 0, xn, dA_uB[0]);
        grad_normal<T, Q>(cx, c, 1, yn, dA_uB[1]);
        grad_normal<T, Q>(cx, c, 2, zn, dA_uB[2]);
        fd_regularize<T, Q>(cx, F, c, rho, dA_uB);
    }
}

// ------------------------------------------------------------- set-up, transfer, analysis

struct BoxL {             ///< a box in storage coordinates of one slab: X in [X0,X1], y, z inclusive
    int X0, X1, y0, y1, z0, z1;
    __host__ __device__ int ny() const { return y1 - y0 + 1; }
    __host__ __device__ int nz() const { return z1 - z0 + 1; }
    __host__ __device__ long long cells() const { return (long long)(X1 - X0 + 1) * ny() * nz(); }
};

__device__ __forceinline__ int box_cell(const Geo& g, const BoxL& b, long long k, int& X, int& y, int& z) {
    const int bz = b.nz(), by = b.ny();
    const int kz = (int)(k % bz);
    const long long r = k / bz;
    const int ky = (int)(r % by);
    const int kx = (int)(r / by);
    X = b.X0 + kx; y = b.y0 + ky; z = b.z0 + kz;
    return (X * g.ny + y) * g.nz + z;
}

/// initializeAtEquilibrium(lattice, box, rho, u) -> iniCellAtEquilibrium core/cell.hh:124-132.
/// field == nullptr: constant (rho, u); otherwise field = [cells of the box][rho,ux,uy,uz] in double
/// (converted to T like the reference's functional overload does).
template <class T, int Q>
__global__ void init_equilibrium_kernel(CellCtx<T> cx, T* F, BoxL box, T rho, T ux, T uy, T uz, const double* field) {
    using L = LB<T, Q>;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= box.cells()) return;
    int X, y, z;
    const int c = box_cell(cx.g, box, k, X, y, z);
    T u[3] = { ux, uy, uz };
    if (field) { rho = (T)field[4 * k]; u[0] = (T)field[4 * k + 1]; u[1] = (T)field[4 * k + 2]; u[2] = (T)field[4 * k + 3]; }
    T f[Q];
    L::ini_equilibrium(cx.dyns[cx.kind[c] & KIND_MASK].type, f, rho, u);
    const Nbr n = make_nbr(cx.g, X, y, z);
    CellIO<T, Q>::store(F, cx.g, n, c, cx.swapped, f);
}

/// CoProcessor3D::send / receive wire format (coProcessors/coProcessor3D.h:79-92): raw T,
/// [x][y][z][q], values as stored (f - t_i, Cell::serialize core/cell.hh:103-111).
template <class T, int Q, bool UP>
__global__ void transfer_kernel(CellCtx<T> cx, T* F, BoxL box, T* wire) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= box.cells()) return;
    int X, y, z;
    const int c = box_cell(cx.g, box, k, X, y, z);
    const Nbr n = make_nbr(cx.g, X, y, z);
    T f[Q];
    if (UP) {
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = wire[k * Q + i];
        CellIO<T, Q>::store(F, cx.g, n, c, cx.swapped, f);
    } else {
        CellIO<T, Q>::load(F, cx.g, n, c, cx.swapped, f);
#pragma unroll
        for (int i = 0; i < Q; ++i) wire[k * Q + i] = f[i];
    }
}

/// rho and u of every cell of a box through Cell::computeDensity / computeVelocity; out = [cells][4] in T
template <class T, int Q>
__global__ void density_velocity_kernel(CellCtx<T> cx, BoxL box, T* out) {
    using L = LB<T, Q>;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= box.cells()) return;
    int X, y, z;
    const int c = box_cell(cx.g, box, k, X, y, z);
    CellView<T, Q> v;
    fetch_cell<T, Q>(cx, c, true, true, v);
    T u[3];
    out[4 * k] = L::cell_density(v.d, v.hasBc, v.b, v.f);
    L::cell_velocity(v.d, v.hasBc, v.b, v.f, u);
    out[4 * k + 1] = u[0]; out[4 * k + 2] = u[1]; out[4 * k + 3] = u[2];
}

/// BoxRhoBarJfunctional3D::processGenericBlocks (dataProcessors/dataAnalysisFunctional3D.hh:528-548): Dynamics::computeRhoBarJ
/// of every cell of a box into the fields  ext = rhoBar [cells] | j_x | j_y | j_z  (stride apart, storage cell index)
template <class T, int Q>
__device__ __forceinline__ void rhobar_j_cell(const CellCtx<T>& cx, int c, int X, int y, int z, T* ext, long long stride);

template <class T, int Q>
__global__ void __launch_bounds__(128) rhobar_j_kernel(CellCtx<T> cx, BoxL box, T* ext, long long stride) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= box.cells()) return;
    int X, y, z;
    const int c = box_cell(cx.g, box, k, X, y, z);
    rhobar_j_cell<T, Q>(cx, c, X, y, z, ext, stride);
}

/// the same for a list of cells (those that carry boundary dynamics: the fused sweep recomputes the others itself)
template <class T, int Q>
__global__ void __launch_bounds__(128) rhobar_j_list_kernel(CellCtx<T> cx, const int* cells, int n, T* ext, long long stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = cells[t];
    if (c < 0) return;
    int X, y, z;
    decode_cell(cx.g, c, X, y, z);
    rhobar_j_cell<T, Q>(cx, c, X, y, z, ext, stride);
}

/// external-moment sweep: cells whose moments the fused sweep must read from the fields -- boundary dynamics on top,
/// or outside the bulk box (the envelope layer)
static __global__ void flag_ext_cells_kernel(uint8_t* kind, const int* bcIdx, Geo g, BoxL bulk, int nCells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    int X, y, z;
    decode_cell(g, c, X, y, z);
    const bool inBulk = X >= bulk.X0 && X <= bulk.X1 && y >= bulk.y0 && y <= bulk.y1 && z >= bulk.z0 && z <= bulk.z1;
    if (!inBulk || bcIdx[c] >= 0) kind[c] |= (uint8_t)KIND_PLANE;
}

template <class T, int Q>
__device__ __forceinline__ void rhobar_j_cell(const CellCtx<T>& cx, int c, int X, int y, int z, T* ext, long long stride) {
    using L = LB<T, Q>;
    const int kb = cx.kind[c];
    const Dyn<T> d = cx.dyns[kb & KIND_MASK];
    const int bi = cx.bcIdx[c];
    const bool hasBc = bi >= 0;
    Bc<T> b;
    if (hasBc) b = cx.bcs[bi];
    else { b.kind = -1; b.dir = 0; b.ori = 1; b.u[0] = b.u[1] = b.u[2] = (T)0; b.rhoBar = (T)0; }
    T f[Q];
    if (L::rhoBar_j_reads_pops(d.type, hasBc, b.kind)) {
        const Nbr n = make_nbr(cx.g, X, y, z);
        CellIO<T, Q>::load(cx.F, cx.g, n, c, cx.swapped, f);
    } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = (T)0;
    }
    T rhoBar, j[3];
    L::cell_rhoBar_j(d, hasBc, b, f, rhoBar, j);
    ext[c] = rhoBar; ext[stride + c] = j[0]; ext[2 * stride + c] = j[1]; ext[3 * stride + c] = j[2];
}

/// the fields across the ABI: wire = T [cells of the box][rhoBar, jx, jy, jz]
template <class T, bool UP>
__global__ void ext_transfer_kernel(Geo g, BoxL box, T* ext, long long stride, T* wire) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= box.cells()) return;
    int X, y, z;
    const int c = box_cell(g, box, k, X, y, z);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (UP) ext[m * stride + c] = wire[4 * k + m];
        else wire[4 * k + m] = ext[m * stride + c];
    }
}

/// BoxSumEnergyFunctional3D (dataProcessors/dataAnalysisFunctional3D.hh:97-123: sum of |u|^2, double
/// accumulator) and BoxSumRhoBarFunctional3D (:62-76).  what: 0 = |u|^2, 1 = rhoBar.
/// partials[blockIdx.x] = block sum; the host adds the partials in order.
template <class T, int Q>
__global__ void __launch_bounds__(256) box_sum_kernel(CellCtx<T> cx, BoxL box, int what, double* partials) {
    using L = LB<T, Q>;
    __shared__ double sh[8];
    double s = 0.;
    const long long total = box.cells();
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        int X, y, z;
        const int c = box_cell(cx.g, box, k, X, y, z);
        CellView<T, Q> v;
        fetch_cell<T, Q>(cx, c, what == 1, what == 0, v);
        if (what == 0) {
            T u[3];
            L::cell_velocity(v.d, v.hasBc, v.b, v.f, u);
            const T uNormSqr = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
            s += (double)uNormSqr;
        } else {
            s += (double)L::cell_rhoBar(v.d, v.hasBc, v.b, v.f);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = sh[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) a += sh[w];
        partials[blockIdx.x] = a;
    }
}

/// setBoundaryVelocity (what = 0; dataInitializerFunctional3D.hh:830-845 -> StoreVelocityDynamics::defineVelocity, boundaryDynamics.hh:333-339)
/// and setBoundaryDensity (what = 1; StoreDensityDynamics::defineDensity :239-243) on the records of the cells of a box, on the
/// device: the time loop of an unsteady inlet changes values every step, and a host walk over the records does not belong there.
template <class T, int Q>
__global__ void set_bc_values_kernel(const int* __restrict__ bcIdx, Bc<T>* bcs, Geo g, BoxL box, int what, T u0, T u1, T u2, T rhoBar) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= box.cells()) return;
    int X, y, z;
    const int c = box_cell(g, box, k, X, y, z);
    const int bi = bcIdx[c];
    if (bi < 0) return;
    Bc<T>& b = bcs[bi];
    const bool isRho = LB<T, Q>::is_rho_plane(b.kind);
    if ((what == 0) == isRho) return;
    if (what == 0) { b.u[0] = u0; b.u[1] = u1; b.u[2] = u2; }
    else b.rhoBar = rhoBar;
}

/// Scatter the sparse (cell -> boundary record) map into the dense per-cell index.
static __global__ void scatter_bc_index_kernel(const int* cells, int n, int* bcIdx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = cells[t];
    if (c >= 0) bcIdx[c] = t;          // dead records (cell -1: defineDynamics replaced the composite) are not indexed
}

}  // namespace plb200
