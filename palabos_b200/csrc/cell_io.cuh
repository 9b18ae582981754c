// palabos_b200/csrc/cell_io.cuh -- storage layout, addressing and the argument block of the sweep (shared by every kernel).
//
// Storage (per slab, all in HBM): structure-of-arrays populations  F[slot][X][y][z]  with z fastest
// (the reference's index order, atomicBlock/blockLattice3D.hh:342-349), X = local x + 1 so that
// planes X = 0 and X = nxl+1 are the ghost planes towards the neighbouring slabs; one byte per cell
// with the index of its dynamics; one int32 per cell with the index of its boundary record or -1.
//
// Streaming is IN PLACE with the AA pattern (one read and one write per population per step):
//   even step  (layout NORMAL -> SWAPPED): read slot i of the own cell, collide, write f*_i into
//              slot opp(i) of the own cell.  Purely local.
//   odd step   (layout SWAPPED -> NORMAL): population i that arrives at cell x sits in slot opp(i)
//              of cell x - c_i; read it there, collide, write f*_i into slot i of cell x + c_i.
//              Each storage location is read and written by exactly one thread.
// A population whose neighbour lies outside a non-periodic edge of the global domain stays on its
// cell in the opposite slot -- the reference's rule for the rim of the sweep domain (collide()+revert()
// followed by the bounds-checked swaps of boundaryStream(), atomicBlock/blockLattice3D.hh:173-185,
// :403-427).  In the SWAPPED layout the post-stream population (i, x) of the reference is therefore found at
//   valid(x - c_i) ? (slot opp(i), cell x - c_i) : (slot i, cell x)             ("accessor")
// which is how the boundary, transfer and analysis kernels address populations after an even step.
#pragma once
#include <cfloat>
#include <cuda_runtime.h>
#include "lb_core.cuh"

namespace plb200 {

// Launch shape and cache policy of the sweep, per value type and velocity set: the measured optima on B200 at 512^3
// (profiles/r01_sweep_lab.log, profiles/r02_sweep_lab.md).  512-thread blocks for both halves; many more blocks than are
// resident (grid = SMs x factor), which staggers the load / compute / store phases of the resident blocks.
//   registers   double: 128 per thread, one block per SM (D3Q19 needs 38 for the populations alone);
//               float: 64, two blocks per SM -- the fp32 sweeps are short of warps, not of registers
//               (D3Q19 BGK 4.52 -> 5.43 TB/s, D3Q27 Smagorinsky 3.88 -> 4.31 TB/s)
//   L2 prefetch one tile ahead in the odd (neighbour) half: pays for D3Q19 fp64 only (6.96 -> 6.59 ms); elsewhere the 2 x Q
//               extra instructions cost more than the latency they hide (D3Q27 fp64 odd half 10.97 -> 10.19 ms without,
//               D3Q27 fp32 8.32 -> 7.16 ms)
// The macros exist for tools/sweep_lab.cu, which explores the alternatives.
#ifndef PLB_SWEEP_THREADS_EVEN
#define PLB_SWEEP_THREADS_EVEN 512
#endif
#ifndef PLB_SWEEP_THREADS_ODD
#define PLB_SWEEP_THREADS_ODD 512
#endif
#ifndef PLB_SWEEP_LD
#define PLB_SWEEP_LD(p) (*(p))
#endif
#ifndef PLB_SWEEP_ST
#define PLB_SWEEP_ST(p, v) (*(p) = (v))
#endif
__host__ __device__ constexpr int sweep_threads(bool odd) { return odd ? PLB_SWEEP_THREADS_ODD : PLB_SWEEP_THREADS_EVEN; }
/// register budget per thread that the launch bounds ask the compiler to keep
template <class T>
__host__ __device__ constexpr int sweep_regs() {
#ifdef PLB_SWEEP_REGS
    return PLB_SWEEP_REGS;
#else
    return sizeof(T) == 4 ? 64 : 128;
#endif
}
template <class T>
__host__ __device__ constexpr int sweep_minblocks(bool odd) {
    return 65536 / sweep_regs<T>() / sweep_threads(odd) > 0 ? 65536 / sweep_regs<T>() / sweep_threads(odd) : 1;
}
/// L2 prefetch distance in tiles of the block (0 = off)
template <class T, int Q>
__host__ __device__ constexpr int sweep_prefetch(bool odd) {
#ifdef PLB_SWEEP_PF_ODD
    return odd ? PLB_SWEEP_PF_ODD : 0;
#else
    return (odd && sizeof(T) == 8 && Q == 19) ? 1 : 0;
#endif
}
/// blocks per SM in the grid (profiles/r02_sweep_lab.md: flat within 1 % between 32 and 128, falling off on both sides)
template <class T, int Q>
__host__ __device__ constexpr int sweep_grid_factor(bool odd) {
#ifdef PLB_SWEEP_GRID
    return PLB_SWEEP_GRID;
#else
    return sizeof(T) == 4 ? (odd ? 16 : 64) : (Q == 19 ? 128 : (odd ? 64 : 128));
#endif
}
constexpr int SWEEP_MAX_THREADS = 1024;

enum XMode : int { XMODE_EDGE = 0, XMODE_GHOST = 1, XMODE_WRAP = 2 };

/// Per-cell kind byte: bits 0-5 = index into the dynamics table; bit 7 (external-moment sweep only) = the moments of the
/// cell are not those of a bulk cell: boundary dynamics on top, or envelope layer (flag_ext_cells_kernel).
/// bit 6 = "handoff" cell: a plane boundary cell that sits alone among the 32 cells of a warp (a z-face: one per row).  The
/// plane pass leaves its completed populations in a compact buffer (indexed by boundary record) instead of scattering 19
/// isolated 8-byte stores into the lattice, and the next sweep takes them from there (sweep_kernel).
constexpr int KIND_MASK = 0x3f;
constexpr int KIND_HANDOFF = 0x40;
constexpr int KIND_PLANE = 0x80;

struct Geo {
    int nxl, ny, nz;      ///< owned planes, rows, columns
    int P;                ///< ny*nz cells per x-plane
    long long S;          ///< slot stride in elements = (nxl+2)*P
    int perY, perZ;       ///< in-kernel index wrap along y / z
    int lowMode, highMode;///< what lies beyond X = 1 / X = nxl (XMode)
};

struct Nbr {
    int oxm, oxp, oym, oyp, ozm, ozp;
    bool vxm, vxp, vym, vyp, vzm, vzp;
};

__device__ __forceinline__ Nbr make_nbr(const Geo& g, int X, int y, int z) {
    Nbr n;
    n.vzm = (z > 0) || g.perZ;          n.ozm = (z > 0) ? -1 : (g.nz - 1);
    n.vzp = (z < g.nz - 1) || g.perZ;   n.ozp = (z < g.nz - 1) ? 1 : -(g.nz - 1);
    n.vym = (y > 0) || g.perY;          n.oym = (y > 0) ? -g.nz : (g.ny - 1) * g.nz;
    n.vyp = (y < g.ny - 1) || g.perY;   n.oyp = (y < g.ny - 1) ? g.nz : -(g.ny - 1) * g.nz;
    if (X > 1) { n.vxm = true; n.oxm = -g.P; }
    else { n.vxm = g.lowMode != XMODE_EDGE; n.oxm = g.lowMode == XMODE_WRAP ? (g.nxl - 1) * g.P : -g.P; }
    if (X < g.nxl) { n.vxp = true; n.oxp = g.P; }
    else { n.vxp = g.highMode != XMODE_EDGE; n.oxp = g.highMode == XMODE_WRAP ? -(g.nxl - 1) * g.P : g.P; }
    return n;
}

template <int SX, int SY, int SZ>
__device__ __forceinline__ bool nbr_valid(const Nbr& n) {
    bool v = true;
    if (SX > 0) v = v && n.vxp;
    if (SX < 0) v = v && n.vxm;
    if (SY > 0) v = v && n.vyp;
    if (SY < 0) v = v && n.vym;
    if (SZ > 0) v = v && n.vzp;
    if (SZ < 0) v = v && n.vzm;
    return v;
}
template <int SX, int SY, int SZ>
__device__ __forceinline__ int nbr_off(const Nbr& n) {
    int o = 0;
    if (SX > 0) o += n.oxp;
    if (SX < 0) o += n.oxm;
    if (SY > 0) o += n.oyp;
    if (SY < 0) o += n.oym;
    if (SZ > 0) o += n.ozp;
    if (SZ < 0) o += n.ozm;
    return o;
}

/// Storage index to which f*_I of cell c streams (odd step), i.e. where the reference's post-stream
/// population (I, c + c_I) lives in the NORMAL layout -- or the opposite slot of c at a domain edge.
template <class T, int Q, int I>
__device__ __forceinline__ long long stream_target(const Geo& g, const Nbr& n, int c) {
    using L = LB<T, Q>;
    constexpr int sx = L::c(I, 0), sy = L::c(I, 1), sz = L::c(I, 2);
    return nbr_valid<sx, sy, sz>(n) ? (long long)I * g.S + (c + nbr_off<sx, sy, sz>(n))
                                    : (long long)L::opp(I) * g.S + c;
}

/// Storage index of the reference's population (I, c) in the given layout (see file header).
template <class T, int Q, int I>
__device__ __forceinline__ long long pop_index(const Geo& g, const Nbr& n, int c, bool swapped) {
    using L = LB<T, Q>;
    if (!swapped || I == 0) return (long long)I * g.S + c;
    return stream_target<T, Q, L::opp(I)>(g, n, c);
}

__device__ __forceinline__ void decode_cell(const Geo& g, int c, int& X, int& y, int& z) {
    X = c / g.P;
    const int r = c - X * g.P;
    y = r / g.nz;
    z = r - y * g.nz;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <class T, int Q, int I = 0>
struct CellIO {
    __device__ __forceinline__ static void load(const T* __restrict__ F, const Geo& g, const Nbr& n, int c, bool swapped, T* f) {
        f[I] = F[pop_index<T, Q, I>(g, n, c, swapped)];
        CellIO<T, Q, I + 1>::load(F, g, n, c, swapped, f);
    }
    __device__ __forceinline__ static void store(T* __restrict__ F, const Geo& g, const Nbr& n, int c, bool swapped, const T* f) {
        F[pop_index<T, Q, I>(g, n, c, swapped)] = f[I];
        CellIO<T, Q, I + 1>::store(F, g, n, c, swapped, f);
    }
    // the sweep's own forms
    __device__ __forceinline__ static void load_local(const T* __restrict__ F, const Geo& g, int c, T* f) {
        f[I] = PLB_SWEEP_LD(F + ((long long)I * g.S + c));
        CellIO<T, Q, I + 1>::load_local(F, g, c, f);
    }
    __device__ __forceinline__ static void store_local_swapped(T* __restrict__ F, const Geo& g, int c, const T* f) {
        const long long k = (long long)LB<T, Q>::opp(I) * g.S + c;
        PLB_SWEEP_ST(F + k, f[I]);
        CellIO<T, Q, I + 1>::store_local_swapped(F, g, c, f);
    }
    __device__ __forceinline__ static void load_pull(const T* __restrict__ F, const Geo& g, const Nbr& n, int c, T* f) {
        const long long k = stream_target<T, Q, LB<T, Q>::opp(I)>(g, n, c);
        f[I] = PLB_SWEEP_LD(F + k);
        CellIO<T, Q, I + 1>::load_pull(F, g, n, c, f);
    }
    __device__ __forceinline__ static void store_push(T* __restrict__ F, const Geo& g, const Nbr& n, int c, const T* f) {
        const long long k = stream_target<T, Q, I>(g, n, c);
        PLB_SWEEP_ST(F + k, f[I]);
        CellIO<T, Q, I + 1>::store_push(F, g, n, c, f);
    }
    // interior cells (every neighbour exists at its plain offset): the slot base plus the neighbour offset
    // is the same for the whole launch, the address is that uniform base plus the cell index.  (Passing the
    // Q byte offsets in the constant bank makes the addresses cheaper still -- 2 instead of ~7 instructions
    // -- but the leaner instruction stream issues the 19 loads in a tighter burst and measured SLOWER,
    // 6.84 vs 6.67 ms at 512^3: profiles/r01_sweep_lab.log.)
    __device__ __forceinline__ static long long interior_target(const Geo& g) {
        using L = LB<T, Q>;
        return (long long)I * g.S + (L::c(I, 0) * g.P + L::c(I, 1) * g.nz + L::c(I, 2));
    }
    __device__ __forceinline__ static void load_pull_interior(const T* __restrict__ F, const Geo& g, int c, T* f) {
        const T* base = F + CellIO<T, Q, LB<T, Q>::opp(I)>::interior_target(g);
        f[I] = PLB_SWEEP_LD(base + c);
        CellIO<T, Q, I + 1>::load_pull_interior(F, g, c, f);
    }
    __device__ __forceinline__ static void store_push_interior(T* __restrict__ F, const Geo& g, int c, const T* f) {
        T* base = F + interior_target(g);
        PLB_SWEEP_ST(base + c, f[I]);
        CellIO<T, Q, I + 1>::store_push_interior(F, g, c, f);
    }
    __device__ __forceinline__ static void prefetch_interior(const T* __restrict__ F, const Geo& g, int c) {
        prefetch_l2(F + interior_target(g) + c);
        CellIO<T, Q, I + 1>::prefetch_interior(F, g, c);
    }
    // the same through host-computed per-slot bases (SweepArgs::ldBase / stBase)
    __device__ __forceinline__ static void load_base(T* const* __restrict__ base, int c, T* f) {
        f[I] = PLB_SWEEP_LD(base[I] + c);
        CellIO<T, Q, I + 1>::load_base(base, c, f);
    }
    __device__ __forceinline__ static void store_base(T* const* __restrict__ base, int c, const T* f) {
        PLB_SWEEP_ST(base[I] + c, f[I]);
        CellIO<T, Q, I + 1>::store_base(base, c, f);
    }
    /// BounceBack::collide (swap f_i <-> f_{i+q/2}, core/dynamics.hh:783-791) folded into the store: no register moves
    __device__ __forceinline__ static void store_base_swapped(T* const* __restrict__ base, int c, const T* f) {
        { constexpr int O = LB<T, Q>::opp(I); PLB_SWEEP_ST(base[I] + c, f[O]); }
        CellIO<T, Q, I + 1>::store_base_swapped(base, c, f);
    }
    __device__ __forceinline__ static void prefetch_base(T* const* __restrict__ base, int c) {
        prefetch_l2(base[I] + c);
        CellIO<T, Q, I + 1>::prefetch_base(base, c);
    }
};

/// L2 prefetch of the cell that this thread will process `delta` cells further on: same slots and
/// (for the odd half) the same neighbour offsets as the current cell -- a hint, exact away from rims.
template <class T, int Q, bool ODD, int I = 0>
struct CellPrefetch {
    __device__ __forceinline__ static void run(const T* __restrict__ F, const Geo& g, const Nbr& n, int c, int delta) {
        const long long k = ODD ? stream_target<T, Q, I>(g, n, c) : (long long)I * g.S + c;
        prefetch_l2(F + k + delta);
        CellPrefetch<T, Q, ODD, I + 1>::run(F, g, n, c, delta);
    }
};
template <class T, int Q, bool ODD>
struct CellPrefetch<T, Q, ODD, Q> {
    __device__ __forceinline__ static void run(const T*, const Geo&, const Nbr&, int, int) {}
};

template <class T, int Q>
struct CellIO<T, Q, Q> {
    __device__ __forceinline__ static void load(const T*, const Geo&, const Nbr&, int, bool, T*) {}
    __device__ __forceinline__ static void store(T*, const Geo&, const Nbr&, int, bool, const T*) {}
    __device__ __forceinline__ static void load_local(const T*, const Geo&, int, T*) {}
    __device__ __forceinline__ static void store_local_swapped(T*, const Geo&, int, const T*) {}
    __device__ __forceinline__ static void load_pull(const T*, const Geo&, const Nbr&, int, T*) {}
    __device__ __forceinline__ static void store_push(T*, const Geo&, const Nbr&, int, const T*) {}
    __device__ __forceinline__ static void load_pull_interior(const T*, const Geo&, int, T*) {}
    __device__ __forceinline__ static void store_push_interior(T*, const Geo&, int, const T*) {}
    __device__ __forceinline__ static void prefetch_interior(const T*, const Geo&, int) {}
    __device__ __forceinline__ static void load_base(T* const*, int, T*) {}
    __device__ __forceinline__ static void store_base(T* const*, int, const T*) {}
    __device__ __forceinline__ static void store_base_swapped(T* const*, int, const T*) {}
    __device__ __forceinline__ static void prefetch_base(T* const*, int) {}
};

// ---------------------------------------------------------------------------------- the sweep


struct Item {
    int cell;            ///< storage cell index
    int bc;              ///< plane items: index of the cell's boundary record (-1: none)
    signed char cls;     ///< ProcClass
    signed char a, b, c; ///< edge: plane, normal1, normal2; corner: xNormal, yNormal, zNormal
};

/// The slab on the other side of a slab face (envelope exchange).  The rim cells of the sweep store the
/// populations that cross the face straight into that slab's storage: another slab of this process, or the slab of another
/// rank mapped through CUDA IPC -- the stores then travel over NVLink while the rest of the sweep runs.  F == nullptr: no
/// slab there, or the copy-based transport (ghost planes + NCCL send/recv + unpack_halo_kernel) is in use.
template <class T>
struct PeerSlab {
    T* F;
    long long S;         ///< its slot stride
    int nxl;             ///< its number of owned planes
};

/// cells [begin, end) of a launch, processed as tiles tile0, tile0 + 1, ... of the launch's tile sequence
struct TileRange { int begin, end, tile0; };

template <class T>
struct SweepArgs {
    T* F;
    const uint8_t* kind;
    const Dyn<T>* dyns;       ///< dynamics table in global memory
    const int* bcIdx;         ///< boundary record index per cell (external-moment sweep: read for KIND_PLANE cells)
    const Bc<T>* bcs;         ///< boundary records
    Geo g;
    TileRange range[3];       ///< the planes next to a neighbouring slab first (their stores have the longest way), then the rest
    int nRanges, nTiles;
    T omega, preFactor;       ///< parameters of dynamics-table entry 0 (the lattice's background dynamics)
    int gatherStats;          ///< BlockStatistics on/off (MultiBlock3D::toggleInternalStatistics)
    double* partials;         ///< [gridDim.x][4] per-block partial sums
    unsigned int* ticket;     ///< arrival counter for the last-block reduction
    double* statsOut;         ///< [4] sum rhoBar, sum uSqr, max uSqr, cell count of this launch
    const T* hbuf;            ///< completed populations of the handoff cells, [Q][hStride] by boundary record; nullptr: nothing to take over
    long long hStride;
    const T* ext;             ///< MAIN == DYN_EXTERNAL: rhoBar [cells] followed by j_x, j_y, j_z [cells] each (extStride apart)
    long long extStride;
    int extBcFromFields;      ///< DYN_EXTERNAL_FUSED: the moments of cells with boundary dynamics are stored in the fields (else implied)
    // Per-slot base pointers of the fast paths (every neighbour at its plain offset), computed on the host: cell c finds population
    // I at ldBase[I] + c and leaves f*_I at stBase[I] + c.  Even half: slot I / slot opp(I) of the cell itself; odd half: slot
    // opp(I) of cell c - c_I / slot I of cell c + c_I.  Two instructions per access (constant bank + one 64-bit add) instead of
    // the 64-bit multiply-add chains `I * S + offset` costs when S, P and nz are run-time values -- the sweep is short of
    // issue slots, not of bandwidth, whenever its path between two gathers grows (profiles/r02_ncu_metrics_planes.md).
    T* ldBase[27];
    T* stBase[27];
    PeerSlab<T> lo, hi;       ///< neighbouring slabs below X = 1 / above X = nxl that take direct stores
    int nRimTiles;            ///< > 0: tiles [0, nRimTiles) hold the rim planes of a face towards ANOTHER RANK; when the last of them
    unsigned int* rimTicket;  ///<      is done, sigValue (the step count) goes to the neighbours' flags: "my stores of this step have landed"
    unsigned int* sigLo;
    unsigned int* sigHi;
    unsigned int sigValue;
};

/// completePopulations of a plane boundary cell with (direction, orientation) turned into compile-time
/// constants, as the reference's templates <direction, orientation> have them
/// (WrappedLocalBoundaryFunctional3D::process boundaryCondition/wrappedLocalBoundaryProcessor3D.hh:42-62).
template <class T, int Q, bool INC>
__device__ __forceinline__ void complete_by_face_as(int baseType, Bc<T> b, T* f) {
    using L = LB<T, Q>;
    switch (b.dir * 2 + (b.ori > 0 ? 1 : 0)) {
    case 0: b.dir = 0; b.ori = -1; L::template complete_populations<INC>(baseType, b, f); break;
    case 1: b.dir = 0; b.ori = +1; L::template complete_populations<INC>(baseType, b, f); break;
    case 2: b.dir = 1; b.ori = -1; L::template complete_populations<INC>(baseType, b, f); break;
    case 3: b.dir = 1; b.ori = +1; L::template complete_populations<INC>(baseType, b, f); break;
    case 4: b.dir = 2; b.ori = -1; L::template complete_populations<INC>(baseType, b, f); break;
    default: b.dir = 2; b.ori = +1; L::template complete_populations<INC>(baseType, b, f); break;
    }
}
/// the concave edges and corners out of line: rare, and their loops over the composing planes are run-time
template <class T, int Q>
__device__ __noinline__ void complete_inner_cell(int baseType, const Bc<T>& b, T* f) {
    if (baseType == DYN_INCBGK) LB<T, Q>::template complete_inner<true>(baseType, b, f);
    else LB<T, Q>::template complete_inner<false>(baseType, b, f);
}
template <class T, int Q>
__device__ __forceinline__ void complete_by_face(int baseType, const Bc<T>& b, T* f) {
    if (LB<T, Q>::is_inner(b.kind)) { complete_inner_cell<T, Q>(baseType, b, f); return; }
    // IncBGK as base dynamics (velIsJ, invRho0) is its own instantiation: the usual one stays as lean as it was
    if (baseType == DYN_INCBGK) complete_by_face_as<T, Q, true>(baseType, b, f);
    else complete_by_face_as<T, Q, false>(baseType, b, f);
}

/// Dynamics::computeRhoBarJ of a cell with boundary dynamics from its (completed) populations and its record; out of line,
/// so that the caller's register file is not sized for this rarely taken path
template <class T, int Q>
__device__ __noinline__ void boundary_moments(const Dyn<T>* __restrict__ dyns, int k, const Bc<T>* __restrict__ bcs, int bi, const T* f, T* rhoBar_j) {
    T rb, j[3];
    LB<T, Q>::cell_rhoBar_j(dyns[k], true, bcs[bi], f, rb, j);
    rhoBar_j[0] = rb; rhoBar_j[1] = j[0]; rhoBar_j[2] = j[1]; rhoBar_j[3] = j[2];
}

}  // namespace plb200
