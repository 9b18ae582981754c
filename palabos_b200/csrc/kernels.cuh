// palabos_b200/csrc/kernels.cuh -- device kernels of the collide-and-stream path.
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

// Launch shape and cache policy of the sweep.  The defaults are the measured optimum on B200 for
// D3Q19 fp64 at 512^3 (profiles/r01_sweep_lab.log): 512-thread blocks for both halves, the odd
// (neighbour) half with an L2 prefetch one grid stride ahead; both use many more blocks than are
// resident (grid = SMs x factor), which staggers the load / compute / store phases of the resident blocks.  The macros exist for tools/sweep_lab.cu, which explores the alternatives.
#ifndef PLB_SWEEP_THREADS_EVEN
#define PLB_SWEEP_THREADS_EVEN 512
#endif
#ifndef PLB_SWEEP_THREADS_ODD
#define PLB_SWEEP_THREADS_ODD 512
#endif
#ifndef PLB_SWEEP_PF_EVEN       // L2 prefetch distance in grid strides (0 = off)
#define PLB_SWEEP_PF_EVEN 0
#endif
#ifndef PLB_SWEEP_PF_ODD
#define PLB_SWEEP_PF_ODD 1
#endif
#ifndef PLB_SWEEP_GRID_EVEN     // blocks per SM in the grid
#define PLB_SWEEP_GRID_EVEN 64
#endif
#ifndef PLB_SWEEP_GRID_ODD
#define PLB_SWEEP_GRID_ODD 128
#endif
#ifndef PLB_SWEEP_LD
#define PLB_SWEEP_LD(p) (*(p))
#endif
#ifndef PLB_SWEEP_ST
#define PLB_SWEEP_ST(p, v) (*(p) = (v))
#endif
__host__ __device__ constexpr int sweep_threads(bool odd) { return odd ? PLB_SWEEP_THREADS_ODD : PLB_SWEEP_THREADS_EVEN; }
#ifndef PLB_SWEEP_REGS          // register budget per thread that the launch bounds ask the compiler to keep
#define PLB_SWEEP_REGS 128
#endif
__host__ __device__ constexpr int sweep_minblocks(bool odd) {
    return 65536 / PLB_SWEEP_REGS / sweep_threads(odd) > 0 ? 65536 / PLB_SWEEP_REGS / sweep_threads(odd) : 1;
}
__host__ __device__ constexpr int sweep_prefetch(bool odd) { return odd ? PLB_SWEEP_PF_ODD : PLB_SWEEP_PF_EVEN; }
__host__ __device__ constexpr int sweep_grid_factor(bool odd) { return odd ? PLB_SWEEP_GRID_ODD : PLB_SWEEP_GRID_EVEN; }
constexpr int SWEEP_MAX_THREADS = 1024;

enum XMode : int { XMODE_EDGE = 0, XMODE_GHOST = 1, XMODE_WRAP = 2 };

/// Per-cell kind byte: bits 0-5 = index into the dynamics table, bit 7 = the cell carries a plane
/// boundary record whose completePopulations is fused into the sweep (see "lazy completion" below).
constexpr int KIND_MASK = 0x3f;
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
    const int* bcIdx;         ///< boundary record index per cell (read for KIND_PLANE cells only)
    const Bc<T>* bcs;         ///< boundary records
    int applyPlanes;          ///< the plane completion of the previous step is pending: apply it on load
    int pendingStamp;         ///< ... under the record values it was deferred with (complete_pending); -1: the current ones
    Geo g;
    TileRange range[3];       ///< the planes next to a neighbouring slab first (their stores have the longest way), then the rest
    int nRanges, nTiles;
    T omega, preFactor;       ///< parameters of dynamics-table entry 0 (the lattice's background dynamics)
    int gatherStats;          ///< BlockStatistics on/off (MultiBlock3D::toggleInternalStatistics)
    double* partials;         ///< [gridDim.x][4] per-block partial sums
    unsigned int* ticket;     ///< arrival counter for the last-block reduction
    double* statsOut;         ///< [4] sum rhoBar, sum uSqr, max uSqr, cell count of this launch
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
    // parking of sparse plane boundary cells (see sweep_kernel): what a warp still holds in its stash when its block ends goes
    // to the mailbox of (SM, warp) in global memory, where the next block on that SM picks it up
    int drain;                ///< this launch is the drain launch: block b flushes the mailboxes of SM b, nothing else
    int* mbCount;             ///< [MAILBOX_SMS * warps] entries waiting (-1: being written); nullptr: never park
    int* mbCell;              ///< [MAILBOX_SMS * warps][STASH_CAP]
    T* mbF;                   ///< [MAILBOX_SMS * warps][Q][STASH_CAP]
};

/// Deterministic block reduction of the BlockStatistics accumulators (core/blockStatistics.cpp:205-229:
/// gatherAverage / gatherMax / incrementStats, always in double) followed by a fixed-order reduction
/// over blocks done by whichever block arrives last.
__device__ __forceinline__ void reduce_stats(double sRho, double sU, double mU, double cnt,
                                             double* partials, unsigned int* ticket, double* out) {
    __shared__ double sh[4][SWEEP_MAX_THREADS / 32];
    __shared__ bool isLast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sRho += __shfl_down_sync(0xffffffffu, sRho, o);
        sU += __shfl_down_sync(0xffffffffu, sU, o);
        mU = fmax(mU, __shfl_down_sync(0xffffffffu, mU, o));
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { sh[0][warp] = sRho; sh[1][warp] = sU; sh[2][warp] = mU; sh[3][warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = sh[0][0], b = sh[1][0], m = sh[2][0], n = sh[3][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; m = fmax(m, sh[2][w]); n += sh[3][w]; }
        double* p = partials + 4 * (size_t)blockIdx.x;
        p[0] = a; p[1] = b; p[2] = m; p[3] = n;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    // fixed order: thread k sums blocks k, k+B, k+2B, ...; then the same tree as above
    double a = 0., b = 0., m = -DBL_MAX, n = 0.;
    for (unsigned int blk = threadIdx.x; blk < gridDim.x; blk += blockDim.x) {
        const volatile double* p = partials + 4 * (size_t)blk;
        a += p[0]; b += p[1]; m = fmax(m, p[2]); n += p[3];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
        n += __shfl_down_sync(0xffffffffu, n, o);
    }
    __syncthreads();
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; sh[2][warp] = m; sh[3][warp] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = sh[0][0]; b = sh[1][0]; m = sh[2][0]; n = sh[3][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; m = fmax(m, sh[2][w]); n += sh[3][w]; }
        out[0] = a; out[1] = b; out[2] = m; out[3] = n;
        *ticket = 0u;
    }
}

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
template <class T, int Q>
__device__ __forceinline__ void complete_by_face(int baseType, const Bc<T>& b, T* f) {
    // IncBGK as base dynamics (velIsJ, invRho0) is its own instantiation: the usual one stays as lean as it was
    if (baseType == DYN_INCBGK) complete_by_face_as<T, Q, true>(baseType, b, f);
    else complete_by_face_as<T, Q, false>(baseType, b, f);
}

/// Out-of-line form for the sweep and the on-the-fly readers: the populations travel through local
/// memory so that the caller's register file is not sized for this rarely taken path.
/// pendingStamp: number of the sweep that applies the outstanding completion (the caller itself, if it is that sweep): a record
/// changed since the completion was deferred carries this stamp and answers with its previous values; -1: current values
template <class T, int Q>
__device__ __forceinline__ void complete_pending_inline(int baseType, const Bc<T>* __restrict__ bcs, int bi, int pendingStamp, T* f) {
    Bc<T> b = bcs[bi];
    if (b.stamp == pendingStamp) { b.u[0] = b.uPrev[0]; b.u[1] = b.uPrev[1]; b.u[2] = b.uPrev[2]; b.rhoBar = b.rhoBarPrev; }
    complete_by_face<T, Q>(baseType, b, f);
}
template <class T, int Q>
__device__ __noinline__ void complete_pending(int baseType, const Bc<T>* __restrict__ bcs, int bi, int pendingStamp, T* f) {
    complete_pending_inline<T, Q>(baseType, bcs, bi, pendingStamp, f);
}

/// Dynamics::computeRhoBarJ of a cell with boundary dynamics from its (completed) populations and its record; out of line
/// like complete_pending, for the same reason
template <class T, int Q>
__device__ __noinline__ void boundary_moments(const Dyn<T>* __restrict__ dyns, int k, const Bc<T>* __restrict__ bcs, int bi, const T* f, T* rhoBar_j) {
    T rb, j[3];
    LB<T, Q>::cell_rhoBar_j(dyns[k], true, bcs[bi], f, rb, j);
    rhoBar_j[0] = rb; rhoBar_j[1] = j[0]; rhoBar_j[2] = j[1]; rhoBar_j[3] = j[2];
}

/// Stores of the rim cells into the neighbouring slab (PeerSlab).
///   even half: a cell writes f*_i into slot opp(i) of ITSELF; the neighbour's next (odd) half pulls the slots that point at
///              it from its ghost plane, so those go there as well -- plane X = 1 sends the f*_i with c_ix = -1 to the high
///              ghost plane of the slab below, plane X = nxl those with c_ix = +1 to the low ghost plane of the slab above.
///   odd half:  a cell writes f*_i into slot i of cell x + c_i; where that cell belongs to the neighbouring slab the store
///              goes to its plane nxl (below) or 1 (above) directly -- unless the cell does not exist (y / z rim of a
///              non-periodic lattice): then the on-site rule keeps the population at home, as everywhere else.
template <class T, int Q, int I = 1>
struct RimStore {
    template <int CX>
    __device__ __forceinline__ static void even(const PeerSlab<T>& p, int P, int r, const T* f) {
        using L = LB<T, Q>;
        if (L::c(I, 0) == CX) {
            const long long Xg = CX < 0 ? p.nxl + 1 : 0;
            p.F[(long long)L::opp(I) * p.S + (Xg * P + r)] = f[I];
        }
        RimStore<T, Q, I + 1>::template even<CX>(p, P, r, f);
    }
    __device__ __forceinline__ static void odd(T* __restrict__ F, const Geo& g, const Nbr& n, int c, int X,
                                               const PeerSlab<T>& lo, const PeerSlab<T>& hi, const T* f) {
        using L = LB<T, Q>;
        constexpr int sx = L::c(I, 0), sy = L::c(I, 1), sz = L::c(I, 2);
        if (sx < 0 && X == 1 && lo.F != nullptr && nbr_valid<sx, sy, sz>(n)) {
            lo.F[(long long)I * lo.S + ((long long)lo.nxl * g.P + (c - g.P) + nbr_off<0, sy, sz>(n))] = f[I];
        } else if (sx > 0 && X == g.nxl && hi.F != nullptr && nbr_valid<sx, sy, sz>(n)) {
            hi.F[(long long)I * hi.S + ((long long)g.P + (c - g.nxl * g.P) + nbr_off<0, sy, sz>(n))] = f[I];
        } else {
            { T* dst = F + stream_target<T, Q, I>(g, n, c); PLB_SWEEP_ST(dst, f[I]); }
        }
        RimStore<T, Q, I + 1>::odd(F, g, n, c, X, lo, hi, f);
    }
};
template <class T, int Q>
struct RimStore<T, Q, Q> {
    template <int CX>
    __device__ __forceinline__ static void even(const PeerSlab<T>&, int, int, const T*) {}
    __device__ __forceinline__ static void odd(T*, const Geo&, const Nbr&, int, int, const PeerSlab<T>&, const PeerSlab<T>&, const T*) {}
};

/// "Parking" of sparse plane boundary cells (lazy completion, see sweep_kernel): entries per warp, and the number of flagged
/// cells among a warp's 32 up to which they are parked rather than completed in line
constexpr int STASH_CAP = 32;
constexpr int STASH_MAX_NEW = 8;
constexpr int MAILBOX_SMS = 256;     ///< upper bound of %smid + 1
template <class T, int Q>
__host__ __device__ constexpr size_t sweep_stash_bytes(int threads) {
    return (size_t)(threads / 32) * STASH_CAP * (Q * sizeof(T) + sizeof(int));
}
__device__ __forceinline__ unsigned int sm_id() {
    unsigned int r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

/// The fused collide-and-stream sweep (BlockLattice3D::collideAndStream, atomicBlock/blockLattice3D.hh:253-312,
/// with Cell::collide -> Dynamics::collide inlined as a device functor and gatherStatistics,
/// core/latticeStatistics.h:36-45, reduced per block).
///   MAIN     dynamics type of table entry 0, compiled in
///   GENERIC  false: every cell is entry 0, BounceBack or NoDynamics; true: any table entry per cell
///   ODD      which half of the AA pattern (see file header)
///   PLANES   the lattice has plane boundary cells whose completion is fused into the sweep ("lazy
///            completion"): completePopulations of such a cell is a function of its own post-stream
///            populations and its record only, so instead of a pass over the face cells after the
///            step, the next sweep applies it to the populations it has just gathered.
///
/// Work distribution.  A block owns the tiles blockIdx.x, blockIdx.x + gridDim.x, ... (tile = 512 consecutive cells) and
/// hands them out to its WARPS, 32 consecutive cells ("piece") at a time, from a counter in shared memory.  Pieces differ in
/// cost -- row ends and rim planes take the general path of the odd half, boundary rows are completed in line, DRAM latency
/// varies -- and with a fixed warp <-> piece mapping the warps of a block drifted apart and waited for the slowest at the
/// end of every block with the SM nearly empty: 25-30 % of all warp cycles (ncu, profiles/r02_ncu_sweep_d.md).  Measured at
/// 512^3, boundary-free launches: 6.62 -> 6.38 ms (97.7 % of the measured HBM peak).
///
/// Lazy completion, lane-efficiently.  The x- and y-faces fill whole warps (z is the fastest index) and are completed in
/// line.  A z-face puts ONE plane cell into a warp, and the completion is ~2000 instructions: in line it would nearly
/// double the instruction count of the sweep.  A warp therefore PARKS such a cell -- its gathered populations go to the
/// warp's stash in shared memory and the cell is skipped -- and when the stash holds a warp's worth the parked cells are
/// completed, collided and scattered one per lane, as one more "piece" (the other warps keep streaming meanwhile: the
/// piece counter absorbs the delay).  What a warp still holds when its block runs out of pieces goes to the mailbox of
/// (SM, warp) in global memory, where the next block on that SM picks it up; a drain launch of 148 small blocks empties the
/// mailboxes after the sweep.  Nothing is read twice from the lattice and no list exists on the host.  The AA pattern allows
/// the delay: within a step every storage location is read and written on behalf of one cell only, whenever and by
/// whichever thread.
/// History (profiles/r01_ncu_full_summary_d.md, r02_bench_n1_b_stash_first.json, r02_bench_n1_d.json, 512^3): round 1 listed
/// these cells and their sector mates on the host for extra blocks of the same grid (+1.7 GB of DRAM reads per launch, step
/// 7.04 ms); flushing every stash at the end of its block left two warps busy for ~10 us with the SM empty, 64-128 times per
/// SM and launch (7.90 ms); deferring all of them to a second launch made that launch latency-bound (0.5 ms for 0.5 M cells).
template <class T, int Q, int MAIN, bool GENERIC, bool ODD, bool PLANES>
__global__ void __launch_bounds__(sweep_threads(ODD), sweep_minblocks(ODD)) sweep_kernel(const SweepArgs<T> a) {
    using L = LB<T, Q>;
    constexpr int SWEEP_THREADS = sweep_threads(ODD);
    constexpr int WARPS = SWEEP_THREADS / 32;
    extern __shared__ __align__(16) unsigned char sweepSmem[];
    __shared__ int pieceCounter;
    const Geo& g = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sRho = 0., sU = 0., mU = -DBL_MAX, cnt = 0.;
    T* stF = nullptr;       // [Q][STASH_CAP] populations of this warp's parked cells
    int* stC = nullptr;     // [STASH_CAP] their cell indices
    if (PLANES) {
        stF = reinterpret_cast<T*>(sweepSmem) + (size_t)warp * STASH_CAP * Q;
        stC = reinterpret_cast<int*>(reinterpret_cast<T*>(sweepSmem) + (size_t)WARPS * STASH_CAP * Q) + warp * STASH_CAP;
    }

    // outstanding completePopulations of a plane boundary cell (k = its kind byte)
    auto complete = [&](int c, int k, T* f) {
        const int bi = a.bcIdx[c];
        if (bi < 0) return;
        // in line, populations in registers: a warp busy with a completion is a warp that does not stream, and with 16 warps
        // per SM the sweep's bandwidth follows the number of streaming warps -- through local memory (the out-of-line form the
        // on-the-fly readers use) one completion held its warp for ~35 us (profiles/r02_ncu_metrics_planes.md)
        complete_pending_inline<T, Q>(a.dyns[k & KIND_MASK].type, a.bcs, bi, a.pendingStamp, f);
    };
    // Cell::collide, scatter, gatherStatistics
    auto collide_scatter = [&](int c, int k, bool interior, int X, const Nbr& n, T* f) {
        const int kRaw = k;
        (void)kRaw;
        k &= KIND_MASK;
        T rhoBar = (T)0, uSqr = (T)0;
        bool took, bounced = false;
        if constexpr (MAIN == DYN_EXTERNAL || MAIN == DYN_EXTERNAL_FUSED) {
            // ExternalRhoJcollideAndStream3D (basicDynamics/dynamicsProcessor3D.hh:41-76): rhoBar and j come from the fields.
            // DYN_EXTERNAL_FUSED: nobody has touched the fields since BoxRhoBarJfunctional3D filled them at the end of the
            // previous step, so they hold Dynamics::computeRhoBarJ of the very populations this thread has just gathered
            // (and, for a plane boundary cell, completed) -- recomputed here instead of stored and re-read: the moment
            // pass and 64 B per cell of traffic go away.  Cells flagged KIND_PLANE in this mode: boundary dynamics on top
            // (moments by the boundary formulas, from the record) or envelope layer (no record: zero for ever); when
            // the boundary records have changed since, their moments were stored under the old values (extBcFromFields).
            static_assert(GENERIC, "the external-moment sweep reads the dynamics table");
            const Dyn<T> d = a.dyns[k];
            T rb, j[3];
            if (MAIN == DYN_EXTERNAL || ((kRaw & KIND_PLANE) && a.extBcFromFields)) {
                rb = a.ext[c];
                j[0] = a.ext[a.extStride + c]; j[1] = a.ext[2 * a.extStride + c]; j[2] = a.ext[3 * a.extStride + c];
            } else if (kRaw & KIND_PLANE) {
                const int bi = a.bcIdx[c];
                if (bi < 0) {
                    rb = (T)0; j[0] = j[1] = j[2] = (T)0;
                } else {
                    T tmp[Q], m[4];
#pragma unroll
                    for (int i = 0; i < Q; ++i) tmp[i] = f[i];
                    boundary_moments<T, Q>(a.dyns, k, a.bcs, bi, tmp, m);
                    rb = m[0]; j[0] = m[1]; j[1] = m[2]; j[2] = m[3];
                }
            } else {
                Bc<T> none;
                none.kind = -1; none.dir = 0; none.ori = 1; none.u[0] = none.u[1] = none.u[2] = (T)0; none.rhoBar = (T)0;
                L::cell_rhoBar_j(d, false, none, f, rb, j);
            }
            L::collide_external(d, f, rb, j);
            took = false;
        } else if (!GENERIC) {
            if (k == 0) {
                took = L::template collide_as<MAIN>(f, a.omega, a.preFactor, rhoBar, uSqr);
            } else {
                took = false;
                if (a.dyns[k].type == DYN_BOUNCEBACK) {
                    if (interior || !ODD) bounced = true;       // the swap happens in the store below
                    else L::collide_bounceback(f);
                }
            }
        } else {
            const Dyn<T> d = a.dyns[k];
            took = L::collide_any(d, f, rhoBar, uSqr);
        }
        if (ODD && !interior) {
            PLB_SWEEP_ST(a.F + c, f[0]);
            RimStore<T, Q>::odd(a.F, g, n, c, X, a.lo, a.hi, f);
        } else {
            if (bounced) CellIO<T, Q>::store_base_swapped(a.stBase, c, f);
            else CellIO<T, Q>::store_base(a.stBase, c, f);
            if (!ODD) {
                const bool rimLo = a.lo.F != nullptr && c < 2 * g.P, rimHi = a.hi.F != nullptr && c >= g.nxl * g.P;
                if (rimLo || rimHi) {
                    if (bounced) L::collide_bounceback(f);      // what goes to the neighbouring slab is the swapped set
                    if (rimLo) RimStore<T, Q>::template even<-1>(a.lo, g.P, c - g.P, f);
                    if (rimHi) RimStore<T, Q>::template even<+1>(a.hi, g.P, c - g.nxl * g.P, f);
                }
            }
        }
        if (took) {
            sRho += (double)rhoBar;
            sU += (double)uSqr;
            mU = fmax(mU, (double)uSqr);
            cnt += 1.;
        }
    };

    const bool planes = PLANES && a.applyPlanes;
    const bool parking = PLANES && planes && a.mbCount != nullptr;
    if (threadIdx.x == 0) pieceCounter = 0;
    __syncthreads();
    const int nMyTiles = (!a.drain && (int)blockIdx.x < a.nTiles) ? (a.nTiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nPieces = nMyTiles * WARPS;
    auto next_piece = [&]() {
        int p = 0;
        if (lane == 0) p = atomicAdd(&pieceCounter, 1);
        return __shfl_sync(0xffffffffu, p, 0);
    };

    int nStash = 0;                 // warp-uniform
    if (PLANES && parking) {
        // what the previous block on this SM (drain launch: any block on SM blockIdx.x) left in this warp's mailbox
        const unsigned int home = a.drain ? blockIdx.x : sm_id();
        const int slot = (int)home * WARPS + warp;
        int n = 0;
        if (lane == 0 && home < (unsigned int)MAILBOX_SMS) {
            n = *(volatile int*)(a.mbCount + slot);
            if (n > 0 && atomicCAS(a.mbCount + slot, n, 0) != n) n = 0;
            if (n < 0) n = 0;
        }
        n = __shfl_sync(0xffffffffu, n, 0);
        if (n > 0) {
            __threadfence();
            if (lane < n) {
                stC[lane] = a.mbCell[(size_t)slot * STASH_CAP + lane];
#pragma unroll
                for (int i = 0; i < Q; ++i) stF[i * STASH_CAP + lane] = a.mbF[((size_t)slot * Q + i) * STASH_CAP + lane];
            }
            nStash = n;
            __syncwarp();
        }
    }

    // The loop.  One iteration is a piece, a flush of the stash, or a report about a rim piece.  The common case -- a piece
    // without plane boundary cells -- runs straight through gather / collide / scatter ("lean"): with 16 warps per SM the sweep
    // sits exactly at the point where every instruction between two gathers costs bandwidth (time follows the instruction
    // count of this path 1:1, profiles/r02_ncu_metrics_planes.md), so everything else is kept out of it.  The cells that
    // need completePopulations first -- a whole warp of an x- or y-face, or the parked ones -- share the second copy of
    // collide / scatter ("heavy").
    bool rimPending = false;        // this warp has worked on a piece of a rim tile and has not reported it yet
    bool flushNow = false;          // the mailbox could not take the leftovers: complete them here
    int p = next_piece();
    for (;;) {
        int c = 0, k = 0, X = 0, y = 0, z = 0;
        bool heavy = false;
        T f[Q];
        Nbr n;
        if (PLANES && nStash > 0 && (nStash > STASH_CAP - STASH_MAX_NEW || rimPending || flushNow || (p >= nPieces && a.drain))) {
            // ---- the parked cells, one per lane
            heavy = lane < nStash;
            if (heavy) {
                c = stC[lane];
#pragma unroll
                for (int i = 0; i < Q; ++i) f[i] = stF[i * STASH_CAP + lane];
                k = a.kind[c];
                if (ODD) { decode_cell(g, c, X, y, z); n = make_nbr(g, X, y, z); }
            }
            nStash = 0;
            __syncwarp();
        } else if (rimPending) {
            // ---- the stores of this warp's piece of a rim tile have been issued: perform them system-wide, count the piece in,
            // and let the last piece of the last rim tile tell the neighbouring ranks (they wait on the flag with a stream
            // memory operation before anything of theirs touches the planes this launch has stored into)
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                const unsigned int total = (unsigned int)a.nRimTiles * WARPS;
                if (atomicAdd(a.rimTicket, 1u) == total - 1u) {
                    __threadfence_system();
                    if (a.sigLo) *(volatile unsigned int*)a.sigLo = a.sigValue;
                    if (a.sigHi) *(volatile unsigned int*)a.sigHi = a.sigValue;
                    *a.rimTicket = 0u;
                }
            }
            rimPending = false;
            continue;
        } else if (p < nPieces) {
            // ---- a piece: tile and range (block-uniform per tile)
            const int tile = (int)blockIdx.x + (p / WARPS) * (int)gridDim.x;
            int rBegin = a.range[0].begin, rEnd = a.range[0].end, rTile0 = 0;
            if (a.nRanges > 1 && tile >= a.range[1].tile0) { rBegin = a.range[1].begin; rEnd = a.range[1].end; rTile0 = a.range[1].tile0; }
            if (a.nRanges > 2 && tile >= a.range[2].tile0) { rBegin = a.range[2].begin; rEnd = a.range[2].end; rTile0 = a.range[2].tile0; }
            const int pieceBegin = rBegin + (tile - rTile0) * SWEEP_THREADS + (p % WARPS) * 32;
            c = pieceBegin + lane;
            bool active = c < rEnd;
            bool interior = false;
            if (tile < a.nRimTiles) rimPending = true;
            if (ODD) {
                // decode the first cell of the piece; if its 32 cells lie in one row of interior y, clear of the row ends, in a
                // plane whose x-neighbours are planes of this slab (or ghost planes of the copy-based transport), every
                // neighbour of every cell exists at its plain offset and the whole warp takes the fast path
                decode_cell(g, pieceBegin, X, y, z);
                interior = z > 0 && z + 32 < g.nz && y > 0 && y < g.ny - 1
                        && (X > 1 || (g.lowMode == XMODE_GHOST && a.lo.F == nullptr))
                        && (X < g.nxl || (g.highMode == XMODE_GHOST && a.hi.F == nullptr));
                z += lane;
                if (z >= g.nz) decode_cell(g, c, X, y, z);
            }
            if (active) {
                if (ODD && !interior) {
                    n = make_nbr(g, X, y, z);
                    CellIO<T, Q>::load_pull(a.F, g, n, c, f);
                } else {
                    CellIO<T, Q>::load_base(a.ldBase, c, f);
                }
                if (sweep_prefetch(ODD) > 0) {
                    const long long delta = (long long)sweep_prefetch(ODD) * gridDim.x * SWEEP_THREADS;
                    if (c + delta < rEnd) {
                        if (ODD && !interior) CellPrefetch<T, Q, ODD>::run(a.F, g, n, c, (int)delta);
                        else CellIO<T, Q>::prefetch_base(a.stBase, c + (int)delta);
                    }
                }
                k = a.kind[c];
            }
            p = next_piece();
            if (PLANES && planes) {
                const bool flagged = active && (k & KIND_PLANE);
                const unsigned int m = __ballot_sync(0xffffffffu, flagged);
                if (m != 0u) {
                    if (parking && __popc(m) <= STASH_MAX_NEW) {
                        if (flagged) {
                            const int e = nStash + __popc(m & ((1u << lane) - 1u));
                            stC[e] = c;
#pragma unroll
                            for (int i = 0; i < Q; ++i) stF[i * STASH_CAP + e] = f[i];
                        }
                        nStash += __popc(m);
                        __syncwarp();
                    } else {
                        heavy = flagged;              // a boundary row: completed in line, by the heavy copy below
                        if (ODD && interior && heavy) n = make_nbr(g, X, y, z);     // (the heavy copy takes the general path)
                    }
                    active = active && !flagged;
                }
            }
            if (active) collide_scatter(c, k, interior, X, n, f);         // lean
        } else {
            if (PLANES && nStash > 0) {
                // ---- out of pieces with cells still parked: leave them to the next block on this SM
                const unsigned int home = sm_id();
                const int slot = (int)home * WARPS + warp;
                int ok = 0;
                if (lane == 0 && home < (unsigned int)MAILBOX_SMS) ok = atomicCAS(a.mbCount + slot, 0, -1) == 0 ? 1 : 0;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (!ok) { flushNow = true; continue; }          // occupied (cannot happen with one block per SM): complete here
                if (lane < nStash) {
                    a.mbCell[(size_t)slot * STASH_CAP + lane] = stC[lane];
#pragma unroll
                    for (int i = 0; i < Q; ++i) a.mbF[((size_t)slot * Q + i) * STASH_CAP + lane] = stF[i * STASH_CAP + lane];
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicExch(a.mbCount + slot, nStash);
                nStash = 0;
            }
            break;
        }
        if (PLANES && heavy) {
            complete(c, k, f);
            collide_scatter(c, k, false, X, n, f);
        }
    }
    if (a.gatherStats) reduce_stats(sRho, sU, mU, cnt, a.partials, a.ticket, a.statsOut);
}

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
    bool pending;        ///< plane completion is pending in storage: readers of KIND_PLANE cells apply it on the fly
    int pendingStamp;    ///< ... under the record values it was deferred with (complete_pending)
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
        if (cx.pending && (kb & KIND_PLANE) && v.hasBc) complete_pending<T, Q>(v.d.type, cx.bcs, bi, cx.pendingStamp, v.f);
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
}

/// One stage of the level-0 data processors (AtomicBlock3D::executeInternalProcessors,
/// atomicBlock/atomicBlock3D.cpp:195-222).  The host has ordered the (processor, cell) pairs into
/// stages such that running a stage's items concurrently equals the reference's sequential order.
/// Plane items (local completion of the populations) and finite-difference items (outer edges and
/// corners, which read neighbours) run in separate kernels.
///
/// WrappedLocalBoundaryFunctional3D::process boundaryCondition/wrappedLocalBoundaryProcessor3D.hh:42-62.
/// The switch on (direction, orientation) lets the compiler specialise completePopulations for each
/// of the six faces, as the reference's templates <direction, orientation> do.
template <class T, int Q>
__global__ void __launch_bounds__(128) plane_stage_kernel(CellCtx<T> cx, T* F, const Item* items, int nItems) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nItems) return;
    const int c = items[t].cell;
    const int bi = cx.bcIdx[c];
    if (bi < 0) return;
    Bc<T> b = cx.bcs[bi];
    // materialising a deferred completion: a record changed since answers with the values the completion was deferred under
    if (b.stamp == cx.pendingStamp) { b.u[0] = b.uPrev[0]; b.u[1] = b.uPrev[1]; b.u[2] = b.uPrev[2]; b.rhoBar = b.rhoBarPrev; }
    const int baseType = cx.dyns[cx.kind[c] & KIND_MASK].type;
    int X, y, z;
    decode_cell(cx.g, c, X, y, z);
    const Nbr n = make_nbr(cx.g, X, y, z);
    T f[Q];
    CellIO<T, Q>::load(F, cx.g, n, c, cx.swapped, f);
    complete_by_face<T, Q>(baseType, b, f);
    CellIO<T, Q>::store(F, cx.g, n, c, cx.swapped, f);
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
        if (cx.pending && (kb & KIND_PLANE) && hasBc) complete_pending<T, Q>(d.type, cx.bcs, bi, cx.pendingStamp, f);
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
/// device: the time loop of an unsteady inlet changes values every step, and neither a host walk over the records nor an eager
/// pass over the pending plane cells belongs there.  nextSweep: see Bc::stamp.
template <class T, int Q>
__global__ void set_bc_values_kernel(const int* __restrict__ bcIdx, Bc<T>* bcs, Geo g, BoxL box, int what, T u0, T u1, T u2, T rhoBar, int nextSweep) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= box.cells()) return;
    int X, y, z;
    const int c = box_cell(g, box, k, X, y, z);
    const int bi = bcIdx[c];
    if (bi < 0) return;
    Bc<T>& b = bcs[bi];
    const bool isRho = LB<T, Q>::is_rho_plane(b.kind);
    if ((what == 0) == isRho) return;
    if (b.stamp != nextSweep) {
        b.uPrev[0] = b.u[0]; b.uPrev[1] = b.u[1]; b.uPrev[2] = b.u[2]; b.rhoBarPrev = b.rhoBar;
        b.stamp = nextSweep;
    }
    if (what == 0) { b.u[0] = u0; b.u[1] = u1; b.u[2] = u2; }
    else b.rhoBar = rhoBar;
}

/// Mark the cells of the plane items whose completion is fused into the sweep.
static __global__ void flag_plane_cells_kernel(const Item* items, int n, uint8_t* kind) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) kind[items[t].cell] |= (uint8_t)KIND_PLANE;
}

/// Scatter the sparse (cell -> boundary record) map into the dense per-cell index.
static __global__ void scatter_bc_index_kernel(const int* cells, int n, int* bcIdx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = cells[t];
    if (c >= 0) bcIdx[c] = t;          // dead records (cell -1: defineDynamics replaced the composite) are not indexed
}

}  // namespace plb200
