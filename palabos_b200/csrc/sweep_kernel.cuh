// palabos_b200/csrc/sweep_kernel.cuh -- the collide-and-stream sweep kernel (compiled by the inst_sweep_*.cu translation units only;
// storage layout and addressing: cell_io.cuh).
#pragma once
#include "cell_io.cuh"

namespace plb200 {

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

/// The fused collide-and-stream sweep (BlockLattice3D::collideAndStream, atomicBlock/blockLattice3D.hh:253-312,
/// with Cell::collide -> Dynamics::collide inlined as a device functor and gatherStatistics,
/// core/latticeStatistics.h:36-45, reduced per block).
///   MAIN     dynamics type of table entry 0, compiled in
///   GENERIC  false: every cell is entry 0, BounceBack or NoDynamics; true: any table entry per cell
///   ODD      which half of the AA pattern (see cell_io.cuh)
///
/// Work distribution.  A block owns the tiles blockIdx.x, blockIdx.x + gridDim.x, ... (tile = 512 consecutive cells) and
/// hands them out to its WARPS, 32 consecutive cells ("piece") at a time, from a counter in shared memory.  Pieces differ in
/// cost -- row ends and rim planes take the general path of the odd half, DRAM latency varies -- and with a fixed
/// warp <-> piece mapping the warps of a block drifted apart and waited for the slowest at the end of every block with
/// the SM nearly empty: 25-30 % of all warp cycles (ncu, profiles/r02_ncu_sweep_d.md).  Measured at 512^3: 6.62 -> 6.38 ms
/// (97.7 % of the measured HBM peak).
///
/// The sweep does nothing but gather / collide / scatter.  With 16 warps per SM (128 registers x 512 threads) it sits exactly
/// at the point where every instruction between two gathers costs bandwidth -- time follows the instruction count of this
/// loop 1:1 (profiles/r02_ncu_metrics_planes.md) -- so the completion of the plane boundary cells, which rounds 1 and 2 had
/// fused into it in three different ways (lists of sparse cells: 7.04 ms per step at 512^3; per-warp stash in shared memory
/// flushed at block end: 7.90 ms; stash + per-SM mailbox + drain launch: 7.6 ms, with a 1 KB stack frame in the rare path),
/// is a pass of its own again (plane_stage_kernel, kernels.cuh).
template <class T, int Q, int MAIN, bool GENERIC, bool ODD>
__global__ void __launch_bounds__(sweep_threads(ODD), sweep_minblocks<T>(ODD)) sweep_kernel(const SweepArgs<T> a) {
    using L = LB<T, Q>;
    constexpr int SWEEP_THREADS = sweep_threads(ODD);
    constexpr int WARPS = SWEEP_THREADS / 32;
    __shared__ int pieceCounter;
    const Geo& g = a.g;
    const int lane = threadIdx.x & 31;
    double sRho = 0., sU = 0., mU = -DBL_MAX, cnt = 0.;

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
            // -- recomputed here instead of stored and re-read: the moment pass and 64 B per cell of traffic go away.
            // Cells flagged KIND_PLANE in this mode: boundary dynamics on top (moments by the boundary formulas, from
            // the record) or envelope layer (no record: zero for ever); when the boundary records have changed since,
            // their moments were stored under the old values (extBcFromFields).
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

    if (threadIdx.x == 0) pieceCounter = 0;
    __syncthreads();
    const int nMyTiles = ((int)blockIdx.x < a.nTiles) ? (a.nTiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nPieces = nMyTiles * WARPS;
    auto next_piece = [&]() {
        int p = 0;
        if (lane == 0) p = atomicAdd(&pieceCounter, 1);
        return __shfl_sync(0xffffffffu, p, 0);
    };

    bool rimPending = false;        // this warp has worked on a piece of a rim tile and has not reported it yet
    int p = next_piece();
    for (;;) {
        if (rimPending) {
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
        }
        if (p >= nPieces) break;
        // ---- a piece: tile and range (block-uniform per tile)
        const int tile = (int)blockIdx.x + (p / WARPS) * (int)gridDim.x;
        int rBegin = a.range[0].begin, rEnd = a.range[0].end, rTile0 = 0;
        if (a.nRanges > 1 && tile >= a.range[1].tile0) { rBegin = a.range[1].begin; rEnd = a.range[1].end; rTile0 = a.range[1].tile0; }
        if (a.nRanges > 2 && tile >= a.range[2].tile0) { rBegin = a.range[2].begin; rEnd = a.range[2].end; rTile0 = a.range[2].tile0; }
        const int pieceBegin = rBegin + (tile - rTile0) * SWEEP_THREADS + (p % WARPS) * 32;
        const int c = pieceBegin + lane;
        const bool active = c < rEnd;
        bool interior = false;
        int X = 0, y = 0, z = 0;
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
        T f[Q];
        Nbr n;
        int k = 0;
        if (active) {
            if (ODD && !interior) {
                n = make_nbr(g, X, y, z);
                CellIO<T, Q>::load_pull(a.F, g, n, c, f);
            } else {
                CellIO<T, Q>::load_base(a.ldBase, c, f);
            }
            if (sweep_prefetch<T, Q>(ODD) > 0) {
                const long long delta = (long long)sweep_prefetch<T, Q>(ODD) * gridDim.x * SWEEP_THREADS;
                if (c + delta < rEnd) {
                    if (ODD && !interior) CellPrefetch<T, Q, ODD>::run(a.F, g, n, c, (int)delta);
                    else CellIO<T, Q>::prefetch_base(a.stBase, c + (int)delta);
                }
            }
            k = a.kind[c];
            if ((k & KIND_HANDOFF) && a.hbuf != nullptr) {
                // a handoff cell: the plane pass has left its completed populations in the buffer (what was gathered is the raw set)
                const T* h = a.hbuf + a.bcIdx[c];
#pragma unroll
                for (int i = 0; i < Q; ++i) f[i] = h[(long long)i * a.hStride];
            }
        }
        p = next_piece();           // (the gather is in flight)
        if (active) collide_scatter(c, k, interior, X, n, f);
    }
    if (a.gatherStats) reduce_stats(sRho, sU, mU, cnt, a.partials, a.ticket, a.statsOut);
}

}  // namespace plb200
