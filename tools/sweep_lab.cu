// tools/sweep_lab.cu -- development tool (not part of the product): times the production sweep kernel
// of palabos_b200/csrc/sweep_kernel.cuh under different launch shapes / cache policies, selected at
// compile time with -DPLB_SWEEP_THREADS=.. -DPLB_SWEEP_MINBLOCKS=.. -DLAB_POLICY=..
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -Ipalabos_b200/csrc tools/sweep_lab.cu -o lab
//   ./lab <n> <iters> [gridFactor] [q27]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#ifndef LAB_POLICY
#define LAB_POLICY 0
#endif
#if LAB_POLICY == 1      // streaming loads and stores
#define PLB_SWEEP_LD(p) __ldcs(p)
#define PLB_SWEEP_ST(p, v) __stcs(p, v)
#elif LAB_POLICY == 2    // streaming stores only
#define PLB_SWEEP_ST(p, v) __stcs(p, v)
#elif LAB_POLICY == 3    // streaming loads only
#define PLB_SWEEP_LD(p) __ldcs(p)
#elif LAB_POLICY == 4    // last-use loads, write-through stores
#define PLB_SWEEP_LD(p) __ldlu(p)
#define PLB_SWEEP_ST(p, v) __stwt(p, v)
#elif LAB_POLICY == 5    // L1-bypassing loads
#define PLB_SWEEP_LD(p) __ldcg(p)
#define PLB_SWEEP_ST(p, v) __stcg(p, v)
#endif
#include "sweep_kernel.cuh"

using namespace plb200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

#ifndef LAB_T
#define LAB_T double
#endif
#ifndef LAB_Q
#define LAB_Q 19
#endif
#ifndef LAB_DYN
#define LAB_DYN DYN_BGK
#endif
typedef LAB_T T;
constexpr int Q = LAB_Q;

__global__ void fill(T* F, long long n, int q) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        F[k] = (T)(1e-3 * ((k * 2654435761u) % 1000) / 1000.0 - 5e-4);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 512;
    const int iters = argc > 2 ? atoi(argv[2]) : 10;
    const int gfE = argc > 3 ? atoi(argv[3]) : sweep_grid_factor<T, Q>(false);   // blocks per SM in the grid (even half)
    const int gfO = argc > 4 ? atoi(argv[4]) : sweep_grid_factor<T, Q>(true);    // (odd half)
    Geo g;
    g.nxl = n; g.ny = n; g.nz = n; g.P = n * n; g.S = (long long)(n + 2) * n * n;
    g.perY = 0; g.perZ = 0; g.lowMode = XMODE_EDGE; g.highMode = XMODE_EDGE;
    T* F; uint8_t* kind; Dyn<T>* dyns; double *partials, *stats; unsigned int* ticket;
    const long long nF = (long long)Q * g.S;
    CK(cudaMalloc(&F, nF * sizeof(T)));
    CK(cudaMalloc(&kind, g.S));
    CK(cudaMemset(kind, 0, g.S));
    CK(cudaMalloc(&dyns, 256 * sizeof(Dyn<T>)));
    CK(cudaMemset(dyns, 0, 256 * sizeof(Dyn<T>)));
    CK(cudaMalloc(&partials, 64 << 20));
    CK(cudaMalloc(&stats, 64));
    CK(cudaMalloc(&ticket, 4));
    CK(cudaMemset(ticket, 0, 4));
    fill<<<148 * 8, 256>>>(F, nF, Q);
    CK(cudaDeviceSynchronize());
    SweepArgs<T> a;
    a.F = F; a.kind = kind; a.dyns = dyns; a.bcIdx = nullptr; a.bcs = nullptr; a.g = g;
    a.lo.F = a.hi.F = nullptr; a.lo.S = a.hi.S = 0; a.lo.nxl = a.hi.nxl = 0; a.nRimTiles = 0; a.rimTicket = nullptr; a.sigLo = a.sigHi = nullptr; a.sigValue = 0;
    a.ext = nullptr; a.extStride = 0; a.extBcFromFields = 0;
    const int cellBegin = g.P, cellEnd = (n + 1) * g.P;
    for (int k = 0; k < 3; ++k) { a.range[k].begin = a.range[k].end = a.range[k].tile0 = 0; }
    a.range[0].begin = cellBegin; a.range[0].end = cellEnd; a.nRanges = 1;
    a.omega = (T)1.2388; a.preFactor = (T)0.01; a.gatherStats = 1; a.partials = partials; a.ticket = ticket; a.statsOut = stats;
    auto kEven = sweep_kernel<T, Q, LAB_DYN, false, false>;
    auto kOdd = sweep_kernel<T, Q, LAB_DYN, false, true>;
    int occE = 0, occO = 0, sms = 0;
    constexpr int TE = sweep_threads(false), TO = sweep_threads(true);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occE, kEven, TE, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occO, kOdd, TO, 0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaFuncAttributes fe, fo;
    CK(cudaFuncGetAttributes(&fe, kEven)); CK(cudaFuncGetAttributes(&fo, kOdd));
    auto gridOf = [&](int gf, int threads) { const int nTiles = (cellEnd - cellBegin + threads - 1) / threads; a.nTiles = nTiles; return gf * sms < nTiles ? gf * sms : nTiles; };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // per-slot base pointers of the fast paths, as LatticeImpl::launch_sweep sets them
    auto bases = [&](bool odd) {
        using L = LB<T, Q>;
        for (int i = 0; i < 27; ++i) a.ldBase[i] = a.stBase[i] = nullptr;
        for (int i = 0; i < Q; ++i) {
            const int o = L::opp(i);
            if (!odd) { a.ldBase[i] = F + (long long)i * g.S; a.stBase[i] = F + (long long)o * g.S; }
            else {
                const long long off = (long long)L::c(i, 0) * g.P + (long long)L::c(i, 1) * g.nz + L::c(i, 2);
                a.ldBase[i] = F + (long long)o * g.S - off; a.stBase[i] = F + (long long)i * g.S + off;
            }
        }
    };
    double msE = 0, msO = 0;
    for (int it = 0; it < iters + 2; ++it) {
        float t;
        CK(cudaEventRecord(e0));
        bases(false);
        kEven<<<gridOf(gfE, TE), TE>>>(a);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&t, e0, e1));
        if (it >= 2) msE += t;
        CK(cudaEventRecord(e0));
        bases(true);
        kOdd<<<gridOf(gfO, TO), TO>>>(a);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&t, e0, e1));
        if (it >= 2) msO += t;
    }
    CK(cudaGetLastError());
    msE /= iters; msO /= iters;
    const double bytes = (double)n * n * n * 2 * Q * sizeof(T);
    printf("T=%zu Q=%d dyn=%d n=%d policy=%d | even: thr %d pf %d gf %d regs %d occ %d | odd: thr %d pf %d gf %d regs %d occ %d | even %.3f ms %.0f GB/s | odd %.3f ms %.0f GB/s | avg %.0f GB/s %.0f MLUPS\n",
           sizeof(T), Q, (int)LAB_DYN, n, LAB_POLICY, TE, sweep_prefetch<T, Q>(false), gfE, fe.numRegs, occE, TO, sweep_prefetch<T, Q>(true), gfO, fo.numRegs, occO,
           msE, bytes / msE / 1e6, msO, bytes / msO / 1e6, 2 * bytes / (msE + msO) / 1e6, 2.0 * n * n * n / (msE + msO) / 1e3);
    return 0;
}
