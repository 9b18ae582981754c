// palabos_b200/csrc/sweep_launch.cuh -- the sweep kernels are compiled in translation units of their own
// (inst_sweep_*.cu: one per value type, velocity set and AA half), so that the build spreads over all cores;
// the engine only sees this declaration.
#pragma once
#include <cstdlib>
#include <string>
#include "engine.h"
#include "kernels.cuh"

namespace plb200 {

/// sweep_kernel<T,Q,MAIN,GENERIC,ODD,PLANES><<<grid, sweep_threads(ODD), 0, stream>>>(a)
template <class T, int Q, int MAIN, bool GENERIC, bool ODD, bool PLANES>
void launch_sweep_kernel(int grid, cudaStream_t stream, const SweepArgs<T>& a);

#ifdef PLB_SWEEP_DEFINE_LAUNCH
template <class T, int Q, int MAIN, bool GENERIC, bool ODD, bool PLANES>
void launch_sweep_kernel(int grid, cudaStream_t stream, const SweepArgs<T>& a) {
    // PLANES: the per-warp stash of parked plane boundary cells lives in dynamic shared memory (80 KB for D3Q19 fp64, 113 KB for
    // D3Q27 fp64 at 16 warps; the sweep runs one block per SM anyway -- 128 registers x 512 threads)
    size_t smem = PLANES ? sweep_stash_bytes<T, Q>(sweep_threads(ODD)) : 0;
    static const long dummy = std::getenv("PLB_B200_DUMMY_SMEM") ? std::atol(std::getenv("PLB_B200_DUMMY_SMEM")) : 0;   // tuning runs
    if (!PLANES && dummy > 0) smem = (size_t)dummy;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(sweep_kernel<T, Q, MAIN, GENERIC, ODD, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) throw PlbError(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): ") + cudaGetErrorString(e));
    }
    sweep_kernel<T, Q, MAIN, GENERIC, ODD, PLANES><<<grid, sweep_threads(ODD), smem, stream>>>(a);
}
/// every combination the engine launches (LatticeImpl::launch_sweep): the five collision models compiled in as the
/// background dynamics, and the generic form that reads the model from the dynamics table; with and without fused planes;
/// the external-moment sweep (moments from the rhoBar / j fields, plane boundaries completed eagerly)
#define PLB_INSTANTIATE_SWEEPS(T, Q, ODD)                                                        \
    template void launch_sweep_kernel<T, Q, DYN_BGK, false, ODD, false>(int, cudaStream_t, const SweepArgs<T>&);    \
    template void launch_sweep_kernel<T, Q, DYN_BGK, false, ODD, true>(int, cudaStream_t, const SweepArgs<T>&);     \
    template void launch_sweep_kernel<T, Q, DYN_RLB, false, ODD, false>(int, cudaStream_t, const SweepArgs<T>&);    \
    template void launch_sweep_kernel<T, Q, DYN_RLB, false, ODD, true>(int, cudaStream_t, const SweepArgs<T>&);     \
    template void launch_sweep_kernel<T, Q, DYN_TRT, false, ODD, false>(int, cudaStream_t, const SweepArgs<T>&);    \
    template void launch_sweep_kernel<T, Q, DYN_TRT, false, ODD, true>(int, cudaStream_t, const SweepArgs<T>&);     \
    template void launch_sweep_kernel<T, Q, DYN_SMAGO, false, ODD, false>(int, cudaStream_t, const SweepArgs<T>&);  \
    template void launch_sweep_kernel<T, Q, DYN_SMAGO, false, ODD, true>(int, cudaStream_t, const SweepArgs<T>&);   \
    template void launch_sweep_kernel<T, Q, DYN_INCBGK, false, ODD, false>(int, cudaStream_t, const SweepArgs<T>&); \
    template void launch_sweep_kernel<T, Q, DYN_INCBGK, false, ODD, true>(int, cudaStream_t, const SweepArgs<T>&);  \
    template void launch_sweep_kernel<T, Q, DYN_BGK, true, ODD, false>(int, cudaStream_t, const SweepArgs<T>&);     \
    template void launch_sweep_kernel<T, Q, DYN_BGK, true, ODD, true>(int, cudaStream_t, const SweepArgs<T>&);      \
    template void launch_sweep_kernel<T, Q, DYN_EXTERNAL, true, ODD, false>(int, cudaStream_t, const SweepArgs<T>&);     \
    template void launch_sweep_kernel<T, Q, DYN_EXTERNAL_FUSED, true, ODD, false>(int, cudaStream_t, const SweepArgs<T>&); \
    template void launch_sweep_kernel<T, Q, DYN_EXTERNAL_FUSED, true, ODD, true>(int, cudaStream_t, const SweepArgs<T>&);
#endif

}  // namespace plb200
