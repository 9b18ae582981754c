// palabos_b200/csrc/sweep_launch.cuh -- the sweep kernels are compiled in translation units of their own
// (inst_sweep_*.cu: one per value type, velocity set and AA half), so that the build spreads over all cores;
// the engine only sees this declaration.
#pragma once
#include <cstdlib>
#include <string>
#include "engine.h"
#include "cell_io.cuh"

namespace plb200 {

/// sweep_kernel<T,Q,MAIN,GENERIC,ODD><<<grid, sweep_threads(ODD), 0, stream>>>(a)
template <class T, int Q, int MAIN, bool GENERIC, bool ODD>
void launch_sweep_kernel(int grid, cudaStream_t stream, const SweepArgs<T>& a);

#ifdef PLB_SWEEP_DEFINE_LAUNCH
}  // namespace plb200
#include "sweep_kernel.cuh"
namespace plb200 {
template <class T, int Q, int MAIN, bool GENERIC, bool ODD>
void launch_sweep_kernel(int grid, cudaStream_t stream, const SweepArgs<T>& a) {
    sweep_kernel<T, Q, MAIN, GENERIC, ODD><<<grid, sweep_threads(ODD), 0, stream>>>(a);
}
/// every combination the engine launches (LatticeImpl::launch_sweep): the five collision models compiled in as the
/// background dynamics, the generic form that reads the model from the dynamics table, and the external-moment sweep
/// (moments from the rhoBar / j fields, or recomputed from the gathered populations)
#define PLB_INSTANTIATE_SWEEPS(T, Q, ODD)                                                        \
    template void launch_sweep_kernel<T, Q, DYN_BGK, false, ODD>(int, cudaStream_t, const SweepArgs<T>&);    \
    template void launch_sweep_kernel<T, Q, DYN_RLB, false, ODD>(int, cudaStream_t, const SweepArgs<T>&);    \
    template void launch_sweep_kernel<T, Q, DYN_TRT, false, ODD>(int, cudaStream_t, const SweepArgs<T>&);    \
    template void launch_sweep_kernel<T, Q, DYN_SMAGO, false, ODD>(int, cudaStream_t, const SweepArgs<T>&);  \
    template void launch_sweep_kernel<T, Q, DYN_INCBGK, false, ODD>(int, cudaStream_t, const SweepArgs<T>&); \
    template void launch_sweep_kernel<T, Q, DYN_BGK, true, ODD>(int, cudaStream_t, const SweepArgs<T>&);     \
    template void launch_sweep_kernel<T, Q, DYN_EXTERNAL, true, ODD>(int, cudaStream_t, const SweepArgs<T>&);     \
    template void launch_sweep_kernel<T, Q, DYN_EXTERNAL_FUSED, true, ODD>(int, cudaStream_t, const SweepArgs<T>&);
#endif

}  // namespace plb200
