// palabos_b200/csrc/inst_sweep_f32_q19_even.cu -- the even half of the sweep for T = float, D3Q19 (see sweep_launch.cuh)
#define PLB_SWEEP_DEFINE_LAUNCH
#include "sweep_launch.cuh"

namespace plb200 {
PLB_INSTANTIATE_SWEEPS(float, 19, false)
}  // namespace plb200
