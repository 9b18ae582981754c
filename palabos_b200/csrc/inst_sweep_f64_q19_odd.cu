// palabos_b200/csrc/inst_sweep_f64_q19_odd.cu -- the odd half of the sweep for T = double, D3Q19 (see sweep_launch.cuh)
#define PLB_SWEEP_DEFINE_LAUNCH
#include "sweep_launch.cuh"

namespace plb200 {
PLB_INSTANTIATE_SWEEPS(double, 19, true)
}  // namespace plb200
