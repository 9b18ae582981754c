// palabos_b200/csrc/inst_f64_q19.cu -- explicit instantiation of the engine for T = double, D3Q19
// (one translation unit per (T, Q) so that the four compile in parallel).
#include "engine_impl.cuh"

namespace plb200 {
template class LatticeImpl<double, 19>;
LatticeBase* make_lattice_f64_q19(const CreateDesc& d) { return new LatticeImpl<double, 19>(d); }
}  // namespace plb200
