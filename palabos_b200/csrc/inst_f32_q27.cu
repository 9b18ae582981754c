// palabos_b200/csrc/inst_f32_q27.cu -- explicit instantiation of the engine for T = float, D3Q27
// (one translation unit per (T, Q) so that the four compile in parallel).
#include "engine_impl.cuh"

namespace plb200 {
template class LatticeImpl<float, 27>;
LatticeBase* make_lattice_f32_q27(const CreateDesc& d) { return new LatticeImpl<float, 27>(d); }
}  // namespace plb200
