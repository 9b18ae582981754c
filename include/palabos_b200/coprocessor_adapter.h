// include/palabos_b200/coprocessor_adapter.h -- plb::B200CoProcessor3D<T>: the reference's accelerator plug-in interface
// (plb::CoProcessor3D<T>, /root/reference/src/coProcessors/coProcessor3D.h:59-116) implemented on libplb_b200.so.
//
// What global::defaultCoProcessor3D<T>() (coProcessor3D.h:179-186) should return instead of the commented-out
// D3Q19CudaCoProcessor3D (coProcessor3D.hh:82-434).  Include AFTER plb::CoProcessor3D<T>, plb::Box3D and plb::plint are declared:
// either the reference's own headers (a Palabos build; examples/ref_coprocessor.cpp does exactly that with the unmodified
// sources) or the API mirror palabos3D.h of this repository.  See INTEGRATION.md section 1.
#ifndef PALABOS_B200_COPROCESSOR_ADAPTER_H
#define PALABOS_B200_COPROCESSOR_ADAPTER_H

#include <cstddef>
#include <vector>
#include "../plb_b200.h"

namespace plb {

/// Contract of the interface (coProcessor3D.h:56-58, :79-92): BGK on D3Q19, raw T, 19 populations fastest, then z, y, x; values as
/// Cell::serialize writes them (f_i - t_i).  Return convention 1 = success, 0 = failure, no exceptions.
template <typename T>
class B200CoProcessor3D : public CoProcessor3D<T> {
public:
    virtual int addDomain(plint nx, plint ny, plint nz, T omega, int& domainHandle) {
        return plb_b200_coproc_add_domain((int)sizeof(T), (int)nx, (int)ny, (int)nz, (double)omega, &domainHandle);
    }
    virtual int send(int domainHandle, Box3D const& subDomain, std::vector<char> const& data) {
        const int b[6] = {(int)subDomain.x0, (int)subDomain.x1, (int)subDomain.y0, (int)subDomain.y1, (int)subDomain.z0, (int)subDomain.z1};
        return plb_b200_coproc_send(domainHandle, b, data.empty() ? 0 : &data[0], data.size());
    }
    virtual int receive(int domainHandle, Box3D const& subDomain, std::vector<char>& data) const {
        const int b[6] = {(int)subDomain.x0, (int)subDomain.x1, (int)subDomain.y0, (int)subDomain.y1, (int)subDomain.z0, (int)subDomain.z1};
        // "it is the responsibility of the receive method to resize the data vector" (coProcessor3D.h:100-101); 19: the interface is D3Q19 only
        data.resize((std::size_t)subDomain.nCells() * 19 * sizeof(T));
        return plb_b200_coproc_receive(domainHandle, b, data.empty() ? 0 : &data[0], data.size());
    }
    virtual int collideAndStream(int domainHandle) { return plb_b200_coproc_collide_and_stream(domainHandle); }
};

}  // namespace plb
#endif
