// examples/coprocessor3d.cpp -- the reference's accelerator plug-in interface plb::CoProcessor3D<T>
// (/root/reference/src/coProcessors/coProcessor3D.h:59-116) driven the way the reference drives it
// (coProcessorCommunication3D.hh:36-124, multiBlockLattice3D.hh:326-329): addDomain for a pure BGK block, send the block's
// cells in Cell::serialize order, N x collideAndStream(handle), receive.  global::defaultCoProcessor3D<T>() is the singleton
// the reference's code calls; here it is plb::B200CoProcessor3D<T> (include/palabos_b200/coprocessor_adapter.h).
//
//   coprocessor3d <nx> <ny> <nz> <omega> <steps> <out.bin> [f32]
// writes the received populations ([nx][ny][nz][19] of T, f_i - t_i) to out.bin; tests/test_cpp_api.py compares them with
// the CPU oracle stepped from the same initial populations.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "palabos_b200/palabos3D.h"

using namespace plb;

template <typename T>
static int run(plint nx, plint ny, plint nz, double omega, int steps, const char* out) {
    CoProcessor3D<T>& co = global::defaultCoProcessor3D<T>();
    int handle = -1;
    if (!co.addDomain(nx, ny, nz, (T)omega, handle)) { std::fprintf(stderr, "addDomain: %s\n", plb_b200_last_error()); return 1; }
    const Box3D whole(0, nx - 1, 0, ny - 1, 0, nz - 1);
    // deterministic initial populations (the same formula in tests/test_cpp_api.py)
    std::vector<char> data((size_t)whole.nCells() * 19 * sizeof(T));
    T* f = reinterpret_cast<T*>(&data[0]);
    for (size_t k = 0; k < (size_t)whole.nCells() * 19; ++k) f[k] = (T)(1e-3 * std::sin(0.37 * (double)k) / 19.0);
    if (!co.send(handle, whole, data)) { std::fprintf(stderr, "send: %s\n", plb_b200_last_error()); return 1; }
    // a sub-domain send as well: the interface takes any box of the domain
    const Box3D sub(1, 2, 1, 3, 0, nz - 1);
    std::vector<char> part((size_t)sub.nCells() * 19 * sizeof(T));
    T* g = reinterpret_cast<T*>(&part[0]);
    for (size_t k = 0; k < (size_t)sub.nCells() * 19; ++k) g[k] = (T)(2e-3 * std::cos(0.11 * (double)k) / 19.0);
    if (!co.send(handle, sub, part)) { std::fprintf(stderr, "send(sub): %s\n", plb_b200_last_error()); return 1; }
    for (int s = 0; s < steps; ++s)
        if (!co.collideAndStream(handle)) { std::fprintf(stderr, "collideAndStream: %s\n", plb_b200_last_error()); return 1; }
    std::vector<char> back;                                    // "it is the responsibility of the receive method to resize"
    if (!co.receive(handle, whole, back)) { std::fprintf(stderr, "receive: %s\n", plb_b200_last_error()); return 1; }
    if (back.size() != data.size()) { std::fprintf(stderr, "receive resized to %zu bytes, expected %zu\n", back.size(), data.size()); return 1; }
    // error behaviour of the interface: 0, no exception
    std::vector<char> dummy;
    if (co.receive(handle + 1000, whole, dummy) != 0 || co.collideAndStream(handle + 1000) != 0) { std::fprintf(stderr, "unknown handle accepted\n"); return 1; }
    FILE* fp = std::fopen(out, "wb");
    if (!fp || std::fwrite(&back[0], 1, back.size(), fp) != back.size()) { std::fprintf(stderr, "cannot write %s\n", out); return 1; }
    std::fclose(fp);
    plb_b200_coproc_remove_domain(handle);
    std::printf("coprocessor3d: %ldx%ldx%ld, omega %.6f, %d steps, %zu bytes received\n", (long)nx, (long)ny, (long)nz, omega, steps, back.size());
    return 0;
}

int main(int argc, char* argv[]) {
    if (argc < 7) { std::fprintf(stderr, "usage: coprocessor3d nx ny nz omega steps out.bin [f32]\n"); return 2; }
    const plint nx = std::atol(argv[1]), ny = std::atol(argv[2]), nz = std::atol(argv[3]);
    const double omega = std::atof(argv[4]);
    const int steps = std::atoi(argv[5]);
    if (argc > 7 && std::strcmp(argv[7], "f32") == 0) return run<float>(nx, ny, nz, omega, steps, argv[6]);
    return run<double>(nx, ny, nz, omega, steps, argv[6]);
}
