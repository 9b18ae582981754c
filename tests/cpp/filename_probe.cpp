// tests/cpp/filename_probe.cpp -- probe of plb::FileName / createFileName, compiled against the mirror header (tests) and, with -DUSE_REF, against the
// reference's own headers (to make tests/golden/filename_probe.txt)
#include <cstdio>
#include <string>
#ifdef USE_REF
#include "core/globalDefs.h"
#include "io/plbFiles.h"
#include "io/imageWriter.h"
#else
#include "palabos_b200/palabos3D.h"
#endif
using namespace plb;
static void show(FileName f) { std::printf("[%s|%s|%s] %s\n", f.getPath().c_str(), f.getName().c_str(), f.getExt().c_str(), f.get().c_str()); }
int main() {
    show(FileName("dir/sub/name.ext"));
    show(FileName("name"));
    show(FileName("name").defaultExt("plb"));
    show(FileName("name.dat").defaultExt("plb"));
    show(FileName("name").defaultPath("./"));
    show(FileName("a/name").defaultPath("./tmp/"));
    show(FileName("name.plb").setExt(".dat"));
    show(FileName("x/y.tar.gz").setExt(""));
    show(FileName("/abs/file.").setName("other"));
    show(FileName(".hidden"));
    std::printf("%s %s %s\n", createFileName("ckpt", 5, 8).c_str(), createFileName("u", 123456789, 6).c_str(), createFileName("out/vtk", 0, 1).c_str());
    return 0;
}
