// palabos_b200/csrc/capi.cu -- the extern "C" surface declared in include/plb_b200.h.
// Every entry point catches C++ exceptions and returns 1/0 like the reference's CoProcessor3D
// methods (coProcessors/coProcessor3D.h:77,93,103,115).
#include <cstdio>
#include <fcntl.h>
#include <unistd.h>
#include <algorithm>
#include <cstring>
#include <functional>
#include <vector>
#include <map>
#include <mutex>
#include <string>
#include "../../include/plb_b200.h"
#include "engine.h"
#include "nccl_dyn.h"

using namespace plb200;

/// desc holds the USER's lattice; off = width of the envelope layer in front of it inside the engine lattice
/// (plb_b200_create_enveloped; 0 otherwise).  Every box that crosses the ABI is shifted by it.
struct plb_b200_lattice { LatticeBase* impl; CreateDesc desc; int off[3]; int per[3]; bool enveloped; };

namespace {
struct SBox {
    int v[6];
    SBox(const plb_b200_lattice* L, const int b[6]) {
        if (!b) throw PlbError("null box");
        for (int k = 0; k < 6; ++k) v[k] = b[k] + L->off[k / 2];
        // a box must lie inside the user's lattice: the envelope layer is not addressable
        if (L->enveloped && !(b[0] >= 0 && b[1] < L->desc.nx && b[2] >= 0 && b[3] < L->desc.ny && b[4] >= 0 && b[5] < L->desc.nz))
            throw PlbError("box outside the lattice");
    }
    operator const int*() const { return v; }
};
}  // namespace

static thread_local std::string g_err;

static int fail(const std::string& m) { g_err = m; return 0; }
namespace plb200 { int set_last_error(const std::string& m) { return fail(m); } }

#define PLB_TRY(L, ...)                                            \
    if (!(L) || !(L)->impl) return fail("null lattice handle");    \
    try { __VA_ARGS__; return 1; }                                 \
    catch (const std::exception& e) { return fail(e.what()); }     \
    catch (...) { return fail("unknown error"); }

extern "C" {

const char* plb_b200_last_error(void) { return g_err.c_str(); }
int plb_b200_abi_version(void) { return PLB_B200_ABI_VERSION; }

plb_b200_lattice* plb_b200_create_distributed(int q, int prec, int nx, int ny, int nz, int dynamics, double p1, double p2,
                                              int rank, int nRanks, int slabsPerRank) {
    try {
        if ((q != 19 && q != 27) || (prec != 4 && prec != 8) || nx < 1 || ny < 1 || nz < 1)
            throw PlbError("bad lattice descriptor");
        CreateDesc d;
        d.q = q; d.prec = prec; d.nx = nx; d.ny = ny; d.nz = nz; d.dynamics = dynamics; d.p1 = p1; d.p2 = p2;
        d.rank = rank; d.nRanks = nRanks; d.slabsPerRank = slabsPerRank;
        LatticeBase* impl = prec == 8 ? (q == 19 ? make_lattice_f64_q19(d) : make_lattice_f64_q27(d))
                                      : (q == 19 ? make_lattice_f32_q19(d) : make_lattice_f32_q27(d));
        plb_b200_lattice* L = new plb_b200_lattice;
        L->impl = impl;
        L->desc = d;
        L->off[0] = L->off[1] = L->off[2] = 0;
        L->per[0] = L->per[1] = L->per[2] = 0;
        L->enveloped = false;
        return L;
    } catch (const std::exception& e) { fail(e.what()); return nullptr; }
    catch (...) { fail("unknown error"); return nullptr; }
}

plb_b200_lattice* plb_b200_create(int q, int prec, int nx, int ny, int nz, int dynamics, double p1, double p2) {
    return plb_b200_create_distributed(q, prec, nx, ny, nz, dynamics, p1, p2, 0, 1, 1);
}

plb_b200_lattice* plb_b200_create_enveloped(int q, int prec, int nx, int ny, int nz, int dynamics, double p1, double p2,
                                            int px, int py, int pz, int rank, int nRanks, int slabsPerRank) {
    const int per[3] = {px != 0, py != 0, pz != 0};
    const int off[3] = {per[0] ? 0 : 1, per[1] ? 0 : 1, per[2] ? 0 : 1};
    if (nx < 1 || ny < 1 || nz < 1) { fail("bad lattice descriptor"); return nullptr; }
    plb_b200_lattice* L = plb_b200_create_distributed(q, prec, nx + 2 * off[0], ny + 2 * off[1], nz + 2 * off[2],
                                                      dynamics, p1, p2, rank, nRanks, slabsPerRank);
    if (!L) return nullptr;
    try {
        L->desc.nx = nx; L->desc.ny = ny; L->desc.nz = nz;
        for (int k = 0; k < 3; ++k) { L->off[k] = off[k]; L->per[k] = per[k]; }
        L->enveloped = true;
        L->impl->set_periodicity(per[0], per[1], per[2]);
        const int bulk[6] = {0, nx - 1, 0, ny - 1, 0, nz - 1};
        L->impl->enable_external_rhoj(SBox(L, bulk));
        return L;
    } catch (const std::exception& e) { fail(e.what()); plb_b200_destroy(L); return nullptr; }
}

void plb_b200_destroy(plb_b200_lattice* L) {
    if (!L) return;
    try { delete L->impl; } catch (...) {}
    delete L;
}

int plb_b200_comm_unique_id(char id128[128]) {
    try {
        if (!id128) throw PlbError("null id buffer");
        NcclApi& n = NcclApi::get();
        ncclUniqueId uid;
        n.check(n.GetUniqueId(&uid), "ncclGetUniqueId");
        std::memcpy(id128, &uid, 128);
        return 1;
    } catch (const std::exception& e) { return fail(e.what()); }
}
int plb_b200_comm_init(plb_b200_lattice* L, const char id128[128]) {
    PLB_TRY(L, { if (!id128) throw PlbError("null id buffer"); L->impl->comm_init(id128); })
}

int plb_b200_slab_range(int nx, int nSlabs, int slab, int* x0, int* x1) {
    if (nx < 1 || nSlabs < 1 || nSlabs > nx || slab < 0 || slab >= nSlabs || !x0 || !x1) return fail("bad slab query");
    slab_range(nx, nSlabs, slab, *x0, *x1);
    return 1;
}

// ---- saveBinaryBlock / loadBinaryBlock (io/serializerIO_3D.cpp:37-66): the file is base64 (io/base64.hh: standard
// alphabet, '=' padding, no line breaks) of the data size as a 64-bit integer, followed by base64 of the block
// serializer's bytes in forward ordering -- for a lattice exactly the wire format of plb_b200_download.  Streamed
// through x-plane chunks so that a 20 GB lattice does not need a 20 GB host copy.
namespace {
const char B64[65] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

struct B64Encoder {
    FILE* fp; unsigned char carry[3]; int nCarry = 0; std::vector<char> out;
    explicit B64Encoder(FILE* f) : fp(f) {}
    void block(const unsigned char* d) {
        out.push_back(B64[d[0] >> 2]); out.push_back(B64[((d[0] & 0x03) << 4) | ((d[1] & 0xf0) >> 4)]);
        out.push_back(B64[((d[1] & 0x0f) << 2) | ((d[2] & 0xc0) >> 6)]); out.push_back(B64[d[2] & 0x3f]);
    }
    void write(const unsigned char* d, size_t n) {
        size_t pos = 0;
        out.clear(); out.reserve(n / 3 * 4 + 8);
        while (nCarry > 0 && nCarry < 3 && pos < n) carry[nCarry++] = d[pos++];
        if (nCarry == 3) { block(carry); nCarry = 0; }
        for (; pos + 3 <= n; pos += 3) block(d + pos);
        while (pos < n) carry[nCarry++] = d[pos++];
        if (!out.empty() && std::fwrite(out.data(), 1, out.size(), fp) != out.size()) throw PlbError("write error");
    }
    void finish() {
        if (nCarry == 0) return;
        for (int k = nCarry; k < 3; ++k) carry[k] = 0;
        char o[4] = {B64[carry[0] >> 2], B64[((carry[0] & 0x03) << 4) | ((carry[1] & 0xf0) >> 4)],
                     nCarry == 2 ? B64[((carry[1] & 0x0f) << 2) | ((carry[2] & 0xc0) >> 6)] : '=', '='};
        if (std::fwrite(o, 1, 4, fp) != 4) throw PlbError("write error");
        nCarry = 0;
    }
};

struct B64Decoder {
    FILE* fp; std::vector<unsigned char> pending;   // decoded bytes not yet handed out
    explicit B64Decoder(FILE* f) : fp(f) {}
    static int val(int c) {
        if (c >= 'A' && c <= 'Z') return c - 'A';
        if (c >= 'a' && c <= 'z') return c - 'a' + 26;
        if (c >= '0' && c <= '9') return c - '0' + 52;
        if (c == '+') return 62;
        if (c == '/') return 63;
        if (c == '=') return 0;
        return -1;
    }
    void read(unsigned char* dst, size_t n) {
        size_t have = std::min(n, pending.size());
        std::memcpy(dst, pending.data(), have);
        pending.erase(pending.begin(), pending.begin() + have);
        const size_t need = n - have;
        if (!need) return;
        const size_t groups = (need + 2) / 3;
        std::vector<char> in(groups * 4);
        size_t got = 0;
        while (got < in.size()) {                    // operator>> of the reference skips white space
            const int c = std::fgetc(fp);
            if (c == EOF) throw PlbError("binary block file ends early");
            if (c == ' ' || c == '\n' || c == '\r' || c == '\t') continue;
            in[got++] = (char)c;
        }
        std::vector<unsigned char> dec(groups * 3);
        for (size_t g = 0; g < groups; ++g) {
            int v[4];
            for (int k = 0; k < 4; ++k) { v[k] = val((unsigned char)in[4 * g + k]); if (v[k] < 0) throw PlbError("not a base64 character in binary block file"); }
            dec[3 * g] = (unsigned char)(v[0] << 2 | v[1] >> 4);
            dec[3 * g + 1] = (unsigned char)(v[1] << 4 | v[2] >> 2);
            dec[3 * g + 2] = (unsigned char)(((v[2] << 6) & 0xc0) | v[3]);
        }
        std::memcpy(dst + have, dec.data(), need);
        pending.assign(dec.begin() + need, dec.end());
    }
};

/// x-plane chunks of about 64 MB
void for_chunks(plb_b200_lattice* L, const std::function<void(const int box[6], size_t bytes)>& fn) {
    const CreateDesc& d = L->desc;
    if (d.nRanks != 1) throw PlbError("saveBinaryBlock / loadBinaryBlock: one rank only (the reference gathers on the main processor)");
    const size_t planeBytes = (size_t)d.ny * d.nz * d.q * d.prec;
    const int planes = (int)std::max<size_t>(1, (size_t(64) << 20) / planeBytes);
    for (int x = 0; x < d.nx; x += planes) {
        const int box[6] = {x, std::min(d.nx, x + planes) - 1, 0, d.ny - 1, 0, d.nz - 1};
        fn(box, (size_t)(box[1] - box[0] + 1) * planeBytes);
    }
}
}  // namespace

int plb_b200_save_binary_block(plb_b200_lattice* L, const char* path) {
    if (!L || !L->impl) return fail("null lattice handle");
    FILE* fp = nullptr;
    try {
        if (!path) throw PlbError("null path");
        fp = std::fopen(path, "wb");
        if (!fp) throw PlbError(std::string("Could not open binary file ") + path + " for saving");
        const CreateDesc& d = L->desc;
        const uint64_t total = (uint64_t)d.nx * d.ny * d.nz * d.q * d.prec;
        B64Encoder header(fp);
        header.write((const unsigned char*)&total, sizeof total);
        header.finish();
        B64Encoder enc(fp);
        std::vector<unsigned char> buf;
        for_chunks(L, [&](const int box[6], size_t bytes) {
            buf.resize(bytes);
            L->impl->download(SBox(L, box), buf.data(), bytes);
            enc.write(buf.data(), bytes);
        });
        enc.finish();
        if (std::fclose(fp) != 0) { fp = nullptr; throw PlbError("write error"); }
        return 1;
    } catch (const std::exception& e) { if (fp) std::fclose(fp); return fail(e.what()); }
}

int plb_b200_load_binary_block(plb_b200_lattice* L, const char* path) {
    if (!L || !L->impl) return fail("null lattice handle");
    FILE* fp = nullptr;
    try {
        if (!path) throw PlbError("null path");
        fp = std::fopen(path, "rb");
        if (!fp) throw PlbError(std::string("Could not open binary file ") + path + " for reading");
        const CreateDesc& d = L->desc;
        const uint64_t total = (uint64_t)d.nx * d.ny * d.nz * d.q * d.prec;
        uint64_t stored = 0;
        B64Decoder header(fp);
        header.read((unsigned char*)&stored, sizeof stored);
        if (stored != total) throw PlbError("binary block file holds " + std::to_string(stored) + " bytes, the lattice needs " + std::to_string(total));
        B64Decoder dec(fp);
        std::vector<unsigned char> buf;
        for_chunks(L, [&](const int box[6], size_t bytes) {
            buf.resize(bytes);
            dec.read(buf.data(), bytes);
            L->impl->upload(SBox(L, box), buf.data(), bytes);
        });
        std::fclose(fp);
        return 1;
    } catch (const std::exception& e) { if (fp) std::fclose(fp); return fail(e.what()); }
}

// ---- parallelIO::save / load (io/multiBlockWriter3D.cpp:226-238, io/multiBlockReader3D.cpp:251-259), static variables

namespace {
struct Component { int box[6]; long long begin, end; };      ///< a block of the file: its bulk and its byte range in the .dat

/// the blocks a lattice is saved as: its x-slabs (all ranks), clipped to the user's lattice
std::vector<Component> slab_components(const plb_b200_lattice* L) {
    const CreateDesc& d = L->desc;
    const int G = d.nRanks * d.slabsPerRank, implNx = d.nx + 2 * L->off[0];
    std::vector<Component> comps;
    long long pos = 0;
    for (int g = 0; g < G; ++g) {
        int x0, x1;
        slab_range(implNx, G, g, x0, x1);
        x0 = std::max(0, x0 - L->off[0]); x1 = std::min(d.nx - 1, x1 - L->off[0]);
        if (x1 < x0) continue;
        Component c;
        c.box[0] = x0; c.box[1] = x1; c.box[2] = 0; c.box[3] = d.ny - 1; c.box[4] = 0; c.box[5] = d.nz - 1;
        c.begin = pos;
        pos += (long long)(x1 - x0 + 1) * d.ny * d.nz * d.q * d.prec;
        c.end = pos;
        comps.push_back(c);
    }
    return comps;
}

/// text between <tag ...> and </tag>, starting the search at `from`; npos-safe
bool xml_field(const std::string& xml, const std::string& tag, size_t from, std::string& out, size_t* next = nullptr) {
    const size_t a = xml.find("<" + tag, from);
    if (a == std::string::npos) return false;
    const size_t b = xml.find('>', a);
    const size_t c = xml.find("</" + tag + ">", b);
    if (b == std::string::npos || c == std::string::npos) return false;
    out = xml.substr(b + 1, c - b - 1);
    if (next) *next = c + tag.size() + 3;
    return true;
}
std::vector<long long> xml_ints(const std::string& text) {
    std::vector<long long> v;
    const char* p = text.c_str();
    char* e = nullptr;
    for (;;) {
        const long long x = std::strtoll(p, &e, 10);
        if (e == p) break;
        v.push_back(x);
        p = e;
    }
    return v;
}
std::string trimmed(const std::string& t) {
    const size_t a = t.find_first_not_of(" \t\r\n"), b = t.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
}
}  // namespace

int plb_b200_parallel_save(plb_b200_lattice* L, const char* base_path, const char* data_file_entry, int global_id) {
    if (!L || !L->impl) return fail("null lattice handle");
    FILE* fp = nullptr;
    int fd = -1;
    try {
        if (!base_path) throw PlbError("null path");
        const CreateDesc& d = L->desc;
        const std::string base(base_path), dat = base + ".dat";
        const std::vector<Component> comps = slab_components(L);
        // every rank writes the bulks of its own slabs at their offsets (the reference: writeRawData, MPI-IO or gathered).  Nobody
        // truncates on opening -- the ranks arrive in any order --; rank 0 cuts the file to its final length at the end, which
        // removes the tail of an older, longer file and never touches what another rank has written
        fd = ::open(dat.c_str(), O_CREAT | O_WRONLY, 0644);
        if (fd < 0) throw PlbError("Could not open file " + dat + " for saving");
        const size_t planeBytes = (size_t)d.ny * d.nz * d.q * d.prec;
        const int planes = (int)std::max<size_t>(1, (size_t(64) << 20) / planeBytes);
        std::vector<unsigned char> buf;
        for (const Component& c : comps) {
            for (int x = c.box[0]; x <= c.box[1]; x += planes) {
                const int box[6] = {x, std::min(c.box[1], x + planes - 1), 0, d.ny - 1, 0, d.nz - 1};
                const size_t bytes = L->impl->local_bytes(SBox(L, box));
                if (bytes == 0) continue;                                  // another rank's slab
                if (bytes != (size_t)(box[1] - box[0] + 1) * planeBytes) throw PlbError("a slab is split over ranks");
                buf.resize(bytes);
                L->impl->download(SBox(L, box), buf.data(), bytes);
                off_t at = (off_t)(c.begin + (long long)(x - c.box[0]) * (long long)planeBytes);
                for (size_t done = 0; done < bytes;) {
                    const ssize_t w = ::pwrite(fd, buf.data() + done, bytes - done, at + (off_t)done);
                    if (w <= 0) throw PlbError("write error on " + dat);
                    done += (size_t)w;
                }
            }
        }
        if (d.rank == 0 && ::ftruncate(fd, (off_t)(comps.empty() ? 0 : comps.back().end)) != 0) throw PlbError("write error on " + dat);
        if (::close(fd) != 0) { fd = -1; throw PlbError("write error on " + dat); }
        fd = -1;
        if (d.rank == 0) {
            // writeXmlSpec io/multiBlockWriter3D.cpp:52-142, without the dynamics dictionary and the processor list: both are
            // optional to the reader (multiBlockReader3D.cpp:66-120) and describe the set-up, which lives in the program here
            const std::string plb = base + ".plb";
            fp = std::fopen(plb.c_str(), "wb");
            if (!fp) throw PlbError("Could not open file " + plb + " for saving");
            std::fprintf(fp, "<?xml version=\"1.0\" ?>\n<Block3D>\n    <General>\n        <Family> BlockLattice3D </Family>\n"
                             "        <Datatype> %s </Datatype>\n        <Descriptor> D3Q%d </Descriptor>\n        <cellDim> %d </cellDim>\n"
                             "        <dynamicContent> False </dynamicContent>\n        <globalId> %d </globalId>\n    </General>\n",
                         d.prec == 8 ? "double" : "float", d.q, d.q, global_id);
            std::fprintf(fp, "    <Structure>\n        <BoundingBox> 0 %d 0 %d 0 %d </BoundingBox>\n        <EnvelopeWidth> 1 </EnvelopeWidth>\n"
                             "        <GridLevel> 0 </GridLevel>\n        <NumComponents> %d </NumComponents>\n    </Structure>\n",
                         d.nx - 1, d.ny - 1, d.nz - 1, (int)comps.size());
            std::fprintf(fp, "    <Data>\n        <File> %s </File>\n", data_file_entry ? data_file_entry : dat.c_str());
            for (size_t k = 0; k < comps.size(); ++k) {
                const int* b = comps[k].box;
                if (comps.size() == 1) std::fprintf(fp, "        <Component> %d %d %d %d %d %d </Component>\n", b[0], b[1], b[2], b[3], b[4], b[5]);
                else std::fprintf(fp, "        <Component id=\"%d\"> %d %d %d %d %d %d </Component>\n", (int)k, b[0], b[1], b[2], b[3], b[4], b[5]);
            }
            std::fprintf(fp, "        <Offsets>");
            for (const Component& c : comps) std::fprintf(fp, " %lld", c.end);
            std::fprintf(fp, " </Offsets>\n    </Data>\n</Block3D>\n");
            if (std::fclose(fp) != 0) { fp = nullptr; throw PlbError("write error on " + plb); }
        }
        return 1;
    } catch (const std::exception& e) { if (fp) std::fclose(fp); if (fd >= 0) ::close(fd); return fail(e.what()); }
}

int plb_b200_parallel_load(plb_b200_lattice* L, const char* base_path) {
    if (!L || !L->impl) return fail("null lattice handle");
    FILE* fp = nullptr;
    try {
        if (!base_path) throw PlbError("null path");
        const CreateDesc& d = L->desc;
        const std::string base(base_path), plb = base + ".plb";
        fp = std::fopen(plb.c_str(), "rb");
        if (!fp) throw PlbError("Could not open file " + plb + " for reading");
        std::string xml;
        char chunk[4096];
        for (size_t n; (n = std::fread(chunk, 1, sizeof chunk, fp)) > 0;) xml.append(chunk, n);
        std::fclose(fp); fp = nullptr;
        // readXmlSpec io/multiBlockReader3D.cpp:122-207
        std::string t;
        if (!xml_field(xml, "Family", 0, t) || trimmed(t) != "BlockLattice3D") throw PlbError(plb + " does not hold a BlockLattice3D");
        if (!xml_field(xml, "Datatype", 0, t) || trimmed(t) != (d.prec == 8 ? "double" : "float")) throw PlbError(plb + ": data type does not match the lattice");
        if (xml_field(xml, "Descriptor", 0, t) && trimmed(t) != "D3Q" + std::to_string(d.q)) throw PlbError(plb + ": descriptor does not match the lattice");
        if (!xml_field(xml, "cellDim", 0, t) || xml_ints(t) != std::vector<long long>{d.q}) throw PlbError(plb + ": cellDim does not match the lattice");
        if (xml_field(xml, "dynamicContent", 0, t) && trimmed(t) != "False" && trimmed(t) != "false" && trimmed(t) != "0")
            throw PlbError(plb + " was saved with dynamic content (serialized dynamics objects): only static variables can be loaded");
        if (!xml_field(xml, "BoundingBox", 0, t) || xml_ints(t) != std::vector<long long>{0, d.nx - 1, 0, d.ny - 1, 0, d.nz - 1})
            throw PlbError(plb + ": bounding box does not match the lattice");
        if (xml_field(xml, "IndexOrdering", 0, t) && trimmed(t) != "zIsFastest") throw PlbError("Backward ordering not accepted while reading data.");
        if (!xml_field(xml, "Offsets", 0, t)) throw PlbError(plb + ": no offsets");
        const std::vector<long long> offsets = xml_ints(t);
        std::vector<Component> comps;
        size_t pos = 0;
        while (xml_field(xml, "Component", pos, t, &pos)) {
            const std::vector<long long> b = xml_ints(t);
            if (b.size() != 6) throw PlbError(plb + ": malformed component");
            Component c;
            for (int k = 0; k < 6; ++k) c.box[k] = (int)b[k];
            c.begin = comps.empty() ? 0 : comps.back().end;
            if (comps.size() >= offsets.size()) throw PlbError("Number of offsets does not match number of components in XML file.");
            c.end = offsets[comps.size()];
            const long long want = (long long)(c.box[1] - c.box[0] + 1) * (c.box[3] - c.box[2] + 1) * (c.box[5] - c.box[4] + 1) * d.q * d.prec;
            if (c.box[0] < 0 || c.box[1] >= d.nx || c.box[2] < 0 || c.box[3] >= d.ny || c.box[4] < 0 || c.box[5] >= d.nz || c.box[1] < c.box[0] ||
                c.box[3] < c.box[2] || c.box[5] < c.box[4] || c.end - c.begin != want) throw PlbError(plb + ": component does not fit the lattice");
            comps.push_back(c);
        }
        if (comps.size() != offsets.size() || comps.empty()) throw PlbError("Number of offsets does not match number of components in XML file.");
        // the data file: next to the .plb unless the entry carries a path of its own
        if (!xml_field(xml, "File", 0, t)) throw PlbError(plb + ": no data file");
        std::string dat = trimmed(t);
        if (dat.find('/') == std::string::npos) {
            const size_t slash = base.rfind('/');
            if (slash != std::string::npos) dat = base.substr(0, slash + 1) + dat;
        }
        fp = std::fopen(dat.c_str(), "rb");
        if (!fp) throw PlbError("Could not open file " + dat + " for reading");
        std::vector<unsigned char> buf;
        for (const Component& c : comps) {
            const size_t planeBytes = (size_t)(c.box[3] - c.box[2] + 1) * (c.box[5] - c.box[4] + 1) * d.q * d.prec;
            const int planes = (int)std::max<size_t>(1, (size_t(64) << 20) / planeBytes);
            for (int x = c.box[0]; x <= c.box[1]; ) {
                // planes [x, x1] of this component that lie on this rank (a contiguous run: slabs are x-ranges)
                int x1 = x;
                const int one[6] = {x, x, c.box[2], c.box[3], c.box[4], c.box[5]};
                const bool mine = L->impl->local_bytes(SBox(L, one)) != 0;
                while (x1 < c.box[1] && x1 - x + 1 < planes) {
                    const int nxt[6] = {x1 + 1, x1 + 1, c.box[2], c.box[3], c.box[4], c.box[5]};
                    if ((L->impl->local_bytes(SBox(L, nxt)) != 0) != mine) break;
                    ++x1;
                }
                if (mine) {
                    const int box[6] = {x, x1, c.box[2], c.box[3], c.box[4], c.box[5]};
                    const size_t bytes = (size_t)(x1 - x + 1) * planeBytes;
                    buf.resize(bytes);
                    if (fseeko(fp, (off_t)(c.begin + (long long)(x - c.box[0]) * (long long)planeBytes), SEEK_SET) != 0 ||
                        std::fread(buf.data(), 1, bytes, fp) != bytes) throw PlbError("read error on " + dat);
                    L->impl->upload(SBox(L, box), buf.data(), bytes);
                }
                x = x1 + 1;
            }
        }
        std::fclose(fp);
        return 1;
    } catch (const std::exception& e) { if (fp) std::fclose(fp); return fail(e.what()); }
}

int plb_b200_set_stream(plb_b200_lattice* L, void* s) { PLB_TRY(L, L->impl->set_stream((cudaStream_t)s)) }
int plb_b200_set_periodicity(plb_b200_lattice* L, int px, int py, int pz) {
    PLB_TRY(L, {
        if (L->enveloped && ((px != 0) != L->per[0] || (py != 0) != L->per[1] || (pz != 0) != L->per[2]))
            throw PlbError("an enveloped lattice keeps the periodicity it was created with");
        L->impl->set_periodicity(px, py, pz);
    })
}
int plb_b200_toggle_internal_statistics(plb_b200_lattice* L, int on) { PLB_TRY(L, L->impl->toggle_statistics(on != 0)) }
int plb_b200_define_dynamics_box(plb_b200_lattice* L, const int box[6], int dyn, double p1, double p2) {
    PLB_TRY(L, L->impl->define_dynamics_box(SBox(L, box), dyn, p1, p2))
}
int plb_b200_define_dynamics_mask(plb_b200_lattice* L, const int32_t* mask, int flag, int dyn, double p1, double p2) {
    PLB_TRY(L, {
        if (!L->enveloped) { L->impl->define_dynamics_mask(mask, flag, dyn, p1, p2); }
        else {
            // the mask of the user's lattice inside one of the engine lattice; the envelope layer never carries the flag
            if (!mask) throw PlbError("null mask");
            const CreateDesc& d = L->desc;
            const int NY = d.ny + 2 * L->off[1], NZ = d.nz + 2 * L->off[2], NX = d.nx + 2 * L->off[0];
            std::vector<int32_t> big((size_t)NX * NY * NZ, flag == 0 ? 1 : 0);
            for (int x = 0; x < d.nx; ++x) for (int y = 0; y < d.ny; ++y)
                std::memcpy(&big[((size_t)(x + L->off[0]) * NY + (y + L->off[1])) * NZ + L->off[2]],
                            &mask[((size_t)x * d.ny + y) * d.nz], (size_t)d.nz * sizeof(int32_t));
            L->impl->define_dynamics_mask(big.data(), flag, dyn, p1, p2);
        }
    })
}
int plb_b200_add_velocity_boundary(plb_b200_lattice* L, int model, int dir, int ori, const int box[6]) {
    PLB_TRY(L, L->impl->add_velocity_boundary(model, dir, ori, SBox(L, box)))
}
int plb_b200_add_pressure_boundary(plb_b200_lattice* L, int model, int dir, int ori, const int box[6]) {
    PLB_TRY(L, L->impl->add_pressure_boundary(model, dir, ori, SBox(L, box)))
}
int plb_b200_add_external_velocity_edge(plb_b200_lattice* L, int model, int plane, int n1, int n2, const int box[6]) {
    PLB_TRY(L, L->impl->add_external_velocity_edge(model, plane, n1, n2, SBox(L, box)))
}
int plb_b200_add_external_velocity_corner(plb_b200_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z) {
    PLB_TRY(L, L->impl->add_external_velocity_corner(model, xn, yn, zn, x + L->off[0], y + L->off[1], z + L->off[2]))
}
int plb_b200_add_internal_velocity_edge(plb_b200_lattice* L, int model, int plane, int n1, int n2, const int box[6]) {
    PLB_TRY(L, L->impl->add_internal_velocity_edge(model, plane, n1, n2, SBox(L, box)))
}
int plb_b200_add_internal_velocity_corner(plb_b200_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z) {
    PLB_TRY(L, L->impl->add_internal_velocity_corner(model, xn, yn, zn, x + L->off[0], y + L->off[1], z + L->off[2]))
}
int plb_b200_set_velocity_condition_on_block_boundaries(plb_b200_lattice* L, int model, const int block[6], const int app[6]) {
    PLB_TRY(L, L->impl->set_velocity_condition_on_block_boundaries(model, SBox(L, block), SBox(L, app)))
}
int plb_b200_set_boundary_velocity(plb_b200_lattice* L, const int box[6], const double u[3]) {
    PLB_TRY(L, L->impl->set_boundary_velocity(SBox(L, box), u))
}
int plb_b200_set_boundary_density(plb_b200_lattice* L, const int box[6], double rho) {
    PLB_TRY(L, L->impl->set_boundary_density(SBox(L, box), rho))
}
int plb_b200_initialize_at_equilibrium(plb_b200_lattice* L, const int box[6], double rho, const double u[3]) {
    PLB_TRY(L, L->impl->initialize_at_equilibrium(SBox(L, box), rho, u))
}
int plb_b200_initialize_at_equilibrium_field(plb_b200_lattice* L, const int box[6], const double* ru) {
    PLB_TRY(L, L->impl->initialize_at_equilibrium_field(SBox(L, box), ru))
}
int plb_b200_initialize(plb_b200_lattice* L) { PLB_TRY(L, L->impl->initialize()) }

int plb_b200_upload(plb_b200_lattice* L, const int box[6], const void* data, size_t bytes) {
    PLB_TRY(L, L->impl->upload(SBox(L, box), data, bytes))
}
int plb_b200_download(plb_b200_lattice* L, const int box[6], void* data, size_t bytes) {
    PLB_TRY(L, L->impl->download(SBox(L, box), data, bytes))
}
size_t plb_b200_local_bytes(plb_b200_lattice* L, const int box[6]) {
    if (!L || !L->impl) { fail("null lattice handle"); return 0; }
    try { return L->impl->local_bytes(SBox(L, box)); } catch (const std::exception& e) { fail(e.what()); return 0; }
}

int plb_b200_collide_and_stream(plb_b200_lattice* L, int n) { PLB_TRY(L, L->impl->collide_and_stream(n)) }
int plb_b200_execute_internal_processors(plb_b200_lattice* L, int n) { PLB_TRY(L, L->impl->execute_internal_processors(n)) }
int plb_b200_apply_rhobar_j(plb_b200_lattice* L, const int box[6]) { PLB_TRY(L, L->impl->apply_rhobar_j(SBox(L, box))) }
int plb_b200_upload_rhobar_j(plb_b200_lattice* L, const int box[6], const void* data, size_t bytes) {
    PLB_TRY(L, L->impl->transfer_rhobar_j(SBox(L, box), const_cast<void*>(data), bytes, true))
}
int plb_b200_download_rhobar_j(plb_b200_lattice* L, const int box[6], void* data, size_t bytes) {
    PLB_TRY(L, L->impl->transfer_rhobar_j(SBox(L, box), data, bytes, false))
}
int plb_b200_synchronize(plb_b200_lattice* L) { PLB_TRY(L, L->impl->synchronize()) }

int plb_b200_get_stored_statistics(plb_b200_lattice* L, double out3[3], uint64_t* n) {
    PLB_TRY(L, { if (!out3) throw PlbError("null output pointer"); L->impl->stored_statistics(out3, n); })
}
int plb_b200_compute_average_energy(plb_b200_lattice* L, const int box[6], double* out) {
    PLB_TRY(L, { if (!out) throw PlbError("null output pointer"); *out = L->impl->average_energy(SBox(L, box)); })
}
int plb_b200_compute_average_density(plb_b200_lattice* L, const int box[6], double* out) {
    PLB_TRY(L, { if (!out) throw PlbError("null output pointer"); *out = L->impl->average_density(SBox(L, box)); })
}
int plb_b200_compute_density_velocity(plb_b200_lattice* L, const int box[6], void* out, size_t bytes) {
    PLB_TRY(L, L->impl->density_velocity(SBox(L, box), out, bytes))
}

uint64_t plb_b200_launch_count(plb_b200_lattice* L) { return (L && L->impl) ? L->impl->launch_count() : 0; }
size_t plb_b200_device_bytes(plb_b200_lattice* L) { return (L && L->impl) ? L->impl->device_bytes() : 0; }
int plb_b200_boundary_stages(plb_b200_lattice* L, int* nStages, int64_t* nItems) {
    PLB_TRY(L, L->impl->boundary_stages(nStages, nItems))
}

int plb_b200_profile_sweeps(plb_b200_lattice* L, int on) { PLB_TRY(L, L->impl->profile_sweeps(on != 0)) }
int plb_b200_profile_read(plb_b200_lattice* L, double* totalMs, int* nLaunches) { PLB_TRY(L, L->impl->profile_read(totalMs, nLaunches)) }
const char* plb_b200_halo_transport(plb_b200_lattice* L) { return (L && L->impl) ? L->impl->halo_transport() : ""; }

// ---- plb::CoProcessor3D<T> surface (coProcessors/coProcessor3D.h:59-116) -------------------------

static std::mutex g_coprocMutex;
static std::map<int, plb_b200_lattice*> g_coproc;
static int g_nextHandle = 0;

static plb_b200_lattice* coproc(int handle) {
    std::lock_guard<std::mutex> lock(g_coprocMutex);
    auto it = g_coproc.find(handle);
    return it == g_coproc.end() ? nullptr : it->second;
}

int plb_b200_coproc_add_domain(int prec, int nx, int ny, int nz, double omega, int* handle) {
    if (!handle) return fail("null handle pointer");
    plb_b200_lattice* L = plb_b200_create(19, prec, nx, ny, nz, PLB_B200_DYN_BGK, omega, 0.);
    if (!L) return 0;
    // the co-processor contract has no statistics (the reference re-evaluates them on the host copy)
    L->impl->toggle_statistics(false);
    std::lock_guard<std::mutex> lock(g_coprocMutex);
    *handle = g_nextHandle++;
    g_coproc[*handle] = L;
    return 1;
}
int plb_b200_coproc_send(int handle, const int sub[6], const char* data, size_t bytes) {
    plb_b200_lattice* L = coproc(handle);
    if (!L) return fail("unknown domain handle");
    return plb_b200_upload(L, sub, data, bytes);
}
int plb_b200_coproc_receive(int handle, const int sub[6], char* data, size_t bytes) {
    plb_b200_lattice* L = coproc(handle);
    if (!L) return fail("unknown domain handle");
    return plb_b200_download(L, sub, data, bytes);
}
int plb_b200_coproc_collide_and_stream(int handle) {
    plb_b200_lattice* L = coproc(handle);
    if (!L) return fail("unknown domain handle");
    // "the collideAndStream() operation is blocking" (coProcessor3D.h:110-113)
    return plb_b200_collide_and_stream(L, 1) && plb_b200_synchronize(L);
}
int plb_b200_coproc_remove_domain(int handle) {
    plb_b200_lattice* L = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_coprocMutex);
        auto it = g_coproc.find(handle);
        if (it == g_coproc.end()) return fail("unknown domain handle");
        L = it->second;
        g_coproc.erase(it);
    }
    plb_b200_destroy(L);
    return 1;
}

}  // extern "C"
