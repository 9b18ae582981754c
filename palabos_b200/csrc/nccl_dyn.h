// palabos_b200/csrc/nccl_dyn.h -- NCCL entry points resolved at run time (dlopen), so that the
// library has no link-time dependency on a particular libnccl: inside a Python process the copy
// bundled with torch is already mapped; a C++ host puts libnccl.so.2 on LD_LIBRARY_PATH or names
// it in PLB_B200_NCCL_LIB.  Only the calls the envelope exchange needs.
#pragma once
#include <dlfcn.h>
#include <cstdlib>
#include <string>
#include <nccl.h>
#include "engine.h"

namespace plb200 {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);

    static NcclApi& get() {
        static NcclApi api = load();
        return api;
    }

    void check(ncclResult_t r, const char* what) const {
        if (r != ncclSuccess) throw PlbError(std::string(what) + ": " + GetErrorString(r));
    }

private:
    template <class F>
    static void sym(void* h, const char* name, F& f) {
        f = reinterpret_cast<F>(dlsym(h, name));
        if (!f) throw PlbError(std::string("libnccl lacks ") + name);
    }
    static NcclApi load() {
        void* h = nullptr;
        if (const char* p = std::getenv("PLB_B200_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
#ifdef PLB_NCCL_FALLBACK_PATH
        if (!h) h = dlopen(PLB_NCCL_FALLBACK_PATH, RTLD_NOW | RTLD_GLOBAL);
#endif
        if (!h) throw PlbError("cannot load libnccl.so.2 (set PLB_B200_NCCL_LIB): " + std::string(dlerror()));
        NcclApi a;
        sym(h, "ncclGetUniqueId", a.GetUniqueId);
        sym(h, "ncclCommInitRank", a.CommInitRank);
        sym(h, "ncclCommDestroy", a.CommDestroy);
        sym(h, "ncclGroupStart", a.GroupStart);
        sym(h, "ncclGroupEnd", a.GroupEnd);
        sym(h, "ncclSend", a.Send);
        sym(h, "ncclRecv", a.Recv);
        sym(h, "ncclAllReduce", a.AllReduce);
        sym(h, "ncclAllGather", a.AllGather);
        sym(h, "ncclGetErrorString", a.GetErrorString);
        return a;
    }
};

}  // namespace plb200
