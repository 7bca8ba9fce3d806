// Thin C-ABI wrappers over NCCL for a single-process host that drives several B200s itself (a C++ / Go / Java
// application binding libise directly): communicator over the visible devices, the per-iteration all-reduce of the
// k-means [k*d sums | k counts] buffer, and the all-gather of per-shard top-k lists.  The Python host uses
// torch.distributed (one process per GPU) instead -- same NCCL underneath.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") (the copy PyTorch ships is found when it is already loaded,
// otherwise the system one), so libise.so carries no link-time dependency on it.
#include <dlfcn.h>

#include <vector>

#include "common.cuh"

namespace {

typedef void* nccl_comm_t;
typedef int (*fn_CommInitAll)(nccl_comm_t*, int, const int*);
typedef int (*fn_CommDestroy)(nccl_comm_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char* (*fn_ErrStr)(int);

constexpr int kNcclInt8 = 0, kNcclFloat32 = 7, kNcclSum = 0;     // nccl.h: ncclDataType_t / ncclRedOp_t

struct NcclApi {
    void* handle = nullptr;
    fn_CommInitAll CommInitAll = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_Group GroupStart = nullptr, GroupEnd = nullptr;
    fn_ErrStr GetErrorString = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_LAZY | RTLD_LOCAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        a.CommInitAll = (fn_CommInitAll)dlsym(a.handle, "ncclCommInitAll");
        a.CommDestroy = (fn_CommDestroy)dlsym(a.handle, "ncclCommDestroy");
        a.AllReduce = (fn_AllReduce)dlsym(a.handle, "ncclAllReduce");
        a.AllGather = (fn_AllGather)dlsym(a.handle, "ncclAllGather");
        a.GroupStart = (fn_Group)dlsym(a.handle, "ncclGroupStart");
        a.GroupEnd = (fn_Group)dlsym(a.handle, "ncclGroupEnd");
        a.GetErrorString = (fn_ErrStr)dlsym(a.handle, "ncclGetErrorString");
        a.ok = a.CommInitAll && a.CommDestroy && a.AllReduce && a.AllGather && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}

}  // namespace

struct ise_comm {
    std::vector<int> devs;
    std::vector<nccl_comm_t> comms;
};

#define ISE_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != 0)                                                                                     \
            ISE_FAIL(std::string(#expr) + " -> " + (nccl().GetErrorString ? nccl().GetErrorString(_r) : "NCCL error")); \
    } while (0)

ISE_EXPORT int ise_comm_init_all(int ndev, const int* devs, ise_comm** out) {
    ISE_CHECK_ARG(out != nullptr && ndev >= 1);
    *out = nullptr;
    if (!nccl().ok) ISE_FAIL("libnccl.so.2 could not be loaded (dlopen): no collectives available");
    int visible = 0;
    ISE_CUDA(cudaGetDeviceCount(&visible));
    ISE_CHECK_ARG(ndev <= visible);
    ise_comm* c = new ise_comm();
    c->devs.resize(ndev);
    for (int i = 0; i < ndev; ++i) c->devs[i] = devs ? devs[i] : i;
    c->comms.resize(ndev, nullptr);
    int r = nccl().CommInitAll(c->comms.data(), ndev, c->devs.data());
    if (r != 0) {
        delete c;
        ISE_FAIL(std::string("ncclCommInitAll -> ") + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
    }
    *out = c;
    return 0;
}

ISE_EXPORT void ise_comm_destroy(ise_comm* comm) {
    if (!comm) return;
    for (nccl_comm_t c : comm->comms)
        if (c) nccl().CommDestroy(c);
    delete comm;
}

ISE_EXPORT int ise_comm_size(const ise_comm* comm) { return comm ? (int)comm->devs.size() : 0; }

// in place: bufs[i] (device pointer on device i of the communicator) <- sum over i of bufs[i]; streams[i] is that
// device's stream (cudaStream_t as void*, NULL entries / NULL array = default stream).  Asynchronous.
ISE_EXPORT int ise_allreduce_sum_f32(ise_comm* comm, float* const* bufs, int64_t count, void* const* streams) {
    ISE_CHECK_ARG(comm && bufs && count >= 0);
    if (count == 0) return 0;
    const int n = (int)comm->devs.size();
    ISE_NCCL(nccl().GroupStart());
    for (int i = 0; i < n; ++i) {
        int r = nccl().AllReduce(bufs[i], bufs[i], (size_t)count, kNcclFloat32, kNcclSum, comm->comms[i],
                                 (cudaStream_t)(streams ? streams[i] : nullptr));
        if (r != 0) {
            nccl().GroupEnd();
            ISE_FAIL(std::string("ncclAllReduce -> ") + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
        }
    }
    ISE_NCCL(nccl().GroupEnd());
    return 0;
}

// recv[i] (device i, n * bytes_per_rank bytes) <- concatenation over ranks of send[j] (bytes_per_rank bytes each):
// the exchange of per-shard (distance, id) lists before ise_topk_merge.  Asynchronous.
ISE_EXPORT int ise_allgather(ise_comm* comm, const void* const* send, void* const* recv, int64_t bytes_per_rank,
                             void* const* streams) {
    ISE_CHECK_ARG(comm && send && recv && bytes_per_rank >= 0);
    if (bytes_per_rank == 0) return 0;
    const int n = (int)comm->devs.size();
    ISE_NCCL(nccl().GroupStart());
    for (int i = 0; i < n; ++i) {
        int r = nccl().AllGather(send[i], recv[i], (size_t)bytes_per_rank, kNcclInt8, comm->comms[i],
                                 (cudaStream_t)(streams ? streams[i] : nullptr));
        if (r != 0) {
            nccl().GroupEnd();
            ISE_FAIL(std::string("ncclAllGather -> ") + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
        }
    }
    ISE_NCCL(nccl().GroupEnd());
    return 0;
}
