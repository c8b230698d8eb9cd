// b2sv: multi-GPU plumbing (see comm.hpp). NCCL is resolved at run time with dlopen.
#include "comm.hpp"
#include "state.hpp"

#include <algorithm>
#include <cstring>
#include <dlfcn.h>

namespace b2sv {

namespace {
// minimal NCCL ABI (nccl.h: ncclUniqueId is 128 opaque bytes; results are ints, 0 = success)
struct NcclId {
    char internal[128];
};
using ncclComm_t = void *;
enum { kNcclFloat64 = 8, kNcclInt8 = 0, kNcclSum = 0 };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi &nccl() {
    static NcclApi api;
    if (api.h)
        return api;
    // prefer a libnccl already mapped into the process (torch's bundled one), then the system one
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
        if (api.h)
            break;
    }
    for (const char *n : names) {
        if (api.h)
            break;
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    B2_ABORT_IF(!api.h, std::string("cannot load libnccl.so.2: ") + dlerror());
#define LOAD(sym)                                                                               \
    api.sym = reinterpret_cast<decltype(api.sym)>(dlsym(api.h, "nccl" #sym));                   \
    B2_ABORT_IF(!api.sym, "libnccl is missing nccl" #sym)
    LOAD(GetUniqueId);
    LOAD(CommInitRank);
    LOAD(CommDestroy);
    LOAD(AllReduce);
    LOAD(Send);
    LOAD(Recv);
    LOAD(GroupStart);
    LOAD(GroupEnd);
    LOAD(GetErrorString);
#undef LOAD
    return api;
}
#define NCCL_CHECK(call)                                                                        \
    do {                                                                                        \
        int r__ = (call);                                                                       \
        if (r__ != 0)                                                                           \
            B2_ABORT(std::string("NCCL error: ") + nccl().GetErrorString(r__) + " in " #call);  \
    } while (0)
} // namespace

struct Comm {
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    void *staging = nullptr;
    size_t staging_bytes = 0;
};

void comm_unique_id(void *out128) {
    NcclId id;
    NCCL_CHECK(nccl().GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
}

Comm *comm_create(int rank, int world, const void *unique_id, int device) {
    B2_ABORT_IF(!unique_id, "sharded state needs an NCCL unique id");
    auto *c = new Comm;
    c->rank = rank;
    c->world = world;
    c->device = device;
    NcclId id;
    std::memcpy(&id, unique_id, sizeof(id));
    CUDA_CHECK(cudaSetDevice(device));
    NCCL_CHECK(nccl().CommInitRank(&c->comm, world, id, rank));
    return c;
}
void comm_destroy(Comm *c) {
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->staging)
        cudaFree(c->staging);
    if (c->comm)
        nccl().CommDestroy(c->comm);
    delete c;
}
void comm_allreduce_sum(Comm *c, double *d_buf, int n, cudaStream_t stream) {
    NCCL_CHECK(nccl().AllReduce(d_buf, d_buf, static_cast<size_t>(n), kNcclFloat64, kNcclSum,
                                c->comm, stream));
}

// Exchange rank bit j with local index bit l: amplitudes with (rank_bit, local_bit) = (0,1) on the
// lower rank trade places with (1,0) on the partner rank r ^ (1<<j). The half to send consists of
// 2^(n_local-1-l) contiguous runs of 2^l amplitudes; it goes through a bounded staging buffer in
// chunks (the shard may fill most of HBM, so there is no room for a second copy).
static void swap_global_local(Comm *c, State &s, int j, int l) {
    CUDA_CHECK(cudaSetDevice(c->device));
    const int partner = c->rank ^ (1 << j);
    const int my_bit = (c->rank >> j) & 1;
    const size_t ab = s.amp_bytes();
    const uint64_t run = uint64_t(1) << l;                          // amplitudes per run
    const uint64_t nruns = uint64_t(1) << (s.num_local() - 1 - l);  // runs in the half
    const size_t want = size_t(256) << 20;
    const uint64_t chunk = std::min<uint64_t>(run, want / ab);      // amplitudes per transfer
    if (c->staging_bytes < chunk * ab) {
        if (c->staging)
            CUDA_CHECK(cudaFree(c->staging));
        CUDA_CHECK(cudaMalloc(&c->staging, chunk * ab));
        c->staging_bytes = chunk * ab;
    }
    char *base = static_cast<char *>(s.data());
    cudaStream_t st = s.stream();
    for (uint64_t r = 0; r < nruns; r++) {
        // run r of the half where local bit l == (1 - my_bit)
        const uint64_t start = (r << (l + 1)) | (uint64_t(1 - my_bit) << l);
        for (uint64_t off = 0; off < run; off += chunk) {
            char *p = base + (start + off) * ab;
            NCCL_CHECK(nccl().GroupStart());
            NCCL_CHECK(nccl().Send(p, chunk * ab, kNcclInt8, partner, c->comm, st));
            NCCL_CHECK(nccl().Recv(c->staging, chunk * ab, kNcclInt8, partner, c->comm, st));
            NCCL_CHECK(nccl().GroupEnd());
            CUDA_CHECK(cudaMemcpyAsync(p, c->staging, chunk * ab, cudaMemcpyDeviceToDevice, st));
        }
    }
}

void comm_localize(Comm *c, State &s, std::vector<Prim> &prims) {
    // v1 policy: find the rank bits that some primitive targets non-diagonally; bring each one in
    // by swapping it with a high local bit that no primitive of this batch targets (or, failing
    // that, any high local bit), run the batch with the relabelled bits, and swap back afterwards.
    const int nl = s.num_local();
    uint64_t need = 0, targeted = 0;
    for (const Prim &p : prims) {
        const uint64_t tm = p.target_mask();
        targeted |= tm;
        need |= tm >> nl;
    }
    if (!need)
        return;
    B2_ABORT("internal: global-qubit targets must be handled by State::apply_prims_sharded");
    (void)c;
    (void)swap_global_local;
}

} // namespace b2sv
