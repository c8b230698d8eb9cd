// b2sv: multi-GPU plumbing (see comm.hpp). NCCL is resolved at run time with dlopen.
#include "comm.hpp"
#include "kernels.cuh"
#include "state.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
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
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
    LOAD(AllGather);
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
    double *d_token = nullptr; // one double used for stream-ordered barriers
    bool use_peer = true;      // NVLink peer-memory swap kernel (CUDA IPC); else NCCL send/recv
    uint64_t swaps = 0, swap_bytes = 0;
    // cross-rank barriers through IPC-mapped flag words (kernels.cu k_flag_barrier): kFlagChannels
    // independent channels of `world` words each, so barriers on different streams cannot mix up
    unsigned long long *flags = nullptr;
    std::vector<void *> flag_peers; // flag arrays of all ranks
    unsigned long long epoch[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
constexpr int kFlagChannels = 8;

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
    CUDA_CHECK(cudaMalloc(&c->d_token, sizeof(double) * 2));
    CUDA_CHECK(cudaMemset(c->d_token, 0, sizeof(double) * 2));
    if (const char *e = getenv("B2SV_SWAP"))
        c->use_peer = std::string(e) != "nccl";
    return c;
}
void comm_destroy(Comm *c) {
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->flags) {
        comm_unmap_peers(c, c->flag_peers);
        cudaFree(c->flags);
    }
    if (c->staging)
        cudaFree(c->staging);
    if (c->d_token)
        cudaFree(c->d_token);
    if (c->comm)
        nccl().CommDestroy(c->comm);
    delete c;
}
void comm_allreduce_sum(Comm *c, double *d_buf, int n, cudaStream_t stream) {
    NCCL_CHECK(nccl().AllReduce(d_buf, d_buf, static_cast<size_t>(n), kNcclFloat64, kNcclSum,
                                c->comm, stream));
}
void comm_barrier(Comm *c, cudaStream_t stream, int channel) {
    if (c->use_peer && c->flags && c->world <= 32) {
        B2_ASSERT(channel >= 0 && channel < kFlagChannels);
        FlagPeers fp{};
        for (int r = 0; r < c->world; r++)
            fp.p[r] = static_cast<unsigned long long *>(c->flag_peers[r]) + channel * c->world;
        launch_flag_barrier(c->flags + channel * c->world, fp, c->rank, c->world, ++c->epoch[channel],
                            stream);
        return;
    }
    NCCL_CHECK(nccl().AllReduce(c->d_token, c->d_token + 1, 1, kNcclFloat64, kNcclSum, c->comm,
                                stream));
}
// Collective, once per communicator: the flag words of every rank, IPC-mapped.
void comm_setup_flags(Comm *c, cudaStream_t stream) {
    if (c->flags || !c->use_peer)
        return;
    CUDA_CHECK(cudaSetDevice(c->device));
    // a 2 MiB allocation of its own: IPC handles map whole driver blocks, small cudaMallocs share one
    const size_t bytes = size_t(2) << 20;
    B2_ASSERT(sizeof(unsigned long long) * kFlagChannels * c->world <= bytes);
    CUDA_CHECK(cudaMalloc(&c->flags, bytes));
    CUDA_CHECK(cudaMemset(c->flags, 0, bytes));
    CUDA_CHECK(cudaDeviceSynchronize());
    comm_map_peers(c, c->flags, c->flag_peers, stream);
    if (!c->use_peer) { // mapping failed somewhere: every rank is on the NCCL path now
        cudaFree(c->flags);
        c->flags = nullptr;
        c->flag_peers.clear();
    }
}
int comm_rank(const Comm *c) { return c->rank; }
int comm_world(const Comm *c) { return c->world; }
void comm_stats(const Comm *c, uint64_t *swaps, uint64_t *bytes) {
    *swaps = c->swaps;
    *bytes = c->swap_bytes;
}
void comm_reset_stats(Comm *c) { c->swaps = c->swap_bytes = 0; }
void comm_count_exchange(Comm *c) { c->swaps++; }
bool comm_uses_peer(const Comm *c) { return c->use_peer; }

// Maps every rank's buffer into this process (CUDA IPC): out[r] = pointer usable in kernels here.
// Collective: every rank calls it with its own buffer, in the same order.
void comm_map_peers(Comm *c, void *my_buffer, std::vector<void *> &out, cudaStream_t stream) {
    out.assign(c->world, nullptr);
    out[c->rank] = my_buffer;
    if (!c->use_peer)
        return;
    CUDA_CHECK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t mine;
    cudaError_t e = cudaIpcGetMemHandle(&mine, my_buffer);
    char *d_all = nullptr;
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    CUDA_CHECK(cudaMalloc(&d_all, hs * (c->world + 1)));
    // slot `world` holds this rank's handle (all zeros on failure, detected by everybody)
    if (e != cudaSuccess) {
        cudaGetLastError();
        std::memset(&mine, 0, hs);
    }
    CUDA_CHECK(cudaMemcpyAsync(d_all + hs * c->world, &mine, hs, cudaMemcpyHostToDevice, stream));
    NCCL_CHECK(nccl().AllGather(d_all + hs * c->world, d_all, hs, kNcclInt8, c->comm, stream));
    std::vector<cudaIpcMemHandle_t> all(c->world);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all, hs * c->world, cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    CUDA_CHECK(cudaFree(d_all));
    bool ok = true;
    const cudaIpcMemHandle_t zero{};
    for (int r = 0; r < c->world; r++)
        ok = ok && std::memcmp(&all[r], &zero, hs) != 0;
    for (int r = 0; ok && r < c->world; r++) {
        if (r == c->rank)
            continue;
        e = cudaIpcOpenMemHandle(&out[r], all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ok = false;
        }
    }
    // all ranks must agree, otherwise one would wait in a barrier the other never enters
    double flag = ok ? 0.0 : 1.0, total = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(c->d_token, &flag, sizeof(double), cudaMemcpyHostToDevice, stream));
    NCCL_CHECK(nccl().AllReduce(c->d_token, c->d_token + 1, 1, kNcclFloat64, kNcclSum, c->comm, stream));
    CUDA_CHECK(cudaMemcpyAsync(&total, c->d_token + 1, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    flag = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(c->d_token, &flag, sizeof(double), cudaMemcpyHostToDevice, stream));
    if (total != 0.0) {
        comm_unmap_peers(c, out);
        out.assign(c->world, nullptr);
        out[c->rank] = my_buffer;
        c->use_peer = false; // every rank takes the NCCL send/recv path from now on
    }
}
void comm_unmap_peers(Comm *c, std::vector<void *> &ptrs) {
    for (int r = 0; r < static_cast<int>(ptrs.size()); r++)
        if (r != c->rank && ptrs[r]) {
            cudaIpcCloseMemHandle(ptrs[r]);
            ptrs[r] = nullptr;
        }
}

// One exchange of k rank bits with k local bits (all-to-all inside the 2^k-rank group), in place.
void comm_exchange(Comm *c, void *data, const std::vector<void *> &peers, int dtype, int n_local,
                   const std::vector<std::pair<int, int>> &jl, cudaStream_t st, int channel,
                   int max_ctas, bool fat, uint64_t slice_mask, uint64_t slice_value) {
    CUDA_CHECK(cudaSetDevice(c->device));
    const int k = static_cast<int>(jl.size());
    if (k == 0)
        return;
    const int n_slice = __builtin_popcountll(slice_mask);
    bool peer_ok = c->use_peer && k <= kMaxExchangeBits && n_slice <= 4 && n_local - k - 1 - n_slice >= 0;
    B2_ASSERT(peer_ok || slice_mask == 0);
    for (int r = 0; r < c->world && peer_ok; r++)
        peer_ok = peers[r] != nullptr;
    if (!peer_ok) { // NCCL path: one bit at a time through the staging buffer
        for (const auto &p : jl)
            comm_swap_bits(c, data, peers, dtype, n_local, p.first, p.second, st);
        return;
    }
    const size_t ab = dtype == 1 ? 16 : 8;
    if (channel == 0) // exchanges done slice by slice are counted by the caller (comm_count_exchange)
        c->swaps++;
    c->swap_bytes += ((uint64_t(1) << (n_local - n_slice)) - (uint64_t(1) << (n_local - n_slice - k))) * ab;
    ExchangeParams p{};
    p.k = k;
    p.n_local = n_local;
    uint64_t lmask = 0;
    for (int i = 0; i < k; i++) {
        p.lpos[i] = jl[i].second;
        lmask |= bit(jl[i].second);
        p.a |= static_cast<uint32_t>((c->rank >> jl[i].first) & 1) << i;
    }
    // selector: the highest local bit that is neither exchanged nor a slice bit
    p.selbit = n_local - 1;
    while (p.selbit >= 0 && ((lmask | slice_mask) & bit(p.selbit)))
        p.selbit--;
    B2_ASSERT(p.selbit >= 0 && (lmask & slice_mask) == 0);
    const uint64_t fix = lmask | bit(p.selbit) | slice_mask;
    for (int b = 0; b < n_local; b++)
        if (fix & bit(b))
            p.fixpos[p.nfix++] = b;
    p.nfree = n_local - p.nfix;
    const size_t sub_offset_bytes = static_cast<size_t>(slice_value) * ab; // slice bits as an index offset
    for (uint32_t b = 0; b < (1u << k); b++) {
        int r = c->rank;
        for (int i = 0; i < k; i++)
            r = (r & ~(1 << jl[i].first)) | (static_cast<int>((b >> i) & 1u) << jl[i].first);
        p.peer[b] = static_cast<char *>(peers[r]) + sub_offset_bytes;
    }
    comm_barrier(c, st, channel); // every rank has finished what it queued on its shard
    launch_exchange(dtype, static_cast<char *>(data) + sub_offset_bytes, p, max_ctas, fat, st);
    comm_barrier(c, st, channel); // every rank's exchange kernel has finished writing into this shard
}

// Exchange rank bit j with local bit l: amplitudes with (rank_bit, local_bit) = (0,1) on the lower
// rank trade places with (1,0) on the partner rank r ^ (1<<j).
//   peer path : one kernel per rank swaps its half of the pairs in place through the partner's
//               IPC-mapped shard (NVLink loads + stores), bracketed by stream-ordered barriers;
//   NCCL path : the half to send consists of 2^(n_local-1-l) contiguous runs of 2^l amplitudes and
//               goes through a bounded staging buffer in chunks (the shard may fill most of HBM).
void comm_swap_bits(Comm *c, void *data, const std::vector<void *> &peers, int dtype, int n_local,
                    int j, int l, cudaStream_t st) {
    CUDA_CHECK(cudaSetDevice(c->device));
    const int partner = c->rank ^ (1 << j);
    const int my_bit = (c->rank >> j) & 1;
    const size_t ab = dtype == 1 ? 16 : 8;
    c->swaps++;
    c->swap_bytes += (uint64_t(1) << (n_local - 1)) * ab;
    if (c->use_peer && peers[partner]) {
        comm_barrier(c, st, 0); // the partner has finished everything queued on its shard
        launch_peer_swap(dtype, data, peers[partner], n_local, l, my_bit, st);
        comm_barrier(c, st, 0); // the partner's kernel has finished writing into this shard
        return;
    }
    const uint64_t run = uint64_t(1) << l;                  // amplitudes per run
    const uint64_t nruns = uint64_t(1) << (n_local - 1 - l); // runs in the half
    const size_t want = size_t(256) << 20;
    const uint64_t chunk = std::min<uint64_t>(run, want / ab); // amplitudes per transfer
    if (c->staging_bytes < chunk * ab) {
        if (c->staging)
            CUDA_CHECK(cudaFree(c->staging));
        CUDA_CHECK(cudaMalloc(&c->staging, chunk * ab));
        c->staging_bytes = chunk * ab;
    }
    char *base = static_cast<char *>(data);
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

} // namespace b2sv
