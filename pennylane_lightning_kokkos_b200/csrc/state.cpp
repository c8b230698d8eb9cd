// b2sv: State implementation (see state.hpp).
#include "state.hpp"

#include <mutex>
#include "comm.hpp"
#include "shard_plan.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace b2sv {

namespace {
void order_after(cudaStream_t waiter, cudaStream_t signaler) {
    if (waiter == signaler)
        return;
    cudaEvent_t ev;
    CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(ev, signaler));
    CUDA_CHECK(cudaStreamWaitEvent(waiter, ev, 0));
    CUDA_CHECK(cudaEventDestroy(ev));
}
int log2_exact(int x) {
    int g = 0;
    while ((1 << g) < x)
        g++;
    B2_ABORT_IF((1 << g) != x, "world size must be a power of two");
    return g;
}
// ---- recycling of device buffers of short-lived states -----------------------------------------------
// An adjoint Jacobian clones the state two or more times per call; cudaMalloc / cudaFree / cudaMallocHost
// of those clones cost milliseconds, comparable with the whole sweep on 20-26 qubit states. Freed state
// buffers of up to 2 GiB (and the small reduction buffers that go with every state) are therefore kept
// in a small per-process pool and handed to the next state of the same size on the same device.
struct SpareState {
    int device;
    size_t bytes;
    void *p;
};
struct SpareAux {
    int device;
    double *d_partials, *d_out, *h_out;
};
std::mutex g_pool_mu;
std::vector<SpareState> g_spare_states;
std::vector<SpareAux> g_spare_aux;
constexpr size_t kPoolMaxBytes = size_t(2) << 30;
constexpr size_t kPoolMaxStates = 6, kPoolMaxAux = 8;

void *pool_take_state(int device, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_spare_states.size(); i++)
        if (g_spare_states[i].device == device && g_spare_states[i].bytes == bytes) {
            void *p = g_spare_states[i].p;
            g_spare_states.erase(g_spare_states.begin() + i);
            return p;
        }
    return nullptr;
}
bool pool_give_state(int device, size_t bytes, void *p) { // caller: the buffer is idle
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (bytes > kPoolMaxBytes || g_spare_states.size() >= kPoolMaxStates)
        return false;
    g_spare_states.push_back({device, bytes, p});
    return true;
}
// gives every pooled state buffer back to the driver (called when an allocation fails)
void pool_release_states() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (const SpareState &sp : g_spare_states) {
        cudaSetDevice(sp.device);
        cudaFree(sp.p);
    }
    g_spare_states.clear();
    cudaSetDevice(cur);
}
bool pool_take_aux(int device, SpareAux *out) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_spare_aux.size(); i++)
        if (g_spare_aux[i].device == device) {
            *out = g_spare_aux[i];
            g_spare_aux.erase(g_spare_aux.begin() + i);
            return true;
        }
    return false;
}
bool pool_give_aux(const SpareAux &a) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_spare_aux.size() >= kPoolMaxAux)
        return false;
    g_spare_aux.push_back(a);
    return true;
}
} // namespace

CsrDevice::~CsrDevice() {
    cudaSetDevice(device);
    cudaFree(data);
    cudaFree(ind);
    cudaFree(ptr);
    cudaFree(ptr32);
}

std::shared_ptr<CsrDevice> csr_upload(int device, const cplx *data, const uint64_t *indices,
                                      const uint64_t *indptr, size_t nnz, size_t nrows) {
    CUDA_CHECK(cudaSetDevice(device));
    B2_ABORT_IF(nrows >= (uint64_t(1) << 32), "CSR matrices with 2^32 or more rows are not supported");
    B2_ABORT_IF(indptr[nrows] != nnz, "CSR indptr does not match the number of non-zeros");
    auto m = std::make_shared<CsrDevice>();
    m->device = device;
    m->nrows = nrows;
    m->nnz = nnz;
    std::vector<uint32_t> ind32(nnz);
    for (size_t i = 0; i < nnz; i++) {
        B2_ABORT_IF(indices[i] >= nrows, "CSR column index out of range");
        ind32[i] = static_cast<uint32_t>(indices[i]);
    }
    CUDA_CHECK(cudaMalloc(&m->data, sizeof(double2) * std::max<size_t>(nnz, 1)));
    CUDA_CHECK(cudaMalloc(&m->ind, sizeof(uint32_t) * std::max<size_t>(nnz, 1)));
    CUDA_CHECK(cudaMalloc(&m->ptr, sizeof(uint64_t) * (nrows + 1)));
    CUDA_CHECK(cudaMemcpy(m->data, data, sizeof(double2) * nnz, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(m->ind, ind32.data(), sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(m->ptr, indptr, sizeof(uint64_t) * (nrows + 1), cudaMemcpyHostToDevice));
    if (nnz < (uint64_t(1) << 32)) {
        std::vector<uint32_t> p32(nrows + 1);
        for (size_t i = 0; i <= nrows; i++) {
            B2_ABORT_IF(indptr[i] > nnz || (i > 0 && indptr[i] < indptr[i - 1]), "CSR indptr is not monotone");
            p32[i] = static_cast<uint32_t>(indptr[i]);
        }
        CUDA_CHECK(cudaMalloc(&m->ptr32, sizeof(uint32_t) * (nrows + 1)));
        CUDA_CHECK(cudaMemcpy(m->ptr32, p32.data(), sizeof(uint32_t) * (nrows + 1), cudaMemcpyHostToDevice));
    }
    const double avg = nrows ? double(nnz) / double(nrows) : 1.0;
    int L = 1;
    while (L < 32 && L < avg)
        L <<= 1;
    m->lanes = L;
    return m;
}

State::State(int num_qubits, int dtype, int device, int rank, int world, const void *nccl_id)
    : n_(num_qubits), dtype_(dtype), device_(device), rank_(rank), world_(world) {
    init_common(nccl_id);
}

State::State(const State &like, int share_stream)
    : n_(like.n_), dtype_(like.dtype_), device_(like.device_), rank_(like.rank_),
      world_(like.world_) {
    comm_ = like.comm_;
    if (comm_ || share_stream) { // every state of a communicator lives on one stream, so NCCL calls stay ordered
        stream_ = like.stream_;
        owns_stream_ = false;
    }
    fuse_ = like.fuse_;
    init_common(nullptr);
}

void State::init_common(const void *nccl_id) {
    B2_ABORT_IF(dtype_ != 0 && dtype_ != 1, "dtype must be B2SV_C64 (0) or B2SV_C128 (1)");
    B2_ABORT_IF(n_ < 1 || n_ > 60, "number of qubits out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    B2_ABORT_IF(e != cudaSuccess || ndev == 0,
                "no CUDA device available: b2sv has no CPU fallback (" +
                    std::string(cudaGetErrorString(e)) + ")");
    B2_ABORT_IF(device_ < 0 || device_ >= ndev, "device_id out of range");
    CUDA_CHECK(cudaSetDevice(device_));
    gbits_ = log2_exact(world_);
    B2_ABORT_IF(rank_ < 0 || rank_ >= world_, "rank out of range");
    B2_ABORT_IF(n_ - gbits_ < 1, "too few qubits for this many ranks");
    n_local_ = n_ - gbits_;
    tile_config(dtype_, &B_, &R_);
    n_eff_ = std::max(n_local_, B_);
    B2_ABORT_IF(world_ > 1 && n_local_ < B_ + gbits_,
                "sharded states need at least tile_bits + log2(world) local qubits");
    if (!stream_)
        CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    { // stream-ordered temporaries (cudaMallocAsync): keep up to 1 GiB cached instead of returning
      // everything to the driver at every synchronisation
        static std::mutex mu;
        static uint64_t configured_devices = 0;
        std::lock_guard<std::mutex> lk(mu);
        if (device_ < 64 && !((configured_devices >> device_) & 1)) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device_) == cudaSuccess) {
                uint64_t thr = uint64_t(1) << 30;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            configured_devices |= uint64_t(1) << device_;
        }
    }
    // sharded states are IPC-mapped by their peers: always a fresh cudaMalloc for those
    d_state_ = world_ > 1 ? nullptr : pool_take_state(device_, alloc_length() * amp_bytes());
    if (!d_state_) {
        e = cudaMalloc(&d_state_, alloc_length() * amp_bytes());
        if (e != cudaSuccess) { // the pooled buffers of smaller states may be what is in the way
            cudaGetLastError();
            pool_release_states();
            e = cudaMalloc(&d_state_, alloc_length() * amp_bytes());
        }
        B2_ABORT_IF(e != cudaSuccess, std::string("cannot allocate the state vector: ") +
                                          cudaGetErrorString(e));
    }
    SpareAux aux;
    if (pool_take_aux(device_, &aux)) {
        d_partials_ = aux.d_partials;
        d_out_ = aux.d_out;
        h_out_ = aux.h_out;
    } else {
        CUDA_CHECK(cudaMalloc(&d_partials_, sizeof(double) * kReduceBlocks * kMaxReduceVals));
        CUDA_CHECK(cudaMalloc(&d_out_, sizeof(double) * 64));
        CUDA_CHECK(cudaMallocHost(&h_out_, sizeof(double) * 64));
    }
    if (world_ > 1) {
        if (!comm_)
            comm_ = std::shared_ptr<Comm>(comm_create(rank_, world_, nccl_id, device_), comm_destroy);
        l2p_.resize(n_);
        comm_map_peers(comm_.get(), d_state_, peers_, stream_);
        comm_setup_flags(comm_.get(), stream_);
    }
    reset();
}

State::~State() {
    cudaSetDevice(device_);
    if (stream_)
        cudaStreamSynchronize(stream_);
    for (const TraceRec &r : trace_) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    if (trace_t0_)
        cudaEventDestroy(trace_t0_);
    if (comm_) {
        // Not collective (destruction order is up to the garbage collector and to exceptions, and
        // need not be the same on every rank): every exchange ends with a barrier that this rank's
        // stream has passed, so once the stream is idle no peer is still reading or writing this
        // shard, and peers only ever touch it inside an exchange all ranks take part in.
        comm_unmap_peers(comm_.get(), peers_);
    }
    comm_.reset();
    for (void *p : scratch_)
        if (!pool_give_state(device_, alloc_length() * amp_bytes(), p))
            cudaFree(p);
    // the stream was synchronised above: the buffers are idle and may serve the next state
    if (world_ > 1 || !d_state_ || !pool_give_state(device_, alloc_length() * amp_bytes(), d_state_))
        cudaFree(d_state_);
    if (!d_partials_ || !d_out_ || !h_out_ || !pool_give_aux({device_, d_partials_, d_out_, h_out_})) {
        cudaFree(d_partials_);
        cudaFree(d_out_);
        if (h_out_)
            cudaFreeHost(h_out_);
    }
    if (xstream_) {
        cudaStreamSynchronize(xstream_);
        cudaStreamDestroy(xstream_);
    }
    if (stream_ && owns_stream_)
        cudaStreamDestroy(stream_);
}

void State::sync() const {
    CUDA_CHECK(cudaSetDevice(device_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}

// ---- initialisation / copies ----------------------------------------------------------------------
void State::reset() { set_basis_state(0); }
void State::init_zeros() {
    CUDA_CHECK(cudaSetDevice(device_));
    reset_layout();
    touch();
    CUDA_CHECK(cudaMemsetAsync(d_state_, 0, alloc_length() * amp_bytes(), stream_));
}
void State::set_basis_state(uint64_t index) {
    CUDA_CHECK(cudaSetDevice(device_));
    reset_layout();
    touch();
    B2_ABORT_IF(n_ < 64 && index >= (uint64_t(1) << n_), "basis-state index out of range");
    const uint64_t owner = index >> n_local_;
    const uint64_t local = (owner == static_cast<uint64_t>(rank_)) ? (index & (local_length() - 1))
                                                                    : ~uint64_t(0);
    launch_set_basis(dtype_, d_state_, alloc_length(), local, stream_);
    launches++;
}
void State::set_state_vector(const uint64_t *indices, const cplx *values, size_t n) {
    CUDA_CHECK(cudaSetDevice(device_));
    init_zeros();
    std::vector<uint64_t> idx;
    std::vector<double2> val;
    for (size_t i = 0; i < n; i++) {
        B2_ABORT_IF(n_ < 64 && indices[i] >= (uint64_t(1) << n_), "state index out of range");
        if ((indices[i] >> n_local_) == static_cast<uint64_t>(rank_)) {
            idx.push_back(indices[i] & (local_length() - 1));
            val.push_back(make_double2(values[i].real(), values[i].imag()));
        }
    }
    if (idx.empty())
        return;
    uint64_t *d_idx;
    double2 *d_val;
    CUDA_CHECK(cudaMallocAsync(&d_idx, sizeof(uint64_t) * idx.size(), stream_));
    CUDA_CHECK(cudaMallocAsync(&d_val, sizeof(double2) * val.size(), stream_));
    CUDA_CHECK(cudaMemcpyAsync(d_idx, idx.data(), sizeof(uint64_t) * idx.size(),
                               cudaMemcpyHostToDevice, stream_));
    CUDA_CHECK(cudaMemcpyAsync(d_val, val.data(), sizeof(double2) * val.size(),
                               cudaMemcpyHostToDevice, stream_));
    touch();
    launch_scatter(dtype_, d_state_, d_idx, d_val, idx.size(), stream_);
    launches++;
    CUDA_CHECK(cudaFreeAsync(d_idx, stream_));
    CUDA_CHECK(cudaFreeAsync(d_val, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_)); // host vectors go out of scope
}
void State::set_state_on_wires(const std::vector<int64_t> &wires, const cplx *values) {
    CUDA_CHECK(cudaSetDevice(device_));
    const std::vector<int> bits = wires_to_bits(wires, n_);
    const int k = static_cast<int>(bits.size());
    B2_ABORT_IF(k < 1 || k > 40, "state preparation needs between 1 and 40 wires");
    init_zeros(); // identity layout: logical bit = physical bit
    const size_t count = size_t(1) << k;
    double2 *d_val;
    CUDA_CHECK(cudaMallocAsync(&d_val, sizeof(double2) * count, stream_));
    CUDA_CHECK(cudaMemcpyAsync(d_val, values, sizeof(double2) * count, cudaMemcpyHostToDevice, stream_));
    launch_scatter_wires(dtype_, d_state_, d_val, bits.data(), k, n_local_, static_cast<uint64_t>(rank_),
                         stream_);
    launches++;
    CUDA_CHECK(cudaFreeAsync(d_val, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_)); // the caller's buffer may go away
}
void State::h2d(const void *host, size_t length) {
    CUDA_CHECK(cudaSetDevice(device_));
    B2_ABORT_IF(length != local_length(), "HostToDevice: length does not match the state vector");
    reset_layout();
    touch();
    CUDA_CHECK(cudaMemcpyAsync(d_state_, host, length * amp_bytes(), cudaMemcpyHostToDevice, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}
void State::d2h(void *host, size_t length) const {
    CUDA_CHECK(cudaSetDevice(device_));
    B2_ABORT_IF(length != local_length(), "DeviceToHost: length does not match the state vector");
    normalize_layout();
    CUDA_CHECK(cudaMemcpyAsync(host, d_state_, length * amp_bytes(), cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}
void State::get_amplitudes(const uint64_t *indices, size_t n, cplx *out) const {
    CUDA_CHECK(cudaSetDevice(device_));
    if (n == 0)
        return;
    for (size_t i = 0; i < n; i++)
        B2_ABORT_IF(n_ < 64 && indices[i] >= (uint64_t(1) << n_), "state index out of range");
    normalize_layout();
    uint64_t *d_idx;
    double2 *d_val;
    CUDA_CHECK(cudaMallocAsync(&d_idx, sizeof(uint64_t) * n, stream_));
    CUDA_CHECK(cudaMallocAsync(&d_val, sizeof(double2) * n, stream_));
    CUDA_CHECK(cudaMemcpyAsync(d_idx, indices, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, stream_));
    launch_gather(dtype_, d_state_, d_idx, n, n_local_, static_cast<uint64_t>(rank_), d_val, stream_);
    if (comm_)
        comm_allreduce_sum(comm_.get(), reinterpret_cast<double *>(d_val), static_cast<int>(2 * n), stream_);
    CUDA_CHECK(cudaMemcpyAsync(out, d_val, sizeof(double2) * n, cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaFreeAsync(d_idx, stream_));
    CUDA_CHECK(cudaFreeAsync(d_val, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
}

// ---- tracing ---------------------------------------------------------------------------------------
State::TraceScope::TraceScope(const State &state, int k, cudaStream_t stream)
    : s(state), kind(k), st(stream ? stream : state.stream_) {
    if (!s.tracing_)
        return;
    cudaEventCreate(&e0);
    cudaEventRecord(e0, st);
}
State::TraceScope::~TraceScope() {
    if (!e0)
        return;
    cudaEvent_t e1;
    cudaEventCreate(&e1);
    cudaEventRecord(e1, st);
    s.trace_.push_back({kind, e0, e1});
}
void State::trace_begin() {
    CUDA_CHECK(cudaSetDevice(device_));
    for (const TraceRec &r : trace_) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    trace_.clear();
    if (!trace_t0_)
        CUDA_CHECK(cudaEventCreate(&trace_t0_));
    CUDA_CHECK(cudaEventRecord(trace_t0_, stream_));
    tracing_ = true;
}
std::vector<State::TraceOut> State::trace_end() {
    CUDA_CHECK(cudaSetDevice(device_));
    tracing_ = false;
    CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<TraceOut> out;
    for (const TraceRec &r : trace_) {
        float a = 0.f, d = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&a, trace_t0_, r.e0));
        CUDA_CHECK(cudaEventElapsedTime(&d, r.e0, r.e1));
        out.push_back({r.kind, a, d});
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    trace_.clear();
    return out;
}

void State::copy_from(const State &o) {
    CUDA_CHECK(cudaSetDevice(device_));
    B2_ABORT_IF(o.n_ != n_ || o.dtype_ != dtype_ || o.world_ != world_ || o.device_ != device_,
                "state vectors are not compatible");
    order_after(stream_, o.stream_);
    CUDA_CHECK(cudaMemcpyAsync(d_state_, o.d_state_, alloc_length() * amp_bytes(),
                               cudaMemcpyDeviceToDevice, stream_));
    order_after(o.stream_, stream_);
    l2p_ = o.l2p_;
    touch();
    bytes_moved += 2 * state_bytes();
}
std::unique_ptr<State> State::clone() const {
    // sharded: collective (every rank clones in the same order); shares communicator and stream
    auto c = std::make_unique<State>(*this, 0);
    c->copy_from(*this);
    return c;
}
std::unique_ptr<State> State::clone_on_stream() const {
    auto c = std::make_unique<State>(*this, 1);
    c->copy_from(*this);
    return c;
}
void State::swap_buffer(void *&other) {
    std::swap(d_state_, other);
    touch();
}
void *State::acquire_scratch() const {
    CUDA_CHECK(cudaSetDevice(device_));
    if (!scratch_.empty()) {
        void *p = scratch_.back();
        scratch_.pop_back();
        return p;
    }
    void *p = pool_take_state(device_, alloc_length() * amp_bytes());
    if (p)
        return p;
    cudaError_t e = cudaMalloc(&p, alloc_length() * amp_bytes());
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_release_states();
        e = cudaMalloc(&p, alloc_length() * amp_bytes());
    }
    B2_ABORT_IF(e != cudaSuccess, std::string("cannot allocate a scratch state vector: ") +
                                      cudaGetErrorString(e));
    return p;
}
void State::release_scratch(void *p) const {
    if (scratch_.size() < 2) {
        scratch_.push_back(p);
    } else {
        cudaStreamSynchronize(stream_);
        if (!pool_give_state(device_, alloc_length() * amp_bytes(), p))
            cudaFree(p);
    }
}

// ---- sharded layout -------------------------------------------------------------------------------
void State::reset_layout() const {
    for (size_t q = 0; q < l2p_.size(); q++)
        l2p_[q] = static_cast<int>(q);
}
uint64_t State::phys_mask(uint64_t m) const {
    if (l2p_.empty())
        return m;
    uint64_t r = 0;
    while (m) {
        const int q = __builtin_ctzll(m);
        r |= bit(l2p_[q]);
        m &= m - 1;
    }
    return r;
}
// One exchange: rank-bit positions <-> local positions, all pairs at once (comm.cpp comm_exchange).
void State::exchange_phys(const std::vector<std::pair<int, int>> &pairs) const {
    if (pairs.empty())
        return;
    B2_ASSERT(comm_);
    std::vector<std::pair<int, int>> jl;
    for (const auto &pr : pairs) {
        B2_ASSERT(pr.first >= n_local_ && pr.first < n_ && pr.second >= 0 && pr.second < n_local_);
        jl.emplace_back(pr.first - n_local_, pr.second);
    }
    {
        TraceScope ts(*this, 2);
        comm_exchange(comm_.get(), d_state_, peers_, dtype_, n_local_, jl, stream_, 0, exchange_ctas());
    }
    std::vector<int> p2l(n_);
    for (int q = 0; q < n_; q++)
        p2l[l2p_[q]] = q;
    for (const auto &pr : pairs)
        std::swap(l2p_[p2l[pr.first]], l2p_[p2l[pr.second]]);
}
int State::exchange_ctas() const {
    const char *e = getenv("B2SV_EXCHANGE_CTAS"); // read on every call: benchmarks sweep it
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : sm_count_current_device() * 8;
}
ShardPlanConfig State::shard_plan_config() const {
    ShardPlanConfig cfg;
    cfg.n = n_;
    cfg.n_local = n_local_;
    cfg.min_victim_pos = std::min(5, std::max(0, n_local_ - gbits_ - 1));
    if (const char *e = getenv("B2SV_EXCHANGE_BATCH"))
        cfg.batch = atoi(e) != 0;
    return cfg;
}
void State::ensure_local(uint64_t logical_mask) const {
    if (!comm_)
        return;
    CUDA_CHECK(cudaSetDevice(device_));
    std::vector<std::pair<int, int>> pairs;
    uint64_t used = 0; // local positions already chosen as victims
    for (int q = 0; q < n_; q++) {
        if (!((logical_mask >> q) & 1) || l2p_[q] < n_local_)
            continue;
        // victim: the highest local position whose logical owner is not wanted
        int victim = -1;
        for (int pos = n_local_ - 1; pos >= 0 && victim < 0; pos--) {
            if ((used >> pos) & 1)
                continue;
            for (int o = 0; o < n_; o++)
                if (l2p_[o] == pos && !((logical_mask >> o) & 1))
                    victim = pos;
        }
        B2_ABORT_IF(victim < 0, "operation acts on more qubits than one shard holds");
        used |= bit(victim);
        pairs.emplace_back(l2p_[q], victim);
    }
    exchange_phys(pairs);
}
void State::normalize_layout() const {
    if (!comm_)
        return;
    CUDA_CHECK(cudaSetDevice(device_));
    // rank bits first: logical bit P must sit at physical position P for P >= n_local
    {
        std::vector<int> l2p = l2p_;
        for (const ShardStep &stp : plan_normalize(l2p, shard_plan_config()))
            exchange_phys(stp.swaps);
        B2_ASSERT(l2p == l2p_);
    }
    // then sort the local positions with SWAPs (free address-map permutations inside tile passes)
    std::vector<Prim> prims;
    std::vector<int> cur(l2p_.begin(), l2p_.end());
    for (int q = 0; q < n_local_; q++) {
        if (cur[q] == q)
            continue;
        int other = -1;
        for (int o = 0; o < n_local_; o++)
            if (cur[o] == q)
                other = o;
        // exchange the contents of physical positions cur[q] and q
        const std::vector<int> bits = {cur[q], q};
        lower_gate("SWAP", bits, false, {}, prims);
        std::swap(cur[q], cur[other]);
    }
    if (!prims.empty())
        const_cast<State *>(this)->run_local(prims);
    reset_layout();
}
void State::comm_stats(uint64_t *swaps, uint64_t *bytes, int *peer) const {
    *swaps = *bytes = 0;
    *peer = 0;
    if (comm_) {
        b2sv::comm_stats(comm_.get(), swaps, bytes);
        *peer = comm_uses_peer(comm_.get()) ? 1 : 0;
    }
}

// Sharded gate application (planner: shard_plan.cpp): runs of primitives whose non-diagonal targets are
// shard-local, separated by exchanges that bring every global qubit with pending work into the shard
// at once.
void State::apply_prims_sharded(std::vector<Prim> pending) {
    std::vector<int> l2p = l2p_; // the plan works on a copy; the layout moves step by step below
    const std::vector<ShardStep> steps = plan_sharded(std::move(pending), l2p, shard_plan_config());
    const int c = pipeline_bits();
    bool any_exchange = false;
    for (const ShardStep &stp : steps)
        any_exchange = any_exchange || stp.is_exchange;
    if (c > 0 && any_exchange) {
        run_sharded_pipelined(steps, c);
    } else {
        for (const ShardStep &stp : steps) {
            if (stp.is_exchange)
                exchange_phys(stp.swaps);
            else
                run_local(stp.prims);
        }
    }
    B2_ASSERT(l2p_ == l2p);
}

// Slices only pay when a slice is still a long stream for the persistent tile kernel.
int State::pipeline_bits() const {
    if (!comm_ || !comm_uses_peer(comm_.get()) || !fuse_)
        return 0;
    int c = 2;
    if (const char *e = getenv("B2SV_PIPE_BITS"))
        c = std::max(0, std::min(4, atoi(e)));
    int min_sub = 24; // B2SV_PIPE_MIN_SUB: tests force slicing on small states
    if (const char *e = getenv("B2SV_PIPE_MIN_SUB"))
        min_sub = atoi(e);
    if (n_local_ - c < min_sub || n_local_ - c < B_ + gbits_ + 1 || n_eff_ != n_local_)
        return 0;
    return c;
}

// Pipelined execution of a sharded plan. Around every exchange a REGION of tasks is formed: the
// exchange(s) and the tile passes next to them that leave at least c local index bits untouched
// (not in their tile, not exchanged). Those c bits cut the shard into 2^c slices; inside the region
// every task runs slice by slice -- passes on the state's stream, exchanges on a second stream --
// in skewed order (slice q runs phase p at step q + p), so the NVLink transfer of one slice is in
// flight while the HBM passes of the neighbouring slices run. The slice bits are chosen per region,
// so the pass that needs a given qubit never has to wait for a whole-shard exchange next to it.
// Tasks outside regions are whole-shard launches on the state's stream.
void State::run_sharded_pipelined(const std::vector<ShardStep> &steps, int c) {
    touch();
    const int Q = 1 << c;
    const uint64_t rank_bits = uint64_t(rank_) << n_local_;
    if (!xstream_)
        CUDA_CHECK(cudaStreamCreateWithFlags(&xstream_, cudaStreamNonBlocking));
    // SMs the exchange kernel occupies while tile passes run beside it. The passes are the longer
    // stream at every world size measured (profiles/r2_trace*.json), and their speed follows the
    // number of SMs they keep, so the exchange gets just enough SMs to stay hidden behind them.
    int x_sms = 16;
    if (const char *e = getenv("B2SV_EXCHANGE_SMS"))
        x_sms = std::max(4, std::min(64, atoi(e)));
    const int x_sms_extra = 4; // exchanges of two or more bits move more data per pass beside them
    const int pass_ctas = sm_count_current_device() - x_sms - (gbits_ >= 2 ? x_sms_extra : 0);

    // ---- tasks: the tile passes of every run and the exchanges, in program order
    struct Task {
        const Pass *pass = nullptr;      // tile pass / generic-matrix pass
        const ShardStep *xchg = nullptr; // exchange
        uint64_t bits = 0;               // local index bits the task must see whole
        int region = -1;
    };
    std::vector<std::vector<Pass>> schedules;
    schedules.reserve(steps.size());
    std::vector<Task> tasks;
    const SchedConfig cfg = sched_config();
    const uint64_t all_local = bit(n_local_) - 1;
    for (const ShardStep &stp : steps) {
        Task t;
        if (stp.is_exchange) {
            t.xchg = &stp;
            for (const auto &pr : stp.swaps)
                t.bits |= bit(pr.second);
            tasks.push_back(t);
            continue;
        }
        schedules.push_back(build_schedule(stp.prims, cfg));
        for (const Pass &ps : schedules.back()) {
            t.pass = &ps;
            t.bits = all_local; // generic-matrix kernels are never sliced
            if (!ps.is_matk) {
                t.bits = 0;
                for (int j = 0; j < B_; j++)
                    t.bits |= bit(ps.hdr.tile_bits[j]);
            }
            tasks.push_back(t);
        }
    }
    // ---- regions
    struct Region {
        size_t begin, end;
        uint64_t slice_mask;
    };
    std::vector<Region> regions;
    {
        // slice bits above the contiguous low bits every tile holds: a slice is made of whole runs
        const int min_slice_pos = std::min(cfg.low, std::max(0, n_local_ - c - 1));
        const uint64_t cand0 = all_local & ~(bit(min_slice_pos) - 1);
        constexpr int kMaxSide = 5; // passes taken on either side of an exchange
        size_t prev_end = 0;
        for (size_t x = 0; x < tasks.size(); x++) {
            if (!tasks[x].xchg || tasks[x].region >= 0)
                continue;
            uint64_t mask = cand0 & ~tasks[x].bits;
            if (__builtin_popcountll(mask) < c)
                continue;
            size_t lo = x, hi = x + 1;
            int left_n = 0, right_n = 0;
            bool progress = true;
            while (progress) {
                progress = false;
                if (hi < tasks.size() && right_n < kMaxSide) {
                    const uint64_t m2 = mask & ~tasks[hi].bits;
                    if (__builtin_popcountll(m2) >= c) {
                        mask = m2;
                        right_n = tasks[hi].xchg ? 0 : right_n + 1;
                        hi++;
                        progress = true;
                    } else {
                        right_n = kMaxSide;
                    }
                }
                if (lo > prev_end && left_n < kMaxSide && !tasks[lo - 1].xchg) {
                    const uint64_t m2 = mask & ~tasks[lo - 1].bits;
                    if (__builtin_popcountll(m2) >= c) {
                        mask = m2;
                        left_n++;
                        lo--;
                        progress = true;
                    } else {
                        left_n = kMaxSide;
                    }
                }
            }
            // drop trailing passes that follow the last exchange by more than kMaxSide (none by
            // construction) and keep the c highest remaining bits as the slice bits
            uint64_t slice = 0;
            for (int b = n_local_ - 1, k = 0; b >= 0 && k < c; b--)
                if (mask & bit(b)) {
                    slice |= bit(b);
                    k++;
                }
            for (size_t t = lo; t < hi; t++)
                tasks[t].region = static_cast<int>(regions.size());
            regions.push_back({lo, hi, slice});
            prev_end = hi;
        }
    }

    // ---- issue machinery
    std::vector<cudaEvent_t> events;
    struct Guard {
        cudaEvent_t ev = nullptr;
        cudaStream_t st = nullptr;
    };
    std::vector<Guard> last(Q);
    auto new_event = [&](cudaStream_t st) {
        cudaEvent_t ev;
        CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(ev, st));
        events.push_back(ev);
        return ev;
    };
    auto wait_for = [&](cudaStream_t st, int q) {
        if (last[q].ev && last[q].st != st)
            CUDA_CHECK(cudaStreamWaitEvent(st, last[q].ev, 0));
    };
    auto params = std::make_unique<PassParams>();
    auto load_params = [&](const Pass &ps) {
        params->hdr = ps.hdr;
        std::memcpy(params->ops, ps.ops.data(), sizeof(DevOp) * ps.ops.size());
        if (!ps.dense.empty())
            std::memcpy(params->dense, ps.dense.data(), sizeof(DevDense) * ps.dense.size());
    };
    auto deposit = [&](int q, uint64_t slice_mask) { // slice number -> index offset
        uint64_t v = 0;
        int k = 0;
        for (int b = 0; b < n_local_; b++)
            if (slice_mask & bit(b)) {
                if ((q >> k) & 1)
                    v |= bit(b);
                k++;
            }
        return v;
    };
    auto to_jl = [&](const ShardStep &x) {
        std::vector<std::pair<int, int>> jl;
        for (const auto &pr : x.swaps)
            jl.emplace_back(pr.first - n_local_, pr.second);
        return jl;
    };
    auto advance_layout = [&](const ShardStep &x) {
        std::vector<int> p2l(n_);
        for (int q = 0; q < n_; q++)
            p2l[l2p_[q]] = q;
        for (const auto &pr : x.swaps)
            std::swap(l2p_[p2l[pr.first]], l2p_[p2l[pr.second]]);
    };

    size_t i = 0;
    while (i < tasks.size()) {
        if (tasks[i].region < 0) { // whole-shard launch on the state's stream
            if (tasks[i].pass)
                upload_and_run(std::vector<Pass>{*tasks[i].pass});
            else
                exchange_phys(tasks[i].xchg->swaps);
            i++;
            continue;
        }
        const Region &rg = regions[tasks[i].region];
        B2_ASSERT(rg.begin == i);
        // phases of passes separated by exchanges
        std::vector<std::vector<const Pass *>> phase(1);
        std::vector<const ShardStep *> xafter; // exchange that follows phase p (nullptr for the last)
        for (size_t t = rg.begin; t < rg.end; t++) {
            if (tasks[t].xchg) {
                xafter.push_back(tasks[t].xchg);
                phase.emplace_back();
            } else {
                phase.back().push_back(tasks[t].pass);
            }
        }
        xafter.push_back(nullptr);
        const int P = static_cast<int>(phase.size());
        { // everything queued so far precedes every slice of the region
            cudaEvent_t e0 = new_event(stream_);
            for (int q = 0; q < Q; q++)
                last[q] = {e0, stream_};
        }
        // Phase-major order: all slices of phase 0, then all slices of phase 1, ... The exchange of
        // slice q (second stream) starts as soon as the passes of phase p on q are done and runs
        // while the state's stream works through the remaining slices of that phase; by the time
        // phase p + 1 reaches slice q its exchange has had Q - 1 slice-times to finish. The exchange
        // stream never waits behind passes of a later phase.
        for (int p = 0; p < P; p++) {
            int k_bits = 0;
            if (xafter[p])
                k_bits = static_cast<int>(xafter[p]->swaps.size());
            for (int q = 0; q < Q; q++) {
                const uint64_t off = deposit(q, rg.slice_mask);
                for (const Pass *ps : phase[p]) {
                    load_params(*ps);
                    uint64_t tile_mask = 0;
                    for (int j = 0; j < B_; j++)
                        tile_mask |= bit(ps->hdr.tile_bits[j]);
                    fill_tile_id_segments(params->hdr, tile_mask | rg.slice_mask, n_local_);
                    wait_for(stream_, q);
                    {
                        TraceScope ts(*this, 0);
                        launch_tile_pass(dtype_, static_cast<char *>(d_state_) + off * amp_bytes(), *params,
                                         n_local_ - c, rank_bits | off, stream_, pass_ctas);
                    }
                    launches++;
                    last[q] = {new_event(stream_), stream_};
                }
                if (xafter[p]) {
                    wait_for(xstream_, q);
                    {
                        TraceScope ts(*this, 2, xstream_);
                        comm_exchange(comm_.get(), d_state_, peers_, dtype_, n_local_, to_jl(*xafter[p]),
                                      xstream_, 1, x_sms + (k_bits >= 2 ? x_sms_extra : 0), true,
                                      rg.slice_mask, off);
                    }
                    last[q] = {new_event(xstream_), xstream_};
                }
            }
        }
        for (int q = 0; q < Q; q++) // whatever follows on the state's stream sees the finished region
            wait_for(stream_, q);
        for (int p = 0; p < P; p++) {
            sweeps += phase[p].size();
            bytes_moved += 2 * state_bytes() * phase[p].size();
            if (xafter[p]) {
                comm_count_exchange(comm_.get());
                advance_layout(*xafter[p]);
            }
        }
        i = rg.end;
    }
    for (cudaEvent_t ev : events)
        cudaEventDestroy(ev);
}

// ---- gates ----------------------------------------------------------------------------------------
void State::lower(const GateOp &op, bool flip_inverse, std::vector<Prim> &out) const {
    const std::vector<int> bits = wires_to_bits(op.wires, n_);
    const bool inv = op.inverse ^ flip_inverse;
    if (op.name == "Identity")
        return;
    if (lower_gate(op.name, bits, inv, op.params, out))
        return;
    // not a named gate: the matrix path (reference StateVectorKokkos.hpp:592-598)
    B2_ABORT_IF(op.matrix.empty(),
                "operation '" + op.name + "' is not a named gate and no matrix was provided");
    lower_matrix(bits, inv, op.matrix, out);
}

void State::apply_gate(const GateOp &op) {
    std::vector<Prim> prims;
    lower(op, false, prims);
    apply_prims(std::move(prims));
}

void State::apply_ops(const std::vector<GateOp> &ops, bool adjoint) {
    std::vector<Prim> prims;
    auto flush = [&]() {
        if (!prims.empty())
            apply_prims(std::move(prims));
        prims.clear();
    };
    if (!adjoint) {
        for (const auto &op : ops) {
            lower(op, false, prims);
            if (!fuse_)
                flush();
        }
    } else {
        for (auto it = ops.rbegin(); it != ops.rend(); ++it) {
            lower(*it, true, prims);
            if (!fuse_)
                flush();
        }
    }
    flush();
}

void State::apply_ops_cached(const std::vector<GateOp> &ops, std::shared_ptr<PlanCache> &cache) {
    const char *off = getenv("B2SV_PLAN_CACHE");
    if (comm_ || !fuse_ || ops.empty() || (off && atoi(off) == 0)) {
        apply_ops(ops, false);
        return;
    }
    CUDA_CHECK(cudaSetDevice(device_));
    const SchedConfig cfg = sched_config();
    std::ostringstream ks;
    ks << n_ << ':' << dtype_ << ':' << cfg.B << ':' << cfg.R << ':' << cfg.low << ':' << cfg.max_heavy << ':'
       << cfg.factor << ':' << cfg.fuse_store << ':' << n_eff_ << ':' << cfg.bulk << ':' << cfg.bulk_min_run_bits;
    const std::string key = ks.str();
    if (!cache || cache->key != key) {
        std::vector<Prim> prims;
        for (const auto &op : ops)
            lower(op, false, prims);
        auto pc = std::make_shared<PlanCache>();
        pc->key = key;
        if (!prims.empty())
            pc->passes = build_schedule(prims, cfg);
        pc->graphable = true;
        for (const Pass &ps : pc->passes)
            pc->graphable = pc->graphable && !ps.is_matk;
        cache = std::move(pc);
    }
    if (cache->passes.empty())
        return;
    // small states are launch-bound: replay the passes as one CUDA graph
    int graph_max_bits = 24;
    if (const char *e = getenv("B2SV_GRAPH_MAX_BITS"))
        graph_max_bits = atoi(e);
    if (!cache->graphable || tracing_ || n_eff_ > graph_max_bits) {
        upload_and_run(cache->passes);
        return;
    }
    if (!cache->graph || cache->graph_state != d_state_ || cache->graph_stream != stream_) {
        if (cache->graph) {
            cudaGraphExecDestroy(cache->graph);
            cache->graph = nullptr;
        }
        const uint64_t s0 = sweeps, l0 = launches, b0 = bytes_moved;
        cudaGraph_t g = nullptr;
        CUDA_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
        try {
            upload_and_run(cache->passes);
        } catch (...) {
            cudaStreamEndCapture(stream_, &g);
            if (g)
                cudaGraphDestroy(g);
            throw;
        }
        CUDA_CHECK(cudaStreamEndCapture(stream_, &g));
        sweeps = s0, launches = l0, bytes_moved = b0; // capturing launched nothing
        cudaError_t e = cudaGraphInstantiate(&cache->graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { // no graph: plain launches
            cudaGetLastError();
            cache->graph = nullptr;
            cache->graphable = false;
            upload_and_run(cache->passes);
            return;
        }
        cache->graph_state = d_state_;
        cache->graph_stream = stream_;
    }
    touch();
    CUDA_CHECK(cudaGraphLaunch(cache->graph, stream_));
    sweeps += cache->passes.size();
    launches += 1;
    bytes_moved += 2 * state_bytes() * cache->passes.size();
}

void State::apply_ops_to_all(const std::vector<State *> &states, const std::vector<GateOp> &ops) {
    if (states.empty() || ops.empty())
        return;
    State &first = *states[0];
    bool same = first.fuse_ && !first.comm_;
    for (State *s : states)
        same = same && !s->comm_ && s->n_ == first.n_ && s->dtype_ == first.dtype_ &&
               s->device_ == first.device_ && s->fuse_;
    if (!same) {
        for (State *s : states)
            s->apply_ops(ops, false);
        return;
    }
    CUDA_CHECK(cudaSetDevice(first.device_));
    std::vector<Prim> prims;
    for (const auto &op : ops)
        first.lower(op, false, prims);
    if (prims.empty())
        return;
    const std::vector<Pass> passes = build_schedule(prims, first.sched_config());
    for (State *s : states)
        s->upload_and_run(passes);
}

double State::apply_generator(const std::string &name, const std::vector<int64_t> &wires) {
    std::vector<Prim> prims;
    double scale = 0.0;
    const bool ok = lower_generator(name, wires_to_bits(wires, n_), prims, &scale);
    B2_ABORT_IF(!ok, "Generator does not exist for " + name); // SV.hpp:692-693
    apply_prims(std::move(prims));
    return scale;
}

void State::apply_prims(std::vector<Prim> prims) {
    if (prims.empty())
        return;
    CUDA_CHECK(cudaSetDevice(device_));
    if (comm_)
        apply_prims_sharded(std::move(prims));
    else
        run_local(prims);
}

SchedConfig State::sched_config() const {
    SchedConfig cfg;
    cfg.B = B_;
    cfg.R = R_;
    cfg.low = default_tile_low();
    if (const char *e = getenv("B2SV_MAX_HEAVY"))
        cfg.max_heavy = std::max(1, atoi(e));
    cfg.SW = dtype_ == 1 ? 3 : 4;
    cfg.SH = dtype_ == 1 ? 0 : 1;
    cfg.f32 = dtype_ != 1;
    if (const char *e = getenv("B2SV_FACTOR"))
        cfg.factor = atoi(e) != 0;
    if (const char *e = getenv("B2SV_FUSE_STORE")) // 0: always go through the store phase
        cfg.fuse_store = atoi(e) != 0;
    cfg.n_local = n_local_;
    cfg.n_alloc = n_eff_;
    cfg.fuse = fuse_;
    // Bulk async tile loads (plain-layout passes, cp.async.bulk) are opt-in, B2SV_BULK=1: on one GPU
    // they are neutral for runs >= 4 KiB and 10 % slower for 512-byte runs (profiles/r2_tile_bulk_ab.md),
    // and a 4-GPU run with sliced passes beside exchanges failed after ~20 steps with them (cause not
    // found) -- the default keeps every pass on the swizzled cp.async loader.
    cfg.bulk = false;
    if (const char *e = getenv("B2SV_BULK"))
        cfg.bulk = atoi(e) != 0;
    if (const char *e = getenv("B2SV_BULK_MIN_RUN")) // experiments: 5 = every eligible pass (512-byte copies)
        cfg.bulk_min_run_bits = std::max(cfg.low, atoi(e));
    return cfg;
}

void State::run_local(const std::vector<Prim> &prims) {
    upload_and_run(build_schedule(prims, sched_config()));
}

void State::upload_and_run(const std::vector<Pass> &passes) {
    touch();
    // Pass descriptors travel as kernel parameters (copied by the runtime at launch), so there is
    // no staging buffer, no H2D copy and nothing to keep alive after the launch call returns.
    const uint64_t rank_bits = uint64_t(rank_) << n_local_;
    auto params = std::make_unique<PassParams>();
    last_upload_bytes_ = 0;
    for (size_t i = 0; i < passes.size(); i++) {
        const Pass &ps = passes[i];
        if (ps.is_matk) {
            const int k = static_cast<int>(ps.matk.bits.size());
            for (int b : ps.matk.bits)
                B2_ABORT_IF(b >= n_local_, "matrix operations on global (rank) qubits are not supported");
            double2 *d_mat;
            const size_t bytes = sizeof(double2) * ps.matk.mat.size();
            CUDA_CHECK(cudaMallocAsync(&d_mat, bytes, stream_));
            CUDA_CHECK(cudaMemcpyAsync(d_mat, ps.matk.mat.data(), bytes, cudaMemcpyHostToDevice, stream_));
            {
                TraceScope ts(*this, 1);
                launch_matk(dtype_, d_state_, n_eff_, d_mat, ps.matk.bits.data(), k, stream_);
            }
            CUDA_CHECK(cudaFreeAsync(d_mat, stream_));
            CUDA_CHECK(cudaStreamSynchronize(stream_)); // the host matrix lives in `passes`
            last_upload_bytes_ += bytes;
        } else {
            B2_ASSERT(ps.ops.size() <= static_cast<size_t>(kMaxOpsPerPass));
            params->hdr = ps.hdr;
            std::memcpy(params->ops, ps.ops.data(), sizeof(DevOp) * ps.ops.size());
            B2_ASSERT(ps.dense.size() <= static_cast<size_t>(kMaxDense));
            if (!ps.dense.empty())
                std::memcpy(params->dense, ps.dense.data(), sizeof(DevDense) * ps.dense.size());
            {
                TraceScope ts(*this, 0);
                launch_tile_pass(dtype_, d_state_, *params, n_eff_, rank_bits, stream_);
            }
            last_upload_bytes_ += sizeof(DevPassHeader) + sizeof(DevOp) * ps.ops.size() +
                                  sizeof(DevDense) * ps.dense.size();
        }
        sweeps++;
        launches++;
        bytes_moved += 2 * state_bytes();
    }
}

// ---- reductions -----------------------------------------------------------------------------------
void State::finish_reduce(int nv, double *out) const {
    launch_finalize(d_partials_, kReduceBlocks, nv, d_out_, stream_);
    reduce_launches += 2;
    if (comm_)
        comm_allreduce_sum(comm_.get(), d_out_, nv, stream_);
    CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, sizeof(double) * nv, cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    for (int i = 0; i < nv; i++)
        out[i] = h_out_[i];
}
double State::norm2() const {
    CUDA_CHECK(cudaSetDevice(device_));
    launch_norm2(dtype_, d_state_, local_length(), d_partials_, stream_);
    bytes_moved += state_bytes();
    double r;
    finish_reduce(1, &r);
    return r;
}
void State::inner_product_buf(const void *x, const void *y, double *re, double *im) const {
    CUDA_CHECK(cudaSetDevice(device_));
    launch_dot(dtype_, x, y, local_length(), d_partials_, stream_);
    bytes_moved += 2 * state_bytes();
    double r[2];
    finish_reduce(2, r);
    if (re)
        *re = r[0];
    if (im)
        *im = r[1];
}
void State::inner_product(const State &o, double *re, double *im) const {
    B2_ABORT_IF(o.n_ != n_ || o.dtype_ != dtype_ || o.device_ != device_,
                "state vectors are not compatible");
    if (!same_layout(o)) {
        normalize_layout();
        o.normalize_layout();
    }
    order_after(stream_, o.stream_);
    inner_product_buf(d_state_, o.d_state_, re, im);
}
double State::expval_pauli(uint64_t x, uint64_t z, cplx ph) const {
    CUDA_CHECK(cudaSetDevice(device_));
    ensure_local(x); // X / Y factors pair amplitudes: those qubits must be shard-local
    x = phys_mask(x);
    z = phys_mask(z);
    // Z factors on rank bits contribute a per-rank constant sign
    if (__builtin_popcountll((uint64_t(rank_) << n_local_) & z) & 1)
        ph = -ph;
    launch_pauli_expval(dtype_, d_state_, local_length(), x, z & (local_length() - 1), ph.real(),
                        ph.imag(), d_partials_, stream_);
    double r;
    finish_reduce(1, &r);
    return r;
}
double State::expval_named(const std::string &name, const std::vector<int64_t> &wires) const {
    CUDA_CHECK(cudaSetDevice(device_));
    if (name == "Identity")
        return norm2(); // EVF.hpp:13-28
    B2_ABORT_IF(wires.size() != 1, "named observables act on exactly one wire");
    const int tq = wires_to_bits(wires, n_)[0];
    if (name == "PauliZ" && n_local_ >= 12 && n_local_ <= 40 && n_eff_ == n_local_)
        return expval_z_all()[wires[0]];
    if (name != "PauliZ")
        ensure_local(bit(tq));
    const int t = phys_bit(tq);
    const double s = 0.70710678118654752440;
    double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (name == "PauliX") { // EVF.hpp:30-58
        m[2] = 1;
        m[4] = 1;
    } else if (name == "PauliY") { // EVF.hpp:60-92
        m[3] = -1;
        m[5] = 1;
    } else if (name == "PauliZ") { // EVF.hpp:94-121
        m[0] = 1;
        m[6] = -1;
    } else if (name == "Hadamard") { // EVF.hpp:123-153
        m[0] = s;
        m[2] = s;
        m[4] = s;
        m[6] = -s;
    } else {
        B2_ABORT("unknown named observable '" + name + "'");
    }
    if (t >= n_local_) // PauliZ on a rank bit: a per-rank sign
        return expval_pauli(0, bit(tq), 1.0);
    launch_expval_1q(dtype_, d_state_, local_length(), t, m, d_partials_, stream_);
    double r;
    finish_reduce(1, &r);
    return r;
}
const std::vector<double> &State::expval_z_all() const {
    if (zcache_version_ == version_ && !zcache_.empty())
        return zcache_;
    CUDA_CHECK(cudaSetDevice(device_));
    B2_ABORT_IF(n_local_ < 12 || n_local_ > 40 || n_eff_ != n_local_,
                "internal: the all-Z reduction needs between 12 and 40 local qubits");
    launch_expval_z_all(dtype_, d_state_, n_local_, d_partials_, stream_);
    bytes_moved += state_bytes();
    launch_finalize(d_partials_, kReduceBlocks, kZAllValsHost, d_out_, stream_);
    reduce_launches += 2;
    std::vector<double> loc(kZAllValsHost);
    if (!comm_) {
        CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, sizeof(double) * kZAllValsHost, cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        for (int i = 0; i < kZAllValsHost; i++)
            loc[i] = h_out_[i];
        zcache_.assign(n_, 0.0);
        for (int q = 0; q < n_; q++)
            zcache_[n_ - 1 - q] = loc[0] - 2.0 * loc[1 + q]; // wire w <-> index bit n - 1 - w
    } else {
        // per rank: [tot, P1(logical bit 0), ...]; a logical bit that sits on a rank bit contributes
        // tot or nothing, according to this rank's value of that bit
        CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, sizeof(double) * kZAllValsHost, cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        std::vector<double> mine(n_ + 1, 0.0);
        mine[0] = h_out_[0];
        for (int q = 0; q < n_; q++) {
            const int p = l2p_[q];
            mine[1 + q] = p < n_local_ ? h_out_[1 + p] : (((rank_ >> (p - n_local_)) & 1) ? h_out_[0] : 0.0);
        }
        double *d_tmp;
        CUDA_CHECK(cudaMallocAsync(&d_tmp, sizeof(double) * (n_ + 1), stream_));
        CUDA_CHECK(cudaMemcpyAsync(d_tmp, mine.data(), sizeof(double) * (n_ + 1), cudaMemcpyHostToDevice, stream_));
        comm_allreduce_sum(comm_.get(), d_tmp, n_ + 1, stream_);
        CUDA_CHECK(cudaMemcpyAsync(mine.data(), d_tmp, sizeof(double) * (n_ + 1), cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaFreeAsync(d_tmp, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        zcache_.assign(n_, 0.0);
        for (int q = 0; q < n_; q++)
            zcache_[n_ - 1 - q] = mine[0] - 2.0 * mine[1 + q];
    }
    zcache_version_ = version_;
    return zcache_;
}
double State::expval_pauli_sum(const std::vector<PauliTerm> &terms_in) const {
    CUDA_CHECK(cudaSetDevice(device_));
    if (terms_in.empty())
        return 0.0;
    uint64_t xall = 0;
    for (const PauliTerm &t : terms_in)
        xall |= t.x;
    ensure_local(xall); // X / Y factors pair amplitudes: those qubits must be shard-local
    std::vector<PauliTerm> terms;
    terms.reserve(terms_in.size());
    const uint64_t rank_bits = uint64_t(rank_) << n_local_;
    for (const PauliTerm &t : terms_in) {
        PauliTerm u = t;
        u.x = phys_mask(t.x);
        const uint64_t zp = phys_mask(t.z);
        u.z = zp & (local_length() - 1);
        if (__builtin_popcountll(rank_bits & zp) & 1) { // Z factors on rank bits: a per-rank sign
            u.cr = -u.cr;
            u.ci = -u.ci;
        }
        terms.push_back(u);
    }
    std::stable_sort(terms.begin(), terms.end(),
                     [](const PauliTerm &a, const PauliTerm &b) { return a.x < b.x; });
    PauliTerm *d_terms;
    CUDA_CHECK(cudaMallocAsync(&d_terms, sizeof(PauliTerm) * terms.size(), stream_));
    CUDA_CHECK(cudaMemcpyAsync(d_terms, terms.data(), sizeof(PauliTerm) * terms.size(), cudaMemcpyHostToDevice, stream_));
    launch_pauli_sum_expval(dtype_, d_state_, local_length(), d_terms, static_cast<int>(terms.size()),
                            d_partials_, stream_);
    CUDA_CHECK(cudaFreeAsync(d_terms, stream_));
    {
        size_t nx = 1;
        for (size_t i = 1; i < terms.size(); i++)
            nx += terms[i].x != terms[i - 1].x;
        bytes_moved += nx * state_bytes();
    }
    double r;
    finish_reduce(1, &r); // synchronises: `terms` may go out of scope
    return r;
}
bool State::pauli_sum_apply_sharded(const std::vector<PauliTerm> &terms_in) {
    if (!comm_ || !comm_uses_peer(comm_.get()) || world_ > 64 || n_eff_ != n_local_)
        return false;
    for (int r = 0; r < world_; r++)
        if (!peers_[r])
            return false;
    CUDA_CHECK(cudaSetDevice(device_));
    std::vector<PauliTerm> terms;
    terms.reserve(terms_in.size());
    const uint64_t local_mask = local_length() - 1;
    for (const PauliTerm &t : terms_in) {
        PauliTerm u = t;
        const uint64_t xp = phys_mask(t.x), zp = phys_mask(t.z);
        const int src = rank_ ^ static_cast<int>(xp >> n_local_); // the shard that holds index i ^ x
        u.x = xp & local_mask;
        u.z = zp & local_mask;
        u.src = static_cast<uint32_t>(src);
        // Z factors on rank bits see the SOURCE index: its rank bits are those of `src`
        if (__builtin_popcountll((uint64_t(src) << n_local_) & zp) & 1) {
            u.cr = -u.cr;
            u.ci = -u.ci;
        }
        terms.push_back(u);
    }
    PeerPtrs pp{};
    for (int r = 0; r < world_; r++)
        pp.p[r] = peers_[r];
    void *out = acquire_scratch();
    PauliTerm *d_terms;
    CUDA_CHECK(cudaMallocAsync(&d_terms, sizeof(PauliTerm) * terms.size(), stream_));
    CUDA_CHECK(cudaMemcpyAsync(d_terms, terms.data(), sizeof(PauliTerm) * terms.size(), cudaMemcpyHostToDevice, stream_));
    comm_barrier(comm_.get(), stream_, 0); // every rank's shard holds the state the terms are applied to
    launch_pauli_sum_apply_sharded(dtype_, pp, out, local_length(), d_terms, static_cast<int>(terms.size()), stream_);
    comm_barrier(comm_.get(), stream_, 0); // nobody still reads this shard: it may be overwritten now
    // the shard itself is what the peers have mapped, so the result is copied back (no buffer swap)
    CUDA_CHECK(cudaMemcpyAsync(d_state_, out, local_length() * amp_bytes(), cudaMemcpyDeviceToDevice, stream_));
    CUDA_CHECK(cudaFreeAsync(d_terms, stream_));
    touch();
    launches += 2;
    bytes_moved += (terms.size() + 3) * state_bytes();
    CUDA_CHECK(cudaStreamSynchronize(stream_)); // `terms` goes out of scope
    release_scratch(out);
    return true;
}
double State::expval_matrix(const std::vector<int64_t> &wires, const std::vector<cplx> &mat) const {
    CUDA_CHECK(cudaSetDevice(device_));
    std::vector<int> bits = wires_to_bits(wires, n_);
    const size_t k = bits.size();
    B2_ABORT_IF(k == 0, "matrix observable needs at least one wire");
    B2_ABORT_IF(mat.size() != (size_t(1) << (2 * k)), "matrix size does not match the number of wires");
    {
        uint64_t m = 0;
        for (int b : bits)
            m |= bit(b);
        ensure_local(m);
        for (int &b : bits)
            b = phys_bit(b);
    }
    double r;
    if (k == 1) { // MK.hpp:283-300
        double m[8];
        for (int i = 0; i < 4; i++) {
            m[2 * i] = mat[i].real();
            m[2 * i + 1] = mat[i].imag();
        }
        launch_expval_1q(dtype_, d_state_, local_length(), bits[0], m, d_partials_, stream_);
        finish_reduce(1, &r);
    } else if (k == 2) { // MK.hpp:302-320
        double2 *d_m = reinterpret_cast<double2 *>(d_out_ + 16); // 16 complex = 32 doubles
        CUDA_CHECK(cudaMemcpyAsync(d_m, mat.data(), sizeof(double2) * 16, cudaMemcpyHostToDevice, stream_));
        launch_expval_2q(dtype_, d_state_, local_length(), bits[0], bits[1], d_m, d_partials_, stream_);
        finish_reduce(1, &r);
    } else { // MK.hpp:340-345: copy, apply, Re<psi|O psi>
        void *tmp = acquire_scratch();
        CUDA_CHECK(cudaMemcpyAsync(tmp, d_state_, alloc_length() * amp_bytes(), cudaMemcpyDeviceToDevice, stream_));
        double2 *d_mat;
        CUDA_CHECK(cudaMallocAsync(&d_mat, sizeof(double2) * mat.size(), stream_));
        CUDA_CHECK(cudaMemcpyAsync(d_mat, mat.data(), sizeof(double2) * mat.size(), cudaMemcpyHostToDevice, stream_));
        launch_matk(dtype_, tmp, n_eff_, d_mat, bits.data(), static_cast<int>(k), stream_);
        CUDA_CHECK(cudaFreeAsync(d_mat, stream_));
        inner_product_buf(d_state_, tmp, &r, nullptr);
        release_scratch(tmp);
    }
    return r;
}
double State::expval_csr(const CsrDevice &m) const {
    CUDA_CHECK(cudaSetDevice(device_));
    if (world_ > 1) {
        // Sharded: the matrix is resident on every rank (device-resident CSR of the whole operator), rank r
        // streams the non-zeros of ITS rows, psi[col] is read in place from the shard that holds it
        // (NVLink through the IPC mappings), the partial sums are all-reduced.
        B2_ABORT_IF(!comm_uses_peer(comm_.get()) || world_ > 64 || n_eff_ != n_local_,
                    "CSR expectation values on sharded states need the peer-mapped shards");
        B2_ABORT_IF(m.nrows != (uint64_t(1) << n_), "CSR matrix dimension does not match the state vector");
        normalize_layout(); // rows and columns are plain indices: rank = top bits
        const uint64_t r0 = uint64_t(rank_) << n_local_, r1 = r0 + local_length();
        uint64_t jr[2];
        CUDA_CHECK(cudaMemcpyAsync(&jr[0], m.ptr + r0, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaMemcpyAsync(&jr[1], m.ptr + r1, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        PeerPtrs pp{};
        for (int r = 0; r < world_; r++) {
            B2_ABORT_IF(!peers_[r], "CSR expectation values on sharded states need the peer-mapped shards");
            pp.p[r] = peers_[r];
        }
        comm_barrier(comm_.get(), stream_, 0); // every shard holds its final amplitudes
        if (jr[1] > jr[0])
            launch_csr_expval_sharded(dtype_, d_state_, pp, n_local_, m.data, m.ind, m.ptr, r0, r1, jr[0], jr[1],
                                      d_partials_, stream_);
        else
            CUDA_CHECK(cudaMemsetAsync(d_partials_, 0, sizeof(double) * kReduceBlocks, stream_));
        comm_barrier(comm_.get(), stream_, 0); // nobody still reads this shard
        bytes_moved += (jr[1] - jr[0]) * 20 + state_bytes();
        double r;
        finish_reduce(1, &r); // all-reduces
        return r;
    }
    B2_ABORT_IF(m.nrows != local_length(), "CSR matrix dimension does not match the state vector");
    if (m.nnz > 0 && (getenv("B2SV_CSR_ROWS") == nullptr))
        launch_csr_expval_stream(dtype_, d_state_, m.data, m.ind, m.ptr, m.ptr32, m.nrows, m.nnz, d_partials_,
                                 stream_);
    else
        launch_csr_expval(dtype_, d_state_, m.data, m.ind, m.ptr, m.nrows, m.lanes, d_partials_, stream_);
    bytes_moved += m.nnz * 20 + state_bytes();
    double r;
    finish_reduce(1, &r);
    return r;
}
void State::pauli_dot_im_to(const State &bra, uint64_t x, uint64_t z, cplx ph, double factor,
                            double *d_dst) const {
    CUDA_CHECK(cudaSetDevice(device_));
    B2_ABORT_IF(bra.n_ != n_ || bra.dtype_ != dtype_ || bra.device_ != device_,
                "state vectors are not compatible");
    if (!same_layout(bra)) {
        normalize_layout();
        bra.normalize_layout();
    }
    ensure_local(x); // identical layouts -> identical swap decisions on both states
    bra.ensure_local(x);
    const uint64_t xp = phys_mask(x), zp = phys_mask(z);
    if (__builtin_popcountll((uint64_t(rank_) << n_local_) & zp) & 1)
        ph = -ph;
    order_after(stream_, bra.stream_);
    launch_pauli_dot(dtype_, bra.d_state_, d_state_, local_length(), xp, zp & (local_length() - 1),
                     ph.real(), ph.imag(), d_partials_, stream_);
    launch_finalize_scaled(d_partials_, kReduceBlocks, 2, 1, factor, d_dst, stream_);
    order_after(bra.stream_, stream_);
    reduce_launches += 2;
    bytes_moved += 2 * state_bytes();
}
void State::transition_1q_to(const State &bra, const int *bits, int nb, double *d_scratch,
                             double *d_dst) const {
    CUDA_CHECK(cudaSetDevice(device_));
    B2_ABORT_IF(bra.n_ != n_ || bra.dtype_ != dtype_ || bra.device_ != device_,
                "state vectors are not compatible");
    // sharded states: `bits` are PHYSICAL, shard-local positions in a layout both states share (the
    // caller made the wires local); every rank sums over its shard and the sums are all-reduced
    B2_ABORT_IF(world_ > 1 && !same_layout(bra), "internal: transition sums need identical layouts");
    for (int j = 0; j < nb; j++)
        B2_ABORT_IF(bits[j] < 0 || bits[j] >= n_local_, "internal: transition sums need shard-local bits");
    order_after(stream_, bra.stream_);
    if (n_eff_ >= 11 && n_eff_ == n_local_)
        launch_transition_tile(dtype_, bra.d_state_, d_state_, n_local_, bits, nb, d_scratch, stream_);
    else
        launch_transition_1q(dtype_, bra.d_state_, d_state_, local_length(), bits, nb, d_scratch, stream_);
    launch_finalize(d_scratch, kReduceBlocks, kTransitionVals, d_dst, stream_);
    if (comm_)
        comm_allreduce_sum(comm_.get(), d_dst, kTransitionVals, stream_);
    order_after(bra.stream_, stream_);
    reduce_launches += 2;
    bytes_moved += 2 * state_bytes();
}
void State::dot_im_to(const State &bra, double factor, double *d_dst) const {
    CUDA_CHECK(cudaSetDevice(device_));
    B2_ABORT_IF(bra.n_ != n_ || bra.dtype_ != dtype_ || bra.device_ != device_,
                "state vectors are not compatible");
    if (!same_layout(bra)) {
        normalize_layout();
        bra.normalize_layout();
    }
    order_after(stream_, bra.stream_);
    launch_dot(dtype_, bra.d_state_, d_state_, local_length(), d_partials_, stream_);
    launch_finalize_scaled(d_partials_, kReduceBlocks, 2, 1, factor, d_dst, stream_);
    order_after(bra.stream_, stream_);
    reduce_launches += 2;
    bytes_moved += 2 * state_bytes();
}
void State::allreduce_device(double *d_buf, int n) const {
    if (comm_)
        comm_allreduce_sum(comm_.get(), d_buf, n, stream_);
}
void State::axpy(cplx alpha, const State &x) {
    CUDA_CHECK(cudaSetDevice(device_));
    if (!same_layout(x)) {
        normalize_layout();
        x.normalize_layout();
    }
    order_after(stream_, x.stream_);
    launch_axpy(dtype_, alpha.real(), alpha.imag(), x.d_state_, d_state_, local_length(), stream_);
    touch();
    launches++;
    bytes_moved += 3 * state_bytes();
    order_after(x.stream_, stream_);
}

// ---- probabilities / sampling -----------------------------------------------------------------------
void State::probs(const std::vector<int64_t> &wires_in, double *out) const {
    CUDA_CHECK(cudaSetDevice(device_));
    if (comm_)
        normalize_layout(); // wire w <-> index bit n-1-w again, rank = top bits
    std::vector<int64_t> wires = wires_in;
    bool all_sorted = wires.empty();
    if (!wires.empty() && static_cast<int>(wires.size()) == n_) {
        all_sorted = true;
        for (int i = 0; i < n_; i++)
            all_sorted = all_sorted && wires[i] == i;
    }
    double *d_p;
    if (all_sorted && !comm_) { // MK.hpp:389-408
        const uint64_t len = local_length();
        CUDA_CHECK(cudaMallocAsync(&d_p, sizeof(double) * len, stream_));
        launch_probs_full(dtype_, d_state_, len, d_p, stream_);
        CUDA_CHECK(cudaMemcpyAsync(out, d_p, sizeof(double) * len, cudaMemcpyDeviceToHost, stream_));
    } else { // MK.hpp:418-517, result directly in the requested wire order
        if (wires.empty())
            for (int i = 0; i < n_; i++)
                wires.push_back(i);
        // Output digit j reports wire wires[argsort[argsort[j]]]: the reference's transposition
        // (MeasuresFunctors.hpp:162-191, MK.hpp:493-509) uses argsort where rank is meant, and
        // its own literals pin that (src/tests/Test_StateVectorKokkos_Measure.cpp:21-46). For sorted
        // wires -- all the Python layer allows (lightning_kokkos.py:488-495) -- it is the identity.
        std::vector<size_t> arg(wires.size());
        for (size_t i = 0; i < arg.size(); i++)
            arg[i] = i;
        std::stable_sort(arg.begin(), arg.end(), [&](size_t a, size_t b) { return wires[a] < wires[b]; });
        std::vector<int64_t> eff(wires.size());
        for (size_t j = 0; j < eff.size(); j++)
            eff[j] = wires[arg[arg[j]]];
        const std::vector<int> bits = wires_to_bits(eff, n_);
        const int m = static_cast<int>(bits.size());
        const uint64_t nb = uint64_t(1) << m;
        // Sharded: every rank bins its shard into the full 2^m histogram (its rank supplies the
        // values of the global wires), then the histograms are summed over the ranks.
        B2_ABORT_IF(comm_ && nb > (uint64_t(1) << 30),
                    "probs over more than 30 wires of a sharded state does not fit one all-reduce");
        CUDA_CHECK(cudaMallocAsync(&d_p, sizeof(double) * nb, stream_));
        CUDA_CHECK(cudaMemsetAsync(d_p, 0, sizeof(double) * nb, stream_));
        launch_probs_marginal(dtype_, d_state_, local_length(), bits.data(), m,
                              uint64_t(rank_) << n_local_, d_p, stream_);
        if (comm_)
            comm_allreduce_sum(comm_.get(), d_p, static_cast<int>(nb), stream_);
        CUDA_CHECK(cudaMemcpyAsync(out, d_p, sizeof(double) * nb, cudaMemcpyDeviceToHost, stream_));
    }
    CUDA_CHECK(cudaFreeAsync(d_p, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    reduce_launches += 1;
}

void State::generate_samples(size_t shots, uint64_t seed, uint64_t *out) const {
    CUDA_CHECK(cudaSetDevice(device_));
    if (shots == 0)
        return;
    if (comm_)
        normalize_layout();
    const uint64_t len = local_length();
    const uint64_t csz = uint64_t(1) << kSampleChunkBits;
    const uint64_t nchunks = (len + csz - 1) / csz;
    double *d_chunk;
    unsigned long long *d_s;
    CUDA_CHECK(cudaMallocAsync(&d_chunk, sizeof(double) * (nchunks + 1), stream_));
    CUDA_CHECK(cudaMallocAsync(&d_s, sizeof(unsigned long long) * shots * n_, stream_));
    launch_chunk_sums(dtype_, d_state_, len, d_chunk, stream_);
    launch_scan_chunks(d_chunk, nchunks, stream_);
    ShardCdf sh{};
    if (comm_) {
        // every rank learns every rank's probability mass (one small all-reduce), and builds the same
        // prefix sums, so the ranks' intervals tile (0, total] without gaps or overlaps
        B2_ABORT_IF(world_ > 64, "sharded sampling supports up to 64 ranks");
        B2_ABORT_IF(shots * static_cast<size_t>(n_) > (size_t(1) << 30),
                    "too many sample bits for one all-reduce");
        double mass[64] = {0};
        CUDA_CHECK(cudaMemsetAsync(d_out_, 0, sizeof(double) * 64, stream_));
        CUDA_CHECK(cudaMemcpyAsync(d_out_ + rank_, d_chunk + nchunks, sizeof(double),
                                   cudaMemcpyDeviceToDevice, stream_));
        comm_allreduce_sum(comm_.get(), d_out_, world_, stream_);
        CUDA_CHECK(cudaMemcpyAsync(mass, d_out_, sizeof(double) * world_, cudaMemcpyDeviceToHost, stream_));
        CUDA_CHECK(cudaStreamSynchronize(stream_));
        double run = 0.0;
        for (int r = 0; r < world_; r++) {
            if (r == rank_)
                sh.offset = run;
            run += mass[r];
            if (r == rank_)
                sh.upper = run;
        }
        sh.sharded = 1;
        sh.global_total = run;
        sh.index_or = uint64_t(rank_) << n_local_;
    }
    launch_sample(dtype_, d_state_, len, d_chunk, nchunks, n_, shots, seed, sh, d_s, stream_);
    if (comm_) {
        launch_bits_to_f64(d_s, shots * n_, stream_);
        comm_allreduce_sum(comm_.get(), reinterpret_cast<double *>(d_s), static_cast<int>(shots * n_), stream_);
        launch_f64_to_bits(d_s, shots * n_, stream_);
    }
    CUDA_CHECK(cudaMemcpyAsync(out, d_s, sizeof(unsigned long long) * shots * n_,
                               cudaMemcpyDeviceToHost, stream_));
    CUDA_CHECK(cudaFreeAsync(d_chunk, stream_));
    CUDA_CHECK(cudaFreeAsync(d_s, stream_));
    CUDA_CHECK(cudaStreamSynchronize(stream_));
    reduce_launches += 3;
}

} // namespace b2sv
