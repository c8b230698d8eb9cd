// b2sv: the device-resident state vector and its host-side driver.
// Counterpart of the reference's StateVectorKokkos<P> (reference simulator/StateVectorKokkos.hpp:109)
// and of the drivers in MeasuresKokkos<P> (reference simulator/MeasuresKokkos.hpp:14).
#pragma once
#include "ir.hpp"
#include "kernels.cuh"
#include "schedule.hpp"

#include <memory>
#include <string>
#include <vector>

namespace b2sv {

struct Comm; // NCCL communicator wrapper (comm.cpp); nullptr for single-GPU states

// One gate of an op list in the reference's wire format (reference AdjointDiffKokkos.hpp:40-56)
struct GateOp {
    std::string name;
    std::vector<int64_t> wires;
    bool inverse = false;
    std::vector<double> params;
    std::vector<cplx> matrix; // used when `name` is not a named gate
};

struct CsrDevice {
    int device = 0;
    double2 *data = nullptr;
    uint32_t *ind = nullptr;
    uint64_t *ptr = nullptr;
    uint32_t *ptr32 = nullptr; // 32-bit copy of the row pointers when nnz < 2^32
    uint64_t nrows = 0, nnz = 0;
    int lanes = 1;
    ~CsrDevice();
};
std::shared_ptr<CsrDevice> csr_upload(int device, const cplx *data, const uint64_t *indices,
                                      const uint64_t *indptr, size_t nnz, size_t nrows);

// What a State keeps about an op list it has applied before (owned by the op-list handle): the tile
// passes of the fused schedule and, for small states, a CUDA graph of their launches. Repeated
// circuits (optimisation loops re-applying one handle) skip lowering, scheduling and per-kernel
// launch overhead.
struct PlanCache {
    std::string key;           // state geometry + scheduler settings the plan was built for
    std::vector<Pass> passes;
    bool graphable = false;    // no generic-matrix pass (those synchronise the stream)
    cudaGraphExec_t graph = nullptr;
    void *graph_state = nullptr; // device buffer and stream the graph was captured for
    cudaStream_t graph_stream = nullptr;
    ~PlanCache() {
        if (graph)
            cudaGraphExecDestroy(graph);
    }
};

class State {
  public:
    State(int num_qubits, int dtype, int device, int rank = 0, int world = 1,
          const void *nccl_id = nullptr);
    // a second state on the communicator and stream of `like` (collective when sharded)
    explicit State(const State &like, int share_stream);
    ~State();
    State(const State &) = delete;
    State &operator=(const State &) = delete;

    // ---- geometry
    int num_qubits() const { return n_; }       // total wires
    int num_local() const { return n_local_; }  // index bits held on this GPU
    int dtype() const { return dtype_; }
    int device() const { return device_; }
    int rank() const { return rank_; }
    int world() const { return world_; }
    uint64_t local_length() const { return uint64_t(1) << n_local_; }
    uint64_t alloc_length() const { return uint64_t(1) << n_eff_; }
    size_t amp_bytes() const { return dtype_ == 1 ? 16 : 8; }
    void *data() const { return d_state_; }
    cudaStream_t stream() const { return stream_; }
    void sync() const;

    // ---- initialisation / copies
    void reset();
    void init_zeros();
    void set_basis_state(uint64_t index);
    void set_state_vector(const uint64_t *indices, const cplx *values, size_t n);
    // all zeros, then the 2^k amplitudes `values` on the given wires (the other wires in |0>)
    void set_state_on_wires(const std::vector<int64_t> &wires, const cplx *values);
    void h2d(const void *host, size_t length);
    void d2h(void *host, size_t length) const;
    // amplitudes at global flat indices, as complex128 (sampled read; sharded: all-reduced)
    void get_amplitudes(const uint64_t *indices, size_t n, cplx *out) const;
    void copy_from(const State &other);
    std::unique_ptr<State> clone() const;
    void swap_buffer(void *&other_buffer); // exchange the device buffer with a scratch buffer
    void *acquire_scratch() const;         // a buffer of alloc_length() amplitudes
    void release_scratch(void *p) const;

    // ---- gates
    void apply_gate(const GateOp &op);
    void apply_ops(const std::vector<GateOp> &ops, bool adjoint);
    // the same, forward direction, remembering the schedule (and a CUDA graph) in `cache`
    void apply_ops_cached(const std::vector<GateOp> &ops, std::shared_ptr<PlanCache> &cache);
    // the same op list applied to several states of identical shape (adjoint sweep: lambda and every
    // H_lambda): lowered and scheduled once, the passes launched on each state
    static void apply_ops_to_all(const std::vector<State *> &states, const std::vector<GateOp> &ops);
    double apply_generator(const std::string &name, const std::vector<int64_t> &wires);
    void apply_prims(std::vector<Prim> prims);
    void lower(const GateOp &op, bool flip_inverse, std::vector<Prim> &out) const;
    // ---- sharded layout: logical index bit q lives at physical bit l2p_[q]; physical bits
    // >= num_local() are the rank bits. Gates that target a rank bit pull it into the shard by a
    // global<->local qubit swap over NVLink and leave it there (lazy, never undone until needed).
    int phys_bit(int logical) const { return l2p_.empty() ? logical : l2p_[logical]; }
    uint64_t phys_mask(uint64_t logical_mask) const;
    void ensure_local(uint64_t logical_mask) const; // make these logical bits shard-local
    void normalize_layout() const;                  // back to the identity layout
    bool same_layout(const State &o) const { return l2p_ == o.l2p_; }
    void comm_stats(uint64_t *swaps, uint64_t *bytes, int *peer) const;
    void set_fusion(bool f) { fuse_ = f; }
    bool fusion() const { return fuse_; }

    // ---- reductions (all return after synchronising; sharded states all-reduce)
    double norm2() const;
    void inner_product(const State &other, double *re, double *im) const; // <this|other>
    void inner_product_buf(const void *x, const void *y, double *re, double *im) const;
    double expval_named(const std::string &name, const std::vector<int64_t> &wires) const;
    double expval_matrix(const std::vector<int64_t> &wires, const std::vector<cplx> &m) const;
    double expval_csr(const CsrDevice &m) const;
    double expval_pauli(uint64_t x, uint64_t z, cplx ph) const;
    // <Z_w> for every wire w from ONE read pass; cached until the state changes, so a run of
    // ExpectationValue("PauliZ", [w]) calls (lightning_kokkos.py:554-559) costs one kernel and one sync
    const std::vector<double> &expval_z_all() const;
    // Re <psi| sum_t c_t P_t |psi>, Pauli words as (x, z, coefficient * i^nY) in logical bits
    double expval_pauli_sum(const std::vector<PauliTerm> &terms) const;
    // sharded states: this <- sum_t c_t P_t this, the partner amplitudes of terms with X / Y factors on
    // rank bits read from the peers' shards in place (terms in logical bits); false if the peers'
    // shards are not mapped (NCCL-only path), nothing done then
    bool pauli_sum_apply_sharded(const std::vector<PauliTerm> &terms);
    void touch() { version_++; } // the amplitudes changed: cached measurements are stale
    void axpy(cplx alpha, const State &x);
    // Adjoint-sweep reductions that stay on the device (no host synchronisation):
    //   *d_dst = factor * Im <bra| P |this>   with P|j> = ph (-1)^popc(j&z) |j^x>  (logical bits)
    //   *d_dst = factor * Im <bra|this>
    // Sharded states write the per-rank partial; the caller all-reduces the whole Jacobian once.
    void pauli_dot_im_to(const State &bra, uint64_t x, uint64_t z, cplx ph, double factor,
                         double *d_dst) const;
    void dot_im_to(const State &bra, double factor, double *d_dst) const;
    // Single-qubit transition sums <bra| . |this> for nb <= kTransitionBits index bits in one read
    // pass (kernels.cu k_transition_1q); d_scratch: kReduceBlocks x kTransitionVals doubles, result
    // (kTransitionVals doubles) lands in d_dst. Sharded states: `bits` are physical, shard-local positions in
    // a layout both states share; the sums are all-reduced over the ranks.
    void transition_1q_to(const State &bra, const int *bits, int nb, double *d_scratch,
                          double *d_dst) const;
    bool sharded() const { return world_ > 1; }
    void allreduce_device(double *d_buf, int n) const;
    // clone that lives on this state's stream (so kernels touching both need no cross-stream events)
    std::unique_ptr<State> clone_on_stream() const;

    // ---- probabilities / sampling
    void probs(const std::vector<int64_t> &wires, double *out) const;
    void generate_samples(size_t shots, uint64_t seed, uint64_t *out) const;

    // ---- tracing: CUDA events around every pass / exchange launched on this state's stream
    struct TraceOut {
        int kind;        // 0 tile pass, 1 generic matrix kernel, 2 global<->local exchange
        double start_ms; // relative to trace_begin
        double dur_ms;
    };
    void trace_begin();
    std::vector<TraceOut> trace_end();

    // ---- bookkeeping
    uint64_t sweeps = 0, launches = 0;
    mutable uint64_t reduce_launches = 0;
    // algorithmic bytes the kernels launched on this state have moved (passes 2S, read passes S per
    // vector, copies 2S, ...) and, on the state an adjoint Jacobian was taken of, the bytes that call
    // moved over all its work vectors
    mutable uint64_t bytes_moved = 0, last_adjoint_bytes = 0;
    uint64_t state_bytes() const { return local_length() * amp_bytes(); }

    int tile_bits() const { return B_; }
    size_t last_upload_bytes() const { return last_upload_bytes_; }

  private:
    void finish_reduce(int nv, double *out) const;
    void upload_and_run(const std::vector<Pass> &passes);
    SchedConfig sched_config() const;
    void run_local(const std::vector<Prim> &prims);       // prims in PHYSICAL bits, all targets local
    void apply_prims_sharded(std::vector<Prim> prims);    // prims in logical bits
    // pipelined execution of a sharded plan: the shard is cut into 2^c slices by its top c local
    // bits; passes whose tile holds none of those bits and the exchanges run slice by slice, passes
    // on the state's stream and exchanges on a second one, so that the NVLink transfer of one slice
    // overlaps the HBM passes over the others
    void run_sharded_pipelined(const std::vector<struct ShardStep> &steps, int c);
    int pipeline_bits() const;
    // rank-bit positions <-> local positions, all pairs in ONE exchange
    void exchange_phys(const std::vector<std::pair<int, int>> &pairs) const;
    int exchange_ctas() const;
    struct ShardPlanConfig shard_plan_config() const;
    void reset_layout() const;
    void init_common(const void *nccl_id);

    struct TraceRec {
        int kind;
        cudaEvent_t e0, e1;
    };
    struct TraceScope { // records a pair of events around the launches made while it is alive
        const State &s;
        cudaEvent_t e0 = nullptr;
        int kind;
        cudaStream_t st;
        TraceScope(const State &state, int k, cudaStream_t stream = nullptr);
        ~TraceScope();
    };
    uint64_t version_ = 1;
    mutable uint64_t zcache_version_ = 0;
    mutable std::vector<double> zcache_;
    mutable bool tracing_ = false;
    mutable cudaEvent_t trace_t0_ = nullptr;
    mutable std::vector<TraceRec> trace_;

    int n_, n_local_, n_eff_, dtype_, device_;
    int rank_, world_, gbits_;
    int B_, R_;
    bool fuse_ = true;
    void *d_state_ = nullptr;
    cudaStream_t stream_ = nullptr;
    mutable cudaStream_t xstream_ = nullptr; // exchanges of pipelined sharded runs
    size_t last_upload_bytes_ = 0; // descriptor bytes handed to the device by the last apply
    // reductions
    double *d_partials_ = nullptr, *d_out_ = nullptr, *h_out_ = nullptr;
    mutable std::vector<void *> scratch_;
    std::shared_ptr<Comm> comm_;
    bool owns_stream_ = true;
    mutable std::vector<int> l2p_;     // empty for single-GPU states
    mutable std::vector<void *> peers_; // IPC-mapped shards of all ranks (this rank: own buffer)
};

} // namespace b2sv
