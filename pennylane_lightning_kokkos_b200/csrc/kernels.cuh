// b2sv: host-callable launchers of the stand-alone CUDA kernels (kernels.cu, tile_kernel.cu).
// dtype: 0 = complex64, 1 = complex128. All launches are asynchronous on `stream`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace b2sv {

constexpr int kReduceBlocks = 1184; // 148 SMs x 8 resident CTAs of 256 threads
constexpr int kReduceThreads = 256;
constexpr int kMaxReduceVals = 48; // the all-Z reduction carries 41 values per block

// ---- tile executor (tile_kernel.cu)
void tile_config(int dtype, int *B, int *R);
struct PassParams;
// state / n_eff may describe a sub-range of a shard (its top index bits fixed, supplied through
// rank_bits); max_ctas > 0 caps the persistent grid (SMs left to a concurrent exchange kernel)
void launch_tile_pass(int dtype, void *state, const PassParams &pass, int n_eff,
                      uint64_t rank_bits, cudaStream_t stream, int max_ctas = 0);
void tile_prof_read(unsigned long long out[16]);

// ---- state management
void launch_set_basis(int dtype, void *state, uint64_t len, uint64_t index, cudaStream_t st);
void launch_scatter(int dtype, void *state, const uint64_t *d_idx, const double2 *d_val, size_t n,
                    cudaStream_t st);
// state[sum_j bit_j(v) << h_pos[j]] = d_val[v] for v < 2^k (h_pos[0] <-> MSB of v), this shard's part
void launch_scatter_wires(int dtype, void *state, const double2 *d_val, const int *h_pos, int k,
                          int n_local, uint64_t rank, cudaStream_t st);
// sampled read: d_out[k] = state[d_idx[k]] as complex128 (zero where the index is another rank's)
void launch_gather(int dtype, const void *state, const uint64_t *d_idx, size_t n, int n_local,
                   uint64_t rank, double2 *d_out, cudaStream_t st);
void launch_axpy(int dtype, double ar, double ai, const void *x, void *y, uint64_t len,
                 cudaStream_t st);
// generic k-qubit matrix (row-major, device, complex128), bits[0] = MSB of the local index
void launch_matk(int dtype, void *state, int n_eff, const double2 *d_mat, const int *h_bits, int k,
                 cudaStream_t st);

// ---- multi-GPU: in-place exchange of rank bit j with local bit l through NVLink peer memory.
// `mine` / `peer` are the two shards (peer = IPC-mapped pointer of rank ^ (1 << j)); this rank
// swaps the pairs (mine[i | my_half], peer[i | peer_half]) whose selector bit equals `which`, the
// partner does the other half, so each NVLink direction carries S/4 reads + S/4 writes.
void launch_peer_swap(int dtype, void *mine, void *peer, int n_local, int lbit, int my_bit,
                      cudaStream_t st);

// k-bit exchange (all-to-all inside a group of 2^k ranks), see kernels.cu k_exchange
constexpr int kMaxExchangeBits = 4;
struct ExchangeParams {
    int k;                          // bits exchanged (1..kMaxExchangeBits)
    int n_local;                    // local index bits
    int nfix;                       // k + 1 + number of slice bits
    int nfree;                      // local index bits that are enumerated: n_local - nfix
    int fixpos[kMaxExchangeBits + 1 + 4]; // the k local bits, the selector bit and the bits that select
                                    // the slice of the shard this launch works on, ascending
    int lpos[kMaxExchangeBits];     // local bit paired with group-value bit i
    int selbit;                     // which half of a pair of sub-blocks this rank moves
    uint32_t a;                     // this rank's value on the k rank bits
    void *peer[1 << kMaxExchangeBits]; // shard of the rank with group value b (entry a unused)
};
// fat = true: CTAs of 1024 threads, one per SM, max_ctas of them (an exchange that runs beside tile
// passes occupies exactly max_ctas SMs); false: 256-thread CTAs, up to max_ctas
void launch_exchange(int dtype, void *mine, const ExchangeParams &p, int max_ctas, bool fat,
                     cudaStream_t st);
struct FlagPeers {
    unsigned long long *p[64];
};
void launch_flag_barrier(unsigned long long *mine, const FlagPeers &peers, int rank, int world,
                         unsigned long long epoch, cudaStream_t st);

// ---- reductions: every kernel writes kReduceBlocks x nv partials; finalize sums them (fixed order)
void launch_norm2(int dtype, const void *state, uint64_t len, double *d_partials, cudaStream_t st);
void launch_dot(int dtype, const void *x, const void *y, uint64_t len, double *d_partials,
                cudaStream_t st); // 2 values: Re<x|y>, Im<x|y>
void launch_expval_1q(int dtype, const void *state, uint64_t len, int tbit, const double *m8,
                      double *d_partials, cudaStream_t st);
void launch_expval_2q(int dtype, const void *state, uint64_t len, int bit_a, int bit_b,
                      const double2 *d_m16, double *d_partials, cudaStream_t st);
// <psi| P |psi> for a Pauli word: P|j> = ph * (-1)^popc(j & z) |j ^ x>, ph = i^nY
void launch_pauli_expval(int dtype, const void *state, uint64_t len, uint64_t x, uint64_t z,
                         double phr, double phi, double *d_partials, cudaStream_t st);
void launch_finalize(const double *d_partials, int nblocks, int nv, double *d_out,
                     cudaStream_t st);
// every single-qubit <Z> in one read pass (needs 12 <= n_bits <= 40): kReduceBlocks x kZAllValsHost
// partials; after launch_finalize(nv = kZAllValsHost): [0] = sum |a|^2, [1 + b] = sum over the
// indices with bit b set
constexpr int kZAllValsHost = 41;
void launch_expval_z_all(int dtype, const void *state, int n_bits, double *d_partials, cudaStream_t st);
// 2 values: Re, Im of <bra| P |ket>, P|j> = ph * (-1)^popc(j & z) |j ^ x>
void launch_pauli_dot(int dtype, const void *bra, const void *ket, uint64_t len, uint64_t x,
                      uint64_t z, double phr, double phi, double *d_partials, cudaStream_t st);
// *d_dst = scale * sum_b partials[b * nv + which]   (device-resident Jacobian entries)
void launch_finalize_scaled(const double *d_partials, int nblocks, int nv, int which, double scale,
                            double *d_dst, cudaStream_t st);

struct PauliTerm {
    uint64_t x, z;
    double cr, ci; // coefficient * i^nY
    uint32_t src;  // sharded application: which rank's shard holds the partner amplitudes of this term
    uint32_t pad_;
};
struct PeerPtrs {
    const void *p[64];
};
// sharded form of launch_pauli_sum_apply: term t reads its partner amplitudes from shard peers.p[src]
// (this rank's own buffer or a peer's IPC-mapped shard over NVLink); x, z are shard-local masks
void launch_pauli_sum_apply_sharded(int dtype, const PeerPtrs &peers, void *out, uint64_t len,
                                    const PauliTerm *d_terms, int nterms, cudaStream_t st);
// out = sum_t coef_t P_t in   (in != out)
// 1 value: Re <psi| sum_t coef_t P_t |psi>; terms sorted by x
void launch_pauli_sum_expval(int dtype, const void *state, uint64_t len, const PauliTerm *d_terms,
                             int nterms, double *d_partials, cudaStream_t st);
void launch_pauli_sum_apply(int dtype, const void *in, void *out, uint64_t len,
                            const PauliTerm *d_terms, int nterms, cudaStream_t st);

// ---- CSR (device-resident, 32- or 64-bit indices chosen by the caller: here always 64-bit ptr,
// 32-bit column indices when nrows < 2^32)
void launch_csr_expval(int dtype, const void *state, const double2 *d_data, const uint32_t *d_ind,
                       const uint64_t *d_ptr, uint64_t nrows, int lanes_per_row,
                       double *d_partials, cudaStream_t st);
// expectation value as a stream over the non-zeros (no per-row reduction); d_ptr32 (32-bit row
// pointers, for nnz < 2^32) is used when not null, else d_ptr64
void launch_csr_expval_stream(int dtype, const void *state, const double2 *d_data, const uint32_t *d_ind,
                              const uint64_t *d_ptr64, const uint32_t *d_ptr32, uint64_t nrows, uint64_t nnz,
                              double *d_partials, cudaStream_t st);
// sharded state: rank-local rows [row_begin, row_end), their non-zeros [j_begin, j_end); psi[col] is
// read from the shard that holds it (peers.p[col >> n_local], IPC-mapped)
void launch_csr_expval_sharded(int dtype, const void *state, const PeerPtrs &peers, int n_local,
                               const double2 *d_data, const uint32_t *d_ind, const uint64_t *d_ptr64,
                               uint64_t row_begin, uint64_t row_end, uint64_t j_begin, uint64_t j_end,
                               double *d_partials, cudaStream_t st);
void launch_csr_spmv(int dtype, const void *x, void *y, const double2 *d_data,
                     const uint32_t *d_ind, const uint64_t *d_ptr, uint64_t nrows,
                     int lanes_per_row, cudaStream_t st);

// Single-qubit transition sums <bra| . |ket> for nb <= kTransitionBits index bits in one read pass.
// d_partials: kReduceBlocks x kTransitionVals doubles; finalize with launch_finalize(..., nv =
// kTransitionVals). Layout of the result: D (re, im), then per bit Z_t, X_t, W_t (re, im each), see
// kernels.cu k_transition_1q.
constexpr int kTransitionBits = 6;
constexpr int kTransitionVals = 2 + 6 * kTransitionBits;
void launch_transition_1q(int dtype, const void *bra, const void *ket, uint64_t len,
                          const int *h_bits, int nb, double *d_partials, cudaStream_t st);

// Tiled form (needs n_bits >= 11): every requested bit is served from one read of the two vectors.
void launch_transition_tile(int dtype, const void *bra, const void *ket, int n_bits,
                            const int *h_bits, int nb, double *d_partials, cudaStream_t st);

// ---- probabilities and sampling
void launch_probs_full(int dtype, const void *state, uint64_t len, double *d_out, cudaStream_t st);
// d_out must be zeroed; bitpos[j] = index bit of requested wire j (wire order = output bit order,
// first wire = MSB of the output index)
// index_or: ORed into the local index before the bits are read (rank << n_local on sharded states)
void launch_probs_marginal(int dtype, const void *state, uint64_t len, const int *h_bitpos, int m,
                           uint64_t index_or, double *d_out, cudaStream_t st);
constexpr int kSampleChunkBits = 12;
void launch_chunk_sums(int dtype, const void *state, uint64_t len, double *d_chunk,
                       cudaStream_t st);
void launch_scan_chunks(double *d_chunk, uint64_t nchunks, cudaStream_t st); // exclusive, + total
// where this shard sits in the global cumulative distribution (all zero / false on one GPU)
struct ShardCdf {
    int sharded;          // 0: the state is not sharded, the fields below are ignored
    double offset;        // probability mass held by the lower ranks
    double upper;         // offset + this rank's mass, computed the same way on every rank
    double global_total;  // sum over all ranks
    uint64_t index_or;    // rank << n_local
};
void launch_sample(int dtype, const void *state, uint64_t len, const double *d_chunk_cdf,
                   uint64_t nchunks, int num_qubits, size_t shots, uint64_t seed,
                   const ShardCdf &sh, unsigned long long *d_out, cudaStream_t st);
void launch_bits_to_f64(unsigned long long *d, size_t n, cudaStream_t st); // in place
void launch_f64_to_bits(unsigned long long *d, size_t n, cudaStream_t st); // in place

} // namespace b2sv
