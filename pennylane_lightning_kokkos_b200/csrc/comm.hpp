// b2sv: multi-GPU plumbing -- one process per GPU, rank = top log2(world) index bits.
// NCCL is loaded lazily with dlopen (comm.cpp) so the single-GPU path has no NCCL dependency.
#pragma once
#include "ir.hpp"

#include <cuda_runtime.h>
#include <utility>
#include <vector>

namespace b2sv {

struct Comm;

void comm_unique_id(void *out128);
Comm *comm_create(int rank, int world, const void *nccl_unique_id, int device);
void comm_destroy(Comm *c);
int comm_rank(const Comm *c);
int comm_world(const Comm *c);
// sum-all-reduce `n` doubles in place on the device
void comm_allreduce_sum(Comm *c, double *d_buf, int n, cudaStream_t stream);
// stream-ordered barrier across ranks: IPC flag words (channel = independent set of flags, one per
// stream that may hold a barrier in flight) or, without peer mapping, a one-word NCCL all-reduce
void comm_barrier(Comm *c, cudaStream_t stream, int channel = 0);
void comm_setup_flags(Comm *c, cudaStream_t stream); // collective, idempotent
// CUDA-IPC mapping of every rank's shard (collective); falls back to the NCCL path on all ranks
// together if any mapping fails
void comm_map_peers(Comm *c, void *my_buffer, std::vector<void *> &out, cudaStream_t stream);
void comm_unmap_peers(Comm *c, std::vector<void *> &ptrs);
bool comm_uses_peer(const Comm *c);
// exchange rank bit j (0 = lowest rank bit) with local index bit l, in place
void comm_swap_bits(Comm *c, void *data, const std::vector<void *> &peers, int dtype, int n_local,
                    int j, int l, cudaStream_t stream);
// exchange k rank bits with k local bits in one all-to-all inside the 2^k-rank group, in place:
// jl[i] = (rank bit j_i, local bit l_i). max_ctas bounds the SMs the exchange kernel may occupy.
void comm_exchange(Comm *c, void *data, const std::vector<void *> &peers, int dtype, int n_local,
                   const std::vector<std::pair<int, int>> &jl, cudaStream_t stream, int channel,
                   int max_ctas, bool fat = false, uint64_t slice_mask = 0, uint64_t slice_value = 0);
// (slice_mask / slice_value: work only on the amplitudes whose local index has these bits at these
//  values -- the same slice of every shard; the caller walks the slices)
void comm_stats(const Comm *c, uint64_t *swaps, uint64_t *bytes);
void comm_reset_stats(Comm *c);
void comm_count_exchange(Comm *c);

} // namespace b2sv
