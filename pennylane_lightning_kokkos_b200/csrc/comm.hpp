// b2sv: multi-GPU plumbing -- one process per GPU, rank = top log2(world) index bits.
// NCCL is loaded lazily with dlopen (comm.cpp) so the single-GPU path has no NCCL dependency.
#pragma once
#include "ir.hpp"

#include <cuda_runtime.h>

namespace b2sv {

class State;
struct Comm;

void comm_unique_id(void *out128);
Comm *comm_create(int rank, int world, const void *nccl_unique_id, int device);
void comm_destroy(Comm *c);
// sum-all-reduce `n` doubles in place on the device
void comm_allreduce_sum(Comm *c, double *d_buf, int n, cudaStream_t stream);
// Rewrites `prims` so that no C1Q / MATK target sits on a rank bit, performing the required
// global<->local qubit swaps on `state` (pairwise half-shard exchanges over NVLink).
void comm_localize(Comm *c, State &state, std::vector<Prim> &prims);

} // namespace b2sv
