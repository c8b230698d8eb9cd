// b2sv: planner for sharded states -- splits a primitive-op list (logical index bits) into runs that
// execute inside the shards and global<->local exchanges between them.
//
// The state is sharded by its top g physical index bits (the rank). A primitive can run when its
// non-diagonal targets sit on shard-local physical bits; controls and phases on rank bits are
// CTA-uniform predicates and never move data. When the frontier needs a qubit that currently sits
// on a rank bit, ALL global qubits with a pending non-diagonal use are brought into the shard by
// ONE exchange -- an all-to-all inside the 2^k-rank group that moves (1 - 2^-k) of the shard,
// instead of k pairwise swaps of half a shard each -- and the qubits sent out are the local ones
// whose next non-diagonal use lies farthest ahead (Belady).
// Host-only code: no CUDA, usable by the CPU tests (b2sv_plan_sharded).
#pragma once
#include "ir.hpp"

#include <utility>

namespace b2sv {

struct ShardStep {
    bool is_exchange = false;
    std::vector<Prim> prims;                // run: primitives in PHYSICAL bits of the current layout
    std::vector<std::pair<int, int>> swaps; // exchange: (rank-bit physical position, local position)
};

struct ShardPlanConfig {
    int n = 0;          // logical = physical index bits in total
    int n_local = 0;    // physical bits below n_local are shard-local
    int min_victim_pos = 5; // local positions below this are not sent out (short HBM / NVLink runs)
    int max_victim_pos = 64; // local positions from here up are never sent out (the bits that slice a
                             // shard for pipelined execution must stay put)
    bool batch = true;  // false: one bit per exchange, as many exchanges as needed (A/B measurements)
};

// l2p: logical bit -> physical bit, updated in place to the layout after the last step.
std::vector<ShardStep> plan_sharded(std::vector<Prim> pending, std::vector<int> &l2p,
                                    const ShardPlanConfig &cfg);
// Exchanges that bring the layout back to the identity (logical q at physical q for the rank bits;
// the local positions are then sorted by free SWAPs inside one tile pass, see State::normalize_layout).
std::vector<ShardStep> plan_normalize(std::vector<int> &l2p, const ShardPlanConfig &cfg);

Prim prim_to_physical(const Prim &p, const std::vector<int> &l2p);

} // namespace b2sv
