// b2sv: fusion scheduler -- partitions a primitive-op list into HBM passes ("sweeps") and, inside a
// pass, into register rounds for the tile executor (tile_kernel.cu).
//
// A pass owns a tile of B index bits = the `low` contiguous low bits (coalescing) + (B - low) freely
// chosen bits. Every C1Q in the pass has its target among the tile bits; controls and DIAG masks may
// touch any bit (bits outside the tile are CTA-uniform predicates). Inside a pass each thread keeps
// 2^R amplitudes in registers; a round fixes which R tile bits are register-resident.
#pragma once
#include "ir.hpp"

namespace b2sv {

constexpr int kMaxRounds = 24;
constexpr int kMaxOpsPerPass = 80;
constexpr int kMaxTileBits = 16;
constexpr int kMaxRegBits = 5;

enum OpKind : uint8_t { KIND_GENERAL = 0, KIND_REAL = 1, KIND_PERM = 2, KIND_DIAG = 3 };

// Device-side op record (read from shared memory by every thread; 16-byte aligned).
struct alignas(16) DevOp {
    double m[8];        // C1Q: m00,m01,m10,m11 as (re,im); DIAG: p0,p1 as (re,im)
    uint64_t gcm, gcv;  // control over bits outside the tile (CTA-uniform, tested on tile_base)
    uint64_t gpm;       // DIAG parity over bits outside the tile
    uint32_t lcm, lcv;  // control over tile-local, non-register bits (per thread)
    uint32_t lpm;       // DIAG parity over tile-local, non-register bits
    uint32_t slot_act;  // bit s: register slot s satisfies the register part of the control
    uint32_t slot_par;  // DIAG: parity contribution of register slot s
    uint8_t kind;       // OpKind
    uint8_t tslot;      // C1Q: which register bit (0..R-1) is the target
    int16_t jac;        // adjoint: Jacobian accumulator slot fed by this op, -1 = none
};
static_assert(sizeof(DevOp) == 112, "DevOp layout");

struct alignas(16) DevPassHeader {
    int32_t n_ops;
    int32_t n_rounds;
    int32_t low_bits;                 // number of contiguous low bits in the tile
    int32_t reserved;
    uint8_t tile_bits[kMaxTileBits];  // ascending bit positions; tile_bits[j]=j for j<low_bits
    uint8_t round_regbits[kMaxRounds][8]; // tile-local positions held in registers, ascending
    uint16_t round_begin[kMaxRounds + 1]; // op index ranges per round
    uint16_t pad_[3];
};

struct Pass {
    bool is_matk = false;
    Prim matk;              // when is_matk
    DevPassHeader hdr{};    // otherwise
    std::vector<DevOp> ops;
    std::vector<int> tags;  // per op: Prim::tag
};

struct SchedConfig {
    int B = 12;        // tile bits
    int R = 4;         // register bits
    int low = 5;       // contiguous low bits forced into every tile
    int n_local = 0;   // bits >= n_local cannot be targets (rank bits when sharded)
    int n_alloc = 0;   // index bits of the allocation (>= B; small states are zero-padded)
    bool fuse = true;  // false: one pass per primitive group (reference schedule)
};

// Pre-pass: merge runs of uncontrolled single-bit primitives on the same bit into one 2x2.
std::vector<Prim> fuse_single_qubit(const std::vector<Prim> &prims);
// Main entry.
std::vector<Pass> build_schedule(const std::vector<Prim> &prims, const SchedConfig &cfg);

} // namespace b2sv
